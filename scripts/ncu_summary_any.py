"""Selected metrics of every launch of an `ncu --set full` capture -> profiles/<tag>.json (runs here, no GPU needed).

    python scripts/ncu_summary_any.py gpurun_out/<capture>.ncu-rep profiles/<tag>.json "<note>"
"""
import csv
import io
import json
import subprocess
import sys

KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum',
        'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__cycles_elapsed.avg']


def main():
  rep, out_path, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else '')
  out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
  rows = list(csv.reader(io.StringIO(out)))
  header, units, data = rows[0], rows[1], rows[2:]
  col = {h: i for i, h in enumerate(header)}
  kernels = []
  for r in data:
    k = {}
    for name in KEEP:
      if name in col:
        u = units[col[name]]
        k[name] = r[col[name]] + (f' {u}' if u else '')
    kernels.append(k)
  json.dump({'note': note + ' Cold-cache serialised replay: compare shares and byte counts, not absolute times.',
             'kernels': kernels}, open(out_path, 'w'), indent=1)
  for k in kernels:
    print(k['Kernel Name'][:70], k.get('gpu__time_duration.sum'), k.get('dram__bytes_read.sum'), k.get('dram__bytes_write.sum'),
          k.get('lts__t_sector_hit_rate.pct'))


if __name__ == '__main__':
  main()
