"""Summarises an `ncu --set full` capture of one training step (scripts/profile_fwd.py train) into profiles/.

    ncu --set full --clock-control none --import-source on -k regex:'mlp_pp_kernel|wgrad_kernel' -s 12 -c 6 \
        -o gpurun_out/r02_train_full python scripts/profile_fwd.py train 4          # on the GPU box (gpurun)
    python scripts/ncu_summary.py gpurun_out/r02_train_full.ncu-rep r02                # here (no GPU needed)

Writes profiles/<tag>_ncu_full_train_step.json (selected metrics of every captured launch) and
profiles/ncu_dram_bytes.json (dram__bytes_read.sum + dram__bytes_write.sum per launch and kernel class: bench.py's
`roofline.traffic`).  The six MLP kernels of a step launch in this order: PropMLP forward chain, NerfMLP forward chain,
NerfMLP dgrad chain, NerfMLP weight gradients, PropMLP dgrad chain, PropMLP weight gradients.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORDER = ['chain_fwd_prop', 'chain_fwd_nerf', 'chain_bwd_nerf', 'wgrad_nerf', 'chain_bwd_prop', 'wgrad_prop']
KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__cluster_dim_x', 'smsp__cycles_active.avg']


def to_bytes(v, unit):
  v = float(str(v).replace(',', ''))
  u = (unit or '').lower()
  mult = {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'tbyte': 1e12}.get(u, 1)
  return v * mult


def main():
  rep, tag = sys.argv[1], sys.argv[2]
  out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
  rows = list(csv.reader(io.StringIO(out)))
  header, units, data = rows[0], rows[1], rows[2:]
  col = {h: i for i, h in enumerate(header)}
  kernels, dram = [], {}
  # the capture may start anywhere inside a step (e.g. view_extra_wgrad_kernel also matches the regex "wgrad_kernel"):
  # find the rotation of the launch order whose chain / wgrad pattern fits the captured names
  names = ['wgrad' if 'wgrad_kernel' in r[col['Kernel Name']] else 'chain' for r in data]
  expect = ['wgrad' if c.startswith('wgrad') else 'chain' for c in ORDER]
  rot = next((k for k in range(len(ORDER)) if len(data) == len(ORDER) and
              all(names[i] == expect[(i + k) % len(ORDER)] for i in range(len(ORDER)))), None)
  for i, r in enumerate(data):
    k = {}
    for name in KEEP:
      if name in col:
        u = units[col[name]]
        k[name] = r[col[name]] + (f' {u}' if u else '')
    total = to_bytes(r[col['dram__bytes_read.sum']], units[col['dram__bytes_read.sum']]) + \
        to_bytes(r[col['dram__bytes_write.sum']], units[col['dram__bytes_write.sum']])
    k['dram_bytes_total'] = total
    if rot is not None:
      k['class'] = ORDER[(i + rot) % len(ORDER)]
      dram[k['class']] = total
    kernels.append(k)
  note = (f'ncu --set full --clock-control none of the six MLP kernels of one training step (config A, 4096 rays, '
          f'scripts/profile_fwd.py train), capture {os.path.basename(rep)}. Cold-cache serialised replay: compare shares '
          f'and byte counts, not absolute times.')
  path = os.path.join(ROOT, 'profiles', f'{tag}_ncu_full_train_step.json')
  json.dump({'note': note, 'kernels': kernels}, open(path, 'w'), indent=1)
  if dram:
    json.dump({'source': f'profiles/{tag}_ncu_full_train_step.json (ncu --set full, one training step, 4096 rays)',
               'bytes_per_launch': dram}, open(os.path.join(ROOT, 'profiles', 'ncu_dram_bytes.json'), 'w'), indent=1)
  print('wrote', path, {k: round(v / 1e9, 3) for k, v in dram.items()})


if __name__ == '__main__':
  main()
