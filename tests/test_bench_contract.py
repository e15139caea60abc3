"""bench.py --impl reference (the CPU arm of the measurement contract) runs without a GPU and prints one JSON line with
the contract's keys; the CUDA arm refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
  r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                      '--ref-rays', '8'], capture_output=True, text=True, timeout=300, cwd=ROOT)
  assert r.returncode == 0, r.stderr[-2000:]
  line = json.loads(r.stdout.strip().splitlines()[-1])
  for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
            'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
    assert k in line, k
  assert line['impl'] == 'reference' and line['unit'] == 'rays/s' and line['higher_is_better'] is True
  assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
  assert line['e2e'] == {'value': line['value'], 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
  assert line['value'] > 0 and 'workload' in line['config']


def test_cuda_arm_fails_loudly_without_a_device():
  if torch.cuda.is_available():
    return
  r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '1'],
                     capture_output=True, text=True, timeout=300, cwd=ROOT)
  assert r.returncode != 0 and 'no CPU fallback' in (r.stderr + r.stdout)
