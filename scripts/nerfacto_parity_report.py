"""Relative errors of the vanilla-NeRF torch twin on the GPU against the reference's own outputs
(tests/golden/nerfacto_nerf.npz): every output, the loss and every gradient tensor, in both precision modes.
Usage (GPU box): python scripts/nerfacto_parity_report.py > gpurun_out/nerfacto_parity.json"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import nerfacto_helpers as H   # noqa: E402

DEV = 'cuda:0'


def rel(a, b):
  a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
  return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def main():
  gold = np.load(H.GOLDEN)
  rep = {}
  for precision in ('tc_split', 'bf16_tc'):
    for name, case in H.CASES.items():
      r = {}
      _, model, crit = H.build(name, precision=precision, device=DEV)
      model.last_bins = {}
      batch = H.load_batch(gold, name, DEV)
      if case['train'] and f'{name}/bins/fine' in gold.files and os.environ.get('OVERRIDE_FINE', '1') == '1':
        model.bins_override = {'fine': torch.from_numpy(gold[f'{name}/bins/fine']).to(DEV)}   # see tests/test_gpu_nerfacto_nerf.py
      if int(gold[f'{name}/n_jitter']):
        model.jitter_override = {'coarse': torch.from_numpy(gold[f'{name}/jitter/0']).to(DEV),
                                 'fine': torch.from_numpy(gold[f'{name}/jitter/1']).to(DEV)}
      model.train(case['train'])
      if case['train']:
        outputs = model(batch=batch, curr_step=1, perturb=case['perturb'])
      else:
        with torch.no_grad():
          outputs = model(batch=batch, curr_step=1, perturb=case['perturb'], chunk_size=32)
      for k, v in outputs.items():
        r[f'out/{k}'] = rel(v.detach().cpu().numpy(), gold[f'{name}/out/{k}'])
      for ft in (('coarse', 'fine') if f'{name}/bins/fine' in gold.files else ()):
        b = torch.cat(model.last_bins[ft]).cpu().numpy()
        d = np.abs(b - gold[f'{name}/bins/{ft}'])
        r[f'bins/{ft}'] = {'max_abs': float(d.max()), 'median_abs': float(np.median(d))}
      if case['train']:
        n = case['n_rays']
        loss, info, _ = crit(outputs=outputs, batch=batch, data_shape=(n // 16, 4, 4), is_finetune=False, extra_infos={})
        r['loss'] = [float(loss), float(gold[f'{name}/loss'])]
        loss.backward()
        for pname, p in model.named_parameters():
          g = p.grad.detach().cpu().numpy().astype(np.float64).reshape(-1)
          want = gold[f'{name}/gsum/{pname}']
          e = {'norm': [float(np.linalg.norm(g)), float(want[0])],
               'proj_err_over_norm': float(abs(g @ H.projection_vector(g.size, pname) - want[1]) / max(want[0], 1e-30))}
          if f'{name}/grad/{pname}' in gold.files:
            e['rel_l2'] = rel(p.grad.detach().cpu().numpy(), gold[f'{name}/grad/{pname}'])
          r[f'grad/{pname}'] = e
      rep[f'{precision}/{name}'] = r
  print(json.dumps(rep, indent=1))


if __name__ == '__main__':
  main()
