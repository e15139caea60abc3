"""Relative errors of the nerfacto twin on the GPU against the reference's nerfacto.py run on the restated tcnn encodings
(tests/golden/nerfacto_hash.npz).  Usage (GPU box): python scripts/nerfacto_hash_report.py > gpurun_out/nerfacto_hash_parity.json"""
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
  gold = np.load(H.GOLDEN_HASH)
  rep = {}
  for name, case in H.HASH_CASES.items():
    r = {}
    _, model, crit = H.build_hash(name, device=DEV)
    batch = H.load_hash_batch(gold, name, DEV)
    nj = int(gold[f'{name}/n_jitter'])
    if nj:
      model.jitter_override = [torch.from_numpy(gold[f'{name}/jitter/{i}']).to(DEV) for i in range(nj)]
    model.train(case['train'])
    if case['train']:
      outputs = model(batch=batch, curr_step=case['step'], perturb=case['perturb'])
    else:
      with torch.no_grad():
        outputs = model(batch=batch, curr_step=case['step'], perturb=case['perturb'], chunk_size=32)
    for k, v in outputs.items():
      if isinstance(v, list):
        for i, t in enumerate(v):
          r[f'out/{k}/{i}'] = rel(t.detach().cpu().numpy(), gold[f'{name}/out/{k}/{i}'])
      else:
        r[f'out/{k}'] = rel(v.detach().cpu().numpy(), gold[f'{name}/out/{k}'])
    if case['train']:
      n = case['n_rays']
      loss, info, _ = crit(outputs=outputs, batch=batch, data_shape=(n // 16, 4, 4), is_finetune=False, extra_infos={})
      r['loss'] = [float(loss.detach()), float(gold[f'{name}/loss'])]
      for k, v in info.items():
        r[f'info/{k}'] = [float(v), float(gold[f'{name}/info/{k}'])]
      loss.backward()
      for pname, p in model.named_parameters():
        if p.numel() == 0:
          continue
        g = p.grad.detach().cpu().numpy().astype(np.float64).reshape(-1)
        want = gold[f'{name}/gsum/{pname}']
        e = {'norm': [float(np.linalg.norm(g)), float(want[0])],
             'proj_err_over_norm': float(abs(g @ H.projection_vector(g.size, pname) - want[1]) / max(want[0], 1e-30))}
        if f'{name}/grad/{pname}' in gold.files:
          e['rel_l2'] = rel(p.grad.detach().cpu().numpy(), gold[f'{name}/grad/{pname}'])
        r[f'grad/{pname}'] = e
    rep[name] = r
  print(json.dumps(rep, indent=1))


if __name__ == '__main__':
  main()
