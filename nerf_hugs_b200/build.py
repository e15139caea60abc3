"""Builds nerf_hugs_b200/libhugs_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch headers).

Usage: python -m nerf_hugs_b200.build   (also called by __graft_entry__.build()).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libhugs_b200.so')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMMON = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
          '--expt-relaxed-constexpr', '-Xptxas', '-v']
# (source, extra flags).  -fmad=false: sampling/compositing/exact-IPE mirror the oracle's unfused fp32 arithmetic.
SOURCES = [
    ('api.cu', []),
    ('sampling.cu', ['-fmad=false']),
    ('composite.cu', ['-fmad=false']),
    ('nerfacto_ops.cu', ['-fmad=false']),
    ('mlp_simt.cu', ['-fmad=false']),
    ('mlp_tc.cu', []),
    ('mlp_pp.cu', []),
    ('wgrad_tc.cu', []),
    ('dense_tc.cu', []),
    ('layered.cu', []),
    ('field_chain.cu', []),
    ('hashfield.cu', []),
    ('optim.cu', []),
    ('raygen.cu', []),
]


def _stale(obj, src):
  if not os.path.exists(obj):
    return True
  t = os.path.getmtime(obj)
  deps = [src] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
  deps.append(os.path.join(HERE, '..', 'include', 'hugs_b200.h'))
  deps.append(os.path.abspath(__file__))
  return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=False, force=False):
  nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
  objdir = os.path.join(HERE, 'build')
  os.makedirs(objdir, exist_ok=True)
  objs, rebuilt = [], False
  for src, extra in SOURCES:
    s = os.path.join(CSRC, src)
    o = os.path.join(objdir, src.replace('.cu', '.o'))
    objs.append(o)
    if force or _stale(o, s):
      cmd = [nvcc] + ARCH + COMMON + extra + ['-c', s, '-o', o]
      r = subprocess.run(cmd, capture_output=True, text=True)
      log = r.stdout + r.stderr
      with open(o + '.log', 'w') as f:
        f.write(' '.join(cmd) + '\n' + log)
      if r.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError(f'nvcc failed on {src}')
      if verbose:
        print(log)
      rebuilt = True
  if rebuilt or not os.path.exists(OUT):
    cmd = [nvcc] + ARCH + ['-shared', '-o', OUT] + objs + ['-lcudart_static', '-ldl', '-lrt', '-lpthread']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
      sys.stderr.write(r.stdout + r.stderr)
      raise RuntimeError('link failed')
  return OUT


if __name__ == '__main__':
  print(build(verbose='-v' in sys.argv, force='-f' in sys.argv))
