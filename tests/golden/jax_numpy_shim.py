"""A `jax` / `flax` / `gin` stand-in backed by NumPy, so that the REFERENCE'S OWN Mip-NeRF 360 source
(`/root/reference/MipNeRF360/internal/{math,stepfun,coord,render,models}.py`, and functions cut out of
`train_utils.py`) can be executed in this container, where JAX is not installable.

TEST INFRASTRUCTURE ONLY (used by `tests/golden/make_golden_mipnerf360.py` to generate fixtures; never imported
by the product or at test time on the GPU box).

What is faithful and what is not:
  * every array stays float32 (creation functions default to float32, Python scalars are weakly typed under NumPy 2 /
    NEP 50 exactly like under JAX), `%` is the NumPy remainder JAX uses, `jnp.interp` / `jax.nn.softmax` /
    `jax.nn.softplus` / `jax.nn.sigmoid` follow the published jax implementations op by op;
  * `jax.linearize` is a forward-mode (dual number) JVP over the handful of primitives `coord.contract` uses, with
    the textbook JVP rules (float32);
  * reductions use NumPy's summation order and transcendental functions are libm's, not XLA's: results agree with a
    real JAX run to fp32 rounding, not bit for bit (the tolerances of the tests that read these fixtures say so);
  * `jax.random` is not reproduced: a `KeyStream` hands out caller-provided uniform draws in call order.
"""
import contextlib
import dataclasses
import sys
import types

import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------------------------
# forward-mode dual numbers (for jax.linearize of coord.contract)
# ----------------------------------------------------------------------------------------------
class Dual:
  """primal + eps * tangent, float32.  Supports exactly what coord.contract needs."""
  __array_priority__ = 1000

  def __init__(self, p, t):
    self.p = np.asarray(p, F32)
    self.t = np.asarray(t, F32)

  @property
  def shape(self):
    return self.p.shape

  def __pow__(self, k):
    assert k == 2
    return Dual(self.p ** 2, F32(2) * self.p * self.t)

  def __mul__(self, o):
    if isinstance(o, Dual):
      return Dual(self.p * o.p, self.t * o.p + self.p * o.t)
    return Dual(self.p * o, self.t * o)

  __rmul__ = __mul__

  def __truediv__(self, o):
    if isinstance(o, Dual):
      q = self.p / o.p
      return Dual(q, (self.t - q * o.t) / o.p)
    return Dual(self.p / o, self.t / o)

  def __sub__(self, o):
    if isinstance(o, Dual):
      return Dual(self.p - o.p, self.t - o.t)
    return Dual(self.p - o, self.t)

  def __rsub__(self, o):
    return Dual(o - self.p, -self.t)

  def __add__(self, o):
    if isinstance(o, Dual):
      return Dual(self.p + o.p, self.t + o.t)
    return Dual(self.p + o, self.t)

  __radd__ = __add__

  def __le__(self, o):
    return self.p <= (o.p if isinstance(o, Dual) else o)


def _primal(x):
  return x.p if isinstance(x, Dual) else x


# ----------------------------------------------------------------------------------------------
# jax.numpy
# ----------------------------------------------------------------------------------------------
class _Jnp(types.ModuleType):
  """numpy with jax.numpy's dtype defaults; unknown names fall through to numpy."""

  def __getattr__(self, name):
    if name.startswith('__'):
      raise AttributeError(name)
    return getattr(np, name)


jnp = _Jnp('jax.numpy')
jnp.ndarray = np.ndarray
jnp.float32 = np.float32
jnp.int32 = np.int32
jnp.pi = np.pi
jnp.inf = np.inf
jnp.finfo = np.finfo
jnp.linalg = np.linalg


class JaxInt(np.ndarray):
  """int32 array with jax's promotion: int32 (op) float32 -> float32, int32 / int -> float32 (NumPy would give float64)."""

  def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
    others_float = any(isinstance(i, (np.ndarray, np.floating, float)) and np.asarray(i).dtype.kind == 'f'
                       for i in inputs if not isinstance(i, JaxInt))
    to_float = others_float or ufunc in (np.true_divide, np.sqrt, np.exp, np.log)
    conv = []
    for i in inputs:
      if isinstance(i, JaxInt):
        i = i.view(np.ndarray)
        if to_float:
          i = i.astype(F32)
      conv.append(i)
    out = getattr(ufunc, method)(*conv, **kwargs)
    if isinstance(out, np.ndarray) and out.dtype.kind == 'i':
      return out.astype(np.int32).view(JaxInt)
    if isinstance(out, np.ndarray) and out.dtype == np.float64:
      raise TypeError('float64 crept into a jax.numpy expression')
    return out


def _f32_default(x):
  x = np.asarray(x)
  if x.dtype == np.float64:
    return x.astype(F32)
  if x.dtype == np.int64:
    return x.astype(np.int32).view(JaxInt)
  return x


jnp.array = lambda x, dtype=None: np.asarray(x, dtype) if dtype is not None else _f32_default(x)
jnp.asarray = jnp.array
jnp.zeros = lambda shape, dtype=F32: np.zeros(shape, dtype)
jnp.ones = lambda shape, dtype=F32: np.ones(shape, dtype)
jnp.eye = lambda n, dtype=F32: np.eye(n, dtype=dtype)
jnp.arange = lambda *a, **k: _f32_default(np.arange(*a, **k))
jnp.linspace = lambda *a, **k: np.linspace(*a, **k).astype(F32)   # float64 linspace rounded once (canonical choice)
jnp.full_like = lambda a, v: np.full_like(a, v)
jnp.copy = np.copy


def _matmul(a, b, precision=None):
  del precision   # float32 BLAS: nothing below fp32 to choose from on a CPU
  return np.matmul(a, b)


jnp.matmul = _matmul


def _sum(x, axis=None, keepdims=False):
  if isinstance(x, Dual):
    return Dual(np.sum(x.p, axis=axis, keepdims=keepdims), np.sum(x.t, axis=axis, keepdims=keepdims))
  return np.sum(x, axis=axis, keepdims=keepdims)


def _maximum(a, b):
  if isinstance(a, Dual) or isinstance(b, Dual):
    pa, pb = _primal(a), _primal(b)
    ta = a.t if isinstance(a, Dual) else np.zeros_like(pb)
    tb = b.t if isinstance(b, Dual) else np.zeros_like(pa)
    return Dual(np.maximum(pa, pb), np.where(pa > pb, ta, tb))
  return np.maximum(a, b)


def _sqrt(x):
  if isinstance(x, Dual):
    s = np.sqrt(x.p)
    return Dual(s, x.t / (F32(2) * s))
  return np.sqrt(x)


def _where(c, a, b):
  if isinstance(a, Dual) or isinstance(b, Dual):
    pa, pb = _primal(a), _primal(b)
    ta = a.t if isinstance(a, Dual) else np.zeros_like(pa)
    tb = b.t if isinstance(b, Dual) else np.zeros_like(pb)
    return Dual(np.where(c, pa, pb), np.where(c, ta, tb))
  return np.where(c, a, b)


jnp.sum, jnp.maximum, jnp.sqrt, jnp.where = _sum, _maximum, _sqrt, _where


def _interp(x, xp, fp, left=None, right=None, period=None):
  """jax.numpy.interp (jax/_src/numpy/lax_numpy.py::_interp), float32 throughout."""
  assert left is None and right is None and period is None
  x, xp, fp = np.asarray(x, F32), np.asarray(xp, F32), np.asarray(fp, F32)
  i = np.clip(np.searchsorted(xp, x, side='right'), 1, len(xp) - 1)
  df = fp[i] - fp[i - 1]
  dx = xp[i] - xp[i - 1]
  delta = x - xp[i - 1]
  epsilon = np.spacing(np.finfo(xp.dtype).eps)
  dx0 = np.abs(dx) <= epsilon
  with np.errstate(all='ignore'):
    f = np.where(dx0, fp[i - 1], fp[i - 1] + (delta / np.where(dx0, F32(1), dx)) * df)
  f = np.where(x < xp[0], fp[0], f)
  f = np.where(x > xp[-1], fp[-1], f)
  return f.astype(F32)


jnp.interp = _interp
jnp.vectorize = np.vectorize


# ----------------------------------------------------------------------------------------------
# jax
# ----------------------------------------------------------------------------------------------
jax = types.ModuleType('jax')
jax.numpy = jnp


def _vmap(fn, in_axes=0, out_axes=0):
  def mapped(*args):
    axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
    n = args[0].shape[axes[0]]
    outs = [fn(*[np.take(a, i, axis=ax) for a, ax in zip(args, axes)]) for i in range(n)]
    return np.stack(outs, axis=out_axes)
  return mapped


def _linearize(fn, x):
  x = np.asarray(x, F32)
  y = fn(x)

  def lin(v):
    return fn(Dual(x, v)).t
  return y, lin


def _custom_jvp(fn):
  fn.defjvp = lambda g: g
  return fn


jax.vmap = _vmap
jax.linearize = _linearize
jax.custom_jvp = _custom_jvp
jax.process_index = lambda: 0
jax.process_count = lambda: 1
jax.device_count = lambda: 1
jax.local_device_count = lambda: 1

jax.lax = types.ModuleType('jax.lax')
jax.lax.stop_gradient = lambda x: x
jax.lax.Precision = types.SimpleNamespace(HIGHEST='highest')

jax.nn = types.ModuleType('jax.nn')


def _softmax(x, axis=-1):
  """jax.nn.softmax: exp(x - max) / sum(exp(x - max))."""
  with np.errstate(all='ignore'):
    un = np.exp(x - np.max(x, axis=axis, keepdims=True))
    return un / np.sum(un, axis=axis, keepdims=True)


jax.nn.softmax = _softmax
jax.nn.softplus = lambda x: np.logaddexp(x, F32(0))                      # jax.nn.softplus = logaddexp(x, 0)
jax.nn.sigmoid = lambda x: (F32(1) / (F32(1) + np.exp(-x))).astype(F32)  # lax.logistic
jax.nn.relu = lambda x: np.maximum(x, F32(0))
jax.nn.initializers = types.SimpleNamespace(he_uniform=lambda: 'he_uniform', glorot_uniform=lambda: 'glorot_uniform')


def _tree_map(f, tree, *rest):
  if isinstance(tree, dict):
    return {k: _tree_map(f, v, *[r[k] for r in rest]) for k, v in tree.items()}
  if dataclasses.is_dataclass(tree):
    return type(tree)(**{fl.name: _tree_map(f, getattr(tree, fl.name), *[getattr(r, fl.name) for r in rest])
                         for fl in dataclasses.fields(tree)})
  if isinstance(tree, (list, tuple)):
    return type(tree)(_tree_map(f, v, *[r[i] for r in rest]) for i, v in enumerate(tree))
  return f(tree, *rest)


def _tree_leaves(tree):
  if isinstance(tree, dict):
    return [l for v in tree.values() for l in _tree_leaves(v)]
  if isinstance(tree, (list, tuple)):
    return [l for v in tree for l in _tree_leaves(v)]
  return [tree]


def _tree_reduce(f, tree, initializer=None):
  import functools
  leaves = _tree_leaves(tree)
  return functools.reduce(f, leaves) if initializer is None else functools.reduce(f, leaves, initializer)


jax.tree_util = types.SimpleNamespace(tree_map=_tree_map, tree_reduce=_tree_reduce, tree_leaves=_tree_leaves)


class KeyStream:
  """Stands in for a jax PRNG key: `random.split` hands the stream on, `random.uniform(key, shape, maxval=m)`
  returns the next caller-provided array of unit-uniform draws times `m`."""

  def __init__(self, draws):
    self.draws = list(draws)
    self.splits = 0

  def take(self, shape):
    d = np.asarray(self.draws.pop(0), F32)
    assert d.shape == tuple(shape), (d.shape, shape)
    return d


class _Key:
  def __init__(self, stream, index):
    self.stream, self.index = stream, index


def _split(rng, num=2):
  assert num == 2
  if isinstance(rng, _Key):      # a sub-key split again (MLP.__call__): no draws are taken from it at default config
    return _Key(rng.stream, -1), rng
  assert isinstance(rng, KeyStream)
  rng.splits += 1
  return _Key(rng, rng.splits - 1), rng


def _uniform(key, shape=(), minval=0., maxval=1.):
  return (key.stream.take(shape) * F32(maxval - minval) + F32(minval)).astype(F32)


jax.random = types.ModuleType('jax.random')
jax.random.split = _split
jax.random.uniform = _uniform
jax.random.normal = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError('random.normal'))


# ----------------------------------------------------------------------------------------------
# flax.linen (just enough for models.Model / models.MLP)
# ----------------------------------------------------------------------------------------------
_scope_stack = []     # [(module instance, {class name: next index})]
_params = [None]      # the parameter tree of the apply() in flight


class _Module:
  def __init_subclass__(cls, **kw):
    super().__init_subclass__(**kw)
    dataclasses.dataclass(cls, eq=False)
    for name in ('__call__',):
      fn = cls.__dict__.get(name)
      if fn is not None and not getattr(fn, '_scoped', False):
        setattr(cls, name, _scoped(fn))

  def __post_init__(self):
    if _scope_stack:
      parent, counters = _scope_stack[-1]
      base = type(self).__name__.lstrip('_')
      idx = counters.get(base, 0)
      counters[base] = idx + 1
      self._path = parent._path + (f'{base}_{idx}',)
    else:
      self._path = ()
    if hasattr(self, 'setup'):
      self.setup()

  def _tree(self):
    node = _params[0]
    for k in self._path:
      node = node[k]
    return node

  def apply(self, variables, *args, **kwargs):
    _params[0] = variables['params']
    try:
      return self(*args, **kwargs)
    finally:
      _params[0] = None


def _scoped(fn):
  def wrapped(self, *a, **k):
    _scope_stack.append((self, {}))      # auto-name counters restart on every call (flax: weight sharing)
    try:
      return fn(self, *a, **k)
    finally:
      _scope_stack.pop()
  wrapped._scoped = True
  return wrapped


class _Dense(_Module):
  features: int = 0
  kernel_init: object = None

  def __call__(self, x):
    p = self._tree()
    k, b = np.asarray(p['kernel'], F32), np.asarray(p['bias'], F32)
    assert k.shape == (x.shape[-1], self.features), (self._path, k.shape, x.shape, self.features)
    return (np.matmul(x, k) + b).astype(F32)


class _Embed(_Module):
  num_embeddings: int = 0
  features: int = 0

  def __call__(self, idx):
    return np.asarray(self._tree()['embedding'], F32)[np.asarray(idx)]


nn = types.ModuleType('flax.linen')
nn.Module = _Module
nn.compact = lambda fn: fn
nn.Dense = _Dense
nn.Embed = _Embed
nn.relu, nn.softplus, nn.sigmoid = jax.nn.relu, jax.nn.softplus, jax.nn.sigmoid

flax = types.ModuleType('flax')
flax.linen = nn
flax.struct = types.SimpleNamespace(dataclass=dataclasses.dataclass)
flax.core = types.ModuleType('flax.core')
flax.core.FrozenDict = dict
flax.core.freeze = lambda x: x


# ----------------------------------------------------------------------------------------------
# gin
# ----------------------------------------------------------------------------------------------
gin = types.ModuleType('gin')


gin.bindings = {}    # {'NerfMLP': {'net_width': 256, ...}}: what `NerfMLP.net_width = 256` in a .gin file does


def _configurable(x=None, **kw):
  def wrap(cls):
    def make(*a, **k):
      merged = dict(gin.bindings.get(cls.__name__, {}))
      merged.update(k)
      return cls(*a, **merged)
    make.__name__ = cls.__name__
    make.cls = cls
    return make
  if callable(x):
    return wrap(x)
  return wrap


gin.configurable = _configurable
gin.config = types.SimpleNamespace(external_configurable=lambda *a, **k: None)


class _Anything(types.ModuleType):
  """Module whose every attribute is a harmless placeholder (for imports the executed code paths never touch)."""

  def __getattr__(self, k):
    if k.startswith('__'):
      raise AttributeError(k)
    return type(k, (), {})


@contextlib.contextmanager
def installed(extra_stubs=()):
  """Temporarily registers the stand-ins in sys.modules."""
  mods = {'jax': jax, 'jax.numpy': jnp, 'jax.nn': jax.nn, 'jax.lax': jax.lax, 'jax.random': jax.random,
          'flax': flax, 'flax.linen': nn, 'flax.core': flax.core, 'gin': gin}
  for n in extra_stubs:
    mods[n] = _Anything(n)
  saved = {n: sys.modules.get(n) for n in mods}
  sys.modules.update(mods)
  try:
    yield
  finally:
    for n, m in saved.items():
      if m is None:
        sys.modules.pop(n, None)
      else:
        sys.modules[n] = m
