"""Shim loader that imports the UNMODIFIED reference hot-path modules.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is on the product path;
only ``tests/``, ``tests/golden/make_golden.py`` and the checker legs of
``__graft_entry__.smoke()`` / ``bench.py`` may import it.

The reference (``/root/reference/putting_dune``) cannot be imported as-is in
this image: shapely, jax, etils, tensorflow, skimage, msgpack_numpy and the
CI-generated ``putting_dune_pb2`` are absent.  This
module installs minimal stand-ins into ``sys.modules`` so that the reference's
own ``simulator.py``, ``graphene.py``, ``geometry.py``, ``microscope_utils.py``
and ``imaging.py`` execute their own code unchanged.  It is used only in the
build container (the GPU box has no ``/root/reference``) to

* validate ``oracle/pdune_oracle.py`` (the restatement), and
* generate the golden vectors committed under ``tests/golden/``.

Nothing here copies reference source; it only provides the third-party names
the reference imports.
"""

from __future__ import annotations

import importlib
import os
import pathlib
import sys
import types

import numpy as np

def _default_root() -> str:
  # /root/reference in the build container; on the GPU box the plain copy of
  # the reference's package that __graft_entry__.build() leaves under
  # baseline/_ref (git-ignored, travels with gpurun).
  if os.path.isdir('/root/reference/putting_dune'):
    return '/root/reference'
  here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  return os.path.join(here, 'baseline', '_ref')


REFERENCE_ROOT = os.environ.get('PDUNE_REFERENCE_ROOT', _default_root())


def reference_available() -> bool:
  return os.path.isdir(os.path.join(REFERENCE_ROOT, 'putting_dune'))


class _Point:
  """Value-semantics stand-in for ``shapely.geometry.Point``.

  The reference needs: ``Point(x, y)``, ``Point((x,
  y))``, ``Point(ndarray[2])``; ``.x``/``.y``; ``np.asarray(p.coords)`` of
  shape (1, 2); equality/hash by value.
  """

  __slots__ = ('_xy',)

  def __init__(self, *args):
    if len(args) == 1:
      xy = np.asarray(args[0], dtype=np.float64).reshape(-1)
    else:
      xy = np.asarray(args, dtype=np.float64).reshape(-1)
    if xy.size != 2:
      raise ValueError(f'Point needs two coordinates, got {xy.size}')
    self._xy = (float(xy[0]), float(xy[1]))

  @property
  def x(self) -> float:
    return self._xy[0]

  @property
  def y(self) -> float:
    return self._xy[1]

  @property
  def coords(self):
    return [self._xy]

  def __eq__(self, other):
    return isinstance(other, _Point) and self._xy == other._xy

  def __hash__(self):
    return hash(self._xy)

  def __repr__(self):
    return f'POINT ({self._xy[0]} {self._xy[1]})'


class _AnyAttrModule(types.ModuleType):
  """Module whose unknown attributes are fresh dummy classes (for *_pb2)."""

  def __getattr__(self, name):
    if name.startswith('__'):
      raise AttributeError(name)
    cls = type(name, (), {'__init__': lambda self, *a, **k: None})
    setattr(self, name, cls)
    return cls


def _module(name: str, **attrs) -> types.ModuleType:
  mod = types.ModuleType(name)
  mod.__dict__.update(attrs)
  sys.modules[name] = mod
  return mod


_loaded = None


def load_reference():
  """Returns the reference ``putting_dune`` package with hot-path modules."""
  global _loaded
  if _loaded is not None:
    return _loaded
  if not reference_available():
    raise RuntimeError(
        f'reference not found under {REFERENCE_ROOT}; the shimmed reference '
        'only exists in the build container'
    )
  import scipy.stats  # pylint: disable=g-import-not-at-top

  # shapely
  if 'shapely' not in sys.modules:
    geo = _module('shapely.geometry', Point=_Point)
    _module('shapely', geometry=geo)
  # jax: only jnp.asarray/cos/sin and jax.scipy.stats.multivariate_normal are
  # reached by the hot path (float64 scipy instead of float32 XLA).
  if 'jax' not in sys.modules:
    jstats = _module('jax.scipy.stats',
                     multivariate_normal=scipy.stats.multivariate_normal)
    jscipy = _module('jax.scipy', stats=jstats)
    jax = _module('jax', numpy=np, scipy=jscipy)
    sys.modules['jax.numpy'] = np
    del jax
  if 'etils' not in sys.modules:
    epath = _module('etils.epath', Path=pathlib.Path)
    _module('etils', epath=epath)
  if 'msgpack_numpy' not in sys.modules:
    # graphene.py:28 `import msgpack_numpy as msgpack`: the restatement of
    # its wire encoding, so that the reference's own (de)serialisation runs
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg_dir = os.path.join(here, 'putting-dune_b200', 'putting_dune_b200')
    import importlib.util  # pylint: disable=g-import-not-at-top
    spec = importlib.util.spec_from_file_location(
        'msgpack_numpy', os.path.join(pkg_dir, 'msgpack_numpy_codec.py'))
    codec = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(codec)
    sys.modules['msgpack_numpy'] = codec
  if 'tensorflow' not in sys.modules:
    _module('tensorflow')
  if 'skimage' not in sys.modules:
    exposure = _module('skimage.exposure')
    util = _module('skimage.util')
    _module('skimage', exposure=exposure, util=util)

  if REFERENCE_ROOT not in sys.path:
    sys.path.insert(0, REFERENCE_ROOT)
  pkg = importlib.import_module('putting_dune')
  # The generated putting_dune_pb2 is not in the tree; pdune_oracle_proto
  # restates putting_dune.proto as a descriptor and has the official protobuf
  # runtime build the same message classes, so the reference's own to_proto /
  # from_proto run for real.  tf.make_tensor_proto / tf.make_ndarray (images
  # inside observations) are stood in for there as well.
  try:
    from oracle import pdune_oracle_proto  # pylint: disable=g-import-not-at-top
    pb2 = pdune_oracle_proto.build_pb2()
    tf = sys.modules['tensorflow']
    tf.make_tensor_proto = pdune_oracle_proto.make_tensor_proto
    tf.make_ndarray = pdune_oracle_proto.make_ndarray
  except ImportError:
    pb2 = _AnyAttrModule('putting_dune.putting_dune_pb2')
  sys.modules['putting_dune.putting_dune_pb2'] = pb2
  pkg.putting_dune_pb2 = pb2

  mods = types.SimpleNamespace()
  for name in ('constants', 'geometry', 'microscope_utils', 'graphene',
               'imaging', 'simulator', 'simulator_observers'):
    setattr(mods, name, importlib.import_module(f'putting_dune.{name}'))
  mods.Point = _Point
  _loaded = mods
  return mods


def extract_reference_function(relpath: str, name: str, namespace: dict,
                               class_name: str | None = None):
  """Compiles ONE function of a reference file that cannot be imported whole.

  ``rate_learning/data_utils.py`` and ``rate_learning/learn_rates.py`` import
  jax/haiku/flax/optax/TF at module level, none of which exist here.  The two
  functions on the hot path (``standardize_beam_and_neighbors`` and
  ``LearnedTransitionRatePredictor.predict``) are plain numpy; this picks the
  function's own AST node out of the unmodified source file and executes just
  that node in ``namespace`` (which must provide the module globals the
  function body reads, e.g. ``np``, ``geometry``, ``constants``).
  """
  import ast  # pylint: disable=g-import-not-at-top

  path = os.path.join(REFERENCE_ROOT, 'putting_dune', relpath)
  tree = ast.parse(open(path).read(), filename=path)
  body = tree.body
  if class_name is not None:
    for node in body:
      if isinstance(node, ast.ClassDef) and node.name == class_name:
        body = node.body
        break
    else:
      raise KeyError(class_name)
  for node in body:
    if isinstance(node, ast.FunctionDef) and node.name == name:
      node.decorator_list = []
      node.returns = None
      for a in node.args.args + node.args.kwonlyargs:
        a.annotation = None
      mod = ast.Module(body=[node], type_ignores=[])
      ast.fix_missing_locations(mod)
      code = compile(mod, path, 'exec')
      ns = dict(namespace)
      exec(code, ns)  # pylint: disable=exec-used
      return ns[name]
  raise KeyError(name)


def reference_standardize_beam_and_neighbors():
  """The reference's ``data_utils.standardize_beam_and_neighbors`` itself."""
  mods = load_reference()
  return extract_reference_function(
      'rate_learning/data_utils.py', 'standardize_beam_and_neighbors',
      {'np': np, 'geometry': mods.geometry})


def reference_learned_predict(packaged_model, use_voltage=False,
                              use_current=False):
  """The reference's ``LearnedTransitionRatePredictor.predict`` bound to a
  stand-in ``self`` whose ``packaged_model`` is the supplied callable
  (context float array [1, D] -> array [1, 4])."""
  mods = load_reference()
  std = reference_standardize_beam_and_neighbors()
  fn = extract_reference_function(
      'rate_learning/learn_rates.py', 'predict',
      {'np': np, 'constants': mods.constants,
       'data_utils': types.SimpleNamespace(
           standardize_beam_and_neighbors=std)},
      class_name='LearnedTransitionRatePredictor')
  fake_self = types.SimpleNamespace(
      packaged_model=packaged_model,
      config=types.SimpleNamespace(use_voltage=use_voltage,
                                   use_current=use_current))
  return lambda grid, beam, si, nbrs: fn(fake_self, grid, beam, si, nbrs)


def load_reference_env_stack():
  """Imports the reference's RL layer (putting_dune_environment, adapters,
  goals, feature constructors, agents, run_helpers, eval_lib) for pinning the
  episode oracle (BASELINE configs[4]).  Adds stand-ins for dm_env,
  matplotlib, frozendict and the plotting module (none are on the path that
  is exercised)."""
  import collections
  import enum
  mods = load_reference()
  if 'dm_env' not in sys.modules:
    class StepType(enum.IntEnum):
      FIRST = 0
      MID = 1
      LAST = 2

    class TimeStep(collections.namedtuple(
        'TimeStep', ['step_type', 'reward', 'discount', 'observation'])):
      def first(self): return self.step_type == StepType.FIRST
      def mid(self): return self.step_type == StepType.MID
      def last(self): return self.step_type == StepType.LAST

    class Environment:
      pass

    class Array:
      def __init__(self, shape, dtype, name=None):
        self.shape, self.dtype, self.name = tuple(shape), np.dtype(dtype), name

    class BoundedArray(Array):
      def __init__(self, shape, dtype, minimum, maximum, name=None):
        super().__init__(shape, dtype, name)
        self.minimum, self.maximum = np.asarray(minimum), np.asarray(maximum)

    specs = _module('dm_env.specs', Array=Array, BoundedArray=BoundedArray)
    _module(
        'dm_env', specs=specs, StepType=StepType, TimeStep=TimeStep,
        Environment=Environment,
        restart=lambda obs: TimeStep(StepType.FIRST, None, None, obs),
        transition=lambda reward, observation, discount=1.0: TimeStep(
            StepType.MID, reward, discount, observation),
        termination=lambda reward, observation: TimeStep(
            StepType.LAST, reward, 0.0, observation),
        truncation=lambda reward, observation, discount=1.0: TimeStep(
            StepType.LAST, reward, discount, observation))
  if 'matplotlib' not in sys.modules:
    plt = _module('matplotlib.pyplot')
    _module('matplotlib', pyplot=plt)
  if 'frozendict' not in sys.modules:
    _module('frozendict', frozendict=dict)
  if 'putting_dune.plotting_utils' not in sys.modules:
    pu = types.ModuleType('putting_dune.plotting_utils')
    sys.modules['putting_dune.plotting_utils'] = pu
    sys.modules['putting_dune'].plotting_utils = pu
  for name in ('action_adapters', 'goals', 'feature_constructors',
               'putting_dune_environment', 'run_helpers', 'eval_lib'):
    setattr(mods, name, importlib.import_module(f'putting_dune.{name}'))
  mods.agent_lib = importlib.import_module('putting_dune.agents.agent_lib')
  return mods
