"""CPU tests of the boundary: the C-ABI library loads and exports every
symbol that include/pdune_b200.h declares (no compute without a GPU)."""

import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'pdune_b200.h')


def _declared_functions():
  text = open(HEADER).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(pd_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_entry_points():
  names = _declared_functions()
  for must in ('pd_reset', 'pd_rates', 'pd_apply_control', 'pd_step_and_image',
               'pd_step_and_image_host', 'pd_rollout',
               'pd_get_atoms_in_bounds', 'pd_get_silicon_position',
               'pd_build_lattice'):
    assert must in names


def test_library_exports_every_declared_symbol():
  from putting_dune_b200 import _native as nat
  lib = ctypes.CDLL(str(nat.LIB_PATH))
  for name in _declared_functions():
    assert hasattr(lib, name), f'{name} declared in header but not exported'
  assert nat.lib.pd_abi_version() == 2  # 2: pd_rate_config.prior


def test_python_binding_covers_header():
  from putting_dune_b200 import _native as nat
  import putting_dune_b200  # binds optional entry points  # noqa: F401
  assert set(_declared_functions()) <= set(nat.exported_symbols())


def test_lattice_size_host_arithmetic():
  from putting_dune_b200 import _native as nat
  n, r = ctypes.c_int32(), ctypes.c_int32()
  nat.check(nat.lib.pd_lattice_size(50, ctypes.byref(n), ctypes.byref(r)))
  assert (n.value, r.value) == (1881, 57)
  from oracle import pdune_oracle as po
  for cols in (4, 10, 11, 12, 50, 64):
    nat.check(nat.lib.pd_lattice_size(cols, ctypes.byref(n), ctypes.byref(r)))
    assert n.value == po.hexagonal_grid(cols).shape[0]
  with pytest.raises(nat.NativeError):
    nat.check(nat.lib.pd_lattice_size(2, ctypes.byref(n), ctypes.byref(r)))


def test_no_cpu_fallback_without_device():
  import torch
  if torch.cuda.is_available():
    pytest.skip('a GPU is present')
  import putting_dune_b200 as pd
  with pytest.raises(RuntimeError, match='no CPU fallback'):
    pd.BatchedSimulator(4)
  with pytest.raises(RuntimeError, match='no CPU fallback'):
    pd.graphene.PristineSingleDopedGraphene().reset(
        pd.graphene.PhiloxKey(0))


def test_synthetic_data_host_side():
  """rate_learning/data_utils.py host logic that needs no device: the
  NETWORK mode's weight initialiser (Haiku hk.Linear defaults:
  TruncatedNormal(stddev = 1 / sqrt(fan_in)) cut at two standard deviations,
  zero biases; learn_rates.py:80-99 with hidden (1, 64),
  data_utils.py:196-201) and the loud failure without a CUDA device."""
  import torch
  from putting_dune_b200.rate_learning import data_utils
  net = data_utils.init_network(3, 4, num_states=3)
  assert net['w0'].shape == (4, 1) and net['w1'].shape == (1, 64)
  assert net['w2'].shape == (64, 4) and net['b2'].shape == (4,)
  for k, fan_in in (('w0', 4), ('w1', 1), ('w2', 64)):
    w = net[k]
    assert w.dtype == np.float32
    assert np.abs(w).max() <= 2.0 / np.sqrt(fan_in) + 1e-6
  assert not net['b0'].any() and not net['b1'].any() and not net['b2'].any()
  big = data_utils.init_network(3, 4, num_states=5, hidden=(32, 200))
  assert abs(big['w1'].std() * np.sqrt(32) - 0.88) < 0.05  # truncated normal
  again = data_utils.init_network(3, 4, num_states=3)
  np.testing.assert_array_equal(net['w1'], again['w1'])
  assert data_utils.SyntheticDataType('network') == \
      data_utils.SyntheticDataType.NETWORK
  if not torch.cuda.is_available():
    for mode in ('prior', 'network'):
      with pytest.raises(RuntimeError, match='no CPU fallback'):
        data_utils.generate_synthetic_data(num_data=4, data_seed=1, mode=mode)


def test_product_never_imports_oracle():
  pkg = os.path.join(ROOT, 'putting-dune_b200')
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      if f.endswith(('.py', '.cu', '.cuh', '.h', '.sh')):
        text = open(os.path.join(dirpath, f)).read()
        assert 'import oracle' not in text and 'from oracle' not in text, f
        assert 'oracle/' not in text, f
