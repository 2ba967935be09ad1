"""CPU tests of the boundary: the C-ABI library loads and exports every
symbol that include/pdune_b200.h declares (no compute without a GPU)."""

import ctypes
import os
import re

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
  assert nat.lib.pd_abi_version() == 1


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


def test_product_never_imports_oracle():
  pkg = os.path.join(ROOT, 'putting-dune_b200')
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      if f.endswith(('.py', '.cu', '.cuh', '.h', '.sh')):
        text = open(os.path.join(dirpath, f)).read()
        assert 'import oracle' not in text and 'from oracle' not in text, f
        assert 'oracle/' not in text, f
