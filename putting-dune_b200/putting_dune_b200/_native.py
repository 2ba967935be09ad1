"""ctypes binding of ``libpdune_b200.so`` (C ABI in ``include/pdune_b200.h``).

There is no CPU fallback: if the CUDA library is missing or was not built,
importing this module raises, and so does every product entry point.
"""

from __future__ import annotations

import ctypes as C
import os
import pathlib

_HERE = pathlib.Path(__file__).resolve().parent
LIB_PATH = pathlib.Path(
    os.environ.get('PDUNE_B200_LIB', _HERE.parent / 'lib' / 'libpdune_b200.so'))

PD_OK = 0
RATE_SIMPLE, RATE_PRIOR, RATE_LEARNED, RATE_CONSTANT, RATE_GMM = 0, 1, 2, 3, 4
ENV_BAD_RATE, ENV_LOG_OVERFLOW, ENV_NOT_RESET = 1, 2, 4
STREAM_KMC, STREAM_RESET = 0, 1
ACTION_DIRECT, ACTION_RELATIVE_TO_SILICON = 0, 1
(ADAPTER_DIRECT, ADAPTER_DELTA, ADAPTER_RELATIVE,
 ADAPTER_RELATIVE_MATERIAL) = range(4)
FEATURES_MICROSCOPE, FEATURES_MATERIAL = 0, 1
STEP_FIRST, STEP_MID, STEP_LAST = 0, 1, 2
(RENDER_CLEAN, RENDER_BLUR, RENDER_POISSON, RENDER_JITTER, RENDER_UNIFORM,
 RENDER_EXPONENTIAL, RENDER_GAUSSIAN, RENDER_FINAL) = range(8)

_p = C.c_void_p


class PdLattice(C.Structure):
  _fields_ = [('n_cols', C.c_int32), ('n_sites', C.c_int32),
              ('base_xy', _p), ('nbr', _p)]


class PdState(C.Structure):
  _fields_ = [('n_envs', C.c_int64), ('seed', C.c_uint64),
              ('env_offset', C.c_uint32), ('reserved_', C.c_uint32),
              ('si_idx', _p), ('lattice', _p), ('fov', _p), ('fov_scale', _p),
              ('image_params', _p), ('episode', _p), ('ctrl_count', _p),
              ('frame_count', _p), ('sim_time_us', _p), ('n_events', _p),
              ('n_transitions', _p), ('status', _p)]


class PdMlp(C.Structure):
  _fields_ = [('context_dim', C.c_int32), ('hidden1', C.c_int32),
              ('hidden2', C.c_int32), ('batchnorm', C.c_int32),
              ('bn_scale', _p), ('bn_offset', _p), ('bn_mean', _p),
              ('bn_var', _p), ('w0', _p), ('b0', _p), ('w1', _p), ('b1', _p),
              ('w2', _p), ('b2', _p), ('tensor_core', C.c_int32),
              ('reserved_', C.c_int32), ('w1_umma', _p)]


GMM_MAX_MIXTURES = 16


class PdGmm(C.Structure):
  _fields_ = [('n_mixtures', C.c_int32), ('reserved_', C.c_int32),
              ('max_rate', C.c_double),
              ('mixture_weights', C.c_double * GMM_MAX_MIXTURES),
              ('loc_distances', C.c_double * GMM_MAX_MIXTURES),
              ('variances', (C.c_double * 2) * GMM_MAX_MIXTURES)]


class PdPrior(C.Structure):
  """pd_prior: HumanPriorRatePredictor(mean, cov, max_rate)."""
  _fields_ = [('mean', C.c_double * 2), ('cov', (C.c_double * 2) * 2),
              ('max_rate', C.c_double)]


class PdRateConfig(C.Structure):
  _fields_ = [('rate_fn', C.c_int32), ('reserved_', C.c_int32),
              ('mlp', C.POINTER(PdMlp)), ('constant_rates', C.c_float * 3),
              ('reserved2_', C.c_float), ('gmm', C.POINTER(PdGmm)),
              ('prior', C.POINTER(PdPrior))]


class PdStepOut(C.Structure):
  _fields_ = [('elapsed_us', _p), ('transitions', _p), ('events', _p),
              ('recentred', _p), ('si_xy', _p), ('log_capacity', C.c_int32),
              ('reserved_', C.c_int32), ('log_count', _p),
              ('log_elapsed_us', _p), ('log_site', _p), ('log_ctrl', _p)]


class PdEnvConfig(C.Structure):
  _fields_ = [('adapter', C.c_int32), ('features', C.c_int32),
              ('action_dim', C.c_int32), ('step_limit', C.c_int32),
              ('min_dwell_s', C.c_double), ('max_dwell_s', C.c_double),
              ('max_distance_angstroms', C.c_double),
              ('image_duration_us', C.c_int64)]


class PdEnvBuffers(C.Structure):
  _fields_ = [('goal_xy', _p), ('beam_pos', _p), ('elapsed_steps', _p),
              ('needs_reset', _p), ('controls_xy', _p), ('dwell_us', _p),
              ('elapsed_us', _p), ('resetting', _p)]


class PdEpisodeConfig(C.Structure):
  _fields_ = [('dwell_us', C.c_int64), ('image_duration_us', C.c_int64),
              ('timeout_us', C.c_int64), ('step_limit', C.c_int32),
              ('reserved_', C.c_int32), ('argmax_x', C.c_double),
              ('argmax_y', C.c_double)]


class NativeError(RuntimeError):
  pass


def _load() -> C.CDLL:
  if not LIB_PATH.exists():
    raise ImportError(
        f'{LIB_PATH} not found: build the CUDA extension first '
        '(python -c "import __graft_entry__ as g; g.build()" or '
        'putting-dune_b200/build.sh). There is no CPU fallback.')
  return C.CDLL(str(LIB_PATH))


lib = _load()

_LP, _SP, _RP, _OP = (C.POINTER(PdLattice), C.POINTER(PdState),
                      C.POINTER(PdRateConfig), C.POINTER(PdStepOut))
_i32, _i64 = C.c_int32, C.c_int64

_SIGNATURES = {
    'pd_abi_version': ([], C.c_int),
    'pd_last_error': ([], C.c_char_p),
    'pd_device_sm_count': ([C.POINTER(C.c_int)], C.c_int),
    'pd_lattice_size': ([_i32, C.POINTER(_i32), C.POINTER(_i32)], C.c_int),
    'pd_build_lattice': ([_i32, _p, _p, _p], C.c_int),
    'pd_reset': ([_LP, _SP, _p, _p], C.c_int),
    'pd_rates': ([_LP, _SP, _RP, _p, _p, _p, _p], C.c_int),
    'pd_apply_control': ([_LP, _SP, _RP, _p, _p, _i64, _OP, _p], C.c_int),
    'pd_step_and_image': ([_LP, _SP, _RP, _p, _p, _i64, _i32, _i64, _OP, _p],
                          C.c_int),
    'pd_step_and_image_host': ([_LP, _SP, _RP, _p, _p, _i64, _i32, _i64, _p,
                                _p, _OP, _p, _p, _p, _p], C.c_int),
    'pd_rollout': ([_LP, _SP, _RP, _p, _i64, _i32, _i64, _p, _p, _p], C.c_int),
    'pd_rollout_actions': ([_LP, _SP, _RP, _p, _i32, C.c_double, _i64, _i32,
                            _i64, _p, _p, _p], C.c_int),
    'pd_rollout_actions_host': ([_LP, _SP, _RP, _p, _i32, C.c_double, _i64,
                                 _i32, _i64, _p, _p, _p, _p, _p, _p],
                                C.c_int),
    'pd_rollout_actions_host_f32': ([_LP, _SP, _RP, _p, _i32, C.c_double,
                                     _i64, _i32, _i64, _p, _p, _p, _p, _p, _p,
                                     _p, _p], C.c_int),
    'pd_rollout_host': ([_LP, _SP, _RP, _p, _i64, _i32, _i64, _p, _p, _p, _p,
                         _p, _p], C.c_int),
    'pd_env_step': ([_LP, _SP, _RP, C.POINTER(PdEnvConfig),
                     C.POINTER(PdEnvBuffers), _p, _p, _p, _p, _p, _p],
                    C.c_int),
    'pd_run_episodes': ([_LP, _SP, _RP, C.POINTER(PdEpisodeConfig), _p, _p, _p,
                         _p], C.c_int),
    'pd_mlp_apply_model': ([C.POINTER(PdMlp), _i32, _p, _i64, _p, _p], C.c_int),
    'pd_render_workspace_bytes': ([_i32, C.POINTER(_i64)], C.c_int),
    'pd_render_clusters': ([_i32, C.POINTER(_i32)], C.c_int),
    'pd_render_mask': ([_LP, _SP, _p, _i32, _i32, C.c_double, C.c_double, _p,
                        _p], C.c_int),
    'pd_sample_image_params': ([_SP, _p, _i32, _p], C.c_int),
    'pd_render': ([_LP, _SP, _p, _i32, _i32, _i32, _i32, C.c_double, _p, _p,
                   _i64, _p], C.c_int),
    'pd_get_atoms_in_bounds': ([_LP, _SP, _p, _i32, _p, _p, _p, _p, _p],
                               C.c_int),
    'pd_get_silicon_position': ([_LP, _SP, _p, _p], C.c_int),
    'pd_get_grid': ([_LP, _SP, _p, _i32, _p, _p], C.c_int),
    'pd_encode_observations': ([_LP, _SP, _p, _p, _i64, _i32, _p, C.c_float,
                                C.c_float, _i32, _p, _i64, _p, _p, _p, _p,
                                _p], C.c_int),
    'pd_observation_bytes': ([_i32, _i32], _i64),
    'pd_tfrecord_trajectories': ([_i32, _i64, _p, _p, _p, _p, _i64,
                                  C.POINTER(_i64)], C.c_int),
    'pd_crc32c': ([_p, _i64], C.c_uint32),
    'pd_generate_synthetic_data': ([C.c_uint64, _i32, _i64, _i32, _i32,
                                    C.c_float, C.c_float, _p, _p, _p, _p, _p,
                                    _p], C.c_int),
    'pd_generate_synthetic_data_network': (
        [C.c_uint64, _i32, _i64, _i32, _i32, _i32, C.c_float, C.c_float,
         _p, _p, _p, _p, _p, _p, _i32, _i32, _p, _p, _p, _p, _p, _p], C.c_int),
}

class PdFastAudit(C.Structure):
  """pd_fast_audit (include/pdune_b200.h)."""
  _fields_ = [('samples', C.c_int64), ('no_hop', C.c_int64),
              ('hop', C.c_int64), ('unsure', C.c_int64),
              ('wrong_decision', C.c_int64), ('wrong_slot', C.c_int64),
              ('waiting_time_outside_bounds', C.c_int64),
              ('total_rate_error_over_bound', C.c_double),
              ('waiting_time_error_over_bound', C.c_double),
              ('choice_error_over_bound', C.c_double),
              ('draw_error_over_bound', C.c_double)]


class PdRateOpsStats(C.Structure):
  """pd_rate_ops_stats (include/pdune_b200.h)."""
  _fields_ = [('evaluations', C.c_int64),
              ('simple_cast_differs_unguarded', C.c_int64),
              ('simple_cast_differs', C.c_int64),
              ('simple_guard_taken', C.c_int64),
              ('prior_cast_differs', C.c_int64),
              ('simple_max_ulps', C.c_uint32), ('prior_max_ulps', C.c_uint32),
              ('guard_ulps', C.c_uint32), ('reserved_', C.c_uint32)]


_SIGNATURES['pd_rate_ops_audit'] = (
    [_LP, C.c_uint64, _i64, C.c_double, C.POINTER(PdRateOpsStats), _p],
    C.c_int)
_SIGNATURES['pd_rollout_actions_host_packed'] = (
    [_LP, _SP, _RP, _p, _i32, C.c_double, _i64, _i32, _i64, _p, _p], C.c_int)
_SIGNATURES['pd_set_fast_path'] = ([C.c_int], C.c_int)
_SIGNATURES['pd_set_option'] = ([C.c_char_p, C.c_int], C.c_int)
_SIGNATURES['pd_fast_path_audit'] = (
    [_LP, _i32, C.c_uint64, _i64, _i64, C.c_double, C.POINTER(PdFastAudit),
     _p], C.c_int)

for _name, (_args, _res) in _SIGNATURES.items():
  _fn = getattr(lib, _name)
  _fn.argtypes = _args
  _fn.restype = _res


def bind_optional(name, args, res=C.c_int):
  """Binds an entry point declared later in the header (renderer, episodes)."""
  fn = getattr(lib, name)
  fn.argtypes = args
  fn.restype = res
  _SIGNATURES[name] = (args, res)
  return fn


def check(rc: int) -> None:
  if rc != PD_OK:
    msg = lib.pd_last_error()
    raise NativeError(
        f'libpdune_b200 error {rc}: {msg.decode() if msg else "?"}')


def exported_symbols():
  return sorted(_SIGNATURES)
