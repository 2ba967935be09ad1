"""Times the UNMODIFIED reference simulator (BASELINE.json configs[0]).

TEST / MEASUREMENT INFRASTRUCTURE ONLY: used by the CPU-baseline leg of
``bench.py`` and by ``bench.py --impl reference``; nothing here is on the
product path.

What runs is the reference's own code, nothing of this repo's: its
``PuttingDuneSimulator`` (putting_dune/simulator.py:27-182) over its
``PristineSingleDopedGraphene`` (graphene.py:562-694) with its
``HumanPriorRatePredictor().predict`` / ``simple_canonical_rate_function``
(graphene.py:133-229), its sklearn KD-tree neighbour search
(geometry.py:93-111), ``np.random.default_rng(seed)`` as the generator and its
own ``RelativeToSiliconActionAdapter`` (action_adapters.py:131-216) turning
agent actions ~ U(-1, 1)^2 into beam controls with a 1.5 s dwell time -- the
``relative_random`` experiment of experiments/registry.py:263-266, i.e.
BASELINE.md section 4's CPU-baseline plan.  ``oracle/refshim.py`` only
supplies the third-party names that are absent from this image (shapely
Point, jax.scipy.stats -> scipy.stats, dm_env value types); the reference
sources are read from ``baseline/_ref`` (a plain copy of the reference's
package made by ``__graft_entry__.build()``; git-ignored) or /root/reference.
"""

from __future__ import annotations

import datetime as dt
import os
import time

import numpy as np


def reference_available() -> bool:
  from oracle import refshim
  return refshim.reference_available()


def _make_sim(rate: str):
  from oracle import refshim
  mods = refshim.load_reference_env_stack()
  g = mods.graphene
  fn = (g.HumanPriorRatePredictor().predict if rate == 'prior'
        else g.simple_canonical_rate_function)
  material = g.PristineSingleDopedGraphene(
      rate_function=g.PristineSingleSiGrRatePredictor(
          canonical_rate_prediction_fn=fn))
  sim = mods.simulator.PuttingDuneSimulator(material)
  adapter = mods.action_adapters.RelativeToSiliconActionAdapter(
      dwell_time_range=(dt.timedelta(seconds=1.5), dt.timedelta(seconds=1.5)),
      max_distance_angstroms=1.42)
  return mods, sim, adapter


def run_steps(n_steps: int, seed: int = 0, rate: str = 'prior',
              warmup: int = 1):
  """(seconds, env-steps, transitions) of `n_steps` step_and_image calls of
  one reference simulator after `warmup` untimed ones."""
  mods, sim, adapter = _make_sim(rate)
  rng = np.random.default_rng(seed)
  obs = sim.reset(rng)
  adapter.reset()
  transitions = [0]

  class Count(mods.microscope_utils.SimulatorObserver):

    def observe_transition(self, time_since_control_was_applied, grid):
      transitions[0] += 1

  sim.add_observer(Count())
  t0 = None
  with np.errstate(divide='ignore', over='ignore', invalid='ignore'):
    for i in range(warmup + n_steps):
      if i == warmup:
        transitions[0] = 0
        t0 = time.perf_counter()
      action = rng.uniform(-1.0, 1.0, size=2).astype(np.float32)
      controls = adapter.get_action(obs, action)
      obs = sim.step_and_image(rng, controls)
  return time.perf_counter() - t0, n_steps, transitions[0]


def _worker(args):
  n_steps, seed, rate = args
  return run_steps(n_steps, seed, rate)


def throughput(n_steps: int, cores: int, rate: str = 'prior'):
  """env-steps/s of `cores` independent reference simulators (one process
  each, one env each), `n_steps` steps per simulator.  Returns (value,
  per-process busy seconds, transitions per step, sample text)."""
  import multiprocessing as mp
  jobs = [(n_steps, i, rate) for i in range(cores)]
  t0 = time.perf_counter()
  if cores == 1:
    res = [_worker(jobs[0])]
  else:
    with mp.get_context('fork').Pool(cores) as pool:
      res = pool.map(_worker, jobs)
  wall = time.perf_counter() - t0
  busy = max(r[0] for r in res)
  total = sum(r[1] for r in res)
  hops = sum(r[2] for r in res) / max(total, 1)
  sample = (f'{cores} unmodified PuttingDuneSimulator(s) (one env, one '
            f'process each) x {n_steps} step_and_image calls, {rate} rates, '
            f'relative_random controls, dwell 1.5 s; {busy:.1f} s busy / '
            f'{wall:.1f} s wall; {hops:.3f} transitions per step')
  return total / busy, busy, hops, sample


if __name__ == '__main__':
  import sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  sys.path.insert(0, root)
  n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
  for c in (1, os.cpu_count() or 1):
    v, busy, hops, sample = throughput(n, c)
    print(f'{v:.1f} env-steps/s on {c} core(s): {sample}')
