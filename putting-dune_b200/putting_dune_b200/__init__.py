"""putting_dune_b200: B200-native batched simulator for the putting-dune hot
path (one KMC step of Si-doped graphene under beam control + STEM frame)."""

from putting_dune_b200 import _native  # fails loudly if the .so is missing
from putting_dune_b200 import constants
from putting_dune_b200 import geometry
from putting_dune_b200 import microscope_utils
from putting_dune_b200 import engine
from putting_dune_b200 import graphene
from putting_dune_b200 import imaging
from putting_dune_b200 import io
from putting_dune_b200 import episodes
from putting_dune_b200 import putting_dune_environment
from putting_dune_b200 import simulator
from putting_dune_b200 import simulator_observers
from putting_dune_b200.engine import EnvBatch, Lattice, MlpWeights, RateSpec
from putting_dune_b200.putting_dune_environment import (
    BatchedPuttingDuneEnvironment)
from putting_dune_b200.simulator import BatchedSimulator, PuttingDuneSimulator

__all__ = ['BatchedPuttingDuneEnvironment', 'putting_dune_environment',
           'EnvBatch', 'Lattice', 'MlpWeights', 'RateSpec', 'BatchedSimulator',
           'PuttingDuneSimulator', 'constants', 'geometry', 'microscope_utils',
           'engine', 'episodes', 'graphene', 'imaging', 'simulator', 'simulator_observers']
