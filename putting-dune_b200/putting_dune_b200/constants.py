"""Constants of the simulator hot path (reference: putting_dune/constants.py)."""

import numpy as np

CARBON = 6  # constants.py:20
SILICON = 14  # constants.py:21
CARBON_BOND_DISTANCE_ANGSTROMS = 1.42  # constants.py:23
SIGR_PRIOR_RATE_MEAN = np.array((0.85, 0))  # constants.py:26
SIGR_PRIOR_RATE_COV = np.array(((0.1, 0), (0, 0.1)))  # constants.py:27
SIGR_PRIOR_MAX_RATE = np.log(2) / 3  # constants.py:28
GAMMA_PER_SECOND = 0.9967  # constants.py:35
