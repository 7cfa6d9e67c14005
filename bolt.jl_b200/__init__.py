"""bolt.jl_b200 -- B200-native implementation of Bolt.jl's per-k Boltzmann hierarchy + LOS projection hot path.

Import as `bolt_b200` (the directory name is not a Python identifier; see bolt_b200/__init__.py).
"""
from .params import CosmoParams                               # noqa: F401
from .api import (BasicNewtonian, Hierarchy, boltsolve, boltsolve_rsa, source_grid, source_grid_P,   # noqa: F401
                  quadratic_k, log10_k, cltt, clte, clee, plin, spectra, default_context, device_cosmo)
