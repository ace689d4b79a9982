"""apyib_b200 -- B200-native (sm_100a) drop-in for apyib's correlated-wavefunction hot path.

Mirrors the reference's public surface for that path:
    mp2_wfn, ci_wfn, AAT, finite_difference, energy, phase_corrected_energy, compute_parallel_aats
Importing the package loads libapyib_b200.so; there is no CPU fallback.
"""
from . import config                                    # noqa: F401
from ._lib import lib, ApyibB200Error                   # noqa: F401
from .mp2_wfn import mp2_wfn                            # noqa: F401
from .ci_wfn import ci_wfn                              # noqa: F401
from . import utils                                     # noqa: F401

if config.PAIRS_SINGLE_VECTOR:                          # experimental kernel variant (config.py)
    lib.apyib_det_set_pairs_variant(1)

__version__ = "0.1.0"
