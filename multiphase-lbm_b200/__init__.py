"""multiphase-lbm_b200: host-side mirror of the CooLBM multiphase functor surface over the clbm C ABI.

The directory name carries a hyphen (it is the project name), so import it through
`__graft_entry__.load_package()` / `tests/_cases.load_package()`, which register it as
`multiphase_lbm_b200`.
"""
from . import params  # noqa: F401
from .params import *  # noqa: F401,F403
from . import clbm, slab, pulsatile_cases  # noqa: F401,E402
