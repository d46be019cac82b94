"""xrsdkit_b200 -- B200-native scattering-intensity engine behind xrsdkit's System / Population API.

    from xrsdkit_b200.system import System, Population, fit
    from xrsdkit_b200 import batch          # many parameter sets per launch

All numerics run in hand-written sm_100a CUDA kernels (csrc/) reached through the C ABI of
include/xrsd_b200.h; importing the package does not need a GPU, computing does."""
__version__ = "0.1.0"
