"""phantom_b200 -- B200-native SPH derivative engine behind Phantom's hot-path interfaces.

Only what the hot path needs lives here: `csrc/` (CUDA kernels + the C ABI of
include/sphgpu.h) and the host-side mirror of the reference's
build_tree / densityiterate / cons2prim_everything / force interface (`api.py`).
"""
from .params import SphParams, SphScalars, default_params  # noqa: F401
