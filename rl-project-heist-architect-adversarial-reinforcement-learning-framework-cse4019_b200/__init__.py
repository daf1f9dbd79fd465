"""heist_b200 -- B200-native batched Heist Architect environment hot path.

Importing the package is cheap and GPU-free; constructing an env or calling a kernel loads
lib/libheist_b200.so (built in-tree by build.build() / __graft_entry__.build()) and requires a
CUDA device.  There is no CPU fallback.
"""
from .batched_env import STATUS_NAMES, BatchedHeistEnv, EnvironmentConfig, guard_heading_table  # noqa: F401
from .compat import HeistEnvironment  # noqa: F401
from .rollout import RolloutBuffer, compute_gae, normalize_advantages  # noqa: F401
from . import dist  # noqa: F401
from . import build as _build_mod  # noqa: F401
from ._ffi import load as load_library, lib_path  # noqa: F401

build = _build_mod.build
__all__ = ["BatchedHeistEnv", "EnvironmentConfig", "HeistEnvironment", "RolloutBuffer", "compute_gae",
           "normalize_advantages", "STATUS_NAMES", "dist", "build", "load_library", "lib_path"]
from . import synthetic  # noqa: F401,E402
from . import ppo  # noqa: F401,E402
