"""ctypes binding of include/heist_b200.h.  There is no CPU fallback: if the CUDA library is
missing this module raises, and every call checks the int return code."""
import ctypes as C
import os

from . import build as _build

_lib = None

c_vp = C.c_void_p


class HeistParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "grid_rows", "grid_cols", "max_steps", "start_row", "start_col", "vault_row", "vault_col",
        "architect_budget", "max_walls", "max_cams", "max_guards", "max_path")] + \
        [("reward_vault", C.c_double), ("reward_detection", C.c_double), ("reward_step", C.c_double)]


class HeistLayoutArrays(C.Structure):
    _fields_ = [(n, c_vp) for n in (
        "n_walls", "wall_rc", "n_cams", "cam_rc", "cam_f", "cam_range", "n_guards", "guard_len",
        "guard_path", "guard_head", "guard_speed", "guard_range", "guard_fov")]


class HeistStateView(C.Structure):
    _fields_ = [(n, c_vp) for n in (
        "tile", "wall_bits", "vis_bits", "env_static", "env_dyn", "cam_f", "cam_i", "cam_heading",
        "guard_fov", "guard_i", "guard_path", "guard_heading", "guard_idx", "wall_accepted")]


EXPORTS = {
    "heist_abi_version": (C.c_int, []),
    "heist_last_error": (C.c_char_p, []),
    "heist_last_warning": (C.c_char_p, []),
    "heist_debug_ray_dirs": (C.c_int, [C.c_int, c_vp, C.c_int, c_vp, c_vp, c_vp]),
    "heist_create": (C.c_int, [C.POINTER(HeistParams), C.c_int, C.c_int, C.POINTER(c_vp)]),
    "heist_destroy": (C.c_int, [c_vp]),
    "heist_decode_validate": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int, c_vp, c_vp]),
    "heist_set_layout_explicit": (C.c_int, [c_vp, C.POINTER(HeistLayoutArrays), c_vp, c_vp, c_vp]),
    "heist_reset": (C.c_int, [c_vp, c_vp, c_vp]),
    "heist_step": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "heist_step_many": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "heist_step_many_host": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "heist_observe": (C.c_int, [c_vp, c_vp, c_vp]),
    "heist_observation_vectors": (C.c_int, [c_vp, c_vp, c_vp]),
    "heist_step_observe": (C.c_int, [c_vp, c_vp, C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "heist_expand_states": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_int, c_vp, c_vp]),
    "heist_get_state": (C.c_int, [c_vp, C.POINTER(HeistStateView)]),
    "heist_gae": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_double, C.c_double, c_vp, c_vp, C.c_int, c_vp]),
    "heist_architect_reward": (C.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "heist_check_errors": (C.c_int, [c_vp, c_vp]),
    "heist_set_mode": (C.c_int, [c_vp, C.c_int]),
    "heist_cache_stats": (C.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "heist_launch_count": (C.c_int, [c_vp, c_vp]),
}


def lib_path():
    """HEIST_B200_DEBUG=1 selects the range-checked debug build of the same sources."""
    return _build.DBG_LIB_PATH if os.environ.get("HEIST_B200_DEBUG") == "1" else _build.LIB_PATH


def load():
    """dlopen lib/libheist_b200.so and declare every symbol of include/heist_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: the CUDA library has not been built. Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` (needs nvcc). There is no CPU fallback for the environment hot path.")
    L = C.CDLL(path)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(L, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if L.heist_abi_version() != 2:
        raise RuntimeError("libheist_b200.so ABI version mismatch")
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        msg = load().heist_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
