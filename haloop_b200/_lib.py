"""Loader (and in-tree builder) of libha_b200.so, the C-ABI CUDA library (include/ha_b200.h).

There is no CPU fallback: if the library is missing or was not built, importing the ops fails.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.environ.get("HA_B200_SO") or os.path.join(_HERE, "libha_b200.so")   # override: A/B builds
HEADER = os.path.join(ROOT, "include", "ha_b200.h")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "1886",
]


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + [HEADER]


def needs_build():
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a (one object each, in parallel) and link haloop_b200/libha_b200.so."""
    if not force and not needs_build():
        return SO_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = os.path.join(_HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    units = [f for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    deps = [os.path.getmtime(s) for s in _sources() if not s.endswith(".cu")]

    def compile_one(f):
        src, obj = os.path.join(CSRC, f), os.path.join(objdir, f[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max([os.path.getmtime(src)] + deps):
            return obj
        flags = [x for x in NVCC_FLAGS if x != "-shared"]
        subprocess.check_call([nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src])
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(units))) as ex:
        objs = list(ex.map(compile_one, units))
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO_PATH] + objs)
    return SO_PATH


_lib = None

_vp, _i64, _i32, _f32, _sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/ha_b200.h one to one
SIGNATURES = {
    "ha_b200_version": (_i32, []),
    "ha_b200_last_error": (ctypes.c_char_p, []),
    "ha_ctc_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "ha_ctc_fwd": (_i32, [_vp, _i64, _i64, _i32, _i32, _i32, _vp, _i64, _i32, _i32, _vp, _vp, _i32, _i32,
                          _vp, _vp, _sz, _vp]),
    "ha_ctc_bwd": (_i32, [_vp, _i64, _i64, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _i64, _i64, _vp, _sz, _vp]),
    "ha_star_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "ha_star_fwd": (_i32, [_vp, _i64, _i64, _i32, _i32, _i32, _vp, _i64, _i32, _i32, _vp, _vp, _i32, _f32, _i32,
                           _vp, _vp, _sz, _vp]),
    "ha_star_bwd": (_i32, [_vp, _i64, _i64, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _i64, _i64, _vp, _sz, _vp]),
    "ha_rnnt_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "ha_rnnt_fwd": (_i32, [_vp, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _vp, _i64, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _sz, _vp]),
    "ha_rnnt_bwd": (_i32, [_vp, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _i64, _i64, _i64, _vp, _sz, _vp]),
    "ha_rnnt_fg_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "ha_rnnt_fg_fwd": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _i64, _i32, _vp, _vp, _i32, _vp, _vp, _sz, _vp]),
    "ha_rnnt_fg_bwd": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ha_greedy_decode": (_i32, [_vp, _i64, _i64, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "ha_head_ctc_workspace_bytes": (_i32, [_i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "ha_head_ctc_fwd": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _i64, _i32, _i32, _vp, _vp, _i32, _i32,
                               _vp, _vp, _sz, _vp, _sz, _vp]),
    "ha_head_ctc_bwd": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _vp,
                               _vp, _sz, _vp, _sz, _vp]),
    "ha_ctc_beam_search_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "ha_ctc_beam_search": (_i32, [_vp, _i64, _i64, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ha_ctc_viterbi_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "ha_ctc_viterbi": (_i32, [_vp, _i64, _i64, _i32, _i32, _i32, _vp, _i64, _i32, _i32, _vp, _vp, _i32,
                              _vp, _vp, _vp, _sz, _vp]),
}


def lib():
    """The loaded library.  Raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). haloop_b200 has no CPU or PyTorch fallback.")
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class HaB200Error(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        msg = lib().ha_b200_last_error().decode("utf-8", "replace")
        raise HaB200Error(f"{what} failed (code {rc}): {msg}")
