"""Build + ctypes binding of the C-ABI library (include/rlsolver_b200.h).

The shared object is built IN-TREE (rlsolver_b200/_C/librlsolver_b200.so) with nvcc for
sm_100a only.  There is no CPU fallback: if the library is missing or a call fails, the
Python layer raises.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import threading
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OUT_DIR = os.path.join(_HERE, "_C")
SO_PATH = os.path.join(OUT_DIR, "librlsolver_b200.so")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]

_lock = threading.Lock()
_lib = None


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; rlsolver_b200 needs the CUDA toolkit to build its sm_100a kernels")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "rlsolver_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu for sm_100a and link the shared library.  Returns its path."""
    with _lock:
        if not force and not _stale():
            return SO_PATH
        nvcc = _nvcc()
        os.makedirs(OUT_DIR, exist_ok=True)
        objs = []

        def compile_one(src):
            obj = os.path.join(OUT_DIR, os.path.basename(src)[:-3] + ".o")
            cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                print(r.stderr)
            return obj

        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
            objs = list(ex.map(compile_one, sources()))
        tmp = SO_PATH + ".tmp"
        r = subprocess.run([nvcc, "-shared", "-o", tmp, *objs, "-cudart", "static"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        os.replace(tmp, SO_PATH)
        return SO_PATH


# ----------------------------------------------------------------------------- signatures
_i32, _i64, _f32, _vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p
_u32, _u64 = C.c_uint32, C.c_uint64

SIGNATURES = {
    "rlsb_version": (C.c_int, []),
    "rlsb_last_error": (C.c_char_p, []),
    "rlsb_debug_flags": (_i32, [_i32, _i32]),
    "rlsb_graph_create": (C.c_int, [_i32, _i64, _vp, _vp, _vp, _i32, _i32, C.POINTER(_vp)]),
    "rlsb_graph_destroy": (C.c_int, [_vp]),
    "rlsb_graph_num_nodes": (_i32, [_vp]),
    "rlsb_graph_padded_nodes": (_i32, [_vp]),
    "rlsb_graph_num_edges": (_i64, [_vp]),
    "rlsb_graph_num_listed": (_i64, [_vp]),
    "rlsb_graph_num_full": (_i64, [_vp]),
    "rlsb_graph_num_levels": (_i32, [_vp]),
    "rlsb_graph_max_listed_degree": (_i32, [_vp]),
    "rlsb_graph_max_full_degree": (_i32, [_vp]),
    "rlsb_graph_export": (C.c_int, [_vp] * 7),
    "rlsb_graph_sell_sizes": (C.c_int, [_vp, _i32, _vp]),
    "rlsb_graph_sell_export": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "rlsb_pack_spins": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp]),
    "rlsb_unpack_spins": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp]),
    "rlsb_cut_eval": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "rlsb_cut_eval_packed": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "rlsb_graph_is_weighted": (C.c_int, [_vp]),
    "rlsb_cut_eval_weighted": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "rlsb_node_fields_weighted": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "rlsb_cut_edges": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "rlsb_node_cross_counts": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "rlsb_ls_workspace_bytes": (_i64, [_vp, _i64]),
    "rlsb_ls_workspace_offset": (_i64, [_vp, _i64, _i32]),
    "rlsb_ls_begin": (C.c_int, [_vp, _vp, _i64, _vp, _i32, _i32, _f32, _vp, _vp]),
    "rlsb_ls_thresh": (C.c_int, [_vp, _i64, _i32, _vp, _i32, _vp, _vp]),
    "rlsb_ls_search": (C.c_int, [_vp, _i64, _vp, _i32, _vp, _i32, _i32, _vp, _vp, _vp]),
    "rlsb_ls_debug_times": (C.c_int, [_vp]),
    "rlsb_ls_run": (C.c_int, [_vp, _i64, _vp, _i32, _vp, _i32, _vp, _i32, _i32, _vp, _vp, _vp]),
    "rlsb_ls_mask_words": (_i64, [_vp, _i64]),
    "rlsb_ls_noise_masks": (C.c_int, [_vp, _i64, _i32, _u64, _u64, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "rlsb_rng_cursor_advance": (C.c_int, [_vp, _u64, _vp]),
    "rlsb_ls_run_masks": (C.c_int, [_vp, _i64, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "rlsb_ls_fused_search": (C.c_int, [_vp, _i64, _vp, _i32, _u64, _u64, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "rlsb_ls_fused_status": (C.c_int, [_vp, _i64, _vp, _vp, _vp]),
    "rlsb_ls_begin_packed": (C.c_int, [_vp, _vp, _i64, _vp, _i32, _i32, _f32, _vp, _vp]),
    "rlsb_torch_randn": (C.c_int, [_vp, _i64, _u64, _u64, _vp, _i32, _i32, _i32, _vp]),
    "rlsb_flip_sweep": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "rlsb_relaxed_cut": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "rlsb_relaxed_cut_grad": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "rlsb_step_flip": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "rlsb_greedy_best_flip": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i32, _i32, _vp]),
    "rlsb_torch_rand": (C.c_int, [_u64, _u64, _u32, _u32, _i64, _i64, _vp, _vp]),
    "rlsb_torch_randint": (C.c_int, [_u64, _u64, _u32, _u32, _i64, _i64, _u32, _vp, _vp]),
    "rlsb_mcpg_plan_create": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "rlsb_mcpg_plan_destroy": (C.c_int, [_vp]),
    "rlsb_mcpg_plan_num_levels": (_i32, [_vp]),
    "rlsb_mcpg_sweeps": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _u64, _u64, _u32, _u32, _vp, _vp]),
    "rlsb_mcpg_weighted_sweeps": (C.c_int, [_i32, _i64, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp,
                                            _u64, _u64, _u32, _u32, _vp, _vp]),
    "rlsb_metro_sampling": (C.c_int, [_i32, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _u64, _u64, _u32, _u32, _vp,
                                      _i32, _vp]),
    "rlsb_metro_workspace_bytes": (_i64, [_i32, _i64, _i32]),
    "rlsb_metro_sampling_split": (C.c_int, [_i32, _vp, _vp, _vp, _i64, _i32, _i64, _u64, _u64, _u32, _u32, _vp, _vp,
                                            _vp]),
    "rlsb_subset_sampling": (C.c_int, [_vp, _i64, _i32, _i64, _i32, _vp, _vp, _vp, _u64, _u64, _u32, _u32, _vp]),
    "rlsb_qubo_create": (C.c_int, [_vp, _i32, _i32, C.POINTER(_vp), _vp]),
    "rlsb_qubo_destroy": (C.c_int, [_vp]),
    "rlsb_qubo_padded_vars": (_i32, [_vp]),
    "rlsb_qubo_workspace_bytes": (_i64, [_vp, _i64]),
    "rlsb_qubo_energy": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "rlsb_qubo_sweeps": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "rlsb_peco_fields": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "rlsb_peco_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _i32, _i32, _vp,
                                 _i32, _i32, _f32, _f32, _i32, _f32, _i32, _f32, _i32, _vp]),
    "rlsb_peco_compact_step": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp,
                                         _vp, _i64, _i32, _i32, _i32, _i32, _i32, _f32, _i32, _f32, _i32, _vp, _i32, _i32,
                                         _vp]),
    "rlsb_peco_compact_fields": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp]),
    "rlsb_peco_compact_from_dense": (C.c_int, [_vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "rlsb_peco_compact_expand_matrix": (C.c_int, [_vp, _vp, _i64, _i64, _i32, _vp, _i64, _vp]),
    "rlsb_peco_compact_expand_state": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp,
                                                 _i32, _i32, _f32, _i32, _i32, _vp]),
    "rlsb_peco_compact_expand_state_half": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32,
                                                      _vp, _i32, _i32, _f32, _i32, _i32, _vp]),
    "rlsb_peco_gen_er": (C.c_int, [_vp, _i64, _i32, _f32, _u64, _u64, _u32, _u32, _vp]),
    "rlsb_peco_gen_ba": (C.c_int, [_vp, _i64, _i32, _i32, _u64, _u64, _u32, _u32, _vp]),
    "rlsb_peco_gen_pl": (C.c_int, [_vp, _i64, _i32, _i32, _f32, _u64, _u64, _vp]),
    "rlsb_isco_propose": (C.c_int, [_vp, _vp, _i32, _vp, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp,
                                    _i32, _i32, _i64, _vp]),
    "rlsb_isco_accept": (C.c_int, [_vp, _vp, _i32, _vp, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp,
                                   _vp, _i32, _i32, _i64, _vp]),
    "rlsb_select_rows": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp]),
    "rlsb_best_record": (C.c_int, [_vp, _vp, _i64, _i32, _i64, _vp, _vp]),
    "rlsb_best_record_packed": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i64, _vp, _vp]),
    "rlsb_best_pick_strided": (C.c_int, [_vp, _i32, _i32, _i64, _vp, _vp, _vp]),
    "rlsb_copy_rows": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _vp, _vp]),
    "rlsb_mh_rows_round": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i32, _i64, _u64, _u64, _u32, _u32, _vp, _vp, _vp]),
    "rlsb_peer_exchange_handle_bytes": (_i64, []),
    "rlsb_peer_exchange_create": (C.c_int, [_i32, _i32, _i32, C.POINTER(_vp), _vp]),
    "rlsb_peer_exchange_connect": (C.c_int, [_vp, _vp]),
    "rlsb_peer_exchange_best": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "rlsb_peer_exchange_best_packed": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i64, _vp, _vp, _vp]),
    "rlsb_peer_exchange_status": (C.c_int, [_vp, _vp, _vp]),
    "rlsb_peer_exchange_destroy": (C.c_int, [_vp]),
    "rlsb_best_pick": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp]),
    "rlsb_pick_best": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _i32, _vp, _vp, _vp]),
}


DEBUG_PLAIN_MASKS, DEBUG_FULL_CUT, DEBUG_LS_SKIP, DEBUG_LS_TIMES = 1, 2, 4, 8
DEBUG_CARVEOUT_DEFAULT, DEBUG_GEN_PER_DRAW, DEBUG_PECO_WARP_PER_ENV, DEBUG_QUBO_NO_SPLITK = 16, 32, 64, 128
DEBUG_THRESH_PIPE = 256


def debug_flags(set_mask: int = 0, clear_mask: int = 0) -> int:
    """rlsb_debug_flags: diagnostic switches of the library (tests / profiling tools)."""
    return int(lib().rlsb_debug_flags(int(set_mask), int(clear_mask)))


class RlsbError(RuntimeError):
    pass


_EXC = {1: ValueError, 2: RlsbError, 3: NotImplementedError, 4: RlsbError}


def lib():
    """The loaded C-ABI library (ctypes.CDLL).  Builds it if the in-tree .so is missing."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(SO_PATH):
                    raise ImportError(
                        f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(rlsolver_b200 has no CPU / PyTorch fallback)")
                handle = C.CDLL(SO_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)      # AttributeError if the ABI and the header drift
                    fn.restype, fn.argtypes = res, args
                _lib = handle
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().rlsb_last_error().decode("utf-8", "replace")
        raise _EXC.get(status, RlsbError)(f"{what or 'rlsolver_b200'}: {msg} (status {status})")
