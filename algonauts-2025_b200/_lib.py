"""Loader / builder for ``libtribe_b200.so`` (the C ABI declared in ``include/tribe_b200.h``).

The library is built IN-TREE with nvcc for sm_100a (``csrc/libtribe_b200.so``) so that it travels with the repo
snapshot to the GPU box.  There is no CPU fallback: if the library cannot be loaded every op raises.
"""
from __future__ import annotations

import concurrent.futures
import ctypes
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB_PATH = os.path.join(CSRC, "libtribe_b200.so")
SOURCES = ["core.cu", "gemm_sm100.cu", "elementwise.cu", "reduce.cu", "contrastive.cu", "pool.cu", "optim.cu", "losses.cu", "aux.cu", "xgpu.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              f"-I{INCLUDE}", f"-I{CSRC}"]

_lib = None


class TribeError(RuntimeError):
    pass


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise TribeError("nvcc not found: cannot build libtribe_b200.so")


HASH_PATH = os.path.join(CSRC, "libtribe_b200.srchash")


def _source_hash() -> str:
    import hashlib

    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    deps.append(os.path.join(INCLUDE, "tribe_b200.h"))
    h = hashlib.sha1(" ".join(NVCC_FLAGS[:7]).encode())  # path-independent flags only: the GPU box mounts the repo elsewhere
    for d in deps:
        h.update(os.path.basename(d).encode())
        h.update(open(d, "rb").read())
    return h.hexdigest()


def _stale() -> bool:
    """True when the library is missing or was built from different sources (content hash, not mtimes: the snapshot
    copied to the GPU box does not preserve them)."""
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    return open(HASH_PATH).read().strip() != _source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a and link ``libtribe_b200.so`` (no GPU needed: nvcc cross-compiles)."""
    if not force and not _stale():
        return LIB_PATH
    import fcntl

    with open(os.path.join(CSRC, ".build.lock"), "w") as lock:  # ranks of one torchrun job must not build concurrently
        fcntl.flock(lock, fcntl.LOCK_EX)
        if force or _stale():
            _build_locked(verbose)
    return LIB_PATH


def _build_locked(verbose: bool) -> None:
    nvcc = _nvcc()
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise TribeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and (r.stdout or r.stderr):
            print(r.stdout, r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB_PATH + f".tmp{os.getpid()}"
    # the arch flag at the link step too: without it nvcc's device-link stub is an (empty) sm_52 cubin
    r = subprocess.run([nvcc, *NVCC_FLAGS[:2], "-shared", "-o", tmp, *objs, "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise TribeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    with open(HASH_PATH, "w") as f:
        f.write(_source_hash())


c_i32, c_i64, c_f32, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p
c_f64 = ctypes.c_double


class TribeOperand(ctypes.Structure):
    _fields_ = [("ptr", c_vp), ("inner", c_i64), ("rows", c_i64), ("batch", c_i64), ("row_stride", c_i64),
                ("batch_stride", c_i64), ("mn_major", c_i32), ("inner_off", c_i32), ("zin_stride", c_i32),
                ("zdiv", c_i32), ("gather", c_vp)]


class TribeGemm(ctypes.Structure):
    _fields_ = [("a", TribeOperand), ("b", TribeOperand), ("m", c_i32), ("n", c_i32), ("k", c_i32), ("batch", c_i32),
                ("z_inner", c_i32), ("kgroup", c_vp), ("kgroup_len", c_i32), ("d", c_vp), ("d_f32", c_i32),
                ("d_transposed", c_i32), ("ldd", c_i64), ("d_zo_stride", c_i64), ("d_zi_stride", c_i64),
                ("epilogue", c_i32), ("alpha", c_f32), ("bias", c_vp), ("bias_gathered", c_i32),
                ("bias_z_stride", c_i64), ("res", c_vp), ("ld_res", c_i64), ("res_row_mod", c_i32), ("res_batched", c_i32), ("rscale", c_vp),
                ("aux_in", c_vp), ("aux_out", c_vp), ("ld_aux", c_i64), ("rope", c_vp), ("rope_t", c_i32),
                ("rope_dim", c_i32), ("head_dim", c_i32), ("rope_cols", c_i32), ("rope_sign", c_f32),
                ("block_n", c_i32), ("splitk_ws", c_vp), ("splitk_ws_bytes", c_i64),
                ("adam_p", c_vp), ("adam_m", c_vp), ("adam_v", c_vp), ("adam_shadow", c_vp), ("adam_hyper", c_vp), ("adam_keep_grad", c_i32)]


XGPU_MAX_WORLD, XGPU_SLOTS = 16, 64


class TribeXgpuPeers(ctypes.Structure):
    _fields_ = [("ptr", c_vp * XGPU_MAX_WORLD)]


class TribeShardedAdam(ctypes.Structure):
    _fields_ = [("param", c_vp), ("m", c_vp), ("v", c_vp), ("hyper", c_vp), ("grad_mc", c_vp), ("shadow_mc", c_vp), ("param_mc", c_vp),
                ("grad_peer", TribeXgpuPeers), ("shadow_peer", TribeXgpuPeers), ("param_peer", TribeXgpuPeers), ("n", c_i64),
                ("world", c_i32), ("rank", c_i32), ("bcast_master", c_i32), ("max_blocks", c_i32)]


# name -> (argtypes); every function returns int except the three introspection calls.
_SIGS = {
    "tribe_gemm_bf16": [ctypes.POINTER(TribeGemm), c_vp],
    "tribe_gemm_bf16_probe": [ctypes.POINTER(TribeGemm), c_vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32],
    "tribe_gemm_set_sm_limit": [c_i32],
    "tribe_set_pdl": [c_i32],
    "tribe_peek_last_error": [],
    "tribe_take_last_error": [],
    "tribe_attn_scores": [c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_f32, c_i32, c_vp, c_vp, c_i64, c_vp],
    "tribe_attn_fwd": [c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_f32, c_vp, c_i64, c_vp, c_i64, c_i64, c_vp],
    "tribe_ingest_features": [c_vp, c_i32, c_i64, c_i64, c_i64, c_i64, c_i32, c_vp, c_i64, c_i64, c_vp],
    "tribe_scalenorm_fwd": [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, c_f32, c_vp],
    "tribe_sublayer_bwd": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, c_vp],
    "tribe_rope_half": [c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64, c_f32, c_vp],
    "tribe_softmax_fwd": [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp],
    "tribe_softmax_bwd": [c_vp, c_vp, c_vp, c_f32, c_i64, c_i64, c_i64, c_vp],
    "tribe_colsum": [c_vp, c_i32, c_vp, c_i32, c_vp, c_i64, c_i64, c_i64, c_i32, c_vp],
    "tribe_cast_f32_bf16": [c_vp, c_vp, c_i64, c_vp],
    "tribe_axpby_f32": [c_vp, c_vp, c_f32, c_i32, c_i64, c_vp],
    "tribe_adaptive_avg_pool_fwd": [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp],
    "tribe_adaptive_avg_pool_bwd": [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp],
    "tribe_token_pool_fwd": [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp],
    "tribe_token_pool_bwd": [c_vp, c_i32, c_vp, c_i32, c_i64, c_i64, c_i64, c_i64, c_vp],
    "tribe_add_rows_periodic": [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp],
    "tribe_transpose_cast_bot": [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp],
    "tribe_subject_bias_grad": [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp],
    "tribe_check_subjects": [c_vp, c_i64, c_i64, c_vp, c_vp, c_vp],
    "tribe_mse_fwd_bwd": [c_vp, c_vp, c_vp, c_vp, c_f32, c_i64, c_vp, c_vp],
    "tribe_pearson_stats": [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp],
    "tribe_pearson_pick_shift": [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp],
    "tribe_pearson_recenter": [c_vp, c_i64, c_i64, c_vp, c_vp, c_vp],
    "tribe_pearson_finalize": [c_vp, c_i64, c_vp, c_vp, c_vp],
    "tribe_nce_expsums": [c_vp, c_i64, c_i64, c_f32, c_vp, c_vp, c_vp],
    "tribe_nce_loss": [c_vp, c_i64, c_i64, c_f32, c_vp, c_vp, c_vp, c_vp],
    "tribe_nce_grad": [c_vp, c_i64, c_i64, c_f32, c_vp, c_vp, c_vp, c_f32, c_vp, c_i64, c_vp],
    "tribe_cast_bf16_f32": [c_vp, c_vp, c_i64, c_vp],
    "tribe_adam_step": [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_f64, c_f64, c_f64, c_f64, c_f64, c_i64, c_i32, c_vp],
    "tribe_adam_step_dev": [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i32, c_vp],
    "tribe_adam_hyper": [c_vp, c_f64, c_f64, c_f64, c_f64, c_f64, c_i64, c_vp],
    "tribe_adam_hyper_batch": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp],
    "tribe_sharded_adam_step": [ctypes.POINTER(TribeShardedAdam), c_vp],
    "tribe_memcpy_async": [c_vp, c_vp, c_i64, c_vp],
    "tribe_debug_spin": [c_i32, c_i32, c_f64, c_i32, c_vp, c_vp],
    "tribe_xgpu_probe": [ctypes.POINTER(TribeShardedAdam), c_i32, c_i32, c_vp, c_vp],
    "tribe_xgpu_barrier": [ctypes.POINTER(TribeXgpuPeers), c_i32, c_i32, c_i32, c_vp, c_f64, c_vp],
    "tribe_point_loss_fwd_bwd": [c_vp, c_vp, c_vp, c_vp, c_i32, c_f32, c_f32, c_i64, c_vp, c_vp],
    "tribe_pearson_loss_finalize": [c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp],
    "tribe_pearson_loss_bwd": [c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_i64, c_i64, c_i64, c_vp],
    "tribe_gather_windows": [c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp],
    "tribe_ensemble_weights": [c_vp, c_i64, c_i64, c_f32, c_i32, c_vp, c_vp],
    "tribe_ensemble_average": [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp],
    "tribe_mean_lastdim": [c_vp, c_vp, c_i64, c_i64, c_vp],
    "tribe_retrieval_ranks": [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp],
    "tribe_swa_update": [c_vp, c_vp, c_i64, c_i64, c_vp],
    "tribe_scale_dev": [c_vp, c_vp, c_vp, c_i64, c_vp],
    "tribe_memset_zero": [c_vp, c_i64, c_vp],
    "tribe_transpose_last2": [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp],
}
EXPORTS = sorted(list(_SIGS) + ["tribe_last_error", "tribe_abi_version", "tribe_launch_count"])


def load():
    """Return the ctypes handle; builds the library first if sources are newer.  Raises if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    override = os.environ.get("TRIBE_LIB_OVERRIDE")  # A/B measurements against a library built from another commit
    if override:
        path = override
    else:
        if _stale():
            build()
        path = LIB_PATH
    try:
        lib = ctypes.CDLL(path)
    except OSError as e:  # pragma: no cover
        raise TribeError(f"cannot load {LIB_PATH}: {e} — the CUDA extension is required (no CPU fallback)") from e
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = args, ctypes.c_int
    lib.tribe_last_error.restype = ctypes.c_char_p
    lib.tribe_abi_version.restype = ctypes.c_int
    lib.tribe_launch_count.restype = ctypes.c_int64
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().tribe_last_error().decode(errors="replace")
        raise TribeError(f"{what} failed (code {rc}): {msg}")


REPLAYED_LAUNCHES = 0  # kernels launched by CUDA-graph replays (graphed.py): they bypass the C entry points' counter


def launch_count() -> int:
    """Kernels of this library launched so far by this process: direct launches + launches inside replayed graphs."""
    return int(load().tribe_launch_count()) + REPLAYED_LAUNCHES
