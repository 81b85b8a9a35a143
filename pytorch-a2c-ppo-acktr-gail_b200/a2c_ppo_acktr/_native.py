"""ctypes binding of liba2cb.so (C ABI declared in include/a2cb.h).

There is deliberately no CPU or eager-PyTorch fallback: if the shared library is
missing, or a tensor handed to a kernel wrapper does not live on a CUDA device,
the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "liba2cb.so")

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int32
c_i64 = ctypes.c_int64
c_f64 = ctypes.c_double
c_f32 = ctypes.c_float


class NativeError(RuntimeError):
    pass


class GatherField(ctypes.Structure):
    _fields_ = [("src", c_void_p), ("dst", c_void_p), ("row_bytes", c_i64)]


class ColCopyField(ctypes.Structure):
    """Mirror of a2cb_colcopy_t."""
    _fields_ = [("dst", c_void_p), ("src", c_void_p), ("dst_cs", c_i64), ("dst_rs", c_i64), ("src_cs", c_i64),
                ("src_rs", c_i64), ("w", c_i64), ("rows", c_int), ("reserved_", c_int)]


_SIGNATURES = {
    "a2cb_abi_version": (c_int, []),
    "a2cb_copy_columns": (c_int, [ctypes.POINTER(ColCopyField), c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "a2cb_last_error": (ctypes.c_char_p, []),
    "a2cb_launch_count": (c_i64, []),
    "a2cb_add_launches": (None, [c_i64]),
    "a2cb_set_gemm_mode": (c_int, [c_int]),
    "a2cb_get_gemm_mode": (c_int, []),
    "a2cb_returns_scan": (c_int, [c_void_p] * 6 + [c_int, c_i64, c_f64, c_f64, c_int, c_int, c_int, c_void_p]),
    "a2cb_gather_rows": (c_int, [ctypes.POINTER(GatherField), c_int, c_void_p, c_i64, c_int, c_i64, c_i64, c_void_p]),
    "a2cb_upload_env_columns": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_void_p, c_int, c_void_p]),
    "a2cb_gru_seq_variant": (c_int, [c_int, c_int]),
    "a2cb_set_gru_mode": (c_int, [c_int]),
    "a2cb_gru_tc_debug_stamps": (c_int, [c_void_p]),
    "a2cb_gru_cluster_probe": (c_int, [c_int, c_int]),
}

_lib = None
_applied = set()


def register(signatures):
    """Adds ctypes signatures (used by _engine for the descriptor-taking entry points)."""
    _SIGNATURES.update(signatures)
    if _lib is not None:
        _apply_signatures(_lib)


def _apply_signatures(l):
    for name, (res, args) in _SIGNATURES.items():
        if name in _applied:
            continue
        fn = getattr(l, name)
        fn.restype = res
        fn.argtypes = args
        _applied.add(name)


def lib():
    """Loads liba2cb.so once; raises if it has not been built (python build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                "liba2cb.so not found at %s -- build it with `python __graft_entry__.py` or "
                "`python pytorch-a2c-ppo-acktr-gail_b200/build.py`; there is no CPU fallback" % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        _apply_signatures(l)
        _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        raise NativeError("%s failed (%d): %s" % (what, rc, lib().a2cb_last_error().decode()))


def set_gemm_mode(mode):
    """0: exact-fp32 SIMT engine; 1: tcgen05 3xTF32 engine everywhere; 2: automatic choice (default)."""
    check(lib().a2cb_set_gemm_mode(int(mode)), "a2cb_set_gemm_mode")


def set_gru_mode(mode):
    """Kernel family of the persistent GRU recurrence: 2 tcgen05 (default), 1 mma.sync 3xTF32, 0 fp32 SIMT, -1 default."""
    check(lib().a2cb_set_gru_mode(int(mode)), "a2cb_set_gru_mode")


def gru_seq_variant(N=64, H=512):
    """Kernel pair behind a2cb_gru_seq for this shape: "simt" | "mma" | "tc" (None: unsupported)."""
    return {0: "simt", 1: "mma", 2: "tc"}.get(int(lib().a2cb_gru_seq_variant(int(N), int(H))))


def launch_count():
    return int(lib().a2cb_launch_count())


def stream_ptr(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def dptr(t, dtype=None, name="tensor"):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return c_void_p(0)
    if not t.is_cuda:
        raise NativeError("%s must be a CUDA tensor (got %s); the update path has no CPU fallback"
                          % (name, t.device))
    if dtype is not None and t.dtype != dtype:
        raise NativeError("%s must be %s (got %s)" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise NativeError("%s must be contiguous" % name)
    return c_void_p(t.data_ptr())


# ---------------------------------------------------------------------------
# thin wrappers
# ---------------------------------------------------------------------------
def returns_scan(rewards, value_preds, returns, masks, bad_masks, next_value, gamma, gae_lambda,
                 use_gae, use_proper_time_limits, schedule=0):
    T, N = rewards.shape[0], rewards.shape[1]
    f32 = torch.float32
    ptrs = [dptr(rewards, f32, "rewards"), dptr(value_preds, f32, "value_preds"),
            dptr(returns, f32, "returns"), dptr(masks, f32, "masks"),
            dptr(bad_masks, f32, "bad_masks"), dptr(next_value, f32, "next_value")]
    with torch.cuda.device(rewards.device):
        rc = lib().a2cb_returns_scan(
            *ptrs, T, N, float(gamma), float(gae_lambda), int(bool(use_gae)),
            int(bool(use_proper_time_limits)), int(schedule), stream_ptr(rewards.device))
    check(rc, "a2cb_returns_scan")


def gather_rows(pairs, idx, n_out_rows, mode=0, E=0, N=0):
    """pairs: list of (src, dst) tensors; rows are dim 0 of the flattened [rows, ...] views."""
    arr = (GatherField * len(pairs))()
    for i, (src, dst) in enumerate(pairs):
        if src.dtype != dst.dtype:
            raise NativeError("gather_rows: dtype mismatch")
        row_bytes = dst[0].numel() * dst.element_size() if dst.shape[0] else 0
        if dst.shape[0] == 0:
            row_bytes = src[0].numel() * src.element_size()
        arr[i].src = dptr(src, name="gather src").value
        arr[i].dst = dptr(dst, name="gather dst").value
        arr[i].row_bytes = row_bytes
    dev = pairs[0][0].device
    idx_p = dptr(idx, torch.int64, "idx")
    with torch.cuda.device(dev):
        rc = lib().a2cb_gather_rows(arr, len(pairs), idx_p, int(n_out_rows),
                                    int(mode), int(E), int(N), stream_ptr(dev))
    check(rc, "a2cb_gather_rows")


def copy_columns(pairs, dst_cols, src_cols, n_cols):
    """dst[:, dst_cols[j]] = src[:, src_cols[j]] for j < n_cols, for every (dst, src) pair of [rows, columns, ...] device
    tensors (same rows and trailing shape byte for byte; any strides over the first two dimensions, trailing dimensions
    dense) in ONE launch on the current stream.  `dst_cols` / `src_cols`: device int64 tensors with at least n_cols
    entries, or None for column j."""
    if n_cols <= 0 or not pairs:
        return
    arr = (ColCopyField * len(pairs))()
    dev = pairs[0][0].device
    for f, (d, s) in zip(arr, pairs):
        if not (d.is_cuda and s.is_cuda) or d.shape[0] != s.shape[0] or d.dim() < 2 or s.dim() < 2:
            raise NativeError("copy_columns: pairs of CUDA tensors [rows, columns, ...] with the same number of rows")
        wd = d[0, 0].numel() * d.element_size()
        ws = s[0, 0].numel() * s.element_size()
        if wd != ws or not d[0, 0].is_contiguous() or not s[0, 0].is_contiguous():
            raise NativeError("copy_columns: elements must be dense and of the same size (%d vs %d bytes)" % (wd, ws))
        f.dst, f.src = d.data_ptr(), s.data_ptr()
        f.dst_rs, f.dst_cs = d.stride(0) * d.element_size(), d.stride(1) * d.element_size()
        f.src_rs, f.src_cs = s.stride(0) * s.element_size(), s.stride(1) * s.element_size()
        f.w, f.rows = wd, d.shape[0]
    for c in (dst_cols, src_cols):
        if c is not None and (c.dtype != torch.int64 or not c.is_cuda or not c.is_contiguous() or c.numel() < n_cols):
            raise NativeError("copy_columns: column lists are contiguous int64 CUDA tensors with >= n_cols entries")
    with torch.cuda.device(dev):
        rc = lib().a2cb_copy_columns(arr, len(pairs), c_void_p(dst_cols.data_ptr()) if dst_cols is not None else None,
                                     c_void_p(src_cols.data_ptr()) if src_cols is not None else None, int(n_cols),
                                     stream_ptr(dev))
    check(rc, "a2cb_copy_columns")


def upload_env_columns(dst, src_host, envs, stream):
    """Columns `envs` (host int64 list) of the time-major [rows, N, ...] pinned host tensor -> the device tensor of
    the same shape, asynchronously on `stream` (one strided DMA copy per run of adjacent environments)."""
    if src_host.is_cuda or not src_host.is_pinned():
        raise NativeError("upload_env_columns: the source must be a pinned host tensor")
    if tuple(src_host.shape) != tuple(dst.shape) or src_host.dtype != dst.dtype or not src_host.is_contiguous():
        raise NativeError("upload_env_columns: source and destination must share shape, dtype and layout")
    rows, N = dst.shape[0], dst.shape[1]
    col_bytes = dst[0, 0].numel() * dst.element_size()
    e = torch.as_tensor(envs, dtype=torch.int64).contiguous()
    with torch.cuda.device(dst.device):
        rc = lib().a2cb_upload_env_columns(dptr(dst, name="upload dst"), c_void_p(src_host.data_ptr()), col_bytes, rows,
                                           N * col_bytes, c_void_p(e.data_ptr()), e.numel(), c_void_p(stream.cuda_stream))
    check(rc, "a2cb_upload_env_columns")
