"""randlapack_b200 — host-side mirror of RandLAPACK's algorithm objects over librlb200.so.

The product is the sm_100a shared library behind the C-ABI of include/rlb200.h (C++ host code +
hand-written CUDA kernels).  This package is the thin Python veneer the tests and bench.py drive it
through: it mirrors the reference's objects for this path — RNGState, DenseDist/fill_dense, CholQRQ,
RS, RF, QB, RSVD (same constructor arguments, public fields, return codes) — and uses torch only for
device memory, streams and torch.distributed plumbing.

Matrices are torch CUDA tensors in COLUMN-MAJOR storage (shape (m, n), strides (1, m)); `empty_f`,
`to_f` build them.  Nothing here computes on the CPU and nothing imports oracle/.
"""
from __future__ import annotations

import ctypes
import os
import sys
from dataclasses import dataclass

from . import _capi
from ._capi import (STAB_PLUL, STAB_CHOLQRQ, STAB_HQRQ, FAMILY_GAUSSIAN, FAMILY_UNIFORM, AXIS_LONG, AXIS_SHORT,  # noqa: F401
                    LAYOUT_NATURAL, LAYOUT_COLMAJOR, LAYOUT_ROWMAJOR, StackOpts, QRCP_LUQR, QRCP_GEQP3, QRTALL_GEQRF,
                    QRTALL_CHOLQR, QRTALL_GEQRT)

__all__ = ["Context", "RNGState", "DenseDist", "fill_dense", "SparseDist", "fill_sparse", "sketch_general_left", "sketch_general_right", "sketch_general_dense_left", "sketch_general_dense_right", "sketch_general_sparse_left", "sketch_general_sparse_right", "CQRRPT", "BQRRP", "hqrrp", "qr_small", "col_swap", "CholQRQ", "PLUL", "HQRQ", "RS", "RF", "QB", "RSVD", "empty_f",
           "to_f", "Error", "shard_rows"]


class Error(RuntimeError):
    """Raised for negative RLB200_ERR_* codes (what RandLAPACK::Error / RandBLAS::Error / a CUDA abort are in the reference)."""

    def __init__(self, code, msg):
        super().__init__(f"rlb200 error {code}: {msg}")
        self.code = code


def _torch():
    import torch
    return torch


def empty_f(m, n, dtype, device):
    """Uninitialised m x n column-major matrix."""
    torch = _torch()
    return torch.empty((n, m), dtype=dtype, device=device).t()


def to_f(x):
    """Column-major copy of a 2-D tensor (no-op if it already is)."""
    if x.dim() == 2 and x.stride(0) == 1 and (x.stride(1) == x.shape[0] or x.shape[1] <= 1):
        return x
    return x.t().contiguous().t()


def _is_f(x):
    """column-major with leading dimension >= rows (or a contiguous vector)"""
    return x.dim() == 2 and (x.stride(0) == 1 or x.shape[0] <= 1) and (x.stride(1) >= x.shape[0] or x.shape[1] <= 1) \
        or x.dim() == 1 and x.is_contiguous()


def _ld(x):
    return x.stride(1) if x.shape[1] > 1 else max(x.shape[0], 1)


def _suffix(dtype):
    torch = _torch()
    if dtype == torch.float64:
        return "f64"
    if dtype == torch.float32:
        return "f32"
    raise TypeError("T must be float or double, as in the reference")


class RNGState:
    """RandBLAS::RNGState<r123::Philox4x32> (RandBLAS/RandBLAS/base.hh:64-164)."""

    def __init__(self, key=0, counter=(0, 0, 0, 0)):
        if isinstance(key, (tuple, list)):
            self.key = tuple(int(x) & 0xFFFFFFFF for x in key)
        else:   # RNGState(uint64 k): zero key incremented by k (base.hh:119)
            self.key = (int(key) & 0xFFFFFFFF, (int(key) >> 32) & 0xFFFFFFFF)
        self.counter = tuple(int(c) & 0xFFFFFFFF for c in counter)

    def words(self):
        return (ctypes.c_uint32 * 6)(*self.counter, *self.key)

    def assign(self, w):
        self.counter, self.key = (w[0], w[1], w[2], w[3]), (w[4], w[5])

    def copy(self):
        return RNGState(self.key, self.counter)

    def __eq__(self, o):
        return self.counter == o.counter and self.key == o.key

    def __repr__(self):
        return f"RNGState(counter={self.counter}, key={self.key})"


class Context:
    """Opaque rlb200_ctx: device, stream, workspaces, optional row-shard description."""

    def __init__(self, device=None, stream=None):
        torch = _torch()
        self._lib = _capi.load()
        if not torch.cuda.is_available():
            raise Error(_capi.ERR_CUDA, "no CUDA device: librlb200 has no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else device.index)
        with torch.cuda.device(self.device):
            s = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        h = ctypes.c_void_p()
        rc = self._lib.rlb200_create(ctypes.byref(h), self.device.index, ctypes.c_void_p(s))
        if rc:
            raise Error(rc, "rlb200_create failed (needs an sm_100 device)")
        self._h = h
        self._hook = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rlb200_destroy(self._h)
            self._h = None

    __del__ = close

    def check(self, rc):
        if rc < 0:
            raise Error(rc, self._lib.rlb200_last_error(self._h).decode())
        return rc

    def synchronize(self):
        self.check(self._lib.rlb200_synchronize(self._h))

    def launch_count(self, reset=False):
        return self._lib.rlb200_launch_count(self._h, int(reset))

    def timers_enable(self, on=True):
        self.check(self._lib.rlb200_timers_enable(self._h, int(on)))

    def timer_read(self, which, reset=False):
        ms, n = ctypes.c_double(), ctypes.c_int64()
        self.check(self._lib.rlb200_timer_read(self._h, which, ctypes.byref(ms), ctypes.byref(n), int(reset)))
        return ms.value, n.value

    def set_fp64_engine(self, engine):
        """Tall products of the drivers: "i8" (tcgen05 int8 digit slices, default) or "dmma" (fp64 tensor pipe)."""
        self.check(self._lib.rlb200_set_fp64_engine(self._h, {"dmma": 0, "i8": 1}[engine]))

    def set_i8_digits(self, digits):
        """Digits per value of the int8-slice engine: 0 = default (6 for fp64 storage, 4 for fp32), else 3..7 (8*digits - 2 bits)."""
        self.check(self._lib.rlb200_set_i8_digits(self._h, int(digits)))

    def phase_timing(self, on=True):
        """Record the reference's per-phase `times` vector (microseconds) during the next CQRRPT / CQRRT / BQRRP call."""
        self.check(self._lib.rlb200_set_phase_timing(self._h, 1 if on else 0))

    def phase_times(self):
        buf = (ctypes.c_longlong * 32)()
        n = self._lib.rlb200_get_phase_times(self._h, buf, 32)
        return [int(buf[i]) for i in range(max(0, min(n, 32)))]

    def set_i8_fused(self, on):
        """True (default): tall products whose shapes allow it slice the tall operand inside the tensor-core kernel (ozaki_fused.cu);
        False: always stage the digits in HBM (ozaki.cu)."""
        self.check(self._lib.rlb200_set_i8_fused(self._h, 1 if on else 0))

    # ---- row sharding over torch.distributed -------------------------------------------------
    def set_shard(self, row_offset, m_global, group=None, native=None):
        """This rank holds rows [row_offset, row_offset + m_local) of an m_global-row A.  Gram / B^T / norm / R-factor partials are
        sum-allreduced over the shards.  native=True: through the context's own NCCL communicator (ncclAllReduce issued from C++ on the
        context's stream; created on first use, its unique id travels over torch.distributed); native=False: through a callback into
        torch.distributed (`group`; gloo in the CPU tests of this plumbing).  Default: native on NCCL process groups."""
        import torch.distributed as dist
        if native is None:
            native = dist.is_initialized() and dist.get_backend(group) == "nccl" and os.environ.get("RLB200_NATIVE_COMM", "1") != "0"
        if dist.is_initialized():
            self.check(self._lib.rlb200_set_shard_rank(self._h, dist.get_rank(group), dist.get_world_size(group)))
        if native:
            if not getattr(self, "_native_comm", False):
                self.comm_init_from_torch(group)
            self._hook = None
            self.check(self._lib.rlb200_set_shard(self._h, row_offset, m_global, ctypes.cast(None, _capi.ALLREDUCE_FN), None))
        else:
            self._hook = _capi.ALLREDUCE_FN(make_allreduce_hook(group))
            self.check(self._lib.rlb200_set_shard(self._h, row_offset, m_global, self._hook, None))

    def comm_init_from_torch(self, group=None):
        """rlb200_comm_unique_id on rank 0, broadcast of the 128 bytes over torch.distributed, rlb200_comm_init on every rank."""
        torch = _torch()
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        idb = (ctypes.c_ubyte * 128)()
        if rank == 0:
            rc = self._lib.rlb200_comm_unique_id(idb)
            if rc:
                raise Error(rc, "rlb200_comm_unique_id failed (libnccl not found?)")
        t = torch.tensor(list(idb), dtype=torch.uint8, device=self.device if dist.get_backend(group) == "nccl" else "cpu")
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        idb = (ctypes.c_ubyte * 128)(*t.cpu().tolist())
        # NCCL may print its version banner on stdout while a communicator is created: keep the process's stdout clean (bench.py prints
        # exactly one JSON line there) by pointing fd 1 at stderr for the duration of the call
        sys.stdout.flush()
        saved = os.dup(1)
        try:
            os.dup2(2, 1)
            rc = self._lib.rlb200_comm_init(self._h, world, rank, idb)
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        self.check(rc)
        self._native_comm = True

    def clear_shard(self):
        self._hook = None
        self.check(self._lib.rlb200_set_shard(self._h, 0, -1, ctypes.cast(None, _capi.ALLREDUCE_FN), None))


def shard_rows(m_global, world_size, rank, align=128):
    """Row block [r0, r1) of rank `rank`: contiguous, aligned to `align` rows (a multiple of 4 keeps every
    shard's first row on a Philox counter boundary of the odd-p operator, RandBLAS dense_skops.hh:109-123)."""
    per = -(-m_global // world_size)
    per = -(-per // align) * align
    if (world_size - 1) * per >= m_global:
        # a trailing rank would hold no rows: the drivers require m > 0 on every shard, and an empty shard returning early would leave
        # the other ranks waiting in a collective - fail on every rank alike, before any collective
        raise ValueError(f"shard_rows: {m_global} rows cannot be split into {world_size} non-empty blocks of multiples of {align} rows")
    r0 = min(m_global, rank * per)
    r1 = min(m_global, r0 + per)
    return r0, r1


class _CudaBuf:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def make_allreduce_hook(group=None):
    """Build the rlb200_allreduce_fn callback: sum-allreduce `count` elements at `buf` over `group`.
    Device buffers (NCCL) are wrapped zero-copy; host buffers (gloo, used by the CPU tests of this plumbing)
    are wrapped through ctypes."""
    torch = _torch()
    import torch.distributed as dist

    def hook(user, buf, count, elem_size, stream):
        try:
            dt = {8: torch.float64, 4: torch.float32}[elem_size]
            backend = dist.get_backend(group)
            if backend == "nccl":
                t = torch.as_tensor(_CudaBuf(buf, count * elem_size), device="cuda").view(dt)
                ext = torch.cuda.ExternalStream(stream) if stream else torch.cuda.default_stream()
                with torch.cuda.stream(ext):
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            else:
                import numpy as np
                arr = np.ctypeslib.as_array(ctypes.cast(buf, ctypes.POINTER(ctypes.c_uint8)), shape=(count * elem_size,))
                t = torch.from_numpy(arr).view(dt)
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            return 0
        except Exception as e:  # never let an exception unwind through the C frames
            import sys
            print(f"[rlb200 allreduce hook] {type(e).__name__}: {e}", file=sys.stderr)
            return 1
    return hook


# --------------------------------------------------------------------------------------------------
# RandBLAS dense operators
# --------------------------------------------------------------------------------------------------
@dataclass
class DenseDist:
    """RandBLAS::DenseDist (RandBLAS/RandBLAS/dense_skops.hh:228-347)."""
    n_rows: int
    n_cols: int
    family: int = FAMILY_GAUSSIAN
    major_axis: int = AXIS_LONG

    def __post_init__(self):
        if self.n_rows <= 0 or self.n_cols <= 0:
            raise Error(_capi.ERR_ARG, "randblas_require(n_rows > 0 && n_cols > 0)")
        mx, mn = max(self.n_rows, self.n_cols), min(self.n_rows, self.n_cols)
        self.dim_major = mx if self.major_axis == AXIS_LONG else mn
        self.dim_minor = self.n_rows + self.n_cols - self.dim_major
        self.isometry_scale = float(self.dim_minor) ** -0.5
        is_wide, fa_long = self.n_rows < self.n_cols, self.major_axis == AXIS_LONG
        self.natural_layout = LAYOUT_ROWMAJOR if (is_wide and fa_long) or (not is_wide and not fa_long) else LAYOUT_COLMAJOR


def fill_dense(ctx: Context, D: DenseDist, state: RNGState, dtype=None, layout=LAYOUT_NATURAL, sub=None):
    """RandBLAS::fill_dense(D, buff, seed) -> next state (dense_skops.hh:620-623) and fill_dense_unpacked (:560-603).
    Returns (buffer as a 1-D device tensor in the requested layout, next RNGState)."""
    torch = _torch()
    dtype = dtype or torch.float64
    sub_rows, sub_cols, ro, co = sub if sub is not None else (D.n_rows, D.n_cols, 0, 0)
    buf = torch.empty(max(sub_rows * sub_cols, 1), dtype=dtype, device=ctx.device)
    w = state.words()
    fn = getattr(ctx._lib, f"rlb200_fill_dense_{_suffix(dtype)}_dev")
    ctx.check(fn(ctx._h, D.n_rows, D.n_cols, D.family, D.major_axis, layout, sub_rows, sub_cols, ro, co, buf.data_ptr(), w))
    nxt = RNGState()
    nxt.assign(w)
    return buf[: sub_rows * sub_cols], nxt


@dataclass
class SparseDist:
    """RandBLAS::SparseDist (RandBLAS/RandBLAS/sparse_skops.hh:167-282); Axis::Short (SASO) is the default; Axis::Long (LASO) is offered for generation (fill_sparse) and
    for the left sketch with a wide operator (sketch_general_left)."""
    n_rows: int
    n_cols: int
    vec_nnz: int = 4
    major_axis: int = AXIS_SHORT

    def __post_init__(self):
        if self.n_rows <= 0 or self.n_cols <= 0 or self.vec_nnz <= 0:
            raise Error(_capi.ERR_ARG, "randblas_require(n_rows > 0 && n_cols > 0 && vec_nnz > 0)")
        mx, mn = max(self.n_rows, self.n_cols), min(self.n_rows, self.n_cols)
        self.dim_major = mn if self.major_axis == AXIS_SHORT else mx
        self.dim_minor = self.n_rows + self.n_cols - self.dim_major
        if self.vec_nnz > self.dim_major:
            raise Error(_capi.ERR_ARG, "randblas_require(vec_nnz <= dim_major)")
        self.full_nnz = self.vec_nnz * self.dim_minor
        self.isometry_scale = self.vec_nnz ** -0.5 if self.major_axis == AXIS_SHORT else \
            (self.dim_major / (self.vec_nnz * float(self.dim_minor))) ** 0.5


def fill_sparse(ctx: Context, D: SparseDist, state: RNGState, dtype=None, sub=None):
    """RandBLAS::fill_sparse_unpacked (sparse_skops.hh:568-704): -> (nnz, vals, rows, cols [device], returned RNGState)."""
    torch = _torch()
    dtype = dtype or torch.float64
    sub_rows, sub_cols, ro, co = sub if sub is not None else (D.n_rows, D.n_cols, 0, 0)
    fn = getattr(ctx._lib, f"rlb200_fill_sparse_{_suffix(dtype)}_dev")
    cap = ctypes.c_int64(0)
    w = state.words()
    ctx.check(fn(ctx._h, D.n_rows, D.n_cols, D.vec_nnz, D.major_axis, sub_rows, sub_cols, ro, co, ctypes.byref(cap), None, None, None, w))
    vals = torch.empty(max(cap.value, 1), dtype=dtype, device=ctx.device)
    rows = torch.empty(max(cap.value, 1), dtype=torch.int64, device=ctx.device)
    cols = torch.empty(max(cap.value, 1), dtype=torch.int64, device=ctx.device)
    nnz = ctypes.c_int64(0)
    ctx.check(fn(ctx._h, D.n_rows, D.n_cols, D.vec_nnz, D.major_axis, sub_rows, sub_cols, ro, co, ctypes.byref(nnz), vals.data_ptr(),
                 rows.data_ptr(), cols.data_ptr(), w))
    nxt = RNGState()
    nxt.assign(w)
    return nnz.value, vals[: nnz.value], rows[: nnz.value], cols[: nnz.value], nxt


def sketch_general_left(ctx: Context, D, state: RNGState, A, d=None, alpha=1.0, beta=0.0, B=None, ro_s=0, co_s=0):
    """RandBLAS::sketch_general(ColMajor, NoTrans, NoTrans, d, n, m, alpha, S, ro_s, co_s, A, lda, beta, B, ldb) with
    S = D.sample(state) regenerated on device (skge.hh:840-1052).  D is a DenseDist or a SparseDist.  `state` <- S.next_state."""
    m, n = A.shape
    assert _is_f(A)
    d = D.n_rows - ro_s if d is None else d
    if B is None:
        B = empty_f(d, n, A.dtype, A.device)
        beta = 0.0
    w = state.words()
    if isinstance(D, SparseDist):
        # Axis::Long (LASO, sparse_skops.hh:669-684): rows of the wide operator drawn with replacement; Axis::Short: SASO
        name = "sketch_sparse_left" if D.major_axis == AXIS_SHORT else "sketch_sparse_left_laso"
        fn = getattr(ctx._lib, f"rlb200_{name}_{_suffix(A.dtype)}_dev")
        ctx.check(fn(ctx._h, D.n_rows, D.n_cols, D.vec_nnz, d, n, m, alpha, ro_s, co_s, A.data_ptr(), _ld(A), beta, B.data_ptr(), _ld(B), w))
    else:
        fn = getattr(ctx._lib, f"rlb200_sketch_dense_left_{_suffix(A.dtype)}_dev")
        ctx.check(fn(ctx._h, D.n_rows, D.n_cols, D.family, D.major_axis, d, n, m, alpha, ro_s, co_s, A.data_ptr(), _ld(A), beta,
                     B.data_ptr(), _ld(B), w))
    state.assign(w)
    return B


def sketch_general_right(ctx: Context, A, D: DenseDist, state: RNGState, d=None, alpha=1.0, beta=0.0, B=None, ro_s=0, co_s=0):
    """sketch_general(ColMajor, NoTrans, NoTrans, m, d, n, alpha, A, lda, S, ro_s, co_s, beta, B, ldb), dense S (rskge3, skge.hh:308-356)."""
    m, n = A.shape
    assert _is_f(A)
    d = D.n_cols - co_s if d is None else d
    if B is None:
        B = empty_f(m, d, A.dtype, A.device)
        beta = 0.0
    w = state.words()
    fn = getattr(ctx._lib, f"rlb200_sketch_dense_right_{_suffix(A.dtype)}_dev")
    ctx.check(fn(ctx._h, D.n_rows, D.n_cols, D.family, D.major_axis, m, d, n, alpha, A.data_ptr(), _ld(A), ro_s, co_s, beta,
                 B.data_ptr(), _ld(B), w))
    state.assign(w)
    return B


def sketch_general_dense_left(ctx: Context, layout, opS, opA, d, n, m, alpha, D: DenseDist, ro_s, co_s, A, lda, beta, B, ldb, state: RNGState):
    """RandBLAS::sketch_general(layout, opS, opA, d, n, m, alpha, S, ro_s, co_s, A, lda, beta, B, ldb) with a DenseSkOp S = D.sample(state)
    (skge.hh:859-905): B(d x n) = alpha * op(submat(S)) * op(A) + beta * B for every layout (LAYOUT_COLMAJOR | LAYOUT_ROWMAJOR) and
    transposition flag (False / True).  A and B are device buffers (any tensor; data_ptr, lda / ldb in elements).  `state` <- S.next_state."""
    w = state.words()
    fn = getattr(ctx._lib, f"rlb200_sketch_general_dense_left_{_suffix(A.dtype)}_dev")
    ctx.check(fn(ctx._h, int(layout), int(bool(opS)), int(bool(opA)), d, n, m, alpha, D.n_rows, D.n_cols, D.family, D.major_axis, ro_s, co_s,
                 A.data_ptr(), lda, beta, B.data_ptr(), ldb, w))
    state.assign(w)
    return B


def sketch_general_dense_right(ctx: Context, layout, opA, opS, m, d, n, alpha, A, lda, D: DenseDist, ro_s, co_s, beta, B, ldb, state: RNGState):
    """RandBLAS::sketch_general(layout, opA, opS, m, d, n, alpha, A, lda, S, ro_s, co_s, beta, B, ldb) (skge.hh:1031-1076):
    B(m x d) = alpha * op(A) * op(submat(S)) + beta * B, every layout / transposition flag."""
    w = state.words()
    fn = getattr(ctx._lib, f"rlb200_sketch_general_dense_right_{_suffix(A.dtype)}_dev")
    ctx.check(fn(ctx._h, int(layout), int(bool(opA)), int(bool(opS)), m, d, n, alpha, A.data_ptr(), lda, D.n_rows, D.n_cols, D.family,
                 D.major_axis, ro_s, co_s, beta, B.data_ptr(), ldb, w))
    state.assign(w)
    return B


def sketch_general_sparse_left(ctx: Context, layout, opS, opA, d, n, m, alpha, D: SparseDist, ro_s, co_s, A, lda, beta, B, ldb, state: RNGState):
    """sketch_general(layout, opS, opA, d, n, m, alpha, S, ro_s, co_s, A, lda, beta, B, ldb) with a short-axis SparseSkOp S = D.sample(state)
    (skge.hh:907-960), every layout / transposition flag; op(submat(S)) must be wide.  A, B: device buffers with lda / ldb in elements."""
    if D.major_axis != AXIS_SHORT:
        raise Error(_capi.ERR_UNSUPPORTED, "Axis::Long sparse operators are not offered on the device")
    w = state.words()
    fn = getattr(ctx._lib, f"rlb200_sketch_general_sparse_left_{_suffix(A.dtype)}_dev")
    ctx.check(fn(ctx._h, int(layout), int(bool(opS)), int(bool(opA)), d, n, m, alpha, D.n_rows, D.n_cols, D.vec_nnz, ro_s, co_s, A.data_ptr(), lda,
                 beta, B.data_ptr(), ldb, w))
    state.assign(w)
    return B


def sketch_general_sparse_right(ctx: Context, layout, opA, opS, m, d, n, alpha, A, lda, D: SparseDist, ro_s, co_s, beta, B, ldb, state: RNGState):
    """sketch_general(layout, opA, opS, m, d, n, alpha, A, lda, S, ro_s, co_s, beta, B, ldb) with a short-axis SparseSkOp (skge.hh:1078-1131)."""
    if D.major_axis != AXIS_SHORT:
        raise Error(_capi.ERR_UNSUPPORTED, "Axis::Long sparse operators are not offered on the device")
    w = state.words()
    fn = getattr(ctx._lib, f"rlb200_sketch_general_sparse_right_{_suffix(A.dtype)}_dev")
    ctx.check(fn(ctx._h, int(layout), int(bool(opA)), int(bool(opS)), m, d, n, alpha, A.data_ptr(), lda, D.n_rows, D.n_cols, D.vec_nnz, ro_s, co_s,
                 beta, B.data_ptr(), ldb, w))
    state.assign(w)
    return B


def philox_stream(ctx: Context, state: RNGState, n: int):
    torch = _torch()
    out = torch.empty((max(n, 1), 4), dtype=torch.int32, device=ctx.device)
    ctx.check(ctx._lib.rlb200_philox_stream_dev(ctx._h, state.words(), n, out.data_ptr()))
    return out[:n]


def gemm(ctx: Context, transa, transb, alpha, A, B, beta=0.0, C=None, engine="dmma"):
    """blas::gemm(ColMajor, ...) on column-major device tensors, shapes as BLAS defines them.
    engine: "dmma" (fp64 tensor pipe) or "i8" (tcgen05 int8 digit slices)."""
    m = A.shape[1] if transa else A.shape[0]
    k = A.shape[0] if transa else A.shape[1]
    n = B.shape[0] if transb else B.shape[1]
    assert _is_f(A) and _is_f(B)
    if C is None:
        C = empty_f(m, n, A.dtype, A.device)
        beta = 0.0
    fn = getattr(ctx._lib, f"rlb200_gemm_{_suffix(A.dtype)}_dev" if engine == "dmma" else f"rlb200_gemm_{_suffix(A.dtype)}_i8_dev")
    ctx.check(fn(ctx._h, int(transa), int(transb), m, n, k, alpha, A.data_ptr(), _ld(A), B.data_ptr(), _ld(B), beta, C.data_ptr(), _ld(C)))
    return C


# --------------------------------------------------------------------------------------------------
# Algorithm objects (RandLAPACK/comps, RandLAPACK/drivers)
# --------------------------------------------------------------------------------------------------
class _Stab:
    kind = None

    def __init__(self, cond_check=False, verbose=False):
        self.cond_check, self.verbose = cond_check, verbose
        self.chol_fail = False

    def call(self, ctx: Context, A):
        """Stabilization<T>::call(m, k, A) (rl_orth.hh:13-23): in place; returns the reference's int code."""
        assert _is_f(A)
        m, k = A.shape
        cf = ctypes.c_int(0)
        fn = getattr(ctx._lib, f"rlb200_stab_{_suffix(A.dtype)}_dev")
        rc = ctx.check(fn(ctx._h, self.kind, m, k, A.data_ptr(), int(self.cond_check), ctypes.byref(cf)))
        self.chol_fail = bool(cf.value)
        return rc


class CholQRQ(_Stab):
    """RandLAPACK::CholQRQ (rl_orth.hh:25-98)."""
    kind = STAB_CHOLQRQ


class PLUL(_Stab):
    """RandLAPACK::PLUL (rl_orth.hh:166-230)."""
    kind = STAB_PLUL


class HQRQ(_Stab):
    """RandLAPACK::HQRQ (rl_orth.hh:100-164)."""
    kind = STAB_HQRQ


class RS:
    """RandLAPACK::RS(stab, p, q, verbose, cond_check) (rl_rs.hh:31-178)."""

    def __init__(self, stab_obj, p, q, verbose=False, cond_check=False):
        self.Stab_Obj, self.passes_over_data, self.passes_per_stab = stab_obj, p, q
        self.verbose, self.cond_check = verbose, cond_check
        # fold CholQR's m x k triangular solve into the next product ((A^T Y) R^-1 instead of A^T (Y R^-1)); same outputs to round-off
        self.fold_solves = True

    def _opts(self, o: StackOpts):
        o.passes_over_data, o.passes_per_stab, o.stab = self.passes_over_data, self.passes_per_stab, self.Stab_Obj.kind
        o.cond_check = int(self.cond_check or self.Stab_Obj.cond_check)
        o.reserved = 0 if self.fold_solves else 1

    def call(self, ctx: Context, A, k, state: RNGState):
        """-> (rc, Omega n x k).  `state` is advanced in place like the reference's in/out reference."""
        torch = _torch()
        m, n = A.shape
        o = StackOpts()
        self._opts(o)
        Omega = empty_f(n, k, A.dtype, A.device)
        work = empty_f(m, k, A.dtype, A.device) if self.passes_over_data > 0 else None
        w = state.words()
        fn = getattr(ctx._lib, f"rlb200_rs_{_suffix(A.dtype)}_dev")
        rc = ctx.check(fn(ctx._h, m, n, A.data_ptr(), k, Omega.data_ptr(), work.data_ptr() if work is not None else None, w,
                          ctypes.byref(o)))
        state.assign(w)
        return rc, Omega


class RF:
    """RandLAPACK::RF(rs, orth, verbose, cond_check) (rl_rf.hh:31-137)."""

    def __init__(self, rs_obj, orth_obj, verbose=False, cond_check=False):
        self.rs, self.orth, self.verbose, self.cond_check = rs_obj, orth_obj, verbose, cond_check

    def _opts(self, o: StackOpts):
        self.rs._opts(o)
        o.orth_rf = self.orth.kind

    def call(self, ctx: Context, A, k, state: RNGState, Q=None):
        m, n = A.shape
        o = StackOpts()
        self._opts(o)
        if Q is None:
            Q = empty_f(m, k, A.dtype, A.device)
        w = state.words()
        fn = getattr(ctx._lib, f"rlb200_rf_{_suffix(A.dtype)}_dev")
        rc = ctx.check(fn(ctx._h, m, n, A.data_ptr(), k, Q.data_ptr(), w, ctypes.byref(o)))
        state.assign(w)
        return rc, Q


class QB:
    """RandLAPACK::QB(rf, orth, verbose, orth_check) (rl_qb.hh:36-268)."""

    def __init__(self, rf_obj, orth_obj, verbose=False, orth_check=False):
        self.rf, self.orth, self.verbose, self.orth_check = rf_obj, orth_obj, verbose, orth_check

    def _opts(self, o: StackOpts):
        self.rf._opts(o)
        o.orth_qb, o.orth_check = self.orth.kind, int(self.orth_check)

    def call(self, ctx: Context, A, k, block_sz, tol, state: RNGState, Q=None, BT=None):
        """-> (rc, k_out, Q m x k, BT n x k); columns beyond k_out are unspecified."""
        m, n = A.shape
        o = StackOpts()
        self._opts(o)
        o.block_sz = block_sz
        if Q is None:
            Q = empty_f(m, k, A.dtype, A.device)
        if BT is None:
            BT = empty_f(n, k, A.dtype, A.device)
        kk = ctypes.c_int64(k)
        w = state.words()
        fn = getattr(ctx._lib, f"rlb200_qb_{_suffix(A.dtype)}_dev")
        rc = ctx.check(fn(ctx._h, m, n, A.data_ptr(), ctypes.byref(kk), block_sz, tol, Q.data_ptr(), BT.data_ptr(), None, w,
                          ctypes.byref(o)))
        state.assign(w)
        return rc, kk.value, Q, BT


class RSVD:
    """RandLAPACK::RSVD(qb, block_sz) (rl_rsvd.hh:34-154)."""

    def __init__(self, qb_obj, block_sz):
        self.QB_Obj, self.block_sz = qb_obj, block_sz
        self.qb_code = None

    def _opts(self):
        o = StackOpts()
        self.QB_Obj._opts(o)
        o.block_sz = self.block_sz
        return o

    def call(self, ctx: Context, A, k, tol, state: RNGState, U=None, S=None, V=None):
        """Device-resident call -> (rc, k_out, U m x k, S k, V n x k).  A is only read when block_sz == k."""
        torch = _torch()
        m, n = A.shape
        o = self._opts()
        if U is None:
            U = empty_f(m, k, A.dtype, A.device)
        if S is None:
            S = torch.empty(k, dtype=A.dtype, device=A.device)
        if V is None:
            V = empty_f(n, k, A.dtype, A.device)
        kk, qc = ctypes.c_int64(k), ctypes.c_int(0)
        w = state.words()
        fn = getattr(ctx._lib, f"rlb200_rsvd_{_suffix(A.dtype)}_dev")
        rc = ctx.check(fn(ctx._h, m, n, A.data_ptr(), ctypes.byref(kk), tol, U.data_ptr(), S.data_ptr(), V.data_ptr(), None, w,
                          ctypes.byref(o), ctypes.byref(qc)))
        state.assign(w)
        self.qb_code = qc.value
        return rc, kk.value, U, S, V

    def call_host(self, ctx: Context, A_host, k, tol, state: RNGState, U=None, S=None, V=None):
        """The reference-facing form: HOST column-major A in, host U, S, V out (copies inside).
        U, S, V may be preallocated (e.g. pinned) host buffers of the requested k."""
        torch = _torch()
        m, n = A_host.shape
        assert _is_f(A_host) and not A_host.is_cuda
        o = self._opts()
        U = empty_f(m, k, A_host.dtype, "cpu") if U is None else U
        S = torch.empty(k, dtype=A_host.dtype) if S is None else S
        V = empty_f(n, k, A_host.dtype, "cpu") if V is None else V
        kk, qc = ctypes.c_int64(k), ctypes.c_int(0)
        w = state.words()
        fn = getattr(ctx._lib, f"rlb200_rsvd_{_suffix(A_host.dtype)}_host")
        rc = ctx.check(fn(ctx._h, m, n, A_host.data_ptr(), ctypes.byref(kk), tol, U.data_ptr(), S.data_ptr(), V.data_ptr(), w,
                          ctypes.byref(o), ctypes.byref(qc)))
        state.assign(w)
        self.qb_code = qc.value
        return rc, kk.value, U, S, V


class CQRRPT:
    """RandLAPACK::CQRRPT(time_subroutines, eps) (rl_cqrrpt.hh:45-146); public fields nnz (SASO non-zeros per column, default 2), rank.
    qrcp: "geqp3" (the reference's default), "bqrrp" or "hqrrp" (rl_cqrrpt.hh:39-43, 230-247); nb_alg, oversampling, panel_pivoting,
    use_cholqr are the HQRRP fields (:134-137, constructor defaults :60-63)."""

    def __init__(self, time_subroutines=False, eps=None):
        self.timing, self.eps, self.nnz, self.rank, self.qrcp = time_subroutines, eps, 2, None, "geqp3"
        self.nb_alg, self.oversampling, self.panel_pivoting, self.use_cholqr = 64, 10, 1, 0
        self.orthogonalization = False      # rl_cqrrpt.hh:139-142: R keeps the Cholesky factor, Q is completed to n orthonormal columns

    def _set_qrcp(self, ctx):
        kinds = {"geqp3": _capi.CQRRPT_QRCP_GEQP3, "bqrrp": _capi.CQRRPT_QRCP_BQRRP, "hqrrp": _capi.CQRRPT_QRCP_HQRRP}
        if self.qrcp not in kinds:
            raise Error(_capi.ERR_ARG, f"CQRRPT qrcp {self.qrcp!r}: geqp3, bqrrp and hqrrp are the reference's choices")
        ctx.check(ctx._lib.rlb200_set_cqrrpt_qrcp(ctx._h, kinds[self.qrcp]))
        ctx.check(ctx._lib.rlb200_set_cqrrpt_hqrrp_opts(ctx._h, int(self.nb_alg), int(self.oversampling), int(self.panel_pivoting),
                                                         int(self.use_cholqr)))
        ctx.check(ctx._lib.rlb200_set_cqrrpt_orthogonalization(ctx._h, int(bool(self.orthogonalization))))

    def _eps(self, dtype):
        torch = _torch()
        return float(torch.finfo(dtype).eps) ** 0.85 if self.eps is None else self.eps

    def call(self, ctx: Context, A, d_factor, state: RNGState, R=None, J=None):
        """A (m x n device, column-major) is overwritten by Q; -> (rc, R n x n [first rank rows valid], J int64 1-based)."""
        torch = _torch()
        assert _is_f(A)
        m, n = A.shape
        if R is None:
            R = torch.zeros((n, n), dtype=A.dtype, device=A.device).t()
        if J is None:
            J = torch.zeros(n, dtype=torch.int64, device=A.device)
        rank = ctypes.c_int64(0)
        w = state.words()
        fn = getattr(ctx._lib, f"rlb200_cqrrpt_{_suffix(A.dtype)}_dev")
        self._set_qrcp(ctx)
        rc = ctx.check(fn(ctx._h, m, n, A.data_ptr(), _ld(A), R.data_ptr(), _ld(R), J.data_ptr(), d_factor, self._eps(A.dtype), self.nnz,
                          ctypes.byref(rank), w))
        state.assign(w)
        self.rank = rank.value
        return rc, R, J

    def call_host(self, ctx: Context, A_host, d_factor, state: RNGState, R=None, J=None):
        """The reference-facing form: HOST column-major A (overwritten by Q), host R / J out; copies inside."""
        torch = _torch()
        assert _is_f(A_host) and not A_host.is_cuda
        m, n = A_host.shape
        R = torch.zeros((n, n), dtype=A_host.dtype).t() if R is None else R
        J = torch.zeros(n, dtype=torch.int64) if J is None else J
        rank = ctypes.c_int64(0)
        w = state.words()
        fn = getattr(ctx._lib, f"rlb200_cqrrpt_{_suffix(A_host.dtype)}_host")
        self._set_qrcp(ctx)
        rc = ctx.check(fn(ctx._h, m, n, A_host.data_ptr(), _ld(A_host), R.data_ptr(), _ld(R), J.data_ptr(), d_factor,
                          self._eps(A_host.dtype), self.nnz, ctypes.byref(rank), w))
        state.assign(w)
        self.rank = rank.value
        return rc, R, J


class CQRRT:
    """RandLAPACK::CQRRT(time_subroutines, eps) (rl_cqrrt.hh:39-89); public fields nnz (= 2), orthogonalization (False), compute_Q (True)."""

    def __init__(self, time_subroutines=False, eps=None):
        self.timing, self.eps = time_subroutines, eps
        self.nnz, self.orthogonalization, self.compute_Q = 2, False, True

    def call(self, ctx: Context, A, d_factor, state: RNGState, R=None):
        """A (m x n device, column-major) is overwritten by Q; -> (rc, R n x n column-major device)."""
        torch = _torch()
        assert _is_f(A)
        m, n = A.shape
        R = torch.zeros((n, n), dtype=A.dtype, device=A.device).t() if R is None else R
        w = state.words()
        fn = getattr(ctx._lib, f"rlb200_cqrrt_{_suffix(A.dtype)}_dev")
        rc = ctx.check(fn(ctx._h, m, n, A.data_ptr(), _ld(A), R.data_ptr(), _ld(R), d_factor, self.nnz, int(self.orthogonalization),
                          int(self.compute_Q), w))
        state.assign(w)
        return rc, R

    def call_host(self, ctx: Context, A_host, d_factor, state: RNGState, R=None):
        torch = _torch()
        assert _is_f(A_host) and not A_host.is_cuda
        m, n = A_host.shape
        R = torch.zeros((n, n), dtype=A_host.dtype).t() if R is None else R
        w = state.words()
        fn = getattr(ctx._lib, f"rlb200_cqrrt_{_suffix(A_host.dtype)}_host")
        rc = ctx.check(fn(ctx._h, m, n, A_host.data_ptr(), _ld(A_host), R.data_ptr(), _ld(R), d_factor, self.nnz, int(self.orthogonalization),
                          int(self.compute_Q), w))
        state.assign(w)
        return rc, R


class SYPS:
    """RandLAPACK::SYPS(p, q, verbose, cond_check) (rl_syps.hh:21-143): power sketch of a symmetric operator, Householder-QR stabilised."""

    def __init__(self, passes_over_data, passes_per_stab, verbose=False, cond_check=False):
        self.passes_over_data, self.passes_per_stab = passes_over_data, passes_per_stab

    def call(self, ctx: Context, uplo, A, k, state: RNGState):
        """A: m x m device column-major (only the `uplo` triangle is read) -> (0, skop m x k)."""
        assert _is_f(A) and A.shape[0] == A.shape[1]
        m = A.shape[0]
        skop, work = empty_f(m, k, A.dtype, A.device), empty_f(m, k, A.dtype, A.device)
        w = state.words()
        fn = getattr(ctx._lib, f"rlb200_syps_{_suffix(A.dtype)}_dev")
        rc = ctx.check(fn(ctx._h, _uplo(uplo), m, A.data_ptr(), _ld(A), k, self.passes_over_data, self.passes_per_stab, skop.data_ptr(),
                          work.data_ptr(), w))
        state.assign(w)
        return rc, skop


class SYRF:
    """RandLAPACK::SYRF(syps, orth) (rl_syrf.hh:21-118): Q = orth(A * syps(A))."""

    def __init__(self, syps: SYPS, orth: _Stab, verbose=False, cond_check=False):
        self.syps, self.orth = syps, orth

    def _opts(self, error_est_p=0):
        return _capi.Revd2Opts(self.syps.passes_over_data, self.syps.passes_per_stab, self.orth.kind, error_est_p)

    def call(self, ctx: Context, uplo, A, k, state: RNGState):
        assert _is_f(A) and A.shape[0] == A.shape[1]
        m = A.shape[0]
        Q, work = empty_f(m, k, A.dtype, A.device), empty_f(m, k, A.dtype, A.device)
        w = state.words()
        o = self._opts()
        fn = getattr(ctx._lib, f"rlb200_syrf_{_suffix(A.dtype)}_dev")
        rc = ctx.check(fn(ctx._h, _uplo(uplo), m, A.data_ptr(), _ld(A), k, Q.data_ptr(), work.data_ptr(), w, ctypes.byref(o)))
        state.assign(w)
        return rc, Q


class REVD2:
    """RandLAPACK::REVD2(syrf, error_est_power_iters) (rl_revd2.hh:75-246): rank-revealing randomized eigendecomposition of a PSD matrix.
    call -> (rc, k, V m x k, eigvals k); rc 1 / 2 are the reference's std::runtime_error cases, 3 = k_cap reached (include/rlb200.h)."""

    def __init__(self, syrf: SYRF, error_est_power_iters, verbose=False):
        self.syrf, self.error_est_p, self.err = syrf, error_est_power_iters, None

    def _run(self, ctx, host, uplo, A, k, tol, state, k_cap):
        torch = _torch()
        assert _is_f(A) and A.shape[0] == A.shape[1] and A.is_cuda != host
        m = A.shape[0]
        # without an explicit capacity the outputs are sized for 8 k; if the rank estimate outgrows that (code 3) the call is repeated from
        # the caller's state with four times the room (deterministic: the repeated prefix reproduces itself)
        auto, k_cap = k_cap is None, (min(m, max(8 * k, 64)) if k_cap is None else k_cap)
        err = (ctypes.c_double if A.dtype == torch.float64 else ctypes.c_float)(0)
        o = self.syrf._opts(self.error_est_p)
        fn = getattr(ctx._lib, f"rlb200_revd2_{_suffix(A.dtype)}_{'host' if host else 'dev'}")
        while True:
            V = empty_f(m, k_cap, A.dtype, A.device)
            ev = torch.zeros(k_cap, dtype=A.dtype, device=A.device)
            kk = ctypes.c_int64(k)
            w = state.words()
            rc = ctx.check(fn(ctx._h, _uplo(uplo), m, A.data_ptr(), _ld(A), ctypes.byref(kk), k_cap, tol, V.data_ptr(), ev.data_ptr(), w,
                              ctypes.byref(o), ctypes.byref(err)))
            if rc != 3 or not auto or k_cap >= m:
                break
            k_cap = min(m, 4 * k_cap)
        state.assign(w)
        self.err = err.value
        return rc, kk.value, V[:, :kk.value], ev[:kk.value]

    def call(self, ctx: Context, uplo, A, k, tol, state: RNGState, k_cap=None):
        return self._run(ctx, False, uplo, A, k, tol, state, k_cap)

    def call_host(self, ctx: Context, uplo, A_host, k, tol, state: RNGState, k_cap=None):
        return self._run(ctx, True, uplo, A_host, k, tol, state, k_cap)


def _uplo(u):
    if u in (_capi.UPLO_UPPER, "U", "u", "upper", "Upper"):
        return _capi.UPLO_UPPER
    if u in (_capi.UPLO_LOWER, "L", "l", "lower", "Lower"):
        return _capi.UPLO_LOWER
    raise Error(_capi.ERR_ARG, f"uplo {u!r}")


class BQRRP:
    """RandLAPACK::BQRRP(time_subroutines, b_sz) (rl_bqrrp.hh:43-152); public fields block_size, qrcp_wide, qr_tall, rank.
    Defaults are the reference's (luqr + geqrf); BQRRP_GPU's configuration is qr_tall = QRTALL_CHOLQR."""

    def __init__(self, time_subroutines=False, b_sz=256):
        if b_sz <= 0:
            raise Error(_capi.ERR_ARG, "randlapack_require(b_sz > 0)")
        self.timing, self.block_size, self.rank = time_subroutines, b_sz, None
        self.qrcp_wide, self.qr_tall = _capi.QRCP_LUQR, _capi.QRTALL_GEQRF
        self.tol = None      # BQRRP::tol (rl_bqrrp.hh:141); None = the constructor default (machine epsilon of the working type)

    def _set_tol(self, ctx):
        ctx.check(ctx._lib.rlb200_set_bqrrp_tol(ctx._h, 0.0 if self.tol is None else float(self.tol)))

    def call(self, ctx: Context, A, d_factor, state: RNGState, tau=None, J=None):
        """A (m x n device, column-major) is overwritten GEQP3-style -> (rc, tau (n), J int64 1-based)."""
        torch = _torch()
        assert _is_f(A)
        m, n = A.shape
        tau = torch.zeros(n, dtype=A.dtype, device=A.device) if tau is None else tau
        J = torch.zeros(n, dtype=torch.int64, device=A.device) if J is None else J
        rank = ctypes.c_int64(0)
        w = state.words()
        fn = getattr(ctx._lib, f"rlb200_bqrrp_{_suffix(A.dtype)}_dev")
        self._set_tol(ctx)
        rc = ctx.check(fn(ctx._h, m, n, A.data_ptr(), _ld(A), d_factor, self.block_size, self.qrcp_wide, self.qr_tall, tau.data_ptr(),
                          J.data_ptr(), ctypes.byref(rank), w))
        state.assign(w)
        self.rank = rank.value
        return rc, tau, J

    def call_sk(self, ctx: Context, A, A_sk, tau=None, J=None):
        """BQRRP_GPU::call(m, n, A, lda, A_sk, d, tau, J) (rl_bqrrp_gpu.hh:122-133): the d x n sketch A_sk (device, column-major) is an
        input and is overwritten; LUQR pivoting; qr_tall = this object's (QRTALL_GEQRF or QRTALL_CHOLQR)."""
        torch = _torch()
        assert _is_f(A) and _is_f(A_sk) and A_sk.dtype == A.dtype
        m, n = A.shape
        d = A_sk.shape[0]
        assert A_sk.shape[1] == n and _ld(A_sk) == d
        tau = torch.zeros(n, dtype=A.dtype, device=A.device) if tau is None else tau
        J = torch.zeros(n, dtype=torch.int64, device=A.device) if J is None else J
        rank = ctypes.c_int64(0)
        fn = getattr(ctx._lib, f"rlb200_bqrrp_{_suffix(A.dtype)}_dev_sk")
        self._set_tol(ctx)
        rc = ctx.check(fn(ctx._h, m, n, A.data_ptr(), _ld(A), A_sk.data_ptr(), d, self.block_size, self.qr_tall, tau.data_ptr(), J.data_ptr(),
                          ctypes.byref(rank)))
        self.rank = rank.value
        return rc, tau, J

    def call_host(self, ctx: Context, A_host, d_factor, state: RNGState, tau=None, J=None):
        torch = _torch()
        assert _is_f(A_host) and not A_host.is_cuda
        m, n = A_host.shape
        tau = torch.zeros(n, dtype=A_host.dtype) if tau is None else tau
        J = torch.zeros(n, dtype=torch.int64) if J is None else J
        rank = ctypes.c_int64(0)
        w = state.words()
        fn = getattr(ctx._lib, f"rlb200_bqrrp_{_suffix(A_host.dtype)}_host")
        self._set_tol(ctx)
        rc = ctx.check(fn(ctx._h, m, n, A_host.data_ptr(), _ld(A_host), d_factor, self.block_size, self.qrcp_wide, self.qr_tall,
                          tau.data_ptr(), J.data_ptr(), ctypes.byref(rank), w))
        state.assign(w)
        self.rank = rank.value
        return rc, tau, J


def hqrrp(ctx: Context, A, nb_alg, pp, panel_pivoting, qr_type, state: RNGState, tau=None, J=None):
    """RandLAPACK::hqrrp(m, n, A, lda, jpvt, tau, nb_alg, pp, panel_pivoting, qr_type, state, timing) (rl_hqrrp.hh:811-1196): Householder QR
    with randomized pivoting.  A (m x n, column-major; device or host tensor) is overwritten GEQP3-style -> (rc, tau (n), J int64 1-based)."""
    torch = _torch()
    assert _is_f(A)
    m, n = A.shape
    tau = torch.zeros(n, dtype=A.dtype, device=A.device) if tau is None else tau
    J = torch.zeros(n, dtype=torch.int64, device=A.device) if J is None else J
    w = state.words()
    fn = getattr(ctx._lib, f"rlb200_hqrrp_{_suffix(A.dtype)}_{'dev' if A.is_cuda else 'host'}")
    rc = ctx.check(fn(ctx._h, m, n, A.data_ptr(), _ld(A), J.data_ptr(), tau.data_ptr(), int(nb_alg), int(pp), int(panel_pivoting),
                      int(qr_type), w))
    state.assign(w)
    return rc, tau, J


def qr_small(ctx: Context, A, pivot=True):
    """lapack::geqp3 (pivot) / geqrf of a small device matrix, in place -> (J int64 1-based or None, tau)."""
    torch = _torch()
    d, n = A.shape
    J = torch.zeros(n, dtype=torch.int64, device=A.device)
    tau = torch.zeros(max(min(d, n), 1), dtype=A.dtype, device=A.device)
    fn = getattr(ctx._lib, f"rlb200_qr_small_{_suffix(A.dtype)}_dev")
    ctx.check(fn(ctx._h, int(pivot), d, n, A.data_ptr(), _ld(A), J.data_ptr(), tau.data_ptr()))
    return (J if pivot else None), tau[: min(d, n)]


def col_swap(ctx: Context, A, idx):
    """util::col_swap / lapack::lapmt(forward): column i of A <- old column idx[i]-1 (idx: 1-based, host sequence)."""
    m, n = A.shape
    arr = (ctypes.c_int64 * n)(*[int(x) for x in idx])
    fn = getattr(ctx._lib, f"rlb200_col_swap_{_suffix(A.dtype)}_dev")
    ctx.check(fn(ctx._h, m, n, A.data_ptr(), _ld(A), arr))
    return A


def svd_tall(ctx: Context, B):
    """lapack::gesdd(SomeVec) of a tall n x k device matrix: returns (left vectors n x k [in place], S, right vectors k x k)."""
    torch = _torch()
    n, k = B.shape
    S = torch.empty(k, dtype=B.dtype, device=B.device)
    W = empty_f(k, k, B.dtype, B.device)
    fn = getattr(ctx._lib, f"rlb200_svd_tall_{_suffix(B.dtype)}_dev")
    ctx.check(fn(ctx._h, n, k, B.data_ptr(), S.data_ptr(), W.data_ptr()))
    return B, S, W
