"""Device-side driver of the MB-PLS fit path: allocates through PyTorch, launches the sm_100a
kernels through the C ABI (``_cabi``), and -- when features are sharded over several GPUs -- issues
the small NCCL all-reduces.  PyTorch is plumbing here (memory, streams, ``torch.distributed``); every
arithmetic step on the path is one of the kernels in ``csrc/``.

Layouts (see ``include/mbpls_b200.h``): matrices are feature-major ``p x ld`` (``ld = n`` rounded up
to 16), results component-major ``K x p`` / ``K x ld``.
"""
from __future__ import annotations

import ctypes as C
import os
import math
import warnings
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _cabi
from ._cabi import call

F64 = torch.float64


def round_ld(n: int) -> int:
    return max(16, (int(n) + 15) // 16 * 16)


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(device=None) -> torch.device:
    """The product path needs a CUDA device and the built library; fail loudly otherwise."""
    _cabi.load()
    if not torch.cuda.is_available():
        raise _cabi.MbplsCudaError("mbpls_b200 requires a CUDA device (sm_100a); there is no CPU fallback")
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise _cabi.MbplsCudaError("mbpls_b200 requires a CUDA device; got %r" % (device,))
    return device


# --------------------------------------------------------------------------------------------- #
# sharding of the concatenated feature axis: contiguous, near-equal ranges
# irrespective of block boundaries; every rank keeps a (block, local range) table.
# --------------------------------------------------------------------------------------------- #
@dataclass
class ShardMap:
    sizes: List[int]            # global block sizes p_b
    rank: int = 0
    world: int = 1
    lo: int = 0                 # global feature range owned by this rank
    hi: int = 0
    local_ranges: List[tuple] = field(default_factory=list)  # per block: (start, stop) *within the block*
    block_off: List[int] = field(default_factory=list)       # B+1 local offsets

    @staticmethod
    def build(sizes: Sequence[int], rank: int = 0, world: int = 1) -> "ShardMap":
        sizes = [int(s) for s in sizes]
        p = sum(sizes)
        per = -(-p // world)
        lo, hi = min(p, rank * per), min(p, (rank + 1) * per)
        m = ShardMap(sizes=sizes, rank=rank, world=world, lo=lo, hi=hi)
        off, g0 = [0], 0
        for pb in sizes:
            a, b = max(lo, g0), min(hi, g0 + pb)
            if b > a:
                m.local_ranges.append((a - g0, b - g0))
                off.append(off[-1] + (b - a))
            else:
                m.local_ranges.append((0, 0))
                off.append(off[-1])
            g0 += pb
        m.block_off = off
        return m

    @property
    def p_local(self) -> int:
        return self.block_off[-1]

    @property
    def p_global(self) -> int:
        return sum(self.sizes)


def _i32(vals, device):
    return torch.tensor(list(vals), dtype=torch.int32, device=device)


# --------------------------------------------------------------------------------------------- #
# ingest: host / device arrays -> feature-major device matrix
# --------------------------------------------------------------------------------------------- #
_STAGE_BYTES = 64 << 20
_COPY_STREAMS: dict = {}


def _copy_stream(device) -> "torch.cuda.Stream":
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    st = _COPY_STREAMS.get(key)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _COPY_STREAMS[key] = st
    return st


def ingest_feature_major(block, n: int, c0: int, c1: int, dst: torch.Tensor, device) -> None:
    """Copy columns [c0, c1) of one n x p_b block into ``dst`` ((c1-c0) x ld, feature-major).

    Row-major sources (what ``check_array`` yields in the reference, mbpls.py:310) are streamed in row
    chunks through one reusable device staging buffer and transposed by ``mbpls_transpose_in_f64``;
    copies and kernels are stream-ordered, so pinned host sources run at PCIe rate without host syncs.
    Column-major sources (numpy Fortran order, or ``t().contiguous().t()`` torch views) are already
    feature-major and are copied straight in.
    """
    ld = dst.shape[1]
    cols = c1 - c0
    if cols <= 0:
        return
    # float32 sources stay float32 until they are on the device (half the PCIe bytes; the reference widens them on the host,
    # check_array(dtype=float64) at mbpls.py:310); anything else that is not float64 is widened here
    if isinstance(block, torch.Tensor):
        t = block if block.dtype in (F64, torch.float32) else block.to(F64)
    else:
        arr = np.asarray(block)
        if arr.dtype not in (np.float64, np.float32):
            arr = arr.astype(np.float64)
        with warnings.catch_warnings():  # read-only sources (memory maps, pandas views) are only ever read
            warnings.simplefilter("ignore")
            t = torch.from_numpy(arr)
    view = t[:, c0:c1]
    f32 = view.dtype == torch.float32
    tr_in = "mbpls_transpose_in_f32" if f32 else "mbpls_transpose_in_f64"
    if view.stride(0) == 1 and n > 1:  # column-major == feature-major
        dst[:, :n].copy_(view.t(), non_blocking=True)  # (widens float32 on the device side of the copy)
        torch.cuda.current_stream(device).synchronize()
        return
    if view.stride(1) != 1:
        view = view.contiguous()
    if view.is_cuda:
        call(tr_in, ptr(view), view.stride(0), n, cols, ptr(dst), ld, 0, stream_ptr(device))
        torch.cuda.current_stream(device).synchronize()  # `view` may be a temporary
        return
    rows_per = max(1, min(n, _STAGE_BYTES // max(view.element_size() * cols, 1)))
    # two staging buffers: the copy engine fills one (side stream) while the transpose kernel drains the other
    main = torch.cuda.current_stream(device)
    side = _copy_stream(device)
    stages = [torch.empty((rows_per, cols), dtype=view.dtype, device=device) for _ in range(2 if n > rows_per else 1)]
    drained = [None] * len(stages)
    side.wait_stream(main)  # dst / staging allocations and earlier writes are ordered on the main stream
    for idx, r0 in enumerate(range(0, n, rows_per)):
        r1 = min(n, r0 + rows_per)
        i = idx % len(stages)
        with torch.cuda.stream(side):
            if drained[i] is not None:
                side.wait_event(drained[i])
            stages[i][:r1 - r0].copy_(view[r0:r1], non_blocking=True)
            filled = torch.cuda.Event()
            filled.record(side)
        main.wait_event(filled)
        call(tr_in, ptr(stages[i]), cols, r1 - r0, cols, ptr(dst), ld, r0, stream_ptr(device))
        drained[i] = torch.cuda.Event()
        drained[i].record(main)
    main.synchronize()


def try_adopt_feature_major(blocks, n: int, writable: bool):
    """Zero-copy path: the blocks are consecutive column-major CUDA views into one 128-byte aligned
    buffer with a common leading dimension -> return that buffer as the p x ld feature-major matrix,
    or None (-> the caller copies).  Nothing outside the n samples of the views is ever written:

    * ``writable=False`` (predict / transform): the product kernels honour n, so any leading dimension
      works, e.g. the row slice ``X_cm[:n_train]`` of a taller column-major matrix;
    * ``writable=True`` (fit with copy=False, which standardises and deflates in place): the kernels
      stream whole padded features and write them back, so the elements [n, ld) of every feature must be
      padding the caller set aside: ``ld == round_ld(n)`` and already zero (they stay zero).  A view whose
      stride runs over rows of a taller matrix is copied instead.
    """
    if not blocks or not all(isinstance(b, torch.Tensor) and b.is_cuda and b.dtype == F64 and b.dim() == 2 for b in blocks):
        return None
    ld = blocks[0].stride(1)
    if ld % 16 != 0 or ld < n or blocks[0].data_ptr() % 128 != 0:
        return None
    nxt = blocks[0].data_ptr()
    storage = blocks[0].untyped_storage()
    base_storage = storage.data_ptr()
    for b in blocks:
        if b.shape[0] != n or b.shape[1] < 1 or b.stride(0) != 1 or b.stride(1) != ld or b.data_ptr() != nxt \
                or b.untyped_storage().data_ptr() != base_storage:
            return None
        nxt += b.shape[1] * ld * 8
    p = sum(int(b.shape[1]) for b in blocks)
    if nxt - base_storage > storage.nbytes():  # the last feature's tail [n, ld) would lie outside the allocation
        return None
    Xt = blocks[0].as_strided((p, ld), (ld, 1))
    if writable and ld > n:
        if ld != round_ld(n) or bool((Xt[:, n:] != 0).any()):
            return None
    return Xt


_PIN_MIN_BYTES = 1 << 20
TALL_MIN_SAMPLES = 1 << 18   # predict / transform batches at least this long take the TMA-ring product kernel


def to_host(t: torch.Tensor, transpose: bool = False) -> np.ndarray:
    """Device tensor -> numpy array (its transpose when ``transpose``), contiguous in the returned orientation.

    The transposition runs on the device and the copy lands in page-locked memory from PyTorch's caching host allocator,
    so a result array costs one PCIe-rate DMA: no pageable staging inside the driver, no first-touch page faults of a
    fresh allocation, no strided host transposes (657 MB of attributes at the headline size cost 0.3 s that way).  The array
    keeps its pinned block alive; the block returns to the cache when the array is dropped.  Small results, and
    allocations the host cannot pin, take the ordinary pageable copy."""
    if transpose:
        t = t.t()
    t = t.contiguous()
    if not t.is_cuda:
        return t.numpy()
    if t.numel() * t.element_size() >= _PIN_MIN_BYTES:
        try:
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        except RuntimeError:
            h = None
        if h is not None:
            h.copy_(t, non_blocking=True)
            torch.cuda.current_stream(t.device).synchronize()
            return h.numpy()
    return t.cpu().numpy()


def alloc_feature_major(p: int, n: int, device) -> torch.Tensor:
    ld = round_ld(n)
    t = torch.empty((max(p, 1), ld), dtype=F64, device=device)[:p]
    if ld > n:
        t[:, n:].zero_()
    return t


def ingest_blocks(blocks, n: int, shard: ShardMap, device, presharded: bool = False, adopt: bool = False,
                  adopt_writable: bool = True) -> torch.Tensor:
    """blocks: the global blocks (every rank slices its own column range) or, when ``presharded``, this
    rank's local column ranges of them.  ``adopt`` allows the zero-copy path (fit: the caller passed copy=False and the
    matrix is standardised / deflated in place, ``adopt_writable``; predict / transform: read-only)."""
    if adopt and (presharded or shard.world == 1):
        nonempty = [b for b in blocks if b is not None and b.shape[1] > 0]
        Xt = try_adopt_feature_major(nonempty, n, writable=adopt_writable)
        if Xt is not None and Xt.shape[0] == shard.p_local:
            return Xt
    Xt = alloc_feature_major(shard.p_local, n, device)
    for b, blk in enumerate(blocks):
        c0, c1 = shard.local_ranges[b]
        if c1 > c0:
            o0, o1 = shard.block_off[b], shard.block_off[b + 1]
            if presharded:
                ingest_feature_major(blk, n, 0, c1 - c0, Xt[o0:o1], device)
            else:
                ingest_feature_major(blk, n, c0, c1, Xt[o0:o1], device)
    return Xt


# --------------------------------------------------------------------------------------------- #
# preprocessing
# --------------------------------------------------------------------------------------------- #
@dataclass
class ScalerStats:
    mean: torch.Tensor
    var: torch.Tensor
    scale: torch.Tensor
    seen: torch.Tensor   # int64
    zss: torch.Tensor    # nansum(z^2) per feature


def standardize_fit(Xt: torch.Tensor, n: int, mode: int = 0) -> ScalerStats:
    p, ld = Xt.shape
    dev = Xt.device
    st = ScalerStats(*(torch.empty(max(p, 1), dtype=F64, device=dev) for _ in range(3)),
                     torch.empty(max(p, 1), dtype=torch.int64, device=dev),
                     torch.empty(max(p, 1), dtype=F64, device=dev))
    call("mbpls_standardize_fit_f64", ptr(Xt), ld, n, p, ptr(st.mean), ptr(st.var), ptr(st.scale), ptr(st.seen),
         ptr(st.zss), mode, stream_ptr(dev))
    return st


def standardize_apply(Xt: torch.Tensor, n: int, mean: torch.Tensor, scale: torch.Tensor) -> None:
    p, ld = Xt.shape
    call("mbpls_standardize_apply_f64", ptr(Xt), ld, n, p, ptr(mean), ptr(scale), stream_ptr(Xt.device))


@dataclass
class OnePassTables:
    """Split table of the one-pass kernels (csrc/fused.cu): one split per persistent worker, never straddling a block."""
    nsplit: int
    f0: torch.Tensor
    f1: torch.Tensor
    bso: torch.Tensor     # B+1: first split of every block
    blk: torch.Tensor     # block of every split


def onepass_tables(block_off: Sequence[int], ld: int, p: int, device) -> Optional[OnePassTables]:
    nworkers = call("mbpls_fused_total_workers", ld) if p > 0 else 0
    if nworkers <= 0:
        return None
    B = len(block_off) - 1
    of0, of1, obso = make_splits(block_off, 1, nworkers, ctas_per_sm=1, min_feats=16)
    oblk = [b for b in range(B) for _ in range(obso[b + 1] - obso[b])]
    return OnePassTables(len(of0), _i32(of0, device), _i32(of1, device), _i32(obso, device), _i32(oblk, device))


@dataclass
class FirstTrip:
    """What mbpls_fused_standardize_f64 leaves behind for the first trip of the first component."""
    tables: OnePassTables
    w: torch.Tensor
    norm_part: torch.Tensor
    Tnum: torch.Tensor


def standardize_fit_first_trip(Xt: torch.Tensor, n: int, block_off: Sequence[int], u0: torch.Tensor):
    """StandardScaler.fit_transform of X in place AND the complete first trip of the first component (u = u0, the first
    standardised Y column) in one read + one write of X (csrc/fused.cu fused_standardize_kernel).  Returns
    (ScalerStats, FirstTrip), or None when the shape is not covered (features longer than 10,240 samples)."""
    p, ld = Xt.shape
    dev = Xt.device
    if p == 0 or (call("mbpls_fused_uses_clusters", ld) & 1):
        return None
    tb = onepass_tables(block_off, ld, p, dev)
    if tb is None:
        return None
    B = len(block_off) - 1
    st = ScalerStats(*(torch.empty(p, dtype=F64, device=dev) for _ in range(3)), torch.empty(p, dtype=torch.int64, device=dev),
                     torch.empty(p, dtype=F64, device=dev))
    w = torch.empty(p, dtype=F64, device=dev)
    norm_part = torch.zeros(tb.nsplit * B, dtype=F64, device=dev)
    Tnum = torch.zeros((tb.nsplit, ld), dtype=F64, device=dev)
    u0u0 = torch.empty(1, dtype=F64, device=dev)
    call("mbpls_rows_sumsq_f64", ptr(u0), ld, 1, n, ptr(u0u0), stream_ptr(dev))
    call("mbpls_fused_standardize_f64", ptr(Xt), ld, n, ptr(u0), ptr(u0u0), ptr(tb.f0), ptr(tb.f1), ptr(tb.blk), tb.nsplit, B,
         ptr(st.mean), ptr(st.var), ptr(st.scale), ptr(st.seen), ptr(st.zss), ptr(w), ptr(norm_part), ptr(Tnum), ld,
         stream_ptr(dev))
    return st, FirstTrip(tb, w, norm_part, Tnum)


def feature_sumsq(Xt: torch.Tensor, n: int) -> torch.Tensor:
    p, ld = Xt.shape
    out = torch.empty(max(p, 1), dtype=F64, device=Xt.device)
    call("mbpls_feature_sumsq_f64", ptr(Xt), ld, n, p, ptr(out), stream_ptr(Xt.device))
    return out


def segsum(v: torch.Tensor, off_dev: torch.Tensor, nseg: int) -> torch.Tensor:
    out = torch.empty(max(nseg, 1), dtype=F64, device=v.device)
    call("mbpls_segsum_f64", ptr(v), ptr(off_dev), nseg, ptr(out), stream_ptr(v.device))
    return out[:nseg]


def nan_census(Xt: torch.Tensor, n: int, block_off_dev: torch.Tensor, B: int, want_rows: bool = True):
    """Returns (col_nan int32[p], row_flag uint8[B x ldf] or None, inf_flag int32[1]) in one read pass."""
    p, ld = Xt.shape
    dev = Xt.device
    col_nan = torch.zeros(max(p, 1), dtype=torch.int32, device=dev)
    row_flag = torch.zeros((B, ld), dtype=torch.uint8, device=dev) if want_rows else None
    inf_flag = torch.zeros(1, dtype=torch.int32, device=dev)
    call("mbpls_nan_census_f64", ptr(Xt), ld, n, p, ptr(block_off_dev), B, ptr(col_nan), ptr(row_flag), ld,
         ptr(inf_flag), stream_ptr(dev))
    return col_nan[:p], row_flag, inf_flag


def all_finite(Xt: torch.Tensor, n: int) -> bool:
    """check_array's finiteness test (mbpls.py:310,336) as one read pass of the census kernel."""
    p = Xt.shape[0]
    if p == 0 or n == 0:
        return True
    col_nan, _, inf_flag = nan_census(Xt, n, _i32([0, p], Xt.device), 1, want_rows=False)
    return int(col_nan.sum().item()) == 0 and int(inf_flag.item()) == 0


# --------------------------------------------------------------------------------------------- #
# split table for the sample-owning kernels (xw, skinny_gemm)
# --------------------------------------------------------------------------------------------- #
def make_splits(block_off: Sequence[int], n: int, sm_count: int, ctas_per_sm: Optional[int] = None, min_feats: int = 32):
    """Cut every block's local feature range into pieces so that (row chunks x splits) is at most ONE
    wave of resident CTAs (a partial second wave would run alone at the end of the kernel), with the
    pieces as equal as possible (largest-remainder apportionment over the blocks)."""
    if ctas_per_sm is None:
        ctas_per_sm = call("mbpls_xw_ctas_per_sm")
    B = len(block_off) - 1
    sizes = [block_off[b + 1] - block_off[b] for b in range(B)]
    p = sum(sizes)
    row_chunks = max(1, -(-n // 512))
    total = max(1, (sm_count * ctas_per_sm) // row_chunks)
    total = max(1, min(total, max(1, p // min_feats)))
    nonempty = [b for b in range(B) if sizes[b] > 0]
    counts = [0] * B
    if nonempty:
        total = max(total, len(nonempty))
        quota = [sizes[b] * total / p for b in range(B)]
        for b in nonempty:
            counts[b] = max(1, int(quota[b]))
        while sum(counts) > total and any(counts[b] > 1 for b in nonempty):
            b = max((b for b in nonempty if counts[b] > 1), key=lambda b: counts[b] - quota[b])
            counts[b] -= 1
        while sum(counts) < total:
            b = max(nonempty, key=lambda b: quota[b] - counts[b])
            counts[b] += 1
    f0, f1, bso = [], [], [0]
    for b in range(B):
        a, e = block_off[b], block_off[b + 1]
        nb = min(counts[b], max(1, e - a)) if e > a else 0
        for k in range(nb):
            f0.append(a + (e - a) * k // nb)
            f1.append(a + (e - a) * (k + 1) // nb)
        bso.append(len(f0))
    return f0, f1, bso


# --------------------------------------------------------------------------------------------- #
# peer-memory exchange buffers for the in-kernel exchange (csrc/nipals.cu xchg_epilogue_kernel)
# --------------------------------------------------------------------------------------------- #
_SYMM_CACHE: dict = {}


class PeerExchange:
    """One symmetric buffer per rank (torch.distributed._symmetric_memory: CUDA VMM allocations mapped into every process of
    the group over NVLink): 2 slots of `slot_elems` doubles + `world` 8-byte flag words.  `seq` counts the exchanges made
    through it on every rank alike (flags only ever grow, so the buffer is reused from fit to fit)."""

    def __init__(self, group, nred: int, device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.slot_elems = (int(nred) + 15) // 16 * 16
        self.flags_off = 2 * self.slot_elems
        total = self.flags_off + max(16, self.world)
        self.buf = symm_mem.empty(total, dtype=F64, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)  # nobody publishes a flag before every rank has zeroed its buffer
        self.peer_bufs = int(self.hdl.buffer_ptrs_dev)
        self.seq = 0


def peer_exchange(group, nred: int, device):
    """The cached PeerExchange for (group, size), or None when peer memory is not available (-> NCCL all-reduce).
    MBPLS_XCHG=nccl forces the NCCL path."""
    if os.environ.get("MBPLS_XCHG", "").lower() == "nccl":
        return None
    import torch.distributed as dist
    key = (id(group), (int(nred) + 15) // 16 * 16, torch.device(device).index)
    if key not in _SYMM_CACHE:
        ok = 0
        px = None
        try:
            px = PeerExchange(group, nred, device)
            ok = 1
        except Exception as exc:  # no VMM / P2P between these devices, or an older torch: fall back on every rank alike
            _SYMM_CACHE.setdefault("errors", []).append(repr(exc)[:300])
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        _SYMM_CACHE[key] = px if int(flag.item()) == 1 else None
    return _SYMM_CACHE[key]


def allreduce_(t: torch.Tensor, group) -> None:
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(t, group=group)


def norm_kind_of(ord_) -> int:
    """Matrix-norm semantics of ``np.linalg.norm(n x 1 array, ord)`` (mbpls/mbpls.py:887, SURVEY a6')."""
    if ord_ is None or ord_ in (2, -2, "fro", "nuc"):
        return _cabi.NORM_L2
    if ord_ in (1, -1):
        return _cabi.NORM_L1
    if ord_ == np.inf:
        return _cabi.NORM_MAX
    if ord_ == -np.inf:
        return _cabi.NORM_MIN
    raise ValueError("Invalid norm order for matrices.")


# --------------------------------------------------------------------------------------------- #
# NIPALS
# --------------------------------------------------------------------------------------------- #
@dataclass
class NipalsResult:
    Wt: torch.Tensor   # K x p_local   un-normalised block weights (W_non_normal_)
    W: torch.Tensor    # K x p_local   block-normalised weights (W_)
    P: torch.Tensor    # K x p_local   loadings
    Ts: torch.Tensor   # K x ld
    U: torch.Tensor    # K x ld
    Tb: torch.Tensor   # B x K x ld
    V: torch.Tensor    # K x q
    A: torch.Tensor    # K x B  (squared superweights)
    pssb: torch.Tensor  # K x B  local sum of p_j^2 per block
    tt: List[float]
    vv: List[float]
    n_iter: List[int]
    diff: List[float]
    exchange: Optional[str] = None  # multi-GPU: how the per-trip sums crossed the GPUs


def sm_count(device) -> int:
    return torch.cuda.get_device_properties(device).multi_processor_count


def nipals_fit(Xt: torch.Tensor, Yt: torch.Tensor, n: int, block_off: Sequence[int], n_components: int, *,
               u0: torch.Tensor, nanmode: bool = False, row_flag: Optional[torch.Tensor] = None,
               ycol_flag: Optional[torch.Tensor] = None, max_tol: float = 1e-14, norm_kind: int = 0,
               max_iter: int = 1_000_000, group=None, fuse_next_xtu: bool = True, deflate_mode: int = 0,
               trips_per_sync: Optional[int] = None, profile: Optional[dict] = None,
               deflate_last: bool = False, one_pass: Optional[bool] = None,
               one_pass_deflate: Optional[bool] = None, col_nan: Optional[torch.Tensor] = None,
               first_trip: Optional["FirstTrip"] = None, p_widest: Optional[int] = None) -> NipalsResult:
    """Multiblock NIPALS on a (local shard of a) feature-major matrix; deflates ``Xt`` in place.

    Follows mbpls/mbpls.py:821-983; see csrc/nipals.cu for the per-kernel citations.
    """
    dev = Xt.device
    st = stream_ptr(dev)
    p, ld = Xt.shape
    B = len(block_off) - 1
    q = Yt.shape[0]
    K = int(n_components)
    nan = 1 if nanmode else 0
    boff = _i32(block_off, dev)
    f0, f1, bso = make_splits(block_off, n, sm_count(dev))
    nsplit = len(f0)
    sf0, sf1, sbso = _i32(f0, dev), _i32(f1, dev), _i32(bso, dev)
    nparts = call("mbpls_xtu_num_ctas", p) if p > 0 else 0

    def buf(*shape, dtype=F64, zero=False):
        shape = tuple(max(int(s), 1) for s in shape)
        return (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=dev)

    # one-pass kernels (csrc/fused.cu): a trip reads X once; they need the score accumulators of a whole feature in
    # registers, i.e. ld <= 20480 (beyond 10240 a feature is split over the CTA pair of a cluster).  Their "workers" own one
    # split each: the library says how many workers one wave of persistent CTAs holds on this device.
    # NaN-masked data runs through the same kernels with NaN read as zero; every masked denominator is derived from the
    # NaN bit matrix (csrc/nanmask.cu), which needs the per-feature NaN counts of the census (col_nan).
    want_op = one_pass is not False and (not nan or col_nan is not None)
    tables = first_trip.tables if first_trip is not None else (onepass_tables(block_off, ld, p, dev) if want_op else None)
    use_op = tables is not None
    if one_pass is True and not use_op and p > 0:
        raise ValueError("one_pass=True needs a leading dimension of at most 20480 samples (and, for NaN data, the census)")
    use_opd = use_op and one_pass_deflate is not False and deflate_mode == 0
    if use_op:
        nsplit_o, osf0, osf1, osbso, osblk = tables.nsplit, tables.f0, tables.f1, tables.bso, tables.blk
        if first_trip is not None:  # the fused standardisation pass has already run the first trip's pass over X
            Tnum_o, norm_part_o = first_trip.Tnum, first_trip.norm_part
        else:
            Tnum_o = buf(nsplit_o, ld, zero=True)
            norm_part_o = buf(nsplit_o * B, zero=True)
        Tden_o = buf(nsplit_o, ld, zero=True) if nan else None

    w = first_trip.w if first_trip is not None else buf(p)
    norm_part = buf(max(nparts, 1) * B, zero=True)
    Tnum = buf(nsplit, ld, zero=True)
    Tden = buf(nsplit, ld, zero=True) if nan else None
    nred = (2 if nan else 1) * B * ld + B
    red = buf(nred, zero=True)
    # multi-GPU: partials are reduced into red_local and *copied* into red before the all-reduce, so that
    # the no-op trips enqueued after convergence re-create the same global sums instead of re-adding them
    red_local = buf(nred, zero=True) if group is not None else red
    T = buf(B, ld, zero=True)
    u, ts, ts_old = buf(ld, zero=True), buf(ld, zero=True), buf(ld, zero=True)
    a, v = buf(B), buf(q)
    scal = buf(_cabi.SCAL_COUNT, zero=True)
    ctrl = buf(_cabi.CTRL_COUNT, dtype=torch.int32, zero=True)
    pss = buf(p)
    scal_h = torch.empty(_cabi.SCAL_COUNT, dtype=F64).pin_memory()
    ctrl_h = torch.empty(_cabi.CTRL_COUNT, dtype=torch.int32).pin_memory()
    u0u0 = buf(1)
    call("mbpls_rows_sumsq_f64", ptr(u0), ld, 1, n, ptr(u0u0), st)
    bits = rden_u = rden_ts = rden_u0 = None
    ldw = 0
    if nan and use_op:
        # NaN bit matrix (the pattern never changes: deflation keeps NaN, mbpls.py:969) and the masked denominators
        # that do not change either: sum over the observed samples of u0^2 per feature
        ldw = call("mbpls_nan_bitmask_ldw", n)
        bits = torch.empty((p, ldw), dtype=torch.int32, device=dev)
        call("mbpls_nan_bitmask_f64", ptr(Xt), ld, n, p, ptr(bits), ldw, st)
        col_nan = col_nan.to(torch.int32).contiguous()
        rden_u, rden_ts, rden_u0 = buf(p), buf(p), buf(p)
        call("mbpls_masked_colden_f64", ptr(bits), ldw, n, p, ptr(col_nan), ptr(u0), None, ptr(u0u0), 1, ptr(rden_u0), None, st)

    res = NipalsResult(Wt=buf(K, p), W=buf(K, p), P=buf(K, p), Ts=buf(K, ld, zero=True), U=buf(K, ld, zero=True),
                       Tb=buf(B, K, ld, zero=True), V=buf(K, q), A=buf(K, B), pssb=buf(K, B, zero=True),
                       tt=[], vv=[], n_iter=[], diff=[])

    epi = _cabi.EpilogueArgs(n=n, B=B, q=q, nanmode=nan, norm_kind=norm_kind, ldt=ld, ldf=ld, max_tol=float(max_tol),
                             red=red.data_ptr(), Yt=Yt.data_ptr(),
                             row_flag=row_flag.data_ptr() if nan else None,
                             ycol_flag=ycol_flag.data_ptr() if nan else None,
                             T=T.data_ptr(), u=u.data_ptr(), ts=ts.data_ptr(), ts_old=ts_old.data_ptr(),
                             a=a.data_ptr(), v=v.data_ptr(), scal=scal.data_ptr(), ctrl=ctrl.data_ptr(),
                             diff_trace=None, diff_trace_len=0)
    done_p = ptr(ctrl)  # ctrl[MBPLS_CTRL_DONE] is element 0

    # One kernel per trip for "sum the split partials -> sum over the GPUs -> superlevel step" (csrc/nipals.cu
    # xchg_epilogue_kernel): single GPU always; several GPUs when their memory is mapped peer to peer (else NCCL).
    px = peer_exchange(group, nred, dev) if group is not None else None
    fused_epi = os.environ.get("MBPLS_XCHG", "").lower() != "off" and (group is None or px is not None)
    res.exchange = None if group is None else ("peer-memory kernel (NVLink loads, sums in rank order)" if px is not None
                                               else "ncclAllReduce")
    xcount = buf(2, dtype=torch.int32, zero=True)
    xwork = buf(32768, zero=True)  # grid-wide sums of the superlevel step when it runs on all CTAs of the exchange kernel
    xepoch = [0]

    def xchg_args(Tn, Td, ldp, bso_dev, npart_t, nparts_):
        return _cabi.XchgArgs(epi=epi, Tnum=Tn.data_ptr(), Tden=Td.data_ptr() if Td is not None else None, ldp=ldp,
                              block_split_off=bso_dev.data_ptr(), norm_part=npart_t.data_ptr(), n_norm_parts=nparts_,
                              world=px.world if px else 1, rank=px.rank if px else 0, peer_bufs=px.peer_bufs if px else None,
                              slot_elems=px.slot_elems if px else 0, flags_off=px.flags_off if px else 0, seq=0,
                              counters=xcount.data_ptr(), work=xwork.data_ptr())

    x_tp = xchg_args(Tnum, Tden, ld, sbso, norm_part, nparts) if fused_epi else None
    x_op = xchg_args(Tnum_o, Tden_o, ld, osbso, norm_part_o, nsplit_o) if (fused_epi and use_op) else None

    def close_trip(xa, Tn, Td, bso_dev, npart_t, nparts_):
        """split partials -> sums over splits and GPUs -> epilogue"""
        if fused_epi:
            if px is not None:
                px.seq += 1
                xa.seq = px.seq
            xepoch[0] += 1
            xa.epoch = xepoch[0]
            timed("xchg", lambda: call("mbpls_nipals_xchg_epilogue_f64", C.byref(xa), 0, st))
            return
        call("mbpls_nipals_reduce_partials_f64", ptr(Tn), ptr(Td), ld, n, B, ptr(bso_dev), ptr(npart_t), nparts_,
             ptr(red_local), nan, done_p, st)
        if group is not None:
            red.copy_(red_local)
            allreduce_(red, group)
        call("mbpls_nipals_epilogue_f64", C.byref(epi), st)

    if trips_per_sync is None:
        # trips enqueued per readback of the convergence flag; trips launched after convergence are no-ops
        # (a few microseconds each), so batching only removes host round-trips from the critical path
        # (the estimate must be the same on every rank -- shard widths differ -- or the ranks would enqueue different
        # numbers of per-trip all-reduces before the readback: use the widest shard)
        p_est = p
        if p_widest is not None:  # the caller knows the shard map: no collective, no host synchronisation
            p_est = int(p_widest)
        elif group is not None:
            import torch.distributed as dist
            pw = torch.tensor([p], dtype=torch.int64, device=dev)
            dist.all_reduce(pw, op=dist.ReduceOp.MAX, group=group)
            p_est = int(pw.item())
        est_ms = 2.0 * p_est * ld * 8 / 5e12 * 1e3
        trips_per_sync = 2 if est_ms > 1.0 else 4
    # what the closing pass of the previous component left behind for the first trip of the coming one:
    # None, "w" (its weights) or "scores" (weights, squared norms and partial block scores: no pass over X needed)
    w_ready = "scores" if first_trip is not None else None
    cur = torch.cuda.current_stream(dev)

    def timed(key, fn, sink=None):
        """Optionally bracket one launch with CUDA events on the launching stream (bench.py roofline).  `sink`: collect the
        event pair there instead (speculative launches count only if they turn out to have run)."""
        if profile is None:
            fn()
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        fn()
        e1.record(cur)
        (profile.setdefault(key, []) if sink is None else sink).append((e0, e1))

    ready = torch.cuda.Event()
    spec_events: list = []

    def record(k, predicated):
        rec = _cabi.RecordArgs(n=n, p=p, B=B, q=q, nanmode=nan, ldt=ld, T_block_stride=K * ld,
                               block_off=boff.data_ptr(), w=w.data_ptr(), red=red.data_ptr(), T=T.data_ptr(),
                               ts=ts.data_ptr(), u=u.data_ptr(), v=v.data_ptr(), a=a.data_ptr(),
                               Wt_k=res.Wt[k].data_ptr(), W_k=res.W[k].data_ptr(), Ts_k=res.Ts[k].data_ptr(),
                               U_k=res.U[k].data_ptr(), T_k=res.Tb[0, k].data_ptr(), V_k=res.V[k].data_ptr(),
                               A_k=res.A[k].data_ptr(), only_if_done=ctrl.data_ptr() if predicated else None)
        call("mbpls_nipals_record_component_f64", C.byref(rec), st)

    def close_fused(k, fuse, predicated):
        """Bookkeeping + loadings + deflation + the next component's first trip; `predicated`: every launch is a no-op unless the
        component has converged (ctrl[DONE]) -- enqueued behind the trips BEFORE the host knows, so the GPU runs the 3-26 ms
        deflation pass while the host reads the flag back and enqueues the next component instead of idling ~70 us per component
        (gpurun_out/p_timeline_s0125.log: 20 x 69.5 us 'Memcpy DtoH -> record_component')."""
        only = done_p if predicated else None
        record(k, predicated)
        if nan:  # 1 / sum over the observed samples of ts^2 per feature (:923-925); scal[TT] = ts'ts
            call("mbpls_masked_colden_f64", ptr(bits), ldw, n, p, ptr(col_nan), ptr(ts), None, ptr(scal[_cabi.SCAL_TT:]), 0,
                 ptr(rden_ts), None, st)
        timed("deflate", lambda: call("mbpls_fused_deflate_f64", ptr(Xt), ld, n, ptr(ts), ptr(rden_ts),
                                      ptr(u0) if fuse else None, ptr(u0u0) if fuse else None,
                                      ptr(rden_u0) if fuse else None,
                                      ptr(osf0), ptr(osf1), ptr(osblk), nsplit_o, B, ptr(res.P[k]), ptr(pss),
                                      ptr(w) if fuse else None, ptr(norm_part_o) if fuse else None,
                                      ptr(Tnum_o) if fuse else None, ld, only, st), sink=spec_events if predicated else None)
        if nan and fuse:
            call("mbpls_masked_rowden_f64", ptr(bits), ldw, n, ptr(w), ptr(osf0), ptr(osf1), nsplit_o, ptr(Tden_o), ld,
                 None, st)
        # (sums whatever `pss` holds: after a no-op deflation the result is overwritten when the component really closes)
        call("mbpls_segsum_f64", ptr(pss), ptr(boff), B, ptr(res.pssb[k]), st)

    for k in range(K):
        call("mbpls_nipals_begin_component_f64", ptr(u0), n, ptr(u), ptr(scal), ptr(ctrl), st)
        last = (k == K - 1)
        fuse = fuse_next_xtu and not last
        # dense fits on the one-pass kernels close every component but the last speculatively (see close_fused)
        speculate = use_opd and not nan and not (last and not deflate_last) and os.environ.get("MBPLS_SPECULATE", "1") != "0"
        launched = 0
        batch, prev_diff = trips_per_sync, None
        closed = False
        while True:
            for _ in range(max(1, min(batch, max_iter - launched))):
                first = launched == 0
                have_scores = first and w_ready == "scores"  # the deflation pass has already left the first trip's scores
                if have_scores or (use_op and not (first and w_ready == "w")):
                    if not have_scores:
                        if nan:  # 1 / sum over the observed samples of u^2, per feature (mbpls.py:848-852)
                            timed("colden", lambda: call("mbpls_masked_colden_f64", ptr(bits), ldw, n, p, ptr(col_nan), ptr(u), None,
                                                         ptr(scal), 1, ptr(rden_u), done_p, st))
                        timed("trip", lambda: call("mbpls_nipals_fused_trip_f64", ptr(Xt), ld, n, ptr(u), ptr(scal), ptr(rden_u),
                                                   ptr(osf0), ptr(osf1), ptr(osblk), nsplit_o, B, ptr(w), ptr(norm_part_o),
                                                   ptr(Tnum_o), ld, done_p, st))
                        if nan:  # sum over the observed features of w~^2, per sample and split (:867-872)
                            timed("rowden", lambda: call("mbpls_masked_rowden_f64", ptr(bits), ldw, n, ptr(w), ptr(osf0), ptr(osf1),
                                                         nsplit_o, ptr(Tden_o), ld, done_p, st))
                    close_trip(x_op, Tnum_o, Tden_o, osbso, norm_part_o, nsplit_o)
                else:
                    if first and w_ready == "w":
                        call("mbpls_block_sumsq_parts_f64", ptr(w), p, ptr(boff), B, ptr(norm_part), done_p, st)
                    else:
                        timed("xtu", lambda: call("mbpls_nipals_xtu_f64", ptr(Xt), ld, n, p, ptr(u), ptr(scal), ptr(boff), B,
                                                  ptr(w), ptr(norm_part), nan, done_p, st))
                    timed("xw", lambda: call("mbpls_nipals_xw_f64", ptr(Xt), ld, n, ptr(w), ptr(sf0), ptr(sf1), nsplit,
                                             ptr(Tnum), ptr(Tden), ld, nan, done_p, st))
                    close_trip(x_tp, Tnum, Tden, sbso, norm_part, nparts)
                launched += 1
            ctrl_h.copy_(ctrl, non_blocking=True)
            scal_h.copy_(scal, non_blocking=True)
            if speculate:
                ready.record(cur)
                close_fused(k, fuse, True)
                ready.synchronize()
            else:
                cur.synchronize()
            if int(ctrl_h[_cabi.CTRL_ERROR]):
                raise _cabi.MbplsCudaError("a peer GPU did not arrive at the in-kernel exchange of a NIPALS trip (timeout)")
            if int(ctrl_h[_cabi.CTRL_DONE]):
                closed = speculate
                break
            if int(ctrl_h[_cabi.CTRL_TRIPS]) >= max_iter:
                break
            # PLS2 loops converge geometrically: from the last two readings of diff_t estimate how many trips remain and enqueue
            # most of them before the next readback (trips past convergence are no-ops, so an over-estimate costs microseconds;
            # diff_t is bit-identical on every rank, so every rank enqueues the same number of exchanges)
            d = float(scal_h[_cabi.SCAL_DIFF])
            batch = trips_per_sync
            if prev_diff is not None and 0.0 < d < prev_diff and d > max_tol > 0.0:
                rate = (d / prev_diff) ** (1.0 / max(1, last_batch))
                if 0.0 < rate < 0.999:
                    batch = int(min(64, max(trips_per_sync, math.log(max_tol / d) / math.log(rate) - 1)))
            prev_diff, last_batch = d, batch
        res.n_iter.append(int(ctrl_h[_cabi.CTRL_TRIPS]))
        res.diff.append(float(scal_h[_cabi.SCAL_DIFF]))
        res.tt.append(float(scal_h[_cabi.SCAL_TT]))
        res.vv.append(float(scal_h[_cabi.SCAL_VV]))
        if closed:
            if profile is not None and spec_events:
                profile.setdefault("deflate", []).append(spec_events[-1])  # the launch that found the flag set
            spec_events.clear()
            w_ready = "scores" if fuse else None
            continue
        spec_events.clear()
        if last and not deflate_last:
            # the deflated X of the last component is never read (mbpls.py:968-969 is followed by the end of the
            # loop), so only the loadings are computed: one read instead of read + write
            record(k, False)
            timed("loadings", lambda: call("mbpls_nipals_xtu_f64", ptr(Xt), ld, n, p, ptr(ts), None, ptr(boff), B,
                                           ptr(res.P[k]), None, nan, None, st))
            if p > 0:
                call("mbpls_block_sumsq_f64", ptr(res.P[k]), ptr(boff), B, ptr(res.pssb[k]), st)
        elif use_opd:
            close_fused(k, fuse, False)  # NaN data, or the max_iter cap was hit before convergence
            w_ready = "scores" if fuse else None
        else:
            record(k, False)
            timed("deflate", lambda: call("mbpls_loadings_deflate_f64", ptr(Xt), ld, n, p, ptr(ts),
                                          ptr(u0) if fuse else None, ptr(u0u0) if fuse else None, ptr(res.P[k]),
                                          ptr(w) if fuse else None, ptr(pss), nan, deflate_mode, st))
            if p > 0:
                call("mbpls_segsum_f64", ptr(pss), ptr(boff), B, ptr(res.pssb[k]), st)
            w_ready = "w" if fuse else None
    return res


# --------------------------------------------------------------------------------------------- #
# small contractions over the feature axis
# --------------------------------------------------------------------------------------------- #
def gram(A: torch.Tensor, Bm: torch.Tensor, p: int, group=None) -> torch.Tensor:
    """A (K1 x p) . Bm (K2 x p)^T -> K1 x K2, deterministic two-stage reduction (+ all-reduce)."""
    dev = A.device
    K1, K2 = A.shape[0], Bm.shape[0]
    out = torch.zeros((K1, K2), dtype=F64, device=dev)
    if p > 0:
        nch = call("mbpls_gram_num_chunks", p)
        part = torch.empty((nch, K1 * K2), dtype=F64, device=dev)
        call("mbpls_gram_partial_f64", ptr(A), A.stride(0), K1, ptr(Bm), Bm.stride(0), K2, p, ptr(part), stream_ptr(dev))
        call("mbpls_reduce_chunks_f64", ptr(part), nch, K1 * K2, ptr(out), stream_ptr(dev))
    allreduce_(out, group)
    return out


def rows_sumsq(M: torch.Tensor, n: int, group=None) -> torch.Tensor:
    rows = M.shape[0]
    out = torch.zeros(max(rows, 1), dtype=F64, device=M.device)
    if n > 0:
        call("mbpls_rows_sumsq_f64", ptr(M), M.stride(0), rows, n, ptr(out), stream_ptr(M.device))
    allreduce_(out, group)
    return out[:rows]


def right_multiply(inp: torch.Tensor, p: int, rowscale: Optional[torch.Tensor], M: torch.Tensor) -> torch.Tensor:
    """out[c][j] = sum_k inp[k][j] * rowscale[k] * M[k][c]  ->  C x p."""
    K, Cc = M.shape
    out = torch.empty((Cc, max(p, 1)), dtype=F64, device=inp.device)
    M = M.contiguous()
    call("mbpls_right_multiply_f64", ptr(inp), inp.stride(0), K, p, ptr(rowscale), ptr(M), Cc, ptr(out), out.stride(0),
         stream_ptr(inp.device))
    return out[:, :p]


_SKINNY_SPLITS: dict = {}


def skinny_gemm(Xt: torch.Tensor, n: int, Bm: torch.Tensor, block_off: Sequence[int], group=None,
                mean: Optional[torch.Tensor] = None, scale: Optional[torch.Tensor] = None,
                flag: Optional[torch.Tensor] = None, dense: bool = False) -> torch.Tensor:
    """out[c][i] = sum_j nan0(z_ij) * Bm[c][j]  ->  C x ld (feature-major result); z = Xt, or the standardised
    value (Xt - mean_j) / scale_j computed on the fly when mean/scale are given.  dense: the caller guarantees that Xt holds no
    NaN (fit-time products of SIMPLS / UNIPALS / KERNEL), which opens the faster one-output route below."""
    dev = Xt.device
    p, ld = Xt.shape
    Cc = Bm.shape[0]
    out = torch.zeros((Cc, ld), dtype=F64, device=dev)
    if scale is not None and p > 0:
        # (x - mean) / scale . b  ==  (x - mean) . (b / scale): the division moves from every element of X to the C x p
        # coefficients (an fp64 divide per element halves the rate of this memory-bound pass)
        Bm = (Bm[:, :p] / scale[:p].view(1, -1)).contiguous()
        scale = None
    if p > 0 and n >= TALL_MIN_SAMPLES and Cc <= 32 and scale is None and ld % 16 == 0 and os.environ.get("MBPLS_TALL", "1") != "0":
        # tall batch: persistent CTAs fed by a TMA ring, results written directly (csrc/finalize.cu skinny_tall_kernel), four
        # outputs per pass (Ts = X R with 30 components on 1 M samples: 8 passes at 5.9 TB/s = 22 ms, against 2 x 22 ms at
        # 0.7 TB/s through the per-thread-load kernel with 16 outputs per thread)
        for c0 in range(0, Cc, 4):
            c1 = min(Cc, c0 + 4)
            coef = torch.zeros((p, 8), dtype=F64, device=dev)  # row j: {mean_j, b_0j .. b_3j, -, -, -}: rides through the ring with feature j
            if mean is not None:
                coef[:, 0] = mean[:p]
            coef[:, 1:1 + c1 - c0] = Bm[c0:c1, :p].t()
            call("mbpls_skinny_gemm_tall_f64", ptr(Xt), ld, n, p, ptr(coef), c1 - c0, ptr(out[c0]), ld, ptr(flag), stream_ptr(dev))
        allreduce_(out, group)
        return out
    if p > 0 and n > 0:
        # the split table depends on the shapes only: built (two small uploads) once per shape, not once per product --
        # SIMPLS / UNIPALS make two of these products per component and are bound by the host's launch rate at C2
        key = (tuple(int(o) for o in block_off), int(n), dev.index)
        tab = _SKINNY_SPLITS.get(key)
        if tab is None:
            f0, f1, _ = make_splits(block_off, n, sm_count(dev))
            if len(_SKINNY_SPLITS) > 64:
                _SKINNY_SPLITS.clear()
            tab = _SKINNY_SPLITS[key] = (len(f0), _i32(f0, dev), _i32(f1, dev))
        ns, sf0, sf1 = tab
        if dense and Cc == 1 and mean is None and scale is None and flag is None and os.environ.get("MBPLS_SKINNY_XW", "1") != "0":
            # one output, plain product (t = X r of SIMPLS, X X'y of UNIPALS): the NIPALS X w kernel streams it at 6.7 TB/s where
            # the kernel below, laid out for several outputs per thread, reaches 2.3 TB/s (5,000 x 50,000); same split table,
            # the split partials are then added in split order
            part = torch.zeros((ns, ld), dtype=F64, device=dev)
            w1 = Bm[0, :p].contiguous()
            call("mbpls_nipals_xw_f64", ptr(Xt), ld, n, ptr(w1), ptr(sf0), ptr(sf1), ns, ptr(part), None, ld, 0, None, stream_ptr(dev))
            call("mbpls_reduce_chunks_f64", ptr(part), ns, ld, ptr(out), stream_ptr(dev))
            allreduce_(out, group)
            return out
        part = torch.zeros((ns, Cc * ld), dtype=F64, device=dev)
        call("mbpls_skinny_gemm_f64", ptr(Xt), ld, n, ptr(Bm), Bm.stride(0), Cc, ptr(sf0), ptr(sf1),
             ns, ptr(part), ld, ptr(mean), ptr(scale), ptr(flag), stream_ptr(dev))
        call("mbpls_reduce_chunks_f64", ptr(part), ns, Cc * ld, ptr(out), stream_ptr(dev))
    allreduce_(out, group)
    return out


# --------------------------------------------------------------------------------------------- #
# small dense linear algebra on the device (csrc/smalllin.cu)
# --------------------------------------------------------------------------------------------- #
def small_top_eigvec(G: torch.Tensor) -> torch.Tensor:
    """Unit top eigenvector of a symmetric PSD m x m matrix (m <= 64)."""
    m = G.shape[0]
    G = G.contiguous()
    out = torch.empty(m, dtype=F64, device=G.device)
    call("mbpls_small_top_eigvec_f64", ptr(G), G.stride(0), m, ptr(out), stream_ptr(G.device))
    return out


def small_pinv(M: torch.Tensor, rcond: float = 1e-15) -> torch.Tensor:
    """Moore-Penrose pseudo-inverse of an m x m matrix (numpy.linalg.pinv semantics, m <= 64)."""
    m = M.shape[0]
    M = M.contiguous()
    out = torch.empty((m, m), dtype=F64, device=M.device)
    call("mbpls_small_pinv_f64", ptr(M), M.stride(0), m, float(rcond), ptr(out), out.stride(0), stream_ptr(M.device))
    return out


def small_top_sv_product(G: torch.Tensor, H: torch.Tensor) -> torch.Tensor:
    """c such that A c is the top left singular vector of A B', from G = B'B and H = A'A."""
    m = G.shape[0]
    G, H = G.contiguous(), H.contiguous()
    out = torch.empty(m, dtype=F64, device=G.device)
    call("mbpls_small_top_sv_product_f64", ptr(G), G.stride(0), ptr(H), H.stride(0), m, ptr(out), stream_ptr(G.device))
    return out
