"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink) collectives.

The reference never issues a collective; its data-movement primitives are "gather all blocks
of an axis into one task" (axis_partition.apply, axpart.py:445-452), "hand one future to many
tasks" (pm.base_broadcast_apply, pm.py:443-494) and the range-partition shuffle
(pm.shuffle_partitions, pm.py:1937-2052).  Their B200 equivalents:

* TreeReduce combine   -> ``all_reduce`` of the W-vector partial (sum / min / max);
* GroupByReduce reduce -> range-partitioned all-to-all of the PRE-AGGREGATED partial tables
  (<= G rows per GPU, never raw rows), pivots picked TeraSort-style from samples of the
  per-rank sorted keys (cf. ShuffleSortFunctions, dfutils.py:163-332);
* broadcast merge      -> ``all_gather`` of the dim shard columns.

Rows are sharded across ranks (rank r owns a contiguous row range); Map / Binary never
communicate.  Everything here works on plain torch tensors so the same code runs under
``gloo`` on CPU tensors (tests, world_size 2) and ``nccl`` on device tensors.
"""

from __future__ import annotations

import os
from typing import List, Sequence

_T = None


def _torch():
    global _T
    if _T is None:
        import torch

        _T = torch
    return _T


def _dist():
    import torch.distributed as dist

    return dist


_LOCAL_DEPTH = 0


class local_frames:
    """Context manager: frames created and used inside belong to THIS rank alone (not row shards of a job-wide
    frame), so ingest does not cut them by rank and nothing issues a collective -- N ranks then run N independent
    single-GPU pipelines (the per-rank host-to-host legs of ``bench.py``).  Every rank must leave the block before
    the next job-wide operation."""

    def __enter__(self):
        global _LOCAL_DEPTH
        _LOCAL_DEPTH += 1
        return self

    def __exit__(self, *exc):
        global _LOCAL_DEPTH
        _LOCAL_DEPTH -= 1
        return False


def is_distributed() -> bool:
    d = _dist()
    return _LOCAL_DEPTH == 0 and d.is_available() and d.is_initialized() and d.get_world_size() > 1


def world_size() -> int:
    d = _dist()
    return d.get_world_size() if d.is_available() and d.is_initialized() else 1


def rank() -> int:
    d = _dist()
    return d.get_rank() if d.is_available() and d.is_initialized() else 0


def init_from_env(backend: str | None = None) -> bool:
    """Initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / MASTER_*).
    Returns True when running with world_size > 1."""
    t = _torch()
    d = _dist()
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws <= 1:
        return False
    if not d.is_initialized():
        if backend is None:
            backend = "nccl" if t.cuda.is_available() else "gloo"
        if backend == "nccl":
            local = int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0")))
            t.cuda.set_device(local)
            d.init_process_group(backend, device_id=t.device("cuda", local))
        else:
            d.init_process_group(backend)
    return True


def barrier():
    if is_distributed():
        _dist().barrier()


_CONTROL = None


def _control_group():
    """Process group for HOST-side integers every rank must agree on (is the job-wide frame empty? how many rows
    are there?): gloo over CPU tensors, so that asking never waits for the GPU -- a NCCL all_reduce + ``.item()``
    would drain the rank's stream and serialise an otherwise asynchronous pipeline.  Under a gloo job (CPU tests) it
    is the default group.  Created collectively on first use (every rank reaches the first use together: SPMD)."""
    global _CONTROL
    d = _dist()
    if _CONTROL is None:
        _CONTROL = d.new_group(backend="gloo") if d.get_backend() != "gloo" else d.group.WORLD
    return _CONTROL


def control_sum(value: int) -> int:
    """Sum of one host integer over the ranks, on the control group (no device synchronisation)."""
    if not is_distributed():
        return int(value)
    t = _torch().tensor([int(value)], dtype=_torch().int64)
    _dist().all_reduce(t, group=_control_group())
    return int(t.item())


def shard_bounds(nrows: int, r: int | None = None, ws: int | None = None):
    """Contiguous row range [lo, hi) owned by rank r: even split, remainder to the low ranks."""
    ws = world_size() if ws is None else ws
    r = rank() if r is None else r
    base, rem = divmod(int(nrows), ws)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


_REDUCE_OPS = {"sum": "SUM", "min": "MIN", "max": "MAX"}

# ---- data-plane collectives through the C ABI (csrc/comm.cu: NCCL resolved at run time, this library's own
# communicator, issued on the rank's CUDA stream).  torch.distributed stays the CONTROL plane: process bootstrap, the
# 128-byte unique id, and the handful of host-side integers a decision needs; under gloo (CPU tests) it also carries
# the data.
_COMM = None
_COMM_OPS = {"sum": 0, "min": 1, "max": 2}


def _native():
    """The mb200_comm handle when the job runs on NCCL (created on first use), else None."""
    global _COMM
    d = _dist()
    if not (d.is_available() and d.is_initialized()) or d.get_backend() != "nccl":
        return None
    if _COMM is None:
        import ctypes as C

        from . import _lib

        lib = _lib.load()
        path = None
        try:
            import nvidia.nccl as _n  # the copy PyTorch ships (already mapped into the process)

            cand = os.path.join(os.path.dirname(_n.__file__), "lib", "libnccl.so.2")
            path = cand.encode() if os.path.exists(cand) else None
        except Exception:
            pass
        _lib.check(lib.mb200_comm_load(path))
        uid = (C.c_char * 128)()
        payload = [None]
        if d.get_rank() == 0:
            _lib.check(lib.mb200_comm_unique_id(uid))
            payload = [bytes(uid)]
        d.broadcast_object_list(payload, src=0)
        handle = C.c_void_p()
        _lib.check(lib.mb200_comm_init_rank(C.byref(handle), d.get_world_size(), payload[0], d.get_rank()))
        _COMM = (lib, handle)
    return _COMM


def _code(tensor):
    t = _torch()
    if tensor.dtype == t.float64:
        return 0
    if tensor.dtype == t.int64:
        return 1
    if tensor.dtype in (t.uint8, t.bool):
        return 2
    raise TypeError(f"collectives of the B200 path move float64 / int64 / uint8 buffers, not {tensor.dtype}")


def _stream():
    return _torch().cuda.current_stream().cuda_stream


def all_reduce_inplace(tensor, op: str) -> None:
    """Element-wise all_reduce of one (large) device array in place -- the cross-GPU merge of dense
    group tables: NVSwitch carries 2 (W-1)/W x the table once, no keys move."""
    if is_distributed() and tensor.numel():
        nat = _native()
        if nat is not None:
            from . import _lib

            _lib.check(nat[0].mb200_comm_allreduce(nat[1], tensor.data_ptr(), tensor.data_ptr(), tensor.numel(),
                                                   _code(tensor), _COMM_OPS[op], _stream()))  # fmt: skip
            return
        d = _dist()
        d.all_reduce(tensor, op=getattr(d.ReduceOp, _REDUCE_OPS[op]))


def all_gather_small(values, extra: int | None = None):
    """all_gather of a short int64 device vector (+ one host integer appended) -> list of per-rank
    Python int lists.  One collective and one D2H for the handful of scalars a decision needs."""
    t = _torch()
    d = _dist()
    v = values.reshape(-1)
    if extra is not None:
        v = t.cat([v, t.tensor([int(extra)], dtype=v.dtype, device=v.device)])
    out = t.empty(world_size() * v.numel(), dtype=v.dtype, device=v.device)
    d.all_gather_into_tensor(out, v.contiguous())
    return out.reshape(world_size(), v.numel()).tolist()


def all_gather_fixed(vec):
    """all_gather of a device vector that has the SAME length on every rank -> one flat device vector, rank-major
    (no host round trip: the carries of a cumulative function, W numbers per rank)."""
    t = _torch()
    v = vec.reshape(-1).contiguous()
    out = t.empty(world_size() * v.numel(), dtype=v.dtype, device=v.device)
    if not v.numel():
        return out
    nat = _native()
    if nat is not None:
        from . import _lib

        _lib.check(nat[0].mb200_comm_allgather(nat[1], v.data_ptr(), out.data_ptr(), v.numel(), _code(v), _stream()))
        return out
    _dist().all_gather_into_tensor(out, v)
    return out


def exclusive_row_offset(nrows_local: int) -> int:
    """Global position of this rank's first row when every rank holds ``nrows_local`` rows in rank order."""
    if not is_distributed():
        return 0
    t = _torch()
    dev = "cuda" if _dist().get_backend() == "nccl" else "cpu"
    counts = all_gather_small(t.tensor([int(nrows_local)], dtype=t.int64, device=dev))
    return int(sum(c[0] for c in counts[: rank()]))


def dense_split(nkeys: int, r: int | None = None, ws: int | None = None):
    """Slice [lo, hi) of a dense key range [0, nkeys) that rank r emits: equal chunks rounded up to a
    multiple of 4 (the presence map is scanned in 4-byte words); trailing ranks may get (0, 0)."""
    ws = world_size() if ws is None else ws
    r = rank() if r is None else r
    chunk = -(-int(nkeys) // ws)
    chunk = (chunk + 3) & ~3
    lo = r * chunk
    if lo >= nkeys:
        return 0, 0
    return lo, min(lo + chunk, int(nkeys))


def dense_chunk(nkeys: int, ws: int | None = None) -> int:
    """Keys per rank when a dense key range of ``nkeys`` is reduce-scattered: equal chunks, a multiple of 4 (the
    presence map is scanned in 4-byte words); ``ws * chunk >= nkeys`` (the table is padded to that)."""
    ws = world_size() if ws is None else ws
    chunk = -(-int(nkeys) // ws)
    return max(4, (chunk + 3) & ~3)


def reduce_scatter(out, inp, op: str) -> None:
    """``out`` <- this rank's equal slice of the element-wise reduction of ``inp`` over the ranks (NCCL
    reduce-scatter over NVLink: each rank receives (W-1)/W of ONE slice instead of the all_reduce's whole array)."""
    nat = _native()
    if nat is not None:
        from . import _lib

        _lib.check(nat[0].mb200_comm_reduce_scatter(nat[1], inp.data_ptr(), out.data_ptr(), out.numel(), _code(out),
                                                    _COMM_OPS[op], _stream()))  # fmt: skip
        return
    d = _dist()
    d.reduce_scatter_tensor(out, inp, op=getattr(d.ReduceOp, _REDUCE_OPS[op]))


def all_reduce_values(tensors: Sequence, ops: Sequence[str]) -> None:
    """In-place all_reduce of many 1-element tensors: one collective per (dtype, op) bucket.

    TreeReduce combine: W x 8 B per bucket -- latency-bound, so the W scalars are packed."""
    if not is_distributed() or not tensors:
        return
    t = _torch()
    d = _dist()
    buckets = {}
    for i, (x, op) in enumerate(zip(tensors, ops)):
        buckets.setdefault((x.dtype, op), []).append(i)
    for (dtype, op), idxs in buckets.items():
        packed = t.cat([tensors[i].reshape(-1) for i in idxs])
        all_reduce_inplace(packed, op)
        off = 0
        for i in idxs:
            n = tensors[i].numel()
            tensors[i].copy_(packed[off : off + n].reshape(tensors[i].shape))
            off += n


def all_gather_rows(cols: Sequence) -> List:
    """all_gather of row shards (different lengths per rank) -> concatenated columns."""
    if not is_distributed():
        return list(cols)
    t = _torch()
    d = _dist()
    ws = world_size()
    n_local = int(cols[0].shape[0]) if cols else 0
    dev = cols[0].device if cols else "cpu"
    counts = t.zeros(ws, dtype=t.int64, device=dev)
    counts[rank()] = n_local
    d.all_reduce(counts)
    counts = [int(c) for c in counts.tolist()]
    out = []
    for c in cols:
        pieces = [t.empty(k, dtype=c.dtype, device=c.device) for k in counts]
        m = max(counts)
        # all_gather needs equal shapes: pad to the longest shard
        padded = t.zeros(m, dtype=c.dtype, device=c.device)
        padded[:n_local] = c
        nat = _native()
        if nat is not None and m:
            from . import _lib

            flat = t.empty(ws * m, dtype=c.dtype, device=c.device)
            _lib.check(nat[0].mb200_comm_allgather(nat[1], padded.data_ptr(), flat.data_ptr(), m, _code(padded), _stream()))
            gathered = [flat[r * m : (r + 1) * m] for r in range(ws)]
        else:
            gathered = [t.empty(m, dtype=c.dtype, device=c.device) for _ in range(ws)]
            d.all_gather(gathered, padded)
        pieces = [g[:k] for g, k in zip(gathered, counts)]
        out.append(t.cat(pieces))
    return out


def _all_to_all_rows(recv, packed, recv_rows, send_rows) -> None:
    """all_to_all of a row-major [rows, ncols] int64 matrix split by rows: ``mb200_comm_alltoallv`` (grouped
    ncclSend / ncclRecv on the rank's stream) under NCCL, ``all_to_all_single`` under gloo."""
    nat = _native()
    if nat is None:
        _dist().all_to_all_single(recv, packed, output_split_sizes=list(recv_rows), input_split_sizes=list(send_rows))
        return
    import ctypes as C

    from . import _lib

    n = len(send_rows)

    def arr(vals):
        return (C.c_int64 * n)(*[int(v) for v in vals])

    sd = [sum(send_rows[:i]) for i in range(n)]
    rd = [sum(recv_rows[:i]) for i in range(n)]
    row_bytes = int(packed.shape[1]) * 8 if packed.dim() == 2 else 8
    _lib.check(nat[0].mb200_comm_alltoallv(nat[1], packed.data_ptr(), arr(send_rows), arr(sd), recv.data_ptr(),
                                           arr(recv_rows), arr(rd), row_bytes, _stream()))  # fmt: skip


def exchange_rows(columns: Sequence, send_counts: Sequence[int]) -> List:
    """all_to_all of raw rows: ``columns`` are 1-D tensors of equal length (8-byte dtypes) laid out destination by
    destination (``send_counts[r]`` consecutive rows go to rank r).  Returns the received columns: the pieces of
    rank 0, 1, ... in that order.  One collective for the counts, one for the packed rows."""
    if not is_distributed():
        return list(columns)
    t = _torch()
    d = _dist()
    dev = columns[0].device
    sc = t.as_tensor([int(c) for c in send_counts], dtype=t.int64).to(dev)
    rc = t.empty_like(sc)
    d.all_to_all_single(rc, sc)
    rcl = [int(x) for x in rc.tolist()]
    scl = [int(c) for c in send_counts]
    n = int(columns[0].shape[0])
    packed = t.empty((n, len(columns)), dtype=t.int64, device=dev)
    for j, c in enumerate(columns):
        packed[:, j] = c.view(t.int64) if c.dtype != t.int64 else c
    recv = t.empty((sum(rcl), len(columns)), dtype=t.int64, device=dev)
    _all_to_all_rows(recv, packed, rcl, scl)
    outs = []
    for j, c in enumerate(columns):
        col = recv[:, j].contiguous()
        outs.append(col.view(c.dtype) if c.dtype != t.int64 else col)
    return outs


def choose_pivots(sorted_keys, nsamples: int = 256):
    """ws-1 range pivots, identical on every rank, from evenly spaced samples of each rank's
    ascending unique keys (TeraSort-style; reference: pick_pivots_from_samples_for_sort,
    dfutils.py:288-332).  Keys k go to destination  #(pivots <= k)."""
    t = _torch()
    d = _dist()
    ws = world_size()
    g = int(sorted_keys.shape[0])
    dev = sorted_keys.device
    samp = t.full((nsamples,), t.iinfo(t.int64).max, dtype=t.int64, device=dev)
    k = min(nsamples, g)
    if k > 0:
        pos = (t.arange(k, device=dev, dtype=t.float64) + 0.5) * (g / k)
        samp[:k] = sorted_keys[pos.to(t.int64).clamp_(max=g - 1)]
    cnt = t.tensor([k], dtype=t.int64, device=dev)
    all_s = [t.empty_like(samp) for _ in range(ws)]
    all_c = [t.empty_like(cnt) for _ in range(ws)]
    d.all_gather(all_s, samp)
    d.all_gather(all_c, cnt)
    pool = t.cat([s[: int(c.item())] for s, c in zip(all_s, all_c)])
    if pool.numel() == 0:
        return t.zeros(ws - 1, dtype=t.int64, device=dev)
    pool, _ = t.sort(pool)
    q = (t.arange(1, ws, device=dev, dtype=t.float64) * (pool.numel() / ws)).to(t.int64).clamp_(max=pool.numel() - 1)
    return pool[q]


def exchange_by_key_range(sorted_keys, columns: Sequence):
    """Range-partitioned all-to-all of a key-sorted partial table.

    ``sorted_keys``: int64 [g] ascending; ``columns``: tensors [g] (8-byte dtypes).  Every rank
    receives, for its key range, the segments of all ranks (each ascending), concatenated in
    rank order.  Returns (keys, columns).  Message size <= g * 8 * (1 + len(columns)) bytes."""
    if not is_distributed():
        return sorted_keys, list(columns)
    t = _torch()
    d = _dist()
    ws = world_size()
    dev = sorted_keys.device
    pivots = choose_pivots(sorted_keys)
    # destination of key k = number of pivots <= k  -> split points via searchsorted(right=False on pivots)
    cuts = t.searchsorted(sorted_keys, pivots, right=False)  # first index with key >= pivot
    bounds = t.cat([t.zeros(1, dtype=t.int64, device=dev), cuts.to(t.int64),
                    t.tensor([sorted_keys.shape[0]], dtype=t.int64, device=dev)])  # fmt: skip
    send_counts = (bounds[1:] - bounds[:-1]).to(t.int64)
    recv_counts = t.empty_like(send_counts)
    d.all_to_all_single(recv_counts, send_counts)
    sc = [int(x) for x in send_counts.tolist()]
    rc = [int(x) for x in recv_counts.tolist()]
    ncols = 1 + len(columns)
    # pack keys + columns as int64 bit patterns: one collective for the whole table
    packed = t.empty((int(sorted_keys.shape[0]), ncols), dtype=t.int64, device=dev)
    packed[:, 0] = sorted_keys
    for j, c in enumerate(columns):
        packed[:, j + 1] = c.view(t.int64) if c.dtype != t.int64 else c
    recv = t.empty((sum(rc), ncols), dtype=t.int64, device=dev)
    _all_to_all_rows(recv, packed, rc, sc)
    keys = recv[:, 0].contiguous()
    outs = []
    for j, c in enumerate(columns):
        col = recv[:, j + 1].contiguous()
        outs.append(col.view(c.dtype) if c.dtype != t.int64 else col)
    return keys, outs
