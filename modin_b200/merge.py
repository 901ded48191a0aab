"""Broadcast merge shared by both front doors (the Modin plug-in's query compiler and the mirror's).

Reference: ``qc.merge`` (qc.py:657-667) -> ``MergeImpl.row_axis_merge`` (storage_formats/pandas/merge.py:104-252): the
right frame is collapsed into ONE partition (``combine``, merge.py:178), broadcast to every row partition of the left
frame (``broadcast_apply_full_axis``) where ``pandas.merge(left_block, right, ...)`` runs, and the index is reset
(merge.py:236-250).  Here the per-block function is ``functors.DevMerge``; what the reference recomputes on every call
is kept on the (immutable) right frame instead:

* the combined right block -- across GPUs an ``all_gather`` of the dim shards (``partitioning.gather_block``);
* the join table built over its key column, with the key-ordered payload copies the library keeps next to a dense
  table -- built once, probed by every later ``merge`` against the same dim frame.
"""

from __future__ import annotations

import numpy as np
import pandas

from . import dist
from .functors import DevMerge, DevMergePacked


def _combined(right_frame):
    """The right frame as one broadcastable partition, cached on the frame object (frames are immutable values)."""
    comb = getattr(right_frame, "_b200_combined", None)
    if comb is None:
        comb = right_frame.combine()
        comb._b200_join_tables = {}
        try:
            right_frame._b200_combined = comb
        except AttributeError:
            pass
    return comb


def _row_blocks(frame):
    from .block import concat_cols

    return [concat_cols([p.get() for p in row]) if len(row) > 1 else row[0].get() for row in frame._partitions]


def _packed_right(left_frame, comb, left_on, right_on):
    """(right frame in packed form, packing plan) for a join on several int64 key columns.  The plan spans the key
    ranges of the left row blocks AND the right block (``groupkeys.packing_plan``: cached column statistics, agreed
    across ranks), so every key tuple of either side has an image and equal tuples have equal images.  The packed
    right frame holds the right columns that enter the result -- a right key column whose LABEL equals its left
    partner's is the same result column and is left out, as pandas does -- plus the image; it is kept per (keys, plan)
    on the combined right frame, with its join table."""
    from . import groupkeys as gk
    from .block import DeviceBlock

    rblock = comb._partitions[0, 0].get() if comb._partitions.shape == (1, 1) else _row_blocks(comb)[0]
    lkeys = [DeviceBlock([b.column(k) for k in left_on], pandas.Index(range(len(left_on))), nrows=b.nrows)
             for b in _row_blocks(left_frame)]  # fmt: skip
    rkeys = DeviceBlock([rblock.column(k) for k in right_on], pandas.Index(range(len(right_on))), nrows=rblock.nrows)
    plan = gk.packing_plan(lkeys + [rkeys])
    cache = comb.__dict__.setdefault("_b200_packed", {})
    ident = (tuple(left_on), tuple(right_on), tuple(map(tuple, plan)))
    hit = cache.get(ident)
    if hit is None:
        same = {r for l, r in zip(left_on, right_on) if l == r}
        keep = [i for i, lab in enumerate(rblock.columns) if lab not in same]
        cols = [rblock.cols[i] for i in keep] + [gk.pack(rkeys, plan)]
        columns = rblock.columns[keep].append(pandas.Index([gk.PACKED_KEY]))
        block = DeviceBlock(cols, columns, nrows=rblock.nrows, range_start=0, replicated=rblock.replicated)
        pc = comb._partition_mgr_cls._partition_class
        hit = type(comb)(np.array([[pc(block)]], dtype=object), None, columns, [block.nrows], [len(cols)])
        hit._b200_join_tables = {}
        if len(cache) >= 8:
            cache.clear()
        cache[ident] = hit
    return hit, plan


_RESULT_COLUMNS: dict = {}  # id(labels tuple) -> (labels tuple, result column Index): built once per pair of frames


def row_axis_merge(left_qc, right_qc, reset_row_index, **kwargs):
    """``left_qc.merge(right_qc, **kwargs)`` -> new frame (the caller wraps it in its query compiler)."""
    how = kwargs.get("how", "inner")
    on, left_on, right_on = kwargs.get("on"), kwargs.get("left_on"), kwargs.get("right_on")
    if kwargs.get("left_index") or kwargs.get("right_index"):
        raise NotImplementedError("index joins default to pandas in the reference (merge.py:129-136); not on the B200 path")
    if how not in ("left", "inner"):
        raise NotImplementedError(f"merge(how={how!r}) defaults to pandas in the reference; not on the B200 path")

    def as_list(label):
        return list(label) if isinstance(label, (list, tuple)) else ([] if label is None else [label])

    on, left_on, right_on = as_list(on), as_list(left_on), as_list(right_on)
    if on:
        left_on = right_on = on
    if not left_on or not right_on:
        raise NotImplementedError("device merge needs `on` (or `left_on` and `right_on`)")
    if len(left_on) != len(right_on):
        raise ValueError("len(right_on) must equal len(left_on)")
    for lab in left_on:
        if lab not in left_qc.columns:
            raise KeyError(lab)
    for lab in right_on:
        if lab not in right_qc.columns:
            raise KeyError(lab)
    suffixes = kwargs.get("suffixes", ("_x", "_y"))
    left_frame = left_qc._modin_frame
    right_to_broadcast = _combined(right_qc._modin_frame)  # merge.py:178
    if len(left_on) == 1:
        left_on, right_on = left_on[0], right_on[0]
        right_columns, right_dtypes_all = right_qc.columns, right_qc.dtypes
        func = DevMerge(how=how, suffixes=suffixes, left_on=left_on, right_on=right_on,
                        table_cache=right_to_broadcast._b200_join_tables)  # fmt: skip
    else:
        # several int64 key columns: both sides' key tuples are packed into ONE order-preserving int64 over the ranges
        # of BOTH frames (groupkeys.py), the single-key join runs on the image; the right frame's packed form (payload
        # columns + image, with its join table) is kept on the combined right frame
        right_to_broadcast, plan = _packed_right(left_frame, right_to_broadcast, left_on, right_on)
        rblock = right_to_broadcast._partitions[0, 0].get()
        right_columns, right_dtypes_all = rblock.columns, rblock.dtypes
        func = DevMergePacked(left_on, plan, how=how, suffixes=suffixes, table_cache=right_to_broadcast._b200_join_tables)
    labels = func.result_labels(left_qc.columns, right_columns)  # the same tuple object for the same two Indexes
    pay_pos, ll, rl = labels
    memo = _RESULT_COLUMNS.get(id(labels))
    if memo is None or memo[0] is not labels:
        if len(_RESULT_COLUMNS) >= 64:
            _RESULT_COLUMNS.clear()
        memo = _RESULT_COLUMNS[id(labels)] = (labels, pandas.Index(ll + rl))
    new_columns = memo[1]
    right_dtypes = [np.dtype(right_dtypes_all.iloc[i]) for i in pay_pos]
    if how == "left" and any(dt == np.int64 for dt in right_dtypes):
        # pandas turns int64 payload into float64 when ANY left row misses; every row partition on every GPU has to
        # take the same decision or the result's partitions (and ranks) would disagree about the column dtypes
        rblock = right_to_broadcast._partitions[0, 0].get()
        misses = 0
        for row in left_frame._partitions:
            from .block import concat_cols

            blk = concat_cols([p.get() for p in row]) if len(row) > 1 else row[0].get()
            misses += func.count_misses(blk, rblock)
        if dist.is_distributed():
            import torch

            t = torch.tensor([misses], dtype=torch.int64, device=rblock.cols[0].data.device)
            dist.all_reduce_values([t], ["sum"])
            misses = int(t.item())
        func.promote_ints = misses > 0
    new_frame = left_frame.broadcast_apply_full_axis(
        axis=1, func=func, other=right_to_broadcast, keep_partitioning=True, num_splits=1, new_columns=new_columns,
        sync_labels=False,
    )  # fmt: skip
    # merge.py:236-250: the result index is reset to a fresh RangeIndex (metadata only on range-indexed blocks).  A left
    # join against distinct keys keeps the left rows one to one: when the left shard still carries its original
    # job-wide range labels, their start is this rank's offset and no rank has to ask the others for their row counts
    hint = None
    if how == "left" and dist.is_distributed():
        # `_b200_shard_offset` is set by the ingest paths on EVERY rank alike (from_pandas / from_arrow / from_blocks),
        # so all ranks take the same branch -- a collective that only some ranks reach would hang
        off = getattr(left_frame, "_b200_shard_offset", None)
        rblock = right_to_broadcast._partitions[0, 0].get()
        if off is not None and func._table(rblock)[1]:
            hint = off
    return reset_row_index(new_frame, hint) if hint is not None else reset_row_index(new_frame)
