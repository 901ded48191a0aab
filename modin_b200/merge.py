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
from .functors import DevMerge


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


_RESULT_COLUMNS: dict = {}  # id(labels tuple) -> (labels tuple, result column Index): built once per pair of frames


def row_axis_merge(left_qc, right_qc, reset_row_index, **kwargs):
    """``left_qc.merge(right_qc, **kwargs)`` -> new frame (the caller wraps it in its query compiler)."""
    how = kwargs.get("how", "inner")
    on, left_on, right_on = kwargs.get("on"), kwargs.get("left_on"), kwargs.get("right_on")
    if kwargs.get("left_index") or kwargs.get("right_index"):
        raise NotImplementedError("index joins default to pandas in the reference (merge.py:129-136); not on the B200 path")
    if how not in ("left", "inner"):
        raise NotImplementedError(f"merge(how={how!r}) defaults to pandas in the reference; not on the B200 path")

    def one(label, what):
        if isinstance(label, (list, tuple)):
            if len(label) != 1:
                raise NotImplementedError(f"device merge joins on exactly one int64 key column ({what} has {len(label)})")
            label = label[0]
        return label

    on, left_on, right_on = one(on, "on"), one(left_on, "left_on"), one(right_on, "right_on")
    if on is not None:
        left_on = right_on = on
    if left_on is None or right_on is None:
        raise NotImplementedError("device merge needs `on` (or `left_on` and `right_on`)")
    if left_on not in left_qc.columns:
        raise KeyError(left_on)
    if right_on not in right_qc.columns:
        raise KeyError(right_on)
    suffixes = kwargs.get("suffixes", ("_x", "_y"))
    left_frame = left_qc._modin_frame
    right_to_broadcast = _combined(right_qc._modin_frame)  # merge.py:178
    func = DevMerge(how=how, suffixes=suffixes, left_on=left_on, right_on=right_on,
                    table_cache=right_to_broadcast._b200_join_tables)  # fmt: skip
    labels = func.result_labels(left_qc.columns, right_qc.columns)  # the same tuple object for the same two Indexes
    pay_pos, ll, rl = labels
    memo = _RESULT_COLUMNS.get(id(labels))
    if memo is None or memo[0] is not labels:
        if len(_RESULT_COLUMNS) >= 64:
            _RESULT_COLUMNS.clear()
        memo = _RESULT_COLUMNS[id(labels)] = (labels, pandas.Index(ll + rl))
    new_columns = memo[1]
    right_dtypes = [np.dtype(right_qc.dtypes.iloc[i]) for i in pay_pos]
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
