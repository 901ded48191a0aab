"""CPU restatement of the reference's partition-execution hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this package; nothing under ``modin_b200/`` does (the product path
has no CPU branch).

What is restated: Modin's *partitioning + operator-template* logic -- the part of the algorithm
that lives in /root/reference -- with the per-block arithmetic delegated to pandas exactly as the
reference does (the arithmetic itself is the third-party dependency ``pandas`` pinned
``>=2.2,<2.4`` in /root/reference/setup.py:49; this image has pandas 3.0.2).  Each function cites
the reference code it follows.  Parity is PINNED: ``tests/golden/*.npz`` were produced by running
the unmodified reference (PandasOnPython engine, NPartitions=4, five pandas-3 import shims, see
``tests/golden/make_golden.py``) in the build container, and ``tests/test_oracle.py`` checks this
restatement against them bit for bit.  ``sort_values`` is pinned since round 2 (tests/golden/ext5_sort_fold.npz):
under pandas 3 the reference's range-partitioning split finds no group (``grp.get_group(scalar)`` after
``groupby([codes])``, dataframe/utils.py:415-420 -- pandas < 2.4 semantics), so the generator restores exactly that
one behaviour (make_golden.py) and the reference sorts again.  Rows with DISTINCT keys (and NaN keys, which keep
their order) are pinned row for row; rows with EQUAL keys have no defined order in the reference (unseeded pivot
sampling + pandas' default unstable kind per bin), so tie-heavy keys are pinned as "same keys, same rows per run of
equal keys" and the restatement / device path fix the order to the stable one.
"""

from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, List, Optional, Sequence

import numpy as np
import pandas

MODIN_UNNAMED_SERIES_LABEL = "__reduced__"  # modin/utils.py:98
MIN_ROW_PARTITION_SIZE = 32  # modin/config/envvars.py:1149-1190
MIN_COLUMN_PARTITION_SIZE = 32


# ------------------------------------------------------------------ partition grid
def compute_chunksize(axis_len: int, num_splits: int, min_block_size: int) -> int:
    """modin/core/storage_formats/pandas/utils.py:28-58."""
    chunksize = axis_len // num_splits
    if axis_len % num_splits:
        chunksize += 1
    return max(chunksize, min_block_size)


def split_into_partitions(df: pandas.DataFrame, npartitions: int) -> List[List[pandas.DataFrame]]:
    """pm.from_pandas / split_pandas_df_into_partitions (partition_manager.py:1029-1149): a 2-D grid
    of ``iloc`` slices, chunk = max(ceil(n / NPartitions), 32) along both axes."""
    row_chunk = compute_chunksize(df.shape[0], npartitions, MIN_ROW_PARTITION_SIZE)
    col_chunk = compute_chunksize(df.shape[1], npartitions, MIN_COLUMN_PARTITION_SIZE)
    grid = []
    for i in range(0, max(len(df), 1), row_chunk):
        grid.append([df.iloc[i : i + row_chunk, j : j + col_chunk].copy()  # python engine copies on put, partition.py:52
                     for j in range(0, max(len(df.columns), 1), col_chunk)])  # fmt: skip
    return grid


def _pmap(fn: Callable, items: Sequence, threads: int):
    if threads <= 1 or len(items) <= 1:
        return [fn(x) for x in items]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        return list(ex.map(fn, items))


def to_pandas(grid) -> pandas.DataFrame:
    """pm.to_pandas (partition_manager.py:989-1005): concat columns within a row, then rows."""
    rows = [pandas.concat(row, axis=1) if len(row) > 1 else row[0] for row in grid]
    return pandas.concat(rows, axis=0) if len(rows) > 1 else rows[0]


# ------------------------------------------------------------------ Map (alg/map.py:32-70, df.py:2253-2319)
def map_partitions(grid, func: Callable, threads: int = 1):
    """pm.map_partitions (partition_manager.py:708-769): ``func`` on a copy of every block
    (pandas_on_python/partitioning/partition.py:110)."""
    flat = [(i, j, blk) for i, row in enumerate(grid) for j, blk in enumerate(row)]
    outs = _pmap(lambda t: func(t[2].copy()), flat, threads)
    new = [[None] * len(row) for row in grid]
    for (i, j, _), o in zip(flat, outs):
        new[i][j] = o
    return new


def df_abs(df: pandas.DataFrame, npartitions: int, threads: int = 1) -> pandas.DataFrame:
    """qc.abs = Map.register(pandas.DataFrame.abs) (query_compiler.py:2036)."""
    return to_pandas(map_partitions(split_into_partitions(df, npartitions), pandas.DataFrame.abs, threads))


def df_fillna(df, value, npartitions: int, threads: int = 1) -> pandas.DataFrame:
    """qc.fillna scalar/dict branch -> frame.map (query_compiler.py:2710-2813)."""
    return to_pandas(map_partitions(split_into_partitions(df, npartitions), lambda b: b.fillna(value=value), threads))


def df_isna(df, npartitions: int, threads: int = 1) -> pandas.DataFrame:
    return to_pandas(map_partitions(split_into_partitions(df, npartitions), pandas.DataFrame.isna, threads))


# ------------------------------------------------------------------ Binary (alg/binary.py:334-458)
def binary_scalar(df, op: str, other, npartitions: int, threads: int = 1) -> pandas.DataFrame:
    """Scalar / list operand: ``frame.map(func, func_args=(other,), lazy=True)`` (binary.py:449-455)."""
    fn = getattr(pandas.DataFrame, op)
    return to_pandas(map_partitions(split_into_partitions(df, npartitions), lambda b: fn(b, other), threads))


def n_ary_op(frames: Sequence[pandas.DataFrame], func: Callable, npartitions: int, threads: int = 1):
    """PandasDataframe.n_ary_op fast path (dataframe.py:3851-3950 with identical axes, :3750-3758) ->
    pm.n_ary_operation (partition_manager.py:1725-1788): ``out[i,j] = func(left[i,j], *right[i,j])``."""
    grids = [split_into_partitions(f, npartitions) for f in frames]
    left = grids[0]
    flat = [(i, j) for i, row in enumerate(left) for j in range(len(row))]
    outs = _pmap(lambda ij: func(*[g[ij[0]][ij[1]].copy() for g in grids]), flat, threads)
    new = [[None] * len(row) for row in left]
    for (i, j), o in zip(flat, outs):
        new[i][j] = o
    return new


def a_mul_b_add_c(a, b, c, npartitions: int, threads: int = 1) -> pandas.DataFrame:
    """``a * b + c`` the way the reference executes it: two Binary passes with a materialised
    temporary (``mul`` then ``add``, each an n_ary_op or a scalar map)."""

    def one(x, op, y):
        if isinstance(y, pandas.DataFrame):
            return to_pandas(n_ary_op([x, y], lambda l, r: getattr(pandas.DataFrame, op)(l, r), npartitions, threads))
        return binary_scalar(x, op, y, npartitions, threads)

    return one(one(a, "mul", b), "add", c)


# ------------------------------------------------------------------ TreeReduce (alg/tree_reduce.py:33-82)
def _build_treereduce_func(func: Callable) -> Callable:
    """PandasDataframe._build_treereduce_func, axis=0 (dataframe.py:2081-2123): Series -> 1-row frame
    labelled ``__reduced__`` so partials stack along the reduced axis."""

    def _tree_reduce_func(df):
        series_result = func(df)
        if isinstance(series_result, pandas.Series):
            result = pandas.DataFrame(series_result).T
            result.index = [MODIN_UNNAMED_SERIES_LABEL]
        else:
            result = pandas.DataFrame(series_result)
        return result

    return _tree_reduce_func


def tree_reduce(df, map_func: Callable, reduce_func: Optional[Callable], npartitions: int, threads: int = 1):
    """PandasDataframe.tree_reduce (dataframe.py:2208-2250): map every block, then per column partition
    ``pandas.concat`` the partials (deploy_axis_func, axis_partition.py:445-452) and reduce once."""
    grid = split_into_partitions(df, npartitions)
    mfn = _build_treereduce_func(map_func)
    rfn = mfn if reduce_func is None else _build_treereduce_func(reduce_func)
    mapped = map_partitions(grid, mfn, threads)
    ncolparts = len(mapped[0])
    reduced = []
    for j in range(ncolparts):
        stacked = pandas.concat([row[j] for row in mapped], axis=0)
        reduced.append(rfn(stacked))
    out = pandas.concat(reduced, axis=1) if len(reduced) > 1 else reduced[0]
    ser = out.iloc[0]
    ser.name = None
    return ser


def df_sum(df, npartitions, skipna=True, min_count=0, threads: int = 1):
    """qc.sum = TreeReduce.register(pandas.DataFrame.sum) (query_compiler.py:978-984)."""
    f = lambda x: x.sum(axis=0, skipna=skipna, min_count=min_count)  # noqa: E731
    return tree_reduce(df, f, None, npartitions, threads)


def df_prod(df, npartitions, skipna=True, threads: int = 1):
    """qc.prod = TreeReduce.register(pandas.DataFrame.prod) (query_compiler.py:985)."""
    return tree_reduce(df, lambda x: x.prod(axis=0, skipna=skipna), None, npartitions, threads)


def df_count(df, npartitions, threads: int = 1):
    """qc.count = TreeReduce.register(pandas.DataFrame.count, pandas.DataFrame.sum) (query_compiler.py:976)."""
    return tree_reduce(df, lambda x: x.count(axis=0), lambda x: x.sum(axis=0), npartitions, threads)


def df_min(df, npartitions, skipna=True, threads: int = 1):
    return tree_reduce(df, lambda x: x.min(axis=0, skipna=skipna), None, npartitions, threads)  # query_compiler.py:1013-1035


def df_max(df, npartitions, skipna=True, threads: int = 1):
    return tree_reduce(df, lambda x: x.max(axis=0, skipna=skipna), None, npartitions, threads)


def df_mean(df, npartitions, skipna=True, threads: int = 1):
    """qc.mean (query_compiler.py:1037-1096): map = {"sum","count"} rows; reduce = sum both, divide."""

    def map_fn(x):
        return pandas.DataFrame(
            {"sum": x.sum(axis=0, skipna=skipna), "count": x.count(axis=0) if skipna else len(x)}
        ).T

    grid = split_into_partitions(df, npartitions)
    mapped = map_partitions(grid, map_fn, threads)
    outs = []
    for j in range(len(mapped[0])):
        stacked = pandas.concat([row[j] for row in mapped], axis=0)
        sums = stacked.loc["sum"].sum(axis=0, skipna=False) if stacked.loc[["sum"]].shape[0] > 1 else stacked.loc["sum"]
        cnts = stacked.loc["count"].sum(axis=0, skipna=False) if stacked.loc[["count"]].shape[0] > 1 else stacked.loc["count"]
        outs.append(sums / cnts)
    res = pandas.concat(outs) if len(outs) > 1 else outs[0]
    res.name = None
    return res.astype("float64")


# ------------------------------------------------------------------ GroupByReduce (alg/groupby.py)
def groupby_reduce(df, by: str, agg: str, npartitions: int, threads: int = 1, dropna: bool = True) -> pandas.DataFrame:
    """``df.groupby(by).<agg>()`` through GroupByReduce: map = per row block
    ``df.groupby(by, as_index=True, sort=True).<map_agg>()`` (alg/groupby.py:124-208); reduce =
    concat of the partial tables + ``groupby(level=0).<reduce_agg>()`` (alg/groupby.py:211-300).
    Aggregation table: storage_formats/pandas/groupby.py:237-248 (sum->sum, count->sum, size->sum,
    mean -> sum & count then divide :184-234)."""
    grid = split_into_partitions(df, npartitions)
    row_blocks = [pandas.concat(row, axis=1) if len(row) > 1 else row[0] for row in grid]

    def map_fn(block):
        g = block.groupby(by, as_index=True, sort=True, observed=True, dropna=dropna)  # groupby_kwargs reach both phases
        if agg == "sum":
            return g.sum()
        if agg == "count":
            return g.count()
        if agg in ("min", "max"):
            return getattr(g, agg)()
        if agg == "size":
            return g.size().to_frame("size")
        if agg == "mean":
            return pandas.concat({"sum": g.sum(), "count": g.count()}, axis=1)
        raise ValueError(agg)

    partials = _pmap(lambda b: map_fn(b.copy()), row_blocks, threads)
    stacked = pandas.concat(partials, axis=0)
    levels = list(range(len(by))) if isinstance(by, (list, tuple)) else 0  # several key columns -> MultiIndex levels
    if agg in ("min", "max"):  # impl table storage_formats/pandas/groupby.py:237-248: ("min","min"), ("max","max")
        return getattr(stacked.groupby(level=levels, sort=True, dropna=dropna), agg)()
    regrouped = stacked.groupby(level=levels, sort=True, dropna=dropna).sum()
    if agg == "mean":
        return regrouped["sum"] / regrouped["count"]
    if agg == "size":
        return regrouped["size"]
    return regrouped


# ------------------------------------------------------------------ broadcast merge (merge.py:104-252)
def broadcast_merge(left, right, on: str, how: str, npartitions: int, threads: int = 1, suffixes=("_x", "_y")):
    """MergeImpl.row_axis_merge: the right frame collapsed to one partition (merge.py:178) and handed to
    every left ROW partition, ``pandas.merge(left_block, right, how=how, on=on, sort=False)``
    (merge.py:139-168), results concatenated, index reset (merge.py:236-250)."""
    assert how in ("left", "inner")
    grid = split_into_partitions(left, npartitions)
    row_blocks = [pandas.concat(row, axis=1) if len(row) > 1 else row[0] for row in grid]
    right_full = right.copy()
    outs = _pmap(lambda b: pandas.merge(b.copy(), right_full, how=how, on=on, sort=False, suffixes=suffixes),
                 row_blocks, threads)  # fmt: skip
    res = pandas.concat(outs, axis=0)
    return res.reset_index(drop=True)


def cpu_threads() -> int:
    return os.cpu_count() or 1


# ------------------------------------------------------------------ more Map / Reduce registrations
def df_round(df, decimals, npartitions: int, threads: int = 1) -> pandas.DataFrame:
    """qc.round = Map.register(pandas.DataFrame.round) (query_compiler.py:2438)."""
    return to_pandas(map_partitions(split_into_partitions(df, npartitions), lambda b: b.round(decimals), threads))


def df_clip(df, lower, upper, npartitions: int, threads: int = 1) -> pandas.DataFrame:
    """qc.clip = Map.register(pandas.DataFrame.clip) with scalar bounds."""
    return to_pandas(map_partitions(split_into_partitions(df, npartitions), lambda b: b.clip(lower, upper), threads))


def reduce_full_axis(df, func: Callable, npartitions: int):
    """Reduce.register(func) (alg/reduce.py; query_compiler.py:1152-1153 for var / std): the function runs on
    whole COLUMN partitions (every row block of a column partition concatenated), so the result does not depend
    on the row partitioning -- only the column grid is restated here."""
    grid = split_into_partitions(df, npartitions)
    ncol_parts = len(grid[0]) if grid else 0
    outs = [func(pandas.concat([row[j] for row in grid], axis=0)) for j in range(ncol_parts)]
    return pandas.concat(outs) if outs else pandas.Series(dtype="float64")


def df_var(df, npartitions: int, ddof: int = 1, skipna: bool = True):
    return reduce_full_axis(df, lambda x: x.var(axis=0, ddof=ddof, skipna=skipna), npartitions)


def df_std(df, npartitions: int, ddof: int = 1, skipna: bool = True):
    return reduce_full_axis(df, lambda x: x.std(axis=0, ddof=ddof, skipna=skipna), npartitions)


def fold_full_axis(df, func: Callable, npartitions: int) -> pandas.DataFrame:
    """Fold.register(func, shape_preserved=True) (alg/fold.py:32-95 -> PandasDataframe.fold, dataframe.py:2357-2400:
    ``map_axis_partitions(axis, partitions, func, keep_partitioning=True)``): the function runs on whole COLUMN
    partitions (every row block of the column partition concatenated, axis_partition.py:445-452) and the result is
    cut back into the original row lengths -- only the column grid matters for the values."""
    grid = split_into_partitions(df, npartitions)
    ncol_parts = len(grid[0]) if grid else 0
    outs = [func(pandas.concat([row[j] for row in grid], axis=0)) for j in range(ncol_parts)]
    return pandas.concat(outs, axis=1) if outs else df.copy()


def df_cumulative(df, which: str, npartitions: int, skipna: bool = True) -> pandas.DataFrame:
    """qc.cumsum / cummax / cummin = Fold.register(pandas.DataFrame.cumsum / ..., shape_preserved=True)
    (query_compiler.py:2429-2431)."""
    if which not in ("cumsum", "cummax", "cummin"):
        raise ValueError(which)
    return fold_full_axis(df, lambda x: getattr(x, which)(axis=0, skipna=skipna), npartitions)


def df_ffill(df, npartitions: int) -> pandas.DataFrame:
    """qc.fillna(method="ffill") (query_compiler.py:2809-2810: ``self._modin_frame.fold(axis, fillna)``; the pinned
    pandas spells it ``fillna(method="ffill")``, pandas 3 ``ffill()``)."""
    return fold_full_axis(df, lambda x: x.ffill(axis=0), npartitions)


# ------------------------------------------------------------------ sort_values (range-partition shuffle)
def sort_values(df, by: str, ascending: bool, npartitions: int, kind: str = "stable") -> pandas.DataFrame:
    """qc.sort_rows_by_column_values -> PandasDataframe.sort_by (dataframe.py:2741-2791) ->
    _apply_func_to_range_partitioning (dataframe.py:2565-2739): pick npartitions-1 pivots from the key column
    (ShuffleSortFunctions.pick_pivots_from_samples_for_sort, dataframe/utils.py:288-332), route every row of every
    row partition to the range its key falls in (split_partitions_using_pivots_for_sort, utils.py:334-475; NaN keys go
    to the last range), concatenate what arrives in partition order and sort each range with
    ``pandas.DataFrame.sort_values``.  With a stable ``kind`` the concatenation over ranges equals the stable
    sort of the whole frame, whatever the pivots -- which is what makes the result partition-independent."""
    grid = split_into_partitions(df, npartitions)
    row_parts = [pandas.concat(row, axis=1) for row in grid]
    keys = df[by].dropna().to_numpy()
    if len(keys) == 0 or npartitions < 2:
        return df.sort_values(by, ascending=ascending, kind=kind)
    qs = np.linspace(0, 1, npartitions + 1)[1:-1]
    pivots = np.quantile(keys, qs, method="inverted_cdf")
    if not ascending:
        pivots = pivots[::-1]
    ranges = [[] for _ in range(npartitions)]
    for part in row_parts:
        k = part[by].to_numpy()
        if ascending:
            dest = np.searchsorted(pivots, k, side="right")
        else:
            dest = np.searchsorted(-pivots, -k, side="right")
        dest = np.where(np.isnan(k.astype(np.float64)), npartitions - 1, dest)
        for r in range(npartitions):
            sel = part[dest == r]
            if len(sel):
                ranges[r].append(sel)
    out = [pandas.concat(parts).sort_values(by, ascending=ascending, kind=kind) for parts in ranges if parts]
    return pandas.concat(out)


# ------------------------------------------------------------------ more registrations, second batch
_DICT_REDUCE = {"sum": "sum", "count": "sum", "min": "min", "max": "max"}  # storage_formats/pandas/groupby.py:237-248


def groupby_dict_reduce(df, by, spec: dict, npartitions: int) -> pandas.DataFrame:
    """``df.groupby(by).agg({column: function})`` -- qc._groupby_dict_reduce (query_compiler.py:3876-3970): the
    dictionary is split into a map dictionary (per row block) and a reduce dictionary (sum of sums / counts,
    min of mins, max of maxes) applied to the concatenated partial tables."""
    grid = split_into_partitions(df, npartitions)
    row_blocks = [pandas.concat(row, axis=1) if len(row) > 1 else row[0] for row in grid]
    partials = [b.groupby(by, as_index=True, sort=True, observed=True).agg(spec) for b in row_blocks]
    stacked = pandas.concat(partials, axis=0)
    levels = list(range(len(by))) if isinstance(by, (list, tuple)) else 0
    return stacked.groupby(level=levels, sort=True).agg({c: _DICT_REDUCE[f] for c, f in spec.items()})


def df_logical(df_a, df_b, op: str, npartitions: int) -> pandas.DataFrame:
    """Binary.register(pandas.DataFrame.__and__ / __or__ / __xor__) (query_compiler.py:541-571) on co-partitioned frames."""
    return to_pandas(n_ary_op([df_a, df_b], lambda a, b: getattr(a, op)(b), npartitions))


def df_any_all(df, which: str, npartitions: int):
    """qc.any / qc.all = TreeReduce.register(pandas.DataFrame.any / all) (query_compiler.py:986-987)."""
    fn = (lambda x: x.any(axis=0)) if which == "any" else (lambda x: x.all(axis=0))
    return tree_reduce(df, fn, None, npartitions)


def df_isin(df, values, npartitions: int) -> pandas.DataFrame:
    """qc.isin = Map.register(pandas.DataFrame.isin)."""
    return to_pandas(map_partitions(split_into_partitions(df, npartitions), lambda b: b.isin(values)))


def filter_rows(df, mask: pandas.Series, npartitions: int) -> pandas.DataFrame:
    """``df[bool_series]``: every row partition keeps its rows where the co-partitioned mask holds; partition order
    and row labels are preserved (the reference goes through getitem_array -> take_2d_labels_or_positional)."""
    grid = split_into_partitions(df, npartitions)
    out, pos = [], 0
    for row in grid:
        blk = pandas.concat(row, axis=1) if len(row) > 1 else row[0]
        out.append(blk[mask.iloc[pos : pos + len(blk)].to_numpy()])
        pos += len(blk)
    return pandas.concat(out, axis=0)


def drop_duplicates(df, subset, keep: str, ignore_index: bool, npartitions: int) -> pandas.DataFrame:
    """modin/pandas/base.py:1600-1623 -> qc.unique (qc.py:2231-2270; without range partitioning and with a subset it
    defers to BaseQueryCompiler.unique, base/query_compiler.py:2410-2435): ``duplicated(keep)`` over the subset
    columns -- a full-axis ``pandas.DataFrame.duplicated`` (qc.py:3346-3387) -- inverted, boolean row selection with
    that mask, then ``reset_index(drop=True)`` when ``ignore_index``."""
    subset = [subset] if not isinstance(subset, (list, tuple)) else list(subset)
    grid = split_into_partitions(df[subset], npartitions)
    if len(grid[0]) != 1:  # wider subsets are md5-hashed row by row first (qc.py:3369-3376); not restated
        raise NotImplementedError("duplicated() over more than one column partition is not restated")
    # apply_full_axis(axis=0): the row blocks of the column partition are concatenated, pandas sees whole columns
    dup = pandas.concat([row[0] for row in grid], axis=0).duplicated(keep=keep)
    out = filter_rows(df, ~dup, npartitions)
    return out.reset_index(drop=True) if ignore_index else out


def concat_frames(frames: Sequence[pandas.DataFrame], axis: int, ignore_index: bool, npartitions: int) -> pandas.DataFrame:
    """modin/pandas/general.py:441 ``concat`` -> qc.concat (qc.py:482-503) -> PandasDataframe.concat (df.py:3953-4096).
    For frames whose labels along the other axis are equal no reindexing is needed: the partition grids are stacked
    (``pm.concat``, partition_manager.py:943-986) and the labels along ``axis`` are appended; ``ignore_index`` is a
    ``reset_index(drop=True)`` (axis 0) afterwards."""
    grids = [split_into_partitions(f, npartitions) for f in frames]
    if axis == 0:
        stacked = [row for g in grids for row in g]
        out = pandas.concat([pandas.concat(row, axis=1) if len(row) > 1 else row[0] for row in stacked], axis=0)
        return out.reset_index(drop=True) if ignore_index else out
    return pandas.concat([to_pandas(g) for g in grids], axis=1)


def df_astype(df, col_dtypes, npartitions: int) -> pandas.DataFrame:
    """qc.astype (qc.py:2335-2343) -> PandasDataframe.astype (df.py:1707-1810): ``df.astype(col_dtypes)`` mapped over
    the block partitions (``lazy_map_partitions``); a mapping is cut down to the labels of each block."""
    def cast(block):
        if isinstance(col_dtypes, dict):
            return block.astype({c: t for c, t in col_dtypes.items() if c in block.columns})
        return block.astype(col_dtypes)

    return to_pandas(map_partitions(split_into_partitions(df, npartitions), cast))


def df_nunique(df, npartitions: int) -> pandas.Series:
    """qc.nunique (qc.py:1109-1113, range partitioning off): ``Reduce.register(pandas.DataFrame.nunique)`` -- a
    full-axis reduce, every column partition sees whole columns."""
    return reduce_full_axis(df, lambda full: full.nunique(), npartitions)


# ------------------------------------------------------------------ _copartition with reindex (df.py:3709-3848)
def copartition_rows(frames: Sequence[pandas.DataFrame], how: str, sort: Optional[bool], npartitions: int):
    """``PandasDataframe._copartition(axis=0, ...)``: join the row labels (``_join_index_objects``, df.py:1984-2075:
    ``left.join(right, how=how, sort=sort)`` pairwise; ``sort`` defaults to "the labels differ", df.py:3759-3760), then
    every frame whose labels differ from the joined index is gathered along the rows, re-indexed
    (``df.reindex(joined_index, axis=0)``, df.py:2073) and split at the base frame's row lengths (df.py:3799-3840).
    Returns the aligned frames as partition grids plus the joined index."""
    base = frames[0]
    if all(base.index.equals(f.index) for f in frames[1:]):
        return [split_into_partitions(f, npartitions) for f in frames], base.index
    if sort is None:
        sort = not all(base.index.equals(f.index) for f in frames[1:])
    joined = base.index
    for f in frames[1:]:
        joined = joined.join(f.index, how=how, sort=sort)
    aligned = [f if f.index.equals(joined) else f.reindex(joined, axis=0) for f in frames]
    return [split_into_partitions(f, npartitions) for f in aligned], joined


def n_ary_op_aligned(frames: Sequence[pandas.DataFrame], func: Callable, npartitions: int, join_type: str = "outer"):
    """``PandasDataframe.n_ary_op`` (df.py:3851-3950) with the general ``_copartition``: align the rows, then
    ``out[i, j] = func(left[i, j], *right[i, j])`` block by block."""
    grids, _ = copartition_rows(list(frames), join_type, None, npartitions)
    left = grids[0]
    out = [[func(*[g[i][j].copy() for g in grids]) for j in range(len(row))] for i, row in enumerate(left)]
    return to_pandas(out)


def setitem_aligned(df: pandas.DataFrame, label, value: pandas.Series, npartitions: int) -> pandas.DataFrame:
    """``df[label] = series`` (modin/pandas/dataframe.py ``__setitem__`` -> ``qc.insert`` / ``setitem``,
    qc.py:3161-3247): the value is re-indexed on the frame's row labels (a LEFT join of the labels) and appended
    as a column of every row partition."""
    aligned = value if df.index.equals(value.index) else value.reindex(df.index)
    blocks = split_into_partitions(df, npartitions)
    vparts = split_into_partitions(aligned.to_frame(label), npartitions)
    rows = [pandas.concat([pandas.concat(r, axis=1) if len(r) > 1 else r[0], v[0]], axis=1) for r, v in zip(blocks, vparts)]
    return pandas.concat(rows, axis=0)


def filter_rows_aligned(df: pandas.DataFrame, mask: pandas.Series, npartitions: int) -> pandas.DataFrame:
    """``df[bool_series]`` with a mask on other row labels (qc.getitem_array -> ``__getitem_bool``, qc.py:3021-3103:
    ``broadcast_apply(..., join_type="left")``): the mask is brought to the frame's labels first."""
    return filter_rows(df, mask.reindex(df.index) if not df.index.equals(mask.index) else mask, npartitions)


def concat_columns_aligned(frames: Sequence[pandas.DataFrame], npartitions: int) -> pandas.DataFrame:
    """``concat(axis=1)`` of frames with different row labels (``PandasDataframe.concat``, df.py:3952-4096): rows
    co-partitioned with an OUTER, unsorted join of the labels, then the column partitions are lined up."""
    grids, joined = copartition_rows(list(frames), "outer", False, npartitions)
    return pandas.concat([to_pandas(g) for g in grids], axis=1)


def broadcast_merge_general(left, right, how: str, npartitions: int, on=None, left_on=None, right_on=None,
                            suffixes=("_x", "_y")) -> pandas.DataFrame:  # fmt: skip
    """``MergeImpl.row_axis_merge`` (merge.py:104-252) with everything ``pandas.merge`` accepts for the key: ``on`` or
    ``left_on`` / ``right_on``; repeated right keys give one output row per match (many-to-many)."""
    grid = split_into_partitions(left, npartitions)
    row_blocks = [pandas.concat(row, axis=1) if len(row) > 1 else row[0] for row in grid]
    kw = dict(on=on) if on is not None else dict(left_on=left_on, right_on=right_on)
    outs = [pandas.merge(b.copy(), right.copy(), how=how, sort=False, suffixes=suffixes, **kw) for b in row_blocks]
    return pandas.concat(outs, axis=0).reset_index(drop=True)
