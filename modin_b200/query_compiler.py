"""B200QueryCompiler: the hot-path subset of PandasQueryCompiler with device functors registered
through the operator templates.

Reference registrations being mirrored (modin/core/storage_formats/pandas/query_compiler.py):
Binary ops ``:535-624``; ``count/sum/max/min/mean`` TreeReduce ``:976-1096``; ``abs`` ``:2036``;
``isna/notna/negative`` ``:2063-2106``; ``fillna`` ``:2710-2813``; ``groupby_*`` ``:3741-3748``
with GroupbyReduceImpl (storage_formats/pandas/groupby.py:26-248); ``merge`` ``:657-667`` with
MergeImpl.row_axis_merge (merge.py:104-252).

Operations outside this list raise ``NotImplementedError`` instead of defaulting to pandas: the
north-star forbids a CPU fallback on this path, and a silent D2H -> pandas -> H2D round trip would
hide exactly the cost this backend exists to remove.
"""

from __future__ import annotations

import numpy as np
import pandas

from .algebra import Binary, Fold, GroupByReduce, Map, Reduce, TreeReduce
from .dataframe import B200Dataframe
from .functors import (
    DevAstype,
    DevBinary,
    DevBoolReduce,
    DevClip,
    DevCumulative,
    DevFillna,
    DevGroupbyMap,
    DevGroupbyReduce,
    DevLogical,
    DevMap,
    DevMeanMap,
    DevMeanReduce,
    DevReduce,
    DevRound,
    DevVar,
)
from .partitioning import Bound


def _dtypes_sum(dtypes: pandas.Series, *func_args, **func_kwargs):
    """qc.py:961-974: result dtype of sum = common type of the columns."""
    return np.result_type(*dtypes.values) if len(dtypes) else np.dtype("float64")


class B200QueryCompiler:
    """Query compiler holding a ``B200Dataframe`` (qc.py:279-302)."""

    _modin_frame: B200Dataframe
    _shape_hint = None

    def __init__(self, modin_frame: B200Dataframe, shape_hint=None):
        self._modin_frame = modin_frame
        self._shape_hint = shape_hint

    @property
    def __constructor__(self):
        return type(self)

    storage_format = property(lambda self: self._modin_frame.storage_format)
    engine = property(lambda self: self._modin_frame.engine)

    # ---- metadata ----------------------------------------------------------------------------------
    index = property(lambda self: self._modin_frame.index)
    columns = property(lambda self: self._modin_frame.columns)
    dtypes = property(lambda self: self._modin_frame.dtypes)
    frame_has_materialized_dtypes = property(lambda self: self._modin_frame.has_materialized_dtypes)
    frame_has_materialized_columns = property(lambda self: self._modin_frame.has_materialized_columns)

    def get_axis_len(self, axis):
        return len(self._modin_frame) if axis == 0 else len(self.columns)

    # ---- lifecycle (qc.py:366-387) ------------------------------------------------------------------
    def finalize(self):
        self._modin_frame.finalize()

    def execute(self):
        self.finalize()
        self._modin_frame.wait_computations()

    def free(self):
        pass

    @classmethod
    def from_pandas(cls, df, data_cls=B200Dataframe):
        return cls(data_cls.from_pandas(df))

    @classmethod
    def from_arrow(cls, at, data_cls=B200Dataframe):
        return cls(data_cls.from_arrow(at))

    def to_pandas(self):
        return self._modin_frame.to_pandas()

    def to_numpy(self, **kwargs):
        return self._modin_frame.to_numpy(**kwargs)

    def default_to_pandas(self, pandas_op, *args, **kwargs):
        name = getattr(pandas_op, "__name__", str(pandas_op))
        raise NotImplementedError(
            f"`{name}` has no device implementation in modin_b200 and this execution never defaults to pandas"
        )

    def sort_rows_by_column_values(self, columns, ascending=True, **kwargs):
        """qc.py ``sort_rows_by_column_values`` -> ``PandasDataframe.sort_by`` (df.py:2741-2791): one float64 / int64
        key column, NaN last, ties in original row order (pandas ``kind="stable"``; the default quicksort leaves
        tie order unspecified, so a stable result is a valid answer for every ``kind``)."""
        from .functors import DevSortRows

        pos, asc = DevSortRows.resolve(self.columns, columns, ascending, **kwargs)
        return self.__constructor__(self._modin_frame.sort_by(pos, asc, bool(kwargs.get("ignore_index", False))))

    def drop_duplicates(self, subset=None, keep="first", ignore_index=False):
        """What ``BasePandasDataset.drop_duplicates`` (modin/pandas/base.py:1600-1623) asks of the query compiler --
        ``qc.unique(keep, ignore_index, subset)``, qc.py:2231-2270: rows with the first / last occurrence of every
        value of ONE int64 subset column (``subset=None`` means all columns, so it is accepted only for a
        one-column frame)."""
        from .functors import DevDropDuplicates

        pos = DevDropDuplicates.resolve(self.columns, subset, keep)
        return self.__constructor__(self._modin_frame.drop_duplicate_rows(pos, keep, bool(ignore_index)))

    def relabel_columns(self, new_labels):
        """Same buffers, new column labels (metadata only; ``qc.columns = ...`` in the reference)."""
        new_labels = pandas.Index(new_labels)
        if len(new_labels) != len(self.columns):
            raise ValueError("Length mismatch: expected axis has %d elements, new values have %d elements"
                             % (len(self.columns), len(new_labels)))  # fmt: skip
        return self.__constructor__(self._modin_frame.relabel_columns(new_labels))

    def set_column(self, label, value_qc):
        """``df[label] = series``: append (or replace) one column, sharing every buffer (qc.setitem / insert)."""
        value_qc = value_qc.relabel_columns([label])
        labels = list(self.columns)
        if label in labels:
            pos = labels.index(label)
            others = [i for i in range(len(labels)) if i != pos]
            base = self.__constructor__(self._modin_frame.take_2d_labels_or_positional(col_positions=others))
            stacked = self.__constructor__(base._modin_frame.hstack(value_qc._modin_frame))
            order = list(range(pos)) + [len(others)] + list(range(pos, len(others)))
            return self.__constructor__(stacked._modin_frame.take_2d_labels_or_positional(col_positions=order))
        return self.__constructor__(self._modin_frame.hstack(value_qc._modin_frame))

    def head(self, n):
        return self.__constructor__(self._modin_frame.head_rows(n))

    def tail(self, n):
        return self.__constructor__(self._modin_frame.tail_rows(n))

    def isin(self, values, **kwargs):
        """qc.py ``isin = Map.register(pandas.DataFrame.isin, dtypes=np.bool_)``: integer values, int64 columns."""
        from .functors import DevIsin

        return self.__constructor__(self._modin_frame.map(DevIsin(values), dtypes=np.bool_))

    def getitem_row_mask(self, mask_qc):
        """``df[bool_series]`` (qc.getitem_array with a boolean key, qc.py:2907-2960): rows where the mask holds."""
        return self.__constructor__(self._modin_frame.filter_rows(mask_qc._modin_frame))

    def row_all(self):
        """Row-wise AND over bool columns -> one bool column (``(...).all(axis=1)``; used by dropna)."""
        from .functors import DevRowLogical

        return self.__constructor__(self._modin_frame.rowwise_to_column(DevRowLogical("all")))

    def row_any(self):
        from .functors import DevRowLogical

        return self.__constructor__(self._modin_frame.rowwise_to_column(DevRowLogical("any")))

    def getitem_column_array(self, key, numeric=False, ignore_order=False):
        """qc.py:2885-2905."""
        if numeric:
            positions = list(key)
        else:
            positions = [int(p) for p in self.columns.get_indexer_for(list(key))]
            if any(p < 0 for p in positions):
                missing = [k for k, p in zip(key, positions) if p < 0]
                raise KeyError(f"{missing} not in index")
        return self.__constructor__(self._modin_frame.take_2d_labels_or_positional(col_positions=positions))

    # ---- Map (qc.py:2036-2106) ----------------------------------------------------------------------
    abs = Map.register(DevMap("abs"), dtypes="copy")
    negative = Map.register(DevMap("neg"), dtypes="copy")
    isna = Map.register(DevMap("isna"), dtypes=np.bool_)
    notna = Map.register(DevMap("notna"), dtypes=np.bool_)
    round = Map.register(DevRound(), dtypes="copy")  # qc.py:2438
    clip = Map.register(DevClip(), dtypes="copy")  # qc.py `clip = Map.register(pandas.DataFrame.clip)`
    copy_data = Map.register(DevMap("copy"), dtypes="copy")
    _astype_map = Map.register(DevAstype())  # result dtypes are read back from the blocks

    def astype(self, col_dtypes, errors: str = "raise"):
        """qc.py ``astype(col_dtypes, errors)``: ``col_dtypes`` is one dtype for every column or a
        {column label: dtype} mapping.  Only the widening casts of ``DevAstype``; checked here, before any launch,
        so that a refused cast leaves no half-converted frame."""
        if errors != "raise":
            raise NotImplementedError("astype(errors='ignore') is not on the B200 path")
        return self._astype_map(DevAstype.validate(self.dtypes, col_dtypes))

    def fillna(self, **kwargs):
        """qc.py:2710-2813: scalar / dict values are a Map; ``method``/``limit`` would be a Fold."""
        value = kwargs.get("value")
        if kwargs.get("method") is not None:
            if kwargs["method"] not in ("ffill", "pad") or value is not None or kwargs.get("limit") is not None \
                    or kwargs.get("axis") not in (0, "index", None):  # fmt: skip
                raise NotImplementedError("fillna(method=) on the B200 path: forward fill down the rows, no limit=")
            return self._ffill(0)
        if isinstance(value, type(self)):
            return self.__constructor__(
                self._modin_frame.n_ary_op(Bound(DevBinary("fillna")), [value._modin_frame], join_type="left")
            )
        return self.__constructor__(self._modin_frame.map(Bound(DevFillna(), (), kwargs), dtypes=None))

    # ---- Binary (qc.py:535-624) ---------------------------------------------------------------------
    add = Binary.register(DevBinary("add"), infer_dtypes="try_sample")
    radd = Binary.register(DevBinary("radd"), infer_dtypes="try_sample")
    sub = Binary.register(DevBinary("sub"), infer_dtypes="try_sample")
    rsub = Binary.register(DevBinary("rsub"), infer_dtypes="try_sample")
    mul = Binary.register(DevBinary("mul"), infer_dtypes="try_sample")
    rmul = Binary.register(DevBinary("rmul"), infer_dtypes="try_sample")
    truediv = Binary.register(DevBinary("truediv"), infer_dtypes="try_sample")
    rtruediv = Binary.register(DevBinary("rtruediv"), infer_dtypes="try_sample")
    eq = Binary.register(DevBinary("eq"), infer_dtypes="bool")
    ne = Binary.register(DevBinary("ne"), infer_dtypes="bool")
    lt = Binary.register(DevBinary("lt"), infer_dtypes="bool")
    le = Binary.register(DevBinary("le"), infer_dtypes="bool")
    gt = Binary.register(DevBinary("gt"), infer_dtypes="bool")
    ge = Binary.register(DevBinary("ge"), infer_dtypes="bool")
    # logical ops between bool frames (qc.py:541-571) and ~frame (qc.py `invert`)
    __and__ = Binary.register(DevLogical("and"), infer_dtypes="bool")
    __or__ = Binary.register(DevLogical("or"), infer_dtypes="bool")
    __xor__ = Binary.register(DevLogical("xor"), infer_dtypes="bool")
    invert = Map.register(DevMap("not"), dtypes=np.bool_)

    # ---- TreeReduce (qc.py:976-1096) ----------------------------------------------------------------
    count = TreeReduce.register(DevReduce("count"), DevReduce("count", phase="reduce"))
    sum = TreeReduce.register(DevReduce("sum"), DevReduce("sum", phase="reduce"), compute_dtypes=_dtypes_sum)
    prod = TreeReduce.register(DevReduce("prod"), DevReduce("prod", phase="reduce"), compute_dtypes=_dtypes_sum)
    max = TreeReduce.register(DevReduce("max"), DevReduce("max", phase="reduce"))
    min = TreeReduce.register(DevReduce("min"), DevReduce("min", phase="reduce"))
    mean = TreeReduce.register(DevMeanMap(), DevMeanReduce(), compute_dtypes=lambda *a, **k: np.dtype("float64"))
    any = TreeReduce.register(DevBoolReduce("any"), DevBoolReduce("any", phase="reduce"),
                              compute_dtypes=lambda *a, **k: np.dtype("bool"))  # qc.py:986
    all = TreeReduce.register(DevBoolReduce("all"), DevBoolReduce("all", phase="reduce"),
                              compute_dtypes=lambda *a, **k: np.dtype("bool"))  # qc.py:987

    # ---- Reduce (qc.py:1155-1156: std / var = Reduce.register(pandas.DataFrame.std / var)) ---------------------------
    _var_frame = Reduce.register(DevVar(sqrt=False))
    _std_frame = Reduce.register(DevVar(sqrt=True))

    def _var(self, axis=0, skipna=True, ddof=1, numeric_only=False, sqrt=False, **kwargs):
        """The W variances (standard deviations when ``sqrt``) as a host Series: the API layer of the mirror takes
        reductions of this kind as W host numbers (identical on every rank)."""
        qc = (self._std_frame if sqrt else self._var_frame)(axis=axis, skipna=skipna, ddof=ddof, numeric_only=numeric_only)
        return pandas.Series(qc.to_pandas().iloc[0].to_numpy(dtype=np.float64), index=self.columns, dtype="float64")

    def var(self, axis=0, skipna=True, ddof=1, numeric_only=False, **kwargs):
        """Reduced result as a host ``pandas.Series`` (W numbers; identical on every rank)."""
        return self._var(axis, skipna, ddof, numeric_only, sqrt=False)

    def std(self, axis=0, skipna=True, ddof=1, numeric_only=False, **kwargs):
        return self._var(axis, skipna, ddof, numeric_only, sqrt=True)

    # ---- Fold (qc.py:2429-2431; forward fill is fillna(method="ffill"), qc.py:2809-2810) ----------------------------
    cumsum = Fold.register(DevCumulative("sum"), shape_preserved=True)
    cummax = Fold.register(DevCumulative("max"), shape_preserved=True)
    cummin = Fold.register(DevCumulative("min"), shape_preserved=True)
    _ffill = Fold.register(DevCumulative("ffill"), shape_preserved=True)

    # ---- GroupByReduce (qc.py:3741-3748; impl table storage_formats/pandas/groupby.py:237-248) ------
    groupby_sum = GroupByReduce.register(DevGroupbyMap("sum"), DevGroupbyReduce("sum"))
    groupby_count = GroupByReduce.register(DevGroupbyMap("count"), DevGroupbyReduce("count"))
    groupby_size = GroupByReduce.register(DevGroupbyMap("size"), DevGroupbyReduce("size"))
    groupby_mean = GroupByReduce.register(DevGroupbyMap("mean"), DevGroupbyReduce("mean"))
    groupby_min = GroupByReduce.register(DevGroupbyMap("min"), DevGroupbyReduce("min"))
    groupby_max = GroupByReduce.register(DevGroupbyMap("max"), DevGroupbyReduce("max"))

    # ---- merge (qc.py:657-667 -> MergeImpl.row_axis_merge merge.py:104-252) --------------------------
    def merge(self, right, **kwargs):
        from .merge import row_axis_merge

        return self.__constructor__(row_axis_merge(self, right, _reset_row_index, **kwargs))


def group_keys_to_columns(frame):
    """Result frame of a device groupby with ``as_index=False``: keys as leading columns, rows renumbered 0..G-1 over
    all row partitions (and ranks, in rank order = key order).  Metadata only."""
    from . import dist
    from .functors import keys_to_columns

    lengths = [row[0].length() for row in frame._partitions]
    offset = dist.exclusive_row_offset(sum(lengths)) if dist.is_distributed() else 0
    pc = frame._partition_mgr_cls._partition_class
    rows = []
    for row, n in zip(frame._partitions, lengths):
        if len(row) != 1:
            raise NotImplementedError("groupby(as_index=False) results wider than one column partition")
        rows.append([pc(keys_to_columns(row[0].get(), offset))])
        offset += n
    first = rows[0][0].get()
    start = offset - sum(lengths)
    return type(frame)(np.array(rows, dtype=object), pandas.RangeIndex(start, offset), first.columns, lengths,
                       [len(first.cols)], None)  # fmt: skip


def _frame_device(frame):
    """Device of the frame's first non-empty block (works for the mirror frame and for Modin's ``PandasDataframe``)."""
    from .block import current_device

    for p in frame._partitions.flatten():
        b = p.get()
        if b.cols:
            return b.cols[0].data.device
    return current_device()


def _reset_row_index(frame: B200Dataframe, offset_hint=None) -> B200Dataframe:
    """``reset_index(drop=True)`` on range-indexed device blocks: renumber ``range_start`` so that
    the row partitions of this rank form one contiguous RangeIndex (metadata only, no kernel).
    ``offset_hint``: this rank's first global row position when the caller already knows it (a left merge keeps
    the left frame's rows, so the left shard's own range start is the answer) -- saves the collective + D2H that
    counting the other ranks' rows costs."""
    from . import dist
    from .block import DeviceBlock

    lengths = frame.row_lengths
    offset = 0
    if offset_hint is not None:
        offset = int(offset_hint)
    elif dist.is_distributed():
        import torch

        dev = _frame_device(frame)
        mine = torch.zeros(dist.world_size(), dtype=torch.int64, device=dev)
        mine[dist.rank()] = sum(lengths)
        dist.all_reduce_values([mine], ["sum"])
        offset = int(mine[: dist.rank()].sum().item())
    pc = frame._partition_mgr_cls._partition_class
    rows = []
    for row, n in zip(frame._partitions, lengths):
        new_row = []
        for p in row:
            b = p.get()
            new_row.append(pc(DeviceBlock(b.cols, b.columns, nrows=b.nrows, range_start=offset)))
        rows.append(new_row)
        offset += n
    start = offset - sum(lengths)
    return type(frame)(np.array(rows, dtype=object), pandas.RangeIndex(start, offset), frame._columns_cache,
                       lengths, frame._column_widths_cache, frame._dtypes)  # fmt: skip
