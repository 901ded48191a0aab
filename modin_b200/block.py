"""DeviceBlock: the payload of one partition -- a device-resident columnar (Arrow-layout) block.

In the reference a block partition wraps a ``pandas.DataFrame``
(modin/core/dataframe/pandas/partitioning/partition.py:33-76; the Python engine keeps it in
``self._data``, pandas_on_python/partitioning/partition.py:49-60).  Here the payload is one
contiguous fixed-width device buffer per column (float64 / int64 / bool-as-uint8; float64
nulls are NaN exactly as in pandas), plus host-side labels:

* ``columns``  pandas.Index of column labels (host, O(W));
* ``index``    either a ``pandas.RangeIndex`` (O(1), the 1e9-row case) or device "index
  columns" (groupby keys) that become the pandas index on ``to_pandas``; never a host
  object array of n labels.

Blocks are immutable values (same contract as the reference: "Objects of this class are
treated as immutable", pandas_on_python/partitioning/partition.py:43-44), so column buffers
are shared by reference between blocks (``mask`` of whole columns, merge pass-through
columns) instead of being deep-copied at every boundary like PandasOnPython does.

torch is used only as the allocator / stream owner (``torch.empty`` on the device,
``torch.cuda.current_stream()``); all arithmetic goes through libmodin_b200.
"""

from __future__ import annotations

import warnings
from typing import List, Optional, Sequence

import numpy as np
import pandas

from . import _lib

_TORCH = None


def torch_mod():
    global _TORCH
    if _TORCH is None:
        import torch

        _TORCH = torch
    return _TORCH


def device_ready() -> bool:
    """True when a CUDA device is visible to torch (the only way device blocks can exist)."""
    t = torch_mod()
    return bool(t.cuda.is_available())


def current_device():
    t = torch_mod()
    if not t.cuda.is_available():
        raise _lib.B200Error(
            "no CUDA device visible: modin_b200 executes partitions on B200 GPUs only (no CPU fallback)"
        )
    return t.device("cuda", t.cuda.current_device())


def current_stream() -> int:
    return torch_mod().cuda.current_stream().cuda_stream


_NP2CODE = {np.dtype("float64"): _lib.F64, np.dtype("int64"): _lib.I64, np.dtype("bool"): _lib.U8,
            np.dtype("uint8"): _lib.U8}  # fmt: skip


def dtype_code(np_dtype) -> int:
    try:
        return _NP2CODE[np.dtype(np_dtype)]
    except KeyError:
        raise TypeError(
            f"dtype {np_dtype} is not supported on the B200 partition path (float64 / int64 / bool only)"
        ) from None


def _torch_dtype(np_dtype):
    t = torch_mod()
    np_dtype = np.dtype(np_dtype)
    if np_dtype == np.float64:
        return t.float64
    if np_dtype == np.int64:
        return t.int64
    if np_dtype in (np.dtype("bool"), np.dtype("uint8")):
        return t.uint8
    raise TypeError(f"dtype {np_dtype} is not supported on the B200 partition path")


class KeyStats:
    """Key statistics of an int64 column -- ``{min, max, sampled, duplicated}`` as ``mb200_key_range`` defines them
    -- carried as immutable column METADATA.  The kernel that produces a column leaves them behind (the synthetic
    generators, the ingest pass after an H2D copy); a column of unknown origin pays one 8 B/row pass the first time
    a groupby asks, never again.  The quadruple stays on the device until somebody needs the numbers
    (``host()``: one 32-byte D2H, memoised).  ``exact`` is False for bounds inherited from a parent column (row
    slices): still a valid range for a direct-addressed table, possibly wider than the slice's own."""

    __slots__ = ("_dev", "_host", "exact", "job")

    def __init__(self, dev=None, host=None, exact=True):
        self._dev, self._host, self.exact = dev, (tuple(int(v) for v in host) if host is not None else None), exact
        self.job = None  # job-wide (all ranks) statistics of a group of key columns, cached by the first of them

    def pending(self):
        """Device quadruple not read back yet (None once ``host()`` has run)."""
        return self._dev if self._host is None else None

    def resolve(self, values):
        self._host, self._dev = tuple(int(v) for v in values), None

    def host(self):
        if self._host is None:
            self.resolve(self._dev.tolist())
        return self._host

    def as_bounds(self) -> "KeyStats":
        if self._host is not None:
            return KeyStats(host=self._host, exact=False)
        ks = KeyStats(dev=self._dev, exact=False)
        return ks


class DeviceColumn:
    """One fixed-width column: a 1-D device tensor plus the pandas dtype it stands for (and, for int64 columns
    that have been or may become group keys, their ``KeyStats``)."""

    __slots__ = ("data", "dtype", "stats")

    def __init__(self, data, dtype, stats=None):
        self.data = data  # torch tensor, 1-D, contiguous
        self.dtype = np.dtype(dtype)
        self.stats = stats

    def __len__(self):
        return int(self.data.shape[0])

    @property
    def ptr(self) -> int:
        return self.data.data_ptr()

    @property
    def code(self) -> int:
        return dtype_code(self.dtype)

    @classmethod
    def empty(cls, n: int, dtype) -> "DeviceColumn":
        t = torch_mod()
        return cls(t.empty(int(n), dtype=_torch_dtype(dtype), device=current_device()), dtype)

    @classmethod
    def from_numpy(cls, arr: np.ndarray) -> "DeviceColumn":
        t = torch_mod()
        arr = np.ascontiguousarray(arr)
        dtype = arr.dtype
        dtype_code(dtype)
        target = current_device()  # raises (no CPU fallback) before any host work
        host = arr.view(np.uint8) if dtype == np.bool_ else arr
        with warnings.catch_warnings():
            # pandas copy-on-write hands out read-only arrays; they are only read for the H2D copy
            warnings.simplefilter("ignore", UserWarning)
            dev = t.from_numpy(host).to(target, non_blocking=False)
        col = cls(dev, dtype)
        if dtype == np.int64 and len(col) >= _INGEST_STATS_MIN_ROWS:
            # ingest is the producing step of this column: leave its key statistics behind (one device pass over
            # data the PCIe copy has just delivered) so that a later groupby on it needs no pre-pass
            from . import ops

            col.stats = KeyStats(dev=ops.key_range_device([col]))
        return col

    def to_numpy(self) -> np.ndarray:
        host = self.data.cpu().numpy()
        if self.dtype == np.bool_:
            return host.view(np.bool_)
        return host

    def slice(self, start: int, stop: int) -> "DeviceColumn":
        whole = start <= 0 and stop >= len(self)
        stats = self.stats if whole or self.stats is None else self.stats.as_bounds()
        return DeviceColumn(self.data[start:stop], self.dtype, stats)


_INGEST_STATS_MIN_ROWS = 1 << 16  # below this a groupby's own pass over the keys is noise


class NotOnDevicePath(NotImplementedError, AttributeError):
    """A pandas method was looked up on a device block.  Under real Modin that means a query-compiler method the
    plug-in does not override reached a partition with Modin's pandas lambda.  ``NotImplementedError`` because that
    is this package's contract for anything off the path (no pandas fallback); ``AttributeError`` as well so that
    ``hasattr`` / ``getattr(x, name, default)`` probes (numpy, torch, copy, pickle) keep working."""


class DeviceBlock:
    """Columnar device block = payload of one block partition."""

    __slots__ = ("cols", "columns", "index_cols", "index_names", "range_start", "_nrows", "_pending", "index_host",
                 "replicated", "keys_sorted_unique")

    def __init__(
        self,
        cols: Sequence[DeviceColumn],
        columns,
        nrows: Optional[int] = None,
        range_start: int = 0,
        index_cols: Optional[Sequence[DeviceColumn]] = None,
        index_names: Optional[list] = None,
        index_host: Optional[pandas.Index] = None,
        replicated: bool = False,
    ):
        # replicated: under torch.distributed, True when every rank holds this same block (results of
        # collectives); False when the block is this rank's row shard of a larger frame
        self.replicated = bool(replicated)
        # True for a group table straight out of the hash aggregate: device index keys ascending and distinct
        self.keys_sorted_unique = False
        self.cols: List[DeviceColumn] = list(cols)
        self.columns = columns if isinstance(columns, pandas.Index) else pandas.Index(list(columns))
        if len(self.cols) != len(self.columns):
            raise ValueError(f"{len(self.cols)} device columns for {len(self.columns)} labels")
        if nrows is None:
            if self.cols:
                nrows = len(self.cols[0])
            elif index_cols:
                nrows = len(index_cols[0])
            elif index_host is not None:
                nrows = len(index_host)
            else:
                nrows = 0
        self._nrows = int(nrows)
        self._pending = None
        for c in self.cols:
            if len(c) != self._nrows:
                raise ValueError("ragged device block")
        self.range_start = int(range_start)
        self.index_cols = list(index_cols) if index_cols else None
        self.index_names = list(index_names) if index_names else None
        self.index_host = index_host  # only for small blocks (reduction results etc.)

    # ---- row count: normally a host integer; a block whose size is decided ON THE DEVICE (the emit of a group table)
    # carries buffers of the largest possible size plus the device-side count, and learns its row count -- one 16-byte
    # D2H, the buffers trimmed to views -- only when somebody asks.  Until then nothing waits for the GPU.
    @property
    def nrows(self) -> int:
        if self._pending is not None:
            self._resolve()
        return self._nrows

    @nrows.setter
    def nrows(self, value):
        self._nrows = int(value)

    @classmethod
    def with_device_count(cls, cols, columns, count_dev, index_cols=None, index_names=None, check=None):
        """Block over buffers of CAPACITY rows of which the first ``count_dev[0]`` (a device int64) are valid;
        ``check(values: list)`` is called with ``count_dev.tolist()`` when the count is read back (to raise on flags
        stored behind the count)."""
        cap = len(cols[0]) if cols else (len(index_cols[0]) if index_cols else 0)
        blk = cls(cols, columns, nrows=cap, index_cols=index_cols, index_names=index_names)
        blk._pending = (count_dev, check)
        return blk

    def _resolve(self):
        count_dev, check = self._pending
        self._pending = None
        vals = [int(v) for v in count_dev.tolist()]
        if check is not None:
            check(vals)
        n = vals[0]
        for c in list(self.cols) + list(self.index_cols or []):
            c.data = c.data[:n]
        self._nrows = n

    def __getattr__(self, name):
        # only reached when normal lookup fails (every slot is set in __init__)
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)  # protocol probes: plain AttributeError, no story to tell
        raise NotOnDevicePath(
            f"DeviceBlock has no pandas method {name!r}: the operation that reached this partition has no device "
            "implementation in modin_b200 (unsupported operations raise instead of falling back to pandas)"
        )

    # ---- pandas-like surface used by the partition layer ---------------------------------
    def __len__(self):
        return self.nrows

    @property
    def shape(self):
        return (self.nrows, len(self.cols))

    @property
    def dtypes(self) -> pandas.Series:
        return pandas.Series([c.dtype for c in self.cols], index=self.columns)

    @property
    def index(self) -> pandas.Index:
        """Materialise the row labels on the host (only sensible for small blocks)."""
        if self.index_host is not None:
            return self.index_host
        if self.index_cols:
            arrays = [c.to_numpy() for c in self.index_cols]
            names = self.index_names or [None] * len(arrays)
            if len(arrays) == 1:
                return pandas.Index(arrays[0], name=names[0])
            return pandas.MultiIndex.from_arrays(arrays, names=names)
        return pandas.RangeIndex(self.range_start, self.range_start + self.nrows)

    def has_range_index(self) -> bool:
        return self.index_cols is None and self.index_host is None

    # ---- construction / egress --------------------------------------------------------------
    @classmethod
    def from_pandas(cls, df: pandas.DataFrame) -> "DeviceBlock":
        """H2D ingest of one pandas block (pm.from_pandas -> partition.put, pm.py:1029-1066)."""
        if isinstance(df, pandas.Series):
            df = df.to_frame()
        cols = []
        for i in range(df.shape[1]):
            s = df.iloc[:, i]
            arr = s.to_numpy()
            if arr.dtype == object or arr.dtype.kind not in "fib":
                raise TypeError(
                    f"column {df.columns[i]!r} has dtype {s.dtype}; the B200 partition path carries "
                    "float64 / int64 / bool columns only"
                )
            if arr.dtype.kind == "f" and arr.dtype != np.float64:
                arr = arr.astype(np.float64)
            if arr.dtype.kind == "i" and arr.dtype != np.int64:
                arr = arr.astype(np.int64)
            cols.append(DeviceColumn.from_numpy(arr))
        idx = df.index
        if isinstance(idx, pandas.RangeIndex) and idx.step == 1 and idx.name is None:
            return cls(cols, df.columns, nrows=len(df), range_start=idx.start)
        if not isinstance(idx, pandas.MultiIndex) and idx.dtype.kind in "if" and len(idx) > 0:
            arr = idx.to_numpy()
            arr = arr.astype(np.int64) if arr.dtype.kind == "i" else arr.astype(np.float64)
            return cls(cols, df.columns, nrows=len(df), index_cols=[DeviceColumn.from_numpy(arr)],
                       index_names=[idx.name])  # fmt: skip
        return cls(cols, df.columns, nrows=len(df), index_host=idx)

    def to_pandas(self) -> pandas.DataFrame:
        """D2H egress (partition.to_pandas, part.py:330-350)."""
        data = {i: c.to_numpy() for i, c in enumerate(self.cols)}
        df = pandas.DataFrame(data, index=self.index, copy=False)
        df.columns = self.columns
        if not self.cols:
            df = pandas.DataFrame(index=self.index, columns=self.columns)
        return df

    def to_numpy(self) -> np.ndarray:
        return self.to_pandas().to_numpy()

    # ---- structural ops (no arithmetic) -------------------------------------------------------
    def with_cols(self, cols, columns=None) -> "DeviceBlock":
        return DeviceBlock(
            cols,
            self.columns if columns is None else columns,
            nrows=self.nrows,
            range_start=self.range_start,
            index_cols=self.index_cols,
            index_names=self.index_names,
            index_host=self.index_host,
            replicated=self.replicated,
        )

    def set_axis(self, labels, axis=0, **kwargs) -> "DeviceBlock":
        """New labels over the same column buffers.  Under real Modin this is what ``PandasDataframe``'s deferred label
        synchronisation applies to every partition (``apply_idx_objs``: ``df.set_axis(idx, axis="index")`` /
        ``df.set_axis(cols, axis="columns")``, df.py:940-1030), so the block has to answer it like a pandas frame."""
        if axis in (1, "columns"):
            labels = labels if isinstance(labels, pandas.Index) else pandas.Index(list(labels))
            if len(labels) != len(self.cols):
                raise ValueError(f"Length mismatch: Expected axis has {len(self.cols)} elements, "
                                 f"new values have {len(labels)} elements")  # fmt: skip
            return self.with_cols(self.cols, labels)
        if axis not in (0, "index"):
            raise ValueError(f"No axis named {axis} for object type DeviceBlock")
        labels = labels if isinstance(labels, pandas.Index) else pandas.Index(labels)
        if len(labels) != self.nrows:
            raise ValueError(f"Length mismatch: Expected axis has {self.nrows} elements, "
                             f"new values have {len(labels)} elements")  # fmt: skip
        if isinstance(labels, pandas.RangeIndex) and labels.step == 1 and labels.name is None:
            if self.has_range_index() and labels.start == self.range_start:
                return self
            out = DeviceBlock(self.cols, self.columns, nrows=self.nrows, range_start=labels.start)
        elif len(labels) == 0 and not isinstance(labels, pandas.MultiIndex):
            # no rows, no labels to keep: an empty range, or an empty device label column when the labels carry a name
            # (an empty ``Index([], dtype=object)`` would otherwise make this a host-labelled block, which row-shard
            # gathers refuse -- a rank whose shard of a result is empty)
            if labels.name is None:
                out = DeviceBlock(self.cols, self.columns, nrows=0, range_start=0)
            else:
                kind = np.float64 if labels.dtype.kind == "f" else np.int64
                out = DeviceBlock(self.cols, self.columns, nrows=0, index_cols=[DeviceColumn.empty(0, kind)],
                                  index_names=[labels.name])  # fmt: skip
        elif not isinstance(labels, pandas.MultiIndex) and labels.dtype.kind in "if" and len(labels) > 0:
            arr = labels.to_numpy()
            arr = arr.astype(np.int64) if arr.dtype.kind == "i" else arr.astype(np.float64)
            out = DeviceBlock(self.cols, self.columns, nrows=self.nrows, index_cols=[DeviceColumn.from_numpy(arr)],
                              index_names=[labels.name])  # fmt: skip
        else:
            out = DeviceBlock(self.cols, self.columns, nrows=self.nrows, index_host=labels)
        out.replicated = self.replicated
        return out

    def reindex(self, labels=None, index=None, columns=None, axis=None, fill_value=None, copy=None, **kwargs):
        """``df.reindex(joined_index, axis=axis, fill_value=fill_value)``: what ``PandasDataframe._copartition``
        applies to the gathered blocks of a frame whose labels differ from the joined index (``make_reindexer``,
        df.py:2058-2073).  Answered on the device by ``functors.DevReindex``."""
        from .functors import DevReindex

        if kwargs.get("method") is not None or kwargs.get("level") is not None:
            raise NotImplementedError("reindex(method= / level=) is not on the B200 path")
        out = self
        if labels is not None:
            out = DevReindex()(out, labels, axis=0 if axis is None else axis, fill_value=fill_value)
        if index is not None:
            out = DevReindex()(out, index, axis=0, fill_value=fill_value)
        if columns is not None:
            out = DevReindex()(out, columns, axis=1, fill_value=fill_value)
        return out

    def _reindex_with_indexers(self, reindexers, fill_value=None, copy=None, allow_dups=False, **kwargs):
        """pandas' internal ``NDFrame._reindex_with_indexers``, which ``_copartition`` calls when some frame's labels
        repeat (df.py:2064-2072): ``reindexers = {axis: [new_labels, positional_indexer_or_None]}``."""
        from .functors import DevReindex

        if fill_value is not None and not (isinstance(fill_value, float) and np.isnan(fill_value)):
            raise NotImplementedError("reindex(fill_value=) is not on the B200 path")
        out = self
        for axis, (labels, indexer) in reindexers.items():
            if axis in (0, "index"):
                out = DevReindex.with_indexer(out, labels, indexer)
            else:
                labels = labels if isinstance(labels, pandas.Index) else pandas.Index(labels)
                if indexer is None:
                    out = out.with_cols(out.cols, labels)
                else:
                    from . import ops

                    cols = [out.cols[p] if p >= 0 else ops.full_column(out.nrows, np.float64, float("nan")) for p in indexer]
                    out = out.with_cols(cols, labels)
        return out

    def squeeze(self, axis=None) -> "DeviceBlock":
        """A one-column block IS this package's Series, so squeezing changes nothing.  Modin's Binary template calls
        ``right.squeeze()`` on the broadcast operand (``df.mul(series, axis=0)``, alg/binary.py:396-402) before handing
        it to the block function; ``DevBinary`` pairs a one-column right operand with every left column."""
        return self

    def select_columns(self, positions: Sequence[int]) -> "DeviceBlock":
        """Column subset sharing the buffers (mask along axis 1)."""
        return self.with_cols([self.cols[i] for i in positions], self.columns[list(positions)])

    def slice_rows(self, start: int, stop: int) -> "DeviceBlock":
        """Contiguous row range as views of the same buffers (mask along axis 0)."""
        start = max(0, min(start, self.nrows))
        stop = max(start, min(stop, self.nrows))
        cols = [c.slice(start, stop) for c in self.cols]
        icols = [c.slice(start, stop) for c in self.index_cols] if self.index_cols else None
        ihost = self.index_host[start:stop] if self.index_host is not None else None
        return DeviceBlock(cols, self.columns, nrows=stop - start, range_start=self.range_start + start,
                           index_cols=icols, index_names=self.index_names, index_host=ihost,
                           replicated=self.replicated)  # fmt: skip

    @property
    def T(self) -> "DeviceBlock":
        """Transpose of a VECTOR-shaped block (1 x W or n x 1), on device.  Modin's API layer turns the
        1 x W result frame of a reduction into a Series through ``qc.transpose()`` -> ``lambda df: df.T``
        (modin/pandas/dataframe.py ``_reduce_dimension``; df.py:4745-4775).  General 2-D transposes are not
        on this path."""
        t = torch_mod()
        if self.nrows == 1 and self.cols:
            dts = {c.dtype for c in self.cols}
            dtype = self.cols[0].dtype if len(dts) == 1 else np.result_type(*dts)
            parts = [c.data if c.dtype == dtype else c.data.to(_torch_dtype(dtype)) for c in self.cols]
            col = DeviceColumn(t.cat(parts), dtype)
            blk = DeviceBlock([col], self.index, nrows=len(self.cols), index_host=self.columns)
            blk.replicated = self.replicated
            return blk
        if len(self.cols) == 1 and self.nrows <= 4096:
            c = self.cols[0]
            cols = [DeviceColumn(c.data[i : i + 1], c.dtype) for i in range(self.nrows)]
            blk = DeviceBlock(cols, self.index, nrows=1, index_host=self.columns)
            blk.replicated = self.replicated
            return blk
        if not self.cols or self.nrows == 0:
            return DeviceBlock([], self.index, nrows=len(self.cols), index_host=self.columns)
        raise NotImplementedError("general 2-D transposition is not on the B200 path")

    def column(self, label) -> DeviceColumn:
        loc = self.columns.get_loc(label)
        if not isinstance(loc, (int, np.integer)):
            raise KeyError(f"column label {label!r} is not unique")
        return self.cols[int(loc)]

    def __repr__(self):
        return f"DeviceBlock(nrows={self.nrows}, columns={list(self.columns)!r})"


class HostBlock:
    """A pandas block that was NOT copied to the device at ingest (large float64 / int64 frames,
    ``config.HostStreamMinBytes``).  Two things can happen to it: a fusable elementwise call queue followed by
    ``to_pandas`` streams it through the device chunk by chunk (``mb200_map_host``: H2D / kernel / D2H overlapped on
    three streams, the frame never resident as a whole -- partitioning.stream_to_pandas); anything else calls
    ``materialize()`` first, which is the plain H2D ingest.  The host arrays are only read."""

    __slots__ = ("frame", "nrows", "ncols", "_device")

    def __init__(self, frame: pandas.DataFrame):
        self.frame, self.nrows, self.ncols = frame, len(frame), frame.shape[1]
        self._device = None  # the H2D copy, made once and shared by every partition derived from this block

    @staticmethod
    def eligible(df, min_bytes: int) -> bool:
        if min_bytes <= 0 or not isinstance(df, pandas.DataFrame) or df.shape[1] == 0 or df.shape[1] > 32:
            return False
        if df.shape[0] * df.shape[1] * 8 < min_bytes:
            return False
        return all(dt in (np.dtype("float64"), np.dtype("int64")) for dt in df.dtypes)

    def columns_numpy(self):
        """Contiguous 1-D numpy views of the columns (no copy for frames built column by column)."""
        return [np.ascontiguousarray(self.frame.iloc[:, j].to_numpy()) for j in range(self.ncols)]

    def materialize(self) -> "DeviceBlock":
        if self._device is None:
            self._device = DeviceBlock.from_pandas(self.frame)
        return self._device


def concat_rows(blocks: Sequence[DeviceBlock]) -> DeviceBlock:
    """Row-wise concatenation of blocks with identical columns (pandas.concat in
    deploy_axis_func, axpart.py:445-452) -- a D2D copy into fresh column buffers."""
    blocks = [b for b in blocks]
    if len(blocks) == 1:
        return blocks[0]
    out = _concat_rows(blocks)
    out.replicated = all(b.replicated for b in blocks)  # pieces every rank holds in full stay that
    return out


def _concat_rows(blocks: Sequence[DeviceBlock]) -> DeviceBlock:
    from . import ops

    first = blocks[0]
    ncols = len(first.cols)
    cols = []
    for j in range(ncols):
        dts = {b.cols[j].dtype for b in blocks}
        dtype = first.cols[j].dtype
        if len(dts) > 1:  # int + float partials (count next to sum) promote like pandas.concat
            dtype = np.result_type(*dts)
            if dtype != np.float64:
                raise TypeError(f"row-wise concat of {sorted(str(d) for d in dts)} columns is not on the B200 path")
            pieces = [ops.cast_columns_f64([b.cols[j]])[0] if b.cols[j].dtype != dtype else b.cols[j] for b in blocks]
        else:
            pieces = [b.cols[j] for b in blocks]
        cols.append(ops.concat_columns(pieces))
    nrows = sum(b.nrows for b in blocks)
    if all(b.index_cols for b in blocks):
        k = len(first.index_cols)
        icols = [ops.concat_columns([b.index_cols[i] for b in blocks]) for i in range(k)]
        return DeviceBlock(cols, first.columns, nrows=nrows, index_cols=icols, index_names=first.index_names)
    if all(b.index_host is not None for b in blocks):
        ih = blocks[0].index_host
        for b in blocks[1:]:
            ih = ih.append(b.index_host)
        return DeviceBlock(cols, first.columns, nrows=nrows, index_host=ih)
    contiguous = all(b.has_range_index() for b in blocks)
    if contiguous:
        pos = first.range_start
        for b in blocks:
            if b.range_start != pos:
                contiguous = False
                break
            pos += b.nrows
    if contiguous:
        return DeviceBlock(cols, first.columns, nrows=nrows, range_start=first.range_start)
    if all(b.has_range_index() for b in blocks) and cols:
        # ranges that do not run on from each other (row-wise concat of frames, shard-local slices): still numeric
        # labels, so they stay on the device as an int64 index column instead of becoming a host index
        labels = ops.concat_columns([ops.iota(b.range_start, b.nrows) for b in blocks])
        return DeviceBlock(cols, first.columns, nrows=nrows, index_cols=[labels], index_names=[None])
    ih = blocks[0].index
    for b in blocks[1:]:
        ih = ih.append(b.index)
    return DeviceBlock(cols, first.columns, nrows=nrows, index_host=ih)


def concat_cols(blocks: Sequence[DeviceBlock]) -> DeviceBlock:
    """Column-wise concatenation (shares buffers; no copy)."""
    first = blocks[0]
    if len(blocks) == 1:
        return first
    cols = []
    labels = []
    for b in blocks:
        if b.nrows != first.nrows:
            raise ValueError("column concat of blocks with different row counts")
        cols.extend(b.cols)
        labels.extend(list(b.columns))
    return first.with_cols(cols, pandas.Index(labels))
