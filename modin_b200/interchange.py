"""The dataframe interchange protocol (``__dataframe__``) over device blocks -- SURVEY 8(f)-4.

Reference: ``PandasDataframe.__dataframe__`` / ``from_interchange_dataframe`` (dataframe.py:4803-4867) hand out a
``PandasProtocolDataframe`` (modin/core/dataframe/pandas/interchange/dataframe_protocol/dataframe.py:43-180, columns
column.py, buffers buffer.py) whose buffers are host numpy memory, and import foreign frames by converting them to
pandas first (from_dataframe.py).  Here the producer side wraps the DEVICE columns as they are: every row partition is
a chunk, every column exposes ONE contiguous data buffer that lives in HBM (``__dlpack_device__`` = (kDLCUDA, ordinal),
``__dlpack__`` hands the buffer over without a copy), float64 nulls are NaN (``USE_NAN``), int64 / bool columns are
non-nullable -- exactly what another GPU dataframe library needs to adopt the columns zero-copy.  The consumer side
(``from_dataframe``) adopts CUDA buffers through DLPack without a copy and copies host buffers H2D (``allow_copy``).

Only the dtypes the path computes in travel: float64, int64, bool (one byte per value).  Row labels ride in
``metadata["modin.index"]`` like the reference's (a host ``pandas.Index``; device label columns are read back for it).
"""

from __future__ import annotations

import ctypes
import enum
from typing import Any, Dict, Iterable, Optional, Sequence, Tuple

import numpy as np
import pandas

from .block import DeviceBlock, DeviceColumn, torch_mod


class DTypeKind(enum.IntEnum):  # dataframe_protocol/utils.py:29-56
    INT = 0
    UINT = 1
    FLOAT = 2
    BOOL = 20


class ColumnNullType(enum.IntEnum):  # utils.py:59-80
    NON_NULLABLE = 0
    USE_NAN = 1


class DlpackDeviceType(enum.IntEnum):  # utils.py:83-93
    CPU = 1
    CUDA = 2


_DTYPES = {
    np.dtype("float64"): (DTypeKind.FLOAT, 64, "g", "="),
    np.dtype("int64"): (DTypeKind.INT, 64, "l", "="),
    np.dtype("bool"): (DTypeKind.BOOL, 8, "b", "="),
}


class B200Buffer:
    """One contiguous device buffer (buffer.py:39-117)."""

    def __init__(self, tensor):
        self._x = tensor

    @property
    def bufsize(self) -> int:
        return int(self._x.numel() * self._x.element_size())

    @property
    def ptr(self) -> int:
        return int(self._x.data_ptr())

    def __dlpack__(self, *args, **kwargs):
        return self._x.__dlpack__(*args, **kwargs)

    def __dlpack_device__(self) -> Tuple[int, Optional[int]]:
        if self._x.is_cuda:
            return (DlpackDeviceType.CUDA, int(self._x.device.index or 0))
        return (DlpackDeviceType.CPU, None)

    def __repr__(self) -> str:
        return f"B200Buffer({{'bufsize': {self.bufsize}, 'ptr': {self.ptr}, 'device': {self.__dlpack_device__()[0].name}}})"


class B200Column:
    """One column of one chunk (column.py:52-400): a single data buffer, no validity / offsets buffers."""

    def __init__(self, col: DeviceColumn, allow_copy: bool = True):
        if np.dtype(col.dtype) not in _DTYPES:
            raise NotImplementedError(f"interchange of {col.dtype} columns is not on the B200 path")
        self._col = col
        self._allow_copy = allow_copy

    def size(self) -> int:
        return len(self._col)

    @property
    def offset(self) -> int:
        return 0

    @property
    def dtype(self):
        return _DTYPES[np.dtype(self._col.dtype)]

    @property
    def describe_categorical(self):
        raise TypeError("describe_categorical only works on a column with categorical dtype!")

    @property
    def describe_null(self):
        if self._col.dtype == np.float64:
            return ColumnNullType.USE_NAN, None
        return ColumnNullType.NON_NULLABLE, None

    @property
    def null_count(self) -> int:
        if self._col.dtype != np.float64 or len(self._col) == 0:
            return 0
        from . import ops

        _vals, cnts = ops.reduce_columns("count", [self._col], skipna=True)
        return len(self._col) - int(cnts[0].item())

    @property
    def metadata(self) -> Dict[str, Any]:
        return {}

    def num_chunks(self) -> int:
        return 1

    def get_chunks(self, n_chunks: Optional[int] = None) -> Iterable["B200Column"]:
        if n_chunks in (None, 1):
            yield self
            return
        n = len(self._col)
        if n_chunks < 1:
            raise RuntimeError("n_chunks must be a positive integer")
        step = -(-n // n_chunks) if n else 0
        for k in range(n_chunks):
            lo, hi = min(k * step, n), min((k + 1) * step, n)
            yield B200Column(DeviceColumn(self._col.data[lo:hi], self._col.dtype), self._allow_copy)

    def get_buffers(self) -> Dict[str, Any]:
        return {"data": (B200Buffer(self._col.data), self.dtype), "validity": None, "offsets": None}


class B200ProtocolDataframe:
    """``__dataframe__`` object over the row blocks of a frame (dataframe.py:43-180): one chunk per row partition."""

    version = 0

    def __init__(self, blocks: Sequence[DeviceBlock], index: Optional[pandas.Index], nan_as_null: bool = False,
                 allow_copy: bool = True):  # fmt: skip
        self._blocks = list(blocks)
        self._index = index
        self._nan_as_null = nan_as_null
        self._allow_copy = allow_copy
        labels = {tuple(b.columns) for b in self._blocks}
        if len(labels) > 1:
            raise ValueError("row blocks of one frame carry the same columns")

    def __dataframe__(self, nan_as_null: bool = False, allow_copy: bool = True):
        return B200ProtocolDataframe(self._blocks, self._index, nan_as_null, allow_copy)

    @property
    def metadata(self) -> Dict[str, Any]:
        return {"modin.index": self._index}

    def num_columns(self) -> int:
        return len(self._blocks[0].cols) if self._blocks else 0

    def num_rows(self) -> int:
        return sum(b.nrows for b in self._blocks)

    def num_chunks(self) -> int:
        return len(self._blocks)

    def column_names(self) -> Iterable[str]:
        return list(self._blocks[0].columns) if self._blocks else []

    def _whole(self) -> DeviceBlock:
        from .block import concat_rows

        return self._blocks[0] if len(self._blocks) == 1 else concat_rows(self._blocks)

    def get_column(self, i: int) -> B200Column:
        if len(self._blocks) > 1 and not self._allow_copy:
            raise RuntimeError("a column spanning several row partitions needs a copy; iterate get_chunks() instead")
        return B200Column(self._whole().cols[i], self._allow_copy)

    def get_column_by_name(self, name: str) -> B200Column:
        return self.get_column(list(self.column_names()).index(name))

    def get_columns(self) -> Iterable[B200Column]:
        for i in range(self.num_columns()):
            yield self.get_column(i)

    def select_columns(self, indices: Sequence[int]) -> "B200ProtocolDataframe":
        idx = list(indices)
        blocks = [DeviceBlock([b.cols[i] for i in idx], b.columns[idx], nrows=b.nrows, range_start=b.range_start)
                  for b in self._blocks]  # fmt: skip
        return B200ProtocolDataframe(blocks, self._index, self._nan_as_null, self._allow_copy)

    def select_columns_by_name(self, names: Sequence[str]) -> "B200ProtocolDataframe":
        cols = list(self.column_names())
        return self.select_columns([cols.index(n) for n in names])

    def get_chunks(self, n_chunks: Optional[int] = None) -> Iterable["B200ProtocolDataframe"]:
        """The row partitions as they are; ``n_chunks`` must be a multiple of their number (each partition is cut
        into equal views -- no data moves)."""
        k = len(self._blocks)
        if n_chunks is None or n_chunks == k:
            pos = 0
            for b in self._blocks:
                idx = self._index[pos : pos + b.nrows] if self._index is not None else None
                pos += b.nrows
                yield B200ProtocolDataframe([b], idx, self._nan_as_null, self._allow_copy)
            return
        if n_chunks < 1 or k == 0 or n_chunks % k:
            raise RuntimeError("n_chunks must be a multiple of the number of chunks of the frame")
        per = n_chunks // k
        pos = 0
        for b in self._blocks:
            step = -(-b.nrows // per) if b.nrows else 0
            for j in range(per):
                lo, hi = min(j * step, b.nrows), min((j + 1) * step, b.nrows)
                idx = self._index[pos + lo : pos + hi] if self._index is not None else None
                yield B200ProtocolDataframe([b.slice_rows(lo, hi)], idx, self._nan_as_null, self._allow_copy)
            pos += b.nrows


# ---------------------------------------------------------------- consumer
def _adopt_buffer(buf, kind_bits, length: int, offset: int, allow_copy: bool):
    """Device tensor over ``length`` values of a protocol buffer: zero-copy for CUDA buffers (DLPack), an H2D copy for
    host buffers."""
    t = torch_mod()
    kind, bits = kind_bits
    np_dtype = {(DTypeKind.FLOAT, 64): np.float64, (DTypeKind.INT, 64): np.int64, (DTypeKind.BOOL, 8): np.bool_}.get((int(kind), bits))
    if np_dtype is None:
        raise NotImplementedError(f"interchange: dtype kind {kind} with {bits} bits is not on the B200 path "
                                  "(float64, int64, byte-per-value bool)")  # fmt: skip
    dev_type = int(buf.__dlpack_device__()[0])
    if dev_type == DlpackDeviceType.CUDA:
        tdt = {np.float64: t.float64, np.int64: t.int64, np.bool_: t.uint8}[np_dtype]
        flat = t.from_dlpack(buf).reshape(-1)  # the producer's memory, adopted as it is
        if flat.dtype != tdt:
            flat = flat.view(tdt)  # raw bytes (uint8) or a same-width type: reinterpret, no copy
        return DeviceColumn(flat[offset : offset + length], np_dtype)
    if dev_type != DlpackDeviceType.CPU:
        raise NotImplementedError(f"interchange: buffers on DLPack device type {dev_type} are not on the B200 path")
    if not allow_copy:
        raise RuntimeError("host buffers have to be copied to the device (allow_copy=False)")
    item = np.dtype(np_dtype).itemsize
    raw = (ctypes.c_uint8 * (length * item)).from_address(buf.ptr + offset * item) if length else b""
    host = np.frombuffer(raw, dtype=np_dtype, count=length) if length else np.zeros(0, dtype=np_dtype)
    return DeviceColumn.from_numpy(np.array(host, copy=True))


def _column_from_protocol(col, allow_copy: bool) -> DeviceColumn:
    kind, bits, _fmt, endian = col.dtype
    if endian not in ("=", "<", "|"):
        raise NotImplementedError("interchange: big-endian buffers")
    null_kind = int(col.describe_null[0])
    if null_kind not in (ColumnNullType.NON_NULLABLE, ColumnNullType.USE_NAN):
        if col.null_count:
            raise NotImplementedError("interchange: masked / sentinel nulls are not on the B200 path (NaN only)")
    bufs = col.get_buffers()
    data, (dkind, dbits, _f, _e) = bufs["data"]
    return _adopt_buffer(data, (dkind, dbits), col.size(), col.offset, allow_copy)


def blocks_from_dataframe(df, allow_copy: bool = True):
    """Row blocks (one per chunk) of any object that implements ``__dataframe__``; row labels come from the
    producer's ``metadata`` ("modin.index" / "pandas.index") when it carries them, else 0 .. n-1."""
    if not hasattr(df, "__dataframe__"):
        raise ValueError("`df` does not support DataFrame exchange protocol, i.e. `__dataframe__` method")
    proto = df.__dataframe__(allow_copy=allow_copy)
    names = list(proto.column_names())
    meta = getattr(proto, "metadata", {}) or {}
    index = meta.get("modin.index", meta.get("pandas.index"))
    total = proto.num_rows()
    if index is not None and (len(index) != total or (isinstance(index, pandas.RangeIndex) and index.step == 1
                                                      and index.name is None)):  # fmt: skip
        start, index = (index.start if len(index) == total else 0), None
    else:
        start = 0
    blocks, pos = [], 0
    for chunk in proto.get_chunks():
        cols = [_column_from_protocol(chunk.get_column_by_name(n), allow_copy) for n in names]
        n = chunk.num_rows()
        if index is None:
            blocks.append(DeviceBlock(cols, pandas.Index(names), nrows=n, range_start=start + pos))
        else:  # foreign row labels: on the device when they are numbers (like DeviceBlock.from_pandas), else host-side
            idx = index[pos : pos + n]
            if not isinstance(idx, pandas.MultiIndex) and idx.dtype.kind in "if" and n > 0:
                arr = idx.to_numpy()
                arr = arr.astype(np.int64) if arr.dtype.kind == "i" else arr.astype(np.float64)
                blocks.append(DeviceBlock(cols, pandas.Index(names), nrows=n, index_cols=[DeviceColumn.from_numpy(arr)],
                                          index_names=[idx.name]))  # fmt: skip
            else:
                blocks.append(DeviceBlock(cols, pandas.Index(names), nrows=n, index_host=idx))
        pos += n
    if not blocks:
        blocks = [DeviceBlock([], pandas.Index(names), nrows=0, range_start=0)]
    return blocks


__all__ = ["B200Buffer", "B200Column", "B200ProtocolDataframe", "blocks_from_dataframe"]
