"""End-to-end path for HOST-resident frames: pinned host columns -> mb200_map_host -> pinned host
columns.  This is the reference-facing call with host buffers on both sides (a host
``pandas``/Arrow frame goes in, a host frame comes out) and is what ``bench.py`` reports as
``e2e``: H2D copy, kernel and D2H copy are all inside the call, overlapped on three streams by
the native pipeline (csrc/hostpipe.cu).
"""

from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .ops import f64_bits, i64_bits


class PinnedColumn:
    """A float64/int64/bool column in page-locked host memory (cudaMallocHost), exposed as numpy."""

    def __init__(self, nrows: int, dtype=np.float64):
        self.lib = _lib.load()
        self.dtype = np.dtype(dtype)
        self.nbytes = int(nrows) * self.dtype.itemsize
        self.ptr = C.c_void_p()
        _lib.check(self.lib.mb200_alloc_host(C.byref(self.ptr), max(self.nbytes, 1)))
        buf = (C.c_char * max(self.nbytes, 1)).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(nrows))

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.mb200_free_host(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def map_host(op: str, in0: Sequence[PinnedColumn], out: Sequence[PinnedColumn], in1=None, in2=None,
             s0: Optional[List[float]] = None, s1: Optional[List[float]] = None, chunk_rows: int = 1 << 22) -> None:  # fmt: skip
    """``out[c] = op(in0[c], in1[c], in2[c]; s0[c], s1[c])`` for host columns; synchronous."""
    lib = _lib.load()
    n = len(in0[0].array)
    code = _lib.F64 if in0[0].dtype == np.float64 else _lib.I64
    conv = f64_bits if code == _lib.F64 else i64_bits

    def ptrs(cols):
        return _lib.ptr_array([c.ptr.value for c in cols]) if cols is not None else None

    s0a = _lib.u64_array([conv(v) for v in s0]) if s0 is not None else None
    s1a = _lib.u64_array([conv(v) for v in s1]) if s1 is not None else None
    _lib.check(lib.mb200_map_host(_lib.OP[op], code, len(in0), ptrs(in0), ptrs(in1), ptrs(in2), ptrs(out), n, s0a,
                                  s1a, int(chunk_rows)))  # fmt: skip


# ------------------------------------------------------------------ pinned host frames + buffer pool
class _PinnedPool:
    """Page-locked host buffers are expensive to create (cudaHostAlloc maps and pins every page), so freed ones are
    kept and handed out again: the result columns of ``to_pandas`` on the streaming path come from here and return
    when the pandas frame that wraps them is garbage collected."""

    def __init__(self, keep_bytes: int = 32 << 30):
        self.free = {}  # nbytes -> [ptr]
        self.kept = 0
        self.keep_bytes = keep_bytes

    def take(self, nbytes: int) -> int:
        lst = self.free.get(nbytes)
        if lst:
            self.kept -= nbytes
            return lst.pop()
        lib = _lib.load()
        ptr = C.c_void_p()
        _lib.check(lib.mb200_alloc_host(C.byref(ptr), max(nbytes, 1)))
        return ptr.value

    def give(self, ptr: int, nbytes: int) -> None:
        if self.kept + nbytes > self.keep_bytes:
            _lib.load().mb200_free_host(C.c_void_p(ptr))
            return
        self.free.setdefault(nbytes, []).append(ptr)
        self.kept += nbytes


_POOL = _PinnedPool()


class _PinnedOwner:
    """Owns one pooled pinned buffer; numpy arrays made from it keep it alive through ``base``."""

    def __init__(self, nrows: int, dtype):
        self.dtype = np.dtype(dtype)
        self.nrows = int(nrows)
        self.nbytes = max(self.nrows * self.dtype.itemsize, 1)
        self.ptr = _POOL.take(self.nbytes)
        self.__array_interface__ = {"shape": (self.nrows,), "typestr": self.dtype.str, "data": (self.ptr, False),
                                    "version": 3}  # fmt: skip

    def __del__(self):
        try:
            _POOL.give(self.ptr, self.nbytes)
        except Exception:
            pass


def pinned_array(nrows: int, dtype=np.float64) -> np.ndarray:
    """numpy array over page-locked host memory from the pool (uninitialised)."""
    return np.asarray(_PinnedOwner(nrows, dtype))


def pinned_frame(data: dict, index=None):
    """``pandas.DataFrame`` whose columns live in page-locked host memory (each column its own buffer, not
    consolidated): the form of host frame whose H2D / D2H copies run at full PCIe speed.  ``data`` maps labels to
    numpy arrays (copied into pinned buffers) or to ``(nrows, dtype)`` for uninitialised columns."""
    import pandas

    cols = {}
    for label, v in data.items():
        if isinstance(v, tuple):
            cols[label] = pinned_array(*v)
        else:
            a = pinned_array(len(v), v.dtype)
            a[:] = v
            cols[label] = a
    return pandas.DataFrame(cols, index=index, copy=False)


def stream_map(op: str, code: int, in0, out, in1=None, in2=None, s0=None, s1=None, chunk_rows: int = 1 << 22) -> None:
    """``mb200_map_host`` over plain numpy columns (contiguous, 8-byte elements; ``out`` may be bool for predicates)."""
    lib = _lib.load()
    conv = f64_bits if code == _lib.F64 else i64_bits

    def ptrs(cols):
        return _lib.ptr_array([c.ctypes.data for c in cols]) if cols is not None else None

    s0a = _lib.u64_array([conv(v) for v in s0]) if s0 is not None else None
    s1a = _lib.u64_array([conv(v) for v in s1]) if s1 is not None else None
    _lib.check(lib.mb200_map_host(_lib.OP[op], code, len(in0), ptrs(in0), ptrs(in1), ptrs(in2), ptrs(out), len(in0[0]),
                                  s0a, s1a, int(chunk_rows)))  # fmt: skip
