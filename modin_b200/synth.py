"""Synthetic frames: bit-identical numpy twin of csrc/synth.cu + device frame builders.

``value(seed, col, row)`` is a pure counter-based function (splitmix64 finaliser), so a row range
generated on any GPU equals the same range generated here with numpy -- the tests check that, and
the CPU baseline / oracle legs use these numpy generators on the host.

The device builders are the ``from_map``-style ingest of the backend (reference:
BaseIO.from_map, modin/core/io/io.py:184-209; Ray implementation
modin/core/execution/ray/implementations/pandas_on_ray/io/io.py:309-345): each rank generates its
own row shard directly in HBM -- a 64 GB host frame is never built or shipped.
"""

from __future__ import annotations

from typing import List, Optional

import numpy as np
import pandas

K1 = np.uint64(0x9E3779B97F4A7C15)
K2 = np.uint64(0xD1B54A32D192ED03)
K3 = np.uint64(0x8CB92BA72F3D8DD7)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
SQRT3 = 1.7320508075688772
TWO_M32 = 2.3283064365386963e-10


def _mix64(z: np.ndarray) -> np.ndarray:
    z = (z ^ (z >> np.uint64(30))) * _M1
    z = (z ^ (z >> np.uint64(27))) * _M2
    return z ^ (z >> np.uint64(31))


def _z1(seed: int, col: int, rows: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        base = np.uint64(seed) * K1 + np.uint64(col) * K2 + np.uint64(1)
        return _mix64(rows.astype(np.uint64) + base)


def gen_f64(nrows: int, seed: int, col: int, row_offset: int = 0, nan_per_64k: int = 0) -> np.ndarray:
    """numpy twin of mb200_gen_f64 (approx N(0,1): Irwin-Hall of four exact uniforms)."""
    rows = np.arange(row_offset, row_offset + nrows, dtype=np.int64)
    with np.errstate(over="ignore"):
        z1 = _z1(seed, col, rows)
        z2 = _mix64(z1 + K3)
        z3 = _mix64(z2 + K3)
    lo = np.uint64(0xFFFFFFFF)
    u0 = (z1 >> np.uint64(32)).astype(np.float64) * TWO_M32
    u1 = (z1 & lo).astype(np.float64) * TWO_M32
    u2 = (z2 >> np.uint64(32)).astype(np.float64) * TWO_M32
    u3 = (z2 & lo).astype(np.float64) * TWO_M32
    x = (((u0 + u1) + (u2 + u3)) - 2.0) * SQRT3
    if nan_per_64k:
        x[(z3 & np.uint64(0xFFFF)).astype(np.int64) < nan_per_64k] = np.nan
    return x


def gen_i64(nrows: int, seed: int, col: int, modulus: int, row_offset: int = 0) -> np.ndarray:
    """numpy twin of mb200_gen_i64 (uniform integers in [0, modulus), modulus < 2**32)."""
    rows = np.arange(row_offset, row_offset + nrows, dtype=np.int64)
    z1 = _z1(seed, col, rows)
    with np.errstate(over="ignore"):
        return (((z1 >> np.uint64(32)) * np.uint64(modulus)) >> np.uint64(32)).astype(np.int64)


def skew_levels(modulus: int):
    """Cumulative level weights of the skewed key generator (same integer recurrence as csrc/synth.cu)."""
    T = int(modulus).bit_length() - 1
    w, acc, cum = 1 << 20, 0, []
    for _ in range(T + 1):
        cum.append(acc)
        acc += w
        w = w + (w >> 4) + (w >> 7)
    cum.append(acc)
    return np.array(cum, dtype=np.uint64)


def gen_i64_skew(nrows: int, seed: int, col: int, modulus: int, row_offset: int = 0) -> np.ndarray:
    """numpy twin of mb200_gen_i64_skew (Zipf-like keys in [0, modulus): key 0 takes ~11 % at modulus 1e6)."""
    rows = np.arange(row_offset, row_offset + nrows, dtype=np.int64)
    cum = skew_levels(modulus)
    with np.errstate(over="ignore"):
        z1 = _z1(seed, col, rows)
        z2 = _mix64(z1 + K3)
        pick = ((z2 >> np.uint64(32)) * cum[-1]) >> np.uint64(32)
        t = np.searchsorted(cum[1:-1], pick, side="right").astype(np.uint64)  # largest t with cum[t] <= pick
        rng = np.maximum(np.uint64(modulus) >> t, np.uint64(1))
        return (((z1 >> np.uint64(32)) * rng) >> np.uint64(32)).astype(np.int64)


def host_frame(nrows: int, ncols: int, seed: int = 42, row_offset: int = 0, nan_per_64k: int = 0,
               key_modulus: Optional[int] = None, key_seed: int = 43, prefix: str = "c",
               key_skew: bool = False) -> pandas.DataFrame:  # fmt: skip
    """Host (pandas) synthetic frame: float64 columns c0..c{W-1} and optionally an int64 ``key``."""
    data = {}
    if key_modulus:
        gen = gen_i64_skew if key_skew else gen_i64
        data["key"] = gen(nrows, key_seed, 0, key_modulus, row_offset)
    for j in range(ncols):
        data[f"{prefix}{j}"] = gen_f64(nrows, seed, j, row_offset, nan_per_64k)
    return pandas.DataFrame(data, index=pandas.RangeIndex(row_offset, row_offset + nrows))


def device_blocks(nrows: int, ncols: int, seed: int = 42, nan_per_64k: int = 0, key_modulus: Optional[int] = None,
                  key_seed: int = 43, npartitions: int = 1, prefix: str = "c", key_skew: bool = False) -> List:  # fmt: skip
    """This rank's shard of the synthetic frame as ``npartitions`` device blocks, generated in HBM."""
    from . import dist, ops
    from .block import DeviceBlock

    lo, hi = dist.shard_bounds(nrows)
    local = hi - lo
    blocks = []
    per = -(-local // npartitions) if local else 0
    pos = lo
    for _ in range(npartitions):
        n = min(per, hi - pos)
        if n <= 0 and blocks:
            break
        cols, labels = [], []
        if key_modulus:
            cols.append(ops.gen_i64(n, key_seed, 0, key_modulus, pos, skew=key_skew))
            labels.append("key")
        for j in range(ncols):
            cols.append(ops.gen_f64(n, seed, j, pos, nan_per_64k))
            labels.append(f"{prefix}{j}")
        blocks.append(DeviceBlock(cols, pandas.Index(labels), nrows=n, range_start=pos))
        pos += n
    return blocks


def device_frame(nrows: int, ncols: int, **kwargs):
    """``modin_b200.pandas.DataFrame`` over this rank's device-generated shard."""
    from .dataframe import B200Dataframe
    from .pandas import DataFrame
    from .query_compiler import B200QueryCompiler

    blocks = device_blocks(nrows, ncols, **kwargs)
    frame = B200Dataframe.from_blocks(blocks)
    frame._b200_shard_offset = blocks[0].range_start  # every rank generates its own contiguous shard of the job's rows
    return DataFrame(query_compiler=B200QueryCompiler(frame))
