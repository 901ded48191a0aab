"""GPU parity at the BENCHMARK'S OWN SHAPE (``-m gpu``): ``n = 2^26`` rows, ``V = 8`` float64 values,
``G = 1e6`` int64 keys -- the configuration ``bench.py`` times (BASELINE.json configs[3] / configs[4]), on every
table variant the engine can pick there:

* dense (direct-addressed) table with the 64 MB of sums pinned in the persisting-L2 carve-out (the default),
* the same table with the carve-out disabled (``MB200_GB_PERSIST=0``),
* the open-addressed hash table (96 MB: larger than the carve-out, misses L2),
* skewed (Zipf-like) keys, which switch the accumulate kernel to its per-CTA hot-group cache.

Reference = numpy on the very same rows copied back from the device (the generators are pinned to their numpy twin
by ``test_device_generators_match_numpy_twin``): ``np.bincount`` for sizes / counts (bit-exact), ``np.bincount(weights=)``
for the group sums (``|got - ref| <= 4 log2(n) 2^-53 sum|x|`` per group), plus the size-independent property
``sum of group sums == column sums`` against the TreeReduce kernel.  Style follows the reference's
``eval_sum`` / ``eval_size`` (modin/tests/pandas/test_groupby.py:1499, 1559) and ``test_groupby_with_empty_partition``
(modin/tests/core/storage_formats/pandas/test_internals.py:863).

The merge at its benchmark shape (2^26 fact rows x 1e7 dim rows, dense and hashed dim table) is checked bit for bit
against a numpy gather.
"""

import math
import os

import numpy as np
import pandas
import pytest

from modin_b200 import synth

pytestmark = pytest.mark.gpu

EPS = 2.0**-53
N, V, G = 1 << 26, 8, 1_000_000


def _tol(abs_sum, n):
    return 4.0 * max(1.0, math.log2(n)) * EPS * abs_sum + 1e-300


@pytest.fixture(scope="module", params=["uniform", "skewed"])
def frame_and_reference(request):
    """The device frame (2 row partitions, so the fused table absorbs more than one block) and the numpy reference
    computed from the SAME rows (D2H of the generated columns)."""
    from modin_b200 import config

    old = config.NPartitions.get()
    config.NPartitions.put(2)
    df = synth.device_frame(N, V, seed=42, key_modulus=G, npartitions=2, key_skew=request.param == "skewed")
    host = df._to_pandas()
    keys = host["key"].to_numpy()
    vals = [host[f"c{j}"].to_numpy() for j in range(V)]
    del host
    ref = {
        "kind": request.param,
        "size": np.bincount(keys, minlength=G).astype(np.int64),
        "sum": np.stack([np.bincount(keys, weights=v, minlength=G) for v in vals], axis=1),
        "abs": np.stack([np.bincount(keys, weights=np.abs(v), minlength=G) for v in vals], axis=1),
        "min0": pandas.Series(vals[0]).groupby(keys).min().to_numpy(),
    }
    present = ref["size"] > 0
    ref["keys"] = np.nonzero(present)[0].astype(np.int64)
    for k in ("size", "sum", "abs"):
        ref[k] = ref[k][present]
    yield df, ref
    config.NPartitions.put(old)


@pytest.fixture(params=["dense_persist", "dense_nopersist", "hash"])
def table_variant(request):
    from modin_b200 import config

    old = config.GroupbyDenseKeys.get()
    config.GroupbyDenseKeys.put(request.param != "hash")
    if request.param == "dense_nopersist":
        os.environ["MB200_GB_PERSIST"] = "0"
    yield request.param
    os.environ.pop("MB200_GB_PERSIST", None)
    config.GroupbyDenseKeys.put(old)


def test_groupby_at_benchmark_shape(frame_and_reference, table_variant):
    df, ref = frame_and_reference
    g = df.groupby("key")
    got = g.sum()._to_pandas()
    assert np.array_equal(got.index.to_numpy(), ref["keys"]), "group keys (ascending, complete)"
    err = np.abs(got.to_numpy() - ref["sum"])
    tol = _tol(ref["abs"], N)
    assert (err <= tol).all(), f"group sums: max err {err.max()} / tol {tol[err.argmax() // V].max()}"
    # size-independent property against an independent kernel: sum of group sums == column sums (TreeReduce)
    vals = df[[f"c{j}" for j in range(V)]]
    col_sum, col_abs = vals.sum().to_numpy(), vals.abs().sum().to_numpy()
    assert (np.abs(got.to_numpy().sum(axis=0) - col_sum) <= _tol(col_abs, N)).all(), "sum of group sums == column sums"
    sz = g.size()._to_pandas()
    assert np.array_equal(sz.to_numpy().ravel(), ref["size"]) and int(sz.to_numpy().sum()) == N, "group sizes"
    cnt = g.count()._to_pandas()
    assert np.array_equal(cnt.to_numpy(), np.repeat(ref["size"][:, None], V, axis=1)), "group counts (no NaN: == sizes)"
    mn = df[["key", "c0"]].groupby("key").min()._to_pandas()
    assert np.array_equal(mn.to_numpy().ravel().view(np.uint64), ref["min0"].view(np.uint64)), "group min (bit-exact)"


def test_groupby_key_statistics_are_column_metadata():
    """The groupby takes its key range and skew flag from statistics the producing kernel left on the column: no
    pre-pass over the keys, and the numbers equal what a scan of the column finds."""
    from modin_b200 import ops

    n, G2 = 1 << 22, 50_000
    for skew in (False, True):
        col = ops.gen_i64(n, 43, 0, G2, 0, skew=skew)
        assert col.stats is not None
        lo, hi, sampled, dup = col.stats.host()
        k = col.to_numpy()
        assert (lo, hi) == (int(k.min()), int(k.max())) and sampled > 1024
        assert ops.keys_are_skewed(sampled, dup) == skew
        scanned = [int(v) for v in ops.key_range_device([col]).tolist()]
        assert scanned[:2] == [lo, hi] and ops.keys_are_skewed(*scanned[2:]) == skew
    df = synth.device_frame(n, 2, key_modulus=G2, npartitions=2)
    before = ops.key_stats_passes
    r = df.groupby("key").sum()
    r.execute()
    assert ops.key_stats_passes == before, "generated key column paid a statistics pass"
    # a column of unknown origin pays exactly one pass, the first time
    import modin_b200.pandas as bpd

    host = synth.host_frame(1 << 15, 2, key_modulus=777)  # below the ingest-statistics threshold
    d2 = bpd.DataFrame(host)
    before = ops.key_stats_passes
    a = d2.groupby("key").sum()._to_pandas()
    mid = ops.key_stats_passes
    b = d2.groupby("key").sum()._to_pandas()
    assert mid > before and ops.key_stats_passes == mid
    assert np.array_equal(a.index.to_numpy(), b.index.to_numpy())


@pytest.mark.parametrize("join_table", ["dense", "hash"])
def test_merge_at_benchmark_shape(join_table):
    """fact (2^26 rows, key + 2 f64) LEFT / INNER JOIN dim (1e7 rows, key + f64 + int64 payload), 90 % hit rate:
    bit-exact against a numpy gather over the same rows."""
    import modin_b200.pandas as bpd
    from modin_b200 import config

    old = config.NPartitions.get()
    config.NPartitions.put(2)
    if join_table == "hash":
        os.environ["MB200_JOIN_DENSE"] = "0"
    try:
        ndim, keyspace = 10_000_000, 11_111_111
        fact = synth.device_frame(N, 2, seed=42, key_modulus=keyspace, npartitions=2)
        rng = np.random.RandomState(5)
        dim_keys = rng.permutation(keyspace).astype(np.int64)[:ndim]
        d0 = synth.gen_f64(ndim, 11, 0)
        d1 = np.arange(ndim, dtype=np.int64) * 3 - 7
        dim = bpd.DataFrame(pandas.DataFrame({"key": dim_keys, "d0": d0, "d1": d1}))
        fk = fact[["key"]]._to_pandas()["key"].to_numpy()
        row_of = np.full(keyspace, -1, dtype=np.int64)
        row_of[dim_keys] = np.arange(ndim)
        idx = row_of[fk]
        hit = idx >= 0
        assert 0.85 < hit.mean() < 0.95
        left = fact.merge(dim, on="key", how="left")
        assert list(left.columns) == ["key", "c0", "c1", "d0", "d1"] and len(left) == N
        got = left[["d0", "d1"]]._to_pandas()
        want0 = np.where(hit, d0[np.maximum(idx, 0)], np.nan)
        want1 = np.where(hit, d1[np.maximum(idx, 0)].astype(np.float64), np.nan)  # int payload with misses -> float64
        g0, g1 = got["d0"].to_numpy(), got["d1"].to_numpy()
        assert g1.dtype == np.float64
        for g, w, what in ((g0, want0, "d0"), (g1, want1, "d1")):
            same = (g.view(np.uint64) == w.view(np.uint64)) | (np.isnan(g) & np.isnan(w))
            assert same.all(), f"left merge payload {what}: {np.count_nonzero(~same)} rows differ"
        inner = fact.merge(dim, on="key", how="inner")
        assert len(inner) == int(hit.sum())
        gi = inner[["key", "d1"]]._to_pandas()
        assert np.array_equal(gi["key"].to_numpy(), fk[hit]) and np.array_equal(gi["d1"].to_numpy(), d1[idx[hit]])
    finally:
        os.environ.pop("MB200_JOIN_DENSE", None)
        config.NPartitions.put(old)
