"""sort_values, the Fold template (cumsum / cummax / cummin / ffill) and the Reduce template (var / std) through both
front doors against golden vectors of the UNMODIFIED reference (tests/golden/ext5_sort_fold.npz, produced by
tests/golden/make_golden.py: real Modin on its PandasOnPython engine, NPartitions = 4).

Bars.  sort_values: rows with distinct keys (and NaN keys, which keep their order) row for row, labels included; the
tie-heavy key as "same key column, same rows per run of equal keys" -- the reference's order inside a run is not
defined (unseeded pivot sampling, unstable per-bin sort), the device sort is stable.  cummax / cummin / ffill and int64
cumsum: bit for bit.  float cumsum: NaN positions identical, ``|got - ref| <= 4 log2(n) 2^-53 * running sum of |x|``
(tile tree vs pandas' sequential loop).  var / std: two correct two-pass evaluations differ by at most the sum of
their own bounds -- each sums n non-negative squared deviations with relative error <= 4 log2(n) 2^-53 (tree / pairwise
summation; the error of the mean enters squared) -- so ``rtol = 16 log2(n) 2^-53`` (sqrt halves it for std).
"""

import math
import os
import sys

import numpy as np
import pytest

from modin_b200 import synth
from tests.test_alignment_merge import REF, _modin, needs_modin
from tests.test_oracle import _same_rows_per_key_run

EPS = 2.0**-53
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fifth_batch_checks(pdm, real_modin):
    sys.path.insert(0, GOLDEN)
    from make_golden import fifth_batch_frame

    z = dict(np.load(os.path.join(GOLDEN, "ext5_sort_fold.npz"), allow_pickle=False))
    F = fifth_batch_frame(synth)
    n = len(F)
    df = pdm.DataFrame(F)
    P = lambda x: x._to_pandas()  # noqa: E731
    for by in ("c0", "c2", "u"):
        for asc in (True, False):
            tag = f"sort_{by}_{'asc' if asc else 'desc'}"
            r = P(df.sort_values(by, ascending=asc))
            assert np.array_equal(r.index.to_numpy(), z[tag + "_index"]), tag
            assert np.array_equal(r.to_numpy(dtype=np.float64), z[tag], equal_nan=True), tag
    for asc in (True, False):
        tag = f"sort_key_{'asc' if asc else 'desc'}"
        r = P(df.sort_values("key", ascending=asc))
        assert _same_rows_per_key_run(r["key"].to_numpy(), r.index.to_numpy(), z[tag + "_keys"], z[tag + "_index"]), tag
    fl = ["c0", "c1", "c2", "c3"]
    if real_modin:
        fold = {name: (lambda d, name=name: P(getattr(d, name)())) for name in ("cumsum", "cummax", "cummin", "ffill")}
    else:  # the mirror's API layer is frozen: the Fold template is reached through its query compiler
        fold = {name: (lambda d, name=name: getattr(d._query_compiler, name)(0).to_pandas()) for name in ("cumsum", "cummax", "cummin")}
        fold["ffill"] = lambda d: d._query_compiler.fillna(method="ffill").to_pandas()
    got = fold["cumsum"](df[fl]).to_numpy()
    assert np.array_equal(np.isnan(got), np.isnan(z["cumsum"]))
    bound = 4.0 * math.log2(n) * EPS * np.cumsum(np.abs(np.nan_to_num(F[fl].to_numpy())), axis=0) + 1e-300
    assert (np.isnan(got) | (np.abs(got - z["cumsum"]) <= bound)).all()
    for name in ("cummax", "cummin", "ffill"):
        assert np.array_equal(fold[name](df[fl]).to_numpy(), z[name], equal_nan=True), name
    assert np.array_equal(fold["cumsum"](df[["key", "u"]]).to_numpy(), z["cumsum_int"])
    assert np.array_equal(fold["cummax"](df[["key", "u"]]).to_numpy(), z["cummax_int"])
    rtol = 16.0 * math.log2(n) * EPS
    for name in ("var", "std"):
        for ddof in (1, 0):
            if real_modin:
                got = P(getattr(df[fl], name)(ddof=ddof)).to_numpy()
            else:
                got = getattr(df[fl]._query_compiler, name)(ddof=ddof).to_numpy()
            assert np.allclose(got, z[f"{name}_ddof{ddof}"], rtol=rtol, atol=0), (name, ddof)


def _sixth_batch_checks(pdm):
    """Merge on several int64 key columns (packed into one order-preserving int64, ``groupkeys`` / ``DevMergePacked``)
    against the unmodified reference's result: bit for bit, column labels and dtypes included."""
    sys.path.insert(0, GOLDEN)
    from make_golden import sixth_batch_frames

    z = dict(np.load(os.path.join(GOLDEN, "ext6_multikey_merge.npz"), allow_pickle=False))
    fact, dim, dim_dups = sixth_batch_frames(synth)
    P = lambda x: x._to_pandas()  # noqa: E731
    mf = pdm.DataFrame(fact)
    for how in ("left", "inner"):
        r = P(mf.merge(pdm.DataFrame(dim), on=["a", "b"], how=how))
        assert list(r.columns) == list(z[f"on_{how}_cols"]) and [str(t) for t in r.dtypes] == list(z[f"on_{how}_dtypes"]), how
        assert np.array_equal(r.to_numpy(dtype=np.float64), z[f"on_{how}"], equal_nan=True), how
        assert isinstance(r.index, type(fact.index)) and len(r.index) == len(r) and r.index[0] == 0  # fresh RangeIndex
        r = P(mf.merge(pdm.DataFrame(dim_dups), on=["a", "b"], how=how))
        assert np.array_equal(r.to_numpy(dtype=np.float64), z[f"m2m_{how}"], equal_nan=True), f"repeated pairs, {how}"
    r = P(mf.merge(pdm.DataFrame(dim.rename(columns={"a": "k"})), left_on=["a", "b"], right_on=["k", "b"], how="left"))
    assert list(r.columns) == list(z["lr_on_cols"])
    assert np.array_equal(r.to_numpy(dtype=np.float64), z["lr_on"], equal_nan=True)
    with pytest.raises(KeyError):
        mf.merge(pdm.DataFrame(dim), on=["a", "nope"])
    with pytest.raises(NotImplementedError):  # float key columns have no order-preserving int64 packing here
        mf.merge(pdm.DataFrame(dim), left_on=["a", "c0"], right_on=["a", "d"])


def _seventh_batch_checks(pdm):
    """groupby on a float64 key through the real Modin front door (the key runs as its order-preserving int64 image,
    ``groupkeys.float_image``) against the unmodified reference: keys bit for bit -- NaN keys dropped or kept as one
    last group --, counts / sizes / min / max exact, sums and means within the per-group bound."""
    sys.path.insert(0, GOLDEN)
    from make_golden import seventh_batch_frame

    z = dict(np.load(os.path.join(GOLDEN, "ext7_float_keys.npz"), allow_pickle=False))
    F = seventh_batch_frame(synth)
    n = len(F)
    P = lambda x: x._to_pandas()  # noqa: E731
    df = pdm.DataFrame(F)
    vcols = ["c0", "c1", "c2"]
    abs_sums = F[vcols].abs().groupby(F["fk"]).sum().to_numpy()
    counts = np.maximum(F[vcols].notna().groupby(F["fk"]).sum().to_numpy(), 1)
    for agg in ("sum", "count", "mean", "min", "max", "size"):
        r = P(getattr(df.groupby("fk"), agg)())
        assert np.array_equal(r.index.to_numpy(), z[agg + "_keys"]) and r.index.name == "fk", agg
        got = np.asarray(r, dtype=np.float64).reshape(len(r), -1)
        if agg in ("sum", "mean"):
            tol = 4.0 * math.log2(n) * EPS * (abs_sums if agg == "sum" else abs_sums / counts + np.abs(np.nan_to_num(z[agg])) / (2.0 * math.log2(n))) + 1e-300
            assert ((np.isnan(got) & np.isnan(z[agg])) | (np.abs(got - z[agg]) <= tol)).all(), agg
        else:
            assert np.array_equal(got, z[agg], equal_nan=True), agg
    r = P(df.groupby("fk", dropna=False).sum())
    assert np.array_equal(r.index.to_numpy(), z["sum_keepna_keys"], equal_nan=True)
    assert np.allclose(r.to_numpy(), z["sum_keepna"], rtol=0, atol=1e-9, equal_nan=True)
    r = P(df.groupby("fk", as_index=False).mean())
    assert list(r.columns) == list(z["mean_flat_cols"]) and np.allclose(r.to_numpy(), z["mean_flat"], rtol=0, atol=1e-9, equal_nan=True)


def test_fifth_batch_through_the_mirror_on_the_double(cpu_device):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked test")
    import modin_b200.pandas as bpd

    _fifth_batch_checks(bpd, False)
    _sixth_batch_checks(bpd)


@needs_modin
def test_fifth_batch_under_real_modin_on_the_double(cpu_device):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked test")
    _fifth_batch_checks(_modin(), True)
    _sixth_batch_checks(_modin())
    _seventh_batch_checks(_modin())


@pytest.mark.gpu
def test_fifth_batch_on_b200():
    import modin_b200.pandas as bpd

    _fifth_batch_checks(bpd, False)
    if os.path.isdir(os.path.join(REF, "modin")):
        _fifth_batch_checks(_modin(), True)
