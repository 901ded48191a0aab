"""world_size-2 ``gloo`` tests (CPU tensors) of the multi-GPU plumbing in modin_b200/dist.py:
row sharding, the packed all_reduce of TreeReduce partials, the all_gather of row shards and the
range-partitioned all-to-all of groupby partial tables.  The local compute that the GPU kernels
do is stood in for by the CPU oracle / numpy -- only the exchange logic is under test here.
"""

import os
import socket

import numpy as np
import pandas
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from modin_b200 import dist as bdist
from modin_b200 import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, ws, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        ret[rank] = fn(rank, ws)
    finally:
        dist.destroy_process_group()


def _run(fn, ws=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(ws, _free_port(), fn, ret), nprocs=ws, join=True)
    return [ret[r] for r in range(ws)]


def test_shard_bounds_cover_all_rows():
    for n in (0, 1, 7, 8, 1_000_000_007):
        for ws in (1, 2, 3, 8):
            b = [bdist.shard_bounds(n, r, ws) for r in range(ws)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _tree_reduce_job(rank, ws):
    # every rank reduces its shard (numpy stands in for reduce_columns), then the packed all_reduce
    n, W = 10_001, 5
    lo, hi = bdist.shard_bounds(n, rank, ws)
    cols = [synth.gen_f64(hi - lo, 42, j, lo) for j in range(W)]
    sums = [torch.tensor([c.sum()], dtype=torch.float64) for c in cols]
    mins = [torch.tensor([c.min()], dtype=torch.float64) for c in cols]
    cnts = [torch.tensor([len(c)], dtype=torch.int64) for c in cols]
    bdist.all_reduce_values(sums + mins + cnts, ["sum"] * W + ["min"] * W + ["sum"] * W)
    return [float(x) for x in sums], [float(x) for x in mins], [int(x) for x in cnts]


def test_tree_reduce_combine_matches_single_process():
    out = _run(_tree_reduce_job)
    n, W = 10_001, 5
    full = [synth.gen_f64(n, 42, j) for j in range(W)]
    for sums, mins, cnts in out:  # replicated on every rank
        assert np.allclose(sums, [c.sum() for c in full], rtol=0, atol=1e-10)
        assert mins == [c.min() for c in full]
        assert cnts == [n] * W
    assert out[0] == out[1]


def _gather_job(rank, ws):
    n = 1001
    lo, hi = bdist.shard_bounds(n, rank, ws)
    a = torch.from_numpy(synth.gen_f64(hi - lo, 1, 0, lo))
    k = torch.from_numpy(synth.gen_i64(hi - lo, 2, 0, 50, lo))
    ga, gk = bdist.all_gather_rows([a, k])
    return ga.numpy(), gk.numpy()


def test_all_gather_rows_restores_global_order():
    out = _run(_gather_job)
    for ga, gk in out:
        assert np.array_equal(ga.view(np.uint64), synth.gen_f64(1001, 1, 0).view(np.uint64))
        assert np.array_equal(gk, synth.gen_i64(1001, 2, 0, 50))


def _groupby_job(rank, ws):
    """Local partial table (pandas groupby of the shard = stand-in for the hash-aggregate kernel),
    range exchange, local merge of what arrives."""
    import pandas

    n, G, V = 40_000, 3_000, 3
    lo, hi = bdist.shard_bounds(n, rank, ws)
    pdf = synth.host_frame(hi - lo, V, seed=42, row_offset=lo, key_modulus=G, key_seed=43)
    part = pdf.groupby("key").sum()  # ascending unique keys
    keys = torch.from_numpy(part.index.to_numpy().copy())
    cols = [torch.from_numpy(part.iloc[:, j].to_numpy().copy()) for j in range(V)]
    rk, rc = bdist.exchange_by_key_range(keys, cols)
    got = pandas.DataFrame({f"c{j}": rc[j].numpy() for j in range(V)}, index=rk.numpy()).groupby(level=0).sum()
    return got.index.to_numpy(), got.to_numpy()


def test_groupby_range_exchange_equals_global_groupby():
    out = _run(_groupby_job)
    n, G, V = 40_000, 3_000, 3
    pdf = synth.host_frame(n, V, seed=42, key_modulus=G, key_seed=43)
    want = pdf.groupby("key").sum()
    keys = np.concatenate([o[0] for o in out])
    vals = np.concatenate([o[1] for o in out])
    # rank r owns the r-th key range: concatenation in rank order is globally sorted and complete
    assert np.array_equal(keys, want.index.to_numpy())
    assert np.allclose(vals, want.to_numpy(), rtol=0, atol=1e-9)
    assert len(out[0][0]) > 0 and len(out[1][0]) > 0
    assert out[0][0].max() < out[1][0].min()
    # balanced within a factor of two (sampled pivots)
    assert 0.5 < len(out[0][0]) / len(out[1][0]) < 2.0


def _empty_rank_job(rank, ws):
    keys = torch.arange(10, dtype=torch.int64) if rank == 0 else torch.empty(0, dtype=torch.int64)
    vals = [keys.to(torch.float64) * 2.0]
    rk, rc = bdist.exchange_by_key_range(keys, vals)
    return rk.numpy(), rc[0].numpy()


def test_exchange_with_an_empty_rank():
    out = _run(_empty_rank_job)
    keys = np.concatenate([o[0] for o in out])
    vals = np.concatenate([o[1] for o in out])
    assert np.array_equal(np.sort(keys), np.arange(10))
    assert np.array_equal(vals, keys * 2.0)


def test_dense_split_covers_the_range_in_word_aligned_slices():
    for n in (1, 3, 4, 5, 1000, 1_000_000, (1 << 29)):
        for ws in (1, 2, 3, 8):
            sl = [bdist.dense_split(n, r, ws) for r in range(ws)]
            live = [(lo, hi) for lo, hi in sl if hi > lo]
            assert live[0][0] == 0 and live[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(live, live[1:]))
            assert all(lo % 4 == 0 and (hi % 4 == 0 or hi == n) for lo, hi in live)
            assert all(s == (0, 0) for s in sl[len(live):])


def _full_stack_groupby_job(rank, ws):
    """The whole host stack under gloo with the numpy device double: sharded ingest, GroupByReduce template,
    fused dense table + cross-rank merge by collectives (dense=True) or map -> range exchange -> reduce."""
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_double
    import modin_b200.pandas as bpd
    from modin_b200 import config

    out = {}
    with cpu_double.installed():
        config.NPartitions.put(2)
        pdf = synth.host_frame(20_011, 3, seed=42, nan_per_64k=3000, key_modulus=1237, key_seed=43)
        pdf["key"] -= 600  # negative base
        for dense in (True, False):
            config.GroupbyDenseKeys.put(dense)
            g = bpd.DataFrame(pdf).groupby("key")
            for agg in ("sum", "count", "size", "mean", "min", "max"):
                r = getattr(g, agg)()._query_compiler._modin_frame
                blks = [p.get() for p in r._partitions[:, 0]]  # this rank's row partitions, in order
                out[(dense, agg)] = (
                    np.concatenate([b.index_cols[0].data.numpy() for b in blks]),
                    np.concatenate([np.stack([c.data.numpy().astype(np.float64) for c in b.cols], axis=1) for b in blks]),
                )
    return out


def test_full_stack_groupby_dense_and_exchange_paths_agree_with_the_oracle():
    from oracle import reference_path as orc

    out = _run(_full_stack_groupby_job)
    pdf = synth.host_frame(20_011, 3, seed=42, nan_per_64k=3000, key_modulus=1237, key_seed=43)
    pdf["key"] -= 600
    for dense in (True, False):
        for agg in ("sum", "count", "size", "mean", "min", "max"):
            want = orc.groupby_reduce(pdf, "key", agg, 4)
            keys = np.concatenate([o[(dense, agg)][0] for o in out])
            vals = np.concatenate([o[(dense, agg)][1] for o in out])
            assert np.array_equal(keys, want.index.to_numpy()), (dense, agg)  # rank order == key order
            w = want.to_numpy(dtype=np.float64).reshape(len(want), -1)
            assert np.allclose(vals, w, rtol=0, atol=1e-9, equal_nan=True), (dense, agg)
            assert all(len(o[(dense, agg)][0]) > 0 for o in out)


def _full_stack_var_job(rank, ws):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_double
    import modin_b200.pandas as bpd
    from modin_b200 import config

    with cpu_double.installed():
        config.NPartitions.put(2)
        pdf = synth.host_frame(9_001, 3, seed=4, nan_per_64k=3000, key_modulus=31)
        df = bpd.DataFrame(pdf)  # sharded by rank
        vals = df[["c0", "c1", "c2"]]
        spec = {"c1": "max", "c0": "sum", "c2": "count"}
        agg = df.groupby("key").agg(spec)._query_compiler._modin_frame
        blks = [p.get() for p in agg._partitions[:, 0]]
        pdf2 = pdf.copy()
        pdf2["k2"] = synth.gen_i64(len(pdf2), 77, 1, 4) - 2
        mk = bpd.DataFrame(pdf2).groupby(["key", "k2"]).sum()._to_pandas()  # gathers the per-rank key ranges in order
        want2 = pdf2.groupby(["key", "k2"]).sum()
        assert mk.index.equals(want2.index) and np.allclose(mk.to_numpy(), want2.to_numpy(), rtol=0, atol=1e-9)
        mask = (vals > 0.0) & (vals < 2.5)
        assert list(mask.any().to_numpy()) == [True, True, True] and list((vals > 2.5).all().to_numpy()) == [False] * 3
        assert list(mask.sum().to_numpy()) == list(((pdf[["c0", "c1", "c2"]] > 0.0) & (pdf[["c0", "c1", "c2"]] < 2.5)).sum().to_numpy())
        return (vals.var().to_numpy(), vals.std(ddof=0).to_numpy(), vals.mean().to_numpy(),
                np.concatenate([b.index_cols[0].data.numpy() for b in blks]),
                np.concatenate([np.stack([c.data.numpy().astype(np.float64) for c in b.cols], axis=1) for b in blks]))


def test_full_stack_var_std_and_dict_agg_across_ranks():
    out = _run(_full_stack_var_job)
    pdf = synth.host_frame(9_001, 3, seed=4, nan_per_64k=3000, key_modulus=31)
    vals = pdf[["c0", "c1", "c2"]]
    for var, std0, mean, _, _ in out:  # replicated on every rank, equal to the whole-frame statistics
        assert np.allclose(var, vals.var().to_numpy(), rtol=1e-12, atol=0)
        assert np.allclose(std0, vals.std(ddof=0).to_numpy(), rtol=1e-12, atol=0)
        assert np.allclose(mean, vals.mean().to_numpy(), rtol=1e-12, atol=0)
    want = pdf.groupby("key").agg({"c1": "max", "c0": "sum", "c2": "count"})
    keys = np.concatenate([o[3] for o in out])
    got = np.concatenate([o[4] for o in out])
    assert np.array_equal(keys, want.index.to_numpy())
    assert np.allclose(got, want.to_numpy(dtype=np.float64), rtol=0, atol=1e-9)


def _full_stack_late_ops_job(rank, ws):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_double
    import modin_b200.pandas as bpd
    from modin_b200 import config

    def refused(f):
        try:
            f()
        except NotImplementedError:
            return True
        return False

    out = {}
    with cpu_double.installed():
        config.NPartitions.put(2)
        pdf = synth.host_frame(7_001, 3, seed=8, nan_per_64k=4000, key_modulus=29)
        pdf["k2"] = synth.gen_i64(len(pdf), 55, 1, 6) - 3
        df = bpd.DataFrame(pdf)  # sharded by rank; _to_pandas() gathers the shards in rank order
        out["astype"] = df.astype({"key": "float64"})._to_pandas()
        out["filter"] = df[df["c0"] > 0.25]._to_pandas()
        out["dropna"] = df.dropna()._to_pandas()
        out["isin"] = df[df["key"].isin([1, 5, 28])]._to_pandas()
        out["wide"] = bpd.concat([df, df[["c0"]].rename(columns={"c0": "x"})], axis=1)._to_pandas()
        out["head"], out["tail"] = df.head(4000)._to_pandas(), df.tail(4000)._to_pandas()
        # row labels of gathered range-indexed shards: a RangeIndex that does not start at 0, and shard-local slices
        # whose ranges do not run on from each other (explicit labels after the gather)
        shifted = pdf.copy()
        shifted.index = pandas.RangeIndex(1000, 1000 + len(pdf))
        out["shifted"] = bpd.DataFrame(shifted)._to_pandas()
        frame = df._query_compiler._modin_frame
        out["sliced"] = type(df._query_compiler)(frame._slice_local(10, 100))._modin_frame.to_pandas()
        out["nunique"] = df[["key", "k2"]].nunique()  # group tables are merged across the ranks
        out["series_nunique"] = df["key"].nunique()
        # a shard-local answer would be wrong for these: refused on every rank
        out["concat_rows_refused"] = refused(lambda: bpd.concat([df, df]))
        # equal keys sit on different ranks: every rank keeps its survivors, they are gathered, the same pass picks
        # the job-wide first / last -- the result is the whole-frame answer, identical on every rank
        for keep in ("first", "last"):
            out["dd_" + keep] = df.drop_duplicates(subset=["key"], keep=keep)._to_pandas()
        out["dd_ignore"] = df.drop_duplicates(subset=["k2"], ignore_index=True)._to_pandas()
        out["dd_filtered"] = df[df["c0"] > 0.25].drop_duplicates(subset=["key"], keep="last")._to_pandas()
        out["local_rows"] = len(df._query_compiler._modin_frame)
    return out


def test_full_stack_late_ops_across_ranks():
    """astype / row selection / dropna / isin / column concat / head / tail work shard by shard and gather to the
    whole-frame pandas answer; nunique counts groups job-wide; drop_duplicates exchanges the per-rank survivors; row
    concat refuses under torch.distributed instead of answering per shard."""
    import pandas

    out = _run(_full_stack_late_ops_job)
    pdf = synth.host_frame(7_001, 3, seed=8, nan_per_64k=4000, key_modulus=29)
    pdf["k2"] = synth.gen_i64(len(pdf), 55, 1, 6) - 3
    assert sum(o["local_rows"] for o in out) == len(pdf) and all(o["local_rows"] > 0 for o in out)
    want = {
        "astype": pdf.astype({"key": "float64"}),
        "filter": pdf[pdf["c0"] > 0.25],
        "dropna": pdf.dropna(),
        "isin": pdf[pdf["key"].isin([1, 5, 28])],
        "wide": pandas.concat([pdf, pdf[["c0"]].rename(columns={"c0": "x"})], axis=1),
        "head": pdf.head(4000),
        "tail": pdf.tail(4000),
        "shifted": pdf.set_axis(pandas.RangeIndex(1000, 1000 + len(pdf)), axis=0),
        "sliced": pandas.concat([pdf.iloc[lo + 10 : lo + 100] for lo, _ in (bdist.shard_bounds(len(pdf), r, 2) for r in range(2))]),
        "dd_first": pdf.drop_duplicates(subset=["key"], keep="first"),
        "dd_last": pdf.drop_duplicates(subset=["key"], keep="last"),
        "dd_ignore": pdf.drop_duplicates(subset=["k2"], ignore_index=True),
        "dd_filtered": pdf[pdf["c0"] > 0.25].drop_duplicates(subset=["key"], keep="last"),
    }
    for o in out:  # the gathered frame is the same on every rank
        for name, w in want.items():
            g = o[name]
            assert list(g.columns) == list(w.columns) and np.array_equal(g.index.to_numpy(), w.index.to_numpy()), name
            assert [str(t) for t in g.dtypes] == [str(t) for t in w.dtypes], name
            assert np.array_equal(g.to_numpy(dtype=np.float64), w.to_numpy(dtype=np.float64), equal_nan=True), name
        assert list(o["nunique"].index) == ["key", "k2"] and list(o["nunique"]) == list(pdf[["key", "k2"]].nunique())
        assert o["series_nunique"] == pdf["key"].nunique()
        assert o["concat_rows_refused"]


def _full_stack_api_sweep_job(rank, ws):
    """Every case: (device expression on the sharded frame, pandas expression on the whole frame)."""
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_double
    import modin_b200.pandas as bpd
    from modin_b200 import config

    res = {}
    with cpu_double.installed():
        config.NPartitions.put(2)
        pdf = synth.host_frame(6_001, 3, seed=9, nan_per_64k=4000, key_modulus=17)
        other = synth.host_frame(6_001, 3, seed=10, nan_per_64k=0, key_modulus=17)
        rng = np.random.RandomState(3)
        dim = pandas.DataFrame({"key": rng.permutation(17)[:14].astype(np.int64), "d0": rng.randn(14)})
        df, do, dd = bpd.DataFrame(pdf), bpd.DataFrame(other), bpd.DataFrame(dim)
        f = ["c0", "c1", "c2"]
        v, pv, vo, pvo = df[f], pdf[f], do[f], other[f]
        cases = {
            "round": (lambda: v.round(2), lambda: pv.round(2)),
            "clip": (lambda: v.clip(-0.5, 0.5), lambda: pv.clip(-0.5, 0.5)),
            "fillna": (lambda: v.fillna(1.0), lambda: pv.fillna(1.0)),
            "a*b+c": (lambda: v * vo + vo, lambda: pv * pvo + pvo),
            "mul by a column": (lambda: v.mul(df["c1"], axis=0), lambda: pv.mul(pdf["c1"], axis=0)),
            "comparison": (lambda: v < 0.0, lambda: pv < 0.0),
            "logical": (lambda: (v > 0.0) & (v < 1.0), lambda: (pv > 0.0) & (pv < 1.0)),
            "assign": (lambda: df.assign(d=df["c0"] * 2.0), lambda: pdf.assign(d=pdf["c0"] * 2.0)),
            "drop + rename": (lambda: df.drop(columns=["c1"]).rename(columns={"c0": "x"}),
                              lambda: pdf.drop(columns=["c1"]).rename(columns={"c0": "x"})),
            "merge left": (lambda: df.merge(dd, on="key", how="left"), lambda: pdf.merge(dim, on="key", how="left")),
            "merge inner": (lambda: df.merge(dd, on="key", how="inner"), lambda: pdf.merge(dim, on="key", how="inner")),
            "isin": (lambda: df[["key"]].isin([1, 2, 16]), lambda: pdf[["key"]].isin([1, 2, 16])),
            "groupby size": (lambda: df.groupby("key").size(), lambda: pdf.groupby("key").size()),
            "groupby max": (lambda: df.groupby("key").max(), lambda: pdf.groupby("key").max()),
            "filter -> groupby": (lambda: df[df["c0"] > 0.0].groupby("key").sum(), lambda: pdf[pdf["c0"] > 0.0].groupby("key").sum()),
            "filter -> tail": (lambda: df[df["c0"] > 0.0].tail(50), lambda: pdf[pdf["c0"] > 0.0].tail(50)),
            "filter -> head": (lambda: df[df["c0"] > 0.0].head(50), lambda: pdf[pdf["c0"] > 0.0].head(50)),
            "sort -> head": (lambda: df.sort_values("c0").head(20), lambda: pdf.sort_values("c0", kind="stable").head(20)),
            "astype -> sum": (lambda: df.astype({"key": "float64"}).sum(), lambda: pdf.astype({"key": "float64"}).sum()),
            "prod": (lambda: (v.head(40) * 1.1).prod(), lambda: (pv.head(40) * 1.1).prod()),
        }
        for red in ("sum", "mean", "min", "max", "count"):
            cases[red] = ((lambda r=red: getattr(v, r)()), (lambda r=red: getattr(pv, r)()))
        for name, (dev, host) in cases.items():
            g, want = dev(), host()
            g = g._to_pandas() if hasattr(g, "_to_pandas") else g
            labels_ok = list(g.index) == list(want.index)
            vals_ok = g.shape == want.shape and np.allclose(np.asarray(g, dtype=np.float64), np.asarray(want, dtype=np.float64),
                                                            rtol=1e-12, atol=1e-9, equal_nan=True)  # fmt: skip
            res[name] = "ok" if labels_ok and vals_ok else f"values {vals_ok}, labels {labels_ok}, shape {g.shape} vs {want.shape}"
        vc, wvc = df["key"].value_counts()._to_pandas(), pdf["key"].value_counts()  # tie order is unspecified
        res["value_counts"] = "ok" if sorted(zip(vc.index, np.asarray(vc).ravel())) == sorted(zip(wvc.index, wvc)) else "pairs differ"
    return res


def test_full_stack_api_sweep_across_ranks():
    """A regression net for silently wrong multi-GPU answers (the kind the gather's row labels were): compositions
    of the API on a frame sharded over 2 ranks, gathered, against pandas on the whole frame -- values AND labels."""
    out = _run(_full_stack_api_sweep_job)
    for r, res in enumerate(out):
        bad = {k: s for k, s in res.items() if s != "ok"}
        assert not bad and len(res) >= 26, (r, bad)


def _tiny_frames_job(rank, ws):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_double
    import modin_b200.pandas as bpd
    from modin_b200 import config

    res = {}
    with cpu_double.installed():
        config.NPartitions.put(2)
        base = synth.host_frame(64, 3, seed=9, nan_per_64k=0, key_modulus=3)
        dim = pandas.DataFrame({"key": np.array([2, 0, 1], dtype=np.int64), "d0": np.random.RandomState(3).randn(3)})
        f = ["c0", "c1", "c2"]
        cases = {
            "to_pandas": (lambda df: df, lambda p: p),
            "abs": (lambda df: df[f].abs(), lambda p: p[f].abs()),
            "affine": (lambda df: df[f] * 2.0 + 1.0, lambda p: p[f] * 2.0 + 1.0),
            "sum": (lambda df: df[f].sum(), lambda p: p[f].sum()),
            "mean": (lambda df: df[f].mean(), lambda p: p[f].mean()),
            "min": (lambda df: df[f].min(), lambda p: p[f].min()),
            "count": (lambda df: df[f].count(), lambda p: p[f].count()),
            "var": (lambda df: df[f].var(), lambda p: p[f].var()),
            "any": (lambda df: (df[f] > 0.0).any(), lambda p: (p[f] > 0.0).any()),
            "filter": (lambda df: df[df["c0"] > 0.0], lambda p: p[p["c0"] > 0.0]),
            "filter, nothing passes": (lambda df: df[df["c0"] > 100.0], lambda p: p[p["c0"] > 100.0]),
            "dropna": (lambda df: df.dropna(), lambda p: p.dropna()),
            "isin": (lambda df: df[["key"]].isin([1]), lambda p: p[["key"]].isin([1])),
            "groupby sum": (lambda df: df.groupby("key").sum(), lambda p: p.groupby("key").sum()),
            "groupby size": (lambda df: df.groupby("key").size(), lambda p: p.groupby("key").size()),
            "merge left": (lambda df: df.merge(bpd.DataFrame(dim), on="key", how="left"), lambda p: p.merge(dim, on="key", how="left")),
            "merge inner": (lambda df: df.merge(bpd.DataFrame(dim), on="key", how="inner"), lambda p: p.merge(dim, on="key", how="inner")),
            "sort": (lambda df: df.sort_values("c0"), lambda p: p.sort_values("c0", kind="stable")),
            "head": (lambda df: df.head(1), lambda p: p.head(1)),
            "tail": (lambda df: df.tail(3), lambda p: p.tail(3)),
            "head -> filter": (lambda df: (lambda h: h[h["c0"] > 0.0])(df.head(2)), lambda p: (lambda h: h[h["c0"] > 0.0])(p.head(2))),
            "astype": (lambda df: df.astype({"key": "float64"}), lambda p: p.astype({"key": "float64"})),
            "nunique": (lambda df: df[["key"]].nunique(), lambda p: p[["key"]].nunique()),
        }  # fmt: skip
        for n in (0, 1, 2, 5):  # fewer rows than ranks: some shards are empty from the start
            pdf = base.head(n)
            for name, (dev, host) in cases.items():
                want = host(pdf)
                g = dev(bpd.DataFrame(pdf))
                g = g._to_pandas() if hasattr(g, "_to_pandas") else g
                labels_ok = list(g.index) == list(want.index)
                vals_ok = g.shape == want.shape and np.allclose(np.asarray(g, dtype=np.float64), np.asarray(want, dtype=np.float64),
                                                                rtol=1e-12, atol=1e-9, equal_nan=True)  # fmt: skip
                res[f"n={n} {name}"] = "ok" if labels_ok and vals_ok else f"values {vals_ok}, labels {labels_ok}, shape {g.shape} vs {want.shape}"
    return res


@pytest.mark.timeout(300)
def test_tiny_frames_leave_some_ranks_empty():
    """0 / 1 / 2 / 5 rows over THREE ranks.  A rank whose shard is empty from the start used to keep range labels
    through a row filter while the ranks with rows got a device label column; the gather then issued different
    collectives on different ranks and hung.  The ranks now agree on the label layout first (``gather_block``)."""
    out = _run(_tiny_frames_job, ws=3)
    for r, res in enumerate(out):
        bad = {k: s for k, s in res.items() if s != "ok"}
        assert not bad and len(res) == 4 * 23, (r, bad)


def _wide_and_labelled_job(rank, ws):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_double
    import modin_b200.pandas as bpd
    from modin_b200 import config

    res = {}
    with cpu_double.installed():
        config.NPartitions.put(2)
        rng = np.random.RandomState(7)
        fcols = [f"w{i}" for i in range(40)]
        wide = pandas.DataFrame(rng.randn(301, 40), columns=fcols)  # two column partitions on every rank
        wide.iloc[::17, 3] = np.nan
        wide.iloc[::29, 35] = np.nan
        wide.insert(0, "key", rng.randint(0, 7, 301).astype(np.int64))
        pa = synth.host_frame(1003, 3, seed=1, nan_per_64k=3000, key_modulus=11)
        named = pa.set_axis(pandas.Index(np.arange(len(pa))[::-1] * 2, name="rid"), axis=0)
        fidx = pa.set_axis(pandas.Index(np.linspace(0.0, 1.0, len(pa))), axis=0)
        dim = pandas.DataFrame({"key": rng.permutation(11)[:9].astype(np.int64), "d0": rng.randn(9)})
        f = ["c0", "c1", "c2"]
        mk = bpd.DataFrame
        cases = {
            "wide": (lambda: mk(wide), lambda: wide),
            "wide filter": (lambda: (lambda d: d[d["w0"] > 0.0])(mk(wide)), lambda: wide[wide["w0"] > 0.0]),
            "wide dropna": (lambda: mk(wide).dropna(), lambda: wide.dropna()),
            "wide sum": (lambda: mk(wide)[fcols].sum(), lambda: wide[fcols].sum()),
            "wide var": (lambda: mk(wide)[fcols].var(), lambda: wide[fcols].var()),
            "wide a*b+c": (lambda: (lambda d: d[fcols] * d[fcols] + d[fcols])(mk(wide)), lambda: wide[fcols] * wide[fcols] + wide[fcols]),
            "wide groupby": (lambda: mk(wide).groupby("key").sum(), lambda: wide.groupby("key").sum()),
            "wide merge": (lambda: mk(wide).merge(mk(dim), on="key", how="left"), lambda: wide.merge(dim, on="key", how="left")),
            "wide sort": (lambda: mk(wide).sort_values("w5"), lambda: wide.sort_values("w5", kind="stable")),
            "wide head": (lambda: mk(wide).head(200), lambda: wide.head(200)),
            "wide tail": (lambda: mk(wide).tail(200), lambda: wide.tail(200)),
            "wide astype": (lambda: mk(wide).astype({"key": "float64"}), lambda: wide.astype({"key": "float64"})),
            "wide assign": (lambda: (lambda d: d.assign(z=d["w1"] * 2.0))(mk(wide)), lambda: wide.assign(z=wide["w1"] * 2.0)),
            "wide concat columns": (lambda: (lambda d: bpd.concat([d, d[["w0"]].rename(columns={"w0": "zz"})], axis=1))(mk(wide)),
                                    lambda: pandas.concat([wide, wide[["w0"]].rename(columns={"w0": "zz"})], axis=1)),
            "named index": (lambda: mk(named), lambda: named),
            "named index filter": (lambda: (lambda d: d[d["c0"] > 0.0])(mk(named)), lambda: named[named["c0"] > 0.0]),
            "named index sort": (lambda: mk(named).sort_values("c0"), lambda: named.sort_values("c0", kind="stable")),
            "named index head": (lambda: mk(named).head(600), lambda: named.head(600)),
            "named index tail": (lambda: mk(named).tail(600), lambda: named.tail(600)),
            "named index groupby": (lambda: mk(named).groupby("key").sum(), lambda: named.groupby("key").sum()),
            "named index * 2": (lambda: mk(named)[f] * 2.0, lambda: named[f] * 2.0),
            "named index merge": (lambda: mk(named).merge(mk(dim), on="key", how="left"), lambda: named.merge(dim, on="key", how="left")),
            "float index": (lambda: mk(fidx), lambda: fidx),
            "float index filter": (lambda: (lambda d: d[d["c0"] > 0.0])(mk(fidx)), lambda: fidx[fidx["c0"] > 0.0]),
            "float index tail": (lambda: mk(fidx).tail(600), lambda: fidx.tail(600)),
            "float index sort": (lambda: mk(fidx).sort_values("c1"), lambda: fidx.sort_values("c1", kind="stable")),
            "float index head -> filter": (lambda: (lambda h: h[h["c0"] > 0.0])(mk(fidx).head(100)),
                                           lambda: (lambda h: h[h["c0"] > 0.0])(fidx.head(100))),
        }  # fmt: skip
        for name, (dev, host) in cases.items():
            g, want = dev(), host()
            g = g._to_pandas() if hasattr(g, "_to_pandas") else g
            labels_ok = list(g.index) == list(want.index) and g.index.name == want.index.name
            cols_ok = not hasattr(want, "columns") or list(g.columns) == list(want.columns)
            vals_ok = g.shape == want.shape and np.allclose(np.asarray(g, dtype=np.float64), np.asarray(want, dtype=np.float64),
                                                            rtol=1e-12, atol=1e-9, equal_nan=True)  # fmt: skip
            res[name] = "ok" if labels_ok and cols_ok and vals_ok else f"values {vals_ok}, labels {labels_ok}, columns {cols_ok}"
    return res


@pytest.mark.timeout(300)
def test_wide_and_labelled_frames_across_ranks():
    """Two column partitions per rank, and inputs with a named int index / a float index, sharded over 2 ranks."""
    out = _run(_wide_and_labelled_job)
    for r, res in enumerate(out):
        bad = {k: s for k, s in res.items() if s != "ok"}
        assert not bad and len(res) == 27, (r, bad)


def _full_stack_sort_job(rank, ws):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_double
    import modin_b200.pandas as bpd
    from modin_b200 import config

    out = {}
    with cpu_double.installed():
        config.NPartitions.put(2)
        pdf = synth.host_frame(5_003, 2, seed=6, nan_per_64k=5000, key_modulus=13)
        df = bpd.DataFrame(pdf)  # sharded by rank
        for by, asc in (("key", True), ("c0", False)):
            blk = df.sort_values(by, ascending=asc)._query_compiler._modin_frame._partitions[0, 0].get()
            out[(by, asc)] = (blk.index_cols[0].data.numpy().copy(),
                              np.stack([c.data.numpy().astype(np.float64) for c in blk.cols], axis=1))
        blk = df.sort_values("c1", ignore_index=True)._query_compiler._modin_frame._partitions[0, 0].get()
        out["ignore"] = (np.arange(blk.range_start, blk.range_start + blk.nrows),
                         np.stack([c.data.numpy().astype(np.float64) for c in blk.cols], axis=1))
    return out


def test_full_stack_sort_values_across_ranks():
    """The raw-row range shuffle: rank r ends up with the r-th key range, ties in original order."""
    out = _run(_full_stack_sort_job)
    pdf = synth.host_frame(5_003, 2, seed=6, nan_per_64k=5000, key_modulus=13)
    for by, asc in (("key", True), ("c0", False)):
        want = pdf.sort_values(by, ascending=asc, kind="stable")
        idx = np.concatenate([o[(by, asc)][0] for o in out])
        vals = np.concatenate([o[(by, asc)][1] for o in out])
        assert np.array_equal(idx, want.index.to_numpy()), (by, asc)
        assert np.array_equal(vals, want.to_numpy(dtype=np.float64), equal_nan=True), (by, asc)
        assert all(len(o[(by, asc)][0]) > 0 for o in out)
    want = pdf.sort_values("c1", kind="stable", ignore_index=True)
    assert np.array_equal(np.concatenate([o["ignore"][0] for o in out]), want.index.to_numpy())
    assert np.array_equal(np.concatenate([o["ignore"][1] for o in out]), want.to_numpy(dtype=np.float64), equal_nan=True)


def _dedupe_and_packed_merge_sweep_job(rank, ws):
    """Random shapes for the two exchanges written late in round 2: drop_duplicates across ranks (keys of the per-rank
    survivors all-gathered) and the merge on two key columns (keys packed over the ranges of BOTH frames, agreed across
    ranks).  Shards of 0 / 1 / a few rows, all duplicates in one shard, keys of both signs."""
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_double
    import modin_b200.pandas as bpd
    from modin_b200 import config

    out = []
    with cpu_double.installed():
        config.NPartitions.put(2)
        for case, (n, modulus) in enumerate(((0, 5), (1, 5), (2, 1), (7, 3), (50, 4), (333, 40), (2003, 7), (2003, 1500))):
            rng = np.random.RandomState(100 + case)
            pdf = pandas.DataFrame({"key": rng.randint(0, modulus, n).astype(np.int64) * 3 - modulus,
                                    "k2": rng.randint(-2, 3, n).astype(np.int64), "v": rng.randn(n)})
            if n >= 50:
                pdf.loc[: n // 3, "key"] = pdf["key"].iloc[0]  # a long run of one key inside the first shard
            df = bpd.DataFrame(pdf)
            res = {"n": n, "modulus": modulus}
            for keep in ("first", "last"):
                res["dd_" + keep] = df.drop_duplicates(subset=["key"], keep=keep)._to_pandas()
            res["dd_ignore"] = df.drop_duplicates(subset=["k2"], keep="last", ignore_index=True)._to_pandas()
            pairs = sorted({(int(a), int(b)) for a, b in zip(rng.randint(0, modulus, 40) * 3 - modulus, rng.randint(-2, 3, 40))})
            dim = pandas.DataFrame({"key": np.array([p[0] for p in pairs], dtype=np.int64),
                                    "k2": np.array([p[1] for p in pairs], dtype=np.int64), "w": rng.randn(len(pairs))})
            res["dim"] = dim
            if n:
                for how in ("left", "inner"):
                    res["merge_" + how] = df.merge(bpd.DataFrame(dim), on=["key", "k2"], how=how)._to_pandas()
            out.append((pdf, res))
    return out


def test_dedupe_and_packed_merge_sweep_across_three_ranks():
    import pandas

    outs = _run(_dedupe_and_packed_merge_sweep_job, ws=3)
    for rank_out in outs:  # every rank gathers the same job-wide answers
        for pdf, res in rank_out:
            tag = (res["n"], res["modulus"])
            for keep in ("first", "last"):
                want = pdf.drop_duplicates(subset=["key"], keep=keep)
                got = res["dd_" + keep]
                assert list(got.index) == list(want.index) and np.array_equal(got.to_numpy(), want.to_numpy()), (tag, keep)
            want = pdf.drop_duplicates(subset=["k2"], keep="last", ignore_index=True)
            assert list(res["dd_ignore"].index) == list(want.index), tag
            assert np.array_equal(res["dd_ignore"].to_numpy(), want.to_numpy()), tag
            for how in ("left", "inner"):
                if res["n"]:
                    want = pandas.merge(pdf, res["dim"], on=["key", "k2"], how=how)
                    got = res["merge_" + how]
                    assert list(got.columns) == list(want.columns) and got.shape == want.shape, (tag, how)
                    assert np.array_equal(got.to_numpy(dtype=np.float64), want.to_numpy(dtype=np.float64), equal_nan=True), (tag, how)
