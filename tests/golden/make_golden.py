"""Generate the golden fixtures from the UNMODIFIED reference (run in the build container only).

    MODIN_ENGINE=python python tests/golden/make_golden.py

Imports modin from /root/reference (read-only mount) with the PandasOnPython engine and
NPartitions=4, after applying the five pandas-3 import shims listed in SURVEY.md §8(c) (the
image ships pandas 3.0.2, the reference pins pandas<2.4; the shims only restore removed pandas
names / swallow removed keyword arguments -- no reference source is modified or copied).  Inputs
are the seeded synthetic frames of ``modin_b200.synth`` (host twin) at small sizes; outputs are
what ``modin.pandas`` returns.  Everything is stored in one ``.npz`` per case under
tests/golden/.  /root/reference does not exist on the GPU box, which is why the vectors are
committed.
"""

import functools
import os
import sys

import numpy as np
import pandas

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def apply_pandas3_shims():
    import pandas.core.series as pcs
    import pandas.io.parsers.base_parser as bp

    def _stub(*a, **k):
        """stub"""
        raise NotImplementedError

    if not hasattr(pandas, "read_gbq"):
        pandas.read_gbq = _stub
    if not hasattr(pcs, "_coerce_method"):
        def _coerce_method(converter):
            def wrapper(self):
                if len(self) == 1:
                    return converter(self.iloc[0])
                raise TypeError(f"cannot convert the series to {converter}")
            wrapper.__name__ = f"__{converter.__name__}__"
            return wrapper
        pcs._coerce_method = _coerce_method
    if not hasattr(bp.ParserBase, "_validate_usecols_arg"):
        bp.ParserBase._validate_usecols_arg = lambda self, usecols: (usecols, None)
    for cls in (pandas.DataFrame, pandas.Series):
        g = cls.groupby
        cls.groupby = functools.wraps(g)(lambda self, *a, axis=0, _g=g, **k: _g(self, *a, **k))
        f = cls.fillna
        cls.fillna = functools.wraps(f)(lambda self, *a, method=None, downcast=None, _f=f, **k: (
            (self.ffill(**{kk: v for kk, v in k.items() if kk in ("axis", "limit", "inplace")}) if method in ("ffill", "pad")
             else self.bfill(**{kk: v for kk, v in k.items() if kk in ("axis", "limit", "inplace")}))
            if method is not None else _f(self, *a, **k)))
    # pandas < 2.4 (the reference's pin, setup.py:49): a SCALAR names a group of a length-1 list of keys.  pandas 3
    # wants a 1-tuple, so the reference's range-partitioning split (dataframe/utils.py:415-420 ``grp.get_group(key)``
    # after ``groupby([codes])``) finds no group at all and every sort returns an empty frame.
    from pandas.core.groupby.groupby import GroupBy

    if not getattr(GroupBy.get_group, "_mb200_shim", False):
        _gg = GroupBy.get_group

        def get_group(self, name, *a, **k):
            try:
                return _gg(self, name, *a, **k)
            except KeyError:
                if not isinstance(name, tuple) and len(self._grouper.groupings) == 1:
                    return _gg(self, (name,), *a, **k)
                raise

        get_group._mb200_shim = True
        GroupBy.get_group = get_group


def third_batch_frames(synth, n, nb, nan, G):
    """Inputs of the ``ext3`` cases (also used by the tests, so generator and checkers cannot drift apart): two frames
    with the same columns -- int64 keys, an int64 column beyond 2**53, float64 values with NaNs, one bool column."""
    frames = []
    for rows, seed, kseed in ((n, 33, 98), (nb, 34, 97)):
        f = synth.host_frame(rows, 3, seed=seed, nan_per_64k=nan, key_modulus=G)
        f["k2"] = synth.gen_i64(rows, kseed, 1, 7) * 5 - 10
        # high and low bits both set, so the cast to float64 has to round (not exactly representable)
        f["big"] = (synth.gen_i64(rows, kseed + 10, 2, 1 << 20) << 42) + synth.gen_i64(rows, kseed + 11, 3, 1 << 20) - (1 << 61)
        f["flag"] = f["c0"] > 0.0
        frames.append(f)
    return frames


def fourth_batch_frames(synth):
    """Inputs of the ``ext4`` cases (label alignment and many-to-many merge; shared with the tests): ``A`` on labels
    0..n-1; ``B`` on a permutation of labels 300..300+n2-1 (partly overlapping ``A``); ``Bp`` on a permutation of
    ``A``'s own labels; a fact / dim pair whose dim keys repeat."""
    n, n2 = 2003, 1801
    rng = np.random.RandomState(7)
    A = synth.host_frame(n, 3, seed=51, nan_per_64k=2000)
    B = synth.host_frame(n2, 3, seed=52, nan_per_64k=2000)
    B.index = rng.permutation(np.arange(300, 300 + n2))
    Bp = synth.host_frame(n, 3, seed=53)
    Bp.index = rng.permutation(n)
    fact = synth.host_frame(3000, 2, seed=42, key_modulus=200, key_seed=43)
    dim = pandas.DataFrame({"key": np.concatenate([rng.permutation(200)[:150], rng.permutation(200)[:80]]).astype(np.int64)})
    dim["d0"] = synth.gen_f64(len(dim), 11, 0)
    dim["d1"] = np.arange(len(dim), dtype=np.int64) * 3 + 1
    dim_u = dim.drop_duplicates("key").rename(columns={"key": "k"})
    return A, B, Bp, fact, dim, dim_u


def fifth_batch_frame(synth):
    """Input of the ``ext5`` cases (sort_values, Fold, Reduce; shared with the tests): float64 columns with NaN --
    also at the very top, in a long run, and a column that is NaN from row 1500 on --, an int64 key with many ties and
    an int64 column without ties."""
    n = 2003
    f = synth.host_frame(n, 4, seed=61, nan_per_64k=3000, key_modulus=23)
    f.iloc[0:4, f.columns.get_loc("c1")] = np.nan
    f.iloc[700:760, f.columns.get_loc("c2")] = np.nan
    f.iloc[1500:, f.columns.get_loc("c3")] = np.nan
    f["u"] = np.random.RandomState(3).permutation(n).astype(np.int64) * 7 - 5000
    return f


def sixth_batch_frames(synth):
    """Inputs of the ``ext6`` cases (merge on SEVERAL key columns; shared with the tests): a fact frame with two int64
    key columns (one with negative values) and a dim frame holding a subset of the key pairs, some of them twice, with
    a float column whose label collides with the fact's, and an int64 column (promoted to float64 by left-join misses)."""
    n = 4001
    rng = np.random.RandomState(17)
    fact = synth.host_frame(n, 2, seed=71, nan_per_64k=1500)
    fact.insert(0, "b", synth.gen_i64(n, 72, 1, 9) - 4)
    fact.insert(0, "a", synth.gen_i64(n, 73, 0, 31))
    pairs = [(a, b) for a in range(0, 34) for b in range(-5, 4)]
    rng.shuffle(pairs)
    pairs = pairs[:170]
    dim = pandas.DataFrame({"a": np.array([p[0] for p in pairs], dtype=np.int64), "b": np.array([p[1] for p in pairs], dtype=np.int64)})
    dim["c0"] = synth.gen_f64(len(dim), 74, 0)
    dim["d"] = synth.gen_f64(len(dim), 75, 0)
    dim["i"] = np.arange(len(dim), dtype=np.int64) * 5 - 7
    dim_dups = pandas.concat([dim, dim.iloc[:45].assign(d=lambda x: x["d"] + 1.0)], ignore_index=True)
    return fact, dim, dim_dups


def seventh_batch_frame(synth):
    """Input of the ``ext7`` cases (float64 group keys; shared with the tests): ~25 distinct half-integers of both
    signs, both zeros, 5 % NaN keys; float64 values with NaN."""
    n = 6007
    rng = np.random.RandomState(23)
    f = synth.host_frame(n, 3, seed=81, nan_per_64k=2500)
    fk = np.round(rng.randn(n) * 3.0, 0) / 2.0
    fk[rng.rand(n) < 0.05] = np.nan
    fk[:4] = [0.0, -0.0, 0.0, -0.0]
    f.insert(0, "fk", fk)
    return f


def main():
    os.environ["MODIN_ENGINE"] = "python"
    apply_pandas3_shims()
    sys.path.insert(0, "/root/reference")
    import warnings

    warnings.filterwarnings("ignore")
    import modin.config as cfg
    import modin.pandas as mpd

    cfg.NPartitions.put(4)
    from modin_b200 import synth

    only = os.environ.get("GOLDEN_ONLY")  # e.g. GOLDEN_ONLY=ext3_ : leave the other (already committed) files alone

    def save(name, **arrays):
        if only and not name.startswith(only):
            return
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
        print("wrote", name, {k: getattr(v, "shape", None) for k, v in arrays.items()})

    def P(x):  # modin -> pandas
        return x._to_pandas() if hasattr(x, "_to_pandas") else x

    # ---- C1-like: Map / TreeReduce on n x 4 float64 with NaNs
    for n, W, nan in ((1000, 4, 0), (4099, 5, 2000)):
        pdf = synth.host_frame(n, W, seed=42, nan_per_64k=nan)
        mdf = mpd.DataFrame(pdf)
        tag = f"frame_n{n}_w{W}_nan{nan}"
        save(
            tag,
            meta=np.array([n, W, 42, nan]),
            abs=P(mdf.abs()).to_numpy(),
            neg=P(-mdf).to_numpy(),
            isna=P(mdf.isna()).to_numpy(),
            fillna=P(mdf.fillna(1.5)).to_numpy(),
            affine=P(mdf * 1.25 + 0.5).to_numpy(),
            rowvec=P(mdf * list(np.arange(1, W + 1) * 0.5) + list(np.arange(W) * 0.25)).to_numpy(),
            sum=P(mdf.sum()).to_numpy(),
            sum_noskip=P(mdf.sum(skipna=False)).to_numpy(),
            sum_mc1=P(mdf.sum(min_count=1)).to_numpy(),
            mean=P(mdf.mean()).to_numpy(),
            min=P(mdf.min()).to_numpy(),
            max=P(mdf.max()).to_numpy(),
            count=P(mdf.count()).to_numpy(),
            lt0=P(mdf < 0.0).to_numpy(),
        )
        # 3-frame a*b+c
        b = synth.host_frame(n, W, seed=7, nan_per_64k=0)
        c = synth.host_frame(n, W, seed=9, nan_per_64k=0)
        mb, mc = mpd.DataFrame(b), mpd.DataFrame(c)
        save(tag + "_fma3", meta=np.array([n, W, 42, nan, 7, 9]), out=P(mdf * mb + mc).to_numpy(),
             sub=P(mdf - mb).to_numpy(), div=P(mdf / mb).to_numpy(), ge=P(mdf >= mb).to_numpy())

    # ---- more registrations: prod, var / std, round, clip; groupby min / max
    for n, W, nan in ((3001, 4, 1500),):
        pdf = synth.host_frame(n, W, seed=77, nan_per_64k=nan)
        mdf = mpd.DataFrame(pdf)
        small = mpd.DataFrame(pdf.iloc[:60] * 1.25)
        save(
            f"ext_n{n}_w{W}_nan{nan}",
            meta=np.array([n, W, 77, nan]),
            prod60=P(small.prod()).to_numpy(),
            var=P(mdf.var()).to_numpy(),
            var_ddof0=P(mdf.var(ddof=0)).to_numpy(),
            std=P(mdf.std()).to_numpy(),
            var_noskip=P(mdf.var(skipna=False)).to_numpy(),
            round2=P(mdf.round(2)).to_numpy(),
            round0=P(mdf.round(0)).to_numpy(),
            round_m1=P((mdf * 100.0).round(-1)).to_numpy(),
            clip=P(mdf.clip(-0.5, 0.75)).to_numpy(),
            clip_lower=P(mdf.clip(lower=0.0)).to_numpy(),
        )
    for n, G, V, nan in ((7001, 97, 3, 4000),):
        pdf = synth.host_frame(n, V, seed=42, nan_per_64k=nan, key_modulus=G, key_seed=43)
        g = mpd.DataFrame(pdf).groupby("key")
        mn, mx = P(g.min()), P(g.max())
        save(f"ext_groupby_n{n}_g{G}_v{V}_nan{nan}", meta=np.array([n, G, V, nan, 42, 43]), keys=mn.index.to_numpy(),
             min=mn.to_numpy(), max=mx.to_numpy())

    # ---- second batch: logical ops, any / all, bool sums, isin, boolean row selection, dropna, multi-key groupby,
    # dictionary aggregation, value_counts, nunique
    for n, nan in ((3001, 6000),):
        pdf = synth.host_frame(n, 3, seed=21, nan_per_64k=nan, key_modulus=9)
        pdf["k2"] = synth.gen_i64(n, 99, 1, 5) * 10 - 20
        mdf = mpd.DataFrame(pdf)
        v = ["c0", "c1", "c2"]
        band = (mdf[v] > 0.0) & (mdf[v] < 1.0)
        sel = P(mdf[mdf["c0"] > 0.5])
        dn = P(mdf.dropna())
        mk = P(mdf.groupby(["key", "k2"]).sum())
        da = P(mdf.groupby("key").agg({"c1": "max", "c0": "sum", "c2": "count"}))
        vc = P(mdf["key"].value_counts())
        save(
            f"ext2_n{n}_nan{nan}",
            meta=np.array([n, 21, nan, 9]),
            band=P(band).to_numpy(),
            bor=P((mdf[v] > 0.5) | (mdf[v] < -0.5)).to_numpy(),
            bxor=P((mdf[v] > 0.0) ^ (mdf[v] > 1.0)).to_numpy(),
            bnot=P(~band).to_numpy(),
            any=P(band.any()).to_numpy(),
            all_true=P((mdf[v] > -100.0).all()).to_numpy(),
            boolsum=P(band.sum()).to_numpy(),
            boolmean=P(band.mean()).to_numpy(),
            isin=P(mdf[["key", "k2"]].isin([3, 7, -20, 30])).to_numpy(),
            sel_index=sel.index.to_numpy(), sel=sel.to_numpy(dtype=np.float64),
            dropna_index=dn.index.to_numpy(), dropna=dn.to_numpy(dtype=np.float64),
            mk_k1=mk.index.get_level_values(0).to_numpy(), mk_k2=mk.index.get_level_values(1).to_numpy(),
            mk_sum=mk.to_numpy(),
            dict_keys=da.index.to_numpy(), dict_agg=da.to_numpy(dtype=np.float64),
            vc_counts=vc.to_numpy(), vc_keys=vc.index.to_numpy(),
            nunique=np.array([mdf["key"].nunique()]),
        )

    # ---- third batch: drop_duplicates, concat, astype, DataFrame.nunique
    for n, nb, nan, G in ((3001, 1501, 3000, 40),):
        pdf, pb = third_batch_frames(synth, n, nb, nan, G)
        mdf, mb = mpd.DataFrame(pdf), mpd.DataFrame(pb)
        num = ["key", "k2", "c0", "c1", "c2"]  # drop_duplicates on the device path carries no bool columns
        arrays = {}
        for keep in ("first", "last"):
            r = P(mdf[num].drop_duplicates(subset=["key"], keep=keep))
            arrays[f"dd_{keep}_index"], arrays[f"dd_{keep}"] = r.index.to_numpy(), r.to_numpy(dtype=np.float64)
        r = P(mdf[num].drop_duplicates(subset=["k2"], keep="last", ignore_index=True))
        arrays["dd_k2_ignore_index"], arrays["dd_k2_ignore"] = r.index.to_numpy(), r.to_numpy(dtype=np.float64)
        r = P(mdf["k2"].drop_duplicates())
        arrays["dd_series_index"], arrays["dd_series"] = r.index.to_numpy(), r.to_numpy()
        for ig in (False, True):
            r = P(mpd.concat([mdf[num], mb[num], mdf[num]], ignore_index=ig))
            arrays[f"cat0_ig{int(ig)}_index"], arrays[f"cat0_ig{int(ig)}"] = r.index.to_numpy(), r.to_numpy(dtype=np.float64)
        r = P(mpd.concat([mdf[num], mdf[["c0", "c1"]].rename(columns={"c0": "x", "c1": "y"})], axis=1))
        arrays["cat1"], arrays["cat1_cols"] = r.to_numpy(dtype=np.float64), np.array(list(r.columns))
        r = P(mdf.astype("float64"))
        assert all(t == np.float64 for t in r.dtypes)
        arrays["astype_f64"] = r.to_numpy()
        r = P(mdf.astype({"key": np.float64, "flag": "int64"}))
        arrays["astype_dict_dtypes"] = np.array([str(t) for t in r.dtypes])
        arrays["astype_dict"] = r.to_numpy(dtype=np.float64)
        r = P(mdf[["key", "k2", "big"]].astype("float64"))
        arrays["astype_big"] = r.to_numpy()
        r = P(mdf[["key", "k2", "big"]].nunique())
        arrays["nunique"], arrays["nunique_cols"] = r.to_numpy(), np.array(list(r.index))
        save(f"ext3_n{n}_nan{nan}", meta=np.array([n, nb, nan, G]), **arrays)

    # ---- fourth batch: row-label alignment (_copartition with reindex, df.py:3709-3848) and many-to-many merge
    A, B, Bp, fact, dim, dim_u = fourth_batch_frames(synth)
    mA, mB, mBp = mpd.DataFrame(A), mpd.DataFrame(B), mpd.DataFrame(Bp)
    arrays = {}
    for name, r in (("add", mA + mB), ("mul_add", mA * mB + mA), ("lt", mA < mBp)):
        r = P(r)
        arrays[name + "_index"], arrays[name] = r.index.to_numpy(), r.to_numpy(dtype=np.float64)
    x = mA.copy()
    x["d"] = mB["c0"]
    arrays["setitem"] = P(x).to_numpy()
    r = P(mA[mBp["c0"] > 0])
    arrays["mask_index"], arrays["mask"] = r.index.to_numpy(), r.to_numpy()
    r = P(mpd.concat([mA, mBp.rename(columns={"c0": "x", "c1": "y", "c2": "z"})], axis=1))
    arrays["cat1_index"], arrays["cat1"] = r.index.to_numpy(), r.to_numpy()
    mf = mpd.DataFrame(fact)
    for how in ("left", "inner"):
        r = P(mf.merge(mpd.DataFrame(dim), on="key", how=how))
        arrays[f"m2m_{how}"], arrays[f"m2m_{how}_cols"] = r.to_numpy(dtype=np.float64), np.array(list(r.columns))
    r = P(mf.merge(mpd.DataFrame(dim_u), left_on="key", right_on="k", how="left"))
    arrays["lr_on"], arrays["lr_on_cols"] = r.to_numpy(dtype=np.float64), np.array(list(r.columns))
    save("ext4_align_m2m", meta=np.array([2003, 1801]), **arrays)

    # ---- fifth batch: sort_values through the reference's range-partitioning shuffle (dataframe.py:2741-2791 ->
    # 2565-2739), the Fold registrations (qc.py:2429-2431, 2809-2810) and the Reduce-registered var / std
    # (qc.py:1155-1156).  The reference samples its pivots with an UNSEEDED ``df.sample`` and sorts every bin with
    # pandas' default (unstable) kind, so the order of rows with EQUAL keys changes from run to run: tie-free keys
    # are stored row for row, the tie-heavy key as (sorted keys, labels) to be compared per run of equal keys.
    F = fifth_batch_frame(synth)
    mF = mpd.DataFrame(F)
    arrays = {}
    for by in ("c0", "c2", "u"):
        for asc in (True, False):
            r = P(mF.sort_values(by, ascending=asc))
            tag = f"sort_{by}_{'asc' if asc else 'desc'}"
            arrays[tag + "_index"], arrays[tag] = r.index.to_numpy(), r.to_numpy(dtype=np.float64)
    for asc in (True, False):
        r = P(mF.sort_values("key", ascending=asc))
        tag = f"sort_key_{'asc' if asc else 'desc'}"
        arrays[tag + "_index"], arrays[tag + "_keys"] = r.index.to_numpy(), r["key"].to_numpy()
    fl = ["c0", "c1", "c2", "c3"]
    for name in ("cumsum", "cummax", "cummin", "ffill"):
        arrays[name] = P(getattr(mF[fl], name)()).to_numpy()
    arrays["cumsum_int"] = P(mF[["key", "u"]].cumsum()).to_numpy()
    arrays["cummax_int"] = P(mF[["key", "u"]].cummax()).to_numpy()
    for name in ("var", "std"):
        for ddof in (1, 0):
            arrays[f"{name}_ddof{ddof}"] = P(getattr(mF[fl], name)(ddof=ddof)).to_numpy()
    save("ext5_sort_fold", meta=np.array([2003]), **arrays)

    # ---- sixth batch: merge on several key columns (merge.py:139-168 is pandas.merge per block, whatever the key is)
    fact, dim, dim_dups = sixth_batch_frames(synth)
    mf = mpd.DataFrame(fact)
    arrays = {}
    for how in ("left", "inner"):
        r = P(mf.merge(mpd.DataFrame(dim), on=["a", "b"], how=how))
        arrays[f"on_{how}"], arrays[f"on_{how}_cols"] = r.to_numpy(dtype=np.float64), np.array(list(r.columns))
        arrays[f"on_{how}_dtypes"] = np.array([str(t) for t in r.dtypes])
        r = P(mf.merge(mpd.DataFrame(dim_dups), on=["a", "b"], how=how))
        arrays[f"m2m_{how}"] = r.to_numpy(dtype=np.float64)
    r = P(mf.merge(mpd.DataFrame(dim.rename(columns={"a": "k"})), left_on=["a", "b"], right_on=["k", "b"], how="left"))
    arrays["lr_on"], arrays["lr_on_cols"] = r.to_numpy(dtype=np.float64), np.array(list(r.columns))
    save("ext6_multikey_merge", meta=np.array([4001, 170]), **arrays)

    # ---- seventh batch: groupby on a float64 key (alg/groupby.py:124-300 is pandas' groupby per block, whatever the key)
    F7 = seventh_batch_frame(synth)
    m7 = mpd.DataFrame(F7)
    arrays = {}
    for agg in ("sum", "count", "mean", "min", "max", "size"):
        r = P(getattr(m7.groupby("fk"), agg)())
        arrays[agg + "_keys"], arrays[agg] = r.index.to_numpy(), np.asarray(r, dtype=np.float64).reshape(len(r), -1)
    r = P(m7.groupby("fk", dropna=False).sum())
    arrays["sum_keepna_keys"], arrays["sum_keepna"] = r.index.to_numpy(), r.to_numpy()
    r = P(m7.groupby("fk", as_index=False).mean())
    arrays["mean_flat"], arrays["mean_flat_cols"] = r.to_numpy(), np.array(list(r.columns))
    save("ext7_float_keys", meta=np.array([6007]), **arrays)

    # ---- C4-like: groupby on int64 key, float64 values (with NaNs)
    for n, G, V, nan in ((5000, 37, 3, 0), (20011, 1500, 8, 3000)):
        pdf = synth.host_frame(n, V, seed=42, nan_per_64k=nan, key_modulus=G, key_seed=43)
        mdf = mpd.DataFrame(pdf)
        g = mdf.groupby("key")
        s, c, m, z = P(g.sum()), P(g.count()), P(g.mean()), P(g.size())
        save(
            f"groupby_n{n}_g{G}_v{V}_nan{nan}",
            meta=np.array([n, G, V, nan, 42, 43]),
            keys=s.index.to_numpy(),
            sum=s.to_numpy(),
            count=c.to_numpy(),
            mean=m.to_numpy(),
            size=z.to_numpy(),
        )

    # ---- C5-like: broadcast merge fact x dim on int64 key
    for n, nd, hit in ((6000, 500, 1.0), (6000, 500, 0.8)):
        fact = synth.host_frame(n, 3, seed=42, key_modulus=nd, key_seed=43)
        rng = np.random.RandomState(5)
        dim_keys = rng.permutation(nd).astype(np.int64)
        if hit < 1.0:
            dim_keys = dim_keys[: int(nd * hit)]
        dim = pandas.DataFrame({"key": dim_keys, "d0": synth.gen_f64(len(dim_keys), 11, 0),
                                "d1": np.arange(len(dim_keys), dtype=np.int64) * 3})
        mf, md = mpd.DataFrame(fact), mpd.DataFrame(dim)
        left = P(mf.merge(md, on="key", how="left"))
        inner = P(mf.merge(md, on="key", how="inner"))
        save(
            f"merge_n{n}_d{nd}_hit{int(hit * 100)}",
            meta=np.array([n, nd, int(hit * 100)]),
            dim_keys=dim_keys,
            left=left.to_numpy(dtype=np.float64),
            left_cols=np.array(list(left.columns)),
            inner=inner.to_numpy(dtype=np.float64),
        )


if __name__ == "__main__":
    main()
