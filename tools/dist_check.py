"""Multi-GPU parity check, run under torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        tools/dist_check.py

Every rank builds the SAME seeded host frame, ingests only its row shard (pm.from_pandas shards by
rank), runs the hot-path templates through the public API and compares the gathered results with
the CPU oracle.  Exit code 0 = all checks passed on every rank.
"""
import math
import os
import sys

import numpy as np
import pandas

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from modin_b200 import config, dist, synth  # noqa: E402
import modin_b200.pandas as bpd  # noqa: E402
from oracle import reference_path as orc  # noqa: E402

EPS = 2.0**-53


def exact(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))).all())


def close(a, b, abs_sum, n):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    tol = 4 * max(1.0, math.log2(n)) * EPS * np.asarray(abs_sum, dtype=np.float64) + 1e-300
    return a.shape == b.shape and bool(((np.abs(a - b) <= tol) | (a == b) | (np.isnan(a) & np.isnan(b))).all())


def main():
    assert dist.init_from_env("nccl"), "run under torchrun with WORLD_SIZE > 1"
    rank, ws = dist.rank(), dist.world_size()
    config.NPartitions.put(2)
    fails = []

    def check(name, ok):
        if not ok:
            fails.append(name)
        if rank == 0:
            print(("PASS " if ok else "FAIL ") + name, flush=True)

    n, W, G = 200_003, 5, 4_001
    pdf = synth.host_frame(n, W, seed=42, nan_per_64k=700, key_modulus=G)
    vals = pdf.drop(columns="key")
    df = bpd.DataFrame(pdf)  # this rank's shard only
    lo, hi = dist.shard_bounds(n)
    check("local shard length", len(df) == hi - lo)
    dv = df[[f"c{i}" for i in range(W)]]
    NP = 4

    check("map abs (gathered, bit-exact)", exact(dv.abs()._to_pandas().to_numpy(), orc.df_abs(vals, NP).to_numpy()))
    check("fused a*b+c (gathered, bit-exact)",
          exact((dv * 1.25 + 0.5)._to_pandas().to_numpy(), orc.a_mul_b_add_c(vals, 1.25, 0.5, NP).to_numpy()))
    abs_sum = np.nansum(np.abs(vals.to_numpy()), axis=0)
    check("tree_reduce sum (all_reduce)", close(dv.sum().to_numpy(), orc.df_sum(vals, NP).to_numpy(), abs_sum, n))
    check("tree_reduce count", exact(dv.count().to_numpy(), orc.df_count(vals, NP).to_numpy()))
    check("tree_reduce min", exact(dv.min().to_numpy(), orc.df_min(vals, NP).to_numpy()))
    check("tree_reduce max", exact(dv.max().to_numpy(), orc.df_max(vals, NP).to_numpy()))
    cnt = np.maximum(orc.df_count(vals, NP).to_numpy(), 1)
    check("tree_reduce mean", close(dv.mean().to_numpy(), orc.df_mean(vals, NP).to_numpy(), abs_sum / cnt, n))

    from modin_b200 import config as _cfg

    for dense, tag in ((True, "dense table + collectives"), (False, "hash tables + range exchange")):
        _cfg.GroupbyDenseKeys.put(dense)
        g = df.groupby("key")
        want = orc.groupby_reduce(pdf, "key", "sum", NP)
        got_local = g.sum()
        check(f"[{tag}] groupby: every rank owns a non-empty key range", len(got_local) > 0)
        got = got_local._to_pandas()  # gathers the per-rank key ranges in rank order
        gabs = vals.abs().groupby(pdf["key"]).sum().to_numpy()
        check(f"[{tag}] groupby keys globally sorted & complete", np.array_equal(got.index.to_numpy(), want.index.to_numpy()))
        check(f"[{tag}] groupby sum", close(got.to_numpy(), want.to_numpy(), gabs, n))
        check(f"[{tag}] groupby count", exact(g.count()._to_pandas().to_numpy(), orc.groupby_reduce(pdf, "key", "count", NP).to_numpy()))
        check(f"[{tag}] groupby size", exact(g.size()._to_pandas().to_numpy(), orc.groupby_reduce(pdf, "key", "size", NP).to_numpy()))
        gc = np.maximum(orc.groupby_reduce(pdf, "key", "count", NP).to_numpy(), 1)
        check(f"[{tag}] groupby mean", close(g.mean()._to_pandas().to_numpy(), orc.groupby_reduce(pdf, "key", "mean", NP).to_numpy(),
                                    gabs / gc, n))
        for agg in ("min", "max"):
            check(f"[{tag}] groupby {agg}", exact(getattr(g, agg)()._to_pandas().to_numpy(),
                                                 orc.groupby_reduce(pdf, "key", agg, NP).to_numpy()))
    _cfg.GroupbyDenseKeys.put(True)

    rng = np.random.RandomState(5)
    dim_keys = rng.permutation(G).astype(np.int64)[: int(G * 0.9)]
    dim = pandas.DataFrame({"key": dim_keys, "d0": synth.gen_f64(len(dim_keys), 11, 0)})
    mdim = bpd.DataFrame(dim)  # sharded too; combine() all-gathers it
    left = df.merge(mdim, on="key", how="left")._to_pandas()
    wl = orc.broadcast_merge(pdf, dim, "key", "left", NP)
    check("merge left (dim all_gathered)", list(left.columns) == list(wl.columns) and
          exact(left.to_numpy(dtype=np.float64), wl.to_numpy(dtype=np.float64)))
    inner = df.merge(mdim, on="key", how="inner")._to_pandas()
    wi = orc.broadcast_merge(pdf, dim, "key", "inner", NP)
    check("merge inner", exact(inner.to_numpy(dtype=np.float64), wi.to_numpy(dtype=np.float64)))

    t = torch.tensor([len(fails)], dtype=torch.int64, device="cuda")
    torch.distributed.all_reduce(t)
    if rank == 0:
        print(f"dist_check: {'OK' if int(t.item()) == 0 else 'FAILED'} on {ws} ranks", flush=True)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
    sys.exit(0 if int(t.item()) == 0 else 1)


if __name__ == "__main__":
    main()
