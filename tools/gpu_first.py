"""First-contact GPU check: kernel parity vs numpy/pandas on small inputs + raw bandwidth
numbers for the design decisions (TMA vs LDG reduce, groupby RED variants).  Run under gpurun:
    python tools/gpu_first.py [log2_rows]
Prints one JSON line per measurement; exits non-zero on any parity failure.
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modin_b200 import _lib  # noqa: E402

lib = _lib.load()
_lib.check(lib.mb200_device_check(0))
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
FAIL = []


def cols_ptr(ts):
    return _lib.ptr_array([t.data_ptr() for t in ts])


def bits(x):
    return np.float64(x).view(np.uint64).item()


def run_map(op, ins, out, s0=None, s1=None, dtype=_lib.F64):
    n = ins[0][0].numel()
    W = len(ins[0])
    a = cols_ptr(ins[0])
    b = cols_ptr(ins[1]) if len(ins) > 1 else None
    c = cols_ptr(ins[2]) if len(ins) > 2 else None
    o = cols_ptr(out)
    s0a = _lib.u64_array(s0) if s0 is not None else None
    s1a = _lib.u64_array(s1) if s1 is not None else None
    _lib.check(lib.mb200_map(_lib.OP[op], dtype, W, a, b, c, o, n, s0a, s1a, st))


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]


def report(**kw):
    print(json.dumps(kw), flush=True)


def check(name, ok, detail=""):
    if not ok:
        FAIL.append(name)
    report(check=name, ok=bool(ok), detail=str(detail))


# ------------------------------------------------------------------ parity (small, odd sizes)
def parity(only_groupby=False):
    rng = np.random.RandomState(0)
    for n in (() if only_groupby else (1, 7, 4096, 4097, 100003)):
        W = 3
        A = rng.randn(W, n)
        B = rng.randn(W, n)
        Cc = rng.randn(W, n)
        A[0, ::7] = np.nan
        if n > 3:
            A[1, 3] = np.inf
            A[2, 2] = -0.0
        ta = [torch.from_numpy(A[i].copy()).to(dev) for i in range(W)]
        tb = [torch.from_numpy(B[i].copy()).to(dev) for i in range(W)]
        tc = [torch.from_numpy(Cc[i].copy()).to(dev) for i in range(W)]
        out = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(W)]
        run_map("abs", [ta], out)
        got = np.stack([o.cpu().numpy() for o in out])
        check(f"abs n={n}", np.array_equal(got.view(np.uint64), np.abs(A).view(np.uint64)))
        run_map("fma3", [ta, tb, tc], out)
        got = np.stack([o.cpu().numpy() for o in out])
        ref = A * B + Cc
        check(f"fma3 n={n}", np.array_equal(got.view(np.uint64), ref.view(np.uint64)))
        s0 = [bits(1.5), bits(-2.25), bits(3.0)]
        s1 = [bits(0.1), bits(0.2), bits(0.3)]
        run_map("affine", [ta], out, s0, s1)
        got = np.stack([o.cpu().numpy() for o in out])
        ref = A * np.array([1.5, -2.25, 3.0])[:, None] + np.array([0.1, 0.2, 0.3])[:, None]
        check(f"affine n={n}", np.array_equal(got.view(np.uint64), ref.view(np.uint64)))
        run_map("fillna_s", [ta], out, [bits(9.0)] * W)
        got = np.stack([o.cpu().numpy() for o in out])
        check(f"fillna n={n}", np.array_equal(got.view(np.uint64), np.where(np.isnan(A), 9.0, A).view(np.uint64)))
        outb = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(W)]
        run_map("lt", [ta, tb], outb)
        got = np.stack([o.cpu().numpy() for o in outb])
        check(f"lt n={n}", np.array_equal(got.astype(bool), A < B))
        # unaligned view (scalar path)
        if n > 8:
            va = [t[1:] for t in ta]
            vo = [torch.empty(n, dtype=torch.float64, device=dev)[1:] for _ in range(W)]
            run_map("neg", [va], vo)
            got = np.stack([o.cpu().numpy() for o in vo])
            check(f"neg unaligned n={n}", np.array_equal(got.view(np.uint64), (-A[:, 1:]).view(np.uint64)))
        # reductions, both variants
        scratch = torch.empty(lib.mb200_reduce_scratch_bytes(W), dtype=torch.uint8, device=dev)
        oval = torch.empty(W, dtype=torch.float64, device=dev)
        ocnt = torch.empty(W, dtype=torch.int64, device=dev)
        for variant in (0, 1):
            for opn, ref_fn in (("sum", np.nansum), ("min", np.nanmin), ("max", np.nanmax)):
                _lib.check(lib.mb200_reduce_columns(_lib.RED[opn], _lib.F64, W, cols_ptr(ta), n, 1, oval.data_ptr(),
                                                    ocnt.data_ptr(), scratch.data_ptr(), variant, st))
                g = oval.cpu().numpy()
                gc = ocnt.cpu().numpy()
                with np.errstate(all="ignore"):
                    import warnings
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        r = np.array([ref_fn(A[i]) if (~np.isnan(A[i])).any() else (0.0 if opn == "sum" else np.nan)
                                      for i in range(W)])
                rc = (~np.isnan(A)).sum(axis=1)
                tol = 1e-12 * np.maximum(1.0, np.nansum(np.abs(np.where(np.isinf(A), 0, A)), axis=1))
                okv = np.all((np.abs(g - r) <= tol) | (g == r) | (np.isnan(g) & np.isnan(r)))
                check(f"reduce {opn} v{variant} n={n}", okv and np.array_equal(gc, rc), f"{g} vs {r}")
            # skipna=False
            _lib.check(lib.mb200_reduce_columns(_lib.RED["sum"], _lib.F64, W, cols_ptr(ta), n, 0, oval.data_ptr(),
                                                ocnt.data_ptr(), scratch.data_ptr(), variant, st))
            g = oval.cpu().numpy()
            with np.errstate(all="ignore"):
                r = A.sum(axis=1)
            check(f"reduce sum skipna=False v{variant} n={n}",
                  np.all((np.isnan(g) & np.isnan(r)) | (np.abs(g - r) <= 1e-9) | (g == r)), f"{g} vs {r}")
        # int64 sum/min
        I = rng.randint(-1000, 1000, size=(W, n)).astype(np.int64)
        ti = [torch.from_numpy(I[i].copy()).to(dev) for i in range(W)]
        oi = torch.empty(W, dtype=torch.int64, device=dev)
        for variant in (0, 1):
            _lib.check(lib.mb200_reduce_columns(_lib.RED["sum"], _lib.I64, W, cols_ptr(ti), n, 1, oi.data_ptr(),
                                                ocnt.data_ptr(), scratch.data_ptr(), variant, st))
            check(f"reduce i64 sum v{variant} n={n}", np.array_equal(oi.cpu().numpy(), I.sum(axis=1)))
            _lib.check(lib.mb200_reduce_columns(_lib.RED["min"], _lib.I64, W, cols_ptr(ti), n, 1, oi.data_ptr(),
                                                ocnt.data_ptr(), scratch.data_ptr(), variant, st))
            check(f"reduce i64 min v{variant} n={n}", np.array_equal(oi.cpu().numpy(), I.min(axis=1)))

    # sort
    for n in (() if only_groupby else (2, 255, 256, 5000, 300001)):
        k = rng.randint(-2**62, 2**62, size=n).astype(np.int64)
        k[: n // 3] = rng.randint(-5, 5, size=n // 3)
        pl = np.arange(n, dtype=np.int64)
        tk = torch.from_numpy(k.copy()).to(dev)
        tp = torch.from_numpy(pl.copy()).to(dev)
        sb = lib.mb200_sort_scratch_bytes(n)
        sc = torch.empty(sb, dtype=torch.uint8, device=dev)
        _lib.check(lib.mb200_sort_pairs_i64(tk.data_ptr(), tp.data_ptr(), n, sc.data_ptr(), sb, st))
        order = np.argsort(k, kind="stable")
        check(f"sort n={n}", np.array_equal(tk.cpu().numpy(), k[order]) and np.array_equal(tp.cpu().numpy(), pl[order]))

    # groupby
    import pandas as pd
    for variant in ("0", "1", "2"):
        os.environ["MB200_GB_VARIANT"] = variant
        for n, G, V in ((1, 1, 1), (1000, 10, 3), (100003, 5000, 8), (200000, 150000, 11)):
            keys = rng.randint(-G // 2, G - G // 2, size=n).astype(np.int64)
            vals = rng.randn(V, n)
            vals[0, ::5] = np.nan
            tk = torch.from_numpy(keys).to(dev)
            tv = [torch.from_numpy(vals[i].copy()).to(dev) for i in range(V)]
            tab = C.c_void_p()
            flags = _lib.GB_SUM | _lib.GB_COUNT | _lib.GB_SIZE
            _lib.check(lib.mb200_gb_create(C.byref(tab), G + 8, V, flags, st))
            half = n // 2
            _lib.check(lib.mb200_gb_accumulate(tab, tk.data_ptr(), cols_ptr(tv), half, st))
            _lib.check(lib.mb200_gb_accumulate(tab, tk[half:].data_ptr(), cols_ptr([t[half:] for t in tv]), n - half, st))
            ng = C.c_int64()
            ov = C.c_int()
            _lib.check(lib.mb200_gb_ngroups(tab, C.byref(ng), C.byref(ov), st))
            g = ng.value
            okeys = torch.empty(g, dtype=torch.int64, device=dev)
            osum = [torch.empty(g, dtype=torch.float64, device=dev) for _ in range(V)]
            ocnt = [torch.empty(g, dtype=torch.int64, device=dev) for _ in range(V)]
            osz = torch.empty(g, dtype=torch.int64, device=dev)
            sb = lib.mb200_gb_emit_scratch_bytes(g)
            sc = torch.empty(sb, dtype=torch.uint8, device=dev)
            _lib.check(lib.mb200_gb_emit(tab, g, 1, okeys.data_ptr(), cols_ptr(osum), cols_ptr(ocnt), osz.data_ptr(),
                                         sc.data_ptr(), st))
            torch.cuda.synchronize()
            _lib.check(lib.mb200_gb_destroy(tab, st))
            df = pd.DataFrame({"k": keys, **{f"v{i}": vals[i] for i in range(V)}})
            gb = df.groupby("k")
            rs = gb.sum()
            rc = gb.count()
            rz = gb.size()
            ok = (not ov.value) and g == len(rs) and np.array_equal(okeys.cpu().numpy(), rs.index.values)
            if ok:
                gs = np.stack([o.cpu().numpy() for o in osum], axis=1)
                ok = ok and np.allclose(gs, rs.values, rtol=1e-12, atol=1e-12)
                gc = np.stack([o.cpu().numpy() for o in ocnt], axis=1)
                ok = ok and np.array_equal(gc, rc.values) and np.array_equal(osz.cpu().numpy(), rz.values)
            check(f"groupby v{variant} n={n} G={G} V={V}", ok, f"ngroups {g} vs {len(rs)} overflow {ov.value}")


# ------------------------------------------------------------------ bandwidth
def bandwidth(log2n):
    n = 1 << log2n
    W = 8
    cols = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(W)]
    for i, c in enumerate(cols):
        _lib.check(lib.mb200_gen_f64(c.data_ptr(), n, 42, i, 0, 0, st))
    out = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(W)]
    torch.cuda.synchronize()
    s0 = [bits(1.0000001)] * W
    s1 = [bits(0.5)] * W
    for op, nin in (("abs", 1), ("affine", 1)):
        best, med = timeit(lambda: run_map(op, [cols], out, s0, s1))
        gb = n * W * 16 / 1e9
        report(bench=f"map_{op}", rows=n, W=W, ms_best=best, ms_med=med, GBps=gb / best * 1e3, rows_per_s=n / best * 1e3)
    # torch copy reference (same bytes as abs)
    big_a = torch.empty(n * W, dtype=torch.float64, device=dev)
    big_b = torch.empty(n * W, dtype=torch.float64, device=dev)
    best, med = timeit(lambda: big_b.copy_(big_a))
    report(bench="torch_copy", ms_best=best, GBps=n * W * 16 / 1e9 / best * 1e3)
    del big_a, big_b
    # 3-frame fma at W=4 (memory)
    best, med = timeit(lambda: run_map("fma3", [cols[:2], cols[2:4], cols[4:6]], out[:2]))
    report(bench="map_fma3", rows=n, W=2, ms_best=best, GBps=n * 2 * 32 / 1e9 / best * 1e3)
    # reductions
    scratch = torch.empty(lib.mb200_reduce_scratch_bytes(W), dtype=torch.uint8, device=dev)
    oval = torch.empty(W, dtype=torch.float64, device=dev)
    ocnt = torch.empty(W, dtype=torch.int64, device=dev)
    cp = cols_ptr(cols)
    for variant in (0, 1):
        for opn in ("sum", "min"):
            def f():
                _lib.check(lib.mb200_reduce_columns(_lib.RED[opn], _lib.F64, W, cp, n, 1, oval.data_ptr(),
                                                    ocnt.data_ptr(), scratch.data_ptr(), variant, st))
            best, med = timeit(f)
            report(bench=f"reduce_{opn}_v{variant}", rows=n, W=W, ms_best=best, ms_med=med,
                   GBps=n * W * 8 / 1e9 / best * 1e3, rows_per_s=n / best * 1e3)
    ref = torch.stack([c.sum() for c in cols]).cpu().numpy()
    f()
    got = oval.cpu().numpy()
    del out


def bandwidth_groupby(log2n):
    n = 1 << log2n
    W = 8
    cols = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(W)]
    for i, c in enumerate(cols):
        _lib.check(lib.mb200_gen_f64(c.data_ptr(), n, 42, i, 0, 0, st))
    cp = cols_ptr(cols)
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    for G in (1000, 1_000_000):
        _lib.check(lib.mb200_gen_i64(keys.data_ptr(), n, 43, 0, 0, G, st))
        for variant in ("0", "1", "2"):
            os.environ["MB200_GB_VARIANT"] = variant
            tab = C.c_void_p()
            _lib.check(lib.mb200_gb_create(C.byref(tab), G + 16, W, _lib.GB_SUM, st))

            def g():
                _lib.check(lib.mb200_gb_accumulate(tab, keys.data_ptr(), cp, n, st))
            best, med = timeit(g, iters=3, warm=1)
            ng = C.c_int64()
            ov = C.c_int()
            _lib.check(lib.mb200_gb_ngroups(tab, C.byref(ng), C.byref(ov), st))
            report(bench=f"groupby_sum_warm_env{variant}", rows=n, G=G, V=W, ngroups=ng.value, overflow=ov.value,
                   ms_best=best, ms_med=med, GBps=n * 72 / 1e9 / best * 1e3, rows_per_s=n / best * 1e3)
            _lib.check(lib.mb200_gb_destroy(tab, st))

            # cold: a fresh table per pass (create + memset + all inserts + accumulate), as in a real groupby
            def cold():
                t2 = C.c_void_p()
                _lib.check(lib.mb200_gb_create(C.byref(t2), G + 16, W, _lib.GB_SUM, st))
                _lib.check(lib.mb200_gb_accumulate(t2, keys.data_ptr(), cp, n, st))
                _lib.check(lib.mb200_gb_destroy(t2, st))
            best, med = timeit(cold, iters=3, warm=1)
            report(bench=f"groupby_sum_cold_env{variant}", rows=n, G=G, V=W, ms_best=best, ms_med=med,
                   GBps=n * 72 / 1e9 / best * 1e3, rows_per_s=n / best * 1e3)


if __name__ == "__main__":
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
    what = sys.argv[2] if len(sys.argv) > 2 else "all"
    if what in ("all", "parity"):
        parity()
    if what == "gbparity":
        parity(only_groupby=True)
    if what in ("all", "bw"):
        bandwidth(log2n)
    if what in ("all", "gbbw"):
        bandwidth_groupby(log2n)
    report(done=True, failures=FAIL, launches=lib.mb200_launch_count())
    sys.exit(1 if FAIL else 0)
