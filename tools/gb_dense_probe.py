"""Groupby probe: dense (direct-addressed) vs hash table, fresh table per pass, through the C ABI
(run under gpurun).   python tools/gb_dense_probe.py [log2_rows]
"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modin_b200 import _lib  # noqa: E402

lib = _lib.load()
_lib.check(lib.mb200_device_check(0))
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, iters=4, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def main():
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 27
    n, W = 1 << log2n, 8
    cols = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(W)]
    for i, c in enumerate(cols):
        _lib.check(lib.mb200_gen_f64(c.data_ptr(), n, 42, i, 0, 0, st))
    cp = _lib.ptr_array([c.data_ptr() for c in cols])
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    mm = torch.empty(4, dtype=torch.int64, device=dev)
    Gs = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else (16, 256, 1024, 3000, 4000, 16384, 65_536, 1_000_000)
    kinds = sys.argv[3].split(',') if len(sys.argv) > 3 else ("hash", "dense", "dense_nosmem")
    skew = len(sys.argv) > 4 and sys.argv[4] == "skew"
    for G in Gs:
        _lib.check((lib.mb200_gen_i64_skew if skew else lib.mb200_gen_i64)(keys.data_ptr(), n, 43, 0, 0, G, st))

        def krange():
            _lib.check(lib.mb200_key_range(keys.data_ptr(), n, mm.data_ptr(), 1, st))

        tk = timeit(krange)
        lo, hi, sampled, dup = lo_hi_stats = mm.tolist()
        print(json.dumps({"G": G, "skew": skew, "key_stats": lo_hi_stats, "dup_share": round(dup / max(sampled, 1), 4)}), flush=True)
        for variant in ("0",):
            os.environ["MB200_GB_VARIANT"] = variant
            for kind in kinds:
                os.environ["MB200_GB_SMEM"] = "0" if kind == "dense_nosmem" else "1"

                def cold():
                    t2 = C.c_void_p()
                    if kind != "hash":
                        _lib.check(lib.mb200_gb_create_dense(C.byref(t2), 0, G - 1, W, _lib.GB_SUM, None, None, None, None, st))
                    else:
                        _lib.check(lib.mb200_gb_create(C.byref(t2), G + 16, W, _lib.GB_SUM, st))
                    if kind == "dense_hot":
                        _lib.check(lib.mb200_gb_hint_skew(t2, 1))
                    _lib.check(lib.mb200_gb_accumulate(t2, keys.data_ptr(), cp, n, st))
                    _lib.check(lib.mb200_gb_destroy(t2, st))

                tc = timeit(cold)
                tot = tc + (tk if kind != "hash" else 0.0)
                print(json.dumps({"G": G, "variant": variant, "table": kind, "accumulate_ms": round(tc, 3),
                                  "key_range_ms": round(tk, 3), "Grows_incl_prepass": round(n / tot / 1e6, 2),
                                  "frac_incl_prepass": round(n * 72 / tot / 1e6 / 6477.4, 3)}), flush=True)


if __name__ == "__main__":
    main()
