"""Round-2 groupby probe (run under gpurun): fresh table per pass at the C ABI, 2^LOG2 rows x 8 f64 values.

    python tools/gb_r02_probe.py [log2_rows]

Prints one JSON line per (table kind, key distribution, persisting-L2 window on/off): ms per pass (create +
accumulate + destroy, what a real groupby pays), G rows/s and the fraction of the measured HBM peak on the 72 B/row
algorithmic traffic.  Extra knobs are passed through the environment (MB200_GB_*).
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
PEAK = 6477.4
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, iters=5, warm=2):
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
    ts.sort()
    return ts[len(ts) // 2]


def main():
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 27
    n, W, G = 1 << log2n, 8, 1_000_000
    cols = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(W)]
    for i, c in enumerate(cols):
        _lib.check(lib.mb200_gen_f64(c.data_ptr(), n, 42, i, 0, 0, st))
    cp = _lib.ptr_array([c.data_ptr() for c in cols])
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    stats = torch.empty(4, dtype=torch.int64, device=dev)
    only = os.environ.get("PROBE_ONLY", "")
    for dist_name in ("uniform", "skew"):
        gen = lib.mb200_gen_i64_skew if dist_name == "skew" else lib.mb200_gen_i64
        _lib.check(gen(keys.data_ptr(), n, 43, 0, 0, G, stats.data_ptr(), st))
        lo, hi, sampled, dup = stats.tolist()
        skewed = sampled >= 1024 and dup > 0.05 * sampled
        for kind in ("dense", "hash"):
            if only and only not in f"{kind}-{dist_name}":
                continue
            for persist in ("1", "0"):
                os.environ["MB200_GB_PERSIST"] = persist

                def one_pass():
                    tab = C.c_void_p()
                    if kind == "dense":
                        _lib.check(lib.mb200_gb_create_dense(C.byref(tab), lo, hi, W, _lib.GB_SUM, None, None, None, None, st))
                    else:
                        _lib.check(lib.mb200_gb_create(C.byref(tab), G + 16, W, _lib.GB_SUM, st))
                    _lib.check(lib.mb200_gb_hint_skew(tab, 1 if skewed else 0))
                    _lib.check(lib.mb200_gb_accumulate(tab, keys.data_ptr(), cp, n, st))
                    _lib.check(lib.mb200_gb_destroy(tab, st))

                ms = timeit(one_pass)
                print(json.dumps({"rows": n, "G": G, "keys": dist_name, "table": kind, "persist": persist,
                                  "skew_flag": bool(skewed), "ms": round(ms, 3), "Grows_s": round(n / ms / 1e6, 2),
                                  "frac_hbm": round(n * 72 / ms / 1e6 / PEAK, 3),
                                  "env": {k: v for k, v in os.environ.items() if k.startswith("MB200_GB_") and k != "MB200_GB_PERSIST"}}),
                      flush=True)
    os.environ.pop("MB200_GB_PERSIST", None)


if __name__ == "__main__":
    main()
