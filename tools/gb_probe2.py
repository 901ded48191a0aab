"""Groupby probe: persisting-L2 window on/off x kernel variant, fresh table per pass (run under gpurun).
    python tools/gb_probe2.py [log2_rows]
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


def timeit(fn, iters=3, warm=1):
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
    a, b = C.c_int(), C.c_int()
    _lib.check(lib.mb200_l2_persist_info(C.byref(a), C.byref(b)))
    print(json.dumps({"max_persist_bytes": a.value, "max_window_bytes": b.value}), flush=True)
    cols = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(W)]
    for i, c in enumerate(cols):
        _lib.check(lib.mb200_gen_f64(c.data_ptr(), n, 42, i, 0, 0, st))
    cp = _lib.ptr_array([c.data_ptr() for c in cols])
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    for G in (1_000_000, 2_000_000, 4_000_000):
        _lib.check(lib.mb200_gen_i64(keys.data_ptr(), n, 43, 0, 0, G, st))
        for variant in ("0", "1"):
            for persist in ("1", "0"):
                os.environ["MB200_GB_VARIANT"] = variant
                os.environ["MB200_GB_PERSIST"] = persist
                tab = C.c_void_p()
                _lib.check(lib.mb200_gb_create(C.byref(tab), G + 16, W, _lib.GB_SUM, st))

                def warm():
                    _lib.check(lib.mb200_gb_accumulate(tab, keys.data_ptr(), cp, n, st))

                tw = timeit(warm)
                _lib.check(lib.mb200_gb_destroy(tab, st))

                def cold():
                    t2 = C.c_void_p()
                    _lib.check(lib.mb200_gb_create(C.byref(t2), G + 16, W, _lib.GB_SUM, st))
                    _lib.check(lib.mb200_gb_accumulate(t2, keys.data_ptr(), cp, n, st))
                    _lib.check(lib.mb200_gb_destroy(t2, st))

                tc = timeit(cold)
                print(json.dumps({"G": G, "variant": variant, "persist": persist, "warm_ms": round(tw, 3),
                                  "cold_ms": round(tc, 3), "cold_Grows": round(n / tc / 1e6, 2),
                                  "cold_frac": round(n * 72 / tc / 1e6 / 6477.4, 3)}), flush=True)


if __name__ == "__main__":
    main()
