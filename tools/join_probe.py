"""Join probe: dense vs hash dim table, persisting-L2 window variants (run under gpurun).
    python tools/join_probe.py [log2_fact_rows] [ndim]
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
    n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 28)
    nd = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    _lib.check(lib.mb200_gen_i64(keys.data_ptr(), n, 43, 0, 0, nd, None, st))
    rng = np.random.RandomState(5)
    dk = torch.from_numpy(rng.permutation(nd).astype(np.int64)).to(dev)
    pay = torch.randn(nd, dtype=torch.float64, device=dev)
    out = torch.empty(n, dtype=torch.float64, device=dev)
    nm = torch.zeros(1, dtype=torch.int64, device=dev)
    for dense, persist, ordered in (("1", "0", "1"), ("1", "0", "0"), ("0", "0", "0")):
        os.environ["MB200_JOIN_DENSE"] = dense
        os.environ["MB200_JOIN_PERSIST"] = persist
        os.environ["MB200_JOIN_ORDERED"] = ordered
        tab = C.c_void_p()
        _lib.check(lib.mb200_join_build(C.byref(tab), dk.data_ptr(), nd, st))

        def probe():
            _lib.check(lib.mb200_join_probe_gather(tab, keys.data_ptr(), n, 1, _lib.ptr_array([pay.data_ptr()]), _lib.F64,
                                                   _lib.ptr_array([out.data_ptr()]), None if ordered == "1" else nm.data_ptr(), st))

        t = timeit(probe)
        _lib.check(lib.mb200_join_destroy(tab, st))
        print(json.dumps({"dense": dense, "persist": persist, "ordered_payload": ordered, "probe_ms": round(t, 3), "Grows": round(n / t / 1e6, 2)}),
              flush=True)


if __name__ == "__main__":
    main()
