"""Same-box A/B of two builds of libmodin_b200 (groupby accumulate, warm + cold)."""
import ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from modin_b200 import _lib
HERE = os.path.dirname(os.path.abspath(__file__))
dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
def load(path):
    lib = C.CDLL(path)
    for name, (res, args) in _lib._SIGNATURES.items():
        if hasattr(lib, name):
            fn = getattr(lib, name); fn.restype = res; fn.argtypes = args
    return lib
libs = {"new": load(_lib.LIB_PATH), "old": load(os.path.join(HERE, "libold.so"))}
def timeit(fn, iters=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
n, W = 1 << int(sys.argv[1] if len(sys.argv) > 1 else 27), 8
cols = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(W)]
for i, c in enumerate(cols): libs["new"].mb200_gen_f64(c.data_ptr(), n, 42, i, 0, 0, st)
cp = _lib.ptr_array([c.data_ptr() for c in cols]); keys = torch.empty(n, dtype=torch.int64, device=dev)
for G in (65536, 1_000_000):
    libs["new"].mb200_gen_i64(keys.data_ptr(), n, 43, 0, 0, G, st)
    for rep in range(2):
      for name, lib in libs.items():
        for variant in ("0", "1"):
            os.environ["MB200_GB_VARIANT"] = variant; os.environ["MB200_GB_POLICY"] = "none"
            tab = C.c_void_p(); assert lib.mb200_gb_create(C.byref(tab), G + 16, W, 1, st) == 0
            def warm(): assert lib.mb200_gb_accumulate(tab, keys.data_ptr(), cp, n, st) == 0
            tw = timeit(warm); lib.mb200_gb_destroy(tab, st)
            def cold():
                t2 = C.c_void_p(); lib.mb200_gb_create(C.byref(t2), G + 16, W, 1, st)
                lib.mb200_gb_accumulate(t2, keys.data_ptr(), cp, n, st); lib.mb200_gb_destroy(t2, st)
            tc = timeit(cold)
            print(json.dumps({"lib": name, "G": G, "variant": variant, "warm_ms": round(tw, 3), "cold_ms": round(tc, 3)}), flush=True)
