import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from modin_b200 import _lib
HERE = os.path.dirname(os.path.abspath(__file__))
path = _lib.LIB_PATH if os.environ.get("AB_LIB", "new") == "new" else os.path.join(HERE, "libold.so")
lib = C.CDLL(path)
for name, (res, args) in _lib._SIGNATURES.items():
    if hasattr(lib, name):
        fn = getattr(lib, name); fn.restype = res; fn.argtypes = args
dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
n, W, G = 1 << 27, 8, int(os.environ.get("AB_G", "1000000"))
cols = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(W)]
for i, c in enumerate(cols): lib.mb200_gen_f64(c.data_ptr(), n, 42, i, 0, 0, st)
cp = _lib.ptr_array([c.data_ptr() for c in cols]); keys = torch.empty(n, dtype=torch.int64, device=dev)
lib.mb200_gen_i64(keys.data_ptr(), n, 43, 0, 0, G, st)
os.environ["MB200_GB_VARIANT"] = os.environ.get("AB_VARIANT", "1" if os.environ.get("AB_LIB", "new") == "new" else "0")
os.environ["MB200_GB_POLICY"] = "none"
tab = C.c_void_p(); lib.mb200_gb_create(C.byref(tab), G + 16, W, 1, st)
for _ in range(3): lib.mb200_gb_accumulate(tab, keys.data_ptr(), cp, n, st)
torch.cuda.synchronize(); print("ok")
