"""Where does a dense groupby step spend its time besides the accumulate kernel?  Wall-clock per phase with a
device sync after each (run under gpurun).   python tools/gb_phase_probe.py [log2_rows] [G]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from modin_b200 import _lib, ops, synth  # noqa: E402
from modin_b200.block import DeviceColumn  # noqa: E402


def main():
    n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 27)
    G = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    W = 8
    keys = ops.gen_i64(n, 43, 0, G)
    vals = [ops.gen_f64(n, 42, j) for j in range(W)]
    sync = torch.cuda.synchronize
    phases = {}

    def timed(name, fn):
        sync()
        t0 = time.perf_counter()
        r = fn()
        sync()
        phases.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
        return r

    for it in range(6):
        mm = timed("key_range", lambda: ops.key_range_device([keys]))
        lo, hi, s, d = timed("stats_d2h", lambda: [int(v) for v in mm.tolist()])
        table = timed("create", lambda: ops.GroupTable.dense(lo, hi, W, _lib.GB_SUM))
        timed("accumulate", lambda: table.accumulate(keys, vals))
        ng, ov = timed("ngroups", lambda: table.ngroups())
        out = timed("emit", lambda: table.emit(ng, sort=False))
        timed("close", lambda: table.close())
        del out
        # whole step through the public op for comparison
        sync()
        t0 = time.perf_counter()
        r = ops.hash_aggregate([(keys, vals)], _lib.GB_SUM, G)
        sync()
        phases.setdefault("whole_hash_aggregate", []).append((time.perf_counter() - t0) * 1e3)
        del r
    print(json.dumps({k: round(min(v[1:]), 3) for k, v in phases.items()}))


if __name__ == "__main__":
    main()
