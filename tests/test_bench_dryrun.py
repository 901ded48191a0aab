"""``bench.py``'s GPU arm, dry-run on the numpy device double under gloo (tests/bench_dryrun.py): every rank must
reach every collective at the world sizes the driver runs (a leg only some ranks enter is a hang on the real box --
round 2 found one at 8 ranks this way: a rank whose shard a filter had emptied left Modin's device path), the JSON
line must be well formed and every ``checked`` / ``parity_ok`` true.  Nothing is measured."""

import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("ws", [1, 2, 8])
def test_bench_control_flow_on_the_double(ws):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: the real bench runs")
    script = os.path.join(ROOT, "tests", "bench_dryrun.py")
    args = ["--rows", "400000", "--groups", "5000", "--steps", "3", "--warmup", "1", "--e2e-rows", "80000"]
    env = dict(os.environ, MB200_DRYRUN_HANG_DUMP_S="240")
    if ws == 1:
        cmd = [sys.executable, script] + args
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ws}", "--master-addr",
               "127.0.0.1", "--master-port", str(_free_port()), script, "--gpus", str(ws)] + args  # fmt: skip
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, "rank 0 prints exactly one JSON line"
    j = json.loads(lines[0])
    assert j["n_gpus"] == ws and j["checked"] is True
    assert j["parity_ok"] is (True if ws > 1 else None), j.get("parity_failed")
    legs = j["also"]
    assert len(legs) == 9 and all(a.get("checked") is True for a in legs), [(a["metric"][:40], a.get("checked"), a.get("error")) for a in legs]
    assert j["roofline_groupby"]["checked"] is True
    assert j["e2e"]["checked"] is True and j["e2e_groupby"]["checked"] is True
    for key in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
                "config", "roofline", "e2e", "gpu_launches", "clocks"):  # fmt: skip
        assert key in j, key
    assert "modin.pandas" in j["config"]["api"] or not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "modin"))
