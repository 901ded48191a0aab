"""GPU parity of boolean row selection / dropna (``df[mask]``): mask -> ranked compaction -> one gather per column.

Written after this round's GPU budget was spent, so it has only run on the CPU device double so far
(tests/test_host_stack_cpu.py::test_boolean_row_selection_and_dropna).  Every kernel it reaches is covered by
other GPU tests (bool widening and logical ops: test_boolean_pipelines_on_device; compaction and gather: the
inner-merge tests); the file sorts last so that a surprise here cannot hide the rest of the suite behind ``-x``.
"""

import numpy as np
import pytest

from modin_b200 import synth

pytestmark = pytest.mark.gpu


def _exact(got, want, what):
    g, w = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert g.shape == w.shape, f"{what}: shape {g.shape} vs {w.shape}"
    assert ((g.view(np.uint64) == w.view(np.uint64)) | (np.isnan(g) & np.isnan(w))).all(), what


def test_boolean_row_selection_and_dropna_on_device():
    import modin_b200.pandas as bpd
    from modin_b200 import config

    old = config.NPartitions.get()
    config.NPartitions.put(4)
    try:
        pdf = synth.host_frame(100_003, 3, seed=21, nan_per_64k=6000, key_modulus=9)
        df = bpd.DataFrame(pdf)
        for got, want in (
            (df[df["c0"] > 0.5], pdf[pdf["c0"] > 0.5]),
            (df[(df["c0"] > 0.0) & (df["c1"] < 0.0)], pdf[(pdf["c0"] > 0.0) & (pdf["c1"] < 0.0)]),
            (df[df["key"] == 3], pdf[pdf["key"] == 3]),
            (df[df["c2"] > 100.0], pdf[pdf["c2"] > 100.0]),
            (df.dropna(), pdf.dropna()),
            (df.dropna(how="all", subset=["c0", "c1"]), pdf.dropna(how="all", subset=["c0", "c1"])),
        ):
            g = got._to_pandas()
            assert np.array_equal(g.index.to_numpy(), want.index.to_numpy()) and list(g.columns) == list(want.columns)
            _exact(g.to_numpy(dtype=np.float64), want.to_numpy(dtype=np.float64), "row selection")
        sel = df[df["c0"] > 0.0]
        want = pdf[pdf["c0"] > 0.0]
        assert np.allclose(sel.sum().to_numpy(), want.sum().to_numpy(), rtol=0, atol=1e-8)
        _exact(sel.groupby("key").count()._to_pandas().to_numpy(), want.groupby("key").count().to_numpy(), "count after filter")
    finally:
        config.NPartitions.put(old)
