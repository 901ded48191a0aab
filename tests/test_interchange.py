"""The dataframe interchange protocol over device blocks (``modin_b200.interchange``; reference:
``PandasDataframe.__dataframe__`` / ``from_interchange_dataframe``, dataframe.py:4803-4867, and the protocol tests
under modin/tests/interchange/dataframe_protocol/).  Producer: chunks = row partitions, one contiguous buffer per
column, NaN as the float null, the device said in ``__dlpack_device__``.  Consumer: pandas' host buffers are copied
H2D, device buffers are adopted through DLPack without a copy.  On the numpy double the buffers are host memory, so
pandas' own consumer (``pandas.api.interchange.from_dataframe``) can read them -- an independent check of the producer.
"""

import os
import warnings

import numpy as np
import pandas
import pytest

from modin_b200 import synth
from tests.test_alignment_merge import REF, _modin, needs_modin


def _host_frame(n=1003):
    pdf = synth.host_frame(n, 3, seed=8, nan_per_64k=3000, key_modulus=13)
    pdf["flag"] = pdf["c0"] > 0.0
    return pdf


def _frame_of(df):
    return df._query_compiler._modin_frame


def _protocol_checks(pdm, on_gpu):
    from modin_b200 import config
    from modin_b200.interchange import B200ProtocolDataframe, ColumnNullType, DlpackDeviceType, DTypeKind

    pdf = _host_frame()
    df = pdm.DataFrame(pdf)
    fr = _frame_of(df)
    proto = fr.__dataframe__()
    assert isinstance(proto, B200ProtocolDataframe)
    assert proto.num_rows() == len(pdf) and proto.num_columns() == 5 and list(proto.column_names()) == list(pdf.columns)
    assert proto.num_chunks() == fr._partitions.shape[0]
    assert proto.metadata["modin.index"].equals(pdf.index)
    c = proto.get_column_by_name("c1")
    assert c.dtype == (DTypeKind.FLOAT, 64, "g", "=") and c.describe_null == (ColumnNullType.USE_NAN, None)
    assert c.null_count == int(pdf["c1"].isna().sum()) and c.size() == len(pdf) and c.offset == 0
    assert proto.get_column_by_name("key").dtype[:2] == (DTypeKind.INT, 64)
    assert proto.get_column_by_name("key").describe_null[0] == ColumnNullType.NON_NULLABLE
    assert proto.get_column_by_name("flag").dtype[:2] == (DTypeKind.BOOL, 8)
    with pytest.raises(TypeError):
        c.describe_categorical
    buf, bdt = c.get_buffers()["data"]
    assert c.get_buffers()["validity"] is None and c.get_buffers()["offsets"] is None
    assert buf.bufsize == 8 * len(pdf) and bdt == c.dtype
    assert buf.__dlpack_device__()[0] == (DlpackDeviceType.CUDA if on_gpu else DlpackDeviceType.CPU)
    # chunks: the row partitions, then each cut into equal views
    rows = [ch.num_rows() for ch in proto.get_chunks()]
    assert rows == list(fr.row_lengths) and sum(rows) == len(pdf)
    k = proto.num_chunks()
    sub = list(proto.get_chunks(2 * k))
    assert len(sub) == 2 * k and sum(ch.num_rows() for ch in sub) == len(pdf)
    if k > 1:
        with pytest.raises(RuntimeError):
            list(proto.get_chunks(2 * k + 1))  # not a multiple of the number of row partitions
    with pytest.raises(RuntimeError):
        list(proto.get_chunks(0))
    sel = proto.select_columns_by_name(["c2", "key"])
    assert list(sel.column_names()) == ["c2", "key"] and sel.num_rows() == len(pdf)
    # our consumer on our producer: buffers adopted, values and labels intact
    again = type(fr).from_interchange_dataframe(proto)
    assert again.to_pandas().equals(pdf)
    if on_gpu:
        import torch

        one = next(iter(proto.get_chunks()))
        src = one.get_column_by_name("c0").get_buffers()["data"][0]
        assert torch.from_dlpack(src).data_ptr() == src.ptr  # zero-copy hand-over
        blk = again._partitions[0, 0].get()
        assert blk.cols[list(pdf.columns).index("c0")].ptr == src.ptr  # ... and our consumer kept the producer's memory
    else:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            from pandas.api.interchange import from_dataframe

            assert from_dataframe(proto).equals(pdf)  # pandas' own consumer reads the producer
            assert from_dataframe(sel).equals(pdf[["c2", "key"]])
    # consumer on a HOST producer (pandas): H2D copies; refused when copies are not allowed
    host = type(fr).from_interchange_dataframe(pdf)
    assert host.to_pandas().equals(pdf)
    from modin_b200.interchange import blocks_from_dataframe

    with pytest.raises(RuntimeError):
        blocks_from_dataframe(pdf, allow_copy=False)
    with pytest.raises(ValueError):
        blocks_from_dataframe(object())
    with pytest.raises(NotImplementedError):
        type(fr).from_interchange_dataframe(pandas.DataFrame({"s": ["a", "b"]}))
    # foreign row labels ride in the metadata: integers go to the device, anything else stays a host index
    perm = pdf.copy()
    perm.index = np.random.RandomState(0).permutation(len(pdf)) + 50
    assert type(fr).from_interchange_dataframe(_frame_of(pdm.DataFrame(perm)).__dataframe__()).to_pandas().equals(perm)


def test_protocol_through_the_mirror_on_the_double(cpu_device):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked test")
    from modin_b200 import config
    import modin_b200.pandas as bpd

    old = config.NPartitions.get()
    config.NPartitions.put(3)
    try:
        _protocol_checks(bpd, False)
    finally:
        config.NPartitions.put(old)


@needs_modin
def test_protocol_under_real_modin_on_the_double(cpu_device):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked test")
    mpd = _modin(nparts=3)
    _protocol_checks(mpd, False)
    # through Modin's public entry points (modin/pandas/dataframe.py ``__dataframe__``; modin/pandas/io.py:1048 ``from_dataframe``)
    pdf = _host_frame(500)
    mdf = mpd.DataFrame(pdf)
    proto = mdf.__dataframe__()
    assert proto.num_rows() == 500
    from modin.pandas.io import from_dataframe as modin_from_dataframe  # modin.pandas.api.interchange.from_dataframe

    back = modin_from_dataframe(proto)
    from modin_b200.block import DeviceBlock

    assert all(isinstance(p.get(), DeviceBlock) for p in _frame_of(back)._partitions.flatten())
    assert back._to_pandas().equals(pdf)


@pytest.mark.gpu
def test_protocol_on_b200():
    import modin_b200.pandas as bpd

    _protocol_checks(bpd, True)
    if os.path.isdir(os.path.join(REF, "modin")):
        _protocol_checks(_modin(nparts=2), True)
