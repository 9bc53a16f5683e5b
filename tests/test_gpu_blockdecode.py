"""Block decode on the device end to end: LZ4-compressed marshalled vectors -> MO_XCALL_LZ4_DECODE -> MO_XCALL_VECTOR_UNMARSHAL ->
ordinary resident vectors that the operator entry points take.  The marshalled layout is Vector.MarshalBinary's (pkg/container/vector/vector.go:718-764,
restated in matrixone_b200/vector.py::marshal_vector); the compressed bytes come from liblz4 (pyarrow lz4_raw)."""
import numpy as np
import pyarrow as pa
import pytest

from matrixone_b200 import capi, ops
from matrixone_b200.vector import DeviceBuffer, Vector, bitmap_from_bools, marshal_vector, varlena_column, xcall

pytestmark = pytest.mark.gpu


def _decode_on_device(blobs):
    comps = [pa.compress(b, codec="lz4_raw", asbytes=True) for b in blobs]
    src = np.frombuffer(b"".join(comps), dtype=np.uint8)
    desc = np.zeros((len(blobs), 4), dtype=np.int64); so = do = 0
    offs = []
    for i, (c, b) in enumerate(zip(comps, blobs)):
        desc[i] = (so, len(c), do, len(b)); offs.append(do)
        so += len(c); do += (len(b) + 15) // 16 * 16 + 16            # every decoded block starts 16-byte aligned, with slack behind it
    dst = DeviceBuffer(do + 64)
    dsrc = DeviceBuffer.from_numpy(src); ddesc = DeviceBuffer.from_numpy(desc.reshape(-1))
    xcall(capi.XCALL_LZ4_DECODE, [Vector(data_ptr=dst.ptr, data_nbytes=dst.nbytes, length=1), Vector(data_ptr=dsrc.ptr, data_nbytes=src.nbytes, length=1),
                                  Vector(data_ptr=ddesc.ptr, data_nbytes=desc.nbytes, length=1)], len(blobs))
    return dst, offs


@pytest.mark.parametrize("dtype,oid", [(np.int64, capi.T_INT64), (np.float64, capi.T_FLOAT64), (np.int32, capi.T_INT32), (np.uint8, capi.T_UINT8)])
@pytest.mark.parametrize("n", [1, 63, 8192])
def test_fixed_width_block_with_nulls_becomes_a_resident_column(gpu, dtype, oid, n):
    rng = np.random.default_rng(n + oid)
    col = (rng.integers(0, 100, n)).astype(dtype)
    null = rng.random(n) < 0.2
    words = bitmap_from_bools(null)
    blob = marshal_vector(oid, col, n, nulls=words)
    dst, offs = _decode_on_device([blob])
    data = DeviceBuffer(col.nbytes + 16); nulls = DeviceBuffer(len(words) * 8)
    view = ops.vector_unmarshal_device(dst.view(len(blob) + 8, offs[0]), len(blob), data, None, nulls)
    assert (view.vclass, view.oid, view.length, view.data_len, view.area_len, view.size) == (0, oid, n, col.nbytes, 0, col.itemsize)
    assert view.null_count == int(null.sum()) and view.bad == 0
    assert (data.to_numpy(dtype, n) == col).all()
    if null.any():
        assert (nulls.to_numpy(np.uint64, len(words)) == words).all()
    else:
        assert not nulls.to_numpy(np.uint64, len(words)).any()
    if dtype == np.int64:        # the resident column feeds an operator directly: SUM over the non-NULL rows
        got = ops.agg_state_device  # noqa: F841  (the device entry point exists; use the host-result form below)
        res = np.zeros(1, dtype=np.int64); rn = np.zeros(1, dtype=np.uint64)
        xcall(capi.XCALL_AGG(capi.AGG_SUM, capi.T_INT64), [Vector(data=res, nulls=rn, length=1),
                                                            Vector(data_ptr=data.ptr, data_nbytes=col.nbytes, nulls_ptr=nulls.ptr, length=n)], n)
        assert res[0] == int(col[~null].sum())
    for b in (dst, data, nulls):
        b.free()


def test_varlena_block_and_several_blocks_per_call(gpu):
    rng = np.random.default_rng(4)
    rows = [rng.integers(0, 256, int(l), dtype=np.uint8) for l in rng.integers(0, 60, 500)]
    cells, area = varlena_column(rows, dtype=np.uint8)
    blob_v = marshal_vector(capi.T_INT64 + 40, cells, 500, area=area.tobytes(), size=24)       # any varlena oid: the layout does not depend on it
    col = rng.standard_normal(8192)
    blob_f = marshal_vector(capi.T_FLOAT64, col, 8192, sorted_flag=True)
    dst, offs = _decode_on_device([blob_v, blob_f])
    d1 = DeviceBuffer(cells.nbytes + 16); a1 = DeviceBuffer(area.nbytes + 16)
    v1 = ops.vector_unmarshal_device(dst.view(len(blob_v) + 8, offs[0]), len(blob_v), d1, a1, None)
    assert v1.length == 500 and v1.data_len == cells.nbytes and v1.area_len == area.nbytes and v1.null_count == 0
    assert (d1.to_numpy(np.uint8, cells.nbytes) == cells).all() and (a1.to_numpy(np.uint8, area.nbytes) == area).all()
    d2 = DeviceBuffer(col.nbytes)
    v2 = ops.vector_unmarshal_device(dst.view(len(blob_f) + 8, offs[1]), len(blob_f), d2, None, None)
    assert v2.sorted == 1 and (d2.to_numpy(np.float64, 8192) == col).all()


def test_malformed_vector_bytes_fail(gpu):
    blob = bytearray(marshal_vector(capi.T_INT64, np.arange(100, dtype=np.int64), 100))
    blob[21:25] = np.uint32(10_000).tobytes()            # dataLen runs past the buffer
    src = DeviceBuffer.from_numpy(np.frombuffer(bytes(blob) + bytes(16), np.uint8))
    data = DeviceBuffer(1024)
    with pytest.raises(capi.MoError):
        ops.vector_unmarshal_device(src, len(blob), data, None, None)
    good = marshal_vector(capi.T_INT64, np.arange(100, dtype=np.int64), 100)
    src2 = DeviceBuffer.from_numpy(np.frombuffer(good + bytes(16), np.uint8))
    with pytest.raises(capi.MoError):
        ops.vector_unmarshal_device(src2, len(good), DeviceBuffer(64), None, None)      # data buffer too small
