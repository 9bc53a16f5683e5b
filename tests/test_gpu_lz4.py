"""LZ4 block decode on the GPU (csrc/lz4.cu, MO_XCALL_LZ4_DECODE) against liblz4's own output (pyarrow lz4_raw) and the oracle decoder:
byte-exact, many blocks per call, malformed blocks fail the call."""
import numpy as np
import pyarrow as pa
import pytest

import oracle_lib as O
from matrixone_b200 import capi, ops
from test_oracle_lz4 import corpus

pytestmark = pytest.mark.gpu


def _comp(raw):
    return pa.compress(raw, codec="lz4_raw", asbytes=True) if raw else b"\x00"


def test_corpus_round_trip(gpu):
    raws = corpus()
    got = ops.lz4_decode_blocks([_comp(r) for r in raws], [len(r) for r in raws])
    for g, r in zip(got, raws):
        assert g == r


def test_a_table_scan_worth_of_column_blocks(gpu):
    rng = np.random.default_rng(1)
    raws = []
    for i in range(600):       # 8192-row blocks of int64 / float64 / int32 / byte columns with different entropy
        kind = i % 4
        if kind == 0:
            raws.append((rng.integers(0, 1000, 8192) + i).astype(np.int64).tobytes())
        elif kind == 1:
            raws.append(rng.choice(np.round(rng.random(50) * 100, 2), 8192).astype(np.float64).tobytes())
        elif kind == 2:
            raws.append(np.sort(rng.integers(8000, 11000, 8192)).astype(np.int32).tobytes())
        else:
            raws.append(rng.choice([65, 78, 82], 8192).astype(np.uint8).tobytes())
    comps = [_comp(r) for r in raws]
    got = ops.lz4_decode_blocks(comps, [len(r) for r in raws])
    assert all(g == r for g, r in zip(got, raws))
    # the oracle agrees block by block
    for c, r in list(zip(comps, raws))[:20]:
        src = np.frombuffer(c, np.uint8); dst = np.zeros(len(r), np.uint8)
        assert O.go().og_lz4_decode_block(O.p(src), len(c), O.p(dst), len(r)) == len(r) and dst.tobytes() == r


def test_malformed_block_fails_the_call_and_names_the_block(gpu):
    raws = [b"abcdefgh" * 100, bytes(5000), b"hello world " * 50]
    comps = [_comp(r) for r in raws]
    comps[1] = comps[1][: len(comps[1]) // 2]
    with pytest.raises(capi.MoError) as e:
        ops.lz4_decode_blocks(comps, [len(r) for r in raws])
    assert "block 1" in str(e.value)
    with pytest.raises(capi.MoError):
        ops.lz4_decode_blocks([_comp(raws[0])], [len(raws[0]) + 1])        # decodes to a different size than the descriptor says
