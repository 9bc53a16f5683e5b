"""Pins the LZ4 block decoder restatement (og_lz4_decode_block) to liblz4 (pyarrow's lz4_raw codec = the raw block format that
compress.Compress / lz4.CompressBlock produce, pkg/compress/compress.go:26-47) and to the reference's own TestLz4 vector
(pkg/compress/compress_test.go:26-43).  No GPU."""
import numpy as np
import pyarrow as pa
import pytest

import oracle_lib as O


def corpus():
    rng = np.random.default_rng(0)
    out = [np.array([200, 200, 0, 200, 10, 30, 20, 1111], dtype=np.int64).tobytes(),          # compress_test.go:29
           b"", b"a", b"abcd" * 5000, bytes(70000), rng.integers(0, 256, 50000, dtype=np.uint8).tobytes(),
           np.arange(8192, dtype=np.int64).tobytes(), (np.arange(8192, dtype=np.int64) // 37).tobytes(),
           rng.integers(0, 4, 65536, dtype=np.uint8).tobytes(), (b"x" * 300 + b"yz" * 400 + bytes(range(256))) * 20,
           rng.choice([1.25, 2.5, 100.0], 8192).astype(np.float64).tobytes()]
    return out


def test_decoder_matches_liblz4_on_the_corpus():
    for raw in corpus():
        comp = pa.compress(raw, codec="lz4_raw", asbytes=True) if raw else b"\x00"
        src = np.frombuffer(comp, np.uint8); dst = np.zeros(max(len(raw), 1), np.uint8)
        n = O.go().og_lz4_decode_block(O.p(src), len(comp), O.p(dst), len(raw))
        assert n == len(raw) and dst[:n].tobytes() == raw


def test_malformed_blocks_are_rejected():
    raw = b"abcdefgh" * 100
    comp = bytearray(pa.compress(raw, codec="lz4_raw", asbytes=True))
    dst = np.zeros(len(raw), np.uint8)
    for mutate in (lambda c: c[:len(c) // 2], lambda c: c + b"\xff\xff", lambda c: bytes([c[0]]) + b"\x00\x00" + c[3:]):
        bad = np.frombuffer(bytes(mutate(bytes(comp))), np.uint8)
        n = O.go().og_lz4_decode_block(O.p(bad), len(bad), O.p(dst), len(raw))
        assert n != len(raw) or dst.tobytes() != raw
    short = np.zeros(10, np.uint8)
    good = np.frombuffer(bytes(comp), np.uint8)
    assert O.go().og_lz4_decode_block(O.p(good), len(good), O.p(short), 10) == -1       # does not fit
