"""Geometric-normal predictor in the oracle, against a prediction written from the bitstream specification
in tests/drc_writer.py (the reference's own decoder is defective here, Appendix B-17, and holds no fixture for it): with
all-zero corrections the decoded octahedral coordinates ARE the predictions."""
import os
import struct

import numpy as np
import pytest

import drc_writer as W
from oracle import pyoracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def mesh_buffer(attr_section):
    head = b"DRACO" + bytes([2, 2, 1, 1]) + struct.pack("<H", 0) + bytes([2]) + b"\xAA" * 37
    return np.frombuffer(head + attr_section, dtype=np.uint8), len(head)


def normals_section(c_pos, corr_n, flips, nbits, pos_bits=12, normals_decoder=1, scheme="raw", canonical=True, pos_scheme="raw"):
    """Two attribute decoders: positions (parallelogram + wrap) in decoder 0, normals (geometric normal + octahedron
    transform) in decoder `normals_decoder` (0: same decoder and maps as the positions, 1: the second decoder's maps)."""
    hi = (1 << pos_bits) - 1
    pos = W.portable_int(c_pos, 3, 1, 1, pos_scheme, W.wrap_data(0, hi), num_bytes=4)
    nrm = W.portable_int(corr_n, 2, 6, 3 if canonical else 2, scheme, W.geometric_normal_data(flips, nbits, canonical), zig=False,
                         num_bytes=4)
    if normals_decoder == 0:
        sec = bytearray([1, 0xFF, 0, 0])
        sec += W.varint(2) + bytes([0, 9, 3, 0]) + W.varint(0) + bytes([1, 9, 3, 0]) + W.varint(1) + bytes([2, 3])
        sec += pos + nrm + W.quant_params([1.0, 2.0, 3.0], 10.0, pos_bits) + bytes([nbits])
    else:
        sec = bytearray([2, 0xFF, 0, 0, 0, 1, 0])
        sec += W.varint(1) + bytes([0, 9, 3, 0]) + W.varint(0) + bytes([2])
        sec += W.varint(1) + bytes([1, 9, 3, 0]) + W.varint(1) + bytes([3])
        sec += pos + W.quant_params([1.0, 2.0, 3.0], 10.0, pos_bits) + nrm + bytes([nbits])
    return bytes(sec)


def house():
    o = O.decode(np.fromfile(os.path.join(GOLD, "house_04.obj.drc"), dtype=np.uint8))
    assert o.status == 0
    return o


@pytest.mark.parametrize("dec,nbits,pos_bits,canonical", [(0, 10, 12, True), (1, 8, 12, True), (1, 30, 30, True), (1, 2, 12, True),
                                                         (0, 12, 14, False)])
def test_zero_corrections_give_the_predictions(dec, nbits, pos_bits, canonical):
    o = house()
    rng = np.random.default_rng(dec * 7 + nbits)
    n0 = o.maps[0]["data_to_corner"].size
    m = o.maps[dec]
    n = m["data_to_corner"].size
    big = pos_bits >= 30
    c_pos = rng.integers(-(1 << 28), 1 << 28, size=n0 * 3) if big else rng.integers(-40, 41, size=n0 * 3)
    flips = rng.integers(0, 2, size=n)
    sec = normals_section(c_pos, np.zeros(n * 2, dtype=np.int64), flips, nbits, pos_bits, dec, "uncompressed" if big else "raw",
                          canonical, "uncompressed" if big else "raw")
    buf, aoff = mesh_buffer(sec)
    r = O.decode(buf, [o.maps[0], o.maps[1]], aoff, o.n_points)
    assert r.status == 0 and r.attrs[1].pred_method == 6
    want = W.geometric_normal_predictions(m, o.maps[0], r.attrs[0].qints, nbits, flips)
    assert np.array_equal(r.attrs[1].qints, np.asarray(want, dtype=np.int32))
    assert r.attrs[1].out.view(np.float32).size == 3 * n


def test_random_corrections_decode_and_short_flip_block_fails():
    o = house()
    rng = np.random.default_rng(3)
    n0, n1 = o.maps[0]["data_to_corner"].size, o.maps[1]["data_to_corner"].size
    c_pos = rng.integers(-40, 41, size=n0 * 3)
    corr = rng.integers(0, 1 << 10, size=n1 * 2)
    flips = rng.integers(0, 2, size=n1)
    buf, aoff = mesh_buffer(normals_section(c_pos, corr, flips, 10, scheme="tagged"))
    r = O.decode(buf, [o.maps[0], o.maps[1]], aoff, o.n_points)
    assert r.status == 0
    q = r.attrs[1].qints
    assert q.min() >= 0 and q.max() <= (1 << 10) - 1
    nrm = r.attrs[1].out.view(np.float32).reshape(-1, 3)
    assert np.allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-5)
    # no maps for the normals' decoder: the buffer fails, nothing else
    assert O.decode(buf, [o.maps[0]], aoff, o.n_points).status != 0
