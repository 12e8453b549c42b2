"""Oracle vs the reference's own known-answer tests (tests/Draco.UnitTests) and the sample asset's self-checks.

These pin the CPU oracle (oracle/): it is the checker every GPU parity test relies on.
"""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

import drc_writer as W
from oracle import pyoracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- tests/Draco.UnitTests/IO/EncoderBufferTests.cs:27-46: varint 98 and 1739 round-trip ----
@pytest.mark.parametrize("value", [98, 1739, 0, 127, 128, 16383, 16384, (1 << 32) - 1, (1 << 63) + 5])
def test_varint_roundtrip(value):
    enc = np.frombuffer(W.varint(value), dtype=np.uint8).copy()
    pos = C.c_uint64(0)
    out = C.c_uint64(0)
    assert O.lib().orc_varint(enc.ctypes.data, enc.size, C.byref(pos), C.byref(out)) == 0
    assert out.value == value and pos.value == enc.size
    if value == 1739:
        assert bytes(enc) == bytes([0xCB, 0x0D])  # LEB128 of 1739


def test_varint_truncated_is_eof():
    enc = np.array([0x80, 0x80], dtype=np.uint8)
    pos = C.c_uint64(0)
    out = C.c_uint64(0)
    assert O.lib().orc_varint(enc.ctypes.data, enc.size, C.byref(pos), C.byref(out)) == -1


# ---- EncoderBufferTests.cs:7-25: the 9-bit value 0b001100010 written LSB-first reads back the same ----
def test_lsb_first_9_bits():
    value, count = 0b001100010, 9
    # EncoderBuffer.EncodeLeastSignificantBits32 packs bit i of the value at stream bit i
    packed = np.array([value & 0xFF, value >> 8], dtype=np.uint8)
    bitpos = C.c_uint64(0)
    err = C.c_int(0)
    got = O.lib().orc_read_bits_lsb(packed.ctypes.data, packed.size, C.byref(bitpos), count, C.byref(err))
    assert got == value and bitpos.value == 9 and err.value == 0
    # full 32-bit assembly (SURVEY B-4: the C# truncates bits >= 8 through a (byte) cast)
    v32 = 0xDEADBEEF
    packed = np.frombuffer(v32.to_bytes(4, "little"), dtype=np.uint8).copy()
    bitpos = C.c_uint64(0)
    assert O.lib().orc_read_bits_lsb(packed.ctypes.data, 4, C.byref(bitpos), 32, C.byref(err)) == v32


# ---- tests/Draco.UnitTests/IO/ConstantsTests.cs:7-21: int -3 <-> uint 4294967293 ----
def test_reinterpret_cast():
    assert O.lib().orc_reinterpret_i2u(-3) == 4294967293
    assert O.lib().orc_reinterpret_i2u(7) == 7


# ---- tests/Draco.UnitTests/IO/Core/MathUtilitiesTests.cs:7-20 ----
@pytest.mark.parametrize("n,root", [(0, 0), (4, 2), (48722615824, 220732)])
def test_int_sqrt(n, root):
    assert O.lib().orc_int_sqrt(n) == root


def test_zigzag_pairs():
    # BitUtilities.cs:72-81
    for sym, val in [(0, 0), (1, -1), (2, 1), (3, -2), (4, 2), (4294967295, -2147483648), (4294967294, 2147483647)]:
        assert O.lib().orc_zigzag(sym) == val
        if -(1 << 31) < val:
            assert W.zigzag(val) == sym


def test_rans_precision_rule():
    # RAnsSymbolCoding.cs:10-26: clamp(3 * bits / 2, 12, 20)
    expect = {1: 12, 8: 12, 9: 13, 10: 15, 11: 16, 12: 18, 13: 19, 14: 20, 18: 20}
    for mbl, p in expect.items():
        assert O.lib().orc_rans_precision(mbl) == p


# ---- the sample asset: SURVEY.md Appendix C ----
@pytest.fixture(scope="module")
def house():
    b = np.fromfile(os.path.join(GOLD, "house_04.obj.drc"), dtype=np.uint8)
    return b, O.decode(b)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_house_container_and_connectivity(house):
    b, r = house
    assert (r.ver_major, r.ver_minor, r.geom_type, r.method) == (2, 2, 1, 1)
    assert r.n_faces == 2588          # `f` lines of house_04.obj
    assert r.attr_section_off == 1158
    assert r.n_decoders == 3 and r.n_attrs == 3
    # the oracle decodes the whole sample, TexCoordsPortable (attribute 1) included
    assert r.status == 0 and r.end_off == len(b)
    m = r.maps[0]
    assert m["data_to_corner"].size == 1775 and m["opposite"].size == 3 * 2588
    assert m["data_to_corner"][:8].tolist() == [1, 2, 0, 4, 8, 10, 11, 9]
    assert sha(m["data_to_corner"].astype("<u4")) == "742b519742d6189a457257cdb86744cf7aa2c560fcb9971d754c0a77c3d40465"
    assert r.maps[1]["data_to_corner"].size == 3220  # tex-coord entries (6,440 symbols / 2)
    # all faces reference valid points, every point is used
    assert r.faces.max() == r.n_points - 1 and np.unique(r.faces).size == r.n_points


def test_house_position_stream_self_checks(house):
    b, r = house
    a = r.attrs[0]
    assert (a.att_type, a.data_type, a.nc, a.seq_type) == (0, 9, 3, 2)
    assert (a.pred_method, a.transform, a.scheme, a.max_bit_length, a.precision) == (1, 1, 1, 9, 13)
    assert a.table_symbols == 2047 and a.payload_len == 3077 and a.n_entries == 1775
    # rANS self-check: every payload byte consumed, decoder back at its initial state L = 4 * 2^13
    assert a.leftover == 0 and a.final_state == 32768
    assert a.symbols[:12].tolist() == [1967, 1334, 58, 1587, 0, 0, 796, 556, 0, 796, 494, 0] and a.symbols.max() == 2046
    assert (a.xf_a, a.xf_b) == (0, 2047)
    assert a.qbits == 11 and abs(a.qrange - 2009.9021) < 1e-3
    assert np.allclose(a.qmin[:3], [-538.20062, 0.0, -1003.70178], atol=1e-4)


def test_house_position_goldens(house):
    b, r = house
    a = r.attrs[0]
    assert sha(a.symbols.astype("<u4")) == "00823b9eb65a088cb18987b7016f3756f94fbccb5911d1e86912911af2fcb07b"
    assert sha(a.corr.astype("<i4")) == "1e0530c23c261db13732944a656bc2349948b412cab7e7300a9bc4dbfe224b45"
    assert sha(a.qints.astype("<i4")) == "15d5eeb7c1c24707f0bcc20b6ed5e89b6faaa5d46da4fd5f15b5872352e02be6"
    assert sha(a.out) == "028840c055ebfbc5b9a3a04b28d2fc5d0f9cae9c12821f030a815a0826bdcb37"
    q = a.qints.reshape(-1, 3)
    assert q.min(axis=0).tolist() == [0, 0, 0] and q.max(axis=0).tolist() == [1154, 1009, 2047]
    f = a.out.view(np.float32).reshape(-1, 3)
    assert f[0].tolist() == [506.51641845703125, 654.9119262695312, -975.2273559570312]


def test_house_positions_match_the_source_obj(house):
    """Ground truth that does not come from us: every dequantised position lies within half a quantisation
    step of a `v` line of house_04.obj, the mesh the asset was encoded from."""
    from scipy.spatial import cKDTree
    b, r = house
    vs = np.load(os.path.join(GOLD, "house_04_obj_vertices.npy"))
    pos = r.attrs[0].out.view(np.float32).reshape(-1, 3).astype(np.float64)
    d, _ = cKDTree(vs).query(pos, p=np.inf)
    half_step = 2009.9021 / 2047 / 2
    assert d.max() < half_step
    assert d.max() == pytest.approx(0.48928, abs=1e-4)


def test_house_texcoords_match_the_obj(house):
    """SURVEY 8f-3 groundwork: the oracle's TexCoordsPortable predictor (MeshPredictionSchemeTexCoordsPortable
    Decoder / Predictor) on the reference's sample.  Ground truth that does not come from us: every dequantised
    (u, v) lies within half a quantisation step of a `vt` line of house_04.obj, and the rABS-coded orientation
    flags are consumed exactly (a wrong flag order or a wrong parent position would miss by whole texels)."""
    b, r = house
    a = r.attrs[1]
    assert (a.att_type, a.data_type, a.nc, a.seq_type) == (3, 9, 2, 2)
    assert (a.pred_method, a.transform, a.n_entries) == (5, 1, 3220)
    vts = np.load(os.path.join(GOLD, "house_04_obj_texcoords.npy"))
    uv = a.out.view(np.float32).reshape(-1, 2).astype(np.float64)
    half_step = float(a.qrange) / ((1 << a.qbits) - 1) / 2
    d = np.abs(uv[:, None, :] - vts[None, :, :]).max(axis=2).min(axis=1)
    assert d.max() < half_step, (d.max(), half_step)
    # the third attribute (generic, parallelogram) decodes too: the whole file is consumed
    assert r.attrs[2].out.nbytes == 1775 and r.status == 0
    # goldens for the GPU kernel to come (quantized ints and floats of the tex-coord attribute, little endian)
    assert sha(a.qints.astype("<i4")) == "a243b8cf61c7145d3f75eda19721098918e735e925822b6a211ae6686a6d2990"
    assert sha(a.out) == "26bd6cc9ae35d081caba5b2061dd9d6d049a66c9050f7c5b9cfdc345c3c8449a"
