"""Constrained multi-parallelogram predictor in the oracle: round trip against an encoder written from the
bitstream specification in tests/drc_writer.py (the reference's own decoder is defective here, Appendix B-17, and holds
no fixture for the scheme), over the reference sample's real connectivity and over a grid surface."""
import os
import struct

import numpy as np
import pytest

from oracle import pyoracle as O
import drc_writer as W

ERR_PRED = -8  # ORC_ERR_PRED / DCB_ERR_PRED
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def mesh_buffer(attr_section):
    head = b"DRACO" + bytes([2, 2, 1, 1]) + struct.pack("<H", 0) + bytes([2]) + b"\xAA" * 37
    return np.frombuffer(head + attr_section, dtype=np.uint8), len(head)


def cmp_section(values, nc, maps, bits, rng, scheme="raw", p_crease=0.3, data_type=9, drop_flags=0):
    hi = (1 << bits) - 1
    corr, crease = W.cmp_encode(values, nc, maps, 0, hi, rng, p_crease)
    if drop_flags:
        k = max(range(4), key=lambda i: len(crease[i]))
        crease[k] = crease[k][:-drop_flags]
    sec = bytearray([1, 0xFF, 0, 0])
    sec += W.varint(1) + bytes([0, data_type, nc, 0]) + W.varint(0) + bytes([2 if data_type == 9 else 1])
    sec += W.portable_int(corr, nc, 4, 1, scheme, W.cmp_data(crease, 0, hi), num_bytes=4)
    if data_type == 9:
        sec += W.quant_params([0.5] * nc, 3.0, bits)
    return bytes(sec), crease


def house_maps():
    o = O.decode(np.fromfile(os.path.join(GOLD, "house_04.obj.drc"), dtype=np.uint8))
    assert o.status == 0
    return o


@pytest.mark.parametrize("nc,bits,p_crease,scheme", [(3, 12, 0.3, "raw"), (1, 8, 0.0, "tagged"), (2, 10, 1.0, "raw"),
                                                     (4, 30, 0.5, "uncompressed")])
def test_cmp_round_trip_on_the_sample_connectivity(nc, bits, p_crease, scheme):
    o = house_maps()
    rng = np.random.default_rng(nc * 100 + bits)
    m = o.maps[0]
    n = m["data_to_corner"].size
    values = rng.integers(0, 1 << bits, size=n * nc)
    sec, crease = cmp_section(values, nc, m, bits, rng, scheme=scheme, p_crease=p_crease)
    buf, aoff = mesh_buffer(sec)
    r = O.decode(buf, [m], aoff, o.n_points)
    assert r.status == 0
    assert r.attrs[0].pred_method == 4
    assert np.array_equal(r.attrs[0].qints, values.astype(np.int32))
    # the sample's connectivity offers entries with 1, 2, 3 and 4 usable parallelograms
    assert all(len(c) > 0 for c in crease)


def test_cmp_smooth_values_and_seams():
    """Smooth data (small corrections) on the second attribute decoder's maps (attribute seams cut `opposite`)."""
    o = house_maps()
    rng = np.random.default_rng(3)
    m = o.maps[1]
    n = m["data_to_corner"].size
    values = (np.cumsum(rng.integers(-3, 4, size=(n, 2)), axis=0) + 500).clip(0, 1023).ravel()
    sec, _ = cmp_section(values, 2, m, 10, rng, scheme="tagged")
    sec = bytearray(sec)
    buf, aoff = mesh_buffer(bytes(sec))
    r = O.decode(buf, [m], aoff, o.n_points)
    assert r.status == 0 and np.array_equal(r.attrs[0].qints, values.astype(np.int32))


def test_cmp_too_few_flags_fails_the_buffer():
    o = house_maps()
    rng = np.random.default_rng(5)
    m = o.maps[0]
    values = rng.integers(0, 4096, size=m["data_to_corner"].size * 3)
    sec, _ = cmp_section(values, 3, m, 12, rng, drop_flags=1)
    buf, aoff = mesh_buffer(sec)
    assert O.decode(buf, [m], aoff, o.n_points).status == ERR_PRED


def test_cmp_grid_surface():
    from draco_sharp_b200 import synth_gen as G
    w, h = 23, 19
    topo = G.grid_topology(w, h)
    rng = np.random.default_rng(9)
    values = rng.integers(0, 1 << 14, size=w * h * 3)
    sec, crease = cmp_section(values, 3, topo, 14, rng)
    buf, aoff = mesh_buffer(sec)
    r = O.decode(buf, [topo], aoff, w * h)
    assert r.status == 0 and np.array_equal(r.attrs[0].qints, values.astype(np.int32))
