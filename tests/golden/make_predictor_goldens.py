"""Writes tests/golden/mesh_predictors.npz: two small Edgebreaker-mesh buffers over the connectivity of the reference's
sample (house_04.obj.drc) that use the SURVEY 8f-3 predictors the sample itself does not --

  cmp.drc-like buffer   positions by ConstrainedMultiParallelogram + wrap, a one-component generic attribute likewise
  geo.drc-like buffer   positions by Parallelogram + wrap, normals by GeometricNormal + canonicalized octahedron transform

-- together with what they must decode to: the values the bitstream-specification encoder in tests/drc_writer.py
encoded (CMP), its predicted octahedral coordinates (geometric normal, all-zero corrections), and SHA-256 of the output
bytes the CPU oracle produced when the fixture was made.  Run from the repository root:

    python tests/golden/make_predictor_goldens.py

The connectivity maps are not stored: both the oracle and the product's host helper rebuild them from house_04.obj.drc.
"""
import hashlib
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import drc_writer as W  # noqa: E402
from oracle import pyoracle as O  # noqa: E402


def main():
    o = O.decode(np.fromfile(os.path.join(HERE, "house_04.obj.drc"), dtype=np.uint8))
    assert o.status == 0
    m0, m1 = o.maps[0], o.maps[1]
    n0, n1 = m0["data_to_corner"].size, m1["data_to_corner"].size
    rng = np.random.default_rng(20261019)
    head = b"DRACO" + bytes([2, 2, 1, 1]) + struct.pack("<H", 0) + bytes([2]) + b"\xAA" * 37

    # ---- constrained multi-parallelogram ----
    v_pos = (np.cumsum(rng.integers(-12, 13, size=(n0, 3)), axis=0) + 2048).clip(0, 4095).ravel()
    v_gen = (np.cumsum(rng.integers(-2, 3, size=n0)) + 128).clip(0, 255)
    c_pos, f_pos = W.cmp_encode(v_pos, 3, m0, 0, 4095, rng, 0.25)
    c_gen, f_gen = W.cmp_encode(v_gen, 1, m0, 0, 255, rng, 0.25)
    sec = bytearray([1, 0xFF, 0, 0]) + W.varint(2) + bytes([0, 9, 3, 0]) + W.varint(0) + bytes([4, 2, 1, 0]) + W.varint(1) + bytes([2, 1])
    sec += W.portable_int(c_pos, 3, 4, 1, "raw", W.cmp_data(f_pos, 0, 4095))
    sec += W.portable_int(c_gen, 1, 4, 1, "tagged", W.cmp_data(f_gen, 0, 255))
    sec += W.quant_params([1.0, 2.0, 3.0], 10.0, 12)
    cmp_buf = np.frombuffer(head + bytes(sec), dtype=np.uint8)
    r = O.decode(cmp_buf, [m0, m1], len(head), o.n_points)
    assert r.status == 0
    assert np.array_equal(r.attrs[0].qints, v_pos.astype(np.int32)) and np.array_equal(r.attrs[1].qints, v_gen.astype(np.int32))
    cmp_sha = [hashlib.sha256(a.out.tobytes()).hexdigest() for a in r.attrs]

    # ---- geometric normal (normals in the second attributes decoder: its maps carry the attribute seams) ----
    c_p = rng.integers(-30, 31, size=n0 * 3)
    flips = rng.integers(0, 2, size=n1)
    sec = bytearray([2, 0xFF, 0, 0, 0, 1, 0])
    sec += W.varint(1) + bytes([0, 9, 3, 0]) + W.varint(0) + bytes([2])
    sec += W.varint(1) + bytes([1, 9, 3, 0]) + W.varint(1) + bytes([3])
    sec += W.portable_int(c_p, 3, 1, 1, "raw", W.wrap_data(0, 4095)) + W.quant_params([1.0, 2.0, 3.0], 10.0, 12)
    sec += W.portable_int(np.zeros(n1 * 2, dtype=np.int64), 2, 6, 3, "raw", W.geometric_normal_data(flips, 10), zig=False) + bytes([10])
    geo_buf = np.frombuffer(head + bytes(sec), dtype=np.uint8)
    r = O.decode(geo_buf, [m0, m1], len(head), o.n_points)
    assert r.status == 0
    pred = np.asarray(W.geometric_normal_predictions(m1, m0, r.attrs[0].qints, 10, flips), dtype=np.int32)
    assert np.array_equal(r.attrs[1].qints, pred)
    geo_sha = [hashlib.sha256(a.out.tobytes()).hexdigest() for a in r.attrs]

    np.savez_compressed(os.path.join(HERE, "mesh_predictors.npz"), attr_off=len(head), n_points=o.n_points,
                        cmp_buf=cmp_buf, cmp_pos=v_pos.astype(np.int32), cmp_gen=v_gen.astype(np.int32), cmp_sha=np.array(cmp_sha),
                        geo_buf=geo_buf, geo_pred=pred, geo_sha=np.array(geo_sha))
    print("wrote mesh_predictors.npz:", cmp_buf.size, "+", geo_buf.size, "bytes of bitstream;", cmp_sha, geo_sha)


if __name__ == "__main__":
    main()
