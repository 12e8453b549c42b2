"""Sequential meshes (MeshSequentialDecoder.cs:8-118) without a GPU: the product's host connectivity
helper and the container indexer against the oracle and against the faces the test wrote."""
import numpy as np
import pytest

import drc_writer as W
import draco_sharp_b200 as D
from oracle import pyoracle as O


def _attrs(n_points, rng):
    corr = rng.integers(-20, 21, size=n_points * 3)
    return [dict(att_type=0, data_type=9, nc=3, seq_type=2, portable=W.portable_int(corr, 3, 0, 1, "raw", W.wrap_data(0, 4095)),
                 xform=W.quant_params([0.0, 1.0, 2.0], 8.0, 12))]


def _grid_faces(w, h):
    f = []
    for y in range(h - 1):
        for x in range(w - 1):
            a = y * w + x
            f += [(a, a + 1, a + w), (a + 1, a + w + 1, a + w)]
    return f


@pytest.mark.parametrize("w,h,method,scheme", [(4, 3, 0, "raw"), (20, 17, 0, "raw"), (20, 17, 0, "tagged"), (9, 9, 1, "raw"),
                                                (40, 30, 1, "raw"), (300, 300, 1, "raw"), (300, 300, 0, "tagged")])
def test_host_connectivity_of_sequential_meshes(w, h, method, scheme):
    rng = np.random.default_rng(w * 31 + h)
    faces = _grid_faces(w, h)
    n_points = w * h
    buf = np.frombuffer(W.sequential_mesh(faces, n_points, _attrs(n_points, rng), method, scheme), dtype=np.uint8)
    o = O.decode(buf)
    assert o.status == 0 and o.n_points == n_points and np.array_equal(o.faces, np.asarray(faces, dtype=np.uint32))
    bt = D.index_only([buf])
    bi = bt.buffer_info(0)
    assert bi.status == 0 and (bi.geometry_type, bi.encoder_method, bi.needs_connectivity) == (1, 0, 1)
    bt.host_connectivity(0)
    bt.finish()
    bi = bt.buffer_info(0)
    assert bi.status == 0 and bi.n_points == n_points and bi.attr_section_off == o.attr_section_off and bi.n_attrs == 1
    assert np.array_equal(bt.faces(0), o.faces)
    ai = bt.attr_info(0, 0)
    assert (ai.n_entries, ai.pred_method, ai.transform) == (n_points, 0, 1)
    bt.free()


def test_sequential_connectivity_errors_match_the_oracle():
    rng = np.random.default_rng(3)
    faces = _grid_faces(6, 5)
    good = bytearray(W.sequential_mesh(faces, 30, _attrs(30, rng), 0, "raw"))
    cases = []
    bad = bytearray(good); bad[13] = 7                       # connectivity method 7 (:81)
    cases.append(bytes(bad))
    for cut in (11, 12, 13, 14, 20, 40):
        cases.append(bytes(good[:cut]))
    # an index difference that drops below zero (:99): first symbol odd with magnitude 1
    neg = bytearray(b"DRACO" + bytes([2, 2, 1, 0, 0, 0]) + W.varint(1) + W.varint(3) + bytes([0]) + W.symbols_raw([3, 2, 2]))
    cases.append(bytes(neg) + W.point_cloud(3, _attrs(3, rng))[15:])
    for c in cases:
        b = np.frombuffer(c, dtype=np.uint8)
        o = O.decode(b)
        bt = D.index_only([b])
        if bt.buffer_info(0).status == 0:
            bt.host_connectivity(0)
            bt.finish()
        assert o.status < 0 and bt.buffer_info(0).status == o.status, (len(c), o.status, bt.buffer_info(0).status)
        bt.free()
