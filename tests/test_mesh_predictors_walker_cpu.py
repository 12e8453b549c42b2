"""Host logic without a GPU: the container walker of libdracob200.so on Edgebreaker-mesh buffers that use the 8f-3
predictors (constrained multi-parallelogram, geometric normal, tex-coords-portable) -- valid, truncated and bit-flipped --
against the oracle.  The walker locates and validates the prediction data (crease-flag blocks, flip bits) without
decoding a symbol; what only a decode can find (too few flags, bad maps) is the device's business."""
import os

import numpy as np
import pytest

import draco_sharp_b200 as D
import drc_writer as W
from oracle import pyoracle as O
from test_cmp_cpu import mesh_buffer, cmp_section, house_maps
from test_geometric_normal_cpu import normals_section

DEVICE_SIDE = (-8, -10)   # DCB_ERR_PRED (flags run out), DCB_ERR_MAPS: found while decoding, not while indexing


def walker_status(buf, attr_off, n_points, maps):
    bt = D.index_only([buf])
    try:
        bt.set_attr_section(0, attr_off, n_points)
        for d, m in enumerate(maps):
            bt.set_mesh_maps(0, d, m["opposite"], m["corner_to_vertex"], m["data_to_corner"], m["vertex_to_data"])
        bt.finish()
        st = bt.status(0)
        infos = [bt.attr_info(0, a) for a in range(bt.buffer_info(0).n_attrs)] if st == 0 else []
        return st, infos
    finally:
        bt.free()


def variants(buf, attr_off, rng, n_cuts=60, n_flips=60):
    out = [buf]
    for cut in sorted(set(int(c) for c in rng.integers(attr_off, len(buf), n_cuts)) | {len(buf) - 1, len(buf) - 9}):
        out.append(buf[:cut].copy())
    for _ in range(n_flips):
        b = buf.copy()
        i = int(rng.integers(attr_off, len(buf)))
        b[i] ^= 1 << int(rng.integers(0, 8))
        out.append(b)
    return out


@pytest.mark.parametrize("kind", ["cmp", "geo"])
def test_walker_status_parity_on_mesh_predictor_buffers(kind):
    o = house_maps()
    rng = np.random.default_rng(99)
    n0, n1 = o.maps[0]["data_to_corner"].size, o.maps[1]["data_to_corner"].size
    maps = [o.maps[0], o.maps[1]]
    if kind == "cmp":
        sec, _ = cmp_section(rng.integers(0, 4096, size=n0 * 3), 3, o.maps[0], 12, rng)
    else:
        sec = normals_section(rng.integers(-40, 41, size=n0 * 3), rng.integers(0, 1024, size=n1 * 2), rng.integers(0, 2, size=n1), 10)
    buf, aoff = mesh_buffer(sec)
    n_ok = n_bad = 0
    for v in variants(np.array(buf), aoff, rng):
        ref = O.decode(v, maps, aoff, o.n_points)
        st, infos = walker_status(v, aoff, o.n_points, maps)
        if ref.status in DEVICE_SIDE:
            assert st in (0, ref.status), (len(v), st, ref.status)
        else:
            assert st == ref.status, (len(v), st, ref.status)
        if ref.status == 0:
            n_ok += 1
            for ai, ra in zip(infos, ref.attrs):
                assert (ai.pred_method, ai.transform, ai.n_entries, ai.out_bytes) == (ra.pred_method, ra.transform, ra.n_entries, ra.out_bytes)
        else:
            n_bad += 1
    assert n_ok >= 1 and n_bad >= 30
