"""Host logic without a GPU: the container walker of libdracob200.so (dcb_index with a NULL ctx) against the oracle
on valid, truncated and bit-flipped buffers; the same walker source runs on the device behind Tagged bit areas."""
import numpy as np
import pytest

import draco_sharp_b200 as D
from draco_sharp_b200 import synth_gen as G
from oracle import pyoracle as O

from common import cloud
from test_gpu_crafted import crafted_buffers


def _check(bufs, expect_resolved=True):
    bt = D.index_only(bufs)
    n_ok = 0
    for k, b in enumerate(bufs):
        o = O.decode(b)
        bi = bt.buffer_info(k)
        blocked = bi.status == 0 and any(not bt.attr_info(k, a).resolved for a in range(bi.n_attrs))
        if expect_resolved or not blocked:
            # everything the oracle rejects is rejected by the walker with the same code (unless the walk is
            # still parked at a Tagged bit area: the device continues it)
            assert bi.status == o.status, (k, bi.status, o.status)
        if o.status != 0:
            continue
        n_ok += 1
        assert (bi.geometry_type, bi.encoder_method, bi.n_points, bi.n_attrs) == (o.geom_type, o.method, o.n_points, o.n_attrs)
        for a, oa in enumerate(o.attrs):
            ai = bt.attr_info(k, a)
            assert (ai.att_type, ai.data_type, ai.num_components, ai.unique_id, ai.seq_decoder_type, ai.n_entries,
                    ai.out_bytes) == (oa.att_type, oa.data_type, oa.nc, oa.unique_id, oa.seq_type, oa.n_entries, oa.out_bytes)
            assert ai.out_off % 128 == 0
            if not ai.resolved:
                assert not expect_resolved
                continue
            if oa.seq_type != 0:
                assert (ai.pred_method, ai.transform) == (oa.pred_method, oa.transform)
                if oa.n_entries and oa.compressed:
                    assert (ai.scheme, ai.precision_bits) == (oa.scheme, oa.precision)
                if oa.transform in (1, 2, 3):
                    assert ai.xf_a == oa.xf_a
                if oa.transform in (1, 3):
                    assert ai.xf_b == oa.xf_b
            if oa.seq_type == 2:
                assert list(ai.q_min)[:oa.nc] == [np.float32(x) for x in oa.qmin[:oa.nc]]
                assert ai.q_range == np.float32(oa.qrange) and ai.q_bits == oa.qbits
    bt.free()
    return n_ok


def test_walker_matches_oracle_on_raw_buffers():
    bufs = [cloud(n, seed=3 + n, scheme=1, normal_bits=10, colors=1) for n in (0, 1, 5, 1000, 20000)]
    assert _check(bufs) == len(bufs)


def test_walker_status_parity_on_malformed_raw_buffers():
    rng = np.random.default_rng(17)
    g = cloud(2000, seed=4, scheme=1, normal_bits=10, colors=1)
    bufs = [g[:cut].copy() for cut in list(range(0, 64)) + [100, 200, 500, len(g) // 2, len(g) - 20, len(g) - 1]]
    for _ in range(400):
        b = g.copy()
        pos = int(rng.integers(0, min(len(b), 1600)))
        b[pos] ^= np.uint8(1 << int(rng.integers(0, 8)))
        bufs.append(b)
    n_ok = _check(bufs)
    assert n_ok > 50


def test_walker_stops_at_tagged_bit_areas():
    bufs = [cloud(3000, seed=8, scheme=0, normal_bits=10, colors=1)] + crafted_buffers()
    _check(bufs, expect_resolved=False)
    bt = D.index_only(bufs[:1])
    assert bt.buffer_info(0).status == 0
    assert not bt.attr_info(0, 0).resolved  # positions are Tagged: bit area length unknown on the host
    assert bt.points == 3000 and bt.in_bytes == len(bufs[0])


def test_batch_accounting_and_layout():
    sp = G.make_spec(5000, seed=77, scheme=1, colors=1)
    arena, offs, lens, sums, schemes, used = G.synth_batch(sp, 16)
    bufs = [arena[int(o): int(o + l)] for o, l in zip(offs, lens)]
    bt = D.index_only(bufs)
    assert bt.points == 16 * 5000
    assert bt.in_bytes == int(lens.sum())
    out_bytes = 16 * (5000 * 12 + 5000 * 3)
    assert bt.algo_bytes == int(lens.sum()) + out_bytes
    spans = sorted((bt.attr_info(k, a).out_off, bt.attr_info(k, a).out_bytes) for k in range(16) for a in range(2))
    for (o0, n0), (o1, _) in zip(spans, spans[1:]):
        assert o0 + n0 <= o1  # outputs never overlap
    assert bt.out_bytes >= spans[-1][0] + spans[-1][1]
    bt.free()


def test_mesh_buffers_need_connectivity_from_the_host():
    import os
    b = np.fromfile(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "house_04.obj.drc"), dtype=np.uint8)
    bt = D.index_only([b])
    bi = bt.buffer_info(0)
    assert bi.status == 0 and bi.geometry_type == 1 and bi.encoder_method == 1 and bi.needs_connectivity == 1
    o = O.decode(b)
    bt.set_attr_section(0, o.attr_section_off, o.n_points)
    for d, m in enumerate(o.maps):
        bt.set_mesh_maps(0, d, m["opposite"], m["corner_to_vertex"], m["data_to_corner"], m["vertex_to_data"])
    bt.finish()
    bi = bt.buffer_info(0)
    # attribute 1 uses the TexCoordsPortable predictor: indexed like the parallelogram streams
    assert bi.status == 0 and o.status == 0
    at = bt.attr_info(0, 1)
    assert (at.pred_method, at.transform, at.n_entries) == (5, 1, 3220)
    assert (at.xf_a, at.xf_b) == (o.attrs[1].xf_a, o.attrs[1].xf_b)
    ai = bt.attr_info(0, 0)
    assert (ai.pred_method, ai.transform, ai.scheme, ai.precision_bits, ai.n_entries) == (1, 1, 1, 13, 1775)
    assert (ai.xf_a, ai.xf_b) == (0, 2047)
    bt.free()
