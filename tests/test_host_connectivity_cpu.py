"""The product's host Edgebreaker helper (dcb_host_connectivity) against the pinned oracle: same
faces, same corner tables, same traversal maps on the reference's sample asset; malformed connectivity fails with a
status, never a crash.  No GPU needed: connectivity is host work and indexing does not touch the device."""
import hashlib
import os

import numpy as np
import pytest

import draco_sharp_b200 as D
from oracle import pyoracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _house():
    return np.fromfile(os.path.join(GOLD, "house_04.obj.drc"), dtype=np.uint8)


def test_house_connectivity_matches_oracle_and_goldens():
    b = _house()
    o = O.decode(b)
    bt = D.index_only([b])
    bt.host_connectivity(0)
    bt.finish()
    bi = bt.buffer_info(0)
    assert bi.n_points == o.n_points == 3220 and bi.attr_section_off == 1158 and bi.n_attrs == 3
    f = bt.faces(0)
    assert f.shape == (2588, 3) and np.array_equal(f, o.faces)
    names = ("opposite", "corner_to_vertex", "data_to_corner", "vertex_to_data")
    for d, m in enumerate(o.maps):
        for w, nm in enumerate(names):
            got = bt.mesh_map(0, d, w)
            assert np.array_equal(got.view(np.uint32), np.asarray(m[nm]).view(np.uint32)), (d, nm)
    # SHA-256 of the position decoder's data_to_corner map (uint32 LE)
    assert hashlib.sha256(bt.mesh_map(0, 0, 2).tobytes()).hexdigest().startswith("742b5197")
    assert [bt.attr_info(0, a).n_entries for a in range(3)] == [1775, 3220, 1775]
    # the sample's tex-coord attribute uses the TexCoordsPortable predictor: the indexer locates its
    # orientation flags and hands the stream to the tex-coord kernels
    assert bi.status == 0 and o.status == 0
    assert (bt.attr_info(0, 1).pred_method, bt.attr_info(0, 1).transform) == (5, 1)
    bt.free()


def test_truncated_connectivity_fails_with_status():
    b = _house()
    for cut in (12, 13, 20, 40, 100, 400, 900, 1157):
        bt = D.index_only([b[:cut].copy()])
        bt.host_connectivity(0)
        assert bt.buffer_info(0).status < 0, cut
        assert bt.faces(0).shape[0] == 0 or bt.buffer_info(0).status < 0
        bt.free()


def test_corrupted_connectivity_never_crashes():
    b = _house()
    rng = np.random.default_rng(11)
    n_ok = 0
    for it in range(300):
        c = b[:1400].copy()
        for _ in range(int(rng.integers(1, 6))):
            c[int(rng.integers(11, 1158))] = int(rng.integers(0, 256))
        bt = D.index_only([c])
        bt.host_connectivity(0)
        bt.finish()
        st = bt.buffer_info(0).status
        assert -104 <= st <= 0
        n_ok += st == 0
        bt.free()


def test_helper_refuses_point_clouds_and_bad_indices():
    from draco_sharp_b200 import synth_gen as G
    from draco_sharp_b200 import _native as N
    buf, _ = G.synth_cloud(G.make_spec(10, seed=1))
    bt = D.index_only([buf])
    with pytest.raises(D.DracoError):
        bt.host_connectivity(0)
    assert N.lib().dcb_host_connectivity(bt.h, 5) == -100
    bt.free()


def test_positions_only_fixture_is_a_complete_mesh():
    """tests/golden/house_04_positions_only.drc (bench.py --workload c1): the product's host helper walks its
    connectivity, the indexer accepts every attribute, and the oracle decodes it to the golden floats."""
    b = np.fromfile(os.path.join(GOLD, "house_04_positions_only.drc"), dtype=np.uint8)
    assert bytes(b[:1158]) == bytes(_house()[:1158])          # header + connectivity verbatim
    bt = D.index_only([b])
    bt.host_connectivity(0)
    bt.finish()
    bi = bt.buffer_info(0)
    assert bi.status == 0 and bi.n_attrs == 1 and bi.n_points == 3220
    ai = bt.attr_info(0, 0)
    assert ai.n_entries == 1775 and ai.pred_method == 1 and ai.out_bytes == 1775 * 12
    assert bt.faces(0).shape == (2588, 3)
    bt.free()
    r = O.decode(b)
    assert r.status == 0
    assert hashlib.sha256(r.attrs[0].out.tobytes()).hexdigest() == "028840c055ebfbc5b9a3a04b28d2fc5d0f9cae9c12821f030a815a0826bdcb37"
