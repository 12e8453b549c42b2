"""Committed golden vectors for the SURVEY 8f-3 predictors the reference's sample does not use
(tests/golden/mesh_predictors.npz, written by tests/golden/make_predictor_goldens.py): the oracle must decode the stored
bitstreams to the stored values -- the values an independent bitstream-specification encoder encoded (constrained
multi-parallelogram) and predicted (geometric normal) -- and to output bytes with the stored SHA-256."""
import hashlib
import os

import numpy as np

from oracle import pyoracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load():
    z = np.load(os.path.join(GOLD, "mesh_predictors.npz"))
    o = O.decode(np.fromfile(os.path.join(GOLD, "house_04.obj.drc"), dtype=np.uint8))
    assert o.status == 0
    return z, o


def test_oracle_decodes_the_predictor_goldens():
    z, o = load()
    maps, aoff, n = [o.maps[0], o.maps[1]], int(z["attr_off"]), int(z["n_points"])
    r = O.decode(z["cmp_buf"], maps, aoff, n)
    assert r.status == 0 and [a.pred_method for a in r.attrs] == [4, 4]
    assert np.array_equal(r.attrs[0].qints, z["cmp_pos"]) and np.array_equal(r.attrs[1].qints, z["cmp_gen"])
    assert [hashlib.sha256(a.out.tobytes()).hexdigest() for a in r.attrs] == list(z["cmp_sha"])
    r = O.decode(z["geo_buf"], maps, aoff, n)
    assert r.status == 0 and [a.pred_method for a in r.attrs] == [1, 6]
    assert np.array_equal(r.attrs[1].qints, z["geo_pred"])
    assert [hashlib.sha256(a.out.tobytes()).hexdigest() for a in r.attrs] == list(z["geo_sha"])
