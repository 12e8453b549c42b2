"""Encoders (C++ generator, Python writer) -> CPU oracle -> the source values: the oracle decodes what independent
encoders wrote from the bitstream layout.  Also property-based round trips over nc / bit widths / schemes."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import drc_writer as W
from draco_sharp_b200 import synth_gen as G
from oracle import pyoracle as O


@pytest.mark.parametrize("scheme", [-1, 0, 1])
@pytest.mark.parametrize("n", [0, 1, 2, 7, 1000, 20000])
def test_generator_roundtrip(scheme, n):
    sp = G.make_spec(n, seed=99 + n, scheme=scheme, normal_bits=10, colors=1)
    b, tr = G.synth_cloud(sp)
    r = O.decode(b)
    assert r.status == 0 and r.end_off == len(b) and r.n_points == n
    assert np.array_equal(r.attrs[0].qints.reshape(-1, 3), tr["pos_q"])
    assert np.array_equal(r.attrs[1].qints.reshape(-1, 2), tr["nrm_st"])
    assert np.array_equal(r.attrs[2].out.reshape(-1, 3), tr["rgb"])
    if n:
        assert G.word_checksum(r.attrs[0].out) == tr["sums"][0]
        assert G.word_checksum(r.attrs[2].out) == tr["sums"][2]
        for a in r.attrs:
            # rANS self-check: the decoder ends where the encoder began -- after the pending renormalisation every
            # payload byte is consumed and the state is back at L = 4 * 2^precision
            state, left, L = a.final_state, a.leftover, 4 << a.precision
            while state < L and left > 0:
                left -= 1
                state = state * 256 + int(b[a.payload_off + left])
            assert left == 0 and state == L
        # normals: unit length (or the zero vector)
        nrm = r.attrs[1].out.view(np.float32).reshape(-1, 3).astype(np.float64)
        ln = np.linalg.norm(nrm, axis=1)
        assert np.all((np.abs(ln - 1) < 1e-6) | (ln == 0))
    # dequantised floats: q * (range / maxq) + min with two roundings
    q = tr["pos_q"].astype(np.float32)
    delta = np.float32(2.0) / np.float32((1 << 14) - 1)
    want = (q * delta).astype(np.float32) + np.float32(-1.0)
    assert np.array_equal(r.attrs[0].out.view(np.float32).reshape(-1, 3), want)


def _delta_wrap_corr(q, nc, mn, mx):
    """Encoder side of delta + wrap (PredictionSchemeWrapEncodingTransform.cs:45-88)."""
    q = np.asarray(q, dtype=np.int64).reshape(-1, nc)
    md = 1 + mx - mn
    max_corr = md // 2
    min_corr = -max_corr
    if md % 2 == 0:
        max_corr -= 1
    corr = np.zeros_like(q)
    prev = np.clip(np.zeros(nc, dtype=np.int64), mn, mx)
    for i in range(q.shape[0]):
        c = q[i] - prev
        c = np.where(c < min_corr, c + md, np.where(c > max_corr, c - md, c))
        corr[i] = c
        prev = q[i]
    return corr.reshape(-1)


@settings(max_examples=40, deadline=None)
@given(nc=st.integers(1, 4), bits=st.integers(1, 20), n=st.integers(0, 300), scheme=st.sampled_from(["raw", "tagged", "uncompressed"]),
       seed=st.integers(0, 1 << 30))
def test_writer_roundtrip_property(nc, bits, n, scheme, seed):
    rng = np.random.default_rng(seed)
    mx = (1 << bits) - 1
    q = rng.integers(0, mx + 1, size=(n, nc))
    lo, hi = (int(q.min()), int(q.max())) if n else (0, 0)
    corr = _delta_wrap_corr(q, nc, lo, hi)
    port = W.portable_int(corr, nc, 0, 1, scheme, W.wrap_data(lo, hi), num_bytes=4)
    mins = [float(rng.normal()) for _ in range(nc)]
    buf = W.point_cloud(n, [dict(att_type=0, data_type=9, nc=nc, seq_type=2, portable=port,
                                 xform=W.quant_params(mins, 3.5, bits))])
    r = O.decode(buf)
    assert r.status == 0 and r.end_off == len(buf)
    assert np.array_equal(r.attrs[0].qints.reshape(-1, nc), q)
    delta = np.float32(3.5) / np.float32(mx)
    want = (q.astype(np.float32) * delta).astype(np.float32) + np.asarray(mins, dtype=np.float32)
    assert np.array_equal(r.attrs[0].out.view(np.float32).reshape(-1, nc), want.astype(np.float32))


def test_uncompressed_and_generic_attributes():
    rng = np.random.default_rng(5)
    n = 257
    v = rng.integers(-1000, 1000, size=n * 2)
    gen = rng.integers(0, 256, size=n * 6, dtype=np.uint8)
    attrs = [
        dict(att_type=4, data_type=5, nc=2, seq_type=1, portable=W.portable_int(v, 2, -2, -1, "uncompressed", num_bytes=4)),
        dict(att_type=4, data_type=3, nc=3, seq_type=0, portable=gen.tobytes()),
    ]
    r = O.decode(W.point_cloud(n, attrs))
    assert r.status == 0
    assert np.array_equal(r.attrs[0].out.view(np.int32), v.astype(np.int32))
    assert np.array_equal(r.attrs[1].out, gen)


@pytest.mark.parametrize("w,h", [(2, 2), (3, 2), (17, 9), (120, 77)])
@pytest.mark.parametrize("scheme", [-1, 0, 1])
def test_grid_mesh_generator_roundtrip(w, h, scheme):
    """configs[3] generator: corner table + depth-first maps of a grid, parallelogram corrections written by the
    generator, decoded back by the oracle's MeshPredictionSchemeParallelogramDecoder restatement."""
    topo = G.grid_topology(w, h)
    nv = w * h
    assert sorted(topo["vertex_to_data"].tolist()) == list(range(nv))           # every vertex reached exactly once
    assert np.array_equal(topo["corner_to_vertex"][topo["data_to_corner"]][topo["vertex_to_data"]], np.arange(nv))
    opp = topo["opposite"]
    inner = opp != 0xFFFFFFFF
    assert np.array_equal(opp[opp[inner]], np.nonzero(inner)[0])               # opposite is an involution
    assert int((~inner).sum()) == 2 * (w - 1) + 2 * (h - 1)                      # boundary edges
    buf, aoff, sm, sch, q = G.grid_mesh(w, h, topo, seed=w * 31 + h, scheme=scheme, want_q=True)
    r = O.decode(buf, [topo], aoff, nv)
    assert r.status == 0 and np.array_equal(r.attrs[0].qints, q) and G.word_checksum(r.attrs[0].out) == sm
    assert r.attrs[0].pred_method == 1
