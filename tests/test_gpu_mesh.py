"""-m gpu: the parallelogram path (Edgebreaker meshes; connectivity on the host, prediction on the GPU).

Connectivity comes from the oracle's host Edgebreaker restatement, exactly as the C# host would hand its
CornerTable / traversal maps to dcb_set_mesh_maps."""
import hashlib
import os
import struct

import numpy as np
import pytest

import drc_writer as W
from draco_sharp_b200 import _native as N
from oracle import pyoracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _mesh_buffer(attr_section):
    """DRACO v2.2 Edgebreaker mesh header + an opaque connectivity blob + the given ATTRIBUTES section."""
    head = b"DRACO" + bytes([2, 2, 1, 1]) + struct.pack("<H", 0) + bytes([2]) + b"\xAA" * 37  # connectivity: host business
    return head + attr_section, len(head)


def _decode_mesh(dec, buf, attr_off, n_points, maps, flags=0):
    batch = dec.index([buf])
    assert batch.buffer_info(0).needs_connectivity == 1
    batch.set_attr_section(0, attr_off, n_points)
    for d, m in enumerate(maps):
        batch.set_mesh_maps(0, d, m["opposite"], m["corner_to_vertex"], m["data_to_corner"], m["vertex_to_data"])
    batch.finish()
    out, dbg = dec.decode(batch, flags=flags)
    return batch, out, dbg


def test_house_positions_on_gpu_match_goldens(gpu_decoder):
    """The position attribute of the reference's sample asset, bytes verbatim, through the CUDA parallelogram
    path: SHA-256 goldens of SURVEY.md Appendix C (quantized ints and floats), and the oracle."""
    b = np.fromfile(os.path.join(GOLD, "house_04.obj.drc"), dtype=np.uint8)
    o = O.decode(b)
    a = o.attrs[0]
    start = a.table_off - 5                      # pred, transform, compressed, scheme, max_bit_length
    end = a.payload_off + a.payload_len + 8      # + wrap bounds
    portable = bytes(b[start:end])
    xform = bytes(b[end:end + 17])               # decoder 0 holds one attribute: its XFORM_PARAMS follow directly
    section = bytes([1, 0xFF, 0, 0]) + W.varint(1) + bytes([0, 9, 3, 0]) + W.varint(0) + bytes([2]) + portable + xform
    buf, attr_off = _mesh_buffer(section)
    maps = [o.maps[0]]
    ref = O.decode(np.frombuffer(buf, dtype=np.uint8), maps, attr_off, o.n_points)
    assert ref.status == 0 and np.array_equal(ref.attrs[0].out, a.out)
    for flags, key in ((N.DCB_DUMP_QINTS, "qints"), (N.DCB_DUMP_SYMBOLS, "symbols")):
        batch, out, dbg = _decode_mesh(gpu_decoder, buf, attr_off, o.n_points, maps, flags)
        assert batch.status(0) == 0
        ai = batch.attr_info(0, 0)
        assert ai.n_entries == 1775 and ai.pred_method == 1
        got = out[ai.out_off: ai.out_off + ai.out_bytes]
        assert sha(got) == "028840c055ebfbc5b9a3a04b28d2fc5d0f9cae9c12821f030a815a0826bdcb37"
        ints = dbg[ai.dbg_off: ai.dbg_off + 4 * 5325].view(np.int32)
        if key == "qints":
            assert sha(ints) == "15d5eeb7c1c24707f0bcc20b6ed5e89b6faaa5d46da4fd5f15b5872352e02be6"
        else:
            assert sha(ints) == "00823b9eb65a088cb18987b7016f3756f94fbccb5911d1e86912911af2fcb07b"
        batch.free()


@pytest.mark.parametrize("scheme", ["raw", "tagged", "uncompressed"])
def test_parallelogram_random_corrections(gpu_decoder, scheme):
    """Random corrections over the sample's real connectivity, several attributes per decoder, all sources."""
    b = np.fromfile(os.path.join(GOLD, "house_04.obj.drc"), dtype=np.uint8)
    o = O.decode(b)
    rng = np.random.default_rng(5)
    n0, n1 = o.maps[0]["data_to_corner"].size, o.maps[1]["data_to_corner"].size
    c_pos = rng.integers(-30, 31, size=n0 * 3)
    c_gen = rng.integers(-3, 4, size=n0 * 1)
    c_uv = rng.integers(-9, 10, size=n1 * 2)
    c_pos[rng.integers(0, n0 * 3, 20)] = rng.integers(-5000, 5000, 20)  # clamp + wrap corner cases
    sec = bytearray([2, 0xFF, 0, 0, 0, 1, 0])
    sec += W.varint(2) + bytes([0, 9, 3, 0]) + W.varint(0) + bytes([4, 2, 1, 0]) + W.varint(1) + bytes([2, 1])
    sec += W.varint(1) + bytes([3, 9, 2, 0]) + W.varint(2) + bytes([2])
    sec += W.portable_int(c_pos, 3, 1, 1, scheme, W.wrap_data(0, 4095), num_bytes=2)
    sec += W.portable_int(c_gen, 1, 1, 1, scheme, W.wrap_data(0, 255), num_bytes=1)
    sec += W.quant_params([1.0, 2.0, 3.0], 10.0, 12)
    sec += W.portable_int(c_uv, 2, 1, 1, scheme, W.wrap_data(0, 1023), num_bytes=2)
    sec += W.quant_params([0.0, 0.0], 1.0, 10)
    buf, attr_off = _mesh_buffer(bytes(sec))
    maps = [o.maps[0], o.maps[1]]
    ref = O.decode(np.frombuffer(buf, dtype=np.uint8), maps, attr_off, o.n_points)
    assert ref.status == 0
    batch, out, dbg = _decode_mesh(gpu_decoder, buf, attr_off, o.n_points, maps, N.DCB_DUMP_QINTS)
    assert batch.status(0) == 0
    for k, ra in enumerate(ref.attrs):
        ai = batch.attr_info(0, k)
        assert ai.n_entries == ra.n_entries
        assert np.array_equal(dbg[ai.dbg_off: ai.dbg_off + 4 * ra.qints.size].view(np.int32), ra.qints), k
        assert np.array_equal(out[ai.out_off: ai.out_off + ai.out_bytes], ra.out), k
    batch.free()


@pytest.mark.parametrize("scheme,n_flags,uv_bits", [("raw", 3300, 10), ("tagged", 3300, 12), ("uncompressed", 3300, 10),
                                                     ("raw", 700, 10), ("raw", 0, 10), ("uncompressed", 3300, 30)])
def test_texcoords_portable_random_corrections(gpu_decoder, scheme, n_flags, uv_bits):
    """TexCoordsPortable predictor over the sample's real connectivity with random corrections and random
    orientation flags, every symbol source; too few flags must fail the buffer exactly where the oracle fails it
    (:125), and 30-bit coordinates exercise the 64-bit overflow guards (:85-:90)."""
    b = np.fromfile(os.path.join(GOLD, "house_04.obj.drc"), dtype=np.uint8)
    o = O.decode(b)
    rng = np.random.default_rng(17 + n_flags + uv_bits)
    n0, n1 = o.maps[0]["data_to_corner"].size, o.maps[1]["data_to_corner"].size
    big = uv_bits >= 30          # 30-bit positions and coordinates: |prev - next|^2 * |uv| leaves int64 (:85)
    pos_hi = (1 << 30) - 1 if big else 4095
    c_pos = rng.integers(-(1 << 28), 1 << 28, size=n0 * 3) if big else rng.integers(-30, 31, size=n0 * 3)
    hi = (1 << uv_bits) - 1
    c_uv = rng.integers(-(1 << 28), 1 << 28, size=n1 * 2) if big else rng.integers(-9, 10, size=n1 * 2)
    flags = rng.integers(0, 2, size=n_flags)
    sec = bytearray([2, 0xFF, 0, 0, 0, 1, 0])
    sec += W.varint(1) + bytes([0, 9, 3, 0]) + W.varint(0) + bytes([2])
    sec += W.varint(1) + bytes([3, 9, 2, 0]) + W.varint(1) + bytes([2])
    sec += W.portable_int(c_pos, 3, 1, 1, scheme, W.wrap_data(0, pos_hi), num_bytes=4)
    sec += W.quant_params([1.0, 2.0, 3.0], 10.0, 30 if big else 12)
    sec += W.portable_int(c_uv, 2, 5, 1, scheme, W.tex_coords_data(flags, 0, hi), num_bytes=4)
    sec += W.quant_params([0.0, 0.0], 1.0, uv_bits)
    buf, attr_off = _mesh_buffer(bytes(sec))
    maps = [o.maps[0], o.maps[1]]
    ref = O.decode(np.frombuffer(buf, dtype=np.uint8), maps, attr_off, o.n_points)
    batch, out, dbg = _decode_mesh(gpu_decoder, buf, attr_off, o.n_points, maps, N.DCB_DUMP_QINTS)
    assert batch.status(0) == ref.status
    assert (ref.status == 0) == (n_flags >= 3300 and not big)
    if ref.status == 0:
        for k, ra in enumerate(ref.attrs):
            ai = batch.attr_info(0, k)
            assert ai.n_entries == ra.n_entries
            assert np.array_equal(dbg[ai.dbg_off: ai.dbg_off + 4 * ra.qints.size].view(np.int32), ra.qints), k
            assert np.array_equal(out[ai.out_off: ai.out_off + ai.out_bytes], ra.out), k
    batch.free()


def _cmp_attr(values, nc, maps, bits, rng, scheme, p_crease=0.3, drop_flags=0):
    hi = (1 << bits) - 1
    corr, crease = W.cmp_encode(values, nc, maps, 0, hi, rng, p_crease)
    if drop_flags:
        k = max(range(4), key=lambda i: len(crease[i]))
        crease[k] = crease[k][:-drop_flags]
    return W.portable_int(corr, nc, 4, 1, scheme, W.cmp_data(crease, 0, hi), num_bytes=4)


@pytest.mark.parametrize("scheme,p_crease,drop", [("raw", 0.3, 0), ("tagged", 0.3, 0), ("uncompressed", 0.5, 0), ("raw", 0.0, 0),
                                                  ("raw", 1.0, 0), ("raw", 0.3, 1)])
def test_constrained_multi_parallelogram_round_trip(gpu_decoder, scheme, p_crease, drop):
    """ConstrainedMultiParallelogram predictor over the sample's real connectivity: three attributes in two
    decoders (3, 1 and 2 components; the second decoder's maps carry attribute seams), values chosen by the test, encoded
    by the bitstream-specification encoder in tests/drc_writer.py.  The CUDA path must return exactly those values, equal
    the oracle in every output byte, and fail the buffer like the oracle when a flag sequence is one flag short (:93)."""
    b = np.fromfile(os.path.join(GOLD, "house_04.obj.drc"), dtype=np.uint8)
    o = O.decode(b)
    rng = np.random.default_rng(31 + int(p_crease * 10) + drop)
    n0, n1 = o.maps[0]["data_to_corner"].size, o.maps[1]["data_to_corner"].size
    v_pos = rng.integers(0, 4096, size=n0 * 3)
    v_gen = (np.cumsum(rng.integers(-2, 3, size=n0)) + 128).clip(0, 255)
    v_uv = rng.integers(0, 1024, size=n1 * 2)
    sec = bytearray([2, 0xFF, 0, 0, 0, 1, 0])
    sec += W.varint(2) + bytes([0, 9, 3, 0]) + W.varint(0) + bytes([4, 2, 1, 0]) + W.varint(1) + bytes([2, 1])
    sec += W.varint(1) + bytes([3, 9, 2, 0]) + W.varint(2) + bytes([2])
    sec += _cmp_attr(v_pos, 3, o.maps[0], 12, rng, scheme, p_crease, drop)
    sec += _cmp_attr(v_gen, 1, o.maps[0], 8, rng, scheme, p_crease)
    sec += W.quant_params([1.0, 2.0, 3.0], 10.0, 12)
    sec += _cmp_attr(v_uv, 2, o.maps[1], 10, rng, scheme, p_crease)
    sec += W.quant_params([0.0, 0.0], 1.0, 10)
    buf, attr_off = _mesh_buffer(bytes(sec))
    maps = [o.maps[0], o.maps[1]]
    ref = O.decode(np.frombuffer(buf, dtype=np.uint8), maps, attr_off, o.n_points)
    batch, out, dbg = _decode_mesh(gpu_decoder, buf, attr_off, o.n_points, maps, N.DCB_DUMP_QINTS)
    assert batch.status(0) == ref.status
    assert (ref.status == 0) == (drop == 0)
    if ref.status == 0:
        for k, (ra, want) in enumerate(zip(ref.attrs, (v_pos, v_gen, v_uv))):
            ai = batch.attr_info(0, k)
            assert ai.n_entries == ra.n_entries and ai.pred_method == 4
            q = dbg[ai.dbg_off: ai.dbg_off + 4 * ra.qints.size].view(np.int32)
            assert np.array_equal(q, want.astype(np.int32)), k
            assert np.array_equal(q, ra.qints), k
            assert np.array_equal(out[ai.out_off: ai.out_off + ai.out_bytes], ra.out), k
    batch.free()


def test_constrained_multi_parallelogram_with_texcoords_and_grid(gpu_decoder):
    """(1) Positions by ConstrainedMultiParallelogram feeding the TexCoordsPortable predictor of the same buffer (what
    upstream encoders emit at their slowest speeds); (2) a 300 x 300 grid surface whose parallelogram operands lie ~300
    entries back (gathered from the scratch, not the history ring), next to a plain-parallelogram mesh in one batch."""
    from draco_sharp_b200 import synth_gen as G
    b = np.fromfile(os.path.join(GOLD, "house_04.obj.drc"), dtype=np.uint8)
    o = O.decode(b)
    rng = np.random.default_rng(77)
    n0, n1 = o.maps[0]["data_to_corner"].size, o.maps[1]["data_to_corner"].size
    v_pos = (np.cumsum(rng.integers(-9, 10, size=(n0, 3)), axis=0) + 2000).clip(0, 4095).ravel()
    c_uv = rng.integers(-9, 10, size=n1 * 2)
    flags = rng.integers(0, 2, size=3300)
    sec = bytearray([2, 0xFF, 0, 0, 0, 1, 0])
    sec += W.varint(1) + bytes([0, 9, 3, 0]) + W.varint(0) + bytes([2])
    sec += W.varint(1) + bytes([3, 9, 2, 0]) + W.varint(1) + bytes([2])
    sec += _cmp_attr(v_pos, 3, o.maps[0], 12, rng, "raw") + W.quant_params([1.0, 2.0, 3.0], 10.0, 12)
    sec += W.portable_int(c_uv, 2, 5, 1, "raw", W.tex_coords_data(flags, 0, 1023), num_bytes=4) + W.quant_params([0.0, 0.0], 1.0, 10)
    house, house_off = _mesh_buffer(bytes(sec))
    w = h = 300
    topo = G.grid_topology(w, h)
    yy, xx = np.mgrid[0:h, 0:w]
    surf = np.stack([xx * 50, yy * 50, (2000 + 1500 * np.sin(xx / 17.0) * np.cos(yy / 23.0)).astype(np.int64)], axis=-1).reshape(-1, 3)
    v_grid = np.zeros_like(surf)
    v_grid[topo["vertex_to_data"]] = surf          # entry order
    v_grid = v_grid.ravel()
    gsec = bytearray([1, 0xFF, 0, 0]) + W.varint(1) + bytes([0, 9, 3, 0]) + W.varint(0) + bytes([2])
    gsec += _cmp_attr(v_grid, 3, topo, 14, rng, "tagged", 0.2) + W.quant_params([-1.0, -1.0, -1.0], 2.0, 14)
    grid, grid_off = _mesh_buffer(bytes(gsec))
    plain = G.grid_mesh(w, h, topo, seed=5, want_q=True)
    batch = gpu_decoder.index([house, grid, plain[0]])
    batch.set_attr_section(0, house_off, o.n_points)
    for dd in (0, 1):
        m = o.maps[dd]
        batch.set_mesh_maps(0, dd, m["opposite"], m["corner_to_vertex"], m["data_to_corner"], m["vertex_to_data"])
    batch.set_attr_section(1, grid_off, w * h)
    batch.set_attr_section(2, plain[1], w * h)
    for k in (1, 2):
        batch.set_mesh_maps(k, 0, topo["opposite"], topo["corner_to_vertex"], topo["data_to_corner"], topo["vertex_to_data"])
    batch.finish()
    out, dbg = gpu_decoder.decode(batch, flags=N.DCB_DUMP_QINTS)
    refs = [O.decode(np.frombuffer(house, dtype=np.uint8), [o.maps[0], o.maps[1]], house_off, o.n_points),
            O.decode(np.frombuffer(grid, dtype=np.uint8), [topo], grid_off, w * h),
            O.decode(plain[0], [topo], plain[1], w * h)]
    for k, ref in enumerate(refs):
        assert batch.status(k) == ref.status == 0, (k, batch.status(k), ref.status)
        for a, ra in enumerate(ref.attrs):
            ai = batch.attr_info(k, a)
            assert np.array_equal(dbg[ai.dbg_off: ai.dbg_off + 4 * ra.qints.size].view(np.int32), ra.qints), (k, a)
            assert np.array_equal(out[ai.out_off: ai.out_off + ai.out_bytes], ra.out), (k, a)
    assert np.array_equal(refs[0].attrs[0].qints, v_pos.astype(np.int32))
    assert np.array_equal(refs[1].attrs[0].qints, v_grid.astype(np.int32))
    batch.free()


@pytest.mark.parametrize("dec,nbits,pos_bits,canonical,scheme", [(0, 10, 12, True, "raw"), (1, 8, 12, True, "tagged"),
                                                                 (1, 30, 30, True, "uncompressed"), (1, 2, 12, True, "raw"),
                                                                 (0, 12, 14, False, "tagged"), (1, 10, 12, True, "uncompressed")])
def test_geometric_normal_predictor(gpu_decoder, dec, nbits, pos_bits, canonical, scheme):
    """GeometricNormal predictor over the sample's real connectivity, normals in the positions' decoder or
    in the second one (attribute seams), every symbol source.  With all-zero corrections the decoded octahedral
    coordinates must equal the predictions computed by the bitstream-specification restatement in tests/drc_writer.py;
    with random corrections every quantized int and output byte must equal the oracle's (30-bit positions exercise the
    wrapping 64-bit sums and the scaling below 2^29)."""
    from test_geometric_normal_cpu import normals_section
    b = np.fromfile(os.path.join(GOLD, "house_04.obj.drc"), dtype=np.uint8)
    o = O.decode(b)
    rng = np.random.default_rng(dec * 7 + nbits)
    n0 = o.maps[0]["data_to_corner"].size
    m = o.maps[dec]
    n = m["data_to_corner"].size
    big = pos_bits >= 30
    c_pos = rng.integers(-(1 << 28), 1 << 28, size=n0 * 3) if big else rng.integers(-40, 41, size=n0 * 3)
    flips = rng.integers(0, 2, size=n)
    pos_scheme = "uncompressed" if big else "raw"
    maps = [o.maps[0], o.maps[1]]
    for zero in (True, False):
        corr = np.zeros(n * 2, dtype=np.int64) if zero else rng.integers(0, 1 << nbits, size=n * 2)
        sec = normals_section(c_pos, corr, flips, nbits, pos_bits, dec, scheme, canonical, pos_scheme)
        buf, attr_off = _mesh_buffer(sec)
        ref = O.decode(np.frombuffer(buf, dtype=np.uint8), maps, attr_off, o.n_points)
        assert ref.status == 0
        batch, out, dbg = _decode_mesh(gpu_decoder, buf, attr_off, o.n_points, maps, N.DCB_DUMP_QINTS)
        assert batch.status(0) == 0
        for k, ra in enumerate(ref.attrs):
            ai = batch.attr_info(0, k)
            q = dbg[ai.dbg_off: ai.dbg_off + 4 * ra.qints.size].view(np.int32)
            assert np.array_equal(q, ra.qints), (zero, k)
            assert np.array_equal(out[ai.out_off: ai.out_off + ai.out_bytes], ra.out), (zero, k)
            if k == 1 and zero:
                assert ai.pred_method == 6
                want = W.geometric_normal_predictions(m, o.maps[0], ref.attrs[0].qints, nbits, flips)
                assert np.array_equal(q, np.asarray(want, dtype=np.int32))
        batch.free()
        # and without the dump flag (the template instantiation a plain decode uses)
        batch, out2, _ = _decode_mesh(gpu_decoder, buf, attr_off, o.n_points, maps, 0)
        ai = batch.attr_info(0, 1)
        assert np.array_equal(out2[ai.out_off: ai.out_off + ai.out_bytes], ref.attrs[1].out)
        batch.free()


def test_geometric_normal_behind_cmp_positions_on_a_grid(gpu_decoder):
    """A 200 x 200 grid surface the way upstream encoders write meshes at their slowest speeds: positions by
    ConstrainedMultiParallelogram, normals by GeometricNormal in the same decoder; next to a cloud in one batch."""
    from draco_sharp_b200 import synth_gen as G
    w = h = 200
    topo = G.grid_topology(w, h)
    rng = np.random.default_rng(11)
    yy, xx = np.mgrid[0:h, 0:w]
    surf = np.stack([xx * 80, yy * 80, (8000 + 6000 * np.sin(xx / 13.0) * np.cos(yy / 19.0)).astype(np.int64)], axis=-1).reshape(-1, 3)
    v_pos = np.zeros_like(surf)
    v_pos[topo["vertex_to_data"]] = surf
    v_pos = v_pos.ravel()
    n = w * h
    flips = rng.integers(0, 2, size=n)
    corr_n = rng.integers(0, 40, size=n * 2)
    sec = bytearray([1, 0xFF, 0, 0])
    sec += W.varint(2) + bytes([0, 9, 3, 0]) + W.varint(0) + bytes([1, 9, 3, 0]) + W.varint(1) + bytes([2, 3])
    sec += _cmp_attr(v_pos, 3, topo, 14, rng, "raw", 0.2)
    sec += W.portable_int(corr_n, 2, 6, 3, "raw", W.geometric_normal_data(flips, 10), zig=False)
    sec += W.quant_params([-1.0, -1.0, -1.0], 2.0, 14) + bytes([10])
    mesh, aoff = _mesh_buffer(bytes(sec))
    cloud = G.synth_cloud(G.make_spec(4000, seed=8, normal_bits=10, colors=1))[0]
    batch = gpu_decoder.index([cloud, mesh])
    batch.set_attr_section(1, aoff, n)
    batch.set_mesh_maps(1, 0, topo["opposite"], topo["corner_to_vertex"], topo["data_to_corner"], topo["vertex_to_data"])
    batch.finish()
    out, dbg = gpu_decoder.decode(batch, flags=N.DCB_DUMP_QINTS)
    refs = [O.decode(cloud), O.decode(np.frombuffer(mesh, dtype=np.uint8), [topo], aoff, n)]
    for k, ref in enumerate(refs):
        assert batch.status(k) == ref.status == 0, (k, batch.status(k), ref.status)
        for a, ra in enumerate(ref.attrs):
            ai = batch.attr_info(k, a)
            assert np.array_equal(out[ai.out_off: ai.out_off + ai.out_bytes], ra.out), (k, a)
            if ra.seq_type != 0:
                assert np.array_equal(dbg[ai.dbg_off: ai.dbg_off + 4 * ra.qints.size].view(np.int32), ra.qints), (k, a)
    assert np.array_equal(refs[1].attrs[0].qints, v_pos.astype(np.int32))
    nrm = refs[1].attrs[1].out.view(np.float32).reshape(-1, 3)
    assert np.allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-5)
    batch.free()


@pytest.mark.parametrize("kind", ["cmp", "geo"])
def test_mesh_predictor_status_parity_on_malformed_buffers(gpu_decoder, kind):
    """Truncated and bit-flipped constrained-multi-parallelogram / geometric-normal buffers in ONE batch: every buffer
    ends with the oracle's status (parse errors from the walker, run-out-of-flags and map errors from the kernels), the
    intact ones decode bit-exactly next to them."""
    from test_cmp_cpu import cmp_section
    from test_geometric_normal_cpu import normals_section
    from test_mesh_predictors_walker_cpu import variants
    b = np.fromfile(os.path.join(GOLD, "house_04.obj.drc"), dtype=np.uint8)
    o = O.decode(b)
    rng = np.random.default_rng(99)
    n0, n1 = o.maps[0]["data_to_corner"].size, o.maps[1]["data_to_corner"].size
    maps = [o.maps[0], o.maps[1]]
    if kind == "cmp":
        sec, _ = cmp_section(rng.integers(0, 4096, size=n0 * 3), 3, o.maps[0], 12, rng)
    else:
        sec = normals_section(rng.integers(-40, 41, size=n0 * 3), rng.integers(0, 1024, size=n1 * 2), rng.integers(0, 2, size=n1), 10)
    buf, aoff = _mesh_buffer(sec)
    bufs = variants(np.frombuffer(buf, dtype=np.uint8).copy(), aoff, rng)
    refs = [O.decode(v, maps, aoff, o.n_points) for v in bufs]
    batch = gpu_decoder.index(bufs)
    for k in range(len(bufs)):
        batch.set_attr_section(k, aoff, o.n_points)
        for dd, m in enumerate(maps):
            batch.set_mesh_maps(k, dd, m["opposite"], m["corner_to_vertex"], m["data_to_corner"], m["vertex_to_data"])
    batch.finish()
    out, _ = gpu_decoder.decode(batch)
    n_ok = 0
    for k, ref in enumerate(refs):
        assert batch.status(k) == ref.status, (k, len(bufs[k]), batch.status(k), ref.status)
        if ref.status == 0:
            n_ok += 1
            for a, ra in enumerate(ref.attrs):
                ai = batch.attr_info(k, a)
                assert np.array_equal(out[ai.out_off: ai.out_off + ai.out_bytes], ra.out), (k, a)
    assert 1 <= n_ok < len(bufs)
    batch.free()


def test_predictor_goldens_on_gpu_without_the_oracle(gpu_decoder):
    """tests/golden/mesh_predictors.npz through the CUDA path with NOTHING of oracle/ involved: the connectivity maps come
    from the product's own host helper run on the reference's sample (dcb_host_connectivity + dcb_mesh_map), the expected
    values from the fixture (what the bitstream-specification encoder encoded / predicted, SHA-256 of the output bytes)."""
    z = np.load(os.path.join(GOLD, "mesh_predictors.npz"))
    house = np.fromfile(os.path.join(GOLD, "house_04.obj.drc"), dtype=np.uint8)
    hb = gpu_decoder.index([house])
    hb.host_connectivity(0)
    maps = [[hb.mesh_map(0, d, w).copy() for w in range(4)] for d in range(2)]
    hb.finish()
    hb.free()
    aoff, n = int(z["attr_off"]), int(z["n_points"])
    bufs = [z["cmp_buf"], z["geo_buf"]]
    batch = gpu_decoder.index(bufs)
    for k in range(2):
        batch.set_attr_section(k, aoff, n)
        for d in range(2):
            batch.set_mesh_maps(k, d, *maps[d])
    batch.finish()
    out, dbg = gpu_decoder.decode(batch, flags=N.DCB_DUMP_QINTS)
    assert batch.status(0) == 0 and batch.status(1) == 0
    want_q = {(0, 0): z["cmp_pos"], (0, 1): z["cmp_gen"], (1, 1): z["geo_pred"]}
    want_sha = {0: list(z["cmp_sha"]), 1: list(z["geo_sha"])}
    for k in range(2):
        for a in range(2):
            ai = batch.attr_info(k, a)
            assert sha(out[ai.out_off: ai.out_off + ai.out_bytes]) == want_sha[k][a], (k, a)
            if (k, a) in want_q:
                w = want_q[(k, a)]
                assert np.array_equal(dbg[ai.dbg_off: ai.dbg_off + 4 * w.size].view(np.int32), w), (k, a)
    batch.free()


def test_mesh_without_maps_fails_cleanly(gpu_decoder):
    sec = bytes([1, 0xFF, 0, 0]) + W.varint(1) + bytes([0, 9, 3, 0]) + W.varint(0) + bytes([2])
    sec += W.portable_int(np.zeros(30, dtype=np.int64), 3, 1, 1, "raw", W.wrap_data(0, 7)) + W.quant_params([0, 0, 0], 1.0, 3)
    buf, attr_off = _mesh_buffer(sec)
    batch = gpu_decoder.index([buf])
    with pytest.raises(Exception):
        gpu_decoder.decode(batch)          # dcb_index_finish was never run: DCB_ERR_STATE
    batch.set_attr_section(0, attr_off, 10)
    batch.finish()
    assert batch.status(0) == -14          # DCB_ERR_MAPS
    batch.free()


def _house_positions_only():
    """house_04's header + Edgebreaker connectivity, bytes verbatim, followed by an ATTRIBUTES section that keeps
    only the position attribute (also verbatim): a complete, valid Edgebreaker mesh the product decodes alone."""
    b = np.fromfile(os.path.join(GOLD, "house_04.obj.drc"), dtype=np.uint8)
    o = O.decode(b)
    a = o.attrs[0]
    start = a.table_off - 5
    end = a.payload_off + a.payload_len + 8
    section = (bytes([1, 0xFF, 0, 0]) + W.varint(1) + bytes([0, 9, 3, 0]) + W.varint(0) + bytes([2])
               + bytes(b[start:end]) + bytes(b[end:end + 17]))
    return bytes(b[:1158]) + section, o


def test_house_mesh_end_to_end_through_product_host_helper(gpu_decoder):
    """No oracle on the decode path: dcb_host_connectivity (product, host) -> dcb_index_finish -> CUDA
    parallelogram kernels.  Positions and faces against the SHA-256 goldens / the OBJ the asset was made from."""
    buf, o = _house_positions_only()
    (d,) = gpu_decoder.decode_batch([buf])
    assert d.ok and d.points_count == 3220 and d.header.encoder_type == 1
    pos = d.get_named_attribute(0)
    assert pos.unique_entries_count == 1775
    assert sha(pos.buffer) == "028840c055ebfbc5b9a3a04b28d2fc5d0f9cae9c12821f030a815a0826bdcb37"
    assert d.faces.shape == (2588, 3) and np.array_equal(d.faces, o.faces)
    # every decoded vertex within half a quantisation step of a vertex of the source OBJ
    obj = np.load(os.path.join(GOLD, "house_04_obj_vertices.npy")).astype(np.float64)
    v = pos.values.astype(np.float64)
    dist = np.abs(v[:, None, :] - obj[None, :, :]).max(axis=2).min(axis=1)
    assert dist.max() < 0.49094


def test_house_whole_sample_on_gpu(gpu_decoder):
    """The reference's sample asset, bytes verbatim and whole: positions (parallelogram), texture coordinates
    (TexCoordsPortable predictor, SURVEY 8f-3: rABS orientation flags, 64-bit projection, IntSqrt) and the generic uint8
    attribute (parallelogram -> narrowing store) through the CUDA kernels; connectivity from the product's host helper.
    Quantized ints and output bytes against the oracle and the SHA-256 goldens pinned in test_oracle_kats.py."""
    b = np.fromfile(os.path.join(GOLD, "house_04.obj.drc"), dtype=np.uint8)
    o = O.decode(b)
    assert o.status == 0
    for flags in (0, N.DCB_DUMP_QINTS):
        batch = gpu_decoder.index([b])
        batch.host_connectivity(0)
        batch.finish()
        out, dbg = gpu_decoder.decode(batch, flags=flags)
        assert batch.status(0) == 0
        assert batch.buffer_info(0).n_attrs == 3
        for k in range(3):
            ai = batch.attr_info(0, k)
            got = out[ai.out_off: ai.out_off + ai.out_bytes]
            assert np.array_equal(got, o.attrs[k].out.view(np.uint8).ravel()), k
            if flags:
                nv = ai.n_entries * o.attrs[k].nc_portable
                ints = dbg[ai.dbg_off: ai.dbg_off + 4 * nv].view(np.int32)
                assert np.array_equal(ints, o.attrs[k].qints.ravel()), k
        ai = batch.attr_info(0, 1)
        assert (ai.pred_method, ai.n_entries) == (5, 3220)
        assert sha(out[ai.out_off: ai.out_off + ai.out_bytes]) == "26bd6cc9ae35d081caba5b2061dd9d6d049a66c9050f7c5b9cfdc345c3c8449a"
        if flags:
            assert sha(dbg[ai.dbg_off: ai.dbg_off + 4 * 6440]) == "a243b8cf61c7145d3f75eda19721098918e735e925822b6a211ae6686a6d2990"
        ai = batch.attr_info(0, 0)
        assert sha(out[ai.out_off: ai.out_off + ai.out_bytes]) == "028840c055ebfbc5b9a3a04b28d2fc5d0f9cae9c12821f030a815a0826bdcb37"
        batch.free()
    # and through the reference-shaped host mirror: Draco / PointAttribute objects
    (d,) = gpu_decoder.decode_batch([b])
    assert d.ok and d.points_count == 3220
    uv = d.get_named_attribute(3)
    assert uv is not None and uv.unique_entries_count == 3220
    vts = np.load(os.path.join(GOLD, "house_04_obj_texcoords.npy"))
    half_step = float(o.attrs[1].qrange) / ((1 << o.attrs[1].qbits) - 1) / 2
    dist = np.abs(uv.values.astype(np.float64)[:, None, :] - vts[None, :, :]).max(axis=2).min(axis=1)
    assert dist.max() < half_step


@pytest.mark.parametrize("w,h,method,scheme", [(20, 17, 0, "raw"), (20, 17, 0, "tagged"), (300, 300, 1, "raw"), (300, 300, 0, "tagged")])
def test_sequential_meshes_decode_through_the_product(gpu_decoder, w, h, method, scheme):
    """Sequential meshes (MeshSequentialDecoder.cs:8-118): indices by the product's host helper, attributes
    (delta + wrap positions, octahedral normals, 8-bit colours) through the sequential CUDA kernels; next to a point
    cloud and an Edgebreaker mesh in one batch."""
    rng = np.random.default_rng(w + h + method)
    n = w * h
    faces = [(y * w + x, y * w + x + 1, (y + 1) * w + x) for y in range(h - 1) for x in range(w - 1)]
    attrs = [dict(att_type=0, data_type=9, nc=3, seq_type=2,
                  portable=W.portable_int(rng.integers(-20, 21, size=n * 3), 3, 0, 1, scheme, W.wrap_data(0, 4095), num_bytes=2),
                  xform=W.quant_params([0.0, 1.0, 2.0], 8.0, 12)),
             dict(att_type=2, data_type=2, nc=3, seq_type=1,
                  portable=W.portable_int(rng.integers(-3, 4, size=n * 3), 3, 0, 1, "raw", W.wrap_data(0, 255)))]
    mesh = np.frombuffer(W.sequential_mesh(faces, n, attrs, method, scheme), dtype=np.uint8)
    cloud = np.frombuffer(W.point_cloud(n, attrs), dtype=np.uint8)
    house = np.fromfile(os.path.join(GOLD, "house_04.obj.drc"), dtype=np.uint8)
    refs = [O.decode(b) for b in (mesh, cloud, house)]
    assert all(r.status == 0 for r in refs)
    got = gpu_decoder.decode_batch([mesh, cloud, house])
    for d, r in zip(got, refs):
        assert d.ok and d.points_count == r.n_points and len(d.attributes) == len(r.attrs)
        for a, ra in zip(d.attributes, r.attrs):
            assert np.array_equal(np.asarray(a.buffer).view(np.uint8).ravel(), ra.out.view(np.uint8).ravel())
    assert np.array_equal(got[0].faces, refs[0].faces) and np.array_equal(got[0].faces, np.asarray(faces, dtype=np.uint32))
    assert got[1].faces is None and np.array_equal(got[2].faces, refs[2].faces)


def test_mixed_batch_meshes_and_clouds(gpu_decoder):
    """Meshes (host connectivity) and point clouds in one batch; a mesh with broken connectivity fails alone."""
    from draco_sharp_b200 import synth_gen as G
    buf, o = _house_positions_only()
    bad = bytearray(buf)
    bad[20:60] = bytes(40)
    cloud, tr = G.synth_cloud(G.make_spec(5000, seed=3))
    res = gpu_decoder.decode_batch([cloud, buf, bytes(bad), buf, cloud])
    assert [r.ok for r in res] == [True, True, False, True, True]
    assert sha(res[1].attributes[0].buffer) == sha(res[3].attributes[0].buffer) == \
        "028840c055ebfbc5b9a3a04b28d2fc5d0f9cae9c12821f030a815a0826bdcb37"
    assert G.word_checksum(res[0].attributes[0].buffer) == tr["sums"][0]
    assert np.array_equal(res[0].attributes[0].buffer, res[4].attributes[0].buffer)


@pytest.mark.parametrize("w,h,scheme", [(2, 2, -1), (3, 2, 1), (17, 9, 0), (64, 64, 1), (300, 300, -1), (300, 300, 0)])
def test_grid_meshes_match_oracle_and_generator(gpu_decoder, w, h, scheme):
    """BASELINE configs[3] shape at test size: a batch of grid meshes sharing one topology, parallelogram + wrap."""
    from draco_sharp_b200 import synth_gen as G
    topo = G.grid_topology(w, h)
    meshes = [G.grid_mesh(w, h, topo, seed=100 + k, scheme=scheme, want_q=True) for k in range(5)]
    batch = gpu_decoder.index([m[0] for m in meshes])
    for k, m in enumerate(meshes):
        batch.set_attr_section(k, m[1], w * h)
        batch.set_mesh_maps(k, 0, topo["opposite"], topo["corner_to_vertex"], topo["data_to_corner"], topo["vertex_to_data"])
    batch.finish()
    out, dbg = gpu_decoder.decode(batch, flags=N.DCB_DUMP_QINTS)
    for k, (buf, aoff, sm, sch, q) in enumerate(meshes):
        assert batch.status(k) == 0
        ai = batch.attr_info(k, 0)
        ref = O.decode(buf, [topo], aoff, w * h)
        assert ref.status == 0 and np.array_equal(ref.attrs[0].qints, q)
        assert np.array_equal(dbg[ai.dbg_off: ai.dbg_off + 4 * q.size].view(np.int32), q)
        got = out[ai.out_off: ai.out_off + ai.out_bytes]
        assert np.array_equal(got, ref.attrs[0].out) and G.word_checksum(got) == sm
    batch.free()


@pytest.mark.parametrize("scheme", [-1, 0])
def test_million_vertex_mesh_full_size(gpu_decoder, scheme):
    """BASELINE configs[3] at full size: one 1,000 x 1,000 grid mesh (3M-symbol chains, dependencies ~1,000 entries
    back): every quantized int and output byte against the oracle, the output checksum against the generator."""
    from draco_sharp_b200 import synth_gen as G
    w = h = 1000
    topo = G.grid_topology(w, h)
    buf, aoff, sm, sch, q = G.grid_mesh(w, h, topo, seed=77, scheme=scheme, want_q=True)
    batch = gpu_decoder.index([buf])
    batch.set_attr_section(0, aoff, w * h)
    batch.set_mesh_maps(0, 0, topo["opposite"], topo["corner_to_vertex"], topo["data_to_corner"], topo["vertex_to_data"])
    batch.finish()
    out, dbg = gpu_decoder.decode(batch, flags=N.DCB_DUMP_QINTS)
    assert batch.status(0) == 0
    ai = batch.attr_info(0, 0)
    assert np.array_equal(dbg[ai.dbg_off: ai.dbg_off + 4 * q.size].view(np.int32), q)
    got = out[ai.out_off: ai.out_off + ai.out_bytes]
    assert G.word_checksum(got) == sm
    ref = O.decode(buf, [topo], aoff, w * h)
    assert ref.status == 0 and np.array_equal(got, ref.attrs[0].out)
    batch.free()


def test_meshes_and_clouds_through_pipeline_slices(gpu_decoder):
    """Pipeline slices (the device listed three times) with meshes in the batch: host connectivity, maps upload and the
    parallelogram kernels per slice; same bytes as the single-shard decode."""
    import draco_sharp_b200 as D
    from draco_sharp_b200 import synth_gen as G
    mesh, o = _house_positions_only()
    bufs = []
    for k in range(18):
        bufs.append(mesh if k % 3 == 1 else G.synth_cloud(G.make_spec(3000 + 100 * k, seed=40 + k, scheme=k % 2, colors=1))[0])
    one = gpu_decoder.decode_batch(bufs)
    dec = D.DracoBatchDecoder([0, 0, 0])
    try:
        many = dec.decode_batch(bufs)
    finally:
        dec.close()
    for k, (x, y) in enumerate(zip(one, many)):
        assert x.ok and y.ok, k
        assert len(x.attributes) == len(y.attributes)
        for a, b in zip(x.attributes, y.attributes):
            assert np.array_equal(a.buffer, b.buffer), k
        if k % 3 == 1:
            assert np.array_equal(x.faces, y.faces) and sha(y.attributes[0].buffer) == \
                "028840c055ebfbc5b9a3a04b28d2fc5d0f9cae9c12821f030a815a0826bdcb37"
