"""Shared helpers of the parity tests: run the same .drc buffers through the CUDA path (C ABI) and the CPU oracle."""
import numpy as np

from draco_sharp_b200 import _native as N
from draco_sharp_b200 import synth_gen as G
from oracle import pyoracle as O


def gpu_decode_all(dec, buffers, want=("out", "symbols", "qints")):
    """Decode `buffers` on the GPU up to three times (plain, DUMP_SYMBOLS, DUMP_QINTS).
    Returns per buffer: dict(status, attrs=[dict(info, out, symbols, qints)])."""
    res = None
    for what in want:
        flags = {"out": 0, "symbols": N.DCB_DUMP_SYMBOLS, "qints": N.DCB_DUMP_QINTS}[what]
        batch = dec.index(buffers)
        out, dbg = dec.decode(batch, flags=flags)
        cur = []
        for k in range(batch.n_bufs):
            bi = batch.buffer_info(k)
            rec = {"status": bi.status, "n_points": bi.n_points, "attrs": []}
            if bi.status == 0:
                for a in range(bi.n_attrs):
                    ai = batch.attr_info(k, a)
                    ncp = 2 if ai.seq_decoder_type == 3 else ai.num_components
                    nv = ai.n_entries * ncp
                    d = {"info": ai, "ncp": ncp}
                    if what == "out":
                        d["out"] = out[ai.out_off: ai.out_off + ai.out_bytes].copy()
                    else:
                        d[what] = dbg[ai.dbg_off: ai.dbg_off + 4 * nv].view(np.int32).copy()
                    rec["attrs"].append(d)
            cur.append(rec)
        batch.free()
        if res is None:
            res = cur
        else:
            for r, c in zip(res, cur):
                assert r["status"] == c["status"]
                for ra, ca in zip(r["attrs"], c["attrs"]):
                    ra.update({k: v for k, v in ca.items() if k not in ("info", "ncp")})
    return res


def compare_with_oracle(gpu, buffers, maps=None, check_ints=True):
    """Bit-exact comparison of the GPU results with the oracle on the same buffers."""
    n_ok = 0
    for k, (g, buf) in enumerate(zip(gpu, buffers)):
        o = O.decode(buf, maps[k] if maps else None)
        assert g["status"] == o.status, "buffer %d: gpu status %d oracle %d" % (k, g["status"], o.status)
        if o.status != 0:
            continue
        n_ok += 1
        assert len(g["attrs"]) == o.n_attrs
        for a, (ga, oa) in enumerate(zip(g["attrs"], o.attrs)):
            ai = ga["info"]
            assert (ai.att_type, ai.data_type, ai.num_components, ai.unique_id, ai.seq_decoder_type) == \
                   (oa.att_type, oa.data_type, oa.nc, oa.unique_id, oa.seq_type)
            assert ai.n_entries == oa.n_entries
            assert ai.out_bytes == oa.out_bytes, (k, a, ai.out_bytes, oa.out_bytes)
            if oa.seq_type != 0:
                assert (ai.pred_method, ai.transform) == (oa.pred_method, oa.transform)
                if oa.n_entries and oa.compressed:
                    assert ai.scheme == oa.scheme
            if "out" in ga:
                # bit-exact, floats included (north_star allows 1 ulp; we hold 0)
                assert np.array_equal(ga["out"], oa.out), "buffer %d attr %d: output bytes differ" % (k, a)
            if check_ints and oa.seq_type != 0:
                if "symbols" in ga:
                    assert np.array_equal(ga["symbols"].view(np.uint32), oa.symbols), "buffer %d attr %d: symbols differ" % (k, a)
                if "qints" in ga:
                    assert np.array_equal(ga["qints"], oa.qints), "buffer %d attr %d: quantized ints differ" % (k, a)
    return n_ok


def cloud(n, seed=1, scheme=-1, pos_bits=14, normal_bits=0, colors=0, rho=(24, 25)):
    sp = G.make_spec(n, seed=seed, pos_bits=pos_bits, scheme=scheme, normal_bits=normal_bits, colors=colors, rho=rho)
    b, _ = G.synth_cloud(sp, want_truth=False)
    return b
