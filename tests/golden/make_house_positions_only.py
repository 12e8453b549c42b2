"""Generates tests/golden/house_04_positions_only.drc: the reference's sample mesh (header + Edgebreaker connectivity,
bytes verbatim) followed by an ATTRIBUTES section that keeps only the position attribute (portable data and
quantization parameters verbatim).  A complete, valid Edgebreaker mesh that needs no predictor outside the path;
bench.py --workload c1 and the GPU tests decode it end to end.  Run from the repo root."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import drc_writer as W  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

if __name__ == "__main__":
    b = np.fromfile(os.path.join(HERE, "house_04.obj.drc"), dtype=np.uint8)
    o = O.decode(b)
    a = o.attrs[0]
    start = a.table_off - 5                      # pred, transform, compressed, scheme, max_bit_length
    end = a.payload_off + a.payload_len + 8      # + wrap bounds
    section = (bytes([1, 0xFF, 0, 0]) + W.varint(1) + bytes([0, 9, 3, 0]) + W.varint(0) + bytes([2])
               + bytes(b[start:end]) + bytes(b[end:end + 17]))
    out = bytes(b[:1158]) + section
    open(os.path.join(HERE, "house_04_positions_only.drc"), "wb").write(out)
    r = O.decode(np.frombuffer(out, dtype=np.uint8))
    assert r.status == 0 and np.array_equal(r.attrs[0].out, a.out)
    print(len(out), "bytes;", r.attrs[0].n_entries, "position entries,", r.n_points, "points")
