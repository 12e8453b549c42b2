"""Host-side robustness of the C ABI without a GPU: forged counts fail their own buffer before any arena is sized,
no C++ exception crosses the ABI, borrowed map arrays are validated.  """
import os
import struct

import numpy as np

import draco_sharp_b200 as D
from draco_sharp_b200 import _native as N

from common import cloud

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _patch_points(buf, n):
    b = buf.copy()
    b[11:15] = np.frombuffer(struct.pack("<i", n), dtype=np.uint8)  # "DRACO"(5) major minor type method flags(2) | i32 num_points at byte 11
    return b


def test_forged_point_count_fails_its_own_buffer_and_reserves_nothing():
    good = cloud(1000, seed=7, scheme=1)
    forged = _patch_points(good, 0x7FFFFFF0)
    truncated = forged[:60]
    bt = D.index_only([good, forged, truncated])
    assert bt.buffer_info(0).status == 0
    assert bt.buffer_info(1).status == -11          # DCB_ERR_ATTR: count the buffer cannot back
    assert bt.buffer_info(2).status != 0
    # the batch arena holds the good buffer only (12 bytes per point, 128-byte aligned)
    assert bt.out_bytes <= 1000 * 12 + 256, bt.out_bytes
    bt.free()


def test_truncated_buffer_with_a_plausible_count_reserves_nothing():
    good = cloud(5000, seed=9, scheme=1)
    bt = D.index_only([good[:200]])
    assert bt.buffer_info(0).status != 0
    assert bt.out_bytes == 0
    bt.free()


def test_forged_face_count_is_a_status_not_an_abort():
    house = np.fromfile(os.path.join(GOLD, "house_04.obj.drc"), dtype=np.uint8)
    # 39 bytes: header + connectivity header whose varints claim 2^27 faces (ADVICE: std::bad_alloc -> std::terminate)
    def varint(v):
        out = []
        while True:
            b = v & 0x7F
            v >>= 7
            out.append(b | (0x80 if v else 0))
            if not v:
                return out
    hdr = list(house[:11]) + [2]                       # DRACO 2.2 mesh/EB flags | traversal type 2 (valence)
    body = varint(1 << 26) + varint(1 << 27) + [0] + varint(1 << 27) + varint(0) + varint(0)
    forged = np.array(hdr + body + [0] * 8, dtype=np.uint8)
    bt = D.index_only([forged, house])
    bt.host_connectivity(0)
    bt.host_connectivity(1)
    bt.finish()
    assert bt.buffer_info(0).status in (-15, -103, -1)   # connectivity / oom / eof -- its own status, no crash
    assert bt.buffer_info(1).n_attrs == 3                # the neighbour is untouched
    bt.free()


def test_mesh_maps_must_come_in_triangles():
    house = np.fromfile(os.path.join(GOLD, "house_04.obj.drc"), dtype=np.uint8)
    bt = D.index_only([house])
    o = np.zeros(7, dtype=np.uint32)
    d = np.zeros(2, dtype=np.uint32)
    v = np.zeros(3, dtype=np.int32)
    rc = N.lib().dcb_set_mesh_maps(bt.h, 0, 0, o.ctypes.data, o.ctypes.data, 7, d.ctypes.data, 2, v.ctypes.data, 3)
    assert rc == -100  # DCB_ERR_ARG
    bt.free()
