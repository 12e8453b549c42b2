"""Python driver of the synthetic Draco v2.2 bitstream generator (synth/drc_synth.cpp).

Test and benchmark tooling: produces the BASELINE.json shapes.  Not part of the
decode product path.
"""
import ctypes as C
import os

import numpy as np

from . import _native as N


def make_spec(n_points, seed=0xD5AC0000, pos_bits=14, scheme=-1, normal_bits=0, colors=0, rho=(24, 25), color_step=3):
    return N.SynthSpec(seed=seed, n_points=n_points, pos_bits=pos_bits, rho_num=rho[0], rho_den=rho[1], scheme=scheme,
                       normal_bits=normal_bits, colors=colors, color_step=color_step, reserved=0)


def synth_cloud(spec, want_truth=True):
    """One cloud -> (bytes ndarray, truth dict)."""
    n = spec.n_points
    cap = 64 + n * 40 + 1 << 12
    cap = max(1 << 16, n * 48 + 65536)
    out = np.empty(cap, dtype=np.uint8)
    t = N.SynthTruth()
    pos_q = np.zeros((n, 3), dtype=np.int32)
    nrm = np.zeros((n, 2), dtype=np.int32)
    rgb = np.zeros((n, 3), dtype=np.uint8)
    if want_truth:
        t.pos_q = pos_q.ctypes.data
        t.nrm_st = nrm.ctypes.data
        t.rgb = rgb.ctypes.data
    sz = N.synth().synth_cloud(C.byref(spec), out.ctypes.data, cap, C.byref(t))
    if sz < 0:
        raise RuntimeError("synth_cloud: capacity")
    truth = {"pos_q": pos_q, "nrm_st": nrm, "rgb": rgb, "scheme": list(t.scheme), "sums": list(t.sums)}
    return out[:sz].copy(), truth


def synth_batch(spec, n_bufs, n_threads=None, arena=None):
    """A batch packed in one arena (each buffer on a 16-byte boundary).

    Returns (arena ndarray uint8, offs uint64[n], lens uint64[n], sums uint64[n,3], schemes int32[n,3]).
    `arena` may be a caller-provided uint8 ndarray (e.g. a view of pinned memory)."""
    if n_threads is None:
        n_threads = os.cpu_count() or 1
    offs = np.zeros(n_bufs, dtype=np.uint64)
    lens = np.zeros(n_bufs, dtype=np.uint64)
    sums = np.zeros((n_bufs, 3), dtype=np.uint64)
    schemes = np.zeros((n_bufs, 3), dtype=np.int32)
    if arena is None:
        per = 4 * spec.n_points * ((3 if spec.pos_bits else 0) + (2 if spec.normal_bits else 0) + (3 if spec.colors else 0)) // 2 + 4096
        arena = np.empty(max(1 << 16, per * n_bufs), dtype=np.uint8)
    while True:
        used = N.synth().synth_batch(C.byref(spec), n_bufs, n_threads, arena.ctypes.data, arena.nbytes,
                                     offs.ctypes.data, lens.ctypes.data, sums.ctypes.data, schemes.ctypes.data)
        if used >= 0:
            break
        arena = np.empty(int(-used) + 4096, dtype=np.uint8)
    return arena, offs, lens, sums, schemes, int(used)


def grid_topology(w, h):
    """Corner table + depth-first traversal maps of a triangulated w x h grid (BASELINE configs[3]).  The dict has
    the keys dcb_set_mesh_maps / the oracle take; `faces` is corner_to_vertex reshaped."""
    nf, nv = 2 * (w - 1) * (h - 1), w * h
    m = dict(opposite=np.zeros(3 * nf, dtype=np.uint32), corner_to_vertex=np.zeros(3 * nf, dtype=np.uint32),
             data_to_corner=np.zeros(nv, dtype=np.uint32), vertex_to_data=np.zeros(nv, dtype=np.int32))
    n = N.synth().synth_grid_topology(w, h, m["opposite"].ctypes.data, m["corner_to_vertex"].ctypes.data,
                                      m["data_to_corner"].ctypes.data, m["vertex_to_data"].ctypes.data)
    if n != nv:
        raise RuntimeError("grid traversal reached %d of %d vertices" % (n, nv))
    return m


def grid_mesh(w, h, maps, seed=0xD5AC4000, pos_bits=14, scheme=-1, want_q=False):
    """One Edgebreaker-mesh .drc buffer over grid_topology(w, h): positions, parallelogram + wrap, rANS.
    Returns (buffer, attr_section_off, checksum of the expected output floats, scheme used, pos_q or None)."""
    nv = w * h
    cap = 4096 + nv * 3 * 4
    out = np.empty(cap, dtype=np.uint8)
    aoff, sm, sch = C.c_uint64(0), C.c_uint64(0), C.c_int32(0)
    q = np.zeros(nv * 3, dtype=np.int32) if want_q else None
    sz = N.synth().synth_grid_mesh(w, h, seed, pos_bits, scheme, maps["opposite"].ctypes.data,
                                   maps["corner_to_vertex"].ctypes.data, maps["data_to_corner"].ctypes.data,
                                   maps["vertex_to_data"].ctypes.data, out.ctypes.data, cap, C.byref(aoff), C.byref(sm),
                                   q.ctypes.data if want_q else None, C.byref(sch))
    if sz < 0:
        raise RuntimeError("grid mesh needs %d bytes" % -sz)
    return out[:sz].copy(), int(aoff.value), int(sm.value), int(sch.value), q


def word_checksum(a):
    a = np.ascontiguousarray(a)
    return int(N.synth().synth_word_checksum(a.ctypes.data, a.nbytes))
