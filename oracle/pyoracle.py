"""ctypes binding of the CPU ORACLE (oracle/liboracle.so).  TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package (draco_sharp_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")


class OrcAttr(C.Structure):
    _fields_ = [
        ("att_type", C.c_int32), ("data_type", C.c_int32), ("nc", C.c_int32), ("normalized", C.c_int32),
        ("unique_id", C.c_uint32), ("seq_type", C.c_int32), ("decoder_id", C.c_int32), ("pred_method", C.c_int32),
        ("transform", C.c_int32), ("compressed", C.c_int32), ("scheme", C.c_int32), ("nc_portable", C.c_int32),
        ("n_entries", C.c_uint32), ("max_bit_length", C.c_int32), ("precision", C.c_int32),
        ("table_symbols", C.c_uint32), ("table_off", C.c_uint64), ("payload_off", C.c_uint64),
        ("payload_len", C.c_uint64), ("bits_off", C.c_uint64), ("bits_len", C.c_uint64),
        ("final_state", C.c_uint32), ("leftover", C.c_uint64), ("xf_a", C.c_int32), ("xf_b", C.c_int32),
        ("qmin", C.c_float * 4), ("qrange", C.c_float), ("qbits", C.c_int32),
        ("symbols", C.POINTER(C.c_uint32)), ("corr", C.POINTER(C.c_int32)), ("qints", C.POINTER(C.c_int32)),
        ("out", C.POINTER(C.c_uint8)), ("out_bytes", C.c_uint64),
    ]


class OrcMaps(C.Structure):
    _fields_ = [
        ("opposite", C.POINTER(C.c_uint32)), ("corner_to_vertex", C.POINTER(C.c_uint32)), ("n_corners", C.c_uint64),
        ("data_to_corner", C.POINTER(C.c_uint32)), ("n_entries", C.c_uint64),
        ("vertex_to_data", C.POINTER(C.c_int32)), ("n_vertices", C.c_uint64),
    ]


class OrcResult(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("ver_major", C.c_int32), ("ver_minor", C.c_int32), ("geom_type", C.c_int32),
        ("method", C.c_int32), ("flags", C.c_int32), ("n_points", C.c_uint32), ("n_faces", C.c_uint32),
        ("n_decoders", C.c_int32), ("n_attrs", C.c_int32), ("attrs", C.POINTER(OrcAttr)),
        ("faces", C.POINTER(C.c_uint32)), ("attr_section_off", C.c_uint64), ("end_off", C.c_uint64),
        ("n_maps", C.c_int32), ("maps", C.POINTER(OrcMaps)),
    ]


_lib = None


def build():
    subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        l = C.CDLL(LIB_PATH)
        l.orc_decode.restype = C.c_int
        l.orc_decode.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(OrcMaps), C.c_int, C.POINTER(C.POINTER(OrcResult))]
        l.orc_decode_ex.restype = C.c_int
        l.orc_decode_ex.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(OrcMaps), C.c_int, C.c_uint64, C.c_uint32,
                                    C.POINTER(C.POINTER(OrcResult))]
        l.orc_free.restype = None
        l.orc_free.argtypes = [C.POINTER(OrcResult)]
        l.orc_decode_bench.restype = C.c_int
        l.orc_decode_bench.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        l.orc_varint.restype = C.c_int
        l.orc_varint.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        l.orc_read_bits_lsb.restype = C.c_uint32
        l.orc_read_bits_lsb.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_int, C.POINTER(C.c_int)]
        l.orc_zigzag.restype = C.c_int32
        l.orc_zigzag.argtypes = [C.c_uint32]
        l.orc_reinterpret_i2u.restype = C.c_uint32
        l.orc_reinterpret_i2u.argtypes = [C.c_int32]
        l.orc_int_sqrt.restype = C.c_uint64
        l.orc_int_sqrt.argtypes = [C.c_uint64]
        l.orc_rans_precision.restype = C.c_int
        l.orc_rans_precision.argtypes = [C.c_int]
        l.orc_decode_symbols.restype = C.c_int
        l.orc_decode_symbols.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_uint32, C.c_uint32,
                                         C.c_void_p, C.POINTER(OrcAttr)]
        l.orc_fnv1a.restype = C.c_uint64
        l.orc_fnv1a.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        _lib = l
    return _lib


class Attr:
    """Python copy of one decoded attribute (arrays are owned numpy copies)."""

    def __init__(self, a: OrcAttr):
        for name, _ in OrcAttr._fields_:
            if name in ("symbols", "corr", "qints", "out"):
                continue
            v = getattr(a, name)
            setattr(self, name, list(v) if name == "qmin" else v)
        nv = a.n_entries * a.nc_portable
        self.symbols = np.ctypeslib.as_array(a.symbols, shape=(nv,)).copy() if a.symbols and nv else np.zeros(0, np.uint32)
        self.corr = np.ctypeslib.as_array(a.corr, shape=(nv,)).copy() if a.corr and nv else np.zeros(0, np.int32)
        self.qints = np.ctypeslib.as_array(a.qints, shape=(nv,)).copy() if a.qints and nv else np.zeros(0, np.int32)
        self.out = np.ctypeslib.as_array(a.out, shape=(a.out_bytes,)).copy() if a.out and a.out_bytes else np.zeros(0, np.uint8)


class Result:
    def __init__(self, r: OrcResult):
        for name in ("status", "ver_major", "ver_minor", "geom_type", "method", "flags", "n_points", "n_faces",
                     "n_decoders", "n_attrs", "attr_section_off", "end_off", "n_maps"):
            setattr(self, name, getattr(r, name))
        self.attrs = [Attr(r.attrs[i]) for i in range(r.n_attrs)] if r.attrs else []
        self.faces = (np.ctypeslib.as_array(r.faces, shape=(r.n_faces * 3,)).copy().reshape(-1, 3)
                      if r.faces and r.n_faces else np.zeros((0, 3), np.uint32))
        self.maps = []
        for i in range(r.n_maps):
            m = r.maps[i]
            self.maps.append({
                "opposite": np.ctypeslib.as_array(m.opposite, shape=(m.n_corners,)).copy() if m.n_corners else np.zeros(0, np.uint32),
                "corner_to_vertex": np.ctypeslib.as_array(m.corner_to_vertex, shape=(m.n_corners,)).copy() if m.n_corners else np.zeros(0, np.uint32),
                "data_to_corner": np.ctypeslib.as_array(m.data_to_corner, shape=(m.n_entries,)).copy() if m.n_entries else np.zeros(0, np.uint32),
                "vertex_to_data": np.ctypeslib.as_array(m.vertex_to_data, shape=(m.n_vertices,)).copy() if m.n_vertices else np.zeros(0, np.int32),
            })


def decode(buf, maps=None, attr_section_off=0, n_points=0) -> Result:
    """Decode one .drc buffer on the CPU.  maps: optional list of dicts (see Result.maps)."""
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else np.ascontiguousarray(buf, dtype=np.uint8)
    res = C.POINTER(OrcResult)()
    cm = None
    keep = []
    n_maps = 0
    if maps:
        n_maps = len(maps)
        cm = (OrcMaps * n_maps)()
        for i, m in enumerate(maps):
            o = np.ascontiguousarray(m["opposite"], dtype=np.uint32)
            c = np.ascontiguousarray(m["corner_to_vertex"], dtype=np.uint32)
            d = np.ascontiguousarray(m["data_to_corner"], dtype=np.uint32)
            v = np.ascontiguousarray(m["vertex_to_data"], dtype=np.int32)
            keep += [o, c, d, v]
            cm[i].opposite = o.ctypes.data_as(C.POINTER(C.c_uint32))
            cm[i].corner_to_vertex = c.ctypes.data_as(C.POINTER(C.c_uint32))
            cm[i].n_corners = o.size
            cm[i].data_to_corner = d.ctypes.data_as(C.POINTER(C.c_uint32))
            cm[i].n_entries = d.size
            cm[i].vertex_to_data = v.ctypes.data_as(C.POINTER(C.c_int32))
            cm[i].n_vertices = v.size
    lib().orc_decode_ex(a.ctypes.data if a.size else None, a.size, cm, n_maps, attr_section_off, n_points, C.byref(res))
    out = Result(res.contents)
    lib().orc_free(res)
    return out


def decode_bench(arena: np.ndarray, offs, lens):
    """Decode-and-discard over many buffers (GIL released inside the C call): returns (points, out_bytes)."""
    pts = C.c_uint64(0)
    ob = C.c_uint64(0)
    ck = C.c_uint64(0)
    base = arena.ctypes.data
    l = lib()
    for o, n in zip(offs, lens):
        l.orc_decode_bench(base + int(o), int(n), C.byref(pts), C.byref(ob), C.byref(ck))
    return pts.value, ob.value
