"""Generates tests/golden/house_04_obj_vertices.npy and house_04_obj_texcoords.npy from the reference's own sample (run where /root/reference exists).

house_04.obj.drc is the only upstream-produced Draco artefact in the reference repo
(src/Draco.Examples/Samples/); house_04.obj is the mesh it was encoded from.  The `.drc` is copied verbatim
(8,196 bytes, a data asset) and the OBJ `v` / `vt` lines are stored as float64 so the parity tests can run on
machines that do not have /root/reference.
"""
import os
import shutil

import numpy as np

SRC = "/root/reference/src/Draco.Examples/Samples"
HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    shutil.copyfile(os.path.join(SRC, "house_04.obj.drc"), os.path.join(HERE, "house_04.obj.drc"))
    vs = [[float(t) for t in ln.split()[1:4]] for ln in open(os.path.join(SRC, "house_04.obj")) if ln.startswith("v ")]
    np.save(os.path.join(HERE, "house_04_obj_vertices.npy"), np.asarray(vs, dtype=np.float64))
    vts = [[float(t) for t in ln.split()[1:3]] for ln in open(os.path.join(SRC, "house_04.obj")) if ln.startswith("vt ")]
    np.save(os.path.join(HERE, "house_04_obj_texcoords.npy"), np.asarray(vts, dtype=np.float64))
    print(len(vs), "vertices,", len(vts), "texture coordinates")
