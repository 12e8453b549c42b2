"""Diagnostic: one grid mesh of a given side / scheme through the GPU path, stage timings (not a bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import draco_sharp_b200 as D
from draco_sharp_b200 import synth_gen as G

side, scheme, count = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 1
topo = G.grid_topology(side, side)
meshes = [G.grid_mesh(side, side, topo, seed=5 + k, scheme=scheme) for k in range(count)]
dec = D.DracoBatchDecoder([0])
b = dec.index([m[0] for m in meshes])
for k, m in enumerate(meshes):
    b.set_attr_section(k, m[1], side * side)
    b.set_mesh_maps(k, 0, topo["opposite"], topo["corner_to_vertex"], topo["data_to_corner"], topo["vertex_to_data"])
b.finish()
for it in range(2):
    t = time.time()
    out, _ = dec.decode(b)
    dt = time.time() - t
    st = dec.stats()
    ai = b.attr_info(0, 0)
    ok = G.word_checksum(out[ai.out_off: ai.out_off + ai.out_bytes]) == meshes[0][2]
    print("side %d scheme %d x%d: %.3f s wall, kernels %.1f ms (raw %.1f tag %.1f par %.1f para %.1f) launches %d ok %s status %d"
          % (side, meshes[0][3], count, dt, st.ms_total, st.ms_raw, st.ms_tag, st.ms_par, st.ms_para, st.n_launches, ok, b.status(0)), flush=True)
