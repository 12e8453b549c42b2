"""Diagnostic: host-buffer decode of the c2 batch with K pipeline slices, host-side timestamps (DCB_DEBUG_TIMING=1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench as B
import draco_sharp_b200 as D

K = int(sys.argv[1])
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
B.WORKLOADS["c2"] = (nb,) + B.WORKLOADS["c2"][1:]
pinned = {}
def pa(n):
    pinned["in"] = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    return pinned["in"].numpy()
arena, offs, lens, sums, schemes, used = B.make_workload("c2", 0, 2048, pa)
dec = D.DracoBatchDecoder([0] * K)
b = dec.index_arena(arena, offs, lens)
h_out = torch.empty(b.out_bytes, dtype=torch.uint8, pin_memory=True)
b.free()
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    b = dec.index_arena(arena, offs, lens)
    t1 = time.perf_counter()
    dec.decode(b, out_ptr=h_out.data_ptr())
    t2 = time.perf_counter()
    b.free()
    t3 = time.perf_counter()
    print("K=%d: index %.1f ms, decode %.1f ms, free %.1f ms" % (K, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), flush=True)
