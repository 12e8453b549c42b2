"""ASan/UBSan fuzz of the product's HOST code (container indexer, resumable walk, Edgebreaker host helper).

Not a pytest module (the sanitized build takes minutes): run it by hand,

    python tests/fuzz_host_sanitized.py [iterations] [seed]

It (1) builds libdracob200 a second time into a scratch directory with `-Xcompiler -fsanitize=address,undefined
-fno-sanitize-recover=undefined`, (2) re-executes itself in a child with libasan preloaded and DCB_LIB pointing at that
build, and (3) in the child pushes mutated buffers -- the reference's sample mesh, synthetic point clouds of every
symbol scheme, crafted tex-coord / parallelogram / constrained-multi-parallelogram / geometric-normal attribute sections -- through dcb_index -> dcb_host_connectivity ->
dcb_index_finish -> every getter, with no device (index-only batches).  Any sanitizer report aborts the child; the
parent prints its tail and exits non-zero.  The last run's summary is committed as profiles/r2_fuzz_host_asan.txt.
"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build_sanitized(out_dir):
    csrc = os.path.join(ROOT, "draco_sharp_b200", "csrc")
    cus = [os.path.join(csrc, f) for f in sorted(os.listdir(csrc)) if f.endswith(".cu")]
    lib = os.path.join(out_dir, "libdracob200_asan.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-g", "-std=c++17", "-Xcompiler",
             "-fPIC,-fsanitize=address,-fsanitize=undefined,-fno-sanitize-recover=undefined,-fno-omit-frame-pointer"]
    objs = []
    procs = []
    for c in cus:
        o = os.path.join(out_dir, os.path.basename(c)[:-3] + ".o")
        objs.append(o)
        procs.append(subprocess.Popen([nvcc] + flags + ["-c", "-o", o, c]))
    for p in procs:
        if p.wait() != 0:
            raise SystemExit("sanitized build failed")
    subprocess.check_call([nvcc] + flags + ["-shared", "-o", lib] + objs + ["-Xlinker", "-lasan", "-Xlinker", "-lubsan"])
    return lib


def child(iterations, seed):
    import numpy as np
    import drc_writer as W
    import draco_sharp_b200 as D
    from draco_sharp_b200 import _native as N
    from draco_sharp_b200 import synth_gen as G

    rng = np.random.default_rng(seed)
    house = np.fromfile(os.path.join(ROOT, "tests", "golden", "house_04.obj.drc"), dtype=np.uint8)
    seeds = [house]
    for scheme in (0, 1):
        spec = G.make_spec(300, seed=7 + scheme, pos_bits=14, scheme=scheme, normal_bits=10, colors=1)
        arena, offs, lens, _, _, _ = G.synth_batch(spec, 2, n_threads=1)
        for k in range(2):
            seeds.append(np.array(arena[int(offs[k]): int(offs[k]) + int(lens[k])], dtype=np.uint8))
    # a crafted mesh attribute section with a tex-coord predictor behind an opaque connectivity blob
    flags = rng.integers(0, 2, size=40)
    sec = bytearray([2, 0xFF, 0, 0, 0, 1, 0]) + W.varint(1) + bytes([0, 9, 3, 0]) + W.varint(0) + bytes([2])
    sec += W.varint(1) + bytes([3, 9, 2, 0]) + W.varint(1) + bytes([2])
    sec += W.portable_int(rng.integers(-5, 6, size=36), 3, 1, 1, "raw", W.wrap_data(0, 255)) + W.quant_params([0, 0, 0], 1.0, 8)
    sec += W.portable_int(rng.integers(-5, 6, size=24), 2, 5, 1, "tagged", W.tex_coords_data(flags, 0, 255)) + W.quant_params([0, 0], 1.0, 8)
    head = b"DRACO" + bytes([2, 2, 1, 1, 0, 0, 2]) + b"\xAA" * 37
    crafted = np.frombuffer(head + bytes(sec), dtype=np.uint8)
    # ... and one with the other 8f-3 predictors: constrained multi-parallelogram positions (four crease-flag blocks) and
    # geometric normals (flip bits behind the transform data)
    sec2 = bytearray([1, 0xFF, 0, 0]) + W.varint(2) + bytes([0, 9, 3, 0]) + W.varint(0) + bytes([1, 9, 3, 0]) + W.varint(1) + bytes([2, 3])
    crease = [list(rng.integers(0, 2, size=k)) for k in (5, 8, 0, 4)]
    sec2 += W.portable_int(rng.integers(-5, 6, size=36), 3, 4, 1, "raw", W.cmp_data(crease, 0, 255))
    sec2 += W.portable_int(rng.integers(0, 60, size=24), 2, 6, 3, "tagged", W.geometric_normal_data(rng.integers(0, 2, size=12), 8), zig=False)
    sec2 += W.quant_params([0, 0, 0], 1.0, 8) + bytes([8])
    crafted2 = np.frombuffer(head + bytes(sec2), dtype=np.uint8)
    stats = {"decodes": 0, "ok": 0, "failed": 0}
    for it in range(iterations):
        base = seeds[it % len(seeds)] if it % 7 else (crafted if it % 14 else crafted2)
        b = base.copy()
        kind = int(rng.integers(0, 4))
        if kind == 0:      # byte flips
            for _ in range(int(rng.integers(1, 8))):
                b[int(rng.integers(0, b.size))] = int(rng.integers(0, 256))
        elif kind == 1:    # truncation
            b = b[: int(rng.integers(0, b.size + 1))].copy()
        elif kind == 2:    # a varint / count blown up
            i = int(rng.integers(0, max(1, b.size - 5)))
            b[i: i + 5] = np.array([0xFF, 0xFF, 0xFF, 0xFF, 0x0F], dtype=np.uint8)[: b.size - i]
        else:              # splice two seeds
            o = seeds[int(rng.integers(0, len(seeds)))]
            cut = int(rng.integers(0, min(b.size, o.size)))
            b = np.concatenate([b[:cut], o[cut:]])
        bt = D.index_only([b, house])
        for k in range(2):
            bi = bt.buffer_info(k)
            if bi.status == 0 and bi.needs_connectivity:
                try:  # argument errors (an attribute section beyond a truncated buffer ...) come back as exceptions
                    if (base is crafted or base is crafted2) and k == 0:
                        bt.set_attr_section(0, len(head), 12)
                        n = 12
                        idm = np.arange(3 * n, dtype=np.uint32)
                        for d in range(2 if base is crafted else 1):
                            bt.set_mesh_maps(0, d, idm, idm % n, np.arange(n, dtype=np.uint32) * 3, np.arange(n, dtype=np.int32))
                    else:
                        bt.host_connectivity(k)
                except N.DracoError:
                    stats["arg_errors"] = stats.get("arg_errors", 0) + 1
        bt.finish()
        for k in range(2):
            bi = bt.buffer_info(k)
            stats["decodes"] += 1
            stats["ok" if bi.status == 0 else "failed"] += 1
            if bi.status == 0:
                for a in range(bi.n_attrs):
                    bt.attr_info(k, a)
                if bi.geometry_type == 1:
                    bt.faces(k)
        assert bt.buffer_info(1).status == 0, "a malformed neighbour poisoned the intact buffer"
        bt.free()
    print("fuzz_host_sanitized: %d iterations, seed %d: %s, no sanitizer report" % (iterations, seed, stats))


def main():
    iterations = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    if os.environ.get("DCB_FUZZ_CHILD"):
        return child(iterations, seed)
    with tempfile.TemporaryDirectory() as tmp:
        lib = build_sanitized(tmp)
        asan = subprocess.check_output(["gcc", "-print-file-name=libasan.so"], text=True).strip()
        env = dict(os.environ, DCB_FUZZ_CHILD="1", DCB_LIB=lib, LD_PRELOAD=asan,
                   ASAN_OPTIONS="detect_leaks=0:abort_on_error=1:protect_shadow_gap=0", UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
        r = subprocess.run([sys.executable, os.path.abspath(__file__), str(iterations), str(seed)], env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        print(r.stdout[-4000:])
        raise SystemExit(r.returncode)


if __name__ == "__main__":
    main()
