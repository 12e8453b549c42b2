#!/usr/bin/env python3
"""In-order single-warp issue model for a SASS loop body (one basic block), to compare instruction schedules of the
rANS main loops WITHOUT a GPU.  Numbers from scripts/ubench/issue_rate.cu (profiles/r2_ubench_issue_rate.txt):
an integer instruction occupies its pipe (alu or fma) 2 cycles per warp instruction, dependent ALU link 4 cycles
(5 across pipes), dependent LDS ~23 cycles conflict-free (+ wavefronts), one warp per sub-partition so nothing else
issues.  Usage:  sass_sim.py file.sass first_line last_line [iters]   (lines as printed by cuobjdump -sass, 1-based)
The model is calibrated on the round-2 lean loop (measured 177 cycles per symbol)."""
import re
import sys

LAT_ALU, LAT_X, LAT_LDS, LAT_LDG = 4, 5, 27, 400
LDS_ISSUE = 4  # 17 lanes on 32 banks: ~2.4 wavefronts

FMA_OPS = ("IMAD", "FMUL", "FADD", "FFMA")
LSU_OPS = ("LDS", "STS", "LDG", "STG", "LDGSTS", "ATOMS", "LDGDEPBAR", "DEPBAR")


def parse(line):
    s = re.sub(r"/\*.*?\*/", "", line).strip().rstrip(";").strip()
    if not s:
        return None
    pred = None
    m = re.match(r"@(!?)(U?P\d)\s+(.*)", s)
    if m:
        pred = m.group(2)
        s = m.group(3)
    op, _, rest = s.partition(" ")
    ops = [o.strip() for o in rest.split(",")] if rest.strip() else []
    base = op.split(".")[0]

    def regs(o, width=1):
        out = []
        for r in re.findall(r"\b(U?R\d+|U?P\d)\b", o):
            out.append(r)
            if width > 1 and r[0] == "R":
                n = int(r[1:])
                out += ["R%d" % (n + k) for k in range(1, width)]
        return out

    w = 4 if ".128" in op else (2 if (".64" in op or ".WIDE" in op) else 1)
    dst, src = [], []
    if base in ("STG", "STS"):
        src += regs(ops[0], 2 if base == "STG" else 1)
        src += regs(ops[1], w) if len(ops) > 1 else []
    elif base in ("BRA", "BAR", "DEPBAR", "LDGDEPBAR", "NOP", "WARPSYNC", "BSSY", "BSYNC", "EXIT"):
        for o in ops:
            src += regs(o)
    elif base == "LDGSTS":
        for o in ops:
            src += regs(o, 2 if "64" in o else 1)
    elif base == "ISETP" or base == "FSETP" or base == "PLOP3":
        dst += regs(ops[0]) + regs(ops[1])
        for o in ops[2:]:
            src += regs(o)
    elif not ops:
        pass
    else:
        dst += regs(ops[0], w if base in ("LDS", "LDG", "IMAD") else 1)
        k = 1
        if base in ("IADD3", "LEA", "VIADD") or op.startswith("IADD3"):
            while k < len(ops) and re.fullmatch(r"!?U?P[\dT]", ops[k]):
                if ops[k] not in ("PT", "!PT"):
                    dst.append(ops[k].lstrip("!"))
                k += 1
        for o in ops[k:]:
            src += regs(o, 2 if ".64" in o else 1)
    if pred:
        src.append(pred)  # (a predicated write keeps the old value without reading it: WAW order only)
    src = [r for r in src if r not in ("PT", "UPT", "RZ", "URZ")]
    pipe = "lsu" if base in LSU_OPS else ("fma" if base in FMA_OPS else ("none" if base in ("BRA", "NOP") else "alu"))
    return dict(op=op, base=base, dst=dst, src=src, pipe=pipe, text=s)


def simulate(ins, iters=4, verbose=False):
    ready = {}
    rpipe = {}
    pipe_free = {"alu": 0, "fma": 0, "lsu": 0, "none": 0}
    t = 0
    marks = []
    stall = [0.0] * len(ins)
    for it in range(iters):
        for k, i in enumerate(ins):
            t0 = t + 1
            need = t0
            for r in i["src"]:
                if r in ready:
                    lat_ready = ready[r]
                    if rpipe.get(r) not in (None, i["pipe"], "lsu") and i["pipe"] != "lsu":
                        lat_ready += LAT_X - LAT_ALU
                    need = max(need, lat_ready)
            need = max(need, pipe_free[i["pipe"]])
            if it == iters - 1:
                stall[k] = need - t0
            t = need
            if i["base"] == "LDS":
                lat, occ = LAT_LDS, LDS_ISSUE
            elif i["base"] in ("LDG",):
                lat, occ = LAT_LDG, 4
            elif i["base"] in ("STG", "STS", "LDGSTS"):
                lat, occ = 4, 4
            elif i["pipe"] == "none":
                lat, occ = 1, 1
            else:
                lat, occ = LAT_ALU, 2
            pipe_free[i["pipe"]] = t + occ
            for d in i["dst"]:
                ready[d] = t + lat
                rpipe[d] = i["pipe"]
        marks.append(t)
    per_iter = marks[-1] - marks[-2]
    if verbose:
        for k, i in enumerate(ins):
            if stall[k] >= 3:
                print("%5d  +%3d  %s" % (k + 1, stall[k], i["text"]))
    return per_iter


def main():
    path, a, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    syms = int(sys.argv[4]) if len(sys.argv) > 4 else 12
    lines = open(path).read().splitlines()[a - 1:b]
    ins = [p for p in (parse(l) for l in lines) if p]
    cyc = simulate(ins, verbose="-v" in sys.argv)
    n_lds = sum(1 for i in ins if i["base"] == "LDS")
    print("instructions %d  LDS %d  cycles/iteration %d  cycles/symbol %.1f  instr/symbol %.1f" %
          (len(ins), n_lds, cyc, cyc / syms, len(ins) / syms))


if __name__ == "__main__":
    main()
