#!/usr/bin/env python3
"""List the backward-branch loops of one kernel in an object file and run the issue model (sass_sim.py) on each.
Usage: sass_loops.py obj.o <mangled-name-substring> [min_instr] [syms]"""
import re, subprocess, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sass_sim as S
S.LAT_LDS, S.LDS_ISSUE = int(os.environ.get('LAT_LDS', 34)), int(os.environ.get('LDS_ISSUE', 6))
obj, pat = sys.argv[1], sys.argv[2]
min_ins = int(sys.argv[3]) if len(sys.argv) > 3 else 300
syms = int(sys.argv[4]) if len(sys.argv) > 4 else 12
txt = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout.splitlines()
start = [i for i, l in enumerate(txt) if "Function :" in l]
for k, i in enumerate(start):
    if pat not in txt[i]:
        continue
    end = start[k + 1] if k + 1 < len(start) else len(txt)
    body = [l for l in txt[i:end] if re.search(r"/\*[0-9a-f]{4,5}\*/", l)]
    addr = {}
    for n, l in enumerate(body):
        addr[int(re.search(r"/\*([0-9a-f]{4,5})\*/", l).group(1), 16)] = n
    out = os.environ.get("SASS_OUT")
    if out:
        open(out, "w").write("\n".join(body))
    print(txt[i].strip()[:160], "instructions", len(body))
    for n, l in enumerate(body):
        m = re.search(r"BRA\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", l)
        if m and " BRA" in l:
            t = int(m.group(1), 16)
            if t in addr and addr[t] < n and n - addr[t] >= min_ins:
                ins = [p for p in (S.parse(x) for x in body[addr[t]:n + 1]) if p]
                cyc = S.simulate(ins)
                print("  loop lines %d..%d: %d instr, %d LDS, model %.1f cycles/symbol, %.1f instr/symbol" %
                      (addr[t] + 1, n + 1, len(ins), sum(1 for q in ins if q["base"] == "LDS"), cyc / syms, len(ins) / syms))
