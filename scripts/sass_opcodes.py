"""Per-kernel counts of the Blackwell-specific SASS opcodes in the shipped library (evidence that the hot path is
tcgen05 / TMEM / TMA, profiles/*_sass_opcodes.txt).

    python scripts/sass_opcodes.py [path/to/libxgating.so] > profiles/r2_sass_opcodes.txt
"""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "controllable_xgating_b200/libxgating.so"
OPS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMAPF", "UBLKCP", "LDTM", "UTCATOMSWS", "SYNCS", "MUFU.RCP", "MUFU.EX2"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
name, rows = None, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        rows[name] = dict(instr=0, **{o: 0 for o in OPS})
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if name and m:
        rows[name]["instr"] += 1
        for o in OPS:
            if m.group(1).startswith(o):
                rows[name][o] += 1
print("SASS opcode counts per kernel of %s (cuobjdump -sass; sm_100a)" % lib)
print("%-52s %7s " % ("kernel", "instr") + " ".join("%8s" % o[:8] for o in OPS))
for k, r in sorted(rows.items(), key=lambda kv: -kv[1]["UTCHMMA"] * 100000 - kv[1]["instr"]):
    if r["UTCHMMA"] or r["UTMALDG"] or r["UBLKCP"] or r["LDTM"]:
        print("%-52s %7d " % (k[:52], r["instr"]) + " ".join("%8d" % r[o] for o in OPS))
tot = {o: sum(r[o] for r in rows.values()) for o in OPS}
print("%-52s %7d " % ("all %d kernels" % len(rows), sum(r["instr"] for r in rows.values())) + " ".join("%8d" % tot[o] for o in OPS))
