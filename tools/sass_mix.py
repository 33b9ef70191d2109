#!/usr/bin/env python
"""Static SASS opcode mix of one kernel of libdcb.so (which execution pipe the instructions go to)."""
import collections, re, subprocess, sys
pat = sys.argv[1] if len(sys.argv) > 1 else "specILi16ELi9ELi12ELi10"
out = subprocess.run(["cuobjdump", "-sass", "decombinator_b200/libdcb.so"], capture_output=True, text=True).stdout
on, ops = False, []
for line in out.splitlines():
    if "Function :" in line:
        on = pat in line
        continue
    if on:
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if m:
            ops.append(re.sub(r"^@!?U?P\d+\s+", "", m.group(1).strip()))
ALU = ("LOP3", "SHF", "IADD3", "ISETP", "SEL", "PRMT", "LEA", "VIADD", "PLOP3", "VIMNMX", "MOV", "CS2R", "FLO", "BREV", "POPC", "IABS", "SGXT", "BMSK")
FMA = ("IMAD",)
cnt = collections.Counter(o.split()[0].split(".")[0] for o in ops)
alu = sum(v for k, v in cnt.items() if k in ALU)
fma = sum(v for k, v in cnt.items() if k in FMA)
print("static instructions %d  ALU-pipe %d  FMA-pipe %d  LDS %d  control %d" % (len(ops), alu, fma, cnt["LDS"], cnt["BRA"] + cnt["BSSY"] + cnt["BSYNC"]))
print(" ".join("%s:%d" % kv for kv in cnt.most_common(14)))
if len(sys.argv) > 2:
    print("\n".join(ops))
