"""Split a kernel's SASS into straight-line regions (at branches / BSSY / BSYNC) and count
instruction classes per region: how many fp64-pipe instructions, shared / local memory
accesses, MUFU, register moves.  Used for profiles/*_sass_evidence.txt.

    python tools/sass_segments.py <obj-or-so> <function-substring> [min_instructions]
"""
import re
import subprocess
import sys

obj, key = sys.argv[1], sys.argv[2]
min_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
p = False
ins = []
for l in out.splitlines():
    if "Function :" in l:
        p = key in l
    if p:
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
print(len(ins), "instructions")
seg_start, cnt = 0, {}


def flush(i, why):
    global seg_start, cnt
    tot = sum(cnt.values())
    if tot > min_n:
        print("%05x-%05x n=%4d  fp64=%4d lds=%3d ldl/stl=%d mufu=%d mov=%d | %s" % (
            ins[seg_start][0], ins[i][0], tot, cnt.get("D", 0), cnt.get("LDS", 0), cnt.get("L", 0),
            cnt.get("MUFU", 0), cnt.get("MOV", 0), why))
    seg_start, cnt = i + 1, {}


for i, (a, t) in enumerate(ins):
    op = t.split()[0] if not t.startswith("@") else t.split()[1]
    k = ("D" if op.startswith(("DFMA", "DMUL", "DADD")) else "LDS" if op.startswith("LDS")
         else "L" if op.startswith(("LDL", "STL")) else "MUFU" if op.startswith("MUFU")
         else "MOV" if op.startswith(("MOV", "IMAD.MOV")) else "o")
    cnt[k] = cnt.get(k, 0) + 1
    if op.startswith(("BRA", "EXIT", "RET", "BSYNC", "BSSY")):
        flush(i, t[:60])
