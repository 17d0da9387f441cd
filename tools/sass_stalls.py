"""Static single-warp cycle estimate of a kernel's SASS from the control words.

stall = bits [105:109) of each 128-bit instruction (B300_MICROARCH.md): cycles before the
warp may issue its next instruction.  Summing it over a straight-line region gives the
in-order issue time of ONE warp (ignoring scoreboard waits on variable-latency ops).

    python tools/sass_stalls.py <obj-or-so> <function-substring> [lo_addr hi_addr]
"""
import re, subprocess, sys, collections
obj, fn = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
txt = [b for b in blocks if fn in b.split("\n")[0]][0]
lines = txt.split("\n")
ins = []
i = 0
pat = re.compile(r"^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/")
pat2 = re.compile(r"^\s+/\* (0x[0-9a-f]{16}) \*/")
while i < len(lines):
    m = pat.match(lines[i])
    if m and i + 1 < len(lines):
        m2 = pat2.match(lines[i + 1])
        if m2:
            addr = int(m.group(1), 16); text = m.group(2).strip(); hi = int(m2.group(1), 16)
            stall = (hi >> 41) & 0xF
            yld = (hi >> 45) & 1
            wbar = (hi >> 46) & 7; rbar = (hi >> 49) & 7; wmask = (hi >> 52) & 0x3F
            ins.append((addr, text, stall, yld, wbar, rbar, wmask))
            i += 2
            continue
    i += 1
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi_ = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
sel = [x for x in ins if lo <= x[0] <= hi_]
tot = sum(x[2] for x in sel)
def op(t):
    t = t.split()
    if t[0].startswith("@"): t = t[1:]
    return t[0].split(".")[0]
cnt = collections.Counter(op(x[1]) for x in sel)
fp64 = sum(v for k, v in cnt.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print("instrs", len(sel), "fp64", fp64, "sum(stall) =", tot, "-> fp64 pipe util of one warp = %.1f %%" % (100.0 * 2 * fp64 / max(tot, 1)))
hist = collections.Counter(x[2] for x in sel)
print("stall histogram", sorted(hist.items()))
if "-v" in sys.argv:
    for x in sel:
        print("%05x st=%2d wb=%d rb=%d wm=%02x  %s" % (x[0], x[2], x[4], x[5], x[6], x[1]))
