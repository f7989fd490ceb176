"""Developer tool (no GPU needed): opcode histogram of the innermost backward-branch loop that contains a given opcode.
usage: python tools/sass_loop.py <object-or-so> <function-substring> [opcode=MUFU.EX2]"""
import collections
import re
import subprocess
import sys

obj, fn = sys.argv[1], sys.argv[2]
needle = sys.argv[3] if len(sys.argv) > 3 else "MUFU.EX2"
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
body = next(f for f in funcs if fn in f.split("\n", 1)[0])
ins = []
for line in body.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
hits = [i for i, (_, t) in enumerate(ins) if needle in t]
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?(?:\.ANY)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr:
        loops.append((addr[int(m.group(1), 16)], i))
best = None
for lo, hi in loops:
    n = sum(1 for h in hits if lo <= h <= hi)
    if n and (best is None or n > best[2] or (n == best[2] and hi - lo < best[1] - best[0])):
        best = (lo, hi, n)
lo, hi, n = best
cnt = collections.Counter()
for _, t in ins[lo:hi + 1]:
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    op = t.split()[0]
    cnt[".".join(op.split(".")[:2]) if op.startswith(("MUFU", "SYNCS", "LDTM", "STTM", "BAR")) else op.split(".")[0]] += 1
total = hi - lo + 1
print(f"{fn}: loop 0x{ins[lo][0]:x}..0x{ins[hi][0]:x}: {total} instructions, {n} x {needle}")
for op, c in cnt.most_common(40):
    print(f"  {op:16s} {c}")
