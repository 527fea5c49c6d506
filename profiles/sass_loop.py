"""Static SASS view of one kernel from an object file: opcode histogram of an address range.
    python profiles/sass_loop.py obj mangled_substr [lo hi]   (hex addresses; prints block list when omitted)
"""
import re, subprocess, sys, collections
obj, sub = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, lines = None, []
for l in out.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m and cur and sub in cur:
        lines.append((int(m.group(1), 16), m.group(2).strip()))
if len(sys.argv) >= 5:
    lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
    sel = [(a, s) for a, s in lines if lo <= a <= hi]
    h = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", s).split()[0] for a, s in sel)
    print(f"{len(sel)} instructions in [{lo:#x}, {hi:#x}]")
    for k, v in h.most_common():
        print(f"  {k:28s} {v}")
    if "-v" in sys.argv:
        for a, s in sel:
            print(f"{a:05x}  {s}")
else:
    # basic blocks delimited by branch targets / branches, with FP64 density
    targets = set()
    for a, s in lines:
        m = re.search(r"(BRA|BSSY\S*|CALL\S*)\s.*?(0x[0-9a-f]+)", s)
        if m and m.group(1).startswith("BRA"):
            targets.add(int(m.group(2), 16))
    start = lines[0][0]
    cnt = collections.Counter()
    for i, (a, s) in enumerate(lines):
        op = re.sub(r"^@!?U?P\d+\s+", "", s).split()[0]
        cnt[op.split(".")[0]] += 1
        last = i + 1 == len(lines)
        if op.startswith(("BRA", "EXIT", "RET")) or last or lines[i + 1][0] in targets:
            tot = sum(cnt.values())
            if tot >= 12:
                top = ", ".join(f"{k}:{v}" for k, v in cnt.most_common(8))
                print(f"[{start:05x}-{a:05x}] {tot:4d}  {top}   -> {s[:50]}")
            cnt = collections.Counter()
            if not last:
                start = lines[i + 1][0]
