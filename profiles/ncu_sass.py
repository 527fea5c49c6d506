"""Executed-instruction listing of one kernel from an .ncu-rep source page.
    python profiles/ncu_sass.py rep kernel_substring [--hist]
"""
import csv, io, re, subprocess, sys, collections
rep, kname = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
out, started = [], False
for r in rows:
    if r and r[0] == "Kernel Name":
        if started:
            break
        started = kname in r[1].replace("(bool)", "")
        continue
    if started and len(r) > 6 and r[0].startswith("0x"):
        out.append((r[1].strip(), int(r[5]), int(r[4])))
base = max(c for _, c, _ in out[:12])
tot = sum(c for _, c, _ in out)
print(f"# {kname}: {tot} warp instructions, {tot/base:.1f} per warp")
if "--hist" in sys.argv:
    ops, smp = collections.Counter(), collections.Counter()
    for s, c, sm in out:
        op = re.sub(r"^@!?U?P[0-9T]+\s+", "", s).split()[0]
        ops[op] += c / base
        smp[op] += sm
    for op, c in ops.most_common(45):
        print(f"{op:28s} {c:8.2f} per warp   stall samples {smp[op]}")
else:
    for s, c, sm in out:
        if c:
            print(f"{c/base:7.2f} {sm:6d}  {s}")
