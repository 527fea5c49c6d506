"""Executed instructions and stall samples of one kernel grouped by per-instruction execution count
(= code region: hot loop, per-frame overhead, map build, border path, producer ...).
    python profiles/ncu_regions.py rep [kernel_index] [-v COUNT ...]   (-v: list the instructions of those regions)"""
import csv, sys, collections, subprocess, io
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 1
verbose = [int(x) for x in sys.argv[sys.argv.index("-v") + 1:]] if "-v" in sys.argv else []
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
on, seen, base = False, 0, None
g = collections.OrderedDict()
lines = []
for r in rows:
    if r and r[0] == "Kernel Name":
        seen += 1
        on = seen == which
        if on:
            print("#", r[1][:110])
        continue
    if on and len(r) > 6 and r[0].startswith("0x"):
        a = int(r[0], 16)
        base = a if base is None else base
        ex, samp = int(r[5]), int(r[2])
        d = g.setdefault(ex, [0, 0, 0]); d[0] += 1; d[1] += ex; d[2] += samp
        lines.append((a - base, samp, ex, r[1].strip()))
tot = sum(d[1] for d in g.values()); ts = sum(d[2] for d in g.values())
print(f"executed warp instructions {tot}, stall samples {ts}")
for k, d in sorted(g.items(), key=lambda kv: -kv[1][2])[:12]:
    print(f"  executed {k:9d} times: {d[0]:4d} instructions, {100 * d[1] / tot:5.1f} % of executed, {100 * d[2] / ts:5.1f} % of samples")
for off, samp, ex, txt in lines:
    if ex in verbose:
        print(f"{off:05x} {samp:5d} {ex:8d}  {txt[:100]}")
