"""Print every metric of an .ncu-rep whose name matches a regex, per kernel (first launch of each name).
    python profiles/ncu_grep.py rep.ncu-rep 'shared|lsu|stall'
"""
import csv, subprocess, sys, io, re
rep, pat = sys.argv[1], re.compile(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
seen = set()
for r in data:
    name = r[hdr.index("Kernel Name")]
    if name in seen:
        continue
    seen.add(name)
    print("## " + name[:100])
    for i, h in enumerate(hdr):
        if pat.search(h):
            v = r[i]
            try:
                f = float(v.replace(",", ""))
                if f == 0:
                    continue
                v = f"{f:,.3f}".rstrip("0").rstrip(".")
            except ValueError:
                pass
            print(f"  {h:90s} {v:>18s} {units[i]}")
