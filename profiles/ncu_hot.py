"""Top stall-sample instructions of one kernel (first instance) from an .ncu-rep source page.
    python profiles/ncu_hot.py rep kernel_substr [N]
"""
import csv, io, subprocess, sys
rep, sub = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
out, on = [], False
for r in rows:
    if r and r[0] == "Kernel Name":
        if on:
            break
        on = sub in r[1].replace("(bool)", "")
        continue
    if on and len(r) > 6 and r[0].startswith("0x"):
        out.append((int(r[0], 16), r[1].strip(), int(r[2]), int(r[3]), int(r[5])))
base = out[0][0]
tot = sum(o[2] for o in out)
print(f"total samples {tot}, instructions executed {sum(o[4] for o in out)}")
for a, s, smp, ni, ex in sorted(out, key=lambda o: -o[2])[:N]:
    print(f"{a-base:05x} {smp:6d} ({100*smp/tot:4.1f}%) notissued {ni:6d} exec {ex:9d}  {s[:90]}")
