"""Per-instruction stall breakdown of one kernel from an .ncu-rep source page: top instructions for one stall
reason, and shared-memory wavefront excess per LDS.   python profiles/ncu_stalls.py rep kernel_substr stall_long_sb [N]"""
import csv, io, subprocess, sys
rep, sub, col = sys.argv[1], sys.argv[2], sys.argv[3]
N = int(sys.argv[4]) if len(sys.argv) > 4 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, out, on = None, [], False
for r in rows:
    if r and r[0] == "Kernel Name":
        if on:
            break
        on = sub in r[1].replace("(bool)", "")
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if on and hdr and r and r[0].startswith("0x"):
        out.append(r)
ic = hdr.index(col)
iw, ii = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
base = int(out[0][0], 16)
tot = sum(int(r[ic] or 0) for r in out)
print(f"{col}: total {tot}")
for r in sorted(out, key=lambda r: -int(r[ic] or 0))[:N]:
    print(f"{int(r[0],16)-base:05x} {int(r[ic] or 0):6d}  exec {r[hdr.index('Instructions Executed')]:>9s}  wf {r[iw]:>9s} ideal {r[ii]:>9s}  {r[1].strip()[:80]}")
