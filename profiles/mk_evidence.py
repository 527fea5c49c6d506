"""Turn the raw ncu outputs of a round (gpurun_out/) into the committed evidence under profiles/:
    python profiles/mk_evidence.py r1
  <r>_launches.csv          copy of the ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu`
  <r>_launches_summary.txt  per-kernel launch counts / mean time / share of the timed GPU work
  <r>_ncu_full_summary.txt  the few --set full numbers DESIGN.md cites, per kernel
  traffic.json              dram bytes per launch (read + write) of the rectification kernels
"""
import csv, io, json, os, shutil, subprocess, sys, collections
r = sys.argv[1] if len(sys.argv) > 1 else "r1"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
# ---- launch list
src = os.path.join(go, f"{r}_launches.csv")
if os.path.exists(src):
    lines = [l for l in open(src) if l.startswith('"')]
    open(os.path.join(pr, f"{r}_launches.csv"), "w").writelines(lines)
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = collections.OrderedDict()
    for x in rows:
        if x["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        v_us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
        a = agg.setdefault(x["Kernel Name"], [0, 0.0])
        a[0] += 1; a[1] += v_us
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(pr, f"{r}_launches_summary.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 2 --warmup 3 --no-cpu\n")
        f.write(f"# (whole process: warm-up, timed steps, e2e leg and extras; per-launch times are cold-cache and serialised)\n")
        f.write(f"# {'launches':>8s} {'mean us':>10s} {'share':>7s}  kernel\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"  {n:8d} {t / n:10.1f} {100 * t / tot:6.1f}%  {k[:120]}\n")
    print(open(os.path.join(pr, f"{r}_launches_summary.txt")).read())
# ---- traffic
rep = os.path.join(go, f"{r}_all.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    def bytes_of(row, m):
        v, u = float(row[ix[m]].replace(",", "")), units[ix[m]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    names = {"rectify_f32c1_kernel<1>": "c2_f64", "rectify_f32c1_kernel<0>": "c2_f32",
             "rectify_u8c3_kernel<1>": "c3_f64", "rectify_u8c3_kernel<0>": "c3_f32"}
    out = {"source": f"ncu --set full --clock-control none, profiles/prof_kernels.py all ({r}_all.ncu-rep); dram__bytes_read.sum + dram__bytes_write.sum per launch"}
    for row in data:
        for k, key in names.items():
            if k in row[ix["Kernel Name"]].replace("(bool)", "") and key not in out:
                out[key] = int(bytes_of(row, "dram__bytes_read.sum") + bytes_of(row, "dram__bytes_write.sum"))
    json.dump(out, open(os.path.join(pr, "traffic.json"), "w"), indent=1)
    print(out)
