"""Round-2 evidence: turn the raw GPU outputs (gpurun_out/, scratch) into the committed files under profiles/r2/.
    python profiles/mk_evidence_r2.py
inputs  gpurun_out/r2_bench.json, r2_bench_ref.json   plain `python bench.py` / `--impl reference` runs (no profiler)
        gpurun_out/r2_launches.csv                    ncu --metrics gpu__time_duration.sum --clock-control none -c 4000  python bench.py --steps 2 --warmup 3 --no-cpu
        gpurun_out/r2_all.ncu-rep                     ncu --set full --import-source on --clock-control none  python profiles/prof_kernels.py all --reps 1
outputs profiles/r2/bench.json, bench_ref.json, launches.csv, launches_summary.txt, ncu_full_summary.txt,
        ncu_hot_<kernel>.txt (executed-instruction mix + stall reasons of the hot kernels), sass_<kernel>.txt (static
        SASS excerpts from the built object: TMA / mbarrier instructions and the hot loop), traffic.json
"""
import collections, csv, io, json, os, re, shutil, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, out = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles", "r2")
os.makedirs(out, exist_ok=True)

def json_line(path):
    for l in open(path):
        if l.startswith("{"):
            return l
    return None

benches = {}
for a, b in (("r2_bench.json", "bench.json"), ("r2_bench_ref.json", "bench_ref.json"), ("r2_bench_n2.json", "bench_n2.json"),
             ("r2_bench_n4.json", "bench_n4.json"), ("r2_bench_n8.json", "bench_n8.json")):
    if os.path.exists(os.path.join(go, a)):
        l = json_line(os.path.join(go, a))
        if l:
            open(os.path.join(out, b), "w").write(l)
            benches[b] = json.loads(l)
if os.path.exists(os.path.join(go, "r2_topo.txt")):
    shutil.copy(os.path.join(go, "r2_topo.txt"), os.path.join(out, "topo_8gpu.txt"))

# ---- 1/2/4/8 GPUs: the named multi-GPU configs and the host-link ceiling, one table
rows = [(n, benches.get(f)) for n, f in ((1, "bench.json"), (2, "bench_n2.json"), (4, "bench_n4.json"), (8, "bench_n8.json"))]
if all(d for _, d in rows):
    with open(os.path.join(out, "scaling_summary.txt"), "w") as f:
        f.write("# python bench.py (N=1) / torchrun ... bench.py --gpus N --steps 20 --warmup 3 on one 8 x B200 box (N=1: its own 1-GPU box)\n")
        f.write("# device-resident values are CUDA-event times, max over ranks; e2e and the stream are wall clock with host buffers\n")
        f.write("# link = pinned<->device copies, both directions at once, EVERY rank at once (GB/s per direction per GPU)\n")
        f.write(f"{'N':>2s} {'C2 Gpix/s':>10s} {'frac':>6s} {'e2e Gpix/s':>11s} {'link/GPU':>9s} {'link sum':>9s} {'e2e/link':>9s} "
                f"{'C3 stream':>10s} {'C3 dev f64':>11s} {'C3 dev f32':>11s} {'C4 f64 Gpt/s':>13s} {'C4 f32':>8s} {'C5 JtJ us':>10s} {'LM us/it':>9s}\n")
        for n, d in rows:
            e, x = d["e2e"], d["extras"]
            lk = e["link_ceiling"]["duplex_gbs_per_dir_per_gpu"]
            f.write(f"{n:2d} {d['value'] / 1e3:10.1f} {d['roofline']['frac']:6.3f} {e['value'] / 1e3:11.2f} {lk:9.1f} {lk * n:9.1f} {e['link_frac']:9.3f} "
                    f"{x['c3_stream_f64']['mpix_per_s_e2e'] / 1e3:10.2f} {x['c3_device_f64']['mpix_per_s'] / 1e3:11.1f} {x['c3_device_f32']['mpix_per_s'] / 1e3:11.1f} "
                    f"{x['c4_img2world_f64_100M']['gpt_per_s']:13.1f} {x['c4_img2world_f32_100M']['gpt_per_s']:8.1f} "
                    f"{x['c5_reproj_jtj_10k_views']['ms'] * 1e3:10.1f} {x['c5_lm_iterations_10k_views']['us_per_iteration']:9.1f}\n")
        f.write("# C2 / C3 device: weak scaling (every rank its own batch of the same size); C3 stream: 4096 frames, C4: 100 M points,\n")
        f.write("# C5: 10k views -- each split over the ranks (strong); C5 has the 21-double NCCL all-reduce inside the timed region for N > 1\n")

# ---- launch list ------------------------------------------------------------------------------------
src = os.path.join(go, "r2_launches.csv")
if os.path.exists(src):
    lines = [l for l in open(src) if l.startswith('"')]
    open(os.path.join(out, "launches.csv"), "w").writelines(lines)
    agg = collections.OrderedDict()
    for x in csv.DictReader(io.StringIO("".join(lines))):
        if x["Metric Name"] != "gpu__time_duration.sum":
            continue
        v, u = float(x["Metric Value"].replace(",", "")), x["Metric Unit"]
        us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
        a = agg.setdefault(x["Kernel Name"], [0, 0.0])
        a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(out, "launches_summary.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 4000  python bench.py --steps 2 --warmup 3 --no-cpu\n")
        f.write("# whole process (warm-up, timed steps, e2e leg, extras); per-launch times are cold-cache and serialised:\n")
        f.write("# the SHARE column is what must agree with bench.py, not the absolute times\n")
        f.write(f"# {'launches':>8s} {'mean us':>10s} {'share':>7s}  kernel\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"  {n:8d} {t / n:10.1f} {100 * t / tot:6.1f}%  {k[:130]}\n")

# ---- ncu --set full -----------------------------------------------------------------------------------
rep = os.path.join(go, "r2_all.ncu-rep")
W = [("time_us", "gpu__time_duration.sum"), ("dram_rd", "dram__bytes_read.sum"), ("dram_wr", "dram__bytes_write.sum"),
     ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
     ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("warp_inst", "smsp__inst_executed.sum"),
     ("occ_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread"),
     ("l2_hit", "lts__t_sector_hit_rate.pct"), ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
     ("fp64_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
     ("xu_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
     ("fma_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
     ("alu_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
     ("lsu_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
     ("l1tex_pct", "l1tex__throughput.avg.pct_of_peak_sustained_active"),
     ("lts_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"), ("grid", "launch__grid_size"), ("block", "launch__block_size")]
# a later capture of the views kernels alone (python profiles/prof_kernels.py views --reps 1) replaces their entries
rep_views = os.path.join(go, "r2_views.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    if os.path.exists(rep_views):
        rv = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep_views, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
        if rv:                                                  # columns matched by name (the two captures may differ)
            kn, col = hdr.index("Kernel Name"), {h: i for i, h in enumerate(rv[0])}
            data = [r for r in data if "views_kernel" not in r[kn]] + \
                   [[r[col[h]] if h in col else "" for h in hdr] for r in rv[2:]]
    ix = {h: i for i, h in enumerate(hdr)}
    seen, traffic = set(), {"source": "ncu --set full --clock-control none, profiles/prof_kernels.py all (r2_all.ncu-rep); dram__bytes_read.sum + dram__bytes_write.sum per launch"}
    tnames = {"rectify_f32c1_kernel<1>": "c2_f64", "rectify_f32c1_kernel<0>": "c2_f32", "rectify_u8c3_kernel<1>": "c3_f64", "rectify_u8c3_kernel<0>": "c3_f32"}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    with open(os.path.join(out, "ncu_full_summary.txt"), "w") as f:
        f.write("# ncu --set full --import-source on --clock-control none  python profiles/prof_kernels.py all --reps 1   (first instance of each kernel)\n")
        for r in data:
            name = r[ix["Kernel Name"]].replace("(bool)", "")
            if name in seen or "cc::" not in name:          # own kernels only (torch's input generators are not on the path)
                continue
            seen.add(name)
            f.write("## " + name[:140] + "\n")
            for short, m in W:
                if m in ix:
                    v = r[ix[m]]
                    try:
                        v = f"{float(v.replace(',', '')):,.3f}".rstrip("0").rstrip(".")
                    except ValueError:
                        pass
                    f.write(f"  {short:16s} {v:>20s} {units[ix[m]]}\n")
            for k, key in tnames.items():
                if k in name and key not in traffic:
                    rd = float(r[ix["dram__bytes_read.sum"]].replace(",", "")) * scale[units[ix["dram__bytes_read.sum"]]]
                    wr = float(r[ix["dram__bytes_write.sum"]].replace(",", "")) * scale[units[ix["dram__bytes_write.sum"]]]
                    traffic[key] = int(rd + wr)
    json.dump(traffic, open(os.path.join(out, "traffic.json"), "w"), indent=1)
    json.dump(traffic, open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)   # bench.py reads this one

    # ---- executed-instruction mix and stall reasons of the hot kernels (ncu source page)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if os.path.exists(rep_views):           # the later capture first: the first instance of a kernel wins below
        rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep_views, "--page", "source", "--csv"], capture_output=True, text=True).stdout))) + rows
    kern, cur, hdr2 = collections.OrderedDict(), None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = r[1].replace("(bool)", "")
            cur = None if cur in kern else cur
            if cur:
                kern[cur] = []
            continue
        if r and r[0] == "Address":
            hdr2 = r
            continue
        if cur and r and r[0].startswith("0x"):
            kern[cur].append(r)
    for name, ins in kern.items():
        m = re.search(r"(rectify_\w+_kernel<\d>|reproj_jtj_kernel|img2world_kernel<\w+>|world2img_kernel<\w+>)", name)
        if not m or not ins:
            continue
        tag = re.sub(r"[<>]", "_", m.group(1)).rstrip("_")
        ops, stalls = collections.Counter(), collections.Counter()
        tot = sum(int(r[5]) for r in ins)
        for r in ins:
            op = re.sub(r"^@!?U?P[0-9T]+\s+", "", r[1].strip()).split()[0]
            ops[op] += int(r[5])
            for c in range(29, 46):
                stalls[hdr2[c]] += int(r[c])
        with open(os.path.join(out, f"ncu_hot_{tag}.txt"), "w") as f:
            f.write(f"# {name[:120]}\n# executed warp instructions {tot:,}; opcode mix (share of executed) and warp-stall samples by reason\n")
            for op, c in ops.most_common(28):
                f.write(f"  {op:34s} {c:12,d}  {100 * c / tot:5.1f}%\n")
            st = sum(stalls.values())
            f.write("# stall samples\n")
            for k, v in stalls.most_common(10):
                f.write(f"  {k:28s} {v:8d}  {100 * v / st:5.1f}%\n")

# ---- static SASS excerpts from the built object ---------------------------------------------------------
obj = os.path.join(root, "cameracalibrations_b200", "csrc", "_build", "rectify.o")
if os.path.exists(obj):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, fn = None, collections.OrderedDict()
    for l in sass.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            cur = m.group(1); fn[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m and cur:
            fn[cur].append((int(m.group(1), 16), m.group(2).strip()))
    for mangled, lines in fn.items():
        m = re.search(r"rectify_(f32c1|u8c3)_kernelILb([01])", mangled)
        if not m:
            continue
        tag = f"rectify_{m.group(1)}_kernel_{m.group(2)}"
        tma = [(a, s) for a, s in lines if re.search(r"UTMALDG|UTMASTG|UTMACCTL|SYNCS|FENCE", s)]
        # hot loop = the longest basic block that contains STG and LDS
        blocks, start = [], 0
        for i, (a, s) in enumerate(lines):
            if re.search(r"\b(BRA|EXIT|BSYNC|CALL)\b", s):
                blocks.append(lines[start:i + 1]); start = i + 1
        hot = max((b for b in blocks if any("STG" in s for _, s in b) and any("LDS" in s for _, s in b)), key=len, default=[])
        hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", s).split()[0] for _, s in hot)
        with open(os.path.join(out, f"sass_{tag}.txt"), "w") as f:
            f.write(f"# cuobjdump -sass cameracalibrations_b200/csrc/_build/rectify.o : {mangled}\n")
            f.write(f"# {len(lines)} instructions; TMA / mbarrier / fence instructions:\n")
            for a, s in tma:
                f.write(f"  {a:05x}  {s}\n")
            f.write(f"# hot block (all-staged tile, 8 pixels per lane): {len(hot)} instructions at {hot[0][0]:#x}..{hot[-1][0]:#x}\n")
            for k, v in hist.most_common():
                f.write(f"#   {k:26s} {v}\n")
            for a, s in hot:
                f.write(f"  {a:05x}  {s}\n")
print("\n".join(sorted(os.listdir(out))))
