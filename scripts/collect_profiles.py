"""Turn the raw outputs of scripts/gpu_round.sh (gpurun_out/) into the tracked
evidence directory profiles/<name>/: ncu metric summaries, opcode mix, per-source
instruction shares, launch shares, DRAM traffic of the dominant kernel, bench and
test logs.  Run in the authoring container after the GPU call (needs ncu,
cuobjdump, nvdisasm and the same libhermespy_rt.so that was profiled).
usage: collect_profiles.py <name>"""
import collections, csv, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
name = sys.argv[1]
OUT = os.path.join(ROOT, "profiles", name); os.makedirs(OUT, exist_ok=True)
LIB = os.path.join(ROOT, "hermespy-rt_b200", "libhermespy_rt.so")


def sh(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, **kw).stdout


def raw_metrics(rep):
    rows = list(csv.reader(sh(["ncu", "-i", rep, "--page", "raw", "--csv"]).splitlines()))
    H = rows[0]
    return [{h: r[i] for i, h in enumerate(H)} for r in rows[2:]], {h: rows[1][i] for i, h in enumerate(H)}


def fnum(s):
    try: return float(s.replace(",", ""))
    except Exception: return None


def source_shares(rep, kernel_substr, out_path):
    sass = sh(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"])
    rows = list(csv.reader(sass.splitlines()))
    H = rows[1]; data = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name": break
        if len(r) == len(H): data.append(r)
    ix = {h: i for i, h in enumerate(H)}
    tmp = "/tmp/collect_dis"; shutil.rmtree(tmp, ignore_errors=True); os.makedirs(tmp)
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.startswith("hrt_cuda")][0]
    lines = sh(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)]).split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kernel_substr in l)
    cur = None; seq = []
    for l in lines[start + 1:]:
        if l.startswith("//---------------------"): break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): seq.append((cur, l.split("*/")[1].strip()))
    src = {}
    def text(key):
        f, ln = key
        try:
            if f not in src: src[f] = open(os.path.join(ROOT, "hermespy-rt_b200", "csrc", f)).read().split("\n")
            return src[f][ln - 1].strip()[:100]
        except Exception: return ""
    # function of each source line: nearest preceding "HRT_HD ... name(" / "__global__" definition
    def func(key):
        if not key: return "?"
        f, ln = key
        try:
            if f not in src: src[f] = open(os.path.join(ROOT, "hermespy-rt_b200", "csrc", f)).read().split("\n")
        except Exception: return f
        for k in range(ln - 1, -1, -1):
            m = re.match(r"^(?:template.*>\s*)?(?:HRT_HD|__device__ __forceinline__|__global__|static|__device__)[^;]*?\b([A-Za-z_0-9]+)\(", src[f][k])
            if m: return m.group(1)
            m = re.match(r"^([a-z_0-9]+)\(RunDev", src[f][k])
            if m: return m.group(1)
        return f
    agg = collections.Counter(); thr = collections.Counter(); fa = collections.Counter(); ft = collections.Counter()
    n = min(len(data), len(seq))
    for k in range(n):
        c = int(data[k][ix["Instructions Executed"]]); t = int(data[k][ix["Thread Instructions Executed"]])
        agg[seq[k][0]] += c; thr[seq[k][0]] += t
        fn = func(seq[k][0]); fa[fn] += c; ft[fn] += t
    tot = sum(agg.values()); tt = sum(thr.values())
    # share of the issue slots spent on the closest-hit query itself (traversal, slab and
    # triangle tests, their loads and ray set-up) as opposed to the per-path math after it
    INTERSECTION = {"hrt_closest_hit", "hrt_mt_test", "hrt_slab_sorted", "hrt_slab", "hrt_fma_pair", "v3_dot", "v3_cross",
                    "v3_sub", "lds128", "lds32", "smem_base_addr", "node", "tri", "child_at", "child_ref", "cache_word",
                    "hrt_safe_inv", "hrt_octant", "hrt_ray_cull", "hrt_origin_chain", "select_octant", "query", "origin_chain",
                    "hrt_closest_hit_wide", "wide", "select_wide_octant", "scene_tri_off",
                    "query_map", "hrt_rxmap_cells2", "hrt_rxmap_query_depth", "hrt_rxmap_stop", "hrt_mt_self_miss", "hrt_mt_self_nt"}
    share = sum(c for fn, c in fa.items() if fn in INTERSECTION) / max(tot, 1)
    json.dump({"kernel": kernel_substr, "intersection_share_of_issue_slots": share,
               "functions": {fn: c / tot for fn, c in fa.most_common(30)},
               "source": os.path.relpath(out_path, ROOT)},
              open(out_path.replace("_by_source.txt", "_shares.json"), "w"), indent=1)
    with open(out_path, "w") as f:
        f.write(f"kernel {kernel_substr}: {len(data)} SASS instructions, {tot:.4e} warp-instructions executed, "
                f"{tt / tot:.2f} active threads per instruction\n\nby function (share of issued warp-instructions, active threads):\n")
        for fn, c in fa.most_common(25):
            f.write(f"  {100 * c / tot:5.1f}%  thr {ft[fn] / max(c, 1):4.1f}  {fn}\n")
        f.write("\nby source line:\n")
        for key, c in agg.most_common(45):
            if not key: continue
            f.write(f"  {100 * c / tot:5.1f}%  thr {thr[key] / max(c, 1):4.1f}  {key[0]}:{key[1]}  {text(key)}\n")
    return sass


def launch_shares(csv_path, out_path):
    rows = list(csv.reader(l for l in open(csv_path) if l.startswith('"')))
    H = rows[0]; ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    t = collections.Counter(); n = collections.Counter()
    for r in rows[1:]:
        if len(r) != len(H): continue
        k = re.sub(r"\(.*", "", r[ki])[:70]; t[k] += float(r[vi].replace(",", "")) / 1e3; n[k] += 1
    tot = sum(t.values())
    with open(out_path, "w") as f:
        f.write("kernel time shares, ncu --metrics gpu__time_duration.sum (one bench step incl. the counted pass and the fp32 probe)\n")
        for k, v in t.most_common(30):
            f.write(f"  {v:12.1f} us  {n[k]:4d}x  {100 * v / tot:5.1f}%  {k}\n")


for rep, kern, tag in (("prof_scatter.ncu-rep", os.environ.get("HRT_PROF_KERNEL", "_Z9k_scatterILb1ELb0ELb0ELb0ELb1ELb1EE"), "k_scatter"),
                       ("prof_c5.ncu-rep", "_Z9k_scatterILb0ELb0ELb0ELb0ELb1ELb0EE", "k_scatter_c5")):
    p = os.path.join(G, rep)
    if not os.path.exists(p): continue
    open(os.path.join(OUT, f"{tag}_ncu_metrics.txt"), "w").write(sh([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), p]))
    sass = source_shares(p, kern, os.path.join(OUT, f"{tag}_by_source.txt"))
    open("/tmp/collect_sass.csv", "w").write(sass)
    open(os.path.join(OUT, f"{tag}_sass_mix.txt"), "w").write(sh([sys.executable, os.path.join(ROOT, "scripts", "sass_mix.py"), "/tmp/collect_sass.csv"]))
    rows, units = raw_metrics(p)
    tr = []
    for r in rows:
        def val(k):
            v = fnum(r.get(k, "")); u = units.get(k, "")
            if v is None: return None
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
        dur = fnum(r.get("gpu__time_duration.sum", "")); du = units.get("gpu__time_duration.sum", "")
        tr.append({"kernel": r.get("Kernel Name"), "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
                   "duration_ms": dur * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(du, 1) if dur is not None else None})
    json.dump({"source": f"ncu --set full --clock-control none, {rep}", "launches": tr}, open(os.path.join(OUT, f"{tag}_traffic.json"), "w"), indent=1)
    # issue-slot utilisation and lane occupancy of the first captured launch, next to the shares
    sj_path = os.path.join(OUT, f"{tag}_shares.json")
    sj = json.load(open(sj_path))
    if rows:
        sj["issue_active_pct"] = fnum(rows[0].get("smsp__issue_active.avg.pct_of_peak_sustained_active", ""))
        sj["active_threads_per_instruction"] = fnum(rows[0].get("smsp__thread_inst_executed_per_inst_executed.ratio", ""))
        sj["registers_per_thread"] = fnum(rows[0].get("launch__registers_per_thread", ""))
    json.dump(sj, open(sj_path, "w"), indent=1)
    if tag == "k_scatter":
        shutil.copy(os.path.join(OUT, "k_scatter_shares.json"), os.path.join(ROOT, "profiles", "k_scatter_shares.json"))
    if tag == "k_scatter":   # what bench.py reports as roofline.traffic
        json.dump({"source": f"profiles/{name}/{tag}_traffic.json (ncu --set full, one launch of the 8e6-ray bench step)", "launches": tr},
                  open(os.path.join(ROOT, "profiles", "k_scatter_traffic.json"), "w"), indent=1)
if os.path.exists(os.path.join(G, "launches.csv")):
    shutil.copy(os.path.join(G, "launches.csv"), os.path.join(OUT, "launches.csv"))
    launch_shares(os.path.join(G, "launches.csv"), os.path.join(OUT, "launch_shares.txt"))
for f in ("bench_full.json", "bench_ref.json", "bench_small.json", "pytest_gpu.log", "smoke.log", "c5_shard.json", "dense.json", "fp_ops.csv", "fp_ops_bvh.csv",
          "bench_n2.json", "bench_n4.json", "bench_n8.json", "gpu.txt", "c5_full.json"):
    if os.path.exists(os.path.join(G, f)): shutil.copy(os.path.join(G, f), os.path.join(OUT, f))
print("wrote", OUT, sorted(os.listdir(OUT)))
