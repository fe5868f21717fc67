"""Turns the raw outputs of tools/prof_r2.sh (gpurun_out/r2p) + the final bench / test run (gpurun_out/r2z) into the tracked
summaries under profiles/ (round 2)."""
import collections, csv, hashlib, json, os, re, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P, Z, OUT = os.path.join(ROOT, "gpurun_out", "r2p"), os.path.join(ROOT, "gpurun_out", "r2z"), os.path.join(ROOT, "profiles")

def clean(n):
    return re.sub(r"\(.*", "", n).replace("<unnamed>::", "").replace("void ", "")

def load_times(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    return [(clean(r["Kernel Name"]), float(r["Metric Value"].replace(",", "")) / 1e3) for r in csv.DictReader(lines)
            if r.get("Metric Name") == "gpu__time_duration.sum"]

def summarise(rows, title, out, strip_templates=False):
    tot = sum(v for _, v in rows)
    agg = collections.OrderedDict()
    for n, v in rows:
        if strip_templates: n = re.sub(r"<.*", "", n)
        c, t = agg.get(n, (0, 0.0)); agg[n] = (c + 1, t + v)
    with open(os.path.join(OUT, out), "w") as f:
        f.write(f"# {title}\n# total {tot / 1e3:.3f} ms over {len(rows)} launches (per-launch times are cold-cache and serialised under ncu: compare SHARES)\n")
        f.write("kernel,launches,total_us,avg_us,share\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{n},{c},{t:.1f},{t / c:.1f},{t / tot:.4f}\n")

def first_kernel(rows):          # the stem: im2col on the tensor-core path, stem_hx_kernel otherwise
    return [i for i, (n, _) in enumerate(rows) if n.startswith("stem_im2col") or n.startswith("stem_hx")]

b = load_times(os.path.join(P, "launches_bench.csv"))
st = first_kernel(b)
q4 = [i for i in range(len(b) - 3) if all("query_tc_kernel" in b[i + k][0] for k in range(4))]
# a full bench step = stem .. 4th query launch; take the one after the warm-up
starts = [s for s in st if any(s < q <= s + 400 for q in q4)]
s0 = starts[1] if len(starts) > 1 else starts[0]
e0 = min(q for q in q4 if q > s0) + 4
summarise(b[s0:e0], "ncu launch list, ONE bench step (round 2, final build): encoder (one launch per convolution; replayed as one CUDA graph with the hourglass skip branches on side streams in the product path) + 4 x query_tc_kernel over 4 194 304 grid points: CHORE_B200_ENCODER_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 1 --warmup 1 --no-cpu --no-fit", "launches_r2_bench_step_summary.csv")
small = [s for s in st if s > e0 and not any(s < q <= s + 400 for q in q4)]
if small:
    s1 = small[0]
    e1 = next(i for i in range(s1, len(b)) if "query_tc_kernel" in b[i][0]) + 1
    summarise(b[s1:e1], "ncu launch list, ONE step of the north-star workload (round 2, final build): 512x512 image -> encoder -> 20 000-point query (same command as launches_r2_bench_step_summary.csv)", "launches_r2_image20k_step_summary.csv")
f = load_times(os.path.join(P, "launches_fit.csv"))
adam = [i for i, (n, _) in enumerate(f) if n.startswith("adam_step_kernel")]
# one iteration = (first launch after the previous iteration's second Adam) .. (this iteration's second Adam)
a_prev, a_last = adam[-3], adam[-1]
summarise(f[a_prev + 1:a_last + 1], "ncu launch list, ONE fit iteration of the final round-2 build = SMPL-H step (every forward_smpl term) + object-only step, B=1, 6 890 vertices + 20 000 object points, fused adjoint path launched directly (the product path replays these launches as one two-stream CUDA graph): ncu --metrics gpu__time_duration.sum --clock-control none python bench_fit.py --iters 2 --reps 1 --no-graph", "launches_r2_fit_iteration_summary.csv", strip_templates=True)

# per-launch encoder metrics
lines = [l for l in open(os.path.join(P, "encoder_metrics.csv")) if not l.startswith("==")]
byid = collections.OrderedDict()
for r in csv.DictReader(lines):
    d = byid.setdefault(r["ID"], {"name": clean(r["Kernel Name"])})
    d[r["Metric Name"]] = (r["Metric Value"].replace(",", ""), r["Metric Unit"])
L = list(byid.values())
stems = [i for i, d in enumerate(L) if d["name"].startswith("stem_im2col") or d["name"].startswith("stem_hx")]
grp = L[stems[-1]:]
def val(d, k):
    v, u = d[k]
    try: v = float(v)
    except ValueError: return float("nan")
    if k == "gpu__time_duration.sum": return v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
    if "bytes" in k: return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    return v
agg = collections.OrderedDict()
for d in grp:
    if d["name"].startswith("at::"): continue
    e = agg.setdefault((d["name"], d["launch__grid_size"][0]), [0, 0.0, 0.0, 0.0, 0.0, 0.0])
    t = val(d, "gpu__time_duration.sum")
    e[0] += 1; e[1] += t; e[2] += val(d, "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed") * t
    e[3] += val(d, "dram__bytes_read.sum"); e[4] += val(d, "dram__bytes_write.sum"); e[5] += val(d, "lts__t_bytes.sum")
tot = sum(e[1] for e in agg.values())
with open(os.path.join(OUT, "encoder_r2_per_launch_metrics_summary.csv"), "w") as fo:
    fo.write("# one B=1 encode of the final round-2 build, every launch profiled on its own: CHORE_B200_ENCODER_GRAPH=0 ncu --metrics gpu__time_duration.sum,launch__grid_size,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none python tools/time_encoder.py --batches 1 --iters 1\n")
    fo.write(f"# {sum(e[0] for e in agg.values())} launches, {tot / 1e3:.3f} ms summed (cold-cache, serialised: compare shares; the graph replay takes 2.3 ms); tensor_active = time-weighted mean of sm__mem_tensor_cycles_active (% of elapsed); bytes are per launch\n")
    fo.write("kernel,grid,launches,total_us,avg_us,share,tensor_active_pct,dram_read_MB,dram_write_MB,l2_MB\n")
    for (n, g), e in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fo.write(f"{n},{g},{e[0]},{e[1]:.1f},{e[1] / e[0]:.1f},{e[1] / tot:.4f},{e[2] / e[1]:.1f},{e[3] / e[0] / 1e6:.2f},{e[4] / e[0] / 1e6:.2f},{e[5] / e[0] / 1e6:.2f}\n")

# full captures + traffic.json
rows = list(csv.reader(open(os.path.join(P, "query_tc_r2_raw.csv"))))
hdr, r = rows[0], rows[2]
g = lambda k: float(r[hdr.index(k)])
rd, wr, ms = g("dram__bytes_read.sum"), g("dram__bytes_write.sum"), g("gpu__time_duration.sum")
tp = [float(r[i]) for i, h in enumerate(hdr) if h.endswith("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed")]
sha = lambda p: hashlib.sha256(open(os.path.join(ROOT, p), "rb").read()).hexdigest()[:16]
traffic = {
    "_note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (ncu --set full --clock-control none); valid for the kernel source whose sha256 prefix is recorded, bench.py reports null when the source has changed since",
    "query_tc_kernel": {"dram_bytes_per_launch": int(round((rd + wr) * 1e6)), "source": "chore_b200/csrc/query_tc.cu", "source_sha16": sha("chore_b200/csrc/query_tc.cu"),
                        "capture": f"profiles/query_tc_r2_ncu_details.txt (final round-2 build, 4 194 304 grid points: {rd:.2f} MB read + {wr:.2f} MB written, {ms:.2f} ms, tensor pipe active {tp[0] if tp else float('nan'):.1f} % of elapsed)"},
    "conv_hx_kernel": {"dram_bytes_per_launch": None, "source": "chore_b200/csrc/conv_hx.cu", "source_sha16": sha("chore_b200/csrc/conv_hx.cu"),
                       "capture": "profiles/conv_hx_r2_ncu_details.txt + profiles/encoder_r2_per_launch_metrics_summary.csv (B = 1: 1-25 MB read, 0 written per launch -- activations stay in the 126 MB L2)"},
}
json.dump(traffic, open(os.path.join(OUT, "traffic.json"), "w"), indent=1)
for a, b_ in (("query_tc_r2_details.txt", "query_tc_r2_ncu_details.txt"), ("conv_hx_r2_details.txt", "conv_hx_r2_ncu_details.txt"), ("encoder_sweep.json", "encoder_sweep_r2.json")):
    shutil.copy(os.path.join(P, a), os.path.join(OUT, b_))
for a, b_ in (("bench_n1.json", "bench_r2_n1.json"), ("bench_ref.json", "bench_r2_reference_arm.json"), ("parity_report.jsonl", "parity_report_r2.jsonl")):
    if os.path.exists(os.path.join(Z, a)): shutil.copy(os.path.join(Z, a), os.path.join(OUT, b_))
print(open(os.path.join(OUT, "launches_r2_bench_step_summary.csv")).read())
print(open(os.path.join(OUT, "encoder_r2_per_launch_metrics_summary.csv")).read())
print(open(os.path.join(OUT, "launches_r2_fit_iteration_summary.csv")).read()[:1500])
print(json.dumps(traffic["query_tc_kernel"]))
print(open(os.path.join(P, "encoder_sweep.json")).read())
