import csv, sys, collections, re
path, start_kernel, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum": continue
    name = r["Kernel Name"]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
    rows.append((name, ms))
# take the LAST complete group that starts at `start_kernel`
idx = [i for i, (n, _) in enumerate(rows) if start_kernel in n]
mode = sys.argv[4] if len(sys.argv) > 4 else "last"
if mode == "last":
    a, b = idx[-2], idx[-1]
else:
    a, b = idx[-1], len(rows)
grp = rows[a:b]
tot = sum(m for _, m in grp)
agg = collections.OrderedDict()
for n, m in grp:
    n = re.sub(r"\(.*", "", n)
    c, t = agg.get(n, (0, 0.0)); agg[n] = (c + 1, t + m)
print(f"# {title}")
print(f"# total {tot:.3f} ms over {len(grp)} launches (per-launch times are cold-cache and serialised: compare SHARES)")
print("kernel,launches,total_ms,share")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n},{c},{t:.3f},{t/tot:.4f}")
