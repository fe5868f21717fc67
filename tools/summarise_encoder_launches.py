"""Group an ncu launch list (gpu__time_duration.sum + launch__grid_size) of one encode by kernel and grid size.
usage: python tools/summarise_encoder_launches.py <csv> [title]"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
by = {}
for r in csv.DictReader(lines):
    d = by.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r.get("Grid Size")})
    d[r["Metric Name"]] = (r["Metric Value"], r["Metric Unit"])
rows = [by[k] for k in sorted(by, key=int)]
idx = [i for i, r in enumerate(rows) if "stem" in r["name"]]
grp = rows[idx[-1]:]
tot, agg = 0.0, collections.OrderedDict()
for r in grp:
    v, u = r["gpu__time_duration.sum"]
    us = float(v.replace(",", "")) / (1e3 if u in ("ns", "nsecond") else 1.0)
    tot += us
    name = r["name"].split("(")[0].split("::")[-1]
    key = (name, r["grid"])
    c, s = agg.get(key, (0, 0.0))
    agg[key] = (c + 1, s + us)
print(f"# {sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}")
print(f"# last encode of the run: {len(grp)} launches, {tot / 1e3:.3f} ms summed (per-launch times are cold-cache and serialised: compare SHARES)")
print("kernel,grid,launches,total_us,avg_us,share")
for (n, g), (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n},\"{g}\",{c},{s:.1f},{s / c:.1f},{s / tot:.4f}")
