"""Aggregate an ncu launch list (--csv with gpu__time_duration.sum and optionally dram__bytes_read/write.sum) by kernel:
    python scratch/agg_launches.py FILE.csv [top N] [--traffic KERNEL_SUBSTRING WORKLOAD_KEY OUT.json]"""
import collections, csv, json, re, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])  # launches, us, dram bytes
tot = 0.0
for row in rows:
    name = re.sub(r'<.*', '', row['Kernel Name'])[:70]
    v = float(row['Metric Value'].replace(',', ''))
    unit = row.get('Metric Unit', '')
    m = row['Metric Name']
    if m == 'gpu__time_duration.sum':
        us = v / 1e3 if unit in ('ns', 'nsecond') else (v * 1e3 if unit in ('ms', 'msecond') else v)
        agg[name][0] += 1
        agg[name][1] += us
        tot += us
    elif m.startswith('dram__bytes'):
        mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
        agg[name][2] += v * mult
top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 30
print("total us %.1f kernels %d" % (tot, sum(a[0] for a in agg.values())))
mine = sum(a[1] for k, a in agg.items() if 'mpc::' in k or 'tc::' in k or 'knntc::' in k or 'tcb::' in k)
print("ours us %.1f share %.3f" % (mine, mine / tot))
print("%-72s %5s %12s %6s %14s" % ("kernel", "n", "us", "share", "DRAM MB/launch"))
for k, (n, t, by) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-72s %5d %12.1f %5.1f%% %14.2f" % (k, n, t, 100 * t / tot, by / max(n, 1) / 1e6))
if "--traffic" in sys.argv:
    i = sys.argv.index("--traffic")
    sub, key, out = sys.argv[i + 1], sys.argv[i + 2], sys.argv[i + 3]
    sel = [(k, a) for k, a in agg.items() if sub in k]
    n = sum(a[0] for _, a in sel); us = sum(a[1] for _, a in sel); by = sum(a[2] for _, a in sel)
    try:
        cur = json.load(open(out))
    except Exception:
        cur = {}
    cur[key] = [e for e in cur.get(key, []) if isinstance(e, dict) and e.get("kernel") != sub] if isinstance(cur.get(key), list) else []
    cur[key].append({"kernel": sub, "launches_per_step": n, "dram_bytes_per_launch": by / max(n, 1), "kernel_us_per_step": us,
                "step_kernel_us_total": tot, "share_of_step": us / tot,
                "source": "%s (ncu gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum; one eager step of bench.py)" % path})
    json.dump(cur, open(out, "w"), indent=1)
