"""Developer tool: per-kernel totals and shares of an `ncu --metrics gpu__time_duration.sum --csv`
launch list.  usage: python tools/launch_list.py launches.csv [header line ...]"""
import collections, csv, re, sys

rows = []
with open(sys.argv[1]) as f:
    body = [l for l in f if not l.startswith("==")]
for d in csv.DictReader(body):
    rows.append((re.sub(r"\(.*", "", d["Kernel Name"])[:96], float(d["Metric Value"].replace(",", "")) / 1e3))
tot = collections.defaultdict(float)
cnt = collections.Counter()
for n, v in rows:
    tot[n] += v
    cnt[n] += 1
total = sum(tot.values())
for h in sys.argv[2:]:
    print("# " + h)
print(f"# {len(rows)} launches captured; times are cold-cache and serialised: compare SHARES, not absolutes")
print("# total_us share launches kernel")
for n, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v / total < 0.002:
        continue
    print(f"{v:12.1f} {100 * v / total:5.1f}% {cnt[n]:4d} {n}")
ours = sum(v for n, v in tot.items() if "qcp::" in n)
print(f"# qcpinn_b200 kernels: {100 * ours / total:.1f}% of captured GPU time; "
      f"torch glue (start-up eager steps, Adam, copies): {100 - 100 * ours / total:.1f}%")
