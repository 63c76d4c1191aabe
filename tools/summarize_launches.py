"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals / shares."""
import csv, collections, sys, re
path = sys.argv[1]
rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
idx = {h: i for i, h in enumerate(hdr)}
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows:
    if r is hdr or len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("rat::", "")
    unit = r[idx["Metric Unit"]]
    v = float(r[idx["Metric Value"]].replace(",", ""))
    v_us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
    tot[name] += v_us; cnt[name] += 1
total = sum(tot.values())
print(f"# {path}: {sum(cnt.values())} launches, {total/1e3:.3f} ms total (cold-cache, serialised under ncu: compare SHARES)")
print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"| {k} | {cnt[k]} | {v:.1f} | {v/cnt[k]:.1f} | {v/total*100:.1f}% |")
