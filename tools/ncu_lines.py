"""Per-source-line instruction / stall-sample hotspots of one kernel in an ncu report.
   ncu_lines.py <rep> <kernel regex> [launch-skip] [topN]"""
import csv, subprocess, sys, io, collections
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                      "regex:" + rx, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = None
agg = collections.defaultdict(lambda: [0, 0, ""])
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r; ie = r.index("Instructions Executed"); sm = r.index("# Samples"); continue
    if hdr is None or len(r) <= ie or not r[0].isdigit():
        continue
    try:
        n = int(r[ie] or 0); s = int(r[sm] or 0)
    except ValueError:
        continue
    k = (fname, int(r[0]))
    agg[k][0] += n; agg[k][1] += s; agg[k][2] = r[1]
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print(f"total warp-inst {tot:,}  samples {tots:,}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{v[0]/max(tot,1)*100:5.1f}% inst {v[1]/max(tots,1)*100:5.1f}% smp  {k[0]}:{k[1]:<5d} | {v[2].strip()[:100]}")

# optional phase grouping: ncu_lines.py <rep> <regex> <skip> <top> file:lo-hi=name ...
groups = [g for g in sys.argv[5:] if "=" in g]
if groups:
    print("--- phase groups (inst %, sample %)")
    for gspec in groups:
        rng, name = gspec.split("=")
        f, lh = rng.split(":")
        lo, hi = (int(x) for x in lh.split("-"))
        gi = sum(v[0] for k, v in agg.items() if k[0] == f and lo <= k[1] <= hi)
        gs = sum(v[1] for k, v in agg.items() if k[0] == f and lo <= k[1] <= hi)
        print(f"  {name:28s} {gi/max(tot,1)*100:5.1f}% inst  {gs/max(tots,1)*100:5.1f}% smp")
