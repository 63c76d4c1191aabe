"""Write / update profiles/ncu_traffic.json from `ncu --set full` reports: per-launch DRAM traffic of the kernels bench.py
reports a roofline for, keyed by the sha256 of the kernel's source files (bench.py drops an entry whose sources changed).

    python tools/ncu_traffic.py <key> <report.ncu-rep> <kernel regex> <shape> <B> <K> <source file> [<source file> ...]
    NCU_CALLS=<n>: the matching launches belong to n calls of the entry point (e.g. scan + fix-up = one scatter call): the traffic
    is summed per call instead of averaged per launch.
e.g. python tools/ncu_traffic.py gather gpurun_out/prof_gather.ncu-rep k_gather_flat kkbox 4096 5 www24-rat_b200/csrc/gather.cu"""
import csv, hashlib, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
key, rep, rx, shape, B, K = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
sources = sys.argv[7:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
tot, n, dur = 0.0, 0, 0.0
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if not re.search(rx, d["Kernel Name"]):
        continue
    u = dict(zip(hdr, units))
    def bytes_of(m):
        v = float(d[m]); unit = u[m].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
    tot += bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum")
    dur += float(d["gpu__time_duration.sum"])
    n += 1
assert n > 0, f"no kernel matching {rx} in {rep}"
h = hashlib.sha256()
for f in sources:
    h.update(open(os.path.join(ROOT, f), "rb").read())
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
db = json.load(open(path)) if os.path.exists(path) else {}
calls = int(os.environ.get("NCU_CALLS", "0"))
if calls:
    n_div = calls
else:
    n_div = n
db[key] = {"traffic_bytes": tot / n_div, "launches_averaged": n, "kernel": rx, "gpu_time_us_under_ncu": dur / n_div / 1e3 if False else dur / n_div, "shape": shape, "B": B, "K": K,
           "sources": sources, "sha16": h.hexdigest()[:16], "capture": os.path.basename(rep)}
json.dump(db, open(path, "w"), indent=1, sort_keys=True)
print(key, db[key])
