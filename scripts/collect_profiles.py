#!/usr/bin/env python
"""Copy the bench lines scripts/refresh_profiles.sh left in gpurun_out/ into profiles/ and condense the ncu launch list."""
import collections, csv, json, os, re, shutil, sys

TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"          # round prefix of the artefacts

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
for f in sorted(os.listdir(SRC)):
    if (f.startswith(TAG + "_bench") and f.endswith(".json")) or (f.startswith(TAG + "_ncu_") and f.endswith(".txt")) or f == TAG + "_pytest_gpu.txt":
        if f.endswith(".json"):
            json.loads(open(os.path.join(SRC, f)).read().strip())      # must be one valid JSON line
        shutil.copy(os.path.join(SRC, f), os.path.join(DST, f))
        print("copied", f)
path = os.path.join(SRC, "launches.csv")
if os.path.exists(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ik])[:80]
        n, t = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, t + float(r[iv].replace(",", "")) / 1e3)
    with open(os.path.join(DST, TAG + "_launches_summary.csv"), "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none) of `python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu`\n")
        f.write("# cold-cache, serialised per-launch times: compare SHARES, not absolutes\n")
        f.write("kernel,launches,total_us,avg_us\n")
        for k, (n, t) in agg.items():
            f.write('"%s",%d,%.1f,%.1f\n' % (k, n, t, t / n))
    dec = sum(t for k, (n, t) in agg.items() if "rx_decode_kernel" in k)
    acq = sum(t for k, (n, t) in agg.items() if "rx_acquire_kernel" in k)
    print("launch list: decode share of (acquire + decode) = %.1f %%" % (100 * dec / (dec + acq)))
