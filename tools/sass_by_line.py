#!/usr/bin/env python
"""Join an ncu SASS source page (--page source --csv) with nvdisasm line info: executed warp-instructions per CUDA
source line of one kernel. Usage: sass_by_line.py <report.ncu-rep> <lib.so> <mangled-kernel-substring> [n_units]"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, lib, kern = sys.argv[1:4]
units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
sections, cur_sec = [], None           # one section per profiled kernel: ["Kernel Name", name], header row, instruction rows
for r in rows:
    if r and r[0] == "Kernel Name":
        cur_sec = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur_sec)
    elif cur_sec is not None and cur_sec["hdr"] is None:
        cur_sec["hdr"] = r
    elif cur_sec is not None:
        cur_sec["data"].append(r)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
dis, start = None, None
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):     # one cubin per translation unit: find the kernel's
    if kern not in subprocess.run(["cuobjdump", "-elf", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout:
        continue
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
    start = next((i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l), None)
    if start is not None:
        break
assert start is not None, f"kernel {kern} not found in {lib}"
lines = []          # (file:line stack, sass text)
cur = "?"
for l in dis[start + 1:]:
    if l.startswith("//---") and ".text." in l:
        break
    m = re.match(r"\s*//## File \"(.*?)\", line (\d+)(.*)", l)
    if m:
        inl = re.findall(r"inlined at \"(.*?)\", line (\d+)", l)
        cur = os.path.basename(m.group(1)) + ":" + m.group(2) + "".join(" <- " + os.path.basename(a) + ":" + b for a, b in inl)
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m:
        lines.append((cur, m.group(1)))
sec = next((x for x in sections if len(x["data"]) == len(lines)), None)
assert sec is not None, (len(lines), [(x["name"][:40], len(x["data"])) for x in sections])
hdr, data = sec["hdr"], sec["data"]
iex, ismp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
print("kernel:", sec["name"])
agg, smp = collections.Counter(), collections.Counter()
tot = 0
for (loc, _), r in zip(lines, data):
    # attribute to the outermost kernel-level line and to the innermost
    agg[loc] += int(r[iex]); smp[loc] += int(r[ismp]); tot += int(r[iex])
print(f"total warp-instr {tot}  per unit {tot/units:.2f}")
for loc, c in agg.most_common(70):
    print(f"{c:12d} {100*c/tot:6.2f}% {c/units:8.2f}/unit smp {smp[loc]:6d}  {loc}")
