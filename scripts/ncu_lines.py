"""Join an ncu report's SASS-level samples with nvdisasm line info -> per-source-line hot spots.

usage: python scripts/ncu_lines.py <report.ncu-rep> <kernel-substring> [top_n]
(the .so must be the build that was profiled)"""
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kname = sys.argv[1], sys.argv[2]
sect = os.environ.get("SECTION", kname)  # mangled-name substring selecting the nvdisasm section (template instance)
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "cityseer_b200", "libcityseer_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# collect (line) per instruction of the kernel, in order
lines = []
inside = False
cur = None
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        inside = sect in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kname}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci = {h: i for i, h in enumerate(hdr)}
sass = [r for r in rows[hdr_i + 1:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
n = min(len(sass), len(lines))
if len(sass) != len(lines):
    print(f"warning: {len(sass)} profiled instructions vs {len(lines)} disassembled", file=sys.stderr)
agg = {}
tot_s = tot_i = 0.0
for k in range(n):
    r = sass[k]
    s = float(r[ci["# Samples"]] or 0)
    ins = float(r[ci["Instructions Executed"]] or 0)
    lsb = float(r[ci["stall_long_sb"]] or 0)
    key = lines[k]
    a = agg.setdefault(key, [0.0, 0.0, 0.0])
    a[0] += s
    a[1] += ins
    a[2] += lsb
    tot_s += s
    tot_i += ins
srcs = {}
print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.3e}")
for key, (s, ins, lsb) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    if key is None:
        txt = "?"
    else:
        fn = key[0]
        if fn not in srcs:
            p = os.path.join(root, "cityseer_b200", "csrc", fn)
            srcs[fn] = open(p).read().splitlines() if os.path.exists(p) else []
        txt = srcs[fn][key[1] - 1].strip() if key[1] - 1 < len(srcs[fn]) else ""
    print(f"{str(key):34s} samp {s / tot_s * 100:5.1f}%  inst {ins / tot_i * 100:5.1f}%  long_sb {lsb / tot_s * 100:5.1f}% | {txt[:100]}")
if os.environ.get("PHASES"):
    # PHASES="name:lo-hi,name:lo-hi" aggregates samples / instructions per line range of the first kernel file
    for spec in os.environ["PHASES"].split(","):
        nm, rng = spec.split(":")
        lo, hi = [int(x) for x in rng.split("-")]
        s = sum(v[0] for k, v in agg.items() if k and k[0].startswith("cs_") and lo <= k[1] <= hi)
        i = sum(v[1] for k, v in agg.items() if k and k[0].startswith("cs_") and lo <= k[1] <= hi)
        print(f"phase {nm:10s} samples {s / tot_s * 100:5.1f}%  instructions {i / tot_i * 100:5.1f}%")
