"""Per-phase stall-reason totals of an ncu report: python scripts/ncu_stalls.py <rep> <kernel-regex> <section> name:lo-hi,..."""
import csv, glob, io, os, re, subprocess, sys, tempfile
rep, kname, sect, spec = sys.argv[1:5]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "cityseer_b200", "libcityseer_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines, inside, cur = [], False, None
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
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
sass = [r for r in rows[hi + 1:] if r and r[0].startswith("0x")]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
phases = []
for s in spec.split(","):
    nm, rng = s.split(":")
    lo, hi_ = [int(x) for x in rng.split("-")]
    phases.append((nm, lo, hi_))
tot = {nm: {} for nm, _, _ in phases}
allsum = 0.0
for k in range(min(len(sass), len(lines))):
    key = lines[k]
    r = sass[k]
    s = float(r[ci["# Samples"]] or 0)
    allsum += s
    if not key or not key[0].startswith(os.environ.get("SRCFILE", "cs_shortest2")):
        continue
    for nm, lo, hi_ in phases:
        if lo <= key[1] <= hi_:
            d = tot[nm]
            d["samples"] = d.get("samples", 0) + s
            d["inst"] = d.get("inst", 0) + float(r[ci["Instructions Executed"]] or 0)
            for st in stalls:
                d[st] = d.get(st, 0) + float(r[ci[st]] or 0)
for nm, _, _ in phases:
    d = tot[nm]
    if not d:
        continue
    top = sorted(((v, k) for k, v in d.items() if k.startswith("stall_")), reverse=True)[:6]
    print(f"{nm:8s} samples {d['samples'] / allsum * 100:5.1f}%  inst {d['inst']:.3e} | " + "  ".join(f"{k[6:]} {v / d['samples'] * 100:.0f}%" for v, k in top))
