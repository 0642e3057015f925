"""Turn the raw files of scripts/gpu_evidence.sh (gpurun_out/<tag>_*) into the tracked summaries under profiles/.
usage: python scripts/make_profiles.py <tag>   (the in-tree .so must be the build that was profiled)"""
import json
import os
import shutil
import subprocess
import sys

tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, dst = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
KERNELS = {"shortest3": ("cs_k_shortest3", "cs_k_shortest3ILi3", "shortest"), "segment3": ("cs_k_segment3", "cs_k_segment3ILi3", "segment"),
           "simplest": ("cs_k_simplest", "cs_k_simplestILi2", "simplest")}
traffic = {"note": "dram__bytes_read.sum + dram__bytes_write.sum per launch of each dominant kernel, from one ncu --set full "
                   "capture of the bench command (bench.py reads kernels[<name>].dram_bytes_per_launch)", "kernels": {}}
for short, (kname, section, fn) in KERNELS.items():
    rep = os.path.join(src, f"{tag}_{short}.ncu-rep")
    if not os.path.exists(rep):
        continue
    summ = subprocess.run([sys.executable, os.path.join(root, "scripts", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    env = dict(os.environ, SECTION=section)
    lines = subprocess.run([sys.executable, os.path.join(root, "scripts", "ncu_lines.py"), rep, kname, "30"], capture_output=True, text=True, env=env).stdout
    with open(os.path.join(dst, f"{tag}_ncu_{short}.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on -k regex:{kname} -s 1 -c 1 python bench.py --function {fn} --steps 1 --warmup 1 --no-cpu\n")
        f.write("# one launch of the bench workload (see profiles/%s_bench_%s.json for the configuration)\n" % (tag, fn))
        f.write(summ)
        f.write("\n# hottest source lines (scripts/ncu_lines.py)\n")
        f.write(lines)
    rd = wr = None
    for ln in summ.splitlines():
        p = ln.split(",")
        if p[0] == "dram__bytes_read.sum":
            rd = float(p[2]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[p[1]]
        if p[0] == "dram__bytes_write.sum":
            wr = float(p[2]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[p[1]]
    if rd is not None and wr is not None:
        traffic["kernels"][kname] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
                                     "profile": f"profiles/{tag}_ncu_{short}.txt"}
json.dump(traffic, open(os.path.join(dst, "traffic.json"), "w"), indent=1)
for name in os.listdir(src):
    if not name.startswith(tag + "_"):
        continue
    if name.endswith(".json") and "_bench_" in name:
        txt = open(os.path.join(src, name)).read().strip().splitlines()
        if txt:
            open(os.path.join(dst, name), "w").write(txt[-1] + "\n")
    elif name.endswith(".csv") and "_launches_" in name:
        rows = [r for r in open(os.path.join(src, name)) if r.startswith('"') or r.startswith("==PROF==") is False]
        open(os.path.join(dst, name), "w").write("".join(r for r in rows if "gpu__time_duration" in r or r.startswith('"ID"')))
    elif "_sanitizer_" in name:
        keep = [l for l in open(os.path.join(src, name), errors="replace") if "ERROR SUMMARY" in l or " passed" in l or " failed" in l or "=========" in l and "Error" in l]
        with open(os.path.join(dst, f"{tag}_sanitizer.txt"), "a") as f:
            f.write(f"# {name}\n" + "".join(keep[-6:]))
    elif name.endswith("_tests.log"):
        keep = [l for l in open(os.path.join(src, name)) if l.startswith("===") or " passed" in l or " failed" in l]
        open(os.path.join(dst, f"{tag}_gpu_tests.txt"), "w").write("".join(keep))
print("profiles written for", tag)
