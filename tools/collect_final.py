"""Copy the outputs of tools/final_1gpu.sh (gpurun_out/final_*) into profiles/r02/final_<commit>_* and restamp
profiles/dram_traffic.json from the strict T = 7 capture (developer tool; no GPU needed)."""
import csv, glob, json, os, re, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "gpurun_out")
commit = open(os.path.join(out, "final_commit.txt")).read().strip()
dst = os.path.join(ROOT, "profiles", "r02")
os.makedirs(dst, exist_ok=True)
for p in sorted(glob.glob(os.path.join(out, "final_*"))):
    b = os.path.basename(p)
    if b.endswith((".ncu-rep", ".err")) or b == "final_commit.txt" or (b.startswith("final_prof_") and b.endswith(".log")):
        continue
    name = b[len("final_"):]
    name = re.sub(r"^prof_", "ncu_", name).replace("_int_", "_")
    shutil.copy(p, os.path.join(dst, f"final_{commit}_{name}"))
    print("->", f"profiles/r02/final_{commit}_{name}")
src = os.path.join(dst, f"final_{commit}_ncu_jacobi_stream_kernel_7_0_0.csv")
rows = list(csv.reader(open(src)))
h, r = rows[0], rows[2]
d = dict(zip(h, r))
rd, wr = float(d["dram__bytes_read.sum"]) * 1e6, float(d["dram__bytes_write.sum"]) * 1e6
unit = dict(zip(h, rows[1]))
assert unit["dram__bytes_read.sum"] == "Mbyte", unit["dram__bytes_read.sum"]
j = {"8192": rd + wr, "commit": commit,
     "kernel": "jacobi_stream_kernel<7, STRICT, 0> (one T = 7 launch of a viscosity solve, 3rd step of tools/prof_step.py)",
     "dram_read_bytes": rd, "dram_write_bytes": wr, "ncu_duration_us": float(d["gpu__time_duration.sum"]),
     "source": f"profiles/r02/final_{commit}_ncu_jacobi_stream_kernel_7_0_0.csv (ncu --set full --clock-control none)"}
json.dump(j, open(os.path.join(ROOT, "profiles", "dram_traffic.json"), "w"), indent=1)
print(json.dumps(j))
