"""BASELINE config 2: G=1024, 20 iterations, 1000 steps on one B200 -- this library against the
reference's own CUDA kernels rebuilt for sm_100a (oracle/_ref/refgpu_*, built from the reference's
sources by oracle/Makefile).  Reference per-step time = (T(201 steps) - T(1 step)) / 200 from the
program's own wall-clock print, which removes its one-time init/H2D; "asis" keeps the reference's
per-step host zeroing + 3 uploads, "resident" drops them (kernels only).
Also G=8192/K=40 with (T(6) - T(1)) / 5.    usage: config2_race.py [out.json]"""
import json, os, re, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fluidsimulationcuda_b200 import solver as SF

REF = os.path.join(ROOT, "oracle", "_ref")
BLOCKS = {"LOOPUNROLLED-Interleaved": ("16", "16"), "FluidParallelBlockPerElement-Naive": ("32", "16"),
          "FluidParallelBlockPerElement-SM": ("32", "16")}     # the block shapes of the reference's report.txt


def ref_elapsed(name, N, K, steps, mode):
    exe = os.path.join(REF, f"refgpu_{name}_N{N}_K{K}_S{steps}_{mode}")
    if not os.path.exists(exe):
        return None
    best = None
    for _ in range(3):
        out = subprocess.run([exe, *BLOCKS[name]], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600).stdout
        m = re.search(r"elapsed ([0-9.]+) sec", out)
        if not m:
            return None
        t = float(m.group(1))
        best = t if best is None else min(best, t)
    return best


def ours(N, K, steps, warm=20):
    s = SF.StableFluids(N)
    f = [s.new_field() for _ in range(6)]
    s.init_synthetic(1, *f)
    for i in range(warm):
        s.init_sources(10 + i, f[1], f[3], f[5]); s.step(*f, 0.0025, 0.1, 0.016, K)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = s.launch_count
    a.record()
    for i in range(steps):
        s.init_sources(100 + i, f[1], f[3], f[5]); s.step(*f, 0.0025, 0.1, 0.016, K)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps, (s.launch_count - n0) / steps


res = {"gpu": torch.cuda.get_device_name(0), "rows": []}
for (N, K, s_hi, ours_steps) in ((1022, 20, 201, 1000), (8190, 40, 6, 20)):
    ms, launches = ours(N, K, ours_steps)
    row = {"G": N + 2, "iters": K, "ours_ms_per_step": ms, "ours_kernel_launches_per_step": launches, "reference": {}}
    for name in BLOCKS:
        for mode in ("asis", "resident"):
            t1, th = ref_elapsed(name, N, K, 1, mode), ref_elapsed(name, N, K, s_hi, mode)
            if t1 is None or th is None:
                continue
            per = (th - t1) / (s_hi - 1) * 1e3
            row["reference"][f"{name} ({mode})"] = {"ms_per_step": per, "speedup_of_ours": per / ms}
    res["rows"].append(row)
    print(json.dumps(row), flush=True)
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
