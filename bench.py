#!/usr/bin/env python
"""Benchmark of the stable-fluids time step (vel_step + dens_step) -- the BASELINE.json metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--grid G] [--iters K]

One "step" = one pass of the reference's main-loop body (FluidSequential.c:289-312): refresh the
three source fields, vel_step, dens_step, on synthetic fields of the named size.
  N = 1 : G = 8192 (N = 8190 interior), 40 Jacobi iterations per lin_solve -- BASELINE configs[2],
          the configuration the metric is quoted on.
  N > 1 : G = 32768, 40 iterations, row slabs over the N ranks with neighbour halo exchange
          (BASELINE configs[3]); launched by torchrun, one rank per GPU.  The N = 1 line carries the
          single-GPU time of this problem as `scaling_base`, so the strong-scaling base is measured
          in the same run as the G = 8192 headline.
metric = Jacobi cell-updates/s = 5 * iters * N^2 * steps / time (five lin_solves per step); the
whole step (add_source, advect, divergence, gradient subtract, set_bnd) is inside the timed region.

Rank 0 prints ONE JSON line.  --impl reference times the reference's own sequential CPU code
(oracle/_ref, built from the reference's source) on the host cores instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DT, VIS, DIFF = 0.016, 0.0025, 0.1        # the reference's literals (FluidSequential.c:7-9)
HBM_FALLBACK_GBS = 6650.0                  # /opt/skills/guides/B200_PROFILING.md fallback


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx = max(mx, float(s[1]))
                for n, val in zip(names, s[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        # "under load": the upper half of the samples (the sampler also sees the idle lead-in)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own sequential CPU implementation, one thread (it has no threading).
    Each "step" is a bounded sample of the workload: one dens_step (add_source + one `iters`-sweep
    lin_solve + advect = 1/5 of a step's Jacobi work) at the full grid size."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle.pyoracle import Oracle, ReferenceSeq
    G, K = args.grid, args.iters
    N = G - 2
    if ReferenceSeq.available(N, K):
        ref, kind = ReferenceSeq(N, K), "reference"
        step = lambda x, x0, u, v: ref.dens_step(x, x0, u, v, DIFF)
    else:   # no build of the reference for this (N, K): the pinned restatement
        orc, kind = Oracle(), "port"
        step = lambda x, x0, u, v: orc.dens_step(N, x, x0, u, v, DIFF, DT, K)
    o = Oracle(threads=True)
    f = o.init_synthetic(N, 1)
    x, u, v = f["dens"], f["u_prev"] * np.float32(0.5), f["v_prev"] * np.float32(0.5)
    times = []
    for i in range(args.warmup + args.steps):
        x0 = f["dens_prev"].copy()           # live sources every step (early-step values: no subnormals)
        t = time.perf_counter(); step(x, x0, u, v); dt_ = time.perf_counter() - t
        if i >= args.warmup:
            times.append(dt_)
    total = sum(times)
    value = K * N * N * len(times) / total
    sample = f"dens_step (add_source + {K}-sweep lin_solve + advect) at G={G}: 1/5 of a step's Jacobi work per sample"
    print(json.dumps({
        "impl": "reference", "metric": "jacobi_cell_updates_per_s", "value": value, "unit": "cell-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"stable-fluids step (vel_step + dens_step) G={G} (N={N}), {K} Jacobi iterations per lin_solve, "
                               f"the reference's sequential CPU path, sampled as one dens_step per step",
                   "grid": G, "iters": K, "parallelism": "1 host thread (the reference has no threading)"},
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def cpu_baseline(G, K):
    """Reference sequential path (and its SIMD-SSE program) on this box's host cores, bounded sample."""
    import numpy as np
    from oracle.pyoracle import Oracle, ReferenceSeq
    N = G - 2
    out = {}
    try:
        o = Oracle(threads=True)
        f = o.init_synthetic(N, 1)
        x, x0 = f["dens"], f["dens_prev"]
        u, v = f["u_prev"] * np.float32(0.5), f["v_prev"] * np.float32(0.5)
        if ReferenceSeq.available(N, K):
            ref, kind = ReferenceSeq(N, K), "reference"
            t = time.perf_counter(); ref.dens_step(x, x0, u, v, DIFF); dt_ = time.perf_counter() - t
        else:
            kind = "port"
            orc = Oracle()
            t = time.perf_counter(); orc.dens_step(N, x, x0, u, v, DIFF, DT, K); dt_ = time.perf_counter() - t
        out = {"value": K * N * N / dt_, "unit": "cell-updates/s", "cores": 1, "kind": kind,
               "sample": f"one dens_step (add_source + {K}-sweep lin_solve + advect) at G={G}, {dt_:.2f} s, "
                         f"sequential reference, 1 thread (the reference has no threading)",
               "host_cpus": os.cpu_count()}
    except Exception as e:   # the baseline is reported, never fatal
        out = {"value": None, "unit": "cell-updates/s", "cores": 1, "kind": "unavailable", "sample": repr(e)}
    try:   # clearly labelled extra (not a reference path): the pinned restatement with OpenMP on every host core
        f = o.init_synthetic(N, 1)
        x, x0 = f["dens"], f["dens_prev"]
        u, v = f["u_prev"] * np.float32(0.5), f["v_prev"] * np.float32(0.5)
        o.dens_step(N, x, x0, u, v, DIFF, DT, K)            # warm the thread pool and the pages
        x0 = o.init_synthetic(N, 1)["dens_prev"]
        t = time.perf_counter(); o.dens_step(N, x, x0, u, v, DIFF, DT, K); dt_ = time.perf_counter() - t
        out["omp_port"] = {"value": K * N * N / dt_, "unit": "cell-updates/s", "cores": os.cpu_count(), "kind": "port",
                           "sample": f"one dens_step at G={G}, {dt_:.2f} s, oracle restatement built with -fopenmp "
                                     f"(bit-identical to the sequential reference; the reference itself has no threading)"}
    except Exception as e:
        out["omp_port"] = {"value": None, "sample": repr(e)}
    simd = os.path.join(ROOT, "oracle", "_ref", f"ref_simd_N{N}_K{K}")
    if os.path.exists(simd):
        try:
            t = time.perf_counter()
            r = subprocess.run([simd], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=120)
            dt_ = time.perf_counter() - t
            el = [l for l in r.stdout.splitlines() if "elapsed" in l]
            if el:
                dt_ = float(el[-1].split("elapsed")[1].split()[0])
            out["simd_sse"] = {"value": 5 * K * N * N / dt_, "unit": "cell-updates/s", "cores": 1,
                               "sample": f"the reference's SIMD-SSE program, one full step incl. its rand() init at G={G}, "
                                         f"{dt_:.2f} s (timing only: its interior lanes are numerically wrong as shipped)"}
        except Exception as e:
            out["simd_sse"] = {"value": None, "sample": repr(e)}
    return out


def scaling_base(Gs, K, steps=5, timeout_s=240):
    """One GPU on the N > 1 workload (G = 32768: 7 fields of 4 GiB), same step and timing as the headline.  Runs as a
    child process (this file with --grid Gs --skip-extras) under a timeout, so that nothing it does -- an allocation
    failure, a CUDA error, a hang -- can cost the headline line."""
    cmd = [sys.executable, os.path.abspath(__file__), "--gpus", "1", "--grid", str(Gs), "--iters", str(K), "--steps", str(steps),
           "--warmup", "3", "--skip-extras", "--scaling-base", "0"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    try:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout_s, env=env)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"grid": Gs, "value": None, "error": (r.stderr or r.stdout)[-300:]}
        j = json.loads(lines[-1])
        return {"grid": Gs, "iters": K, "n_gpus": 1, "steps": steps, "ms_per_step": j["ms_per_step"], "value": j["value"],
                "unit": j["unit"],
                "note": f"the N > 1 lines divide this G={Gs} problem into row slabs: parallel efficiency at p GPUs = "
                        "value_p / (p * this value)"}
    except Exception as e:
        return {"grid": Gs, "value": None, "error": repr(e)}


def child_json(extra_args, timeout_s):
    """Run this file as a child process (so that nothing the extra measurement does -- an allocation failure, a CUDA error,
    a hang -- can cost the headline line) and return the last JSON object it printed."""
    cmd = [sys.executable, os.path.abspath(__file__)] + extra_args
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "GROUP_RANK",
                                                             "ROLE_RANK", "TORCHELASTIC_RUN_ID")}
    try:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout_s, env=env)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"error": (r.stderr or r.stdout)[-300:]}
        return json.loads(lines[-1])
    except Exception as e:
        return {"error": repr(e)}


REFGPU_BLOCKS = {"LOOPUNROLLED-Interleaved": ("16", "16"), "FluidParallelBlockPerElement-Naive": ("32", "16"),
                 "FluidParallelBlockPerElement-SM": ("32", "16")}     # the block shapes of the reference's report.txt


def refgpu_elapsed(name, N, K, steps, mode, repeats=2):
    """Wall clock the reference's own CUDA program prints ("elapsed ... sec"), best of `repeats` runs; None if not built.
    oracle/_ref/refgpu_* are built from the reference's sources by `make -C oracle refgpu` (sed sets hN, the iteration
    count and the step count; *_resident drops the loop's per-step host zeroing + three uploads)."""
    import re
    exe = os.path.join(ROOT, "oracle", "_ref", f"refgpu_{name}_N{N}_K{K}_S{steps}_{mode}")
    if not os.path.exists(exe):
        return None
    best = None
    for _ in range(repeats):
        out = subprocess.run([exe, *REFGPU_BLOCKS[name]], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120).stdout
        m = re.search(r"elapsed ([0-9.]+) sec", out)
        if not m:
            return None
        best = float(m.group(1)) if best is None else min(best, float(m.group(1)))
    return best


def run_config2_child(args):
    """BASELINE configs[1]: G = 1024, 20 iterations, 1000 steps on one B200 against the reference's smPar / optPar / naivePar
    CUDA programs rebuilt for sm_100a and run on this same GPU; plus the reference's fastest program at the headline size.
    Reference per-step time = (T(many steps) - T(1 step)) / (many - 1) of the program's own wall-clock print, which removes
    its one-time rand() init and uploads (the programs have no per-step timer)."""
    import torch
    from fluidsimulationcuda_b200 import solver as SF
    torch.cuda.set_device(0)
    out = {"workload": "G=1024 (N=1022), 20 Jacobi iterations per lin_solve, 1000 steps, device-resident loop (sf_run_steps, "
                       "synthetic source refresh every step); the working set (9 fields x 4 MiB) is L2-resident: launch- and "
                       "L2-bound, not HBM-bound"}
    N, K, steps = 1022, 20, 1000
    s = SF.StableFluids(N)
    f = [s.new_field() for _ in range(6)]
    s.init_synthetic(1, *f)
    s.run_steps(*f, VIS, DIFF, DT, K, 20, SF.SOURCES_SYNTHETIC, 10)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = s.launch_count
    a.record(); s.run_steps(*f, VIS, DIFF, DT, K, steps, SF.SOURCES_SYNTHETIC, 100); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    out.update({"grid": 1024, "iters": K, "steps": steps, "ms_per_step": ms, "kernel_launches_per_step": (s.launch_count - n0) / steps,
                "value": 5.0 * K * N * N / (ms * 1e-3), "unit": "cell-updates/s", "reference_cuda_programs_same_gpu": {}})
    s.close(); del f
    ref = out["reference_cuda_programs_same_gpu"]
    for name in REFGPU_BLOCKS:
        for mode in ("resident", "asis"):
            t1, th = refgpu_elapsed(name, N, K, 1, mode), refgpu_elapsed(name, N, K, 201, mode)
            if t1 is None or th is None:
                continue
            per = (th - t1) / 200 * 1e3
            ref[f"{name} ({'kernels only' if mode == 'resident' else 'as shipped: per-step host zeroing + 3 uploads'})"] = {
                "ms_per_step": per, "speedup_of_ours": per / ms}
    if not ref:
        out["reference_cuda_programs_same_gpu"] = "not built (oracle/_ref/refgpu_*: `make -C oracle refgpu` where /root/reference exists)"
    # the reference's fastest program (report.txt:45-46) at the headline size, against the headline time passed in
    big = {}
    for mode in ("resident", "asis"):
        t1, th = refgpu_elapsed("LOOPUNROLLED-Interleaved", 8190, 40, 1, mode, 1), refgpu_elapsed("LOOPUNROLLED-Interleaved", 8190, 40, 6, mode, 1)
        if t1 is None or th is None:
            continue
        per = (th - t1) / 5 * 1e3
        big[f"LOOPUNROLLED-Interleaved ({'kernels only' if mode == 'resident' else 'as shipped'})"] = {
            "ms_per_step": per, "speedup_of_ours": (per / args.headline_ms) if args.headline_ms > 0 else None}
    if big:
        out["headline_size_G8192_K40"] = {"ours_ms_per_step": args.headline_ms, "reference_cuda_programs_same_gpu": big}
    print(json.dumps(out))


def reference_schedule(SF, torch, N, K, steps=50):
    """The reference loop's own source rule (FluidSequential.c:298-302: sources act in step 0 only, the three *_prev fields are
    zeroed before every later step).  Density then decays into the 1e-30 .. subnormal range where the exact division takes
    its guarded (binary64) ticks; the headline instead refreshes the sources every step.  ms per step, zeroing included."""
    s = SF.StableFluids(N)
    f = [s.new_field() for _ in range(6)]
    s.init_synthetic(1, *f)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for k in range(steps):
        if k > 0:
            for t in (f[1], f[3], f[5]):
                t.zero_()
        s.step(*f, VIS, DIFF, DT, K)
        ev[k + 1].record()
    torch.cuda.synchronize()
    ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
    dmax = s.reduce_max_abs(f[0])
    s.close()
    pick = [k for k in (1, 2, 5, 10, 25, 50) if k <= steps]
    return {"schedule": "sources in step 0 only, *_prev fields zeroed before every later step (FluidSequential.c:298-302)",
            "grid": N + 2, "iters": K, "ms_at_step": {str(k): ms[k - 1] for k in pick}, "ms_mean_steps_3_to_end": sum(ms[2:]) / len(ms[2:]),
            "max_dens_after": dmax,
            "note": "steps 1-2 include the one-time direct run and graph capture; compare with ms_per_step of the headline "
                    "(sources refreshed every step)"}


def bind_to_gpu_numa_node(torch, local):
    """Pin this rank's host threads (and therefore the first-touch placement of the pinned host buffers it allocates next)
    to the NUMA node its GPU hangs off.  Returns a description for the JSON line; never fatal."""
    bus = None
    try:
        pr = torch.cuda.get_device_properties(local)
        if all(hasattr(pr, a) for a in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bus = f"{int(pr.pci_domain_id):04x}:{int(pr.pci_bus_id):02x}:{int(pr.pci_device_id):02x}.0"
    except Exception:
        bus = None
    try:
        if bus is None:
            q = subprocess.run(["nvidia-smi", f"--id={local}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                               stdout=subprocess.PIPE, text=True, timeout=20).stdout.strip()
            bus = q
        dom = bus.lower()
        if len(dom.split(":")[0]) == 8:      # nvidia-smi prints an 8-digit domain, sysfs uses 4
            dom = dom[4:]
        node = int(open(f"/sys/bus/pci/devices/{dom}/numa_node").read().strip())
        if node < 0:
            return {"numa_node": None, "note": "no NUMA affinity reported for this GPU"}
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed), "pci": dom}
    except Exception as e:
        return {"numa_node": None, "note": repr(e)[:120]}


def slab_parity(torch, dist, rank, world, G=2048, K=20, steps=2):
    """Correctness bit of a multi-GPU line: the SAME peer-slab path (CUDA IPC mappings, fused strip exchange, device-side
    barriers, one graph per GPU) on a problem the CPU oracle finishes in a second, every field of every rank compared
    BITWISE with the oracle after `steps` steps (reference schedule: sources zeroed after step 0)."""
    import numpy as np
    from fluidsimulationcuda_b200.slab import PeerSlabSolver
    N = G - 2
    sim = PeerSlabSolver(N, rank, world, iters=K)
    sim.connect_dist()
    sim.init_synthetic(3)
    for st in range(steps):
        if st > 0:
            sim.zero_sources()
        sim.step(None, VIS, DIFF, DT)
    sim.status()
    torch.cuda.synchronize()
    from oracle.pyoracle import Oracle          # the checker, on every rank's own rows (never timed, never the product path)
    o = Oracle()
    w = o.init_synthetic(N, 3)
    o.run_steps(N, steps, w, VIS, DIFF, DT, K)
    bad = 0
    for k in sim.names:
        got = sim.owned(sim.f[k]).cpu().numpy()
        want = w[k][sim.row_lo:sim.row_hi]
        bad += int((got.view(np.uint32) != want.view(np.uint32)).sum())
    t = torch.tensor([bad], device="cuda", dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    sim.close()
    return {"ok": int(t.item()) == 0, "mismatching_cells": int(t.item()), "grid": G, "iters": K, "steps": steps, "fields": 6,
            "against": "CPU oracle (restatement pinned to the reference's sequential build), bitwise, every rank's owned rows"}


def config5(torch, dist, rank, world, local, G=16384, K=200, steps=3):
    from fluidsimulationcuda_b200.slab import PeerSlabSolver
    from fluidsimulationcuda_b200 import solver as SF
    N = G - 2
    sim = PeerSlabSolver(N, rank, world, iters=K, arithmetic=SF.STRICT)
    sim.connect_dist()
    sim.init_synthetic(1)
    for i in range(2):
        sim.step(100 + i, VIS, DIFF, DT)
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(sim.stream)
    for i in range(steps):
        sim.step(1000 + i, VIS, DIFF, DT)
    b.record(sim.stream)
    sim.status()
    torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([a.elapsed_time(b) / steps], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    sim.close()
    torch.cuda.empty_cache()
    base = None
    if rank == 0:
        j = child_json(["--gpus", "1", "--grid", str(G), "--iters", str(K), "--steps", str(steps), "--warmup", "3", "--skip-extras",
                        "--scaling-base", "0"], 240)
        base = {"n_gpus": 1, "ms_per_step": j.get("ms_per_step"), "value": j.get("value"), "error": j.get("error")}
    dist.barrier()
    out = {"workload": f"G={G} (N={N}), {K} Jacobi iterations per lin_solve, row slabs x{world} (strong scaling of one problem)",
           "grid": G, "iters": K, "n_gpus": world, "steps": steps, "ms_per_step": ms, "value": 5.0 * K * N * N / (ms * 1e-3),
           "unit": "cell-updates/s", "same_run_one_gpu": base}
    if base and base.get("ms_per_step"):
        # halo-exchange time hidden vs exposed.  The halo rows themselves travel inside the Jacobi
        # launches (peer stores by the strip warps while the interior warps compute); what the partition costs on top of the
        # ideal t_1 / p is everything else together: the strip passes in front of interior items, the right-hand-side pushes,
        # the neighbour barriers, the extra launches, and the one-GPU path's overlapped solves that slabs do not have
        ideal = base["ms_per_step"] / world
        out["halo_exchange"] = {"ideal_ms_per_step": ideal, "exposed_ms_per_step": ms - ideal,
                                "note": "ideal = same-run one-GPU time / n_gpus; exposed = this run's time minus ideal (all partition "
                                        "overheads together; the halo rows' transfer itself is inside the Jacobi launches)"}
    return out


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from fluidsimulationcuda_b200 import build
    from fluidsimulationcuda_b200 import solver as SF

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    build.build()
    G, K = args.grid, args.iters
    N = G - 2
    cells = G * G
    peak, peak_how = measured_peak()

    parity = numa = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
        from fluidsimulationcuda_b200.slab import SlabSolver, TorchDistComm, PeerSlabSolver
        if args.slab_comm == "peer" and not args.skip_extras:
            numa = bind_to_gpu_numa_node(torch, local)
            try:      # correctness bit of this line, before anything is timed
                parity = slab_parity(torch, dist, rank, world)
            except Exception as e:
                parity = {"ok": False, "error": repr(e)[:300]}
        if args.slab_comm == "peer":
            # product path: neighbours' slabs mapped over NVLink (CUDA IPC), halo pushes fused into the
            # Jacobi strips, device-side neighbour barriers, one CUDA-graph replay per step per GPU
            sim = PeerSlabSolver(N, rank, world, iters=K, arithmetic=SF.STRICT)
            sim.connect_dist()
        else:
            sim = SlabSolver(N, rank, world, iters=K, arithmetic=SF.STRICT, comm=TorchDistComm())
        sim.init_synthetic(1)
        step = lambda seed: sim.step(seed, VIS, DIFF, DT)
        sync = lambda: (torch.cuda.synchronize(), dist.barrier())
        launches = lambda: sim.launch_count
    else:
        s = SF.StableFluids(N)
        f = [s.new_field() for _ in range(6)]       # dens, dens_prev, u, u_prev, v, v_prev
        s.init_synthetic(1, *f)

        def step(seed):
            s.init_sources(seed, f[1], f[3], f[5])  # the loop's per-step source refresh, on the device
            s.step(*f, VIS, DIFF, DT, K)
        sync = torch.cuda.synchronize
        launches = lambda: s.launch_count

    for i in range(args.warmup):
        step(100 + i)
    sync()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start(); time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # events go on the stream the kernels are launched on (a peer slab owns its stream)
    tstream = sim.stream if (world > 1 and args.slab_comm == "peer") else torch.cuda.current_stream()
    n0 = launches()
    sync()
    ev0.record(tstream)
    for i in range(args.steps):
        step(1000 + i)
    if world > 1:
        sim.check_reach()      # deferred advection-reach verification of the last steps (synchronises)
    ev1.record(tstream)
    sync()
    ms = ev0.elapsed_time(ev1)
    n_launch = launches() - n0
    clocks = sampler.finish() if sampler else None
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tl = torch.tensor([n_launch], device="cuda", dtype=torch.int64)
        dist.all_reduce(tl, op=dist.ReduceOp.SUM)
        n_launch = int(tl.item())
    ms_step = ms / args.steps
    value = 5.0 * K * N * N / (ms_step * 1e-3)
    step_bytes = (60.0 * K + 148.0) * cells          # algorithmic bytes per step
    eff_gbs = step_bytes / (ms_step * 1e-3) / 1e9

    out = {
        "metric": "jacobi_cell_updates_per_s", "value": value, "unit": "cell-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"stable-fluids step (vel_step + dens_step) G={G} (N={N}), {K} Jacobi iterations per lin_solve, "
                               f"STRICT arithmetic (bit-identical to the reference's sequential path)",
                   "grid": G, "iters": K, "parallelism": f"row slabs x{world}" if world > 1 else "single GPU",
                   "l2": "every field (G^2*4 B = %.0f MiB) is larger than the 126 MB L2; 9 fields live" % (cells * 4 / 2**20),
                   "per_step": ("device-side source refresh + vel_step + dens_step, replayed from a CUDA graph whose independent lin_solves (u, v viscosity; density diffusion) are parallel branches" if world == 1 else
                                "device-side source refresh + vel_step + dens_step per slab, one CUDA-graph replay per GPU; halo rows "
                                "stored into the neighbour's ghost rows over NVLink by the Jacobi boundary-strip kernels (peer memory) "
                                "while the interior launch runs; advect gathers through the peer mapping; device-side neighbour barriers"
                                if args.slab_comm == "peer" else
                                "device-side source refresh + vel_step + dens_step per slab; NCCL neighbour halo exchange "
                                "overlapped with the interior Jacobi launch; MAX all-reduce for the advection reach")},
        "full_step_cells_per_s": cells / (ms_step * 1e-3),
        "effective_hbm_gbs": eff_gbs, "effective_hbm_frac_of_measured_peak": eff_gbs / (peak * world),
        "gpu_launches": n_launch, "clocks": clocks,
    }
    if parity is not None:
        out["parity"] = parity
    if world > 1:
        out["scaling_note"] = (f"strong scaling of the G={G} problem over {world} GPUs; the N=1 bench line runs G=8192 (the size the "
                               "metric is quoted on), so compare with the single-GPU time of THIS problem: the `scaling_base` object "
                               "of the N=1 line (measured in that run; profiles/r02/final_c2a13ce_bench.json: 113.4 ms/step at "
                               "G=32768, K=40), and `extra.config5.same_run_one_gpu` for G=16384, K=200")

    if world == 1 and not args.skip_extras:
        # ---- roofline of the dominant kernel: jacobi_stream_kernel, timed per lin_solve with CUDA events on
        # the context's stream (graphs off for this instrumented pass; same kernels, same launch plan)
        sr = SF.StableFluids(N, use_graph=False)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        import numpy as np
        f32 = np.float32
        al = f32(DT) * f32(VIS); al = al * f32(N); al = al * f32(N); be = f32(1) + f32(4) * al
        n0 = sr.launch_count
        tot_ms, sweeps = 0.0, 0
        kind_ms = {"strict": [], "pressure": []}
        for rep in range(3):
            for (kind, b_, x, x0, alpha, beta) in (("strict", 1, f[3], f[2], float(al), float(be)), ("pressure", 0, f[1], f[0], 1.0, 4.0)):
                a.record(); sr.diffuse(b_, x, x0, alpha, beta, K); b.record(); torch.cuda.synchronize()
                if rep > 0:
                    tot_ms += a.elapsed_time(b); sweeps += K
                    kind_ms[kind].append(a.elapsed_time(b))
        jl = (sr.launch_count - n0) // 3 // 2
        alg_bytes_per_launch = 12.0 * cells * K / jl
        achieved = 12.0 * cells * sweeps / (tot_ms * 1e-3) / 1e9
        out["roofline"] = {
            "kernel": "jacobi_stream_kernel (temporally blocked lin_solve)", "bound": "hbm",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_how,
            "traffic": PROFILED_DRAM_BYTES_PER_LAUNCH.get(G), "traffic_capture": PROFILED_DRAM_META or None,
            "algorithmic_bytes_per_launch": alg_bytes_per_launch, "launches_per_lin_solve": jl,
            "avg_launch_ms": tot_ms / (sweeps / K) / jl,
            # the same per kind of solve (what an ncu capture of one launch is to be compared with: a strict T = 7 launch is
            # 7 / (K / launches) of the strict average)
            "by_kind": {k: {"ms_per_lin_solve": sum(v) / len(v), "launches": jl, "avg_launch_ms": sum(v) / len(v) / jl,
                            "achieved_gbs": 12.0 * cells * K / (sum(v) / len(v) * 1e-3) / 1e9} for k, v in kind_ms.items() if v},
            "note": "algorithmic bytes = 12 B per cell per sweep (read x, read x0, write x'); one launch fuses "
                    f"{K}/{jl} sweeps, so achieved exceeds the DRAM peak by design; traffic = ncu dram bytes per launch",
        }
        tr = out["roofline"]["traffic"]
        if tr:   # the same launch seen from DRAM: profiled bytes per launch / event-timed launch duration
            dram = tr / (out["roofline"]["avg_launch_ms"] * 1e-3) / 1e9
            out["roofline"]["dram_gbs"] = dram
            out["roofline"]["dram_frac"] = dram / peak
            out["roofline"]["dram_bytes_per_cell_update"] = tr / (cells * (K / jl))     # 12 without temporal blocking
        sr.close()
        # ---- end to end through the host-buffer entry point (sf_step_host): pinned host fields,
        # H2D of all six fields and D2H of dens,u,v inside the timed region, every step
        try:
            hf = [torch.empty((G, G), dtype=torch.float32).pin_memory() for _ in range(6)]
            for h, d in zip(hf, f):
                h.copy_(d)
            for _ in range(2):
                s.step_host(*hf, VIS, DIFF, DT, K)
            torch.cuda.synchronize()
            n_e2e = max(3, min(args.steps, 5))
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                s.step_host(*hf, VIS, DIFF, DT, K)
            torch.cuda.synchronize()
            dt_ = (time.perf_counter() - t0) / n_e2e
            out["e2e"] = {"value": 5.0 * K * N * N / dt_, "unit": "cell-updates/s", "ms_per_step": dt_ * 1e3,
                          "h2d_bytes_per_step": 6 * cells * 4, "d2h_bytes_per_step": 3 * cells * 4,
                          "api": "sf_step_host (C ABI): six pinned host fields in, dens/u/v out, copies overlapped with compute"}
            del hf
        except Exception as e:
            out["e2e"] = {"value": None, "unit": "cell-updates/s", "error": repr(e)}
        out["cpu_baseline"] = cpu_baseline(G, K)
        # ---- the single-GPU time of the problem the N > 1 lines run (G = 32768), measured in this same run, so
        # that the strong-scaling lines have their own base next to the G = 8192 headline (reported, never fatal)
        if G == 8192 and args.scaling_base:
            out["scaling_base"] = scaling_base(args.scaling_base, K)
        # ---- the other BASELINE configurations that fit one GPU, and the reference's own source schedule (reported, never fatal)
        out["extra"] = {}
        try:
            out["extra"]["reference_schedule"] = reference_schedule(SF, torch, N, K)
        except Exception as e:
            out["extra"]["reference_schedule"] = {"error": repr(e)[:300]}
        if G == 8192 and args.config2:
            f.clear()                               # (the child runs on the same GPU: give the memory back first)
            s.close()
            torch.cuda.empty_cache()
            out["extra"]["config2"] = child_json(["--child", "config2", "--headline-ms", repr(ms_step)], 240)
    elif world > 1 and args.slab_comm == "peer" and not args.skip_extras:
        # ---- roofline of the dominant kernel on a slab: the same two lin_solves as in the N = 1 line, called collectively
        # (every rank solves its slab; strip exchange and neighbour barriers included), CUDA events on the slab's stream,
        # max over ranks.  Per-GPU figures: a slab's algorithmic bytes against one GPU's measured peak.
        try:
            import numpy as np
            f32 = np.float32
            al = f32(DT) * f32(VIS); al = al * f32(N); al = al * f32(N); be = f32(1) + f32(4) * al
            jl = -(-K // 7)
            if (jl & 1) and jl + 1 <= K:
                jl += 1                          # plan_launches (sf_api.cu): an even number of launches of at most 7 sweeps
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            kind_ms = {"strict": [], "pressure": []}
            graphs = sim.ctx.get_option(SF.SF_OPT_USE_GRAPH)
            sim.ctx.set_option(SF.SF_OPT_USE_GRAPH, 0)      # direct launches: a capture between the events would be timed too
            for rep in range(3):
                for (kind, b_, x, x0, alpha, beta) in (("strict", 1, sim.f["u_prev"], sim.f["u"], float(al), float(be)),
                                                       ("pressure", 0, sim.f["dens_prev"], sim.f["dens"], 1.0, 4.0)):
                    sync()
                    a.record(sim.stream)
                    with torch.cuda.stream(sim.stream):
                        sim.ctx.diffuse(b_, x, x0, alpha, beta, K)
                    b.record(sim.stream)
                    sync()
                    if rep > 0:
                        kind_ms[kind].append(a.elapsed_time(b))
            sim.ctx.set_option(SF.SF_OPT_USE_GRAPH, graphs)
            sim.status()
            tk = torch.tensor([sum(kind_ms["strict"]) / 2, sum(kind_ms["pressure"]) / 2], device="cuda", dtype=torch.float64)
            dist.all_reduce(tk, op=dist.ReduceOp.MAX)
            ms_kind = {"strict": float(tk[0].item()), "pressure": float(tk[1].item())}
            slab_cells = cells / world
            tot_ms = ms_kind["strict"] + ms_kind["pressure"]
            achieved = 12.0 * slab_cells * 2 * K / (tot_ms * 1e-3) / 1e9
            out["roofline"] = {
                "kernel": "jacobi_stream_kernel with fused strip exchange (temporally blocked lin_solve on a row slab)", "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_how,
                "traffic": None, "per": "GPU (one slab's algorithmic bytes against one GPU's measured peak; max over ranks of the time)",
                "algorithmic_bytes_per_launch": 12.0 * slab_cells * K / jl, "launches_per_lin_solve": jl,
                "avg_launch_ms": tot_ms / 2 / jl,
                "by_kind": {k: {"ms_per_lin_solve": v, "launches": jl, "avg_launch_ms": v / jl,
                                "achieved_gbs": 12.0 * slab_cells * K / (v * 1e-3) / 1e9} for k, v in ms_kind.items()},
                "note": "a lin_solve here = halo exchange of the right-hand side + the blocked launches (strip warps push halo rows "
                        "to the neighbours over NVLink inside the kernel) + closing neighbour barrier; algorithmic bytes = 12 B per "
                        "cell per sweep, several sweeps per launch, so achieved exceeds the DRAM peak by design",
            }
        except Exception as e:
            out["roofline"] = {"error": repr(e)[:300]}
        # ---- end to end on N GPUs: every rank keeps its slab of the six fields in pinned host memory;
        # per step it uploads them, steps (collectively) and downloads dens, u, v -- all inside the timed region
        try:
            hf = sim.new_host_fields()
            for h, name in zip(hf, sim.names):
                h.copy_(sim.owned(sim.f[name]))
            torch.cuda.synchronize(); dist.barrier()
            sim.step_host(hf, VIS, DIFF, DT)
            torch.cuda.synchronize(); dist.barrier()
            n_e2e = 3
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                sim.step_host(hf, VIS, DIFF, DT)
            sim.status()
            dt_ = (time.perf_counter() - t0) / n_e2e
            tt = torch.tensor([dt_], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt_ = float(tt.item())
            out["e2e"] = {"value": 5.0 * K * N * N / dt_, "unit": "cell-updates/s", "ms_per_step": dt_ * 1e3,
                          "h2d_bytes_per_step": 6 * cells * 4, "d2h_bytes_per_step": 3 * cells * 4,
                          "api": "sf_step_host (C ABI) on connected peer slabs, one call per rank: uploads the rank's owned rows of the "
                                 "six pinned host fields, steps collectively, downloads dens/u/v (bytes are the sum over ranks; "
                                 "max over ranks of the wall time)",
                          "host_numa": numa}
            del hf
        except Exception as e:
            out["e2e"] = {"value": None, "unit": "cell-updates/s", "error": repr(e)}
        # ---- BASELINE configs[4]: pressure-projection stress, G = 16384, 200 iterations, strong scaling -- this run's
        # N-GPU time and, measured by rank 0 in a child process while the others wait, the 1-GPU time of the same problem
        try:
            out["extra"] = {"config5": config5(torch, dist, rank, world, local)}
        except Exception as e:
            out["extra"] = {"config5": {"error": repr(e)[:300]}}
    elif rank == 0:
        out["e2e"] = {"value": None, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                      "note": "not measured in this run (--skip-extras or --slab-comm nccl); see the N=1 line"}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ncu `dram__bytes_read.sum + dram__bytes_write.sum` per jacobi_stream_kernel launch, from the
# committed capture profiles/ (see profiles/README.md); None until a capture exists for that size.
PROFILED_DRAM_BYTES_PER_LAUNCH, PROFILED_DRAM_META = {}, {}
try:
    with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as _f:
        for _k, _v in json.load(_f).items():
            if _k.isdigit():
                PROFILED_DRAM_BYTES_PER_LAUNCH[int(_k)] = _v
            else:
                PROFILED_DRAM_META[_k] = _v      # commit and kernel the capture was taken on
except Exception:
    pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=0, help="full grid width G = N+2 (default 8192 at 1 GPU, 32768 at >1)")
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--slab-comm", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU halo traffic: peer = fused peer-memory pushes + device barriers (default), nccl = NCCL send/recv")
    ap.add_argument("--scaling-base", type=int, default=32768,
                    help="N=1 only: also time one GPU on this grid (the N>1 workload); 0 = skip")
    ap.add_argument("--config2", type=int, default=1, help="N=1 only: also run BASELINE configs[1] (G=1024, K=20, 1000 steps) in a child; 0 = skip")
    ap.add_argument("--child", default="", choices=["", "config2"], help="internal: run one extra measurement and print its JSON object")
    ap.add_argument("--headline-ms", type=float, default=0.0, help="internal (--child config2): this run's G=8192 ms/step")
    ap.add_argument("--skip-extras", action="store_true",
                    help="only the timed region (no roofline / e2e / cpu_baseline passes): for ncu launch lists")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.child == "config2":
        run_config2_child(args)
        return
    if args.impl == "reference":
        if args.grid == 0:
            args.grid = 8192      # the reference arm always samples the single-GPU configuration
        run_reference(args)
        return
    if args.grid == 0:
        args.grid = 8192 if world == 1 else 32768
    args.warmup = max(args.warmup, 3)
    run_ours(args)


if __name__ == "__main__":
    main()
