"""Developer tool: per-warp range timings of the Jacobi launches of one dens_step / one viscosity solve (G=8192, K=40).
Needs the instrumented build:  python -m fluidsimulationcuda_b200.build --out build/libsf_dbg.so -DSF_WARP_TIMES
run:  SF_LIBRARY=$PWD/build/libsf_dbg.so python tools/warp_times.py"""
import ctypes as C, sys
sys.path.insert(0, ".")
import numpy as np, torch
from fluidsimulationcuda_b200 import solver as SF
G, K = 8192, 40
s = SF.StableFluids(G - 2, use_graph=False)
import os
if os.environ.get("SF_SKEW"):
    s.set_option(SF.SF_OPT_WAVE_SKEW, int(os.environ["SF_SKEW"]))
lib = SF.load_library()
f = [s.new_field() for _ in range(6)]
s.init_synthetic(1, *f)
for i in range(8):
    s.init_sources(10 + i, f[1], f[3], f[5]); s.step(*f, 0.0025, 0.1, 0.016, K)
torch.cuda.synchronize()
cap = 400000
buf = torch.zeros(cap * 4, dtype=torch.int64, device="cuda")      # 32 bytes per record
rec = np.dtype([("t0", "<u8"), ("t1", "<u8"), ("band", "<i4"), ("lo", "<i4"), ("hi", "<i4"), ("stolen", "<i4")])

def capture(fn, what):
    assert lib.sf_debug_warp_times(C.c_void_p(buf.data_ptr()), C.c_uint(cap)) == 0
    fn(); torch.cuda.synchronize()
    n = C.c_uint(0); lib.sf_debug_warp_times_count(C.byref(n))
    lib.sf_debug_warp_times(C.c_void_p(0), C.c_uint(0))
    r = np.frombuffer(buf.cpu().numpy().tobytes(), dtype=rec)[:min(n.value, cap)]
    r = np.sort(r, order="t0")
    r = r.copy()
    smid = (r["stolen"] >> 8) & 0xfff; warpid = (r["stolen"] >> 20) & 0xfff
    r["stolen"] &= 0xff
    if "--hw" in sys.argv:       # speed of a range against the hardware warp slot it ran in
        speed = (r["hi"] - r["lo"]) / ((r["t1"] - r["t0"]).astype(np.float64) / 1e3)
        own = r["stolen"] == 0
        print(f"== {what}: rows/us by %warpid (all launches, own ranges):")
        for wslot in sorted(set(warpid[own].tolist())):
            m = own & (warpid == wslot)
            print(f"   warpid {wslot:3d}: n={int(m.sum()):5d} rows/us min {speed[m].min():.2f} mean {speed[m].mean():.2f} max {speed[m].max():.2f}")
        print("   SMs seen:", len(set(smid.tolist())), " ranges per SM min/max:", np.bincount(smid[own]).min(), np.bincount(smid[own]).max())
    # split into launches at gaps: a new launch starts when t0 jumps past every earlier t1
    launches, cur, hi = [], [], 0
    for x in r:
        if cur and x["t0"] > hi + 500:
            launches.append(np.array(cur, dtype=rec)); cur = []
        cur.append(x); hi = max(hi, int(x["t1"]))
    if cur: launches.append(np.array(cur, dtype=rec))
    print(f"== {what}: {n.value} ranges in {len(launches)} launches")
    for k, L in enumerate(launches):
        t_begin, t_end = int(L["t0"].min()), int(L["t1"].max())
        dur = (L["t1"] - L["t0"]).astype(np.int64) / 1e3
        end = (L["t1"].astype(np.int64) - t_begin) / 1e3
        own = L[L["stolen"] == 0]
        print(f" launch {k}: {len(L)} ranges ({int((L['stolen'] == 1).sum())} taken over), span {(t_end - t_begin) / 1e3:7.1f} us; "
              f"range end times us: p10 {np.percentile(end, 10):6.1f} p50 {np.percentile(end, 50):6.1f} p90 {np.percentile(end, 90):6.1f} "
              f"p99 {np.percentile(end, 99):6.1f} max {end.max():6.1f}; own-range duration p50 {np.percentile((own['t1'] - own['t0']) / 1e3, 50):6.1f} "
              f"max {((own['t1'] - own['t0']) / 1e3).max():6.1f}")
        if k in (1, 2):
            last = np.argsort(end)[-8:][::-1]
            for i in last:
                x = L[i]
                print(f"     late: band {x['band']:3d} rows [{x['lo']:5d},{x['hi']:5d}) {'taken ' if x['stolen'] else 'own   '} start {(int(x['t0']) - t_begin) / 1e3:6.1f} dur {dur[i]:6.1f} end {end[i]:6.1f}")
            # where is the time: histogram of own-range durations by band
            byband = {}
            for x in own:
                byband.setdefault(int(x["band"]), []).append((int(x["t1"]) - int(x["t0"])) / 1e3)
            print("     own-range duration by band (max us):", " ".join(f"{b}:{max(v):.0f}" for b, v in sorted(byband.items())))
            bychunk = {}
            for x in own:
                bychunk.setdefault(int(x["lo"]), []).append((int(x["t1"]) - int(x["t0"])) / 1e3)
            print("     own-range duration by first row of the chunk (min/mean/max us):",
                  " ".join(f"{lo}:{min(v):.0f}/{sum(v) / len(v):.0f}/{max(v):.0f}" for lo, v in sorted(bychunk.items())))

import numpy as np
f32 = np.float32
def ab(c):
    a = f32(0.016) * f32(c); a = a * f32(G - 2); a = a * f32(G - 2); return float(a), float(f32(1) + f32(4) * a)
s.init_sources(50, f[1], f[3], f[5])
if "--dens" in sys.argv:
    capture(lambda: s.dens_step(f[0], f[1], f[2], f[4], 0.1, 0.016, K), "dens_step (work-stealing variants)")
al, be = ab(0.0025)
capture(lambda: s.diffuse(1, f[3], f[2], al, be, K), "viscosity solve of u (plain variant)")
capture(lambda: s.diffuse(0, f[5], f[4], 1.0, 4.0, K), "pressure-like solve")
