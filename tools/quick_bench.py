"""Developer timing probe (not the bench contract): step time vs temporal-blocking depth."""
import sys, time, json
sys.path.insert(0, ".")
import torch
from fluidsimulationcuda_b200 import solver as SF

def time_step(N, K, T, arith=0, steps=5, warm=2, chunk=0):
    s = SF.StableFluids(N, sweeps_per_launch=T, arithmetic=arith)
    s.set_option(SF.SF_OPT_CHUNK_ROWS, chunk)
    f = [s.new_field() for _ in range(6)]
    s.init_synthetic(1, *f)
    src = [f[1].clone(), f[3].clone(), f[5].clone()]
    def one():
        f[1].copy_(src[0]); f[3].copy_(src[1]); f[5].copy_(src[2])
        s.step(*f, 0.0025, 0.1, 0.016, K)
    for _ in range(warm): one()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ts = []
    for i in range(steps):
        f[1].copy_(src[0]); f[3].copy_(src[1]); f[5].copy_(src[2])
        ev[0].record(); s.step(*f, 0.0025, 0.1, 0.016, K); ev[1].record()
        torch.cuda.synchronize(); ts.append(ev[0].elapsed_time(ev[1]))
    s.close()
    return min(ts), sum(ts) / len(ts)

def time_solve(N, K, T, alpha, beta, arith=0, reps=5, chunk=0):
    s = SF.StableFluids(N, sweeps_per_launch=T, arithmetic=arith, use_graph=False)
    s.set_option(SF.SF_OPT_CHUNK_ROWS, chunk)
    x, x0 = s.new_field(), s.new_field()
    x.uniform_(0, 1); x0.uniform_(0, 1)
    for _ in range(2): s.diffuse(0, x, x0, alpha, beta, K)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        a.record(); s.diffuse(0, x, x0, alpha, beta, K); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    s.close()
    return min(ts)

if __name__ == "__main__":
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    N = G - 2
    cells = G * G
    print(torch.cuda.get_device_name(0), "G", G, "K", K, flush=True)
    for T in (1, 2, 4, 5, 7, 8):
        for name, (al, be) in (("pressure", (1.0, 4.0)), ("strict", (2683.2, 10733.8))):
            ms = time_solve(N, K, T, al, be)
            print(f"solve T={T} {name:8s} {ms:8.3f} ms  {K*N*N/ms/1e6:9.1f} Mupd/ms-eq  eff GB/s {12.0*K*cells/ms/1e6:9.1f}", flush=True)
        ms = time_solve(N, K, T, 2683.2, 10733.8, arith=1)
        print(f"solve T={T} fast     {ms:8.3f} ms  eff GB/s {12.0*K*cells/ms/1e6:9.1f}", flush=True)
    for T in (4, 5, 8):
        mn, av = time_step(N, K, T)
        B = (60 * K + 148) * cells
        print(f"step  T={T} strict min {mn:8.3f} ms avg {av:8.3f}  eff GB/s {B/mn/1e6:9.1f}  frac {B/mn/1e6/6543.1:6.3f}", flush=True)
    mn, av = time_step(N, K, 8, arith=1)
    print(f"step  T=8 fast   min {mn:8.3f} ms avg {av:8.3f}", flush=True)
