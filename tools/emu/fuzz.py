"""Random-case fuzz of the emulated kernel source against the oracle / the in-place red-black scheme (no GPU).
usage: fuzz.py <default|il2_edge> <seed> <seconds>    -- grid widths 4..128, depths 1..8, chunkings, zero guess, decaying fronts,
outlier rows, red-black levels.  Prints the failing case and exits 1 on the first mismatch."""
import ctypes as C, numpy as np, sys, time, os
HERE=os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0,os.path.join(HERE,'..','..')); sys.path.insert(0,HERE)
import build_emu
from oracle.pyoracle import Oracle, RedBlackCheck
variant=sys.argv[1]; defs={"default":(),"il2_edge":("-DSF_INNER_LOOP=2","-DSF_EDGE_SPLIT=1"),"gg":("-DSF_GUARDED_GROUP=1",),"all":("-DSF_GUARDED_GROUP=1","-DSF_INNER_LOOP=2","-DSF_EDGE_SPLIT=1")}[variant]
L=C.CDLL(build_emu.build("fz_"+variant, defs)); FP=C.POINTER(C.c_float)
L.emu_lin_solve.argtypes=[C.c_int,C.c_int,FP,FP,C.c_float,C.c_float,C.c_int,C.c_int,C.c_int,C.c_int,C.c_int,C.c_float]; L.emu_lin_solve.restype=C.c_int
p=lambda a:a.ctypes.data_as(FP)
o=Oracle(); rb=RedBlackCheck(); rng=np.random.default_rng(int(sys.argv[2])); t0=time.time(); n=0
while time.time()-t0 < float(sys.argv[3]):
    N=int(rng.choice([2,6,10,14,18,22,26,30,46,62,110,114,126]))
    G=N+2; T=int(rng.choice([1,2,3,4,5,6,7,8])); b=int(rng.integers(0,3))
    al,be=[(1.0,4.0),(0.635,3.54),(2683.2,10733.8),(107322.0,429289.0)][int(rng.integers(0,4))]
    K=int(rng.integers(1,3*T+2)); chunk=int(rng.choice([0,0,8,16,20,33]))
    zg=int(al==1.0 and rng.random()<0.5)
    kind=rng.random()
    x=rng.uniform(-1,1,(G,G)).astype(np.float32); x0=rng.uniform(-1,1,(G,G)).astype(np.float32)
    if kind<0.25:   # compact tiny source -> decaying front
        x[...]=0; x0[...]=0; r0=int(rng.integers(1,N)); c0=int(rng.integers(1,N))
        x0[r0:r0+5,c0:c0+9]=rng.uniform(0,1,(G,G)).astype(np.float32)[r0:r0+5,c0:c0+9]*np.float32(10.0**rng.integers(-30,-20))
    elif kind<0.35: # outliers
        x[int(rng.integers(0,G)),int(rng.integers(0,G))]=np.float32(1e32)
    if zg: x[...]=0
    rbmode = (rng.random()<0.3) and not zg
    if rbmode:
        om=float(np.float32(rng.choice([1.0,1.5,0.7,1.9])))
        want=x.copy(); rb.rb_diffuse(N,b,want,x0,al,be,K,om)
        got=x.copy(); rc=L.emu_lin_solve(N,b,p(got),p(x0),al,be,K,6,0,chunk,1,om)
    else:
        want=x.copy(); o.diffuse(N,b,want,x0,al,be,K)
        got=x.copy()
        if zg: got[...]=np.nan
        rc=L.emu_lin_solve(N,b,p(got),p(x0),al,be,K,T,zg,chunk,0,1.0)
    assert rc==0,(N,T,K)
    if not np.array_equal(got.view(np.uint32),want.view(np.uint32)):
        print("MISMATCH",variant,dict(N=N,T=T,b=b,al=al,K=K,chunk=chunk,zg=zg,kind=kind,rb=rbmode),flush=True); sys.exit(1)
    n+=1
print("fuzz",variant,"cases",n,"all identical")
