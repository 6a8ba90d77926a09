"""Randomised campaign: the kernel source run on the CPU (tests/kernel_emulation) against the
oracle on small images filled with awkward field values -- signed zeros, the fast path's range
limits 2^-40 / 2^40 and their neighbours, denormals, the largest finite values, NaN, infinities,
60 decades of dynamic range -- over both dtypes, both modes, every wall combination, both
arithmetic builds and both index widths.  Bitwise comparison, NaN payloads included.

    python tools/emulation_campaign.py <first seed> <seconds> [walk]
    python tools/emulation_campaign.py <first seed> <seconds> replay    (recorded paths: record + replay kernels)
    python tools/emulation_campaign.py <first seed> <seconds> bands     (the same, every pass cut into its own random row bands)

`walk` (default 0) selects the formulation of the pass kernels: 0 per-step, 1 the grouped walk
as the library dispatches it (rlic::Tune), "mix" picks per case among the per-step walk, the
tuned grouped walk and explicit (flavor, walk) pairs of the grouped walk.

Round 1: seeds 100000..233822 (133 823 cases, 300 s on 8 cores): 0 mismatches.
Round 1, `mix` (grouped-walk formulations included): seeds 300000..464805 (164 806 cases, 600 s), 500000..534981
(34 982, after the alignment-compare change) and 600000..1408645 (808 646 cases, 1800 s): 0 mismatches.
Round 2, `replay` (recorded paths: the recording walk + either replay kernel, 2-4 iterations, kernels of 1-140 taps):
seeds 2000000..2227474 (227 475 cases, 900 s) and 3000000..3282116 (282 117 cases, 1500 s): 0 mismatches.
Round 2, `bands` (the same with every pass cut into its own random row bands, as the host path and the slab drivers
launch them): seeds 5100000..5255462 (155 463 cases, 800 s): 0 mismatches.
Round 2, `mix` on the kernels as they ship at the end of the round: seeds 4000000..4258330 (258 331 cases, 800 s): 0 mismatches.
"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT))
import numpy as np, oracle, kernel_emulation as ke
WALLS=[(("closed","closed"),("closed","closed")),(("periodic","periodic"),("periodic","periodic")),(("periodic","periodic"),("closed","closed")),(("closed","closed"),("periodic","periodic"))]
def nasty(rng, shape, dtype):
    fi=np.finfo(dtype)
    pool=np.array([0.0,-0.0,1.0,-1.0,0.5,-0.5,2.0**-40,-(2.0**-40),2.0**-41,2.0**40,-(2.0**40),2.0**41,float(fi.tiny),-float(fi.tiny),float(fi.tiny)/8,float(fi.max),-float(fi.max),float(fi.max)/2,np.nan,np.inf,-np.inf,1e-20,-1e-20,3.0,1/3,-1/3, 1+2.0**-20, 2.0**-60, 2.0**-61],dtype=np.float64)
    style=rng.integers(5)
    with np.errstate(all="ignore"):
        if style==0: a=rng.choice(pool,size=shape)
        elif style==1: a=rng.standard_normal(shape)*10.0**rng.integers(-45,39,size=shape)
        elif style==2: a=rng.choice([-1.0,0.0,1.0,-0.0,0.5,-2.0],size=shape)
        elif style==3:
            a=rng.standard_normal(shape); m=rng.random(shape)<0.15; a[m]=rng.choice(pool,size=int(m.sum()))
        else: a=(rng.random(shape)-0.5)
        return a.astype(dtype)
t0=time.time(); n=0; bad=0
seed0=int(sys.argv[1]); budget=float(sys.argv[2]); walk_arg=sys.argv[3] if len(sys.argv)>3 else "0"
seed=seed0
while time.time()-t0<budget:
    rng=np.random.default_rng(seed); seed+=1
    dtype=[np.float32,np.float64][rng.integers(2)]
    ny,nx=(int(x) for x in rng.integers(1,40,size=2)); klen=int(rng.integers(1,70))
    tex=rng.random((ny,nx)).astype(dtype); u=nasty(rng,(ny,nx),dtype); v=nasty(rng,(ny,nx),dtype)
    k=(rng.random(klen)-0.3).astype(dtype)
    mode=["velocity","polarization"][rng.integers(2)]; walls=WALLS[rng.integers(4)]; its=int(rng.integers(1,4))
    branchless=bool(rng.integers(2)); wide=bool(rng.integers(2))
    how={}
    if walk_arg=="mix":
        pick=int(rng.integers(4))
        if pick==1: how=dict(walk=1)
        elif pick>=2: how=dict(walk=[1,9][int(rng.integers(2))],flavor=int(rng.integers(4)),admit=int(rng.integers(4)))
    elif walk_arg=="replay":
        # recorded paths: pass 1 records, the others replay (what the library does for iterations >= 2);
        # kernels of up to 140 taps (one to three groups of 32 steps per half), either replay kernel
        its=int(rng.integers(2,5)); klen=int(rng.integers(1,141))
        k=(rng.random(klen)-0.3).astype(dtype)
        how=dict(paths=[True,"staged"][int(rng.integers(2))])
    elif walk_arg not in ("0","bands"): how=dict(walk=int(walk_arg))
    if how: branchless=True                      # the grouped walk exists for the default arithmetic
    if "flavor" in how: wide=False; klen=min(klen,60)   # explicit formulations: 32-bit indices only
    k=k[:klen]
    if walk_arg=="bands":
        # the host path and the slab drivers: every pass in its own random row bands (the record is written
        # by one set of launches and read by others), replays by either kernel
        from rlic_b200 import _core
        its=int(rng.integers(2,5)); klen=int(rng.integers(1,100)); k=(rng.random(klen)-0.3).astype(dtype)
        def bands():
            cuts=sorted(set(int(c) for c in rng.integers(1,max(ny,2),size=int(rng.integers(0,4))) if 0<c<ny))
            edges=[0,*cuts,ny]; return [(a,b-a) for a,b in zip(edges,edges[1:])]
        with np.errstate(all="ignore"):
            b=ke.Buffers(dtype,ny,nx,_core.wall_codes(walls),klen); b.pack_field(u,v); b.pad_texture(tex,0)
            rec=b.path_record(klen); src=0
            for it in range(its):
                for rows in bands():
                    if it==0: b.run_pass_paths(src,1-src,k,mode,1,rec,rows=rows,wide=wide)
                    elif not (rng.integers(2) and b.run_pass_paths(src,1-src,k,mode,3,rec,rows=rows,wide=wide)):
                        b.run_pass_paths(src,1-src,k,mode,2,rec,rows=rows,wide=wide)
                src=1-src
            got=b.unpad_texture(src)
            want=oracle.convolve(tex,u,v,kernel=k,uv_mode=mode,boundaries=walls,iterations=its,variant=3)
        n+=1
        if not np.array_equal(got.view(np.uint8),want.view(np.uint8)):
            bad+=1; print("MISMATCH seed",seed-1,dtype.__name__,(ny,nx),klen,mode,walls,its,"bands",flush=True)
            if bad>10: break
        continue
    with np.errstate(all="ignore"):
        got=ke.convolve(tex,u,v,kernel=k,uv_mode=mode,boundaries=walls,iterations=its,branchless=branchless,wide=wide,**how)
        want=oracle.convolve(tex,u,v,kernel=k,uv_mode=mode,boundaries=walls,iterations=its,variant=3 if branchless else 1)
    n+=1
    if not np.array_equal(got.view(np.uint8),want.view(np.uint8)):
        # allow NaN payload differences? report both
        same_val=np.array_equal(got,want,equal_nan=True)
        bad+=1; print("MISMATCH seed",seed-1,dtype.__name__,(ny,nx),klen,mode,walls,its,"branchless",branchless,how,"values equal:",same_val, "ndiff",(got!=want).sum(), flush=True)
        if bad>10: break
print("cases",n,"bad",bad,"seeds",seed0,"..",seed-1)
