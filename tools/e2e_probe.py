import time, ctypes, numpy as np, torch, sys
sys.path.insert(0, '/root/repo')
import rlic_b200
from rlic_b200 import _core, workloads
w = workloads.vortex_noise(4096, iterations=5)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
tex, u, v = pin(w.texture), pin(w.u), pin(w.v)
k = w.kernel
p = ctypes.POINTER(ctypes.c_float)
def call(t,uu,vv,out,its=5):
    rc = _core.lib.rlic_b200_convolve_f32(t.ctypes.data_as(p), uu.ctypes.data_as(p), vv.ctypes.data_as(p), 4096, 4096, k.ctypes.data_as(p), k.size, 0, 0,0,0,0, its, out.ctypes.data_as(p)); _core.check(rc)
def bench(name, fn, n=5):
    fn(); fn()
    t0=time.perf_counter()
    for _ in range(n): fn()
    print(f"{name:50s} {(time.perf_counter()-t0)/n*1e3:8.2f} ms")
out_pinned = pin(np.empty_like(tex))
out_page = np.empty_like(tex); out_page[:] = 0
bench("pinned in, pinned out (reused)", lambda: call(tex,u,v,out_pinned))
bench("pinned in, pageable out (reused, touched)", lambda: call(tex,u,v,out_page))
bench("pinned in, fresh np.empty out", lambda: call(tex,u,v,np.empty_like(tex)))
bench("public API (pinned in)", lambda: rlic_b200.convolve(tex,u,v,kernel=k,iterations=5))
tp,up,vp = w.texture, np.ascontiguousarray(w.u), np.ascontiguousarray(w.v)
bench("pageable in, pinned out", lambda: call(tp,up,vp,out_pinned))
bench("public API (pageable in)", lambda: rlic_b200.convolve(tp,up,vp,kernel=k,iterations=5))
bench("pinned in/out, 1 iteration", lambda: call(tex,u,v,out_pinned,1))
bench("np.empty + touch 64MB", lambda: np.empty_like(tex).fill(0))
bench("validation texture<0 any", lambda: np.any(tex<0))

# RLIC_B200_TRACE=1 python tools/e2e_probe.py: where a call's time goes (lic_api.cu: HostTrace)
import os
if os.environ.get("RLIC_B200_TRACE"):
    print("--- traced calls: pinned in / pinned out, then pageable in / pinned out", file=sys.stderr)
    for _ in range(3):
        call(tex, u, v, out_pinned)
    for _ in range(3):
        call(tp, up, vp, out_pinned)
