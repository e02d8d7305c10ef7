"""Times the stages of the end-to-end path bench.py's `e2e` key measures (host buffers -> C ABI -> optimised parameters + residuals)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from calico_b200 import _capi, synthetic
import bench

cfg = sys.argv[1] if len(sys.argv) > 1 else "C4"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
def gpu_api():
    a = _capi.CApi(); a.set_device(0); return a
truth, prob = synthetic.generate(cfg, gpu_api, noise=True)
opts = bench.bench_options(_capi.Options, iters)
for rep in range(3):
    t = [time.perf_counter()]
    api = gpu_api(); t.append(time.perf_counter())
    p2 = prob.clone(); t.append(time.perf_counter())
    ids = p2.push(api); t.append(time.perf_counter())
    api.upload(); t.append(time.perf_counter())
    summ, log = api.optimize(opts); t.append(time.perf_counter())
    p2.pull(api, ids); t.append(time.perf_counter())
    p2.residuals(api, ids)
    t.append(time.perf_counter())
    st = api.stats()
    api.close(); t.append(time.perf_counter())
    names = ["create", "clone(py)", "push", "upload", "optimize", "pull", "get_residuals", "close"]
    print("rep", rep, " ".join(f"{n}={1e3*(b-a):.1f}ms" for n, a, b in zip(names, t[:-1], t[1:])), f"total={1e3*(t[-1]-t[0]):.1f}ms loop={st.lm_loop_ms:.1f}ms h2d={st.h2d_bytes/1e6:.1f}MB d2h={st.d2h_bytes/1e6:.1f}MB")
