import os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from rapidquadcoptercollisiondetection_b200 import abi, workloads as W
import ctypes as C
# load two builds side by side
def ctx_for(path):
    os.environ["RQCD_LIB"] = path
    import importlib
    import rapidquadcoptercollisiondetection_b200.abi as A
    A._lib = None; A.LIB_PATH = path
    from rapidquadcoptercollisiondetection_b200.api import Context
    return Context(0)
root = os.getcwd()
exact = ctx_for(os.path.join(root, "rapidquadcoptercollisiondetection_b200/librqcd.so"))
fast = ctx_for(os.path.join(root, "variants/librqcd_fmad.so"))
dev = torch.device("cuda", 0)
for name, n, m, seed, oseed in (("c3", 1 << 20, 64, 1, 2), ("c5-slice", 1 << 22, 256, 5, 6)):
    specs = W.static_obstacles(oseed, m // 2, m // 2)
    obs = abi.make_obstacles(specs)
    res = {}
    for tag, ctx in (("exact", exact), ("fmad", fast)):
        ctx.set_obstacles(obs, m)
        d_bc = torch.empty((19, n), dtype=torch.float64, device=dev)
        ctx.synth_bc(seed, 0, n, d_bc.data_ptr(), n)
        mtx = torch.empty((n, m), dtype=torch.uint8, device=dev)
        ver = torch.empty(n, dtype=torch.uint8, device=dev)
        ctx.generate_check_dev(d_bc.data_ptr(), n, n, 0.002, d_verdict=ver.data_ptr(), d_matrix=mtx.data_ptr(), want_summary=True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = torch.cuda.Stream(dev); ctx.set_stream(st.cuda_stream)
        ctx.generate_check_dev(d_bc.data_ptr(), n, n, 0.002, d_verdict=ver.data_ptr(), want_summary=True)
        ev0.record(st)
        for _ in range(3):
            s = ctx.generate_check_dev(d_bc.data_ptr(), n, n, 0.002, d_verdict=ver.data_ptr(), want_summary=True)
        ev1.record(st); torch.cuda.synchronize()
        res[tag] = (mtx, s, n * m * 3 / (ev0.elapsed_time(ev1) * 1e-3))
    diff = int((res["exact"][0] != res["fmad"][0]).sum().item())
    print(name, "pairs", n * m, "verdict differences exact vs fmad:", diff, "| exact %.3e /s fmad %.3e /s" % (res["exact"][2], res["fmad"][2]), res["exact"][1]["pair_counts"], res["fmad"][1]["pair_counts"])
