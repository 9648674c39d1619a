"""Sweep the host-pipeline chunking of dc_score_grad_host on the bench workload (run on the GPU box)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from diffco_b200 import DiffCo, functional as Fn, kernel as K, model as M, _lib

dev = torch.device("cuda", 0)
S, w, q = bench.make_problem(0, bench.BATCH)
robot = M.RevolutePlanarRobot(1.0, 0.3, dof=7)
dc = DiffCo(kernel_func=K.RQKernel(10.0), transform=robot.fkine)
dc.support_points = S.float().to(dev)
dc.support_transformed = robot.fkine(dc.support_points)
dc.gains = w.float().to(dev)
sv, kf = dc._select("gains")
fk = dc._fk_for(sv)
qh = q.float().pin_memory()
oh = torch.empty(len(q), 8).pin_memory()
qd = q.float().to(dev)
od = torch.empty(len(q), 8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts)//2] * 1e3

print("H2D q      us", timeit(lambda: qd.copy_(qh, non_blocking=True)))
print("D2H out    us", timeit(lambda: oh.copy_(od, non_blocking=True)))
print("kernel     us", timeit(lambda: Fn.score_grad(fk, kf.desc, sv, qd, _lib.DC_GRAD_SUM, out=od)))
def unchunked():
    qd.copy_(qh, non_blocking=True); Fn.score_grad(fk, kf.desc, sv, qd, _lib.DC_GRAD_SUM, out=od); oh.copy_(od, non_blocking=True)
print("serial     us", timeit(unchunked))
for rows in (9472, 16384, 18944, 28416, 32768, 37888, 65536):
    for slots in (2, 3):
        pipe = Fn.HostPipeline(dev, chunk_rows=rows, n_slots=slots)
        print(f"pipeline chunk={rows:6d} slots={slots}  us", timeit(lambda: pipe.score_grad(fk, kf.desc, sv, qh, oh)))
for B in (9472, 18944, 16384, 32768):
    sub = qd[:B].contiguous(); so = od[:B]
    print(f"kernel B={B} us", timeit(lambda: Fn.score_grad(fk, kf.desc, sv, sub, _lib.DC_GRAD_SUM, out=so)))
