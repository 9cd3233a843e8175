import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vstab_loader; vstab_loader.load()
from vstab_b200 import _native
from oracle import dis_ref
from tests import cases
h = _native.get_handle(torch.device("cuda", 0))
p, c = cases.make_gray_pair(cases.DIS_CASES[0])
want = dis_ref.calc(p, c)
g = torch.from_numpy(np.stack([p, c])).cuda()
flow, _ = h.dis_flow(g, want_flow=True, grid_step=8)
torch.cuda.synchronize()
f = flow[0].cpu().numpy()
print("cluster_max", os.environ.get("VSTAB_VR_CLUSTER_MAX"), "nan", int(np.isnan(f).sum()), "maxerr", float(np.nanmax(np.abs(f - want))), "exact", float((f == want).mean()))
big = torch.from_numpy(np.stack([p, c] * 61)[:121]).cuda()
for _ in range(2): h.dis_flow(big, want_flow=False, grid_step=8)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): h.dis_flow(big, want_flow=False, grid_step=8)
torch.cuda.synchronize(); print("dis 120 pairs ms", (time.perf_counter() - t0) / 5 * 1e3)
