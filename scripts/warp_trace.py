"""Per-warp timeline of the streaming resampler (needs a library built with -DVSTAB_DBG_TRACE)."""
import os, sys, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vstab_loader
vstab_loader.load()
from vstab_b200 import _native

dev = torch.device("cuda", 0)
h = _native.get_handle(dev)
w, hh, n = 1920, 1080, 32
src = torch.rand((n, hh, w, 3), device=dev)
rng = np.random.default_rng(0)
mats = []
for i in range(n):
    th, s = rng.normal(0, 0.004), 1 + rng.normal(0, 0.003)
    mats.append([s * np.cos(th), -s * np.sin(th), rng.normal(0, 8), s * np.sin(th), s * np.cos(th), rng.normal(0, 6), 0, 0, 1])
fwd = torch.tensor(mats, dtype=torch.float32, device=dev).reshape(n, 1, 9)
dst = torch.empty((n, hh, w, 3), device=dev)
mask_all = torch.zeros((n + 1, hh, w), device=dev)
mask = mask_all[:n]
for _ in range(3):
    h.warp_fused(src, fwd, (w, hh), "bilinear", (0.5, 0.5, 0.5), out=dst, mask_out=mask)
torch.cuda.synchronize()
raw = mask_all[n].reshape(-1).view(torch.int32).cpu().numpy().astype(np.int64) & 0xFFFFFFFF
out = {}
for cta in (0, 37, 74, 111, 148):
    if (cta * 128 + 128) * 64 > raw.size:
        continue
    tr = raw[cta * 128 * 64:(cta * 128 + 128) * 64].reshape(128, 8, 8)
    its = slice(20, 100)
    ws, we, dn, isu = tr[its, :, 0], tr[its, :, 1], tr[its, :, 2], tr[its, :, 3]
    wait = (we - ws) & 0xFFFFFFFF
    comp = (dn - we) & 0xFFFFFFFF
    period = ((we[1:, :] - we[:-1, :]) & 0xFFFFFFFF)
    skew = ((dn.max(axis=1) - dn.min(axis=1)) & 0xFFFFFFFF)
    issued = isu.max(axis=1)
    issue_delay = (issued - dn.max(axis=1)) & 0xFFFFFFFF
    who = isu.argmax(axis=1)
    rows = np.arange(isu.shape[0])
    own_done = dn[rows, who]
    d_atomic = (tr[its, :, 4][rows, who] - own_done) & 0xFFFFFFFF
    d_issue = (tr[its, :, 5][rows, who] - tr[its, :, 4][rows, who]) & 0xFFFFFFFF
    d_req = (issued - tr[its, :, 5][rows, who]) & 0xFFFFFFFF
    last_is_issuer = (dn.argmax(axis=1) == who)
    # load latency: issue at end of tile j -> data seen at wait_end of tile j+2 by the first warp that had to wait
    lat = ((we[2:, :].min(axis=1) - issued[:-2]) & 0xFFFFFFFF)
    out[cta] = {k: [float(np.median(v)), float(np.mean(v)), float(np.percentile(v, 90))] for k, v in
                dict(wait=wait, compute=comp, period=period, skew=skew, issue_delay=issue_delay, issue_to_first_use=lat, d_atomic=d_atomic, d_issue=d_issue, d_req=d_req, last_is_issuer=last_is_issuer.astype(float)).items()}
print(json.dumps(out, indent=1))
