"""Timing of the hypothesis-selection stage at the reference's sizes (test_kitti_config.yaml:
ume_n_samples 2500 hypotheses, pc_corr_max_size 10000 points, corr_num_nn 20).  Run under gpurun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import umeregrobust_b200 as ume
from umeregrobust_b200 import synth, _lib

n_hyp = int(sys.argv[1]) if len(sys.argv) > 1 else 2500
Ns = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
p = synth.make_pair(3, N=40000, C=32, n_kp=min(n_hyp, 2500), model=synth.KITTI, feat_model="field", kp_mode="corresponding")
d = {k: torch.from_numpy(v[None]).cuda() for k, v in p.items() if k.endswith(("pts", "feat", "kp"))}
out = ume.register_hypotheses(d["src_pts"], d["src_feat"], d["src_kp"], d["tgt_pts"], d["tgt_feat"], d["tgt_kp"], 750, 5.0)
T = out["T"][0].contiguous()
if T.shape[0] < n_hyp:
    T = T.repeat((n_hyp + T.shape[0] - 1) // T.shape[0], 1, 1)[:n_hyp].contiguous()
rng = np.random.default_rng(0)
ss, ts = rng.choice(40000, Ns, replace=False), rng.choice(40000, Ns, replace=False)
args = (d["src_pts"][:, ss].contiguous(), d["tgt_pts"][:, ts].contiguous(), d["src_feat"][:, ss].contiguous(), d["tgt_feat"][:, ts].contiguous(), T)
corr = ume.FeatureCorrelator(sigma=1.5, batch=64, n_hypotheses=10)
for _ in range(2):
    sc, best = corr.scores(*args)
torch.cuda.synchronize()
_lib.profile_reset(); _lib.profile_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
R = 3
for _ in range(R):
    sc, best = corr.scores(*args)
e1.record(); torch.cuda.synchronize()
_lib.profile_enable(False)
prof = _lib.profile_read()
ms = e0.elapsed_time(e1) / R
Tb = T[int(best)].cpu().numpy()
ang = np.rad2deg(np.sqrt(((Tb[:3, :3] - p["gt"][:3, :3]) ** 2).sum() / 2))
print("hypotheses %d, Ns=Nt=%d: %.2f ms per pair (corr kernel %.2f ms, knn/spatial-var %.2f ms, grid %.2f ms); queries/s %.3g; best hyp %d, angle to gt %.2f deg"
      % (n_hyp, Ns, ms, prof["corr"][0] / R, prof["knn"][0] / R, prof["grid"][0] / R, n_hyp * Ns / (ms * 1e-3), int(best), ang))
