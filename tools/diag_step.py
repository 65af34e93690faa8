"""Per-call wall/device timing of one hot-path step (diagnostics, run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import umeregrobust_b200 as ume
from umeregrobust_b200 import synth, _lib

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 16
b = synth.make_batch(pairs, seed0=0, n_base=2, N=120000, C=32, n_kp=1024)
d = {k: torch.from_numpy(v).cuda() for k, v in b.items() if k.endswith(("pts", "feat", "kp"))}

def timed(name, fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    print("  %-28s wall %.3f ms  device %.3f ms" % (name, (time.perf_counter() - t0) / reps * 1e3, e0.elapsed_time(e1) / reps))
    return out

for div2 in (False, True):
    ume.config["cell_div2"] = div2
    print("cell_div2 =", div2)
    F, Fc = timed("moments src", lambda: ume.ume_moments(d["src_pts"], d["src_kp"], d["src_feat"], 750, 5.0, return_centered=True))
    G, Gc = timed("moments tgt", lambda: ume.ume_moments(d["tgt_pts"], d["tgt_kp"], d["tgt_feat"], 750, 5.0, return_centered=True))
    Q1 = timed("descriptors", lambda: ume.ume_descriptors(Fc)); Q2 = ume.ume_descriptors(Gc)
    D, am, dm = timed("cdist tc (D+argmin)", lambda: ume.descriptor_cdist(Q1, Q2, want_D=True, want_argmin=True))
    timed("cdist tc (argmin only)", lambda: ume.descriptor_cdist(Q1, Q2, want_D=False, want_argmin=True))
    timed("cdist simt", lambda: ume.descriptor_cdist(Q1, Q2, want_D=True, want_argmin=True, impl=0))
    timed("rigid", lambda: ume.rigid_solve(Fc, Gc, None, am, d["src_kp"], d["tgt_kp"]))
    timed("whole step", lambda: ume.register_hypotheses(d["src_pts"], d["src_feat"], d["src_kp"], d["tgt_pts"], d["tgt_feat"], d["tgt_kp"], 750, 5.0, want_D=True))
