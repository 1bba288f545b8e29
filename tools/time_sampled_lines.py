"""Sampled-line bounces of a compiled model (phases.phase_scan) against the exact line functor of the built-in model1
(phases.model1_phase_scan) on the same 2000-point x 256-temperature scan: time, T_nuc and S differences by S/T band.
usage: python tools/time_sampled_lines.py   (needs a B200)"""
import sys, time, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from cosmotransitions_b200 import jit, phases, potentials
P=2000
rng = np.random.default_rng(20260102)
jitter = rng.uniform(.9, 1.1, (P, 5))
rows = np.stack([potentials.model1_model8(120. * a, 50. * b, 25. * c, .1 * d, .15 * f) for a, b, c, d, f in jitter])
T = np.linspace(40.0, 135.0, 256)
m = jit.model1_compiled()
v = np.sqrt(rows[:, 7]); seed=np.stack([v, v], -1)
for name, fn in (("built-in exact line functor", lambda: phases.model1_phase_scan(rows, T)), ("compiled + sampled lines", lambda: phases.phase_scan(m, rows, T, seed, method="sampled"))):
    fn(); torch.cuda.synchronize()
    t0=time.perf_counter(); r=fn(); torch.cuda.synchronize(); dt=time.perf_counter()-t0
    st=r["status"].cpu().numpy()
    print(name, "%.3f s"%dt, "bounces", r["n_bounces"], "solved", int(((st==0)|(st==1)).sum()), "median Tnuc %.4f"%float(np.nanmedian(r["Tnuc"].cpu().numpy())))
    if name.startswith("built"): ref=r
d=np.abs(r["Tnuc"].cpu().numpy()-ref["Tnuc"].cpu().numpy())
print("Tnuc diff: max %.3g median %.3g K"%(np.nanmax(d), np.nanmedian(d)))
both=((st==0)|(st==1))&((ref["status"].cpu().numpy()==0)|(ref["status"].cpu().numpy()==1))
rel=np.abs(r["S"].cpu().numpy()[both]-ref["S"].cpu().numpy()[both])/ref["S"].cpu().numpy()[both]
print("S rel diff: median %.2e  99%% %.2e  max %.2e  n %d"%(np.median(rel), np.quantile(rel,0.99), rel.max(), both.sum()))
Tg = np.broadcast_to(T[None, :], st.shape)[both]
soT = ref["S"].cpu().numpy()[both] / Tg
for lo, hi in ((0, 300), (300, 1000), (1000, 1e4), (1e4, 1e9)):
    sel = (soT >= lo) & (soT < hi)
    if sel.any():
        print("S/T in [%g, %g): n %d  median %.2e  99%% %.2e  max %.2e" % (lo, hi, sel.sum(), np.median(rel[sel]), np.quantile(rel[sel], 0.99), rel[sel].max()))
