"""Developer check: GPU kernels vs the CPU oracle on the paper sets + a small sweep."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import frontx_b200 as fx
from frontx_b200 import _lib
from oracle import fxo

b = 0.7 - 1e-7; i = 0.025
sets = {
 'LETd1': fx.models.LETd(L=0.004569, E=12930, T=1.505, Dwt=4.660e-4, theta_range=(0.019852, 0.7)),
 'LETd2': fx.models.LETd(Dwt=1.004e-3, L=1.356, E=10010, T=1.224, theta_range=(0.00625, 0.7)),
 'VGn': fx.models.VanGenuchten(n=8.093, l=2.344, Ks=2.079e-6, theta_range=(0.004943, 0.7)),
 'VGm': fx.models.VanGenuchten(m=0.8861, l=2.331, Ks=2.105e-6, theta_range=(0.005623, 0.7)),
 'BC': fx.models.BrooksAndCorey(n=0.2837, l=4.795, Ks=3.983e-6, theta_range=(2.378e-5, 0.7)),
 'LETxs': fx.models.LETxs(Lw=1.651, Ew=230.5, Tw=0.9115, Ls=0.517, Es=493.6, Ts=0.3806, Ks=8.900e-3, theta_range=(0.01176, 0.7)),
}
print(torch.cuda.get_device_name(0))
for name, m in sets.items():
    th = np.array([0.5, 0.025, b])
    g = m.derivatives(th); c = fxo.D(m.model_id, m.params(), th)
    print(name, 'D relerr', np.abs(g/c-1).max(axis=0))
models = fx.ModelBatch.from_models(list(sets.values()) + [fx.D.power_law(k=4), fx.D.philip_exact(), fx.D.constant(1.0)])
bb = np.array([b]*6 + [1.0, 1.0, 1.0]); ii = np.array([i]*6 + [0.1, 0.0, 0.0])
t=time.time(); sol = fx.solve_batch(models, b=bb, i=ii, k_max=512, throw=False); torch.cuda.synchronize(); print('gpu time', time.time()-t)
g = sol.cpu()
ref = fxo.solve_batch(models.model_id, models.params, bb, ii, k_max=512)
names = list(sets) + ['pow4', 'philip', 'const']
for j, nm in enumerate(names):
    print(nm, 'd_dob', g['d_dob'][j], ref['summary'][j,0], 'oi', g['oi'][j], ref['summary'][j,1], 'res', g['residual'][j], ref['summary'][j,5])
    print('   cnt gpu', [int(g[k][j]) for k in ('result','n_shots','n_expansions','n_attempts','n_rejected','n_rhs','n_jac','n_knots')], 'cpu', ref['counters'][j].tolist())
    nk = int(ref['counters'][j,7])
    o = np.linspace(0, ref['summary'][j,1], 200)
    th_c, dth_c = fxo.evaluate(ref['knots'][j,:nk], o)
    th_g, dth_g = sol[j](o), sol[j].d_do(o)
    print('   theta max rel', np.abs(th_g/th_c-1).max(), 'd_do max abs/scale', np.abs(dth_g-dth_c).max()/np.abs(dth_c).max())
# sweep
rng = np.random.default_rng(0); N = 2000
mb = fx.models.VanGenuchten.batch(n=rng.uniform(1.5, 9, N), alpha=np.exp(rng.uniform(np.log(0.5), np.log(5), N)),
        Ks=np.exp(rng.uniform(np.log(1e-7), np.log(1e-4), N)), l=0.5, theta_range=(0.005, 0.7))
t=time.time(); s2 = fx.solve_batch(mb, b=b, i=i, k_max=0, throw=False); torch.cuda.synchronize(); tg=time.time()-t
t=time.time(); r2 = fxo.solve_batch(mb.model_id, mb.params, b, i); tc=time.time()-t
g2 = s2.cpu()
cg = np.stack([g2[k] for k in ('result','n_shots','n_expansions','n_attempts','n_rejected','n_rhs','n_jac','n_knots')],1)
same = (cg == r2['counters']).all(1)
rel = np.abs(g2['d_dob']/r2['summary'][:,0]-1)
print('sweep N', N, 'gpu s', tg, 'cpu s', tc, 'counters identical frac', same.mean(), 'd_dob max rel', rel.max(), 'frac<1e-9', (rel<1e-9).mean())
print('result hist', np.bincount(cg[:,0]), 'shots mean', cg[:,1].mean(), 'attempts mean', cg[:,3].mean(), 'rhs mean', cg[:,5].mean())
relS = np.abs(g2['sorptivity']/r2['summary'][:,3]-1); print('S max rel', relS.max())
relth = np.abs(g2['i']/r2['summary'][:,2]-1); print('theta_oi max rel', relth.max(), 'frac<1e-6', (relth<1e-6).mean())
# throughput
for N in (20000, 100000):
    mb = fx.models.VanGenuchten.batch(n=rng.uniform(1.5, 9, N), alpha=np.exp(rng.uniform(np.log(0.5), np.log(5), N)),
        Ks=np.exp(rng.uniform(np.log(1e-7), np.log(1e-4), N)), l=0.5, theta_range=(0.005, 0.7))
    mb.to_device()
    for rep in range(2):
        torch.cuda.synchronize(); t=time.time(); s3 = fx.solve_batch(mb, b=b, i=i, k_max=0, throw=False); torch.cuda.synchronize(); tg=time.time()-t
        print('N', N, 'time', tg, 'solves/s', N/tg)
pk = __import__('ctypes').c_double()
_lib.check(_lib.lib().fx_measure_fp64_peak(__import__('ctypes').byref(pk), None)); print('fp64 peak TF', pk.value/1e12)
# inverse
o = np.linspace(0, 20, 100); th = np.exp(-o)
isol = fx.InterpolatedSolution(o, th)
tq = np.linspace(1e-6, 1, 100)
print('inverse D err', np.nanmax(np.abs(isol.D(tq) - 0.5*(1-np.log(tq)))), 'sorp', isol.sorptivity(), 'trap', fx.sorptivity(np.linspace(0,20,500), np.exp(-np.linspace(0,20,500)), b=1, i=0))
print('call err', np.abs(isol(o)-th).max())
