import sys, numpy as np
sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import lf_oracle as lo
g = np.load('tests/golden/reference_sim.npz')
nt = len(g['t'])
vel = {k: [g[k][n].astype(np.float64) for n in range(nt)] for k in ('u', 'w')}
vars_ = [[g['b'][n].astype(np.float64) for n in range(nt)]]
fp = lo.bw2_params(2, 1e-4 / 4)
kw = dict(N=(10, 1, 10), H=(3, 0, 3), topology=("Periodic", "Flat", "Bounded"), delta=(1000.0, 1.0, 10.0), filter_params=fp,
          dt=1200.0, T_out=3600.0, velocities="uw", float32_outputs=False)
a = lo.run_offline(g['t'], vel, vars_, ['b'], **kw)
b = lo.run_offline(g['t'], vel, vars_, ['b'], weights_f32=True, **kw)
for name in ("b_Lagrangian_filtered", "xi_u", "xi_w", "u_Lagrangian_filtered", "w_Lagrangian_filtered"):
    x, y = np.stack(a[name]), np.stack(b[name])
    print(f"{name:26s} rel-L2(FP32 weights vs FP64) = {np.linalg.norm(x - y) / np.linalg.norm(x):.3e}")
# a well-resolved smooth 3-D case
import lfb200
from lfb200.synthetic import make_series
gr = lfb200.RectilinearGrid(size=(64, 64, 32), extent=(64.0, 64.0, 32.0), topology=("Periodic", "Periodic", "Bounded"))
times = [0.0, 1.0, 2.0, 3.0, 4.0]
f = make_series(gr, times, var_names=("T", "b"), velocity_names=("u", "v", "w"), T_period=4.0)
fp = lo.bw2_params(2, 2 * np.pi / 4.0)
kw = dict(N=gr.N, H=gr.H, topology=gr.topology, delta=gr.spacing, filter_params=fp, dt=0.125, T_out=1.0, velocities="uvw", float32_outputs=False)
vel = {k: [x.astype(np.float64) for x in f[k]] for k in "uvw"}
vars_ = [[x.astype(np.float64) for x in f[v]] for v in ("T", "b")]
a = lo.run_offline(times, vel, vars_, ['T', 'b'], **kw)
b = lo.run_offline(times, vel, vars_, ['T', 'b'], weights_f32=True, **kw)
for name in ("T_Lagrangian_filtered", "b_Lagrangian_filtered", "xi_u", "xi_w", "u_Lagrangian_filtered"):
    x, y = np.stack(a[name]), np.stack(b[name])
    print(f"synthetic 64x64x32 {name:26s} rel-L2 = {np.linalg.norm(x - y) / np.linalg.norm(x):.3e}")
