import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from informative_path_planning_b200 import aq_library as aq, gpmodel_library as gp, _lib as L
from oracle import aq_oracle as ao
from oracle.gpmodel_oracle import OnlineGPModelOracle
g = np.load('tests/golden/max_vals.npz')
R = (0., 10., 0., 10.)
X, z = g['lt_X'], g['lt_z']
m = gp.OnlineGPModel(R, 1.0, 100.0, noise=0.5); m.add_data(X, z)
rng = np.random.default_rng(0)
F = 200
W = rng.normal(size=(F, 2)); b = 2*np.pi*rng.uniform(size=(F, 1)); eps = rng.normal(size=(F, 1))
Zd = aq.rff_features(W, b, m._X_d, 100.0, F)
Zo = ao.rff_features(W, b, X, 100.0, F)
print('Z maxdiff', np.abs(Zd - Zo).max(), Zd.shape, Zd.flags['C_CONTIGUOUS'])
th = ao.rff_theta(Zo, z, 0.5, eps, F)
th2 = aq._rff_theta(Zd, z, 0.5, eps, F)
print('theta maxdiff', np.abs(th - th2).max(), np.abs(th).max())
tg = aq.RFFTarget(100.0, F, W, th, b)
grid = rng.uniform(1, 9, (90050, 2))
vo = ao.general_target(grid, 100.0, F, W, th, b)
vals, mx, am = tg.eval_device(L.to_device(grid))
vd = vals.cpu().numpy().reshape(-1, 1)
print('vals maxdiff', np.abs(vd - vo).max(), 'argmax', int(am.item()), int(np.argmax(vo)), float(mx.item()), vo.max())
print('scalar', tg.scalar(grid[5]), vo[5], 'grad', tg.gradient(grid[5]))
