import numpy as np, sys, copy, scipy.linalg as sl
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import informative_path_planning_b200 as pkg
from oracle import aq_oracle as aqo, gpmodel_oracle as gpo
from test_gpu_parity import world
aq, L = pkg.aq_library, pkg._lib
R=(0.,10.,0.,10.)
for noise in (0.5,):
    rng = np.random.default_rng(31)
    X, z = world(rng, 150, noise)
    d = pkg.OnlineGPModel(R, 1.0, 100.0, noise=noise); o = gpo.OnlineGPModelOracle(R, 1.0, 100.0, noise=noise)
    d.add_data(X, z); o.add_data(X, z)
    sizes, reps = [4, 1, 7, 4, 3], [2, 1, 2, 2, 2]
    levels=[]; start = rng.uniform(2, 8, 2)
    for k, r in zip(sizes, reps):
        pts = start + np.cumsum(rng.normal(0, 0.3, (k, 2)), axis=0); start = pts[-1]
        levels.append((pts, rng.normal(0, 10.0, (k, 1)), r))
    t=5
    got,info = d.rollout_rewards(L.REWARD_UCB, levels, params=[np.sqrt(aq.ucb_beta(t))])
    def K(A,B):
        dd=((A[:,None,:]-B[None,:,:])**2).sum(-1); return 100*np.exp(-0.5*dd)
    fork=copy.copy(d); Xa, za = X.copy(), z.copy()
    for i,(pts,zo,r) in enumerate(levels):
        md,vd = fork.predict_value(pts); mo,vo = o.predict_value(pts)
        Ky=K(Xa,Xa)+(noise+1e-8)*np.eye(len(Xa)); c=sl.cho_factor(Ky,lower=True)
        kx=K(Xa,pts); mu=kx.T@sl.cho_solve(c,za); var=100-np.einsum('ij,ij->j',kx,sl.cho_solve(c,kx))+noise
        sc = mu.sum()+np.sqrt(aq.ucb_beta(t))*np.abs(var).sum()
        sd = aq.mean_UCB(time=t,xvals=pts,robot_model=fork); so=aqo.mean_UCB(t,pts,o)
        print(i,'fused-scratch %.2e  devseq-scratch %.2e  oraseq-scratch %.2e | mean err dev %.2e ora %.2e var err dev %.2e ora %.2e'%(got[i]-sc, sd-sc, so-sc, np.abs(md.ravel()-mu.ravel()).max(), np.abs(mo.ravel()-mu.ravel()).max(), np.abs(vd.ravel()-var).max(), np.abs(vo.ravel()-var).max()))
        W=fork.woodbury_inv; print('   asym dev %.2e  asym ora %.2e'%(np.abs(W-W.T).max(), np.abs(o._woodbury_inv-o._woodbury_inv.T).max() if hasattr(o,'_woodbury_inv') else -1))
        for _ in range(r):
            fork.add_data(pts,zo); o.add_data(pts,zo); Xa=np.vstack([Xa,pts]); za=np.vstack([za,zo])
