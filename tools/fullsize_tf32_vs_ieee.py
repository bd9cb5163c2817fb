"""Full-size (config 3) check of the TF32 block-order path against the IEEE fp32 path on the same data:
relative reconstruction error after the same number of sweeps.  python tools/fullsize_tf32_vs_ieee.py [sweeps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import rri_nmf_b200 as R
sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device('cuda:0')
cfg = dict(bench.CONFIGS['cfg3'])
X, W0, T0 = bench.gen_shard(torch, cfg, cfg["n"], 0, dev)
res = {}
for math in ('tf32', 'ieee'):
    eng = R.RRIEngine(X, 64, order='hals', math=math)
    W, T = W0.clone(), T0.clone()
    errs = []
    for s in range(sweeps):
        eng.sweeps(W, T, 1, eng.params())
        if s in (0, 1, 4, sweeps - 1):
            errs.append((s + 1, eng.rel_error(W, T)))
    res[math] = (errs, W, T)
    eng.close()
for (s, a), (_, b) in zip(res['tf32'][0], res['ieee'][0]):
    print('sweep %3d  rel_err tf32 %.8f  ieee %.8f  delta %+.2e' % (s, a, b, a - b))
Wt, Tt = res['tf32'][1], res['tf32'][2]
Wi, Ti = res['ieee'][1], res['ieee'][2]
print('||W_tf32 - W_ieee|| / ||W_ieee|| = %.3e   ||T_tf32 - T_ieee|| / ||T_ieee|| = %.3e'
      % (float((Wt - Wi).norm() / Wi.norm()), float((Tt - Ti).norm() / Ti.norm())))
