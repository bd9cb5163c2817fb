"""Device-resident NNDSVD initialisation at config-3 size: the 2*n_iter + 2 passes over X through
the engine's contraction kernel (TF32 tcgen05 / IEEE SIMT) next to library GEMMs.   python tools/bench_init.py [rows]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import rri_nmf_b200 as R
from rri_nmf_b200._device_init import initialize_nmf_torch

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
dev = torch.device('cuda:0')
cfg = dict(bench.CONFIGS['cfg3']); cfg['n'] = rows
X, _, _ = bench.gen_shard(torch, cfg, rows, 0, dev)
k = cfg['k']
for label, math in (('tf32 tcgen05 contraction', 'tf32'), ('ieee simt contraction', 'ieee'), ('library GEMM (torch.matmul)', None)):
    eng = R.RRIEngine(X, k, order='hals', math=math or 'tf32')
    prod = eng.products() if math else None
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        W, T = initialize_nmf_torch(X, k, 'nndsvd', random_state=0, products=prod)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    Wd, Td = W.contiguous(), T.contiguous()
    print('%-30s NNDSVD %dx%d k=%d: %.3f s   rel_err of the initial factors %.5f' % (label, rows, cfg['d'], k, dt, eng.rel_error(Wd, Td)), flush=True)
    eng.close()
