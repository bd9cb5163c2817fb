"""Timing of the masked WRRI path at config-4 shape (or a row subsample): python tools/bench_masked.py [rows]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rri_nmf_b200 as R
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
order = sys.argv[2] if len(sys.argv) > 2 else 'rri'
math = sys.argv[3] if len(sys.argv) > 3 else 'tf32'
d, k = 20000, 50
dev = torch.device('cuda:0')
g = torch.Generator(device=dev); g.manual_seed(0)
U = torch.rand(rows, k, generator=g, device=dev); V = torch.rand(k, d, generator=g, device=dev)
X = U @ V
X += 0.05 * X.mean() * torch.rand(rows, d, generator=g, device=dev)
M = (torch.rand(rows, d, generator=g, device=dev) < 0.05).to(torch.uint8)
W = torch.rand(rows, k, generator=g, device=dev); T = torch.rand(k, d, generator=g, device=dev)
eng = R.RRIEngine(X, k, W_mat=M, order=order, math=math)
p = eng.params(ub_t=1.0)
eng.sweeps(W, T, 1, p); torch.cuda.synchronize()
t0 = time.perf_counter(); eng.sweeps(W, T, 1, p); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print('math=%s' % math, 'masked %s rows=%d d=%d k=%d 5%% mask (u8): %.3f s/sweep = %.3f sweeps/s; rel_err=%.4f' % (order, rows, d, k, dt, 1 / dt, eng.rel_error(W, T)))
