"""Timing of the observed-entries (sparse) WRRI path at config-4 shape, next to the dense masked path on the same
data:  python tools/bench_sparse.py [rows] [order] [sweeps] [refresh_every] [--dense] [--pattern=S]
Prints one JSON line per path; bytes/sweep is the algorithmic figure of DESIGN.md (2k passes of 10 B per entry:
2 B block-local index + 4 B residual read + 4 B residual written)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rri_nmf_b200 as R

args = [a for a in sys.argv[1:] if not a.startswith('--')]
rows = int(args[0]) if len(args) > 0 else 100000
order = args[1] if len(args) > 1 else 'rri'
sweeps = int(args[2]) if len(args) > 2 else 5
refresh_every = int(args[3]) if len(args) > 3 else 1
d, k, density = 20000, 50, 0.05
dev = torch.device('cuda:0')
g = torch.Generator(device=dev); g.manual_seed(0)
U = torch.rand(rows, k, generator=g, device=dev); V = torch.rand(k, d, generator=g, device=dev)
X = U @ V
X += 0.05 * X.mean() * torch.rand(rows, d, generator=g, device=dev)
pattern = [a for a in sys.argv[1:] if a.startswith('--pattern=')]
if pattern:
    # regular pattern (i + j) % S == 0: consecutive entries of a row / column are S apart, so the shared-memory gather
    # of 8 consecutive lanes is conflict-free for odd S and an 8-way conflict for S % 8 == 0 (experiment)
    S = int(pattern[0].split('=')[1])
    ii = torch.arange(rows, device=dev).view(-1, 1); jj = torch.arange(d, device=dev).view(1, -1)
    M = ((ii + jj) % S) == 0
    del ii, jj
else:
    M = torch.rand(rows, d, generator=g, device=dev) < density
W0 = torch.rand(rows, k, generator=g, device=dev); T0 = torch.rand(k, d, generator=g, device=dev)
Xs = (X * M).to_sparse_csr()
nnz = int(Xs.values().numel())
del U, V


def timed(eng, label, extra):
    W, T = W0.clone(), T0.clone()
    p = eng.params(ub_t=1.0, sp_refresh_every=refresh_every)
    eng.sweeps(W, T, 1, p)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(); eng.sweeps(W, T, sweeps, p, want_flags=False); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / sweeps
    out = {'path': label, 'pattern': (pattern[0] if pattern else 'random'), 'order': order, 'refresh_every': refresh_every, 'rows': rows, 'd': d, 'k': k, 'nnz': nnz, 'ms_per_sweep': round(ms, 3),
           'sweeps_per_s': round(1000.0 / ms, 2), 'rel_err': round(eng.rel_error(W, T), 5)}
    out.update(extra(ms))
    print(json.dumps(out), flush=True)
    return W, T


t0 = time.perf_counter()
es = R.RRIEngine(Xs, k, order=order)
torch.cuda.synchronize()
bind_s = time.perf_counter() - t0
alg = 2 * k * 10 * nnz          # 2k passes, each: 2 B index + 4 B residual read + 4 B residual written per entry
Ws, Ts = timed(es, 'sparse', lambda ms: {'bind_s': round(bind_s, 3), 'algorithmic_GB_per_sweep': round(alg / 1e9, 2),
                                          'achieved_GBps': round(alg / ms / 1e6, 1)})
es.close()
if '--dense' in sys.argv:
    ed = R.RRIEngine(X, k, W_mat=M.to(torch.uint8), order=order, math='tf32')
    Wd, Td = timed(ed, 'dense-masked-tf32', lambda ms: {})
    ed.close()
