"""Probe of the tcgen05 TF32 contraction: accuracy and SIGNED bias against an fp64 product, for the two
TMA element types (TFLOAT32 / FLOAT32, env RRI_TMA_F32) and both row-tile multiplicities (RRI_GEMM_MT).
Run on a B200:  python tools/tf32_probe.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
import rri_nmf_b200 as R
dev = torch.device('cuda:0')
torch.manual_seed(0)
for (M, N, K) in [(128, 64, 64), (256, 64, 32), (1000, 64, 2000), (4096, 128, 3108), (300, 10, 500), (257, 50, 1028), (20000, 64, 20000)]:
    A = torch.rand(M, K, device=dev); B = torch.rand(N, K, device=dev)
    eng = R.RRIEngine(torch.zeros(8, 8, device=dev), N, order='hals', math='tf32')
    C = eng.gemm_nt(A, B); torch.cuda.synchronize()
    Cr = A.double() @ B.double().t()
    rel = float((C.double() - Cr).norm() / Cr.norm())
    bias = float(((C.double() - Cr) / Cr).mean())
    print('M=%%d N=%%d K=%%d relfro=%%.3e mean_signed_rel=%%+.3e max_abs_rel=%%.3e' %% (M, N, K, rel, bias, float(((C.double()-Cr)/Cr).abs().max())), flush=True)
    eng.close()
''' % ROOT

for f32 in ('0', '1'):
    for mt in ('1', '2'):
        env = dict(os.environ, RRI_TMA_F32=f32, RRI_GEMM_MT=mt)
        print('=== RRI_TMA_F32=%s RRI_GEMM_MT=%s' % (f32, mt), flush=True)
        try:
            r = subprocess.run([sys.executable, '-c', CHILD], env=env, timeout=120, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
            print(r.stdout[-3000:], 'rc=%d' % r.returncode, flush=True)
        except subprocess.TimeoutExpired as e:
            print('TIMEOUT', (e.stdout or '')[-2000:], flush=True)
