"""Host -> device staging of a large PAGEABLE array: thread count / chunk size sweep of nmf._pageable_to_device next to
the pinned copy and to cudaHostRegister-in-place.   python tools/h2d_probe.py [GB]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import importlib
import rri_nmf_b200
N = importlib.import_module("rri_nmf_b200.nmf")

gb = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
dev = torch.device('cuda:0')
rows = int(gb * 1e9 / (20000 * 4))
a = np.empty((rows, 20000), dtype=np.float32)
a[...] = 1.0
t = torch.from_numpy(a)
nbytes = a.nbytes
torch.cuda.synchronize()
tp = torch.empty(a.shape, dtype=torch.float32).pin_memory()
tp.copy_(t)
for _ in range(2):
    t0 = time.perf_counter(); d = tp.to(dev, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print('pinned              %.3f s  %.1f GB/s' % (dt, nbytes / dt / 1e9), flush=True)
del d, tp
t0 = time.perf_counter(); d = t.to(dev); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print('pageable torch .to  %.3f s  %.1f GB/s' % (dt, nbytes / dt / 1e9), flush=True)
del d
for threads in (4, 8, 12, 16):
    for chunk_mb in (8, 32):
        N._STAGE_THREADS = threads
        N._STAGE_CHUNK_BYTES = chunk_mb << 20
        N._STAGE['bufs'] = None
        N._pageable_to_device(t[:4096], dev)          # allocate the staging buffers (untimed)
        best = 1e9
        for _ in range(2):
            t0 = time.perf_counter(); d = N._pageable_to_device(t, dev); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
            ok = bool((d[-1] == 1).all())
            del d
        print('staged threads=%2d chunk=%3d MB  %.3f s  %.1f GB/s ok=%s' % (threads, chunk_mb, best, nbytes / best / 1e9, ok), flush=True)
rt = torch.cuda.cudart()
t0 = time.perf_counter()
rc = rt.cudaHostRegister(t.data_ptr(), nbytes, 0)
t1 = time.perf_counter()
d = torch.empty(a.shape, dtype=torch.float32, device=dev)
d.copy_(t, non_blocking=True); torch.cuda.synchronize()
t2 = time.perf_counter()
rt.cudaHostUnregister(t.data_ptr())
t3 = time.perf_counter()
print('hostRegister rc=%s: register %.3f s, copy %.3f s (%.1f GB/s), unregister %.3f s, total %.3f s' %
      (rc, t1 - t0, t2 - t1, nbytes / (t2 - t1) / 1e9, t3 - t2, t3 - t0), flush=True)
