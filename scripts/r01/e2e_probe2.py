"""Where the end-to-end step goes: upload, input index build (K0), merge with streamed download."""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'bwt-merge_b200')
import numpy as np, torch
import bwtm_b200
from bwtm_b200 import FMI, MergeParameters, synth
bwtm_b200.set_device(0)
thr = synth.error_threshold(0.01)
A = FMI.synthetic(50_000_000, 42, 150, thr, [(1, 10_000_000)]); B = FMI.synthetic(50_000_000, 42, 150, thr, [(2, 10_000_000)])
ra = torch.from_numpy(A.rle()).pin_memory(); rb = torch.from_numpy(B.rle()).pin_memory()
out = torch.empty(700_000_000, dtype=torch.uint8).pin_memory().numpy()
dev = torch.empty(ra.numel(), dtype=torch.uint8, device="cuda")
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(5):
    t0 = sync(); dev.copy_(ra, non_blocking=True); t1 = sync()
    a = FMI.from_rle(ra.numpy()); t2 = sync()
    b = FMI.from_rle(rb.numpy()); t3 = sync()
    p = MergeParameters(); p.host_output = out; p.slab_symbols = int(__import__("os").environ.get("SLAB", "0"))
    m = FMI.merge(a, b, p); t4 = sync()
    t = m.timings
    print("it %d: raw H2D %.2f ms (%.1f GB/s)  create A %.2f  create B %.2f  merge %.2f (api total %.2f: search %.1f sort %.1f il %.1f enc %.1f idx %.1f)  step %.2f" %
          (it, (t1-t0)*1e3, ra.numel()/(t1-t0)/1e9, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, t.total_seconds*1e3, t.search_seconds*1e3, t.sort_seconds*1e3,
           t.interleave_seconds*1e3, t.encode_seconds*1e3, t.index_seconds*1e3, (t4-t1)*1e3), flush=True)
    m.close()
