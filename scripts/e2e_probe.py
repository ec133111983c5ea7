import sys, time, os
sys.path.insert(0, '.'); sys.path.insert(0, 'bwt-merge_b200')
import numpy as np, torch
import bwtm_b200
from bwtm_b200 import FMI, MergeParameters, synth
bwtm_b200.set_device(0)
thr = synth.error_threshold(0.01)
A = FMI.synthetic(50_000_000, 42, 150, thr, [(1, 10_000_000)]); B = FMI.synthetic(50_000_000, 42, 150, thr, [(2, 10_000_000)])
ra = torch.from_numpy(A.rle()).pin_memory().numpy(); rb = torch.from_numpy(B.rle()).pin_memory().numpy()
out = torch.empty(700_000_000, dtype=torch.uint8).pin_memory().numpy()
A.close(); B.close()
for it in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    a = FMI.from_rle(ra); torch.cuda.synchronize(); t1 = time.perf_counter()
    b = FMI.from_rle(rb); torch.cuda.synchronize(); t2 = time.perf_counter()
    m = FMI.merge(a, b); torch.cuda.synchronize(); t3 = time.perf_counter()
    n = m.download_into(out); torch.cuda.synchronize(); t4 = time.perf_counter()
    m.close(); torch.cuda.synchronize(); t5 = time.perf_counter()
    t = m.timings
    stages = t.search_seconds + t.sort_seconds + t.interleave_seconds + t.encode_seconds + t.index_seconds
    print("it %d: create A %.1f  create B %.1f  merge %.1f (stages %.1f, api total %.1f)  download %.1f  close %.1f ms" %
          (it, (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, stages*1e3, t.total_seconds*1e3, (t4-t3)*1e3, (t5-t4)*1e3), flush=True)
