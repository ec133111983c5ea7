import sys, time, os
sys.path.insert(0, '.'); sys.path.insert(0, 'bwt-merge_b200')
import numpy as np, torch
import bwtm_b200
from bwtm_b200 import FMI, MergeParameters, synth
bwtm_b200.set_device(0)
thr = synth.error_threshold(0.01)
A = FMI.synthetic(50_000_000, 42, 150, thr, [(1, 10_000_000)]); B = FMI.synthetic(50_000_000, 42, 150, thr, [(2, 10_000_000)])
out = torch.empty(700_000_000, dtype=torch.uint8).pin_memory().numpy()
for mode in ("plain", "stream", "stream_small_slab", "plain", "stream"):
    p = MergeParameters()
    if mode.startswith("stream"):
        p.host_output = out
    if mode == "stream_small_slab":
        p.slab_symbols = 1 << 28
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m = FMI.merge(A, B, p, keep_inputs=True); torch.cuda.synchronize(); t1 = time.perf_counter()
        if not mode.startswith("stream"):
            m.download_into(out)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        t = m.timings
        print("%-18s it %d: merge %.1f ms (api %.1f; enc %.1f idx %.1f)  download %.1f ms  total %.1f" %
              (mode, it, (t1-t0)*1e3, t.total_seconds*1e3, t.encode_seconds*1e3, t.index_seconds*1e3, (t2-t1)*1e3, (t2-t0)*1e3), flush=True)
        m.close()
