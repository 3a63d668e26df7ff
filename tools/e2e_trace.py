"""Host-side timeline of the pipelined upload (B200DAB_TRACE=1): python tools/e2e_trace.py"""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ["B200DAB_TRACE"] = "1"
import numpy as np, torch
import bench
from abracadabra_b200 import pipeline
raw, res, nf = bench.make_recording(60.0, 0)
rx = pipeline.Receiver(0)
h = torch.from_numpy(raw).pin_memory()
d = torch.from_numpy(raw).cuda()
n = raw.size // 2
for i in range(3):
    t = time.perf_counter(); rx.decode_raw(h.data_ptr(), n, "u8", False, None, 1); print("host-buffer decode", (time.perf_counter() - t) * 1e3, "ms", file=sys.stderr)
t = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); print("plain pinned H2D of the recording", (time.perf_counter() - t) * 1e3, "ms", file=sys.stderr)
t = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); print("plain pinned H2D of the recording", (time.perf_counter() - t) * 1e3, "ms", file=sys.stderr)
for i in range(2):
    t = time.perf_counter(); rx.decode_raw(d.data_ptr(), n, "u8", True, None, 1); print("device-buffer decode", (time.perf_counter() - t) * 1e3, "ms", file=sys.stderr)
