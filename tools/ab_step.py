"""A/B timing of whole decode steps on ONE box (boxes differ by several per cent, so variants are only comparable inside one run):
every variant = a set of environment variables read by b200dab_create; for each, a fresh receiver decodes the bench workload
(device-resident, profiling events off) `reps` times, interleaved round-robin with the other variants; median / min ms per step.
    python tools/ab_step.py "name1:VAR=1,VAR2=3" "name2:" ...  [--seconds 60] [--reps 12] [--host]
--host: the recording starts in pinned host memory (end-to-end step, one context, sequential)."""
import os
import sys
import time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from abracadabra_b200 import pipeline

args = [a for a in sys.argv[1:] if ":" in a]
secs = float(sys.argv[sys.argv.index("--seconds") + 1]) if "--seconds" in sys.argv else 60.0
reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 12
host = "--host" in sys.argv
raw, res, nf = bench.make_recording(secs, 0)
n = raw.size // 2
d = torch.from_numpy(raw).cuda()
h = torch.from_numpy(raw).pin_memory()
variants = []
for a in args or ["base:"]:
    name, _, envs = a.partition(":")
    env = dict(kv.split("=") for kv in envs.split(",") if kv)
    keep = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    rx = pipeline.Receiver(0)
    for k, v in keep.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    variants.append((name, rx, []))
ref = None
for name, rx, _ in variants:                       # warm-up + results identical across variants
    for _ in range(3):
        dec = rx.decode(d.data_ptr(), "u8", tii=True, n_samples=n, on_device=True)
    sig = (dec.n_frames, [f.tobytes() for f in dec.frames], [(a.subch_index, a.data) for a in dec.aus], dec.fibs.tobytes())
    if ref is None:
        ref = sig
    print(f"{name}: frames {dec.n_frames}, aus {len(dec.aus)}, identical to first variant: {sig == ref}", flush=True)
ptr, on_dev = (h.data_ptr(), False) if host else (d.data_ptr(), True)
for r in range(reps):
    for name, rx, ts in variants:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            rx.decode_raw(ptr, n, "u8", on_dev, None, pipeline.OPT_TII)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3 / 5)
for name, rx, ts in variants:
    print(f"{name:24s} median {np.median(ts):.4f} ms  min {min(ts):.4f} ms  max {max(ts):.4f} ms per step ({'host' if host else 'device'}-resident input)", flush=True)
