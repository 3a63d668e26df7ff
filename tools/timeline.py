"""Kernel timeline of one b200dab_decode on the bench workload (device-resident input): where each kernel starts on the stream,
how long it runs, and the gap to the previous one.  Run on the GPU box: python tools/timeline.py [seconds]"""
import sys
import time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from abracadabra_b200 import pipeline, synth

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
n_frames = int(secs * 2048000 / 196608) + 1
ens = synth.full_ensemble()
raw, _ = synth.synthesize(ens, n_frames, "u8", synth.Impairments(snr_db=25.0, cfo_hz=431.0, seed=7), seed=7, tail=20000)
raw = raw[:int(secs * 2048000) * 2]
import torch
dev = torch.from_numpy(raw).cuda()
rx = pipeline.Receiver(0)
rx.set_profiling(True)
for _ in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rx.decode_raw(dev.data_ptr(), raw.size // 2, "u8", True, flags=pipeline.OPT_TII)
    wall = (time.perf_counter() - t0) * 1e3
tl = rx.kernel_timeline()
end = 0.0
for name, start, dur in tl:
    print(f"{start:8.3f} ms  +{dur:7.3f} ms  gap {start - end:7.3f} ms  {name}")
    end = start + dur
print(f"last kernel ends {end:.3f} ms after the first begins; kernels {sum(d for _, _, d in tl):.3f} ms; host wall time of the call {wall:.3f} ms")
