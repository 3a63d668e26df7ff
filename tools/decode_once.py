"""Two decodes of the bench workload (60 s full ensemble) for profiler captures: python tools/decode_once.py [seconds]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from abracadabra_b200 import pipeline, synth

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
n_frames = int(secs * 2048000 / 196608) + 1
raw, _ = synth.synthesize(synth.full_ensemble(), n_frames, "u8", synth.Impairments(snr_db=25.0, cfo_hz=431.0, seed=7), seed=7, tail=20000)
raw = raw[:int(secs * 2048000) * 2]
dev = torch.from_numpy(raw).cuda()
rx = pipeline.Receiver(0)
rx.set_profiling(True)
for _ in range(2):
    dec = rx.decode(dev.data_ptr(), "u8", tii=True, n_samples=raw.size // 2, on_device=True)
print("frames", dec.n_frames, "fib ok", int(dec.fib_ok.sum()), "aus", len(dec.aus))
for k, v in rx.kernel_times():
    print(f"{k:28s} {v:.4f} ms")
