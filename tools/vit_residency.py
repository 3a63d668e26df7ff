"""k_viterbi(msc) time on the bench workload versus CTA residency (limited through extra dynamic shared memory).
Run on the GPU box: python tools/vit_residency.py [seconds]"""
import os
import sys
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
for kb in (0, 4, 6, 9, 14, 18, 23, 32, 51):
    os.environ["B200DAB_VIT_DSMEM"] = str(kb * 1024)
    rx = pipeline.Receiver(0)
    rx.set_profiling(True)
    for _ in range(3):
        rx.decode(dev.data_ptr(), "u8", tii=True, n_samples=raw.size // 2, on_device=True)
    t = dict(rx.kernel_times())
    tot = sum(v for _, v in rx.kernel_times())
    print(f"dsmem {kb:4d} KB (<= {min(32, 227 // (kb + 4 + 1))} one-warp CTAs/SM): k_viterbi(msc) {t['k_viterbi(msc)']:.4f} ms, fic {t['k_viterbi(fic)']:.4f} ms, trace {t['k_vit_trace(msc)']:.4f}+{t['k_vit_trace(fic)']:.4f} ms, all kernels {tot:.3f} ms", flush=True)
    rx.close()
