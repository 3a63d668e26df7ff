"""k_viterbi(msc) time on the bench workload versus (four-lane threshold, persistent CTAs per SM).
Run on the GPU box: python tools/vit_split_sweep.py [seconds]"""
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
ref = None
for split in (1 << 30, 3000, 2000, 1500, 1000, 0):
    for ctas in (2, 3, 4):
        os.environ["B200DAB_VIT_CTAS"] = str(ctas)
        rx = pipeline.Receiver(0)
        rx.set_viterbi_split(split)
        rx.set_profiling(True)
        for _ in range(3):
            dec = rx.decode(dev.data_ptr(), "u8", tii=True, n_samples=raw.size // 2, on_device=True)
        sig = [f.tobytes() for f in dec.frames]
        if ref is None:
            ref = sig
        same = sig == ref
        ts = []
        for back in range(3):
            ts.append(dict(rx.kernel_times(back)))
        t = {k: float(np.median([x[k] for x in ts])) for k in ts[0]}
        tot = sum(t.values())
        print(f"split {split:>10d}, {ctas} CTAs/SM ({ctas} warps per sub-partition): k_viterbi(msc) {t['k_viterbi(msc)']:.4f} ms, fic {t['k_viterbi(fic)']:.4f} ms, "
              f"trace {t['k_vit_trace(msc)']:.4f}+{t['k_vit_trace(fic)']:.4f} ms, all kernels {tot:.3f} ms, frames identical to first config: {same}", flush=True)
        rx.close()
