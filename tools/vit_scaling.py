"""Throughput of k_viterbi versus the number of resident warps (32 codewords each): separates the ALU-pipe bound from the
dependency-latency bound.  Run on the GPU box: python tools/vit_scaling.py"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from abracadabra_b200 import pipeline

rx = pipeline.Receiver(0)
rx.set_profiling(True)
rng = np.random.default_rng(1)
steps = 774
for warps in (148, 296, 592, 1184, 2368, 4736):
    n_cw = warps * 32
    soft = rng.integers(-60, 60, (n_cw, steps * 4), dtype=np.int8)
    rx.stage_viterbi(soft)
    rx.stage_viterbi(soft)
    t = dict(rx.kernel_times())["k_viterbi(stage)"]
    print(f"warps {warps:5d} ({warps / 592:.2f}/SMSP): {t:.4f} ms  -> {warps * steps / t / 1e3:.1f} M warp-steps/s, "
          f"{t * 1e-3 * 1.965e9 / steps:.0f} cycles/step/warp-slot, {warps * steps * 32 * 64 / t / 1e9:.2f} T ACS/s")
