"""One k_viterbi launch on 296 groups x 774 steps (for ncu)."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from abracadabra_b200 import pipeline
rx = pipeline.Receiver(0)
rng = np.random.default_rng(1)
soft = rng.integers(-60, 60, (296 * 32, 774 * 4), dtype=np.int8)
rx.stage_viterbi(soft)
rx.stage_viterbi(soft)
print("ok")
