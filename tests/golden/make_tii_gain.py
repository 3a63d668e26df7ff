"""Measures the reference core's per-bin TII fold gain (its front-end frequency response summed over the four carrier
quarters) with a flat comb in the null symbol, and writes include/b200dab_tii_gain.h.

Black-box measurement of libdabsdr 4.0.1 (the 384-bin TII spectrum it hands out through getSpectrumTii,
dabsdr.h:383 / src/radiocontrol.cpp:2485).  Run in the build container:  python tests/golden/make_tii_gain.py
"""
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE.parent))
import orc  # noqa: E402
from abracadabra_b200 import dabconst as C, synth  # noqa: E402


def full_comb_symbol(a):
    spec = np.zeros(2048, complex)
    prs, kk = C.prs_carriers(), C.carrier_k()
    for c in range(0, 1536, 2):
        spec[kk[c] % 2048] = a * prs[c]
        spec[kk[c + 1] % 2048] = a * prs[c]
    t = np.fft.ifft(spec) * 2048 / np.sqrt(1536)
    return np.concatenate([t[-608:], t]).astype(np.complex64)


ens = synth.config1_ensemble()
ens.tii = [synth.TiiTx(12, 5, 1.0)]
orig = synth.tii_symbol
synth.tii_symbol = lambda e: full_comb_symbol(0.15)
raw, res = synth.synthesize(ens, 42, "s16", synth.Impairments(snr_db=None, cfo_hz=0.0, rms=7000), tail=150000)
synth.tii_symbol = orig
with tempfile.TemporaryDirectory() as td:
    p = Path(td) / "comb.s16"
    raw.tofile(p)
    _, ev = orc.run_ref(p, "s16", Path(td) / "log.jsonl", tii=1)
sp = np.array([e for e in ev if e["ev"] == "tii"][1]["spectrum"])
g = sp / sp.mean()
lines = ["/* b200dab_tii_gain.h -- per-bin amplitude gain of the reference core's TII spectrum (2 x 192 carrier pairs, each the",
         " * sum over the four 384-carrier quarters), normalised to mean 1.  MEASURED from libdabsdr 4.0.1 as a black box with a",
         " * flat comb in the null symbol (tests/golden/make_tii_gain.py); it is that core's front-end filter ripple (+-6.5 %),",
         " * reproduced here so that TII levels and spectra come out as RadioControl is used to seeing them. */",
         "#ifndef B200DAB_TII_GAIN_H", "#define B200DAB_TII_GAIN_H", "static const float B200DAB_TII_GAIN[384] = {"]
for i in range(0, 384, 12):
    lines.append("    " + ", ".join(f"{v:.4f}f" for v in g[i:i + 12]) + ",")
lines += ["};", "#endif", ""]
(ROOT / "include" / "b200dab_tii_gain.h").write_text("\n".join(lines))
print("gain min/max", g.min(), g.max())
