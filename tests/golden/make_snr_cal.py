"""Calibration of the SNR figure (dabsdrNtfPeriodic_t.snr10, dabsdr.h:153; shown by the application, averaged by the scanner:
src/radiocontrol.cpp:302-319, src/tii/scannerbackend.cpp:819-834) against the reference core as a black box: for AWGN from 2 to 45 dB
the reference's snr10 (median over the periodic notifications of a 30-frame file) next to this path's decision-directed estimate
(10 log10 of the mean over symbols of mean|soft|^2 / var|soft|, the quantity b200dab_decode accumulates from k_ofdm_demod's symbol
statistics; computed here with the CPU checker, which keeps the same statistics).  Writes include/b200dab_snr_cal.h: a monotone
table (raw estimate -> reference scale), interpolated linearly.  Run in the build container:  python tests/golden/make_snr_cal.py"""
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE.parent))
import orc  # noqa: E402
from abracadabra_b200 import synth  # noqa: E402


def main():
    assert orc.have_ref()
    ens = synth.config1_ensemble(seed=5)
    ens.tii = []
    clean, _ = synth.modulate(ens, 30, seed=5)
    rows = []
    for snr in list(range(2, 31)) + [32, 34, 36, 38, 40, 45]:
        x = synth.impair(clean.copy(), synth.Impairments(snr_db=float(snr), cfo_hz=200.0, seed=snr, lead_in=40000), tail=60000)
        raw = synth.quantize(x, "s16", 6000.0)
        with tempfile.TemporaryDirectory() as td:
            p = Path(td) / "s.s16"
            raw.tofile(p)
            _, ev = orc.run_ref(p, "s16", Path(td) / "log.jsonl")
        per = [e["snr10"] for e in ev if e["ev"] == "periodic" and not e.get("null") and e["level"] == 3][1:]
        d = orc.decode(raw, "s16", [])
        if not per or d["nf"] < 10:
            continue
        acc = []
        for f in range(2, min(d["nf"], 24)):
            _, _, stat = orc.demod_frame(raw, "s16", d["t_f"][f], d["cfo"][f], want_carriers=True)
            mu, m2 = stat[:, 0], stat[:, 1]
            acc += list(mu * mu / (m2 - mu * mu))
        rows.append((float(10 * np.log10(np.mean(acc))), float(np.median(per)) / 10.0, snr))
        print(rows[-1], flush=True)
    rows.sort()
    keep = []
    for r in rows:                                   # monotone in both columns
        if not keep or (r[0] > keep[-1][0] + 1e-3 and r[1] >= keep[-1][1]):
            keep.append(r)
    lines = ["/* b200dab_snr_cal.h -- this path's decision-directed SNR estimate mapped onto the scale of the reference core's snr10",
             " * (dabsdr.h:153).  MEASURED from libdabsdr 4.0.1 as a black box over AWGN files from 2 to 45 dB",
             " * (tests/golden/make_snr_cal.py): {raw estimate dB, reference dB}, ascending; interpolate linearly, extrapolate with slope 1. */",
             "#ifndef B200DAB_SNR_CAL_H", "#define B200DAB_SNR_CAL_H", f"#define B200DAB_SNR_CAL_N {len(keep)}",
             "static const float B200DAB_SNR_CAL[B200DAB_SNR_CAL_N][2] = {"]
    lines += ["    " + ", ".join(f"{{{a:.2f}f, {b:.2f}f}}" for a, b, _ in keep[i:i + 6]) + "," for i in range(0, len(keep), 6)]
    lines += ["};", "static inline float b200dab_snr_calibrated(float raw_db)", "{",
              "    const int n = B200DAB_SNR_CAL_N;",
              "    if (raw_db <= B200DAB_SNR_CAL[0][0]) return raw_db + (B200DAB_SNR_CAL[0][1] - B200DAB_SNR_CAL[0][0]);",
              "    for (int i = 1; i < n; i++)",
              "        if (raw_db <= B200DAB_SNR_CAL[i][0]) {",
              "            const float t = (raw_db - B200DAB_SNR_CAL[i - 1][0]) / (B200DAB_SNR_CAL[i][0] - B200DAB_SNR_CAL[i - 1][0]);",
              "            return B200DAB_SNR_CAL[i - 1][1] + t * (B200DAB_SNR_CAL[i][1] - B200DAB_SNR_CAL[i - 1][1]);",
              "        }",
              "    return raw_db + (B200DAB_SNR_CAL[n - 1][1] - B200DAB_SNR_CAL[n - 1][0]);", "}", "#endif", ""]
    (ROOT / "include" / "b200dab_snr_cal.h").write_text("\n".join(lines))


if __name__ == "__main__":
    main()
