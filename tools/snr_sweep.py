#!/usr/bin/env python
"""Sensitivity sweep of the CUDA path next to the reference core (the reference's cliff is between 3 and 4 dB):
one 10 s DAB+ recording (96 kbps EEP-3A), AWGN at 3.0 .. 7.0 dB in 0.5 dB steps, both decoders on the same files.
Per SNR: access units with a good CRC, pre-RS bit errors (rsBitErrors), uncorrectable RS words.

    python tools/snr_sweep.py > profiles/rNN_snr_sweep.json          (needs a GPU and oracle/_ref)
"""
import json
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def sweep(rx, orc, snrs=np.arange(3.0, 7.01, 0.5), n_frames=104):
    from abracadabra_b200 import synth
    ens = synth.config1_ensemble(seed=77)
    ens.tii = []
    clean, truth = synth.modulate(ens, n_frames, seed=77)
    base = synth.impair(clean, synth.Impairments(snr_db=None, cfo_hz=321.0, seed=7, lead_in=40000), tail=120000)
    sent = set(truth["aus"][ens.subchannels[0].subchid])
    rows = []
    with tempfile.TemporaryDirectory() as td:
        for snr in snrs:
            rng = np.random.default_rng(int(snr * 10))
            sg = np.float32(np.sqrt(0.5 * 10 ** (-snr / 10)))
            x = base + (rng.standard_normal(2 * base.size, dtype=np.float32) * sg).view(np.complex64)
            raw = synth.quantize(x, "u8", 28.0)
            p = Path(td) / "s.u8"
            raw.tofile(p)
            _, ev = orc.run_ref(p, "u8", Path(td) / "log.jsonl", service=0, extra=["--tail-frames", "12"])
            aus = [e for e in ev if e["ev"] == "audio" and e["len"] > 0]
            good = [bytes.fromhex(e["data"]) for e in aus if not (e["header"] & 0x80)]      # bit 7 = conceal (dabsdr.h:47-60)
            per = [e for e in ev if e["ev"] == "periodic" and not e.get("null")]
            dec = rx.decode(raw, "u8")
            mine = [a.data for a in dec.aus if a.crc_ok]
            rows.append(dict(snr_db=float(snr), ref_aus_good=len(good), ref_aus_concealed=len(aus) - len(good),
                             ref_rs_bit_errors=int(sum(e["rsBitErr"] for e in per)), ref_rs_uncorrectable=int(sum(e["rsUnc"] for e in per)),
                             gpu_aus_good=len(mine), gpu_aus_bad_crc=len(dec.aus) - len(mine), gpu_rs_bit_errors=int(dec.rs_bit_errors),
                             gpu_rs_uncorrectable=int(dec.rs_uncorrectable), gpu_frames=int(dec.n_frames),
                             gpu_aus_not_transmitted=len([a for a in mine if a not in sent]),
                             ref_good_aus_missing_on_gpu=len(set(good[:-6]) - set(mine))))      # but for the tail the reference flushes out of the zero fill
    return rows, 3 * ((4 * n_frames - 15) // 5)


if __name__ == "__main__":
    import orc
    from abracadabra_b200 import pipeline
    rx = pipeline.Receiver(0)
    rows, possible = sweep(rx, orc)
    rx.close()
    print(json.dumps(dict(what="10 s DAB+ 96 kbps EEP-3A, AWGN, same files through libdabsdr 4.0.1 (fp32 soft bits) and the CUDA path (int8 soft bits, int16 metrics)",
                          access_units_in_the_recording=possible, rows=rows), indent=1))
