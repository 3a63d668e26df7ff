#!/usr/bin/env python
"""Decode a raw IQ recording (.raw u8 / s16 or XML-header .uff as AbracaDABra records them) on the GPU and print what the
receiver found: python tools/decode_file.py <file> [--fmt u8|s16]"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from abracadabra_b200 import pipeline, rawfile  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("file")
ap.add_argument("--fmt", default=None)
a = ap.parse_args()
iq, info = rawfile.read_raw(a.file, a.fmt)
rx = pipeline.Receiver(0)
dec = rx.decode(iq, info.fmt, tii=True, symstat=True)
print(f"{a.file}: {info.num_samples / 2.048e6:.2f} s, {info.fmt}, {info.frequency_khz} kHz, header={info.has_xml_header}")
print(f"frames {dec.n_frames}/{dec.n_slots}, FIB CRC ok {int(dec.fib_ok.sum())}/{dec.fib_ok.size}, CFO {dec.cfo_hz.mean() if dec.n_slots else 0:.1f} Hz, SNR {dec.snr_db:.1f} dB")
if dec.n_frames:
    ens = pipeline.parse_ensemble(dec.fibs, dec.fib_ok)
    print(f"ensemble {ens['ueid']:06X} '{ens['label'].rstrip()}' LTO {ens['lto']}, {len(ens['services'])} services")
    for s in ens["services"]:
        print(f"  {s['sid']:04X} '{s['label'].rstrip()}' sub-channels {s['components']}")
    for i, sc in enumerate(dec.subch):
        n_au = sum(1 for u in dec.aus if u.subch_index == i)
        print(f"  subch {sc.subch_id:2d} @CU {sc.start_cu:3d} +{sc.size_cu:3d} {sc.bitrate:3d} kbps {'DAB+' if sc.kind else 'DAB '}: {dec.frames[i].shape[0]} frames, {n_au} AUs")
    print(f"RS: corrected {dec.rs_corrected} symbols / {dec.rs_bit_errors} bits, uncorrectable {dec.rs_uncorrectable}")
    for w in dec.tii:
        print("TII:", [(m, s, round(l, 3)) for m, s, l in w.ids])
