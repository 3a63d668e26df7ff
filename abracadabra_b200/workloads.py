"""Synthetic workloads of BASELINE.json configs[2..4] (SURVEY.md sec. 8 d), shared by bench.py and the full-size GPU tests:
fixed seeds per unit (seed = 1000 * config + index), so that the reference harness, the CPU checker and the CUDA path all see the
same files.

  config 3: TII batch   -- recordings of 40 transmission frames (+ lead-in), 1-8 transmitters (main 0-69, sub 1-23), amplitudes
                           0.1-1.0, per-TX delay 0-300 samples, one multipath tap, CFO uniform +-5 kHz, SNR 10-30 dB
  config 4: band scan   -- one 10 s recording per Band III channel 5A..13F (src/dabtables.cpp:32-39), different EId / label /
                           TII per channel, written as raw files named so that RawFileInput infers the frequency
                           (src/input/rawfileinput.cpp:170-191)
  config 5: long record -- s16 / f32 full ensemble with linear CFO drift and an SNR staircase 0 -> 20 dB; long versions are tiled
                           from one clean 600-frame signal (2400 CIFs: a multiple of the 8-CIF TII pattern and of the 5-CIF DAB+
                           superframe), noise and rotation generated per tile
"""
from __future__ import annotations


import numpy as np

from . import dabconst as C
from . import synth

TII_FRAMES = 43                      # 40 frames of one TII notification + the frames the core needs to synchronise
BAND3 = list(C.BAND3.items())        # [(channel name, kHz)] in band order


def tii_recording(index: int):
    """config 3, recording `index`: returns (raw u8, [TiiTx], Impairments)"""
    rng = np.random.default_rng(3000 + index)
    n_tx = int(rng.integers(1, 9))
    subs = rng.choice(np.arange(1, 24), n_tx, replace=False)
    txs = [synth.TiiTx(int(rng.integers(0, 70)), int(s), float(rng.uniform(0.1, 1.0)), delay=int(rng.integers(0, 300)),
                       phase=float(rng.uniform(0, 2 * np.pi))) for s in subs]
    txs[0].amp = max(txs[0].amp, 0.6)                     # at least one transmitter is strong
    ens = synth.config1_ensemble(seed=3000 + index)
    ens.tii = txs
    imp = synth.Impairments(snr_db=float(rng.uniform(10, 30)), cfo_hz=float(rng.uniform(-5000, 5000)),
                            taps=[(int(rng.integers(5, 60)), complex(rng.uniform(0.1, 0.4), rng.uniform(-0.3, 0.3)))],
                            seed=3000 + index, lead_in=int(rng.integers(20000, 120000)))
    raw, _ = synth.synthesize(ens, TII_FRAMES, "u8", imp, seed=3000 + index, tail=150000)
    return raw, txs, imp


def tii_null_sets(n_sets: int, seed: int = 3999):
    """config 3, kernel-level variant: n_sets x 40 null symbols as [n_sets, 40, 2048] complex64 (FFT windows, CFO already removed):
    in every set the even nulls carry the TII of 1-8 transmitters (random main / sub, amplitude 0.1-1, delay 0-300 samples, phase),
    the odd ones are the TII-free reference nulls; noise at 10-30 dB below the unit carrier.  500 sets = 10 000 TII-bearing nulls.
    Returns (nulls, [[(main, sub, amp)] per set])."""
    rng = np.random.default_rng(seed)
    out = np.zeros((n_sets, 40, 2048), np.complex64)
    truth = []
    for i in range(n_sets):
        n_tx = int(rng.integers(1, 9))
        subs = rng.choice(np.arange(1, 24), n_tx, replace=False)
        txs = [synth.TiiTx(int(rng.integers(0, 70)), int(s), float(rng.uniform(0.1, 1.0)), delay=int(rng.integers(0, 300)),
                           phase=float(rng.uniform(0, 6.28))) for s in subs]
        txs[0].amp = max(txs[0].amp, 0.6)
        out[i, 0::2] = synth.tii_symbol(synth.EnsembleConfig(tii=txs))[-2048:]
        sigma = np.float32(10 ** (-rng.uniform(10, 30) / 20) * np.sqrt(0.5))
        out[i] += (rng.standard_normal((40, 4096), dtype=np.float32) * sigma).view(np.complex64)
        truth.append([(t.main, t.sub, t.amp) for t in txs])
    return out, truth


def band_scan_ensemble(ch: int) -> synth.EnsembleConfig:
    name, _ = BAND3[ch]
    ens = synth.config1_ensemble(seed=4000 + ch)
    ens.eid = 0x4000 + ch
    ens.label = f"SCAN {name:<3s} MUX"[:16]
    ens.services[0].sid = 0x4100 + ch
    ens.services[0].label = f"SCAN {name} SVC"[:16]
    ens.tii = [synth.TiiTx((7 * ch + 3) % 70, ch % 23 + 1, 1.0), synth.TiiTx((7 * ch + 3) % 70, (ch + 11) % 23 + 1, 0.5, delay=60)]
    return ens


def band_scan_recording(ch: int, seconds: float = 10.0):
    """config 4, channel index ch (0 = 5A .. 37 = 13F): returns (raw u8, SynthResult, channel name, frequency kHz)"""
    name, khz = BAND3[ch]
    n_frames = int(round(seconds * C.FS / C.T_FRAME))
    imp = synth.Impairments(snr_db=18.0 + (ch * 5) % 13, cfo_hz=-3000.0 + 157.0 * ch, seed=4000 + ch, lead_in=30000 + 997 * ch)
    raw, res = synth.synthesize(band_scan_ensemble(ch), n_frames, "u8", imp, seed=4000 + ch, tail=90000)
    return raw, res, name, khz


def staircase_recording(n_frames: int, fmt: str, seed: int = 5000, drift_hz_per_s: float = 200.0 / 60.0, ens=None):
    """config 5 generated in one piece (tests: 120 s): full ensemble, linear CFO drift, SNR floor(21 f / n_frames) dB in frame f.
    Returns (raw, truth dict of synth.modulate)."""
    ens = ens or synth.full_ensemble(seed=2000 + seed)
    clean, truth = synth.modulate(ens, n_frames, seed=seed)
    x = synth.impair(clean, synth.Impairments(snr_db=None, cfo_hz=-800.0, cfo_drift_hz_per_s=drift_hz_per_s, seed=seed, lead_in=41000), tail=90000)
    del clean
    synth.add_noise_staircase(x, 41000, np.floor(np.arange(n_frames) * 21.0 / n_frames), seed=seed + 1)
    raw = synth.quantize(x, fmt, 6000.0 if fmt == "s16" else 1.0)
    return raw, truth


class TiledLongRecording:
    """config 5 at full length without holding 30 GB of synthesis state: one clean 600-frame tile (57.6 s, 2400 CIFs) of the full
    ensemble is modulated once; sample range [a, b) of the long recording is tile (n // tile_len) rotated by the drifting CFO and
    with noise at the SNR of its position, computed with torch on the GPU of the caller and returned as s16 (or f32) on the host.
    (At a tile boundary the time interleaver restarts, so the 15 logical frames across it do not decode -- a property of the test
    signal, identical for the single-pass and the chunked decode.)"""
    TILE_FRAMES = 600

    def __init__(self, seconds: float, seed: int = 5000, drift_hz_total: float = 12000.0, device: str = "cuda"):
        import torch
        self.torch = torch
        self.device = device
        self.n_samples = int(seconds * C.FS)
        self.seconds = seconds
        self.seed = seed
        self.f0 = -6000.0
        self.drift = drift_hz_total / max(seconds, 1.0)        # Hz per second: +-200 Hz/min for an hour
        clean, self.truth = synth.modulate(synth.full_ensemble(seed=2000 + seed), self.TILE_FRAMES, seed=seed)
        self.tile_len = clean.size
        assert self.tile_len == self.TILE_FRAMES * C.T_FRAME
        self.clean = torch.from_numpy(clean.view(np.float32).reshape(-1, 2).copy()).to(device)

    def snr_db(self, t_seconds):
        return np.floor(np.clip(t_seconds / self.seconds, 0, 0.999999) * 21.0)

    def chunk(self, a: int, b: int, fmt: str = "s16") -> np.ndarray:
        """samples [a, b) of the long recording; every sample is a function of its absolute index only (noise is generated per
        absolute block of 2^22 samples), so overlapping chunks carry identical samples in the overlap"""
        torch = self.torch
        b = min(b, self.n_samples)
        out = np.empty((b - a) * 2, np.int16 if fmt == "s16" else np.float32)
        blk = 1 << 22
        for k in range(a // blk, (b - 1) // blk + 1):
            s, e = max(a, k * blk), min(b, (k + 1) * blk)
            gen = torch.Generator(device=self.device)
            gen.manual_seed(self.seed * 1000003 + k)
            nz = torch.randn((blk, 2), device=self.device, generator=gen)[s - k * blk:e - k * blk]
            n = torch.arange(s, e, device=self.device, dtype=torch.float64)
            x = self.clean[(n % self.tile_len).long()]                            # [m, 2] float32
            t = n / C.FS
            ph = torch.remainder(2 * np.pi * (self.f0 * t + 0.5 * self.drift * t * t), 2 * np.pi).float()
            c, sn = torch.cos(ph), torch.sin(ph)
            snr = torch.floor(torch.clamp(t / self.seconds, 0, 0.999999) * 21.0).float()
            sg = torch.sqrt(0.5 * torch.pow(10.0, -snr / 10.0))
            y = torch.stack([x[:, 0] * c - x[:, 1] * sn, x[:, 0] * sn + x[:, 1] * c], 1) + nz * sg[:, None]
            if fmt == "s16":
                y = torch.clamp(torch.round(y * 6000.0), -32768, 32767).to(torch.int16)
            out[(s - a) * 2:(e - a) * 2] = y.reshape(-1).cpu().numpy()
        return out
