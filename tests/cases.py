"""Seeded synthetic inputs shared by the parity tests (sizes the CPU checker finishes in seconds)."""
from functools import lru_cache

import numpy as np

from abracadabra_b200 import synth


@lru_cache(maxsize=8)
def config1(n_frames=10, fmt="u8", snr=30.0, cfo=123.4, seed=11, tii=True, lead_in=49152):
    ens = synth.config1_ensemble(seed=1000 + seed)
    if not tii:
        ens.tii = []
    raw, res = synth.synthesize(ens, n_frames, fmt, synth.Impairments(snr_db=snr, cfo_hz=cfo, seed=seed, lead_in=lead_in),
                                seed=seed, tail=150000)
    return raw, res


@lru_cache(maxsize=4)
def multi(n_frames=8, fmt="u8", snr=25.0, cfo=-2345.6, seed=21):
    """four sub-channels with different protection profiles (EEP-A, EEP-B, UEP)"""
    subs = [synth.SubChannel(1, 0, "dabplus", 96, ("EEP", 3, "A"), seed=seed),
            synth.SubChannel(2, 72, "dabplus", 64, ("EEP", 2, "A"), seed=seed + 1),
            synth.SubChannel(5, 140, "dabplus", 64, ("EEP", 3, "B"), seed=seed + 2),
            synth.SubChannel(9, 200, "mp2", 128, ("UEP", 35), seed=seed + 3),
            synth.SubChannel(11, 300, "dabplus", 32, ("EEP", 1, "A"), seed=seed + 4)]
    svcs = [synth.Service(0x3000 + i, f"MULTI {i}", s.subchid) for i, s in enumerate(subs)]
    ens = synth.EnsembleConfig(eid=0x3003, label="B200 MULTI", subchannels=subs, services=svcs, tii=[synth.TiiTx(3, 9, 1.0)])
    raw, res = synth.synthesize(ens, n_frames, fmt, synth.Impairments(snr_db=snr, cfo_hz=cfo, seed=seed), seed=seed, tail=150000)
    return raw, res
