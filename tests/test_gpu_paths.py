"""GPU tests of the host-side paths around the kernels: chunked/pipelined upload vs one-range decode, lost frames, empty and
truncated inputs, CFO drift with an SNR staircase (BASELINE.json configs[4] in small), a pool of receivers over several
recordings (configs[3], band scan) -- all through the C ABI, checked against the modulator's payload and the CPU checker."""
import numpy as np
import pytest

import cases
from abracadabra_b200 import pipeline, synth

pytestmark = pytest.mark.gpu


def _same(a: pipeline.Decoded, b: pipeline.Decoded):
    assert a.n_frames == b.n_frames and a.n_slots == b.n_slots
    assert np.array_equal(a.t_f, b.t_f) and np.array_equal(a.fibs, b.fibs) and np.array_equal(a.fib_ok, b.fib_ok)
    assert np.allclose(a.cfo_hz, b.cfo_hz, atol=1e-3)
    assert len(a.frames) == len(b.frames) and all(np.array_equal(x, y) for x, y in zip(a.frames, b.frames))
    assert [(u.subch_index, u.cif, u.crc_ok, u.data) for u in a.aus] == [(u.subch_index, u.cif, u.crc_ok, u.data) for u in b.aus]


def test_chunked_host_upload_equals_single_range(receiver):
    """> 128 frames from a host buffer go through the pipelined front end (64-frame chunks with overlap); the same samples
    as a device buffer go through one range.  Frames cut by a chunk boundary must be found exactly once."""
    import torch
    ens = synth.config1_ensemble(seed=5)
    raw, res = synth.synthesize(ens, 150, "u8", synth.Impairments(snr_db=20.0, cfo_hz=-777.0, seed=4, lead_in=12345), seed=4, tail=70000)
    host = receiver.decode(raw, "u8", tii=True)
    d = torch.from_numpy(raw).cuda()
    dev = receiver.decode(d.data_ptr(), "u8", tii=True, n_samples=raw.size // 2, on_device=True)
    assert host.n_frames == 150 and host.n_slots == 150
    _same(host, dev)
    assert np.array_equal(host.frames[0], res.logical[3][:host.frames[0].shape[0]])
    assert [i for i in host.tii[0].ids][0][:2] == (12, 5)


def test_lost_frame_becomes_erasures(receiver, oracle):
    raw, res = cases.config1(14, "u8", 30.0, 250.0, 17)
    raw = raw.copy()
    f0 = res.frame0 + 6 * 196608
    rng = np.random.default_rng(1)
    raw[2 * f0:2 * (f0 + 196608)] = rng.integers(120, 136, 2 * 196608, dtype=np.uint8)     # frame 6 replaced by noise
    dec = receiver.decode(raw, "u8")
    assert dec.n_slots == 14 and dec.n_frames == 13 and dec.t_f[6] == -1
    assert dec.fib_ok.reshape(14, 12)[6].sum() == 0 and dec.fib_ok.reshape(14, 12)[[0, 5, 7, 13]].all()
    sc = res.ens.subchannels[0]
    o = oracle.decode(raw, "u8", [oracle.subch_struct(sc)])
    assert o["nf"] == 14 and np.array_equal(o["t_f"], dec.t_f)
    assert np.array_equal(dec.frames[0], o["frames"][0])           # same erasures, same decisions
    truth = res.logical[sc.subchid][:dec.frames[0].shape[0]]
    bad = np.nonzero((dec.frames[0] != truth).any(axis=1))[0]
    assert set(bad) <= set(range(24 - 15, 28))                      # only rows whose interleaver span touches the lost CIFs


def test_empty_noise_and_truncated_inputs(receiver):
    silence = np.full(2 * 400000, 128, np.uint8)
    d = receiver.decode(silence, "u8", tii=True)
    assert d.n_frames == 0 and d.n_slots == 0 and d.aus == [] and d.frames == []
    noise = np.random.default_rng(3).integers(96, 160, 2 * 600000, dtype=np.uint8)
    assert receiver.decode(noise, "u8").n_frames == 0
    tiny = np.full(2 * 1000, 128, np.uint8)
    assert receiver.decode(tiny, "u8").n_frames == 0
    ens = synth.config1_ensemble()
    raw, res = synth.synthesize(ens, 5, "u8", synth.Impairments(snr_db=30, lead_in=777), tail=0)
    cut = raw[:2 * (777 + 4 * 196608 + 100000)]                     # fifth frame incomplete
    d = receiver.decode(cut, "u8")
    assert d.n_frames == 4 and d.fib_ok.all() and d.frames[0].shape[0] == 1


def test_cfo_drift_and_snr_staircase(receiver):
    """configs[4] in small: s16 and f32 recordings with linear CFO drift; SNR steps 3 -> 20 dB over the recording."""
    ens = synth.config1_ensemble(seed=9)
    n_frames = 40
    clean, truth = synth.modulate(ens, n_frames, seed=9)
    x = synth.impair(clean, synth.Impairments(snr_db=None, cfo_hz=-1500.0, cfo_drift_hz_per_s=400.0, seed=2, lead_in=30000), tail=70000)
    rng = np.random.default_rng(5)
    snr = np.repeat(np.array([3, 5, 7, 9, 12, 15, 18, 20], float), 5)           # per frame
    for f in range(n_frames):
        a, b = 30000 + f * 196608, 30000 + (f + 1) * 196608
        sg = np.sqrt(0.5 * 10 ** (-snr[f] / 10))
        x[a:b] += (sg * (rng.standard_normal(b - a) + 1j * rng.standard_normal(b - a))).astype(np.complex64)
    for fmt, rms in (("s16", 6000.0), ("f32", 1.0)):
        raw = synth.quantize(x, fmt, rms)
        d = receiver.decode(raw, fmt, symstat=True)
        assert d.n_frames == n_frames
        t = (d.t_f - d.t_f[0]) / 2.048e6
        assert np.all(np.abs(d.cfo_hz - (d.cfo_hz[0] + 400.0 * t)) < 12.0)       # tracks the drift frame by frame
        ok = d.fib_ok.reshape(n_frames, 12)
        assert ok[10:].all()                                                     # >= 7 dB: every FIB passes
        fr = d.frames[0]
        want = truth["logical"][3][:fr.shape[0]]
        good_rows = np.arange(4 * 10, fr.shape[0])                               # rows fed only by frames at >= 7 dB
        assert np.array_equal(fr[good_rows], want[good_rows])
        aus = [a for a in d.aus if a.cif >= 40]
        assert aus and all(a.crc_ok for a in aus)
        assert d.firecode_fail <= 3                                              # only the superframes at 3 dB are lost


def test_receiver_pool_band_scan():
    """several recordings (different ensembles / CFO, as different Band III channels) through a pool of two contexts:
    results in input order and identical to one-by-one decoding"""
    import torch
    recs = []
    for i in range(5):
        ens = synth.config1_ensemble(seed=30 + i)
        ens.eid = 0x4000 + i
        ens.tii = [synth.TiiTx(5 + i, 2 + i, 1.0)]
        raw, res = synth.synthesize(ens, 12 + i, "u8", synth.Impairments(snr_db=22.0, cfo_hz=300.0 * (i - 2), seed=i), seed=i, tail=70000)
        recs.append((raw, res))
    pool = pipeline.ReceiverPool(0, depth=2)
    pinned = [torch.from_numpy(r).pin_memory() for r, _ in recs]
    items = [(p.data_ptr(), r.size // 2, "u8", False, None, pipeline.OPT_TII) for p, (r, _) in zip(pinned, recs)]

    def grab(i, rx, res):
        fibs = np.ctypeslib.as_array(res.fibs, (res.n_slots * 12, 32)).copy()
        fr = np.ctypeslib.as_array(res.frames[0], (res.frame_rows[0], 288)).copy()
        return res.n_frames, fibs, fr, (int(res.tii[0].main_id[0]), int(res.tii[0].sub_id[0])) if res.n_tii else None

    out = pool.map_raw(items, grab)
    pool.close()
    for i, ((raw, res), (nf, fibs, fr, tii)) in enumerate(zip(recs, out)):
        assert nf == 12 + i
        assert np.array_equal(fibs.reshape(-1, 96), res.fibs[:4 * nf])
        assert int(fibs[0, 2]) << 8 | int(fibs[0, 3]) == 0x4000 + i              # EId of FIG 0/0: the right recording in the right slot
        assert np.array_equal(fr, res.logical[3][:fr.shape[0]])
    assert [o[3] for o in out] == [(5 + i, 2 + i) for i in range(5)]             # short recordings: one TII window over all frames


def test_sample_clock_offset_and_multipath_gpu(receiver, oracle):
    ens = synth.config1_ensemble(seed=41)
    imp = synth.Impairments(snr_db=18.0, cfo_hz=321.0, taps=[(25, 0.5), (90, 0.3j)], sample_offset_ppm=-60.0, seed=6, lead_in=20000)
    raw, res = synth.synthesize(ens, 14, "u8", imp, seed=6, tail=150000)
    d = receiver.decode(raw, "u8")
    o = oracle.decode(raw, "u8", [oracle.subch_struct(ens.subchannels[0])])
    assert d.n_frames == 14 and d.n_slots == 14 and np.array_equal(d.t_f, o["t_f"])
    assert np.array_equal(d.fibs.reshape(-1, 96), res.fibs[:56]) and d.fib_ok.all()
    assert np.array_equal(d.frames[0], o["frames"][0])
    assert np.array_equal(d.frames[0], res.logical[3][:d.frames[0].shape[0]])


def test_full_size_workload_properties():
    """BASELINE.json configs[1] at its full size (60 s, 12 sub-channels, 246 MB of u8 IQ), checked through size-independent
    properties: every FIB passes its CRC and equals the transmitted one, every sub-channel's logical frames equal the modulator's
    payload (an encode -> channel -> decode round trip over 30 000 codewords), every DAB+ access unit passes its CRC and appears in
    transmission order, RS found nothing to correct at 25 dB, the TII ids are the transmitted ones; decoding the same recording
    from a host buffer (chunked upload) and from a device buffer gives identical bytes."""
    import sys
    from pathlib import Path
    import torch
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    import bench
    raw, res, n_frames = bench.make_recording(60.0, 0)
    rx = pipeline.Receiver(0)
    d = rx.decode(raw, "u8", tii=True)
    assert d.n_frames == n_frames == d.n_slots and d.fib_ok.all()
    assert np.array_equal(d.fibs.reshape(-1, 96), res.fibs[:4 * n_frames])
    assert len(d.subch) == 12 and sum(s.size_cu for s in d.subch) == 854
    for i, sc in enumerate(d.subch):
        want = res.logical[sc.subch_id]
        assert d.frames[i].shape[0] == 4 * n_frames - 15 and np.array_equal(d.frames[i], want[:d.frames[i].shape[0]]), sc.subch_id
    assert d.rs_uncorrectable == 0 and d.firecode_fail == 0 and all(a.crc_ok for a in d.aus)
    for i, sc in enumerate(d.subch):
        if sc.kind != 1:
            continue
        got = [a.data for a in d.aus if a.subch_index == i]
        truth = res.aus[sc.subch_id]
        i0 = truth.index(got[0])
        assert len(got) >= 3 * ((4 * n_frames - 15) // 5 - 1) and got == truth[i0:i0 + len(got)]
    assert all([(m, s) for m, s, _ in w.ids] == [(20, 7), (20, 11)] for w in d.tii) and len(d.tii) == n_frames // 40
    dev = torch.from_numpy(raw).cuda()
    d2 = rx.decode(dev.data_ptr(), "u8", tii=True, n_samples=raw.size // 2, on_device=True)
    assert np.array_equal(d2.fibs, d.fibs) and np.array_equal(d2.t_f, d.t_f) and np.array_equal(d2.cfo_hz, d.cfo_hz)
    assert all(np.array_equal(a, b) for a, b in zip(d.frames, d2.frames))
    rx.close()


def _same_decode(a, b):
    assert a.n_frames == b.n_frames and a.n_slots == b.n_slots
    assert np.array_equal(a.t_f, b.t_f) and np.array_equal(a.cfo_hz, b.cfo_hz) and np.array_equal(a.cir_ratio, b.cir_ratio)
    assert np.array_equal(a.fibs, b.fibs) and np.array_equal(a.fib_ok, b.fib_ok)
    assert len(a.subch) == len(b.subch) and all(bytes(x) == bytes(y) for x, y in zip(a.subch, b.subch))
    for i, (x, y) in enumerate(zip(a.frames, b.frames)):
        assert x.shape == y.shape and np.array_equal(x, y), f"logical frames of sub-channel {i}"
    assert np.array_equal(a.sf, b.sf)
    assert (a.superframes, a.rs_corrected, a.rs_bit_errors, a.rs_uncorrectable, a.firecode_fail) == \
           (b.superframes, b.rs_corrected, b.rs_bit_errors, b.rs_uncorrectable, b.firecode_fail)
    assert len(a.aus) == len(b.aus)
    for x, y in zip(a.aus, b.aus):
        assert (x.subch_index, x.superframe, x.cif, x.crc_ok, x.header, x.data) == (y.subch_index, y.superframe, y.cif, y.crc_ok, y.header, y.data)
    assert len(a.tii) == len(b.tii)
    for x, y in zip(a.tii, b.tii):
        assert x.ids == y.ids and x.n_tii_frames == y.n_tii_frames and np.array_equal(x.spectrum, y.spectrum)


def test_chunked_decode_equals_single_pass(receiver):
    """BASELINE configs[4] at 1/30 of its length: a 120 s full-ensemble recording (s16 and f32) with linear CFO drift and an SNR
    staircase 0 -> 20 dB, decoded in 2, 4 and 7 overlapping time chunks (shard.decode_chunked) -- frames, FIBs, logical frames of
    all 12 sub-channels, superframes, RS statistics, access units and TII windows byte-identical to the single-pass decode."""
    from abracadabra_b200 import shard
    ens = synth.full_ensemble(seed=2300)
    n_frames = 1250
    clean, truth = synth.modulate(ens, n_frames, seed=23)
    x = synth.impair(clean, synth.Impairments(snr_db=None, cfo_hz=-800.0, cfo_drift_hz_per_s=200.0 / 60.0, seed=4, lead_in=41000), tail=90000)
    del clean
    snr = np.floor(np.arange(n_frames) * 21.0 / n_frames)            # 0, 1, ... 20 dB, about 60 frames each
    synth.add_noise_staircase(x, 41000, snr, seed=8)
    for fmt, rms in (("s16", 6000.0), ("f32", 1.0)):
        raw = synth.quantize(x, fmt, rms)
        one = receiver.decode(raw, fmt, tii=True)
        assert one.n_frames >= n_frames - 130                         # the first seconds at 0-1 dB may not synchronise
        assert len(one.subch) == 12 and len(one.aus) > 5000
        # the regimes of SURVEY App. B are all there: nothing at the bottom, clean at the top
        late = [a for a in one.aus if a.cif > 4 * 900]
        assert late and all(a.crc_ok for a in late)
        for n_chunks in ((2, 4, 7) if fmt == "s16" else (4,)):
            st = {}
            got = shard.decode_chunked(raw, fmt, n_chunks, rx=receiver, tii=True, stats=st)
            _same_decode(one, got)
        del raw


def test_error_exit_leaves_the_context_usable(receiver):
    """A caller-given sub-channel that does not fit the CIF makes b200dab_decode fail AFTER the front end and the FIC chain have been
    queued on three streams.  The error exit has to drain them (the caller may free its buffer at once), and the next call on the
    same context has to work and give the same result as before."""
    raw, res = cases.config1(12, "u8", 30.0, 250.0, 21)
    good = receiver.decode(raw, "u8", tii=True)
    assert good.n_frames == 12 and len(good.subch) == 1
    bad = pipeline.SubCh()
    for name, _ in pipeline.SubCh._fields_:
        setattr(bad, name, getattr(good.subch[0], name))
    bad.start_cu = 860                                         # 860 + size > 864
    with pytest.raises(RuntimeError):
        receiver.decode(raw.copy(), "u8", subch=[bad])
    again = receiver.decode(raw, "u8", tii=True)
    _same(good, again)


def test_frames_written_by_the_trace_kernel_and_rs_patch_list(monkeypatch):
    """By default k_vit_trace writes the logical frames to the host result buffer itself, group by group; those of the DAB+ sub-channels
    arrive as traced, and the host applies the RS corrections from a list (offset, byte) that comes back with the other DAB+ results.
    On a recording noisy enough for the RS decoder to work (full ensemble at 5 dB: hundreds of corrected bytes, some uncorrectable
    words) frames, access units and counters must equal those of the plain order (device-to-host copies behind the trace and behind
    the RS correction), also when the list is longer than the prefix that is copied blindly (corrected frames copied again)."""
    ens = synth.full_ensemble(seed=77)
    raw, res = synth.synthesize(ens, 24, "u8", synth.Impairments(snr_db=5.0, cfo_hz=140.0, seed=9, lead_in=30000), seed=9, tail=60000)
    out = []
    for env in ({"B200DAB_FRAME_EXPORT": "0"}, {}, {"B200DAB_RS_PATCH_PREFIX": "3"}, {"B200DAB_TRACE_PDL": "0"}):
        for k in ("B200DAB_FRAME_EXPORT", "B200DAB_RS_PATCH_PREFIX", "B200DAB_TRACE_PDL"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        rx = pipeline.Receiver(0)
        out.append(rx.decode(raw, "u8", tii=True))
        out.append(rx.decode(raw, "u8", tii=True))             # a second call on the same context: counters and lists start over
        rx.close()
    base = out[0]
    assert base.n_frames == 24 and base.rs_corrected > 50, (base.n_frames, base.rs_corrected)
    for d in out[1:]:
        _same(base, d)
        assert (d.rs_corrected, d.rs_bit_errors, d.rs_uncorrectable, d.firecode_fail) == (base.rs_corrected, base.rs_bit_errors, base.rs_uncorrectable, base.firecode_fail)
