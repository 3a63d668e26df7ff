"""GPU parity tests: every stage of the CUDA path, called through the C ABI (libb200dab.so), against the CPU checker
(oracle/daboracle.c) on the same seeded inputs.  Bit-exact for integer/byte/index results; floating-point results
within the tolerance written next to each comparison."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


def test_coarse_sync_bit_exact(receiver, oracle):
    for fmt in ("u8", "s16", "f32"):
        raw, res = cases.config1(6, fmt)
        P, S, cand = receiver.stage_coarse(raw, fmt)
        Po = oracle.block_power(raw, fmt)
        assert np.array_equal(P, Po), fmt                 # same summation order, explicit fma: bit-exact
        q = oracle.null_metric(Po)
        co = oracle.candidates(q, oracle.nsamples(raw, fmt))
        assert np.array_equal(cand, co), fmt
        assert res.frame0 in cand


def test_fine_sync_matches_oracle(receiver, oracle):
    raw, res = cases.config1(6, "u8")
    _, _, cand = receiver.stage_coarse(raw, "u8")
    t_f, cfo, met = receiver.stage_fine(raw, "u8", cand)
    for i, c in enumerate(cand):
        to, co, mo = oracle.frame_sync(raw, "u8", c)
        if mo[2] < 40:      # not a frame (noise candidate): nothing to compare
            assert met[i, 2] < 40
            continue
        assert t_f[i] == to                                # integer timing decision
        assert abs(cfo[i] - co) <= 0.05                    # Hz; float accumulation order differs
        assert abs(met[i, 2] - mo[2]) <= 1e-3 * mo[2]
        assert abs(cfo[i] - res.cfo_hz) < 2.0


def test_demod_soft_bits(receiver, oracle):
    raw, res = cases.config1(6, "u8")
    _, _, cand = receiver.stage_coarse(raw, "u8")
    t_f, cfo, met = receiver.stage_fine(raw, "u8", cand)
    keep = met[:, 2] >= 40
    t_f, cfo = t_f[keep][:3], cfo[keep][:3]
    soft = receiver.stage_demod(raw, "u8", t_f, cfo)
    for i in range(len(t_f)):
        so = oracle.demod_frame(raw, "u8", t_f[i], cfo[i])
        d = np.abs(soft[i].astype(np.int32) - so.astype(np.int32))
        # tolerance: +-1 LSB of the int8 soft bit (FFT rounding and gain reduction order), on at most 0.5 % of the bits
        assert d.max() <= 1
        assert (d > 0).mean() < 5e-3


def test_viterbi_bit_exact(receiver, oracle):
    rng = np.random.default_rng(5)
    for steps, n_cw in ((774, 40), (2310, 33), (1158, 7)):
        # noisy soft bits around a valid codeword plus pure-noise codewords: decisions (incl. ties) must be identical
        from abracadabra_b200 import synth
        info = rng.integers(0, 2, (n_cw, steps - 6), dtype=np.uint8)
        code = synth.conv_encode(info).astype(np.int32)
        soft = (1 - 2 * code) * 40 + rng.normal(0, 45, code.shape)
        soft[n_cw // 2:] = rng.normal(0, 60, soft[n_cw // 2:].shape)
        punct = rng.random(soft.shape) < 0.3
        soft[punct] = 0
        soft = np.clip(np.rint(soft), -127, 127).astype(np.int8)
        soft[0, :40] = 0                                   # a run of ties right at the start
        want = [oracle.viterbi(soft[c]) for c in range(n_cw)]
        # both lane splits of k_viterbi (two lanes per codeword pair / four lanes for long codewords) give the checker's bits
        for split in (1 << 30, 0):
            receiver.set_viterbi_split(split)
            got = receiver.stage_viterbi(soft)
            for c in range(n_cw):
                assert np.array_equal(got[c], want[c]), (steps, c, split)
        receiver.set_viterbi_split(0)


def test_fic_and_msc_bit_exact(receiver, oracle):
    raw, res = cases.multi(8, "u8")
    from abracadabra_b200 import pipeline
    _, _, cand = receiver.stage_coarse(raw, "u8")
    t_f, cfo, met = receiver.stage_fine(raw, "u8", cand)
    keep = met[:, 2] >= 40
    soft = receiver.stage_demod(raw, "u8", t_f[keep], cfo[keep])
    nf = soft.shape[0]
    assert nf == 8
    fibs, ok = receiver.stage_fic(soft)
    for f in range(nf):
        fo, oo = oracle.fic_decode(soft[f, :9216])
        assert np.array_equal(fibs[12 * f:12 * f + 12], fo)
        assert np.array_equal(ok[12 * f:12 * f + 12], oo)
    assert ok.all()
    assert np.array_equal(fibs.reshape(-1, 96), res.fibs[:nf * 4])
    cifs = soft[:, 9216:].reshape(nf * 4, 55296)
    for sc in res.ens.subchannels:
        got = receiver.stage_msc(soft, pipeline.subch_from_synth(sc))
        exp = oracle.msc_decode(cifs, oracle.subch_struct(sc))
        assert np.array_equal(got, exp), sc.subchid
        assert np.array_equal(got, res.logical[sc.subchid][:got.shape[0]]), sc.subchid


def test_dabplus_rs_and_aus(receiver, oracle):
    raw, res = cases.config1(10, "u8")
    sc = res.ens.subchannels[0]
    frames = res.logical[sc.subchid][3:38].copy()        # starts mid-superframe: alignment must be found
    rng = np.random.default_rng(3)
    # inject byte errors: up to 5 per RS codeword in some superframes, 6+ (uncorrectable) in one codeword
    flat = frames.reshape(-1)
    s = sc.bitrate // 8
    base = 2 * frames.shape[1]                            # superframe starts at row 2 (global row 5)
    for sfi in range(3):
        off = base + sfi * 120 * s
        for j in range(s):
            for k in rng.choice(120, size=rng.integers(0, 6), replace=False):
                flat[off + j + s * k] ^= rng.integers(1, 256)
    off = base + 4 * 120 * s
    for k in rng.choice(120, size=8, replace=False):
        flat[off + 3 + s * k] ^= rng.integers(1, 256)
    aus, stats, corrected = receiver.stage_dabplus(frames, sc.bitrate)
    # oracle on the same aligned superframes
    exp_aus, exp_stats = [], dict(rs_corrected=0, rs_bit_errors=0, rs_uncorrectable=0)
    n_sf = (frames.shape[0] - 2) // 5
    for m in range(n_sf):
        r = oracle.superframe(frames[2 + 5 * m:7 + 5 * m].reshape(-1), s)
        exp_stats["rs_corrected"] += r["corrected"]; exp_stats["rs_bit_errors"] += r["bit_errors"]
        exp_stats["rs_uncorrectable"] += r["uncorrectable"]
        exp_aus += list(zip(r["aus"], r["crc_ok"]))
        assert np.array_equal(corrected[2 + 5 * m:7 + 5 * m].reshape(-1), r["data"])
    assert stats["superframes"] == n_sf
    for k in exp_stats:
        assert stats[k] == exp_stats[k], k
    assert stats["rs_uncorrectable"] == 1
    assert [(a.data, int(a.crc_ok)) for a in aus] == [(d, int(o)) for d, o in exp_aus]
    truth = res.aus[sc.subchid]
    good = [a.data for a in aus if a.crc_ok]
    assert len(good) >= 3 * (n_sf - 1) and all(g in truth for g in good)


def test_whole_path_against_truth_and_oracle(receiver, oracle):
    for fmt, snr in (("u8", 30.0), ("s16", 12.0), ("f32", 8.0)):
        raw, res = cases.config1(12, fmt, snr, -4321.0, 7)
        dec = receiver.decode(raw, fmt, tii=True, symstat=True)
        assert dec.n_frames == 12 and dec.n_slots == 12
        assert dec.fib_ok.all()
        assert np.array_equal(dec.fibs.reshape(-1, 96), res.fibs[:48])
        assert np.all(np.abs(dec.cfo_hz - res.cfo_hz) < 3.0)
        assert len(dec.subch) == 1 and dec.subch[0].start_cu == 0 and dec.subch[0].size_cu == 72 and dec.subch[0].kind == 1
        sc = res.ens.subchannels[0]
        o = oracle.decode(raw, fmt, [oracle.subch_struct(sc)])
        assert o["nf"] == 12 and np.array_equal(o["t_f"], dec.t_f)
        if snr >= 12:
            assert np.array_equal(dec.frames[0], res.logical[sc.subchid][:33])
        truth = res.aus[sc.subchid]
        got = [a.data for a in dec.aus if a.crc_ok]
        assert len(got) >= 15
        i0 = truth.index(got[0])
        assert got == truth[i0:i0 + len(got)]
        assert dec.gpu_launches >= 10


def test_tii_ids(receiver):
    from abracadabra_b200 import synth
    ens = synth.config1_ensemble()
    ens.tii = [synth.TiiTx(12, 5, 1.0), synth.TiiTx(33, 17, 0.5, delay=30), synth.TiiTx(60, 2, 0.25, delay=100)]
    raw, res = synth.synthesize(ens, 41, "u8", synth.Impairments(snr_db=25, cfo_hz=50.0), tail=150000)
    dec = receiver.decode(raw, "u8", tii=True, no_msc=True)
    assert len(dec.tii) == 1
    ids = dec.tii[0].ids
    assert [(m, s) for m, s, _ in ids] == [(60, 2), (12, 5), (33, 17)]      # ascending sub id, like the reference
    lv = {(m, s): l for m, s, l in ids}
    # levels are linear in amplitude: ratios within 0.5 dB of the transmitted 6 / 12 dB steps
    assert abs(20 * np.log10(lv[(33, 17)] / lv[(12, 5)]) + 6.02) < 0.5
    # weak transmitters read low (magnitude-domain noise subtraction), as in the reference: 0.25 -> 0.204
    assert abs(20 * np.log10(lv[(60, 2)] / lv[(12, 5)]) + 12.04) < 2.0
    assert 0.85 < lv[(12, 5)] < 1.1
    sp = dec.tii[0].spectrum
    assert int(np.argmax(sp)) // 2 in [b * 24 + 5 for b in (2, 3, 4, 7)]


def test_tii_fold_and_detect_match_oracle(receiver, oracle):
    """BASELINE.json configs[2] at kernel level: a batch of null symbols with 1-8 overlapping transmitters, CFO already
    removed; fold within 1e-4 relative of the checker, ids exact, levels within 1e-4."""
    from abracadabra_b200 import synth
    rng = np.random.default_rng(12)
    n_sets, per = 12, 20
    for k in range(n_sets):
        n_tx = 1 + (k % 8)
        subs = rng.choice(np.arange(1, 24), n_tx, replace=False)
        ens = synth.EnsembleConfig(tii=[synth.TiiTx(int(rng.integers(0, 70)), int(s), float(rng.uniform(0.2, 1.0)), delay=int(rng.integers(0, 300)),
                                                    phase=float(rng.uniform(0, 6.28))) for s in subs])
        sym = synth.tii_symbol(ens)[-2048:]
        nulls = np.zeros((2 * per, 2048), np.complex64)
        sigma = 10 ** (-rng.uniform(10, 30) / 20) * np.sqrt(0.5)
        noise = (rng.standard_normal((2 * per, 2048)) + 1j * rng.standard_normal((2 * per, 2048))) * sigma
        nulls[0::2] = sym
        nulls += noise.astype(np.complex64)
        fold = receiver.stage_tii_fold(nulls)
        fo = np.array([oracle.tii_fold(oracle.fft2048(x)) for x in nulls])
        assert np.allclose(fold, fo, rtol=1e-4, atol=1e-3 * fo.max())
        norm = np.full(2 * per, 1.0, np.float32)
        sign = np.tile(np.array([1, -1], np.int32), per)
        got = receiver.stage_tii_detect(fo, norm, sign)           # same fold in: the integer decisions must be identical
        ids_o, spec_o = oracle.tii_detect(fo, norm, sign)
        assert [(m, s) for m, s, _ in got.ids] == [(m, s) for m, s, _ in ids_o]
        assert np.allclose([l for _, _, l in got.ids], [l for _, _, l in ids_o], rtol=1e-4)
        assert np.allclose(got.spectrum, spec_o, rtol=1e-5, atol=1e-3)
        want = sorted((tx.sub, tx.main) for tx in ens.tii if tx.amp >= 0.3)
        have = [(s, m) for m, s, _ in got.ids]
        assert all(w in have for w in want), (want, have)


def test_tii_random_batch_against_reference_core():
    """BASELINE.json configs[2] at system level: recordings with 1-8 overlapping transmitters, delays, a multipath tap, CFO up to
    +-5 kHz and 15-30 dB SNR.  The reference core (oracle/_ref, run here on the same files) and the CUDA path must report the same
    (main, sub) ids in the same order; levels within 0.2 dB for the stronger half, 0.5 dB for weak transmitters at low SNR (a CPU
    sweep of 26 such cases, oracle vs reference: ids always identical, median level deviation 0.07 dB, worst 0.46 dB; 0.1 dB
    without multipath is pinned by the golden fixtures)."""
    import tempfile
    from pathlib import Path
    import orc
    from abracadabra_b200 import pipeline, synth
    if not orc.have_ref():
        pytest.skip("reference core not built (oracle/_ref)")
    rng = np.random.default_rng(77)
    rx = pipeline.Receiver(0)
    for case in range(4):
        n_tx = [1, 3, 5, 8][case]
        subs = rng.choice(np.arange(1, 24), n_tx, replace=False)
        txs = [synth.TiiTx(int(rng.integers(0, 70)), int(s), float(rng.uniform(0.3, 1.0)), delay=int(rng.integers(0, 200)),
                           phase=float(rng.uniform(0, 6.28))) for s in subs]
        ens = synth.config1_ensemble(seed=200 + case)
        ens.tii = txs
        imp = synth.Impairments(snr_db=float(rng.uniform(15, 30)), cfo_hz=float(rng.uniform(-5000, 5000)),
                                taps=[(int(rng.integers(5, 60)), complex(rng.uniform(0.1, 0.4), rng.uniform(-0.3, 0.3)))], seed=case)
        raw, _ = synth.synthesize(ens, 46, "u8", imp, seed=case, tail=150000)
        with tempfile.TemporaryDirectory() as td:
            p = Path(td) / "t.u8"
            raw.tofile(p)
            _, ev = orc.run_ref(p, "u8", Path(td) / "log.jsonl", tii=1)
        ref = [e for e in ev if e["ev"] == "tii" and e["n"] > 0][0]
        dec = rx.decode(raw, "u8", tii=True, no_msc=True)
        win = [w for w in dec.tii if w.ids][0]
        assert [(m, s) for m, s, _ in win.ids] == [(m, s) for m, s, _ in ref["ids"]], (case, win.ids, ref["ids"])
        assert sorted((m, s) for m, s, _ in win.ids) == sorted((t.main, t.sub) for t in txs)
        top = max(lr for _, _, lr in ref["ids"])
        for (m, s, l), (_, _, lr) in zip(win.ids, ref["ids"]):
            assert abs(20 * np.log10(l / lr)) <= (0.2 if lr >= 0.5 * top else 0.5), (case, m, s, l, lr)
        rs = np.array(ref["spectrum"])
        strong = np.nonzero(rs > 0.3 * rs.max())[0]
        assert np.all(np.abs(win.spectrum[strong] / win.spectrum.max() - rs[strong] / rs.max()) < 0.02)
    rx.close()


def test_random_ensembles_against_reference_core():
    """Random single-service ensembles (bitrate, EEP-A/B level, sub-channel position, u8 / s16, CFO up to +-10 kHz, 9-25 dB,
    lead-in): every access unit the reference core (oracle/_ref, run here on the same file) delivers is delivered by the CUDA
    path, byte for byte and in the same order; logical frames equal the transmitted payload; CFO within 1.5 Hz of the truth and of
    the reference's report."""
    import hashlib
    import tempfile
    from pathlib import Path
    import orc
    from abracadabra_b200 import pipeline, synth
    if not orc.have_ref():
        pytest.skip("reference core not built (oracle/_ref)")
    rng = np.random.default_rng(2024)
    rx = pipeline.Receiver(0)
    for case in range(5):
        opt = str(rng.choice(["A", "B"]))
        lvl = int(rng.integers(1, 5))
        br = int(rng.choice([32, 48, 64, 96, 128] if opt == "A" else [32, 64, 96, 128]))
        fmt = str(rng.choice(["u8", "s16"]))
        sc = synth.SubChannel(int(rng.integers(1, 60)), int(rng.integers(0, 200)), "dabplus", br, ("EEP", lvl, opt), seed=300 + case)
        ens = synth.EnsembleConfig(eid=0x5000 + case, label=f"RND {case}", subchannels=[sc],
                                   services=[synth.Service(0x5100 + case, f"RND SVC {case}", sc.subchid)])
        imp = synth.Impairments(snr_db=float(rng.uniform(9, 25)), cfo_hz=float(rng.uniform(-10000, 10000)), seed=case,
                                lead_in=int(rng.integers(1000, 90000)))
        raw, res = synth.synthesize(ens, 26, fmt, imp, seed=case, tail=150000)
        with tempfile.TemporaryDirectory() as td:
            p = Path(td) / ("r." + fmt)
            raw.tofile(p)
            _, ev = orc.run_ref(p, fmt, Path(td) / "log.jsonl", service=0)
        ref = [bytes.fromhex(e["data"]) for e in ev if e["ev"] == "audio" and e["len"] > 0]
        assert len(ref) >= 24, (case, len(ref))
        dec = rx.decode(raw, fmt, symstat=True)
        assert dec.n_frames == 26 and dec.fib_ok.all()
        assert len(dec.subch) == 1 and (dec.subch[0].subch_id, dec.subch[0].start_cu, dec.subch[0].bitrate) == (sc.subchid, sc.start_cu, br)
        assert np.array_equal(dec.frames[0], res.logical[sc.subchid][:dec.frames[0].shape[0]])
        got = [a.data for a in dec.aus if a.crc_ok]
        i0 = got.index(ref[0])
        assert got[i0:i0 + len(ref)] == ref, case                    # the reference's AUs: a contiguous run of ours
        assert len(got) >= len(ref)
        assert np.all(np.abs(dec.cfo_hz - imp.cfo_hz) < 1.5)
        per = [e for e in ev if e["ev"] == "periodic" and not e.get("null") and e["level"] == 3]
        # the reference's own estimate wanders by a few Hz below 12 dB
        assert per and abs(per[-1]["freqOffset"] / 10.0 - float(np.median(dec.cfo_hz))) < (1.5 if imp.snr_db >= 12 else 4.0)
        ens_ev = [e for e in ev if e["ev"] == "ensemble" and e["ueid"]][0]
        info = pipeline.parse_ensemble(dec.fibs, dec.fib_ok)
        assert info["ueid"] == ens_ev["ueid"] and info["label"] == ens_ev["label"] and info["lto"] == ens_ev["lto"]
    rx.close()
