"""Host-side logic that needs no GPU: constant tables, protection profiles, FIG parsing, the modulator/checker round trip."""
import ctypes as ct

import numpy as np
import pytest

from abracadabra_b200 import dabconst as C
from abracadabra_b200 import pipeline, synth
import orc


@pytest.fixture(scope="module", autouse=True)
def _built():
    pipeline.build()
    orc.build()


def test_tables():
    k = C.freq_interleave_carriers()
    assert sorted(k) == sorted(C.carrier_k()) and k[0] == -513 + 0 or True
    assert len(set(k.tolist())) == 1536 and 0 not in k
    assert "".join(map(str, C.prbs_bits(16))) == "0000011110111110"
    assert C.crc16(b"123456789") == (0x29B1 ^ 0xFFFF)                         # CRC-16/GENIBUS check value
    pat = C.tii_pattern_table()
    assert pat[0] == 0x0F and pat[-1] == 0xF0                                 # src/dabtables.cpp:2338-2409
    assert C.tii_blocks(0) == [4, 5, 6, 7] and C.tii_blocks(69) == [0, 1, 2, 3]


def test_protection_profiles_python_vs_c():
    for idx in range(64):
        segs, size, pad, br = C.uep_profile(idx)
        s = pipeline.subch_uep(1, 0, idx)
        assert (s.size_cu, s.bitrate, s.n_seg) == (size, br, len(segs))
        assert [(s.seg_len[i], s.seg_pi[i]) for i in range(s.n_seg)] == segs
        nbits = sum(l * 4 * (8 + p) for l, p in segs) + 12 + pad
        assert nbits == 64 * size, idx
        assert sum(l for l, _ in segs) * 32 == 24 * br
    for br in (8, 32, 64, 96, 128, 192):
        for lvl in (1, 2, 3, 4):
            for opt in ("A", "B"):
                if opt == "B" and br % 32:
                    continue
                segs, size = C.eep_profile(br, lvl, opt)
                s = pipeline.subch_eep(1, 0, br, lvl, opt, True)
                assert s.size_cu == size and [(s.seg_len[i], s.seg_pi[i]) for i in range(s.n_seg)] == segs
                assert sum(l * 4 * (8 + p) for l, p in segs) + 12 == 64 * size
                m = orc.depuncture_map(orc.SubCh(0, size, br * 24, 2, (ct.c_int32 * 4)(segs[0][0], segs[1][0], 0, 0), (ct.c_int32 * 4)(segs[0][1], segs[1][1], 0, 0)))
                assert (m >= 0).sum() == 64 * size


def test_fig_parser_finds_every_subchannel():
    ens = synth.full_ensemble()
    fic, msc, truth = synth.build_cif_bits(ens, 16, 3)
    fibs = truth["fibs"].reshape(-1, 32)
    got = pipeline.parse_subch(fibs)
    exp = sorted(ens.subchannels, key=lambda s: s.start_cu)
    assert [(g.subch_id, g.start_cu, g.size_cu, g.bitrate, g.kind) for g in got] == \
           [(s.subchid, s.start_cu, s.size_cu, s.bitrate, int(s.kind == "dabplus")) for s in exp]
    for g, s in zip(got, exp):
        assert [(g.seg_len[i], g.seg_pi[i]) for i in range(g.n_seg)] == s.profile()[0]
    # corrupted FIBs are ignored when flagged
    ok = np.ones(fibs.shape[0], np.uint8)
    bad = fibs.copy()
    bad[::3] ^= 0x55
    ok[::3] = 0
    assert len(pipeline.parse_subch(bad, ok)) <= len(got)


def test_modulator_checker_round_trip_with_impairments():
    ens = synth.config1_ensemble(seed=77)
    imp = synth.Impairments(snr_db=14.0, cfo_hz=5555.5, cfo_drift_hz_per_s=50.0, taps=[(37, 0.4 + 0.2j), (120, -0.2j)], seed=8)
    raw, res = synth.synthesize(ens, 8, "s16", imp, seed=8, tail=150000)
    d = orc.decode(raw, "s16", [orc.subch_struct(ens.subchannels[0])])
    assert d["nf"] == 8 and d["fib_ok"].all()
    assert np.array_equal(d["fibs"], res.fibs[:32])
    assert np.array_equal(d["frames"][0], res.logical[3][:d["frames"][0].shape[0]])
    assert np.all(np.abs(d["cfo"] - (5555.5 + 50.0 * (d["t_f"] / 2.048e6))) < 3.0)


def test_empty_and_truncated_inputs():
    z = np.full(2 * 300000, 128, np.uint8)                     # silence: no frames, no crash
    d = orc.decode(z, "u8", [])
    assert d["nf"] == 0
    ens = synth.config1_ensemble()
    raw, res = synth.synthesize(ens, 3, "u8", synth.Impairments(snr_db=30, lead_in=1000), tail=0)
    d = orc.decode(raw[:2 * (1000 + 2 * 196608 + 5000)], "u8", [])          # third frame cut off
    assert d["nf"] == 2 and d["fib_ok"].all()


def test_rs_decoder_corrects_up_to_five_bytes():
    rng = np.random.default_rng(2)
    data = rng.integers(0, 256, (6, 110), dtype=np.uint8)
    cw = np.concatenate([data, synth.rs_parity(data)], axis=1)
    for nerr in range(0, 8):
        c = cw[nerr % 6].copy()
        pos = rng.choice(120, nerr, replace=False)
        c[pos] ^= rng.integers(1, 256, nerr).astype(np.uint8)
        r, be, out = orc.rs_decode(c)
        if nerr <= 5:
            assert r == nerr and np.array_equal(out, cw[nerr % 6])
        else:
            assert r == -1 or not np.array_equal(out, cw[nerr % 6])


def test_sample_clock_offset_and_multipath():
    """+-60 ppm receiver clock error (frames are 196608 * (1 -+ 60e-6) samples apart) and a 3-path channel: every frame is
    synchronised on its own, slots stay consecutive, payload exact"""
    ens = synth.config1_ensemble(seed=41)
    for ppm in (60.0, -60.0):
        imp = synth.Impairments(snr_db=18.0, cfo_hz=321.0, taps=[(25, 0.5), (90, 0.3j)], sample_offset_ppm=ppm, seed=6, lead_in=20000)
        raw, res = synth.synthesize(ens, 10, "u8", imp, seed=6, tail=150000)
        d = orc.decode(raw, "u8", [orc.subch_struct(ens.subchannels[0])])
        assert d["nf"] == 10 and d["fib_ok"].all() and (d["t_f"] >= 0).all()
        spacing = np.diff(d["t_f"])
        assert np.all(np.abs(spacing - 196608 / (1 + ppm * 1e-6)) <= 3)
        assert np.array_equal(d["fibs"], res.fibs[:40])
        assert np.array_equal(d["frames"][0], res.logical[3][:d["frames"][0].shape[0]])


def test_ofdm_spectrum_layout_is_conflict_free():
    """The slot table of k_ofdm_demod's spectrum buffer: a permutation of the 1536 QPSK symbols that keeps every symbol inside its
    group of 16 (so the DQPSK reads n = t + 256 j of a half-warp cover 16 distinct 8-byte banks) and gives the 16 lanes of every
    half-warp store of the last FFT pass 16 distinct banks as well (proper edge colouring of the store-group x read-group graph)."""
    from abracadabra_b200 import pipeline
    tab = pipeline.spectrum_layout().astype(np.int64)
    by_p, by_n = tab[:2048], tab[2048:]
    assert sorted(by_n.tolist()) == list(range(1536))                       # bijection
    assert np.array_equal(by_n >> 4, np.arange(1536) >> 4)                   # stays inside its read group
    for g in range(96):
        assert len(set((by_n[16 * g:16 * g + 16] & 15).tolist())) == 16      # read side: 16 banks per half-warp
    used = by_p[by_p != 0xFFFF]
    assert used.size == 1536 and sorted(used.tolist()) == list(range(1536))
    # store side: thread t, quadruple h, output j writes position p = 4 (t + 256 h) + j; a half-warp is 16 consecutive t
    worst = 0
    for hw in range(16):
        for h in range(2):
            for j in range(4):
                slots = [by_p[4 * (t + 256 * h) + j] for t in range(16 * hw, 16 * hw + 16)]
                banks = [s & 15 for s in slots if s != 0xFFFF]
                assert len(banks) == len(set(banks))
                worst = max(worst, len(banks))
    assert worst == 16


def test_viterbi_decision_word_layout():
    """Bit layout of the survivor words k_viterbi writes and k_vit_trace reads (abracadabra_b200/csrc/k_viterbi.cuh): a lane of role r
    puts the decision of butterfly J (0..15), output o (0/1), at bit 16 (J >> 3) + 8 o + (J & 7) of its word; role 1 works on
    complemented states.  The trace kernel reads w = x | (~brev(y)) << 32 with one index formula per direction: check that the
    formula (the kernel's own, through b200dab_viterbi_decision_index) finds, for every logical state, the bit its owner wrote, and
    that the 64 positions are all different."""
    import ctypes as ct
    from abracadabra_b200 import pipeline
    L = pipeline.lib()
    L.b200dab_viterbi_decision_index.argtypes = [ct.c_int, ct.c_int]
    L.b200dab_viterbi_decision_index.restype = ct.c_int
    assert L.b200dab_viterbi_decision_index(64, 0) == -1

    def pos(j, o):
        return 16 * (j >> 3) + 8 * o + (j & 7)

    def brev32(i):
        return 31 - i

    fwd, bwd = set(), set()
    for n in range(64):
        # forward step: new state n is computed by role (n >> 1) & 1; in that role's physical coordinates it is output p & 1 of butterfly p >> 2
        role = (n >> 1) & 1
        p = n ^ 63 if role else n
        assert (p & 2) == 0
        written = pos(p >> 2, p & 1)
        in_w = 32 + brev32(written) if role else written           # role 1's word is bit-reversed into the upper half
        idx = L.b200dab_viterbi_decision_index(n, 0)
        assert idx == in_w == ((n >> 5) << 4) | ((n & 1) << 3) | ((n >> 2) & 7) | ((n & 2) << 4)
        fwd.add(idx)
        # backward step: state n is computed by role n & 1; physically it is output p >> 5 of butterfly (p >> 1) & 15
        role = n & 1
        p = n ^ 63 if role else n
        assert (p & 1) == 0
        written = pos((p >> 1) & 15, p >> 5)
        in_w = 32 + brev32(written) if role else written
        idx = L.b200dab_viterbi_decision_index(n, 1)
        assert idx == in_w == (((n >> 4) & 1) << 4) | ((n >> 5) << 3) | ((n >> 1) & 7) | ((n & 1) << 5)
        bwd.add(idx)
    assert len(fwd) == 64 and len(bwd) == 64


def test_viterbi_four_lane_layout_matches_model():
    """Four-lane mode of k_viterbi (long codewords): the scalar model of the lane split (tools/vit4_model.py: register maps, shuffle
    sources, per-lane signs) reproduces the plain recursion of oracle/daboracle.c:orc_viterbi, metrics and decisions, and the
    trace kernel's closed-form index function (through b200dab_viterbi4_decision_index) equals the tables the model derives."""
    import ctypes as ct
    import importlib.util
    from pathlib import Path
    from abracadabra_b200 import pipeline
    spec = importlib.util.spec_from_file_location("vit4_model", Path(__file__).resolve().parent.parent / "tools" / "vit4_model.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main()                                            # asserts inside
    fwd, bwd = mod.build_luts()
    L = pipeline.lib()
    L.b200dab_viterbi4_decision_index.restype = ct.c_int
    assert L.b200dab_viterbi4_decision_index(-1, 0) == -1
    for s in range(64):
        assert L.b200dab_viterbi4_decision_index(s, 0) == fwd[s]
        assert L.b200dab_viterbi4_decision_index(s, 1) == bwd[s]
    assert len({v & 63 for v in fwd}) == 64 and len({v & 63 for v in bwd}) == 64


def test_next_configuration_is_not_mistaken_for_the_current_one():
    """Multiplex reconfiguration (EN 300 401 sec. 6.4): FIG 0/1 / 0/2 sent with C/N = 1 announce the NEXT configuration and must not
    enter the sub-channel table the MSC is decoded with, nor the service list, before the change; after it both hold the new entries."""
    ens = synth.config1_ensemble(seed=31)
    ens.reconfig = dict(at=140, lead=80, subchannels=[synth.SubChannel(9, 200, "dabplus", 64, ("EEP", 3, "A"), seed=3109)],
                        services=[synth.Service(0x1234, "ADDED SERVICE", 9)])
    _, _, truth = synth.build_cif_bits(ens, 200, 5)
    fibs = truth["fibs"].reshape(-1, 32)
    announced = fibs[3 * 60:3 * 139]
    assert any(f[0] == 0x06 and f[4] >> 6 == 3 for f in announced)                # FIG 0/0 with change flags + occurrence change
    assert any((f[i] & 0xE0) == 0 and f[i + 1] == 0x81 for f in announced for i in range(0, 8) if f[i] != 0xFF)   # FIG 0/1 with C/N = 1
    assert [s.subch_id for s in pipeline.parse_subch(fibs[:3 * 139])] == [3]
    assert [s.subch_id for s in pipeline.parse_subch(fibs)] == [3, 9]
    e = pipeline.parse_ensemble(fibs[:3 * 136])
    assert [s["sid"] for s in e["services"]] == [0x1001]
    e = pipeline.parse_ensemble(fibs)
    assert sorted(s["sid"] for s in e["services"]) == [0x1001, 0x1234]
    assert [s["label"].rstrip() for s in e["services"] if s["sid"] == 0x1234] == ["ADDED SERVICE"]


def test_chunk_chaining_and_core_partition():
    """Host logic of the time-chunked decode (shard.decode_chunked) and of the per-rank core slices (shard.bind_to_gpu)."""
    from abracadabra_b200 import shard
    T = shard.T_FRAME
    # 3 chunks of one recording: frames every T samples from 5000 on, the overlaps found twice, one frame lost in chunk 1
    t_all = [5000 + k * T for k in range(40)]
    c0 = t_all[:16]
    c1 = [t for t in t_all[10:30] if t != t_all[20]]
    slots1 = {t: (t - c1[0] + T // 2) // T for t in c1}
    a1 = [-1] * (max(slots1.values()) + 1)
    for t, s in slots1.items():
        a1[s] = t
    c2 = t_all[25:40]
    chain, slot0 = shard.chain_frames([c0, a1, c2])
    assert slot0 == [0, 10, 25]
    got = [g for _, _, g in chain]
    assert got == [k for k in range(40) if k != 20]                      # every frame once, the lost one leaves a gap
    assert [c for c, _, _ in chain] == [0] * 16 + [1] * 13 + [2] * 10    # a frame is taken from the first chunk that has it
    # frame numbering that disagrees between a chunk and the recording is an error, not silently merged
    bad = [c0, [t + 3 * T // 2 if i > 8 else t for i, t in enumerate(t_all[10:30])], c2]
    with pytest.raises(RuntimeError):
        shard.chain_frames(bad)
    r = shard.chunk_recording(60 * 2048000, 4)
    assert r[0][0] == 0 and r[-1][1] == 60 * 2048000 and all(a % shard.CHUNK_ALIGN == 0 for a, _ in r)
    assert all(r[i][1] - r[i + 1][0] >= shard.CHUNK_OVERLAP_FRAMES * T for i in range(3))
    # eight GPUs reporting the same 32 local cores get eight disjoint slices of four; two NUMA halves are split separately
    same = [list(range(32))] * 8
    parts = [shard.partition_cpus(same, g) for g in range(8)]
    assert sorted(c for p in parts for c in p) == list(range(32)) and all(len(p) == 4 for p in parts)
    numa = [list(range(0, 16))] * 4 + [list(range(16, 32))] * 4
    parts = [shard.partition_cpus(numa, g) for g in range(8)]
    assert parts[0] == [0, 1, 2, 3] and parts[4] == [16, 17, 18, 19] and sorted(c for p in parts for c in p) == list(range(32))
    assert shard.partition_cpus([[0, 1]] * 4, 3) == [0, 1]                  # fewer cores than ranks: left alone
