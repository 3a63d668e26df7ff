"""Pins the CPU checker (oracle/daboracle.c) against the reference's own receiver core: tests/golden/*.json.gz hold the
observable results of libdabsdr 4.0.1 (run through oracle/_ref/ref_harness by tests/golden/make_golden.py) on seeded
synthetic files; the checker must reproduce them -- AU bytes exactly, CFO and TII levels within the stated tolerance."""
import hashlib

import numpy as np
import pytest

import golden_cases as gc
import orc

SERVICE_SUBCH = {"c1_u8_40db": 0, "c1_u8_4p5db": 0, "c1_s16_12db": 0, "multi_u8_25db": 2}


def _oracle_aus(raw, fmt, sc, frames=None):
    d = orc.decode(raw, fmt, [orc.subch_struct(sc)])
    fr = d["frames"][0]
    s = sc.bitrate // 8
    # superframe alignment: first row whose fire code matches
    best = None
    for r0 in range(5):
        hits = sum(1 for r in range(r0, fr.shape[0] - 4, 5)
                   if orc.lib().orc_firecode(fr[r, 2:11].ctypes.data) == ((int(fr[r, 0]) << 8) | int(fr[r, 1])) and fr[r, :3].any())
        if best is None or hits > best[0]:
            best = (hits, r0)
    aus, be, unc = [], 0, 0
    for r in range(best[1], fr.shape[0] - 4, 5):
        sf = orc.superframe(fr[r:r + 5].reshape(-1), s)
        be += sf["bit_errors"]; unc += sf["uncorrectable"]
        aus += [a for a, ok in zip(sf["aus"], sf["crc_ok"]) if ok]
    return d, aus, be, unc


@pytest.mark.parametrize("name", sorted(SERVICE_SUBCH))
def test_access_units_match_reference(name):
    g = gc.load_golden(name)
    raw, res = gc.build_case(g["spec"])
    if not gc.raw_matches(raw, g):
        pytest.skip("synthetic file differs from the one the fixture was made from (numpy RNG stream changed)")
    sc = res.ens.subchannels[SERVICE_SUBCH[name]]
    d, aus, be, unc = _oracle_aus(raw, g["spec"]["fmt"], sc)
    ref = [e for e in g["events"] if e["ev"] == "audio" and e["len"] > 0]
    assert len(ref) >= 20
    mine = [hashlib.sha1(a).hexdigest() for a in aus]
    ref_sha = [e["sha1"] for e in ref]
    # the reference starts delivering after service selection and stops 4 superframes before the end of the file
    # (SURVEY.md sec. 8 a7): its AU stream must be a contiguous run of the checker's
    i0 = mine.index(ref_sha[0])
    n = min(len(mine) - i0, len(ref_sha))
    assert n >= len(ref_sha) - 3 and mine[i0:i0 + n] == ref_sha[:n]      # the reference squeezes <= 1 more superframe out of the tail
    assert all(e["header"] == 0x70 and e["ascty"] == 63 for e in ref)
    assert unc == 0
    # all of them equal the modulator's payload as well
    truth = [hashlib.sha1(a).hexdigest() for a in res.aus[sc.subchid]]
    assert all(m in truth for m in mine)
    # CFO: reference reports Hz x 10 in the periodic notification
    ref_cfo = [e["freqOffset"] / 10.0 for e in g["events"] if e["ev"] == "periodic" and "freqOffset" in e and e["level"] == 3][1:]
    tol = 1.0 if g["spec"]["snr"] >= 12 else 12.0       # Hz; the reference itself wanders +-2..9 Hz at 4-12 dB
    assert abs(np.median(d["cfo"]) - np.median(ref_cfo)) <= tol
    assert abs(np.median(d["cfo"]) - g["spec"]["cfo"]) <= tol
    # pre-RS error rate: same order as the reference's rsBitErrors (both decoders see the same noise)
    ref_be = sum(e["rsBitErr"] for e in g["events"] if e["ev"] == "periodic" and "rsBitErr" in e)
    if g["spec"]["snr"] >= 7:
        assert be == 0 and ref_be == 0
    else:
        n_ref_sf, n_my_sf = len(ref) / 3.0, len(aus) / 3.0
        assert 0.4 < (be / n_my_sf) / (ref_be / n_ref_sf) < 2.5


def test_mp2_frames_match_reference():
    """DAB (MPEG-1 Layer II) service behind UEP protection: the reference delivers every logical frame whose ISO CRC holds;
    the checker's Viterbi output must contain exactly those bytes"""
    g = gc.load_golden("multi_mp2_u8_25db")
    raw, res = gc.build_case(g["spec"])
    if not gc.raw_matches(raw, g):
        pytest.skip("synthetic file differs from the one the fixture was made from")
    sc = res.ens.subchannels[3]
    assert sc.kind == "mp2" and sc.prot == ("UEP", 35)
    d = orc.decode(raw, "u8", [orc.subch_struct(sc)])
    mine = [hashlib.sha1(bytes(r)).hexdigest() for r in d["frames"][0]]
    ref = [e for e in g["events"] if e["ev"] == "audio"]
    assert len(ref) >= 30 and all(e["ascty"] == 0 and e["len"] == 384 for e in ref)
    ref_sha = [e["sha1"] for e in ref]
    i0 = mine.index(ref_sha[0])
    n = min(len(mine) - i0, len(ref_sha))
    assert n >= len(ref_sha) - 4 and mine[i0:i0 + n] == ref_sha[:n]
    crc = [e for e in g["events"] if e["ev"] == "periodic" and e.get("crcOk")]
    assert crc and all(e["crcErr"] == 0 for e in crc)


@pytest.mark.parametrize("name", ["c1_u8_40db", "tii3_u8_25db"])
def test_tii_matches_reference(name):
    g = gc.load_golden(name)
    raw, res = gc.build_case(g["spec"])
    if not gc.raw_matches(raw, g):
        pytest.skip("synthetic file differs from the one the fixture was made from")
    fmt = g["spec"]["fmt"]
    d = orc.decode(raw, fmt, [])
    assert d["nf"] >= 40
    # the reference emits its first TII list after 40 frames; frames whose first CIF counter is 0 mod 8 carry TII
    fold = np.array([orc.tii_fold(orc.tii_null_spectrum(raw, fmt, d["t_f"][f], d["cfo"][f])) for f in range(40)])
    norm = np.array([orc.prs_rms(raw, fmt, d["t_f"][f], d["cfo"][f]) for f in range(40)])
    sign = np.array([1 if (4 * f) % 8 == 0 else -1 for f in range(40)])
    ids, spec = orc.tii_detect(fold, norm, sign)
    ref = [e for e in g["events"] if e["ev"] == "tii" and e["n"] > 0][0]
    assert [(m, s) for m, s, _ in ids] == [(m, s) for m, s, _ in ref["ids"]]          # ids and their order: exact
    mine_lv = {(m, s): l for m, s, l in ids}
    ref_lv = {(m, s): l for m, s, l in ref["ids"]}
    top = max(ref_lv, key=ref_lv.get)
    for k in ref_lv:
        # RadioControl shows 20 log10(level / maxLevel) (src/radiocontrol.cpp:397-412).  Tolerance: 0.1 dB on the level the
        # reference reports (BASELINE.json north_star); 0.15 dB for a transmitter more than 12 dB below the strongest one
        rel_ref = 20 * np.log10(ref_lv[k] / ref_lv[top]); rel_me = 20 * np.log10(mine_lv[k] / mine_lv[top])
        assert abs(20 * np.log10(mine_lv[k] / ref_lv[k])) <= (0.1 if rel_ref > -12 else 0.15), (k, mine_lv[k], ref_lv[k])
        assert abs(rel_ref - rel_me) <= (0.1 if rel_ref > -12 else 0.15), (k, rel_ref, rel_me)   # weakest (-13.5 dB): 0.15 dB
    rs = np.array(ref["spectrum"])
    strong = np.nonzero(rs > 0.3 * rs.max())[0]
    assert set(strong) <= set(np.nonzero(spec > 0.15 * spec.max())[0])
    # same bins light up; normalised by the maximum as src/tii/tiibackend.cpp:443 does, the lit bins agree within 1 %
    assert np.all(np.abs(spec[strong] / spec.max() - rs[strong] / rs.max()) < 0.01)


def test_reference_structures_in_fixture():
    """what the drop-in must hand to RadioControl for the config-1 ensemble (checked again on the GPU in test_gpu_shim.py)"""
    g = gc.load_golden("c1_u8_40db")
    ens = [e for e in g["events"] if e["ev"] == "ensemble" and e["ueid"]][0]
    assert ens["ueid"] == (0xE1 << 16) | 0x1001 and ens["label"].strip() == "B200 TEST MUX" and ens["freq"] == 227360
    sl = [e for e in g["events"] if e["ev"] == "service_list"][0]
    assert sl["n"] == 1 and sl["items"][0]["sid"] == 0x1001
    sc = [e for e in g["events"] if e["ev"] == "sc_list"][0]["items"][0]
    assert (sc["subchid"], sc["addr"], sc["size"], sc["prot"], sc["ascty"], sc["bitrate"], sc["ps"]) == (3, 0, 72, 8, 63, 96, 2)


def test_reference_pad_fixtures():
    """What the reference core reports for programme associated data (pins abracadabra_b200/csrc/pad.h, see also
    tests/test_rawfile_scanner.py::test_xpad_dynamic_label_and_data_groups)"""
    g = gc.load_golden("c1_mot_u8_30db")
    dg = [e for e in g["events"] if e["ev"] == "dg"]
    sent = [d.hex() for d in gc.mot_data_groups(g["spec"]["xpad_dg"])]
    assert len(dg) >= 12 and all(e["id"] == 0 and e["scid"] == 0xFFFF and e["uatype"] == 2 and e["data"] in sent for e in dg)
    ua = [e for e in g["events"] if e["ev"] == "ua_list"][0]
    assert ua["items"] == [{"type": 2, "len": 2, "data": "0c3c"}]
    per = [e for e in gc.load_golden("c1_dls_u8_30db")["events"] if e["ev"] == "periodic" and e.get("crcOk")]
    assert per and all(e["padBytes"] == 16 * e["crcOk"] for e in per)        # 16 = bytes of the PAD data_stream_element per AU


def test_spectrum_callback_is_power_spectrum_of_raw_blocks():
    """What the reference's signal-spectrum callback delivers (dabsdr.h:365-370, :393; consumer src/signalbackend.cpp:383-431),
    pinned by a fixture with all 2048 bins of the first callbacks: |FFT|^2 (unnormalised, FFT order) of 2048 RAW samples -- SIGNAL
    (type 0): the block starting 2552 samples before the feed position of the callback; NULL / TII (types 1 / 2): 2656 before.
    Sum ratio 1 within 1e-5; bins within 5e-3 of the maximum (the core's CFO tracking turns the block by a fraction of a hertz)."""
    g = gc.load_golden("c1_spectrum_u8_30db")
    raw, res = gc.build_case(g["spec"])
    if not gc.raw_matches(raw, g):
        pytest.skip("synthetic file differs from the one the fixture was made from")
    x = (raw.astype(np.float32) - 128).view(np.float32).reshape(-1, 2)
    x = x[:, 0] + 1j * x[:, 1]
    sp = [e for e in g["events"] if e["ev"] == "spectrum_bins"][1:]          # the very first one is taken before the core has synchronised
    assert len(sp) >= 12 and {e["type"] for e in sp} == {0, 1, 2}
    for e in sp:
        b = np.array(e["bins"])
        w0 = e["pos"] - (2552 if e["type"] == 0 else 2656)
        P = np.abs(np.fft.fft(x[w0:w0 + 2048])) ** 2
        assert abs(b.sum() / P.sum() - 1) < 1e-5
        assert np.abs(P / P.max() - b / b.max()).max() < 5e-3
        # the orc FFT (the arithmetic the CUDA kernel mirrors) gives the same bins
        Q = np.abs(orc.fft2048(x[w0:w0 + 2048].astype(np.complex64))) ** 2
        assert np.abs(Q / Q.max() - P / P.max()).max() < 1e-5


def test_uep_table_index_23_follows_the_standard():
    """the UEP table inside the reference binary has PI2 = 7 for index 23 (80 kbit/s, level 1, 84 CU) where EN 300 401
    table 8 -- and the arithmetic of the sizes -- say 17.  Decided by experiment and documented here: for a sub-channel coded per the
    standard (PI 24 / 17 / 12 / 18) the reference core gets the first segment (192 bits: header and ISO CRC, so its crcOk counter is
    happy) and then garbage from the first bit of the second segment on -- it really de-punctures with 7.  This path (csrc/tables.h)
    follows the standard: the checker decodes the same file to the transmitted frames.  Index 23 is therefore the one place where
    'identical to the reference' is deliberately not the bar."""
    if not orc.have_ref():
        pytest.skip("reference core not built (oracle/_ref)")
    import tempfile
    from pathlib import Path
    from abracadabra_b200 import synth
    sub = synth.SubChannel(5, 0, "mp2", 80, ("UEP", 23), seed=7)
    assert sub.profile()[0][1] == (10, 17)
    ens = synth.EnsembleConfig(eid=0x2323, label="UEP23", subchannels=[sub], services=[synth.Service(0x2323, "UEP 23 SVC", 5)])
    raw, res = synth.synthesize(ens, 24, "u8", synth.Impairments(snr_db=30.0, cfo_hz=10.0, seed=2), seed=2, tail=150000)
    with tempfile.TemporaryDirectory() as td:
        p = Path(td) / "u.u8"
        raw.tofile(p)
        summ, ev = orc.run_ref(p, "u8", Path(td) / "log.jsonl", service=0, extra=["--tail-frames", "12"])
    sc = [e for e in ev if e["ev"] == "sc_list"][0]["items"][0]
    assert (sc["uep_fec"], sc["size"], sc["bitrate"]) == (23, 84, 80)
    au = [np.frombuffer(bytes.fromhex(e["data"]), np.uint8) for e in ev if e["ev"] == "audio" and e["len"] > 0]
    assert len(au) >= 40 and summ["crc_err"] == 0
    truth = res.logical[5]
    for a in au[3:12]:
        f = next(f for f in truth if np.array_equal(f[:16], a[:16]))             # header + first segment are right ...
        diff = np.nonzero(f != a)[0]
        assert diff.size > 100 and diff[0] == 24                                  # ... the rest is not: it breaks at bit 192 = segment 2
    o = orc.decode(raw, "u8", [orc.subch_struct(sub)])
    assert np.array_equal(o["frames"][0], truth[:o["frames"][0].shape[0]])
