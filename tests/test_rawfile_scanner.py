"""Raw/UFF container rules (src/input/rawfileinput.cpp, src/input/inputdevicerecorder.cpp) and scanner CSV (src/tii/scannerbackend.cpp)."""
import numpy as np
import pytest

from abracadabra_b200 import pipeline, rawfile, scanner, synth
import cases


def test_uff_roundtrip(tmp_path):
    rng = np.random.default_rng(1)
    for fmt, dt in (("u8", np.uint8), ("s16", np.int16)):
        iq = rng.integers(0, 200, 2 * 5000).astype(dt)
        p = tmp_path / f"rec_{fmt}.uff"
        rawfile.write_uff(str(p), iq, fmt, 227360)
        blob = p.read_bytes()
        assert len(blob) == 2048 + iq.nbytes and blob[:5] == b"<?xml"
        z = blob.index(b"\x00")
        assert z < 2048 and blob[z:2048] == bytes(2048 - z)            # zero padding up to the data offset
        data, info = rawfile.read_raw(str(p))
        assert info.has_xml_header and info.fmt == fmt and info.frequency_khz == 227360 and info.sample_rate == 2048000
        assert info.num_samples == 5000 and np.array_equal(data, iq)
        assert "AbracaDABra" in info.recorder and info.time.endswith("UTC")


def test_uff_header_fields_reference_layout(tmp_path):
    # the element / attribute names InputDeviceRecorder writes (inputdevicerecorder.cpp:195-263)
    import xml.etree.ElementTree as ET
    p = tmp_path / "a.uff"
    rawfile.write_uff(str(p), np.zeros(64, np.int16), "s16", 174928)
    root = ET.fromstring(p.read_bytes()[:2048].rstrip(b"\x00").decode())
    assert root.tag == "SDR" and [e.tag for e in root] == ["Recorder", "Device", "Time", "Sample", "Datablocks"]
    ch = root.find("Sample/Channels")
    assert ch.get("Bits") == "16" and ch.get("Container") == "int16" and ch.get("Ordering") == "LSB"
    assert [c.get("Value") for c in ch] == ["I", "Q"]
    blk = root.find("Datablocks/Datablock")
    assert blk.get("Count") == "64" and blk.get("Offset") == "2048" and blk.get("Unit") == "Channel" and blk.get("Number") == "1"
    assert blk.find("Frequency").get("Unit") == "kHz" and blk.find("Modulation").get("Value") == "DAB"


def test_foreign_header_variants():
    def hdr(body):
        return (b'<?xml version="1.0"?><SDR>' + body + b"</SDR>").ljust(2048, b"\x00")
    i = rawfile.parse_header(hdr(b'<Sample><Samplerate Value="2048" Unit="kHz"/><Channels Bits="8" Container="uint8"/></Sample>'
                                 b'<Datablocks><Datablock Count="1000"><Frequency Value="202928000" Unit="Hz"/></Datablock></Datablocks>'))
    assert i.sample_rate == 2048000 and i.fmt == "u8" and i.frequency_khz == 202928 and i.num_samples == 500
    with pytest.raises(ValueError):
        rawfile.parse_header(hdr(b'<Sample><Channels Container="float"/></Sample>'))
    with pytest.raises(ValueError):
        rawfile.parse_header(hdr(b'<Sample><Channels Container="int16" Ordering="MSB"/></Sample>'))
    assert rawfile.parse_header(b"\x80" * 2048) is None                # raw samples: no zero byte / no XML
    assert rawfile.parse_header(b"not xml\x00".ljust(2048, b"\x01")) is None


def test_plain_raw_and_filename_frequency(tmp_path):
    assert rawfile.frequency_from_name("/x/2024-01-01_12C_ensemble.raw") == 227360
    assert rawfile.frequency_from_name("/x/5A.raw") == 174928
    assert rawfile.frequency_from_name("/x/scan_13F_x.tar.raw") == 239200
    assert rawfile.frequency_from_name("/x/nothing.raw") == 0
    iq = (np.arange(2 * 3000 + 1) % 251).astype(np.uint8)
    p = tmp_path / "rec_11D_test.raw"
    iq.tofile(p)
    data, info = rawfile.read_raw(str(p))
    assert not info.has_xml_header and info.fmt == "u8" and info.frequency_khz == 222064 and info.num_samples == 3000
    assert np.array_equal(data, iq[:6000])


def test_header_count_larger_than_file(tmp_path):
    p = tmp_path / "t.uff"
    rawfile.write_uff(str(p), np.zeros(2000, np.uint8), "u8", 227360)
    blob = bytearray(p.read_bytes())
    with open(p, "wb") as f:
        f.write(bytes(blob[:2048 + 1000]))                             # truncated recording
    data, info = rawfile.read_raw(str(p))
    assert data.size == 1000 and info.num_samples == 500


def test_parse_ensemble_from_fibs():
    ens = synth.full_ensemble()
    ens.cif_count0 = 40
    _, _, truth = synth.build_cif_bits(ens, 16, 5)
    fibs = truth["fibs"].reshape(-1, 32)
    e = pipeline.parse_ensemble(fibs)
    assert e["eid"] == ens.eid and e["ecc"] == ens.ecc and e["label"] == ens.label.ljust(16)[:16] and e["first_cif_count"] == 40
    assert e["lto"] == ens.lto
    assert [s["sid"] for s in e["services"]] == [s.sid for s in ens.services]
    assert [s["label"] for s in e["services"]] == [s.label.ljust(16)[:16] for s in ens.services]
    assert [s["components"][0][0] for s in e["services"]] == [s.subchid for s in ens.services]
    kind = {sc.subchid: sc.kind for sc in ens.subchannels}
    assert [s["components"][0][1] for s in e["services"]] == [63 if kind[s.subchid] == "dabplus" else 0 for s in ens.services]
    # FIBs flagged bad are skipped: nothing left -> empty ensemble
    e0 = pipeline.parse_ensemble(fibs, np.zeros(fibs.shape[0], np.uint8))
    assert e0["ueid"] == 0 and e0["services"] == [] and e0["first_cif_count"] == -1


def test_scan_row_format():
    r = scanner.ScanRow("2026-01-01 00:00:00", "12C", 227360, 0xE01234, "LABEL           ", 12, 23.44, 7, 3, -3.21)
    assert r.csv() == "2026-01-01 00:00:00;12C;227360;E01234;LABEL;12;23.4;;7;3; -3.2;;;;"
    assert scanner.HEADER.count(";") == r.csv().count(";") == 14
    r = scanner.ScanRow("t", "5A", 174928, 0x1001, "X", 1, 9.96, None, None, None)
    assert r.csv() == "t;5A;174928;  1001;X;1;10.0;;;;;;;;"


@pytest.mark.gpu
def test_scan_uff_files(tmp_path):
    import datetime
    specs = [("12C", 227360, dict(tii=[(7, 3, 1.0), (7, 5, 0.5)])), ("5A", 174928, dict(tii=[(21, 1, 1.0)]))]
    paths = []
    for k, (chn, f, kw) in enumerate(specs):
        ens = synth.config1_ensemble()
        ens.eid = 0x1000 + k
        ens.label = f"SCAN {chn}"
        ens.tii = [synth.TiiTx(*t) for t in kw["tii"]]
        raw, _ = synth.synthesize(ens, 100, "s16", synth.Impairments(snr_db=30, cfo_hz=120.0 * (k + 1), lead_in=5000 + 1000 * k, seed=k,
                                                                     rms=7000.0), seed=k, tail=20000)
        p = tmp_path / f"scan_{k}.uff"
        rawfile.write_uff(str(p), raw, "s16", f)
        paths.append(str(p))
    noise = (np.random.default_rng(3).normal(0, 300, 2 * 196608 * 12)).astype(np.int16)
    p = tmp_path / "x_9B_noise.raw.s16"
    noise.tofile(p)
    paths.append(str(p))
    lines = scanner.scan_files(paths)
    assert lines[0] == scanner.HEADER
    rows = [l.split(";") for l in lines[1:]]
    assert [(r[1], r[2], r[3].strip(), r[4], r[8], r[9]) for r in rows] == [
        ("12C", "227360", "E11000", "SCAN 12C", "7", "3"), ("12C", "227360", "E11000", "SCAN 12C", "7", "5"),
        ("5A", "174928", "E11001", "SCAN 5A", "21", "1")]
    assert float(rows[0][10]) == 0.0 and abs(float(rows[1][10]) + 6.0) < 0.3 and float(rows[2][10]) == 0.0
    assert all(abs(float(r[6]) - 30) < 3 for r in rows) and all(r[5] == "1" for r in rows)


def test_xpad_dynamic_label_and_data_groups():
    """pad.h against the modulator's X-PAD stream and against the reference core's dynamic-label callbacks (golden fixture)."""
    import golden_cases as gc
    g = gc.load_golden("c1_dls_u8_30db")
    ref_dl = [bytes.fromhex(e["data"]) for e in g["events"] if e["ev"] == "dl"]
    assert len(ref_dl) >= 30 and all(e["len"] == len(d) for e, d in zip([e for e in g["events"] if e["ev"] == "dl"], ref_dl))
    # transmitted stream -> decoder: every group comes out once, in order, CRC intact
    sc = synth.SubChannel(3, 0, "dabplus", 96, ("EEP", 3, "A"), seed=1003, dls=list(g["spec"]["dls"]),
                          xpad_dg=[bytes(range(7, 60)), b"\x55" * 130])
    for dg in range(len(sc.xpad_dg)):                       # MSC data groups end with their CRC-16
        body = np.frombuffer(sc.xpad_dg[dg], np.uint8)
        crc = int(synth.crc16_np(body[None, :])[0])
        sc.xpad_dg[dg] = sc.xpad_dg[dg] + bytes([crc >> 8, crc & 0xFF])
    _, aus = synth.dabplus_superframes(sc, 40, np.random.default_rng(5))
    dec = pipeline.PadDecoder()
    assert all(dec.feed(a) for a in aus)
    sent = sc._pad_sent
    n = len(dec.out)
    assert n >= 20 and [(0 if k == "dl" else 12, d) for k, d in sent[:n]] == dec.out
    assert {k for k, _ in dec.out} == {0, 12}
    msgs = pipeline.dynamic_label_text([d for k, d in dec.out if k == 0])
    assert msgs[:3] == list(g["spec"]["dls"])
    # a corrupted group is dropped, the stream recovers
    dec2 = pipeline.PadDecoder()
    bad = bytearray(aus[0]); bad[6] ^= 0x40
    for a in [bytes(bad)] + aus[1:]:
        dec2.feed(a)
    assert dec2.out == dec.out[1:]
    assert not pipeline.PadDecoder().feed(b"\x21" + bytes(40))          # first element is not a DSE: no PAD
    # the reference reports exactly the transmitted dynamic-label groups (from the point where it starts delivering audio)
    sc3 = synth.SubChannel(3, 0, "dabplus", 96, ("EEP", 3, "A"), seed=1003, dls=list(g["spec"]["dls"]))
    synth.dabplus_superframes(sc3, 40, np.random.default_rng(5))
    tx = [d for _, d in sc3._pad_sent]
    k0 = tx.index(ref_dl[0])
    assert tx[k0:k0 + len(ref_dl)] == ref_dl
    dec.close(); dec2.close()
    # DAB (MPEG Layer II) frames: PAD at the end of the frame, split by the ScF-CRC; pinned by the reference's log
    g2 = gc.load_golden("mp2_dls_u8_30db")
    _, res2 = gc.build_case(g2["spec"])
    dec3 = pipeline.PadDecoder()
    assert all(dec3.feed_mp2(bytes(fr)) for fr in res2.logical[9])
    ref2 = [bytes.fromhex(e["data"]) for e in g2["events"] if e["ev"] == "dl"]
    mine2 = [d for k, d in dec3.out if k == 0]
    i0 = mine2.index(ref2[0])
    assert len(ref2) >= 40 and mine2[i0:i0 + len(ref2)] == ref2
    assert pipeline.dynamic_label_text(mine2)[:2] == list(g2["spec"]["dls"])
    dec3.close()


def test_packet_mode_data_groups():
    """Packet mode (csrc/pad.h packet::Assembler) against the modulator's packet stream and the reference core's data group
    callbacks (tests/golden/c1_packet_u8_30db.json.gz)"""
    import golden_cases as gc
    g = gc.load_golden("c1_packet_u8_30db")
    pk = g["spec"]["packet"]
    _, res = gc.build_case(g["spec"])
    sc = [s for s in res.ens.subchannels if s.kind == "packet"][0]
    frames = res.logical[pk["subchid"]]
    for addr in (pk["addr"], pk["other_addr"]):
        dec = pipeline.PadDecoder()
        for fr in frames:
            dec.feed_packets(addr, bytes(fr))
        sent = [d for a, d in sc._packets_sent if a == addr]
        got = [d for k, d in dec.out]
        assert len(got) >= 20 and all(k == -1 for k, _ in dec.out) and got == sent[:len(got)] and len(sent) - len(got) <= 1
        dec.close()
    ref = [e for e in g["events"] if e["ev"] == "dg"]
    assert len(ref) >= 50 and all(e["id"] == -1 and e["scid"] == pk["scid"] and e["uatype"] == 2 for e in ref)
    mine = [d.hex() for a, d in sc._packets_sent if a == pk["addr"]]
    i0 = mine.index(ref[0]["data"])
    assert [e["data"] for e in ref] == mine[i0:i0 + len(ref)]
    # a damaged packet drops its data group only; the following groups come through
    dec = pipeline.PadDecoder()
    bad = frames.copy()
    bad[3, 5] ^= 0x10
    for fr in bad:
        dec.feed_packets(pk["addr"], bytes(fr))
    got = [d.hex() for k, d in dec.out]
    assert len(mine) - 3 <= len(got) < len(mine) and set(got) <= set(mine)
    dec.close()
    # the component list of the reference shows the packet-mode component as the FIGs describe it
    scl = [e for e in g["events"] if e["ev"] == "sc_list"][0]["items"]
    assert scl[1]["tmid"] == 3 and scl[1]["scid"] == pk["scid"] and scl[1]["paddr"] == pk["addr"] and scl[1]["dscty"] == 60 and scl[1]["subchid"] == pk["subchid"]
