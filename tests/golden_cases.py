"""Seeded definitions of the golden-fixture inputs (shared by tests/golden/make_golden.py and the tests)."""
import gzip
import hashlib
import json
from pathlib import Path

from abracadabra_b200 import synth

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"

CASES = {
    # BASELINE.json configs[0] in small: FIC + one DAB+ 96 kbps EEP-3A sub-channel, TII (12,5), u8
    "c1_u8_40db": dict(kind="config1", n_frames=46, fmt="u8", snr=40.0, cfo=123.4, seed=1, tii=[(12, 5, 1.0, 0)], tii_mode=1, service=0),
    # near the decoding threshold: RS has work to do
    "c1_u8_4p5db": dict(kind="config1", n_frames=30, fmt="u8", snr=4.5, cfo=-3456.7, seed=5, tii=[], tii_mode=None, service=0),
    # s16 samples, large negative CFO
    "c1_s16_12db": dict(kind="config1", n_frames=24, fmt="s16", snr=12.0, cfo=-8765.4, seed=9, tii=[], tii_mode=None, service=0),
    # three overlapping transmitters with different amplitude and delay
    "tii3_u8_25db": dict(kind="config1", n_frames=46, fmt="u8", snr=25.0, cfo=50.0, seed=3,
                         tii=[(12, 5, 1.0, 0), (33, 17, 0.5, 30), (60, 2, 0.25, 100)], tii_mode=1, service=None),
    # several sub-channels with different protection (EEP-A/B, UEP); the reference decodes the selected one
    "multi_u8_25db": dict(kind="multi", n_frames=22, fmt="u8", snr=25.0, cfo=-2345.6, seed=21, tii=[(3, 9, 1.0, 0)], tii_mode=None, service=2),
    # same ensemble, the DAB (MPEG-1 Layer II, UEP table index 35) service selected: frames with a valid ISO 11172-3 CRC
    "multi_mp2_u8_25db": dict(kind="multi", n_frames=22, fmt="u8", snr=25.0, cfo=-2345.6, seed=21, tii=[(3, 9, 1.0, 0)], tii_mode=None, service=3),
    # programme associated data: dynamic label segments (X-PAD applications 2/3) inside the data_stream_element of every AU
    "c1_dls_u8_30db": dict(kind="config1", n_frames=40, fmt="u8", snr=30.0, cfo=100.0, seed=3, tii=[], tii_mode=None, service=0,
                           dls=["Hello B200 - now playing: Test Tone One", "Second label", "x"]),
    # MOT-style MSC data groups in the X-PAD (length indicator + applications 12/13), announced by FIG 0/13 as SlideShow (type 2)
    "c1_mot_u8_30db": dict(kind="config1", n_frames=40, fmt="u8", snr=30.0, cfo=100.0, seed=4, tii=[], tii_mode=None, service=0,
                           dls=["MOT test"], xpad_dg=3, xpad_apps=True),
    # DAB (MPEG Layer II) service with a dynamic label in the PAD at the end of every audio frame
    "mp2_dls_u8_30db": dict(kind="mp2pad", n_frames=30, fmt="u8", snr=30.0, cfo=50.0, seed=2, tii=[], tii_mode=None, service=0,
                            dls=["MP2 dynamic label text, long enough for two segments", "short"]),
    # secondary packet-mode component (TMId 3, FIG 0/2 + 0/3 + 0/8 + 0/13) next to the audio: MSC data groups on the data decoder
    "c1_packet_u8_30db": dict(kind="config1", n_frames=40, fmt="u8", snr=30.0, cfo=100.0, seed=5, tii=[], tii_mode=None, service=0,
                              packet=dict(subchid=7, start_cu=80, scid=0x123, scids=1, addr=5, other_addr=9, n_dg=4), data_scids=1),
    # service information the core reports without being asked: FIG 0/13 (user application update), 0/17 (programme type),
    # 0/18 (announcement support), 0/19 (announcement switching: flags raised for CIFs 60..100)
    "c1_ann_u8_30db": dict(kind="config1", n_frames=44, fmt="u8", snr=30.0, cfo=100.0, seed=6, tii=[], tii_mode=None, service=0,
                           pty=10, asu=[0x0003, [5]], announcements=[[60, 100, 5, 0x0002, 3]], user_apps=True,
                           utc=[60600, 23, 59, 57, 100]),
    # signal-spectrum callback (dabsdr.h:365-370, :428) with all 2048 bins of the first callbacks: 2 per frame, SIGNAL and NULL / TII
    "c1_spectrum_u8_30db": dict(kind="config1", n_frames=14, fmt="u8", snr=30.0, cfo=0.0, seed=8, tii=[(12, 5, 1.0, 0)], tii_mode=1, service=None,
                                spectrum_bins=14),
    # multiplex reconfiguration: from CIF 140 on a second sub-channel and service exist; announced 80 CIFs ahead through the FIG 0/0
    # change flags + occurrence change and the next configuration sent with C/N = 1 (EN 300 401 sec. 6.4)
    "c1_reconf_u8_30db": dict(kind="config1", n_frames=70, fmt="u8", snr=30.0, cfo=100.0, seed=31, tii=[], tii_mode=None, service=0,
                              reconfig=dict(at=140, lead=80, subchid=9, start_cu=200, bitrate=64, sid=0x1234, label="ADDED SERVICE")),
    # another ensemble takes over the frequency after 24 frames (same sub-channel organisation): NID_RESET(NEW_EID)
    "two_ens_u8_30db": dict(kind="two", n_frames=24, n_frames_b=26, fmt="u8", snr=30.0, cfo=100.0, seed=7, tii=[], tii_mode=None, service=0),
}


def mot_data_groups(n, seed=9):
    """n MSC data groups with CRC: header (CRC + segment flags, type 4, continuity k), segment number k, 40 random bytes"""
    import numpy as np
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        body = bytes([0x64, (k << 4) & 0xF0, 0x80, k & 0xFF, 0x00, 40]) + rng.integers(0, 256, 40, dtype=np.uint8).tobytes()
        crc = int(synth.crc16_np(np.frombuffer(body, np.uint8)[None, :])[0])
        out.append(body + bytes([crc >> 8, crc & 0xFF]))
    return out


def build_case(spec):
    if spec["kind"] == "config1":
        ens = synth.config1_ensemble(seed=1000 + spec["seed"])
        ens.tii = [synth.TiiTx(m, s, a, delay=d) for m, s, a, d in spec["tii"]]
        if spec.get("dls"):
            ens.subchannels[0].dls = list(spec["dls"])
        if spec.get("pty") is not None:
            ens.services[0].pty = spec["pty"]
        if spec.get("asu"):
            ens.services[0].asu = (spec["asu"][0], list(spec["asu"][1]))
        if spec.get("announcements"):
            ens.announcements = [tuple(a) for a in spec["announcements"]]
        if spec.get("utc"):
            ens.utc = tuple(spec["utc"])
        if spec.get("user_apps"):
            ens.services[0].user_apps = [(2, bytes([0x0C, 0x3C]))]
        if spec.get("packet"):
            pk = spec["packet"]
            dgs = mot_data_groups(pk["n_dg"], seed=11)
            ens.subchannels.append(synth.SubChannel(pk["subchid"], pk["start_cu"], "packet", 32, ("EEP", 3, "A"), seed=1,
                                                    packets=[(pk["addr"], dgs), (pk["other_addr"], dgs[:2])]))
            ens.services[0].data_comps = [synth.DataComponent(pk["subchid"], pk["scid"], pk["scids"], pk["addr"], user_apps=[(2, b"")])]
        if spec.get("reconfig"):
            rc = spec["reconfig"]
            ens.reconfig = dict(at=rc["at"], lead=rc["lead"],
                                subchannels=[synth.SubChannel(rc["subchid"], rc["start_cu"], "dabplus", rc["bitrate"], ("EEP", 3, "A"), seed=3109)],
                                services=[synth.Service(rc["sid"], rc["label"], rc["subchid"])])
        if spec.get("xpad_dg"):
            ens.subchannels[0].xpad_dg = mot_data_groups(spec["xpad_dg"])
            ens.services[0].user_apps = [(2, bytes([0x0C, 0x3C]))]       # SlideShow over X-PAD application 12, DSCTy 60 (MOT)
    elif spec["kind"] == "two":
        import numpy as np
        a = synth.config1_ensemble(seed=1000 + spec["seed"])
        b = synth.config1_ensemble(seed=2000 + spec["seed"])
        b.eid, b.label, b.cif_count0 = 0x2222, "SECOND MUX", 100
        b.services[0].sid, b.services[0].label = 0x2223, "OTHER SVC"
        ca, ta = synth.modulate(a, spec["n_frames"], seed=spec["seed"])
        cb, tb = synth.modulate(b, spec["n_frames_b"], seed=spec["seed"] + 1)
        imp = synth.Impairments(snr_db=spec["snr"], cfo_hz=spec["cfo"], seed=spec["seed"])
        raw = synth.quantize(synth.impair(np.concatenate([ca, cb]), imp, 150000), spec["fmt"], imp.rms)
        res = synth.SynthResult(iq=None, n_frames=spec["n_frames"] + spec["n_frames_b"], fibs=np.concatenate([ta["fibs"], tb["fibs"]]),
                                logical=ta["logical"], aus={3: ta["aus"][3] + tb["aus"][3]}, superframes=ta["superframes"], ens=a,
                                frame0=imp.lead_in, cfo_hz=imp.cfo_hz)
        return raw, res
    elif spec["kind"] == "mp2pad":
        subs = [synth.SubChannel(9, 0, "mp2", 128, ("UEP", 35), seed=spec["seed"] + 3, dls=list(spec["dls"]))]
        ens = synth.EnsembleConfig(eid=0x3003, label="B200 MP2PAD", subchannels=subs, services=[synth.Service(0x3003, "MP2 PAD", 9)])
    else:
        seed = spec["seed"]
        subs = [synth.SubChannel(1, 0, "dabplus", 96, ("EEP", 3, "A"), seed=seed),
                synth.SubChannel(2, 72, "dabplus", 64, ("EEP", 2, "A"), seed=seed + 1),
                synth.SubChannel(5, 140, "dabplus", 64, ("EEP", 3, "B"), seed=seed + 2),
                synth.SubChannel(9, 200, "mp2", 128, ("UEP", 35), seed=seed + 3),
                synth.SubChannel(11, 300, "dabplus", 32, ("EEP", 1, "A"), seed=seed + 4)]
        svcs = [synth.Service(0x3000 + i, f"MULTI {i}", s.subchid) for i, s in enumerate(subs)]
        ens = synth.EnsembleConfig(eid=0x3003, label="B200 MULTI", subchannels=subs, services=svcs,
                                   tii=[synth.TiiTx(m, s, a, delay=d) for m, s, a, d in spec["tii"]])
    imp = synth.Impairments(snr_db=spec["snr"], cfo_hz=spec["cfo"], seed=spec["seed"])
    raw, res = synth.synthesize(ens, spec["n_frames"], spec["fmt"], imp, seed=spec["seed"], tail=150000)
    return raw, res


def load_golden(name):
    with gzip.open(GOLDEN_DIR / f"{name}.json.gz", "rt") as f:
        return json.load(f)


def raw_matches(raw, golden) -> bool:
    return hashlib.sha1(raw.tobytes()).hexdigest() == golden["raw_sha1"]
