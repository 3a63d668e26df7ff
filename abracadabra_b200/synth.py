"""Synthetic DAB mode I ensemble modulator (transmitter side) with known payloads.

Test/benchmark tooling: builds the IQ files the receive path is exercised on (SURVEY.md sec. 7 step 2,
sec. 8(d) configs 1-5).  Everything here follows ETSI EN 300 401 (mode I) and TS 102 563 (DAB+) with the
constants of SURVEY.md Appendix A, each of which libdabsdr 4.0.1 has been shown to decode.

The modulator returns the ground truth alongside the samples: FIBs, per-sub-channel logical frames
(the bytes a receiver obtains after Viterbi + energy dispersal), DAB+ access units, TII ids.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import dabconst as C


# ============================================================================ configuration
@dataclass
class SubChannel:
    subchid: int
    start_cu: int
    kind: str                     # 'dabplus' | 'mp2' | 'data' (random bytes) | 'packet' (packet mode, see packet_frames)
    bitrate: int                  # kbit/s
    prot: Tuple                   # ('EEP', level 1..4, 'A'|'B')  or ('UEP', table index 0..63)
    seed: int = 0
    dls: Optional[List[str]] = None       # dynamic label messages carried in the X-PAD of a DAB+ sub-channel (cycled)
    xpad_dg: Optional[List[bytes]] = None  # MSC data groups (incl. CRC) carried as X-PAD application 12/13 (MOT), cycled
    packets: Optional[List[Tuple[int, List[bytes]]]] = None   # packet mode: [(packet address, [MSC data groups, cycled])]

    def profile(self):
        if self.prot[0] == "EEP":
            segs, size = C.eep_profile(self.bitrate, self.prot[1], self.prot[2])
            return segs, size, 0
        segs, size, pad, br = C.uep_profile(self.prot[1])
        assert br == self.bitrate, "UEP table index does not match bitrate"
        return segs, size, pad

    @property
    def size_cu(self) -> int:
        return self.profile()[1]

    @property
    def frame_bytes(self) -> int:   # logical frame (24 ms) size
        return self.bitrate * 3


@dataclass
class DataComponent:
    """Secondary service component in packet mode (TMId 3): FIG 0/2 + 0/3 + 0/8 + 0/13"""
    subchid: int
    scid: int                     # 12-bit service component identifier
    scids: int                    # identifier within the service (>= 1)
    packet_addr: int              # 10 bit
    dscty: int = 60               # 60 = MOT
    dg_flag: int = 0              # 0 = MSC data groups are used
    user_apps: Optional[List[Tuple[int, bytes]]] = None


@dataclass
class Service:
    sid: int
    label: str
    subchid: int
    user_apps: Optional[List[Tuple[int, bytes]]] = None   # FIG 0/13: (user application type, user application data)
    data_comps: Optional[List[DataComponent]] = None
    pty: Optional[int] = None                              # FIG 0/17: programme type (international code)
    asu: Optional[Tuple[int, List[int]]] = None            # FIG 0/18: (announcement support flags, cluster ids)


@dataclass
class TiiTx:
    main: int
    sub: int
    amp: float = 1.0
    delay: int = 0                # samples
    phase: float = 0.0


@dataclass
class EnsembleConfig:
    eid: int = 0x1001
    ecc: int = 0xE1
    label: str = "B200 TEST MUX"
    lto: int = 2
    subchannels: List[SubChannel] = field(default_factory=list)
    services: List[Service] = field(default_factory=list)
    tii: List[TiiTx] = field(default_factory=list)
    tii_every_null: bool = False  # debugging: put TII in every null instead of every second one
    # FIG 0/19 announcement switching: (first CIF, last CIF, cluster id, ASw flags, sub-channel id); flags 0 outside the interval
    announcements: List[Tuple[int, int, int, int, int]] = field(default_factory=list)
    utc: Optional[Tuple[int, int, int, int, int]] = None   # FIG 0/10 (long form): (MJD, hours, minutes, seconds, ms) at the first CIF
    cif_count0: int = 0           # CIF counter of the first CIF (multiple of 4 keeps frames aligned)
    # multiplex reconfiguration (EN 300 401 sec. 6.4): dict(at=CIF (relative to the first) from which the new configuration holds,
    # lead=CIFs of advance notice, subchannels=[SubChannel added], services=[Service added]).  From at - lead on FIG 0/0 carries the
    # change flags and the occurrence change, and the new configuration is sent with C/N = 1 next to the current one.
    reconfig: Optional[dict] = None

    def check(self):
        used = np.zeros(864, dtype=int)
        for s in self.subchannels:
            used[s.start_cu:s.start_cu + s.size_cu] += 1
            assert s.start_cu + s.size_cu <= 864, "sub-channel exceeds CIF"
        assert used.max() <= 1, "sub-channels overlap"


@dataclass
class Impairments:
    snr_db: Optional[float] = 40.0     # None = no noise
    cfo_hz: float = 0.0
    cfo_drift_hz_per_s: float = 0.0
    taps: List[Tuple[int, complex]] = field(default_factory=list)   # extra multipath taps (delay samples, gain)
    lead_in: int = 49152               # samples of noise before the first frame (24 ms)
    rms: float = 28.0                  # target RMS in output LSB (u8: 28, s16: ~7000)
    sample_offset_ppm: float = 0.0     # receiver sample clock error: the signal is resampled by (1 + ppm * 1e-6)
    seed: int = 1


def config1_ensemble(seed: int = 1000) -> EnsembleConfig:
    """BASELINE.json configs[0]: FIC + one DAB+ 96 kbps EEP-3A sub-channel, TII (12,5)."""
    return EnsembleConfig(
        subchannels=[SubChannel(3, 0, "dabplus", 96, ("EEP", 3, "A"), seed=seed)],
        services=[Service(0x1001, "SERVICE ONE", 3)],
        tii=[TiiTx(12, 5, 1.0)],
    )


def full_ensemble(seed: int = 2000) -> EnsembleConfig:
    """BASELINE.json configs[1]: FIC + 12 sub-channels filling 864 CU (mixed DAB+ EEP and MP2 UEP)."""
    plan = [  # kind, bitrate, prot
        ("dabplus", 96, ("EEP", 3, "A")),    # 72
        ("dabplus", 96, ("EEP", 3, "A")),    # 72
        ("dabplus", 80, ("EEP", 3, "A")),    # 60
        ("dabplus", 72, ("EEP", 3, "A")),    # 54
        ("dabplus", 64, ("EEP", 3, "A")),    # 48
        ("dabplus", 48, ("EEP", 3, "A")),    # 36
        ("dabplus", 64, ("EEP", 2, "A")),    # 64
        ("dabplus", 64, ("EEP", 3, "B")),    # 36
        ("mp2", 128, ("UEP", 35)),           # 96
        ("mp2", 160, ("UEP", 40)),           # 116
        ("mp2", 192, ("UEP", 44)),           # 116
        ("mp2", 128, ("UEP", 34)),           # 84
    ]
    subs, svcs, cu = [], [], 0
    for i, (kind, br, prot) in enumerate(plan):
        sc = SubChannel(i + 1, cu, kind, br, prot, seed=seed + i)
        subs.append(sc)
        svcs.append(Service(0x2000 + i + 1, f"SVC {i + 1:02d} {kind.upper()}", i + 1))
        cu += sc.size_cu
    assert cu <= 864, cu
    ens = EnsembleConfig(eid=0x2002, label="B200 FULL MUX", subchannels=subs, services=svcs,
                         tii=[TiiTx(20, 7, 1.0), TiiTx(20, 11, 0.5, delay=40)])
    ens.check()
    return ens


# ============================================================================ coding primitives
def bytes_to_bits(b: np.ndarray) -> np.ndarray:
    return np.unpackbits(np.asarray(b, dtype=np.uint8), axis=-1)


def bits_to_bytes(bits: np.ndarray) -> np.ndarray:
    return np.packbits(np.asarray(bits, dtype=np.uint8), axis=-1)


def conv_encode(bits: np.ndarray) -> np.ndarray:
    """Rate 1/4 K=7 mother code, zero-terminated.  bits [R, n] -> [R, 4*(n+6)]."""
    bits = np.atleast_2d(bits).astype(np.uint8)
    r, n = bits.shape
    a = np.zeros((r, n + 12), dtype=np.uint8)      # 6 zeros history + data + 6 tail zeros
    a[:, 6:6 + n] = bits
    out = np.zeros((r, n + 6, 4), dtype=np.uint8)
    for j, poly in enumerate(C.POLYS):
        acc = np.zeros((r, n + 6), dtype=np.uint8)
        for d in range(7):                           # tap on a[i-d] is bit (6-d) of the polynomial
            if (poly >> (6 - d)) & 1:
                acc ^= a[:, 6 - d:6 - d + n + 6]
        out[:, :, j] = acc
    return out.reshape(r, 4 * (n + 6))


def encode_protected(frames: np.ndarray, segs, pad: int) -> np.ndarray:
    """Energy-dispersed bits [R, n] -> punctured code bits [R, size]."""
    n = frames.shape[1]
    keep = C.punct_keep_mask(segs, n)
    enc = conv_encode(frames)[:, keep]
    if pad:
        enc = np.concatenate([enc, np.zeros((enc.shape[0], pad), dtype=np.uint8)], axis=1)
    return enc


# GF(256) for RS(120,110) -- poly 0x11D, generator roots alpha^0..alpha^9
def _gf_tables():
    exp = np.zeros(512, dtype=np.int64)
    log = np.zeros(256, dtype=np.int64)
    x = 1
    for i in range(255):
        exp[i] = x
        log[x] = i
        x <<= 1
        if x & 0x100:
            x ^= 0x11D
    exp[255:510] = exp[:255]
    return exp, log


GF_EXP, GF_LOG = _gf_tables()


def _rs_generator():
    g = [1]
    for i in range(10):
        ng = [0] * (len(g) + 1)
        for j, c in enumerate(g):           # multiply by (x + alpha^i); g stored highest degree first
            ng[j] ^= c
            if c:
                ng[j + 1] ^= int(GF_EXP[GF_LOG[c] + i])
        g = ng
    return g                                  # length 11, g[0] = 1


RS_GEN = _rs_generator()


def rs_parity(data: np.ndarray) -> np.ndarray:
    """Systematic RS(120,110) parity; data [R, 110] uint8 -> [R, 10] (shortening zeros are implicit)."""
    data = np.atleast_2d(data).astype(np.int64)
    r = data.shape[0]
    rem = np.zeros((r, 10), dtype=np.int64)
    glog = np.array([GF_LOG[c] if c else -1 for c in RS_GEN[1:]], dtype=np.int64)
    for i in range(data.shape[1]):
        fb = data[:, i] ^ rem[:, 0]
        rem[:, :-1] = rem[:, 1:]
        rem[:, -1] = 0
        nz = fb != 0
        if nz.any():
            lf = GF_LOG[fb[nz]]
            for j in range(10):
                if glog[j] >= 0:
                    rem[nz, j] ^= GF_EXP[lf + glog[j]]
    return rem.astype(np.uint8)


def crc16_np(rows: np.ndarray, poly: int = 0x1021, init: int = 0xFFFF, invert: bool = True) -> np.ndarray:
    """CRC-16 of each row of a uint8 matrix (default: poly 0x1021, init 0xFFFF, inverted)."""
    rows = np.atleast_2d(rows).astype(np.uint32)
    tbl = _crc_table(poly)
    crc = np.full(rows.shape[0], init, dtype=np.uint32)
    for i in range(rows.shape[1]):
        crc = ((crc << 8) & 0xFFFF) ^ tbl[((crc >> 8) ^ rows[:, i]) & 0xFF]
    if invert:
        crc = crc ^ 0xFFFF
    return crc.astype(np.uint16)


_CRC_TABLES: Dict[int, np.ndarray] = {}


def _crc_table(poly: int) -> np.ndarray:
    if poly not in _CRC_TABLES:
        t = np.zeros(256, dtype=np.uint32)
        for i in range(256):
            c = i << 8
            for _ in range(8):
                c = ((c << 1) ^ poly) & 0xFFFF if c & 0x8000 else (c << 1) & 0xFFFF
            t[i] = c
        _CRC_TABLES[poly] = t
    return _CRC_TABLES[poly]


# ============================================================================ payload builders
# ---------------------------------------------------------------------------- PAD (EN 300 401 sec. 7.4, TS 102 563 sec. 5.4)
XPAD_LEN = (4, 6, 8, 12, 16, 24, 32, 48)


def dl_segments(text: str, toggle: int, charset: int = 0) -> List[bytes]:
    """Dynamic-label X-PAD data groups of one message: 2-byte prefix + up to 16 characters + CRC-16."""
    raw = text.encode("latin1")[:128] or b" "
    chunks = [raw[i:i + 16] for i in range(0, len(raw), 16)]
    out = []
    for k, ch in enumerate(chunks):
        first, last = k == 0, k == len(chunks) - 1
        b0 = (toggle << 7) | (first << 6) | (last << 5) | (len(ch) - 1)
        b1 = (charset << 4) if first else (k << 4)
        body = bytes([b0, b1]) + ch
        crc = int(crc16_np(np.frombuffer(body, np.uint8)[None, :])[0])
        out.append(body + bytes([crc >> 8, crc & 0xFF]))
    return out


class XPadStream:
    """Variable-size X-PAD, one application sub-field per AU: dynamic label (application types 2 / 3) and, when given, MSC
    data groups announced by a data group length indicator (application type 1) and carried as types 12 / 13."""

    def __init__(self, dls: Optional[List[str]], dgs: Optional[List[bytes]], dl_len_idx: int = 3, dg_len_idx: int = 6):
        self.q: List[Tuple[int, bytes]] = []          # (application type of the first sub-field, data group bytes)
        self.dls, self.dgs = dls or [], dgs or []
        self.n_dl = self.n_dg = 0
        self.dl_len, self.dg_len = dl_len_idx, dg_len_idx
        self.cur: Optional[Tuple[int, bytes, int]] = None
        self.sent: List[Tuple[str, bytes]] = []       # what a receiver should report, in order

    def _refill(self):
        if self.dls:
            for seg in dl_segments(self.dls[self.n_dl % len(self.dls)], (self.n_dl // len(self.dls)) & 1):
                self.q.append((2, seg))
            self.n_dl += 1
        if self.dgs:
            dg = self.dgs[self.n_dg % len(self.dgs)]
            self.n_dg += 1
            ind = bytes([(len(dg) >> 8) & 0x3F, len(dg) & 0xFF])
            crc = int(crc16_np(np.frombuffer(ind, np.uint8)[None, :])[0])
            self.q.append((1, ind + bytes([crc >> 8, crc & 0xFF])))
            self.q.append((12, dg))

    def next_pad(self) -> bytes:
        """PAD bytes of one AU (X-PAD in transmission order, i.e. reversed, followed by the two F-PAD bytes)."""
        if self.cur is None:
            if not self.q:
                self._refill()
            if not self.q:
                return bytes([0x00, 0x00])            # F-PAD only, no X-PAD
            app, data = self.q.pop(0)
            self.cur = (app, data, 0)
            if app != 1:
                self.sent.append(("dl" if app == 2 else "dg", data))
        app, data, pos = self.cur
        if app == 1:
            n, ci = 4, (0 << 5) | 1                    # length indicator: 4-byte sub-field
        else:
            idx = self.dl_len if app == 2 else self.dg_len
            n = XPAD_LEN[idx]
            ci = (idx << 5) | (app if pos == 0 else app + 1)
        piece = data[pos:pos + n].ljust(n, b"\x00")
        pos += n
        self.cur = None if pos >= len(data) else (app, data, pos)
        xpad = bytes([ci, 0x00]) + piece               # CI list: one CI + end marker
        return xpad[::-1] + bytes([0x20, 0x02])        # F-PAD: variable-size X-PAD, CI flag set


def dabplus_superframes(sc: SubChannel, n_sf: int, rng: np.random.Generator):
    """Build n_sf DAB+ audio superframes for sub-channel `sc` (vectorised over superframes).

    Returns (bytes [n_sf, 120*s], list of access units (bytes, without CRC) in stream order).
    HE-AAC v1 24 kHz core: dac_rate=1, sbr=1, stereo -> 3 AUs per superframe, header byte 0x70.
    """
    s = sc.bitrate // 8
    size = 110 * s
    num_aus, first = 3, 6
    body = size - first
    lens = [body // num_aus] * num_aus
    lens[-1] += body - sum(lens)
    starts = [first]
    for ln in lens[:-1]:
        starts.append(starts[-1] + ln)
    buf = np.zeros((n_sf, size), dtype=np.uint8)
    buf[:, 2] = 0x70
    buf[:, 3] = (starts[1] >> 4) & 0xFF
    buf[:, 4] = ((starts[1] & 0xF) << 4) | ((starts[2] >> 8) & 0xF)
    buf[:, 5] = starts[2] & 0xFF
    payloads = []
    xp = XPadStream(sc.dls, sc.xpad_dg) if (sc.dls or sc.xpad_dg) else None
    pls = [rng.integers(0, 256, (n_sf, lens[a] - 2), dtype=np.uint8) for a in range(num_aus)]
    if xp is not None:
        # every AU starts with a data_stream_element (id 4, instance tag 0, no byte alignment flag) that carries the PAD
        for f in range(n_sf):
            for a in range(num_aus):
                pad = xp.next_pad()
                dse = bytes([4 << 5, len(pad)]) + pad
                pls[a][f, :len(dse)] = np.frombuffer(dse, np.uint8)
                pls[a][f, len(dse)] = 0xE0 | (pls[a][f, len(dse)] & 0x1F)     # next element: id 7 = END
        sc._pad_sent = xp.sent
    for a in range(num_aus):
        pl = pls[a]
        crc = crc16_np(pl)
        buf[:, starts[a]:starts[a] + lens[a] - 2] = pl
        buf[:, starts[a] + lens[a] - 2] = crc >> 8
        buf[:, starts[a] + lens[a] - 1] = crc & 0xFF
        payloads.append(pl)
    fc = crc16_np(buf[:, 2:11], poly=0x782F, init=0, invert=False)      # fire code over bytes 2..10
    buf[:, 0] = fc >> 8
    buf[:, 1] = fc & 0xFF
    aus: List[bytes] = []
    for f in range(n_sf):
        for a in range(num_aus):
            aus.append(payloads[a][f].tobytes())
    # RS: codeword j of a superframe = bytes j + s*i
    data = buf.reshape(n_sf, 110, s).transpose(0, 2, 1).reshape(n_sf * s, 110)
    par = rs_parity(data).reshape(n_sf, s, 10).transpose(0, 2, 1).reshape(n_sf, 10 * s)
    sf = np.concatenate([buf, par], axis=1)
    return sf, aus


def _mp2_crc(crc: int, value: int, bits: int) -> int:
    for i in range(bits - 1, -1, -1):
        b = ((value >> i) & 1) ^ (crc >> 15)
        crc = (crc << 1) & 0xFFFF
        if b:
            crc ^= 0x8005
    return crc


def mp2_frames(sc: SubChannel, n: int, rng: np.random.Generator, valid_crc: bool = True) -> np.ndarray:
    """MPEG-1 Layer II, 48 kHz, stereo frames (24 ms) of the sub-channel's bitrate: valid header, random bit allocation and
    scfsi protected by a correct ISO 11172-3 CRC-16 (poly 0x8005), random remainder (no decodable audio, DAB ScF-CRC not set)."""
    fb = sc.frame_bytes
    frames = rng.integers(0, 256, (n, fb), dtype=np.uint8)
    br_idx = {32: 1, 48: 2, 56: 3, 64: 4, 80: 5, 96: 6, 112: 7, 128: 8, 160: 9, 192: 10, 224: 11, 256: 12,
              320: 13, 384: 14}[sc.bitrate]
    frames[:, 0] = 0xFF
    frames[:, 1] = 0xFC            # MPEG-1, layer II, CRC protected
    frames[:, 2] = (br_idx << 4) | (1 << 2)   # 48 kHz
    frames[:, 3] = 0x00            # stereo
    if not valid_crc:
        return frames
    br_ch = sc.bitrate // 2
    sbl = 8 if br_ch <= 48 else 27
    nbal = (lambda sb: 4 if sb < 2 else 3) if br_ch <= 48 else (lambda sb: 4 if sb < 11 else 3 if sb < 23 else 2)
    for f in range(n):
        bits = np.unpackbits(frames[f])
        crc = _mp2_crc(_mp2_crc(0xFFFF, int(frames[f, 2]), 8), int(frames[f, 3]), 8)
        pos = 48
        nz = 0
        for sb in range(sbl):
            nb = nbal(sb)
            for _ch in range(2):
                v = int("".join(map(str, bits[pos:pos + nb])), 2)
                crc = _mp2_crc(crc, v, nb)
                nz += v != 0
                pos += nb
        for _ in range(nz):
            v = int(bits[pos]) * 2 + int(bits[pos + 1])
            crc = _mp2_crc(crc, v, 2)
            pos += 2
        frames[f, 4], frames[f, 5] = crc >> 8, crc & 0xFF
    if sc.dls or sc.xpad_dg:
        # PAD at the end of the frame: [X-PAD, reversed] [ScF-CRC: 4 bytes at >= 56 kbit/s per channel, else 2] [F-PAD]
        xp = XPadStream(sc.dls, sc.xpad_dg)
        scf = 4 if sc.bitrate // 2 >= 56 else 2
        for f in range(n):
            pad = xp.next_pad()
            tail = pad[:-2] + bytes(scf) + pad[-2:]
            frames[f, fb - len(tail):] = np.frombuffer(tail, np.uint8)
        sc._pad_sent = xp.sent
    return frames


# ---------------------------------------------------------------------------- packet mode (EN 300 401 sec. 5.3.2, 5.3.3)
PACKET_LEN = (24, 48, 72, 96)


def make_packet(plen_code: int, ci: int, first: bool, last: bool, addr: int, data: bytes, command: int = 0) -> bytes:
    n = PACKET_LEN[plen_code]
    assert len(data) <= n - 5
    hdr = bytes([(plen_code << 6) | ((ci & 3) << 4) | (first << 3) | (last << 2) | ((addr >> 8) & 3), addr & 0xFF,
                 (command << 7) | len(data)])
    body = hdr + data.ljust(n - 5, b"\x00")
    crc = int(crc16_np(np.frombuffer(body, np.uint8)[None, :])[0])
    return body + bytes([crc >> 8, crc & 0xFF])


def packet_frames(sc: SubChannel, n: int, rng: np.random.Generator) -> np.ndarray:
    """Logical frames of a packet-mode sub-channel: the data groups of every address are cut into packets of all four
    lengths (cycling through a few frame layouts), interleaved address by address, with padding packets (address 0)."""
    fb = sc.frame_bytes
    layouts = [[0, 0, 1], [3], [2, 0], [1, 1], [0, 0, 0, 0]]            # packet length codes filling 96 bytes
    assert fb % 96 == 0
    state = []                                                         # per address: [addr, groups, next group, offset, ci]
    for addr, groups in sc.packets or []:
        state.append([addr, groups, 0, 0, 0])
    sent = []
    frames = np.zeros((n, fb), np.uint8)
    turn = 0
    for f in range(n):
        out = b""
        while len(out) < fb:
            for code in layouts[(f + len(out) // 96) % len(layouts)]:
                plen = PACKET_LEN[code]
                if not state or (turn % 5) == 4:                       # padding packet
                    out += make_packet(code, 0, True, True, 0, b"")
                    turn += 1
                    continue
                st = state[turn % len(state)] if len(state) > 1 else state[0]
                turn += 1
                addr, groups, gi, off, ci = st
                dg = groups[gi % len(groups)]
                piece = dg[off:off + plen - 5]
                first, last = off == 0, off + len(piece) >= len(dg)
                out += make_packet(code, ci, first, last, addr, piece)
                st[4] = (ci + 1) & 3
                if last:
                    sent.append((addr, dg))
                    st[2], st[3] = gi + 1, 0
                else:
                    st[3] = off + len(piece)
        frames[f] = np.frombuffer(out[:fb], np.uint8)
    sc._packets_sent = sent
    return frames


# ============================================================================ FIC
def _fig(type_: int, body: bytes) -> bytes:
    assert len(body) <= 29
    return bytes([(type_ << 5) | len(body)]) + body


def fig0_0(ens: EnsembleConfig, cif_count: int, change: int = 0, occurrence: int = 0) -> bytes:
    hi, lo = (cif_count // 250) % 20, cif_count % 250
    body = [0x00, ens.eid >> 8, ens.eid & 0xFF, ((change & 3) << 6) | (hi & 0x1F), lo]
    if change:
        body.append(occurrence & 0xFF)
    return _fig(0, bytes(body))


def fig0_1_entry(sc: SubChannel) -> bytes:
    b = bytes([(sc.subchid << 2) | (sc.start_cu >> 8), sc.start_cu & 0xFF])
    if sc.prot[0] == "UEP":
        return b + bytes([sc.prot[1] & 0x3F])
    opt = 0 if sc.prot[2] == "A" else 1
    lvl = sc.prot[1] - 1
    size = sc.size_cu
    return b + bytes([0x80 | (opt << 4) | (lvl << 2) | (size >> 8), size & 0xFF])


def fig0_2_entry(ens: EnsembleConfig, sv: Service) -> bytes:
    sc = next(s for s in list(ens.subchannels) + list((ens.reconfig or {}).get("subchannels", [])) if s.subchid == sv.subchid)
    ascty = 63 if sc.kind == "dabplus" else 0
    comps = bytes([(0 << 6) | ascty, (sc.subchid << 2) | 0x02])
    for dc in sv.data_comps or []:
        comps += bytes([(3 << 6) | (dc.scid >> 6), ((dc.scid & 0x3F) << 2)])          # TMId 3, secondary, no CA
    return bytes([sv.sid >> 8, sv.sid & 0xFF, 1 + len(sv.data_comps or [])]) + comps


def fig0_3_entry(dc: DataComponent) -> bytes:
    return bytes([dc.scid >> 4, (dc.scid & 0xF) << 4, (dc.dg_flag << 7) | dc.dscty, (dc.subchid << 2) | (dc.packet_addr >> 8),
                  dc.packet_addr & 0xFF])


def fig0_8_entry(sv: Service, dc: DataComponent) -> bytes:
    return bytes([sv.sid >> 8, sv.sid & 0xFF, dc.scids & 0x0F, 0x80 | (dc.scid >> 8), dc.scid & 0xFF])


def fig0_13_entry_data(sv: Service, dc: DataComponent) -> bytes:
    b = bytes([sv.sid >> 8, sv.sid & 0xFF, (dc.scids << 4) | len(dc.user_apps)])
    for ua_type, data in dc.user_apps:
        b += bytes([(ua_type >> 3) & 0xFF, ((ua_type & 7) << 5) | len(data)]) + data
    return b


def fig0_13_entry(sv: Service) -> bytes:
    b = bytes([sv.sid >> 8, sv.sid & 0xFF, (0 << 4) | len(sv.user_apps)])
    for ua_type, data in sv.user_apps:
        b += bytes([(ua_type >> 3) & 0xFF, ((ua_type & 7) << 5) | len(data)]) + data
    return b


def fig0_17_entry(sv: Service) -> bytes:
    return bytes([sv.sid >> 8, sv.sid & 0xFF, 0x00, sv.pty & 0x1F])          # static PTy, no language, no complementary code


def fig0_18_entry(sv: Service) -> bytes:
    asu, clusters = sv.asu
    return bytes([sv.sid >> 8, sv.sid & 0xFF, asu >> 8, asu & 0xFF, len(clusters)]) + bytes(clusters)


def fig0_19_entry(cluster: int, asw: int, subchid: int, new: int = 0) -> bytes:
    return bytes([cluster, asw >> 8, asw & 0xFF, (new << 7) | (0 << 6) | (subchid & 0x3F)])


def fig0_10(ens: EnsembleConfig, cif_rel: int) -> bytes:
    mjd, h, m, sec, ms = ens.utc
    t = ((h * 60 + m) * 60 + sec) * 1000 + ms + 24 * cif_rel                    # a CIF lasts 24 ms
    mjd += t // 86400000
    t %= 86400000
    h, m, sec, ms = t // 3600000, (t // 60000) % 60, (t // 1000) % 60, t % 1000
    v = (0 << 31) | (mjd << 14) | (0 << 13) | (0 << 12) | (1 << 11) | (h << 6) | m           # Rfu, MJD, LSI, Rfa, UTC flag = 1, hours, minutes
    return _fig(0, bytes([0x0A, v >> 24, (v >> 16) & 0xFF, (v >> 8) & 0xFF, v & 0xFF, (sec << 2) | (ms >> 8), ms & 0xFF]))


def fig0_9(ens: EnsembleConfig) -> bytes:
    return _fig(0, bytes([0x09, ens.lto & 0x3F, ens.ecc, 0x01]))


def fig1(ext: int, ident: int, label: str) -> bytes:
    txt = label.encode("ascii")[:16].ljust(16, b" ")
    return _fig(1, bytes([ext, ident >> 8, ident & 0xFF]) + txt + bytes([0xFF, 0x00]))


def _pack_rotating(entries: List[bytes], ext_byte: int, room: int, start: int):
    """Take as many entries as fit into `room` bytes (incl. 2 header bytes) starting at index `start` (cyclic)."""
    if not entries:
        return b"", start
    body = bytes([ext_byte])
    taken = 0
    idx = start % len(entries)
    while taken < len(entries) and len(body) + len(entries[idx]) + 1 <= room and len(body) + len(entries[idx]) <= 29:
        body += entries[idx]
        idx = (idx + 1) % len(entries)
        taken += 1
    if taken == 0:
        return b"", start
    return _fig(0, body), idx


def make_fib(figs: bytes) -> bytes:
    assert len(figs) <= 30, len(figs)
    data = figs
    if len(data) < 30:
        data += b"\xFF"
        data = data.ljust(30, b"\x00")
    crc = C.crc16(data)
    return data + bytes([crc >> 8, crc & 0xFF])


class FicBuilder:
    def __init__(self, ens: EnsembleConfig):
        self.ens = ens
        self.e01 = [fig0_1_entry(s) for s in ens.subchannels]
        self.e02 = [fig0_2_entry(ens, sv) for sv in ens.services]
        self.labels = [fig1(0, ens.eid, ens.label)] + [fig1(1, sv.sid, sv.label) for sv in ens.services]
        rc = ens.reconfig
        if rc:
            self.n01 = self.e01 + [fig0_1_entry(s) for s in rc["subchannels"]]         # the next configuration, complete
            self.n02 = self.e02 + [fig0_2_entry(ens, sv) for sv in rc["services"]]
            self.nlabels = self.labels + [fig1(1, sv.sid, sv.label) for sv in rc["services"]]
            self.in01 = self.in02 = 0
        self.e13 = [fig0_13_entry(sv) for sv in ens.services if sv.user_apps]
        self.e13 += [fig0_13_entry_data(sv, dc) for sv in ens.services for dc in (sv.data_comps or []) if dc.user_apps]
        self.e03 = [fig0_3_entry(dc) for sv in ens.services for dc in (sv.data_comps or [])]
        self.e08 = [fig0_8_entry(sv, dc) for sv in ens.services for dc in (sv.data_comps or [])]
        self.e17 = [fig0_17_entry(sv) for sv in ens.services if sv.pty is not None]
        self.e18 = [fig0_18_entry(sv) for sv in ens.services if sv.asu is not None]
        self.i01 = self.i02 = self.ilab = self.i13 = self.i03 = self.i08 = self.i17 = self.i18 = 0

    def cif_fibs(self, cif_count: int) -> bytes:
        """Three FIBs (96 bytes) for one CIF."""
        rc = self.ens.reconfig
        rel = (cif_count - self.ens.cif_count0) % 5000
        announce = bool(rc) and rc["at"] - rc["lead"] <= rel < rc["at"]
        if rc and rel == rc["at"]:                                   # the new configuration is the current one from here on
            self.e01, self.e02, self.labels = self.n01, self.n02, self.nlabels
        f00 = fig0_0(self.ens, cif_count, 3 if announce else 0, (self.ens.cif_count0 + rc["at"]) % 250 if announce else 0)
        if announce and rel % 2 == 1:
            # every second CIF of the notice period carries the NEXT configuration (C/N = 1) instead of the rotating current one
            f01, self.in01 = _pack_rotating(self.n01, 0x81, 30 - len(f00), self.in01)
            f02, self.in02 = _pack_rotating(self.n02, 0x82, 30, self.in02)
            lab = self.labels[self.ilab % len(self.labels)]
            self.ilab += 1
            return make_fib(f00 + f01) + make_fib(f02) + make_fib(lab + fig0_9(self.ens))
        f13 = b""
        if self.e13 and cif_count % 2 == 0:
            f13, self.i13 = _pack_rotating(self.e13, 0x0D, 14, self.i13)
        elif self.e03 and cif_count % 2 == 1:                       # packet-mode components: FIG 0/3 + FIG 0/8 on the odd CIFs
            f03, self.i03 = _pack_rotating(self.e03, 0x03, 7, self.i03)
            f08, self.i08 = _pack_rotating(self.e08, 0x08, 7, self.i08)
            f13 = f03 + f08
        f01, self.i01 = _pack_rotating(self.e01, 0x01, 30 - len(f00) - len(f13), self.i01)
        fib0 = make_fib(f00 + f13 + f01)
        f02, self.i02 = _pack_rotating(self.e02, 0x02, 30, self.i02)
        extra = b""
        if self.e17 and cif_count % 4 == 1:
            f, self.i17 = _pack_rotating(self.e17, 0x11, 30 - len(f02), self.i17)
            extra += f
        if self.e18 and cif_count % 4 == 3:
            f, self.i18 = _pack_rotating(self.e18, 0x12, 30 - len(f02) - len(extra), self.i18)
            extra += f
        if self.ens.utc is not None and cif_count % 4 == 2:
            f = fig0_10(self.ens, cif_count - self.ens.cif_count0)
            if len(f02) + len(extra) + len(f) <= 30:
                extra += f
        if self.ens.announcements and cif_count % 2 == 0:
            rel = cif_count - self.ens.cif_count0
            e19 = [fig0_19_entry(cl, asw if a <= rel <= b else 0, sub) for a, b, cl, asw, sub in self.ens.announcements]
            f, _ = _pack_rotating(e19, 0x13, 30 - len(f02) - len(extra), 0)
            extra += f
        fib1 = make_fib(f02 + extra)
        lab = self.labels[self.ilab % len(self.labels)]
        self.ilab += 1
        fib2 = make_fib(lab + fig0_9(self.ens))
        return fib0 + fib1 + fib2


# ============================================================================ modulator
_PHASE8 = np.exp(1j * np.pi / 4 * np.arange(8)).astype(np.complex64)


@dataclass
class SynthResult:
    iq: np.ndarray                       # complex64 baseband before quantisation (None if not kept)
    n_frames: int
    fibs: np.ndarray                     # [n_cif, 96] uint8
    logical: Dict[int, np.ndarray]       # subchid -> [n_cif, bitrate*3] uint8 logical frames (pre-scrambling)
    aus: Dict[int, List[bytes]]          # subchid -> DAB+ access units in stream order
    superframes: Dict[int, np.ndarray]   # subchid -> [n_sf, 120*s]
    ens: EnsembleConfig = None
    frame0: int = 0                      # sample index of the first null symbol
    cfo_hz: float = 0.0


def build_cif_bits(ens: EnsembleConfig, n_cif: int, seed: int):
    """Payloads -> per-CIF code bits.  Returns (fic_bits [n_cif, 2304], msc_bits [n_cif, 55296], truth dicts)."""
    rng = np.random.default_rng(seed)
    ens.check()
    # ---- FIC
    fb = FicBuilder(ens)
    fibs = np.zeros((n_cif, 96), dtype=np.uint8)
    for c in range(n_cif):
        fibs[c] = np.frombuffer(fb.cif_fibs((ens.cif_count0 + c) % 5000), dtype=np.uint8)
    fic_in = bytes_to_bits(fibs) ^ C.prbs_bits(768)[None, :]
    fic_bits = encode_protected(fic_in, [(21, 16), (3, 15)], 0)
    assert fic_bits.shape[1] == C.FIC_PUNCT_BITS
    # ---- MSC
    msc_pre = rng.integers(0, 2, (n_cif, C.CIF_BITS), dtype=np.uint8)   # filler in unused CUs
    logical, aus, sfs = {}, {}, {}
    added = ens.reconfig["subchannels"] if ens.reconfig else []
    n_cif_all = n_cif
    for sc in list(ens.subchannels) + list(added):
        r0 = ens.reconfig["at"] if sc in added else 0           # a sub-channel of the next configuration starts at the change
        n_cif = max(n_cif_all - r0, 0)
        if n_cif == 0:
            continue
        srng = np.random.default_rng(sc.seed + 77)
        segs, size_cu, pad = sc.profile()
        if sc.kind == "dabplus":
            n_sf = (n_cif + 4) // 5
            sf, au = dabplus_superframes(sc, n_sf, srng)
            frames = sf.reshape(n_sf * 5, sc.frame_bytes)[:n_cif]
            aus[sc.subchid] = au
            sfs[sc.subchid] = sf
        elif sc.kind == "mp2":
            frames = mp2_frames(sc, n_cif, srng)
        elif sc.kind == "packet":
            frames = packet_frames(sc, n_cif, srng)
        else:
            frames = srng.integers(0, 256, (n_cif, sc.frame_bytes), dtype=np.uint8)
        logical[sc.subchid] = frames
        nb = sc.frame_bytes * 8
        scr = bytes_to_bits(frames) ^ C.prbs_bits(nb)[None, :]
        enc = encode_protected(scr, segs, pad)
        assert enc.shape[1] == size_cu * 64, (enc.shape, size_cu)
        msc_pre[r0:, sc.start_cu * 64:(sc.start_cu + size_cu) * 64] = enc
    n_cif = n_cif_all
    # ---- time interleaving: out[r][i] = in[r - D(i mod 16)][i]
    msc = np.zeros_like(msc_pre)
    for j, d in enumerate(C.TIME_DEINT_DELAY):
        if d == 0:
            msc[:, j::16] = msc_pre[:, j::16]
        elif d < n_cif:
            msc[d:, j::16] = msc_pre[:n_cif - d, j::16]
    return fic_bits, msc, dict(fibs=fibs, logical=logical, aus=aus, superframes=sfs)


def tii_symbol(ens: EnsembleConfig) -> np.ndarray:
    """Null-symbol TII signal (2656 samples, useful part last), unit carrier amplitude * amp."""
    spec = np.zeros(C.NFFT, dtype=np.complex128)
    prs = C.prs_carriers()
    kk = C.carrier_k()
    cidx = {int(k): i for i, k in enumerate(kk)}
    for tx in ens.tii:
        s = np.zeros(C.NFFT, dtype=np.complex128)
        for k in C.tii_carriers(tx.main, tx.sub):
            ph = prs[cidx[k]]
            s[k % C.NFFT] += ph
            s[(k + 1) % C.NFFT] += ph
        if tx.delay:
            f = np.fft.fftfreq(C.NFFT, 1.0 / C.NFFT)
            s *= np.exp(-2j * np.pi * f * tx.delay / C.NFFT)
        spec += tx.amp * np.exp(1j * tx.phase) * s
    t = np.fft.ifft(spec) * C.NFFT / np.sqrt(C.NCARR)
    return np.concatenate([t[-(C.T_NULL - C.NFFT):], t]).astype(np.complex64)


def modulate(ens: EnsembleConfig, n_frames: int, seed: int = 1) -> Tuple[np.ndarray, dict]:
    """Return (clean complex64 baseband of n_frames transmission frames at unit mean symbol power, truth)."""
    n_cif = 4 * n_frames
    fic_bits, msc_bits, truth = build_cif_bits(ens, n_cif, seed)
    prs_q = C.prs_quarter_phase()
    fidx = C.freq_interleave_index()        # QPSK symbol n -> carrier index
    bins = C.carrier_bins()
    tii = tii_symbol(ens) if ens.tii else None
    out = np.zeros(n_frames * C.T_FRAME, dtype=np.complex64)
    scale = C.NFFT / np.sqrt(C.NCARR)
    # QPSK phase index (units of pi/4): bits (re, im): 00->1, 10->3, 11->5, 01->7
    qmap = np.array([[1, 7], [3, 5]], dtype=np.int64)
    for f in range(n_frames):
        bits = np.concatenate([fic_bits[4 * f:4 * f + 4].reshape(-1), msc_bits[4 * f:4 * f + 4].reshape(-1)])
        bits = bits.reshape(75, C.BITS_PER_SYM)
        q = qmap[bits[:, :C.NCARR], bits[:, C.NCARR:]]            # [75, 1536] by QPSK symbol index n
        y = np.zeros((75, C.NCARR), dtype=np.int64)
        y[:, fidx] = q                                            # by carrier index
        ph = np.concatenate([2 * prs_q[None, :], y], axis=0)
        ph = np.cumsum(ph, axis=0) & 7                            # differential modulation
        spec = np.zeros((C.NSYM, C.NFFT), dtype=np.complex64)
        spec[:, bins] = _PHASE8[ph]
        t = (np.fft.ifft(spec, axis=1) * scale).astype(np.complex64)
        sym = np.concatenate([t[:, -C.T_CP:], t], axis=1).reshape(-1)
        base = f * C.T_FRAME
        cif0 = ens.cif_count0 + 4 * f
        if tii is not None and (ens.tii_every_null or cif0 % 8 == 0):
            out[base:base + C.T_NULL] = tii
        out[base + C.T_NULL:base + C.T_FRAME] = sym
    return out, truth


def resample_ppm(x: np.ndarray, ppm: float, taps: int = 16) -> np.ndarray:
    """Sample-clock offset: y[n] = x(n * (1 + ppm 1e-6)) by a Hann-windowed-sinc fractional-delay interpolator (16 taps)."""
    r = 1.0 + ppm * 1e-6
    n_out = int((x.size - taps) / r)
    out = np.zeros(n_out, dtype=np.complex64)
    step = 1 << 20
    half = taps // 2
    for s0 in range(0, n_out, step):
        pos = np.arange(s0, min(n_out, s0 + step), dtype=np.float64) * r + half
        k = np.floor(pos).astype(np.int64)
        mu = (pos - k).astype(np.float32)
        acc = np.zeros(pos.size, dtype=np.complex64)
        for m in range(-half + 1, half + 1):
            d = m - mu
            w = np.sinc(d) * (0.5 + 0.5 * np.cos(np.pi * d / half))
            acc += x[k + m] * w.astype(np.float32)
        out[s0:s0 + pos.size] = acc
    return out


def impair(x: np.ndarray, imp: Impairments, tail: int = 0) -> np.ndarray:
    """Apply multipath, CFO (+drift), lead-in, AWGN.  Output has unit mean symbol power before noise."""
    rng = np.random.default_rng(imp.seed)
    if imp.taps:
        y = x.copy()
        for d, g in imp.taps:
            y[d:] += np.complex64(g) * x[:len(x) - d]
        x = y
    if imp.sample_offset_ppm:
        x = resample_ppm(x, imp.sample_offset_ppm)
    x = np.concatenate([np.zeros(imp.lead_in, dtype=np.complex64), x, np.zeros(tail, dtype=np.complex64)])
    n = x.size
    if imp.cfo_hz or imp.cfo_drift_hz_per_s:
        # phase(t) = 2 pi (f0 t + 0.5 a t^2); evaluated in float64 in chunks to bound memory
        step = 1 << 20
        k = np.arange(step, dtype=np.float64)
        for s in range(0, n, step):
            m = min(n, s + step) - s
            # phase(s + k) = phi0 + w k + 0.5 a (k/fs)^2 ; chunk-local terms stay small enough for float32 sincos
            t0 = s / C.FS
            phi0 = 2 * np.pi * (imp.cfo_hz * t0 + 0.5 * imp.cfo_drift_hz_per_s * t0 * t0)
            w = 2 * np.pi * (imp.cfo_hz + imp.cfo_drift_hz_per_s * t0) / C.FS
            ph = np.mod(w * k[:m] + np.pi * imp.cfo_drift_hz_per_s * (k[:m] / C.FS) ** 2, 2 * np.pi).astype(np.float32)
            rot = (np.cos(ph) + 1j * np.sin(ph)).astype(np.complex64) * np.complex64(np.exp(1j * (phi0 % (2 * np.pi))))
            x[s:s + m] *= rot
    if imp.snr_db is not None:
        sigma = np.sqrt(0.5 * 10 ** (-imp.snr_db / 10))
        step = 1 << 22
        for s in range(0, n, step):
            m = min(n, s + step) - s
            nz = rng.standard_normal(2 * m, dtype=np.float32) * np.float32(sigma)
            x[s:s + m] += nz.view(np.complex64)
    return x


def add_noise_staircase(x: np.ndarray, lead_in: int, snr_db_per_frame, seed: int = 5) -> np.ndarray:
    """AWGN whose level changes from transmission frame to transmission frame (in place): frame f of a signal impaired with
    snr_db=None gets SNR snr_db_per_frame[f]; lead-in and tail get the level of the nearest frame (BASELINE configs[4]: SNR sweep)."""
    rng = np.random.default_rng(seed)
    snr = np.asarray(snr_db_per_frame, float)
    edges = [0] + [lead_in + (f + 1) * C.T_FRAME for f in range(len(snr) - 1)] + [x.size]
    for f in range(len(snr)):
        a, b = edges[f], min(edges[f + 1], x.size)
        if a >= b:
            continue
        sg = np.float32(np.sqrt(0.5 * 10 ** (-snr[f] / 10)))
        x[a:b] += (rng.standard_normal(2 * (b - a), dtype=np.float32) * sg).view(np.complex64)
    return x


def quantize(x: np.ndarray, fmt: str, rms: float) -> np.ndarray:
    """Complex baseband -> interleaved IQ in the raw-file formats RawFileInput reads
    (reference: src/input/rawfileinput.h:42-46, conversion rawfileinput.cpp:643-711), plus f32."""
    iq = np.empty(2 * x.size, dtype=np.float32)
    iq[0::2] = x.real
    iq[1::2] = x.imag
    iq *= np.float32(rms)     # symbol power is 1 => I and Q each have rms/sqrt(2); 'rms' is the complex RMS
    if fmt == "u8":
        return np.clip(np.rint(iq) + 128, 0, 255).astype(np.uint8)
    if fmt == "s16":
        return np.clip(np.rint(iq), -32768, 32767).astype(np.int16)
    if fmt == "f32":
        return iq
    raise ValueError(fmt)


def synthesize(ens: EnsembleConfig, n_frames: int, fmt: str = "u8", imp: Optional[Impairments] = None,
               seed: int = 1, tail: int = 0):
    """Full chain: returns (raw interleaved IQ array in `fmt`, SynthResult)."""
    imp = imp or Impairments()
    clean, truth = modulate(ens, n_frames, seed)
    x = impair(clean, imp, tail)
    rms = imp.rms if fmt != "s16" or imp.rms > 200 else imp.rms * 250.0
    raw = quantize(x, fmt, rms if fmt != "f32" else 1.0)
    res = SynthResult(iq=None, n_frames=n_frames, fibs=truth["fibs"], logical=truth["logical"], aus=truth["aus"],
                      superframes=truth["superframes"], ens=ens, frame0=imp.lead_in, cfo_hz=imp.cfo_hz)
    return raw, res
