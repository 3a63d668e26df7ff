"""DAB transmission mode I constants and tables (ETSI EN 300 401, TS 102 563).

These are the standard's tables as validated against libdabsdr 4.0.1 (SURVEY.md
Appendix A).  They are shared by the synthetic modulator (`synth.py`) and by the
host-side mirror of the receiver configuration (`pipeline.py`).  The open-source
restatements in the reference that the same facts can be cross-checked against:
TII pattern /root/reference/src/dabtables.cpp:2338-2444, EEP sizes
/root/reference/src/radiocontrol.cpp:1086-1108, channel list dabtables.cpp:32-39.
"""
from __future__ import annotations

import numpy as np

FS = 2_048_000
T_NULL = 2656
T_SYM = 2552
T_CP = 504
NFFT = 2048
NCARR = 1536
NSYM = 76                       # PRS + 75 data symbols
T_FRAME = T_NULL + NSYM * T_SYM  # 196608
BITS_PER_SYM = 2 * NCARR        # 3072
CIF_BITS = 55296
CU_BITS = 64
FIC_PUNCT_BITS = 2304           # per 768-bit FIC vector
FIB_BYTES = 32

# Band III channel table (centre frequency kHz) -- reference: src/dabtables.cpp:32-39
BAND3 = {
    "5A": 174928, "5B": 176640, "5C": 178352, "5D": 180064, "6A": 181936, "6B": 183648,
    "6C": 185360, "6D": 187072, "7A": 188928, "7B": 190640, "7C": 192352, "7D": 194064,
    "8A": 195936, "8B": 197648, "8C": 199360, "8D": 201072, "9A": 202928, "9B": 204640,
    "9C": 206352, "9D": 208064, "10A": 209936, "10B": 211648, "10C": 213360, "10D": 215072,
    "11A": 216928, "11B": 218640, "11C": 220352, "11D": 222064, "12A": 223936, "12B": 225648,
    "12C": 227360, "12D": 229072, "13A": 230784, "13B": 232496, "13C": 234208, "13D": 235776,
    "13E": 237488, "13F": 239200,
}

# ---------------------------------------------------------------- PRS (EN 300 401 14.3.2)
_H = [
    [0, 2, 0, 0, 0, 0, 1, 1, 2, 0, 0, 0, 2, 2, 1, 1] * 2,
    [0, 3, 2, 3, 0, 1, 3, 0, 2, 1, 2, 3, 2, 3, 3, 0] * 2,
    [0, 0, 0, 2, 0, 2, 1, 3, 2, 2, 0, 2, 2, 0, 1, 3] * 2,
    [0, 1, 2, 1, 0, 3, 3, 2, 2, 3, 2, 1, 2, 1, 3, 2] * 2,
]
_PRS_NEG = [(0, 1), (1, 2), (2, 0), (3, 1), (0, 3), (1, 2), (2, 2), (3, 3), (0, 2), (1, 1), (2, 2), (3, 3),
            (0, 1), (1, 2), (2, 3), (3, 3), (0, 2), (1, 2), (2, 2), (3, 1), (0, 1), (1, 3), (2, 1), (3, 2)]
_PRS_POS = [(0, 3), (3, 1), (2, 1), (1, 1), (0, 2), (3, 2), (2, 1), (1, 0), (0, 2), (3, 2), (2, 3), (1, 3),
            (0, 0), (3, 2), (2, 1), (1, 3), (0, 3), (3, 3), (2, 3), (1, 0), (0, 3), (3, 0), (2, 1), (1, 1)]


def carrier_k() -> np.ndarray:
    """Carrier numbers k in ascending order: -768..-1, 1..768 (index c = 0..1535)."""
    return np.concatenate([np.arange(-768, 0), np.arange(1, 769)])


def carrier_bins() -> np.ndarray:
    """FFT bin of carrier index c (natural FFT order, bin 0 = DC)."""
    return np.mod(carrier_k(), NFFT)


def prs_quarter_phase() -> np.ndarray:
    """PRS phase in units of pi/2 for carrier index c = 0..1535."""
    ph = np.zeros(NCARR, dtype=np.int64)
    for blk, (i, n) in enumerate(_PRS_NEG):
        for j in range(32):
            ph[blk * 32 + j] = (_H[i][j] + n) & 3
    for blk, (i, n) in enumerate(_PRS_POS):
        for j in range(32):
            ph[768 + blk * 32 + j] = (_H[i][j] + n) & 3
    return ph


def prs_carriers() -> np.ndarray:
    return np.exp(1j * np.pi / 2 * prs_quarter_phase())


# ---------------------------------------------------------------- frequency interleaver
def freq_interleave_carriers() -> np.ndarray:
    """k_n for n = 0..1535: QPSK symbol n is sent on carrier k_n (EN 300 401 14.6.1)."""
    pi = 0
    out = []
    for _ in range(1, 2048):
        pi = (13 * pi + 511) % 2048
        if 256 <= pi <= 1792 and pi != 1024:
            out.append(pi - 1024)
    assert len(out) == NCARR
    return np.array(out, dtype=np.int64)


def freq_interleave_index() -> np.ndarray:
    """carrier index c (0..1535, ascending k) of QPSK symbol n."""
    k = freq_interleave_carriers()
    return np.where(k < 0, k + 768, k + 767)


# ---------------------------------------------------------------- convolutional code
POLYS = (0o133, 0o171, 0o145, 0o133)  # MSB = current bit, LSB = oldest (a[i-6])

PI_MASKS = [
    0xC8888888, 0xC888C888, 0xC8C8C888, 0xC8C8C8C8, 0xCCC8C8C8, 0xCCC8CCC8, 0xCCCCCCC8, 0xCCCCCCCC,
    0xECCCCCCC, 0xECCCECCC, 0xECECECCC, 0xECECECEC, 0xEEECECEC, 0xEEECEEEC, 0xEEEEEEEC, 0xEEEEEEEE,
    0xFEEEEEEE, 0xFEEEFEEE, 0xFEFEFEEE, 0xFEFEFEFE, 0xFFFEFEFE, 0xFFFEFFFE, 0xFFFFFFFE, 0xFFFFFFFF,
]
TAIL_MASK24 = 0xCCCCCC

TIME_DEINT_DELAY = (0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15)


def pi_keep(pi: int) -> np.ndarray:
    """bool[32] keep pattern of puncturing index PI (1..24), first mother bit first."""
    m = PI_MASKS[pi - 1]
    return np.array([(m >> (31 - i)) & 1 for i in range(32)], dtype=bool)


# UEP table (EN 300 401 table 8 + table 31); idx: kbps, CU, level, L1..L4, PI1..PI4 (0 = unused), pad
# Row 23 uses the standard's PI2 = 17 (libdabsdr's own table holds 7, which is not self-consistent: SURVEY.md sec. 7).
_UEP = """
0:32/16/5 3,4,17,0 5,3,2,0 0;1:32/21/4 3,3,18,0 11,6,5,0 0;2:32/24/3 3,4,14,3 15,9,6,8 0;3:32/29/2 3,4,14,3 22,13,8,13 0;4:32/35/1 3,5,13,3 24,17,12,17 4;
5:48/24/5 4,3,26,3 5,4,2,3 0;6:48/29/4 3,4,26,3 9,6,4,6 0;7:48/35/3 3,4,26,3 15,10,6,9 4;8:48/42/2 3,4,26,3 24,14,8,15 0;9:48/52/1 3,5,25,3 24,18,13,18 0;
10:56/29/5 6,10,23,3 5,4,2,3 0;11:56/35/4 6,10,23,3 9,6,4,5 0;12:56/42/3 6,12,21,3 16,7,6,9 0;13:56/52/2 6,10,23,3 23,13,8,13 8;
14:64/32/5 6,9,31,2 5,3,2,3 0;15:64/42/4 6,9,33,0 11,6,5,0 0;16:64/48/3 6,12,27,3 16,8,6,9 0;17:64/58/2 6,10,29,3 23,13,8,13 8;18:64/70/1 6,11,28,3 24,18,12,18 4;
19:80/40/5 6,10,41,3 6,3,2,3 0;20:80/52/4 6,10,41,3 11,6,5,6 0;21:80/58/3 6,11,40,3 16,8,6,7 0;22:80/70/2 6,10,41,3 23,13,8,13 8;23:80/84/1 6,10,41,3 24,17,12,18 4;
24:96/48/5 7,9,53,3 5,4,2,4 0;25:96/58/4 7,10,52,3 9,6,4,6 0;26:96/70/3 6,12,51,3 16,9,6,10 4;27:96/84/2 6,10,53,3 22,12,9,12 0;28:96/104/1 6,13,50,3 24,18,13,19 0;
29:112/58/5 14,17,50,3 5,4,2,5 0;30:112/70/4 11,21,49,3 9,6,4,8 0;31:112/84/3 11,23,47,3 16,8,6,9 0;32:112/104/2 11,21,49,3 23,12,9,14 4;
33:128/64/5 12,19,62,3 5,3,2,4 0;34:128/84/4 11,21,61,3 11,6,5,7 0;35:128/96/3 11,22,60,3 16,9,6,10 4;36:128/116/2 11,21,61,3 22,12,9,14 0;37:128/140/1 11,20,62,3 24,17,13,19 8;
38:160/80/5 11,19,87,3 5,4,2,4 0;39:160/104/4 11,23,83,3 11,6,5,9 0;40:160/116/3 11,24,82,3 16,8,6,11 0;41:160/140/2 11,21,85,3 22,11,9,13 0;42:160/168/1 11,22,84,3 24,18,12,19 0;
43:192/96/5 11,20,110,3 6,4,2,5 0;44:192/116/4 11,22,108,3 10,6,4,9 0;45:192/140/3 11,24,106,3 16,10,6,11 0;46:192/168/2 11,20,110,3 22,13,9,13 8;47:192/208/1 11,21,109,3 24,20,13,24 0;
48:224/116/5 12,22,131,3 8,6,2,6 4;49:224/140/4 12,26,127,3 12,8,4,11 0;50:224/168/3 11,20,134,3 16,10,7,9 0;51:224/208/2 11,22,132,3 24,16,10,15 0;52:224/232/1 11,24,130,3 24,20,12,20 4;
53:256/128/5 11,24,154,3 6,5,2,5 0;54:256/168/4 11,24,154,3 12,9,5,10 4;55:256/192/3 11,27,151,3 16,10,7,10 0;56:256/232/2 11,22,156,3 24,14,10,13 8;57:256/280/1 11,26,152,3 24,19,14,18 4;
58:320/160/5 11,26,200,3 8,5,2,6 4;59:320/208/4 11,25,201,3 13,9,5,10 8;60:320/280/2 11,26,200,3 24,17,9,17 0;
61:384/192/5 11,27,247,3 8,6,2,7 0;62:384/280/3 11,24,250,3 16,9,7,10 4;63:384/416/1 12,28,245,3 24,20,14,23 8
"""


def _parse_uep():
    rows = {}
    for item in _UEP.replace("\n", "").split(";"):
        item = item.strip()
        if not item:
            continue
        idx, rest = item.split(":")
        a, ls, pis, pad = rest.split(" ")
        kbps, cu, lvl = (int(x) for x in a.split("/"))
        rows[int(idx)] = dict(bitrate=kbps, size_cu=cu, level=lvl, L=[int(x) for x in ls.split(",")],
                              PI=[int(x) for x in pis.split(",")], pad=int(pad))
    assert len(rows) == 64
    return [rows[i] for i in range(64)]


UEP_TABLE = _parse_uep()


def eep_profile(bitrate: int, level: int, option: str):
    """Return ([(L, PI), ...], size_cu) for EEP `level` 1..4, option 'A' or 'B' (EN 300 401 11.3.2)."""
    if option == "A":
        n = bitrate // 8
        assert bitrate % 8 == 0 and n >= 1
        if level == 1:
            return [(6 * n - 3, 24), (3, 23)], 12 * n
        if level == 2:
            if n == 1:
                return [(5, 13), (1, 12)], 8
            return [(2 * n - 3, 14), (4 * n + 3, 13)], 8 * n
        if level == 3:
            return [(6 * n - 3, 8), (3, 7)], 6 * n
        if level == 4:
            return [(4 * n - 3, 3), (2 * n + 3, 2)], 4 * n
    else:
        m = bitrate // 32
        assert bitrate % 32 == 0 and m >= 1
        pis = {1: (10, 9, 27), 2: (6, 5, 21), 3: (4, 3, 18), 4: (2, 1, 15)}[level]
        return [(24 * m - 3, pis[0]), (3, pis[1])], pis[2] * m
    raise ValueError("bad EEP profile")


def uep_profile(index: int):
    r = UEP_TABLE[index]
    segs = [(l, p) for l, p in zip(r["L"], r["PI"]) if l > 0]
    return segs, r["size_cu"], r["pad"], r["bitrate"]


def punct_keep_mask(segs, n_info_bits: int, pad: int = 0) -> np.ndarray:
    """bool[(n_info+6)*4] keep mask over mother-code bits for a protection profile.

    segs = [(L blocks of 128 mother bits, PI)], then 24 tail mother bits with mask CCCCCC.
    `pad` zero bits are appended after puncturing by the transmitter (UEP only) and are not part of this mask.
    """
    parts = []
    for l, pi in segs:
        parts.append(np.tile(pi_keep(pi), 4 * l))
    tail = np.array([(TAIL_MASK24 >> (23 - i)) & 1 for i in range(24)], dtype=bool)
    parts.append(tail)
    m = np.concatenate(parts)
    assert m.size == (n_info_bits + 6) * 4, (m.size, n_info_bits)
    return m


# ---------------------------------------------------------------- TII
def tii_pattern_table() -> np.ndarray:
    """70 bytes with exactly four bits set, ascending; MSB = comb block b = 0 (EN 300 401 table 42)."""
    vals = [v for v in range(256) if bin(v).count("1") == 4]
    assert len(vals) == 70
    return np.array(vals, dtype=np.uint8)


def tii_blocks(main_id: int):
    pat = int(tii_pattern_table()[main_id])
    return [b for b in range(8) if (pat >> (7 - b)) & 1]


def tii_carriers(main_id: int, sub_id: int):
    """Lower carrier k of each active pair for (main, sub) in mode I; the pair is (k, k+1)."""
    ks = []
    for base in (-768, -384, 1, 385):
        for b in tii_blocks(main_id):
            ks.append(base + 2 * sub_id + 48 * b)
    return ks


# ---------------------------------------------------------------- scrambler / CRC
def prbs_bits(n: int) -> np.ndarray:
    """Energy-dispersal PRBS x^9 + x^5 + 1, all-ones init; first bits 0000 0111 1011 1110."""
    reg = [1] * 9
    out = np.empty(n, dtype=np.uint8)
    for i in range(n):
        b = reg[8] ^ reg[4]
        out[i] = b
        reg = [b] + reg[:8]
    return out


def crc16(data: bytes) -> int:
    """CRC-16-CCITT poly 0x1021 init 0xFFFF, result inverted (reference: src/data/mscdatagroup.cpp:164-185)."""
    crc = 0xFFFF
    for byte in data:
        crc ^= byte << 8
        for _ in range(8):
            crc = ((crc << 1) ^ 0x1021) & 0xFFFF if crc & 0x8000 else (crc << 1) & 0xFFFF
    return crc ^ 0xFFFF


def firecode(data: bytes) -> int:
    """DAB+ fire code, poly 0x782F, init 0, over 9 bytes (TS 102 563 5.2)."""
    crc = 0
    for byte in data:
        crc ^= byte << 8
        for _ in range(8):
            crc = ((crc << 1) ^ 0x782F) & 0xFFFF if crc & 0x8000 else (crc << 1) & 0xFFFF
    return crc
