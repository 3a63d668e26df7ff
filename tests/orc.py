"""ctypes binding of the CPU checker (oracle/_ref/libdaboracle.so) and runner of the reference harness.

Test infrastructure only: imported by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline leg.
"""
from __future__ import annotations

import ctypes as ct
import json
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
REF_DIR = ORACLE_DIR / "_ref"
FMT = {"u8": 0, "s16": 1, "f32": 2}
FMT_BYTES = {"u8": 2, "s16": 4, "f32": 8}


class SubCh(ct.Structure):
    _fields_ = [("start_cu", ct.c_int32), ("size_cu", ct.c_int32), ("n_info", ct.c_int32), ("n_seg", ct.c_int32),
                ("seg_len", ct.c_int32 * 4), ("seg_pi", ct.c_int32 * 4)]


def subch_struct(sc) -> SubCh:
    """sc: abracadabra_b200.synth.SubChannel"""
    segs, size, _pad = sc.profile()
    s = SubCh(sc.start_cu, size, sc.bitrate * 24, len(segs))
    for i, (l, p) in enumerate(segs):
        s.seg_len[i] = l
        s.seg_pi[i] = p
    return s


_lib = None


def build():
    subprocess.run(["make", "-C", str(ORACLE_DIR), "-s"], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        p = REF_DIR / "libdaboracle.so"
        if not p.exists():
            build()
        L = ct.CDLL(str(p))
        vp, i32, i64, u64, f32p = ct.c_void_p, ct.c_int32, ct.c_int64, ct.c_uint64, ct.POINTER(ct.c_float)
        L.orc_block_power.argtypes = [vp, i32, u64, vp]
        L.orc_null_metric.argtypes = [vp, i64, vp]
        L.orc_candidates.argtypes = [vp, i64, u64, vp, i32]
        L.orc_candidates.restype = i32
        L.orc_chain_frames.argtypes = [vp, vp, i32, vp, vp]
        L.orc_chain_frames.restype = i32
        L.orc_frame_sync.argtypes = [vp, i32, u64, i64, ct.POINTER(i64), f32p, vp]
        L.orc_demod_frame.argtypes = [vp, i32, u64, i64, ct.c_float, vp, vp, vp]
        L.orc_fft2048.argtypes = [vp, i32]
        L.orc_depuncture_map.argtypes = [ct.POINTER(SubCh), vp]
        L.orc_depuncture_map.restype = i32
        L.orc_viterbi.argtypes = [vp, i32, vp]
        L.orc_fic_decode.argtypes = [vp, vp, vp]
        L.orc_msc_decode.argtypes = [vp, i32, ct.POINTER(SubCh), vp]
        L.orc_crc16.argtypes = [vp, i32]
        L.orc_crc16.restype = ct.c_uint16
        L.orc_firecode.argtypes = [vp]
        L.orc_firecode.restype = ct.c_uint16
        L.orc_prbs.argtypes = [vp, i32]
        L.orc_rs_decode.argtypes = [vp, ct.POINTER(i32)]
        L.orc_rs_decode.restype = i32
        L.orc_superframe.argtypes = [vp, i32, vp, vp, vp, ct.POINTER(i32), ct.POINTER(i32), ct.POINTER(i32), ct.POINTER(i32)]
        L.orc_superframe.restype = i32
        L.orc_tii_null.argtypes = [vp, i32, u64, i64, ct.c_float, vp]
        L.orc_tii_fold.argtypes = [vp, vp]
        L.orc_tii_detect.argtypes = [vp, vp, vp, i32, ct.c_float, vp, vp, vp, i32, vp]
        L.orc_tii_detect.restype = i32
        L.orc_prs_rms.argtypes = [vp, i32, u64, i64, ct.c_float]
        L.orc_prs_rms.restype = ct.c_float
        L.orc_decode.argtypes = [vp, i32, u64, ct.POINTER(SubCh), i32, vp, vp, ct.POINTER(vp), vp, vp, i32]
        L.orc_decode.restype = i32
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ct.c_void_p)


def nsamples(raw: np.ndarray, fmt: str) -> int:
    return raw.nbytes // FMT_BYTES[fmt]


# ------------------------------------------------------------------ stage wrappers
def block_power(raw, fmt):
    n = nsamples(raw, fmt)
    P = np.zeros(n // 32, dtype=np.float32)
    lib().orc_block_power(_p(raw), FMT[fmt], n, _p(P))
    return P


def null_metric(P):
    q = np.zeros((4, P.size), dtype=np.float32)
    lib().orc_null_metric(_p(P), P.size, _p(q))
    return q


def candidates(q, nsamp, max_cand=None):
    max_cand = max_cand or q.shape[1] // 83 + 2
    st = np.zeros(max_cand, dtype=np.int64)
    n = lib().orc_candidates(_p(q), q.shape[1], nsamp, _p(st), max_cand)
    return st[:n].copy()


def chain_frames(t_f, cir_ratio):
    t_f = np.ascontiguousarray(t_f, dtype=np.int64)
    cir_ratio = np.ascontiguousarray(cir_ratio, dtype=np.float32)
    idx = np.zeros(t_f.size, dtype=np.int32)
    slot = np.zeros(t_f.size, dtype=np.int32)
    n = lib().orc_chain_frames(_p(t_f), _p(cir_ratio), t_f.size, _p(idx), _p(slot))
    return idx[:n].copy(), slot[:n].copy()


def frame_sync(raw, fmt, null_start):
    t = ct.c_int64()
    cfo = ct.c_float()
    m = np.zeros(4, dtype=np.float32)
    lib().orc_frame_sync(_p(raw), FMT[fmt], nsamples(raw, fmt), int(null_start), ct.byref(t), ct.byref(cfo), _p(m))
    return t.value, cfo.value, m


def demod_frame(raw, fmt, t_f, cfo, want_carriers=False):
    soft = np.zeros(75 * 3072, dtype=np.int8)
    car = np.zeros(76 * 1536 * 2, dtype=np.float32) if want_carriers else None
    stat = np.zeros(150, dtype=np.float32)
    lib().orc_demod_frame(_p(raw), FMT[fmt], nsamples(raw, fmt), int(t_f), float(cfo), _p(soft),
                          _p(car) if want_carriers else None, _p(stat))
    if want_carriers:
        return soft, car.view(np.complex64).reshape(76, 1536), stat.reshape(75, 2)
    return soft


def fft2048(x: np.ndarray, inverse=False) -> np.ndarray:
    y = np.ascontiguousarray(x, dtype=np.complex64).copy()
    lib().orc_fft2048(_p(y), int(inverse))
    return y


def depuncture_map(sc: SubCh) -> np.ndarray:
    m = np.zeros(4 * (sc.n_info + 6), dtype=np.int32)
    r = lib().orc_depuncture_map(ct.byref(sc), _p(m))
    assert r >= 0
    return m


def viterbi(soft4: np.ndarray) -> np.ndarray:
    soft4 = np.ascontiguousarray(soft4, dtype=np.int8)
    steps = soft4.size // 4
    bits = np.zeros(steps - 6, dtype=np.uint8)
    lib().orc_viterbi(_p(soft4), steps, _p(bits))
    return bits


def fic_decode(soft_fic: np.ndarray):
    fibs = np.zeros((12, 32), dtype=np.uint8)
    ok = np.zeros(12, dtype=np.uint8)
    lib().orc_fic_decode(_p(np.ascontiguousarray(soft_fic, dtype=np.int8)), _p(fibs), _p(ok))
    return fibs, ok


def msc_decode(soft_cifs: np.ndarray, sc: SubCh) -> np.ndarray:
    soft_cifs = np.ascontiguousarray(soft_cifs, dtype=np.int8)
    n_cif = soft_cifs.shape[0]
    out = np.zeros((max(n_cif - 15, 0), sc.n_info // 8), dtype=np.uint8)
    lib().orc_msc_decode(_p(soft_cifs), n_cif, ct.byref(sc), _p(out))
    return out


def rs_decode(cw: np.ndarray):
    cw = np.ascontiguousarray(cw, dtype=np.uint8).copy()
    be = ct.c_int32()
    r = lib().orc_rs_decode(_p(cw), ct.byref(be))
    return r, be.value, cw


def superframe(sf: np.ndarray, s: int):
    sf = np.ascontiguousarray(sf, dtype=np.uint8).copy()
    off = np.zeros(8, dtype=np.int32)
    ln = np.zeros(8, dtype=np.int32)
    ok = np.zeros(8, dtype=np.int32)
    c, b, u, f = ct.c_int32(), ct.c_int32(), ct.c_int32(), ct.c_int32()
    n = lib().orc_superframe(_p(sf), s, _p(off), _p(ln), _p(ok), ct.byref(c), ct.byref(b), ct.byref(u), ct.byref(f))
    aus = [bytes(sf[off[i]:off[i] + ln[i]]) for i in range(n)]
    return dict(n=n, aus=aus, crc_ok=ok[:n].tolist(), corrected=c.value, bit_errors=b.value, uncorrectable=u.value,
                firecode_ok=f.value, data=sf)


def tii_null_spectrum(raw, fmt, t_f, cfo):
    spec = np.zeros(2048, dtype=np.complex64)
    lib().orc_tii_null(_p(raw), FMT[fmt], nsamples(raw, fmt), int(t_f), float(cfo), _p(spec))
    return spec


def tii_fold(spec: np.ndarray) -> np.ndarray:
    spec = np.ascontiguousarray(spec, dtype=np.complex64)
    fold = np.zeros(384, dtype=np.float32)
    lib().orc_tii_fold(_p(spec), _p(fold))
    return fold


def prs_rms(raw, fmt, t_f, cfo) -> float:
    return float(lib().orc_prs_rms(_p(raw), FMT[fmt], nsamples(raw, fmt), int(t_f), float(cfo)))


def tii_detect(fold, norm, sign, thr_k=6.0):
    fold = np.ascontiguousarray(fold, dtype=np.float32)
    norm = np.ascontiguousarray(norm, dtype=np.float32)
    sign = np.ascontiguousarray(sign, dtype=np.int32)
    mains = np.zeros(24, np.uint8); subs = np.zeros(24, np.uint8); lv = np.zeros(24, np.float32)
    spec = np.zeros(384, np.float32)
    n = lib().orc_tii_detect(_p(fold), _p(norm), _p(sign), fold.shape[0], float(thr_k), _p(mains), _p(subs), _p(lv), 24, _p(spec))
    return [(int(mains[i]), int(subs[i]), float(lv[i])) for i in range(n)], spec


def decode(raw, fmt, subchs, max_frames=None):
    """Whole chain.  Returns dict(nf, fibs [nf*4, 96], fib_ok [nf*12], frames {i: [n_cif-15, bytes]}, t_f, cfo)."""
    n = nsamples(raw, fmt)
    max_frames = max_frames or (n // 196608 + 2)
    fibs = np.zeros((max_frames * 4, 96), dtype=np.uint8)
    ok = np.zeros(max_frames * 12, dtype=np.uint8)
    t_f = np.zeros(max_frames, dtype=np.int64)
    cfo = np.zeros(max_frames, dtype=np.float32)
    arr = (SubCh * max(len(subchs), 1))(*subchs)
    outs = [np.zeros((max(max_frames * 4 - 15, 1), sc.n_info // 8), dtype=np.uint8) for sc in subchs]
    ptrs = (ct.c_void_p * max(len(subchs), 1))(*[o.ctypes.data for o in outs])
    nf = lib().orc_decode(_p(raw), FMT[fmt], n, arr, len(subchs), _p(fibs), _p(ok), ptrs, _p(t_f), _p(cfo), max_frames)
    return dict(nf=nf, fibs=fibs[:nf * 4], fib_ok=ok[:nf * 12], frames=[o[:max(nf * 4 - 15, 0)] for o in outs],
                t_f=t_f[:nf], cfo=cfo[:nf])


# ------------------------------------------------------------------ reference library runner
def have_ref() -> bool:
    return (REF_DIR / "ref_harness").exists() and (REF_DIR / "libdabsdr.so.4.0.1").exists()


def run_ref(path, fmt, log_path=None, tii=None, service=None, second=None, extra=(), harness=None, timeout=600, env=None):
    """Run the harness (by default against the reference core).  Returns (summary dict, [events])."""
    exe = str(harness or (REF_DIR / "ref_harness"))
    cmd = [exe, str(path), fmt, str(log_path) if log_path else "-"]
    if tii is not None:
        cmd += ["--tii", str(tii)]
    if service is not None:
        cmd += ["--service", str(service)]
    if second is not None:
        cmd += ["--second", str(second)]
    cmd += list(extra)
    out = subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=timeout, env=env).stdout
    summary = json.loads(out.strip().splitlines()[-1])
    events = []
    if log_path and os.path.exists(log_path):
        with open(log_path) as f:
            events = [json.loads(line) for line in f if line.strip()]
    return summary, events
