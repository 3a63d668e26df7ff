"""Host-side binding of libb200dab.so (include/b200dab.h) -- the CUDA receive path behind a plain C ABI.

`Receiver.decode()` is the batch form of what AbracaDABra's RadioControl obtains from the closed `dabsdr` core through
the pull/callback interface (reference: lib/linux_x86_64/dabsdr.h:397-429, src/radiocontrol.cpp:81-104, :2235-2609):
raw IQ as RawFileInput reads it (src/input/rawfileinput.cpp:643-711) in, FIBs / logical frames / DAB+ access units /
statistics / TII ids out.  There is no CPU fallback: if the extension is not built or no CUDA device is present this
module raises.
"""
from __future__ import annotations

import ctypes as ct
import subprocess
from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "lib" / "libb200dab.so"
SHIM_PATH = PKG / "lib" / "libdabsdr.so.4"
FMT = {"u8": 0, "s16": 1, "f32": 2}
FMT_BYTES = {"u8": 2, "s16": 4, "f32": 8}
OPT_TII, OPT_NO_MSC, OPT_SYMSTAT, OPT_SPECTRUM = 1, 2, 4, 8
TII_MAX_IDS, TII_BINS = 24, 384
FRAME_SOFT = 75 * 3072


class SubCh(ct.Structure):
    _fields_ = [("subch_id", ct.c_int32), ("start_cu", ct.c_int32), ("size_cu", ct.c_int32), ("bitrate", ct.c_int32),
                ("n_seg", ct.c_int32), ("seg_len", ct.c_int32 * 4), ("seg_pi", ct.c_int32 * 4), ("kind", ct.c_int32),
                ("uep_index", ct.c_int32), ("prot_level", ct.c_int32)]


class Au(ct.Structure):
    _fields_ = [("subch_index", ct.c_int32), ("superframe", ct.c_int32), ("cif", ct.c_int32), ("offset", ct.c_int32),
                ("length", ct.c_int32), ("crc_ok", ct.c_int32), ("header", ct.c_int32)]


class Tii(ct.Structure):
    _fields_ = [("n_ids", ct.c_int32), ("n_tii_frames", ct.c_int32), ("norm", ct.c_float), ("noise", ct.c_float),
                ("main_id", ct.c_uint8 * TII_MAX_IDS), ("sub_id", ct.c_uint8 * TII_MAX_IDS),
                ("level", ct.c_float * TII_MAX_IDS), ("spectrum", ct.c_float * TII_BINS)]


PAD_CB = ct.CFUNCTYPE(None, ct.c_void_p, ct.c_int, ct.POINTER(ct.c_uint8), ct.c_int)


class ServiceInfo(ct.Structure):
    _fields_ = [("sid", ct.c_uint32), ("label", ct.c_char * 17), ("n_components", ct.c_uint8), ("subch_id", ct.c_uint8 * 4),
                ("ascty", ct.c_uint8 * 4)]


class EnsembleInfo(ct.Structure):
    _fields_ = [("ueid", ct.c_uint32), ("lto", ct.c_int32), ("int_table", ct.c_int32), ("alarm", ct.c_int32), ("label", ct.c_char * 17),
                ("first_cif_count", ct.c_int32), ("n_services", ct.c_int32), ("services", ServiceInfo * 64)]


class Sf(ct.Structure):
    _fields_ = [("subch_index", ct.c_int32), ("row", ct.c_int32), ("firecode_ok", ct.c_int32), ("rs_corrected", ct.c_int32),
                ("rs_bit_errors", ct.c_int32), ("rs_uncorrectable", ct.c_int32), ("n_aus", ct.c_int32), ("header", ct.c_int32)]


class Result(ct.Structure):
    _fields_ = [("n_candidates", ct.c_int32), ("n_frames", ct.c_int32), ("n_slots", ct.c_int32),
                ("t_f", ct.POINTER(ct.c_int64)), ("cfo_hz", ct.POINTER(ct.c_float)), ("cir_ratio", ct.POINTER(ct.c_float)),
                ("fibs", ct.POINTER(ct.c_uint8)), ("fib_ok", ct.POINTER(ct.c_uint8)),
                ("n_subch", ct.c_int32), ("subch", ct.POINTER(SubCh)),
                ("frames", ct.POINTER(ct.POINTER(ct.c_uint8))), ("frame_rows", ct.POINTER(ct.c_int32)),
                ("n_aus", ct.c_int32), ("aus", ct.POINTER(Au)),
                ("rs_corrected", ct.c_int64), ("rs_bit_errors", ct.c_int64), ("rs_uncorrectable", ct.c_int64),
                ("firecode_fail", ct.c_int64), ("superframes", ct.c_int64),
                ("n_tii", ct.c_int32), ("tii", ct.POINTER(Tii)),
                ("symstat", ct.POINTER(ct.c_float)), ("snr_db", ct.c_float), ("gpu_launches", ct.c_int32),
                ("tii_fold", ct.POINTER(ct.c_float)), ("tii_norm", ct.POINTER(ct.c_float)), ("tii_sign", ct.POINTER(ct.c_int32)),
                ("spectrum", ct.POINTER(ct.c_float)),
                ("n_sf", ct.c_int32), ("sf", ct.POINTER(Sf)), ("sf_flags", ct.POINTER(ct.POINTER(ct.c_uint8))),
                ("sf_phase", ct.POINTER(ct.c_int32))]


def build(verbose: bool = False) -> None:
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", str(PKG / "csrc"), "-s", "all"], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libb200dab.so failed:\n" + (r.stdout or "") + (r.stderr or ""))


_lib = None


def lib():
    """Load libb200dab.so; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: run abracadabra_b200.pipeline.build() (or __graft_entry__.build())")
        L = ct.CDLL(str(LIB_PATH))
        vp, i32, u64 = ct.c_void_p, ct.c_int32, ct.c_uint64
        L.b200dab_version.restype = i32
        L.b200dab_create.argtypes = [ct.POINTER(vp), i32]
        L.b200dab_destroy.argtypes = [vp]
        L.b200dab_last_error.argtypes = [vp]
        L.b200dab_last_error.restype = ct.c_char_p
        L.b200dab_stream.argtypes = [vp]
        L.b200dab_stream.restype = vp
        L.b200dab_profile_eep.argtypes = [i32, i32, i32, ct.POINTER(SubCh)]
        L.b200dab_profile_uep.argtypes = [i32, ct.POINTER(SubCh)]
        L.b200dab_parse_subch.argtypes = [vp, vp, i32, ct.POINTER(SubCh), i32]
        L.b200dab_spectrum_layout.argtypes = [vp, i32]
        L.b200dab_parse_ensemble.argtypes = [vp, vp, i32, ct.POINTER(EnsembleInfo)]
        L.b200dab_pad_create.restype = vp
        L.b200dab_pad_destroy.argtypes = [vp]
        L.b200dab_pad_reset.argtypes = [vp]
        L.b200dab_pad_feed_au.argtypes = [vp, vp, i32, PAD_CB, vp]
        L.b200dab_pad_feed_mp2.argtypes = [vp, vp, i32, PAD_CB, vp]
        L.b200dab_pad_feed_packets.argtypes = [vp, i32, vp, i32, PAD_CB, vp]
        L.b200dab_decode.argtypes = [vp, vp, u64, i32, i32, ct.POINTER(SubCh), i32, ct.c_uint32, ct.POINTER(Result)]
        L.b200dab_set_profiling.argtypes = [vp, i32]
        L.b200dab_set_superframe_phase.argtypes = [vp, vp, i32]
        L.b200dab_set_viterbi_split.argtypes = [vp, i32]
        L.b200dab_viterbi4_decision_index.argtypes = [i32, i32]
        L.b200dab_get_kernel_times.argtypes = [vp, ct.c_char_p, i32, vp, i32]
        L.b200dab_get_kernel_times_back.argtypes = [vp, i32, ct.c_char_p, i32, vp, i32]
        L.b200dab_get_kernel_timeline.argtypes = [vp, i32, ct.c_char_p, i32, vp, vp, i32]
        L.b200dab_stage_coarse.argtypes = [vp, vp, u64, i32, vp, vp, vp, ct.POINTER(i32), i32]
        L.b200dab_stage_fine.argtypes = [vp, vp, u64, i32, vp, i32, vp, vp, vp]
        L.b200dab_stage_demod.argtypes = [vp, vp, u64, i32, vp, vp, i32, vp]
        L.b200dab_stage_viterbi.argtypes = [vp, vp, i32, i32, vp]
        L.b200dab_stage_fic.argtypes = [vp, vp, i32, vp, vp]
        L.b200dab_stage_msc.argtypes = [vp, vp, i32, ct.POINTER(SubCh), vp]
        L.b200dab_stage_dabplus.argtypes = [vp, vp, i32, i32, ct.POINTER(Au), ct.POINTER(i32), i32, vp]
        L.b200dab_stage_spectrum.argtypes = [vp, vp, u64, i32, vp, i32, vp]
        L.b200dab_stage_tii_fold.argtypes = [vp, vp, i32, vp]
        L.b200dab_stage_tii_detect.argtypes = [vp, vp, vp, vp, i32, ct.c_float, ct.POINTER(Tii)]
        _lib = L
    return _lib


EXPORTS = ["b200dab_version", "b200dab_create", "b200dab_destroy", "b200dab_last_error", "b200dab_stream", "b200dab_host_alloc", "b200dab_host_free",
           "b200dab_set_profiling", "b200dab_get_kernel_times", "b200dab_get_kernel_times_back", "b200dab_get_kernel_timeline", "b200dab_profile_eep", "b200dab_profile_uep", "b200dab_parse_subch", "b200dab_parse_ensemble", "b200dab_spectrum_layout", "b200dab_viterbi_decision_index", "b200dab_viterbi4_decision_index", "b200dab_set_viterbi_split", "b200dab_set_superframe_phase", "b200dab_pad_create", "b200dab_pad_destroy", "b200dab_pad_reset", "b200dab_pad_feed_au", "b200dab_pad_feed_mp2", "b200dab_pad_feed_packets", "b200dab_decode", "b200dab_stage_coarse",
           "b200dab_stage_fine", "b200dab_stage_demod", "b200dab_stage_viterbi", "b200dab_stage_fic", "b200dab_stage_msc",
           "b200dab_stage_dabplus", "b200dab_stage_spectrum", "b200dab_stage_tii_fold", "b200dab_stage_tii_detect"]


def subch_eep(subch_id: int, start_cu: int, bitrate: int, level: int, option: str, dabplus: bool) -> SubCh:
    s = SubCh()
    if lib().b200dab_profile_eep(bitrate, level, 0 if option == "A" else 1, ct.byref(s)) != 0:
        raise ValueError("invalid EEP profile")
    s.subch_id, s.start_cu, s.kind = subch_id, start_cu, int(dabplus)
    return s


def subch_uep(subch_id: int, start_cu: int, table_index: int, dabplus: bool = False) -> SubCh:
    s = SubCh()
    if lib().b200dab_profile_uep(table_index, ct.byref(s)) != 0:
        raise ValueError("invalid UEP index")
    s.subch_id, s.start_cu, s.kind = subch_id, start_cu, int(dabplus)
    return s


def subch_from_synth(sc) -> SubCh:
    """abracadabra_b200.synth.SubChannel -> SubCh"""
    if sc.prot[0] == "EEP":
        return subch_eep(sc.subchid, sc.start_cu, sc.bitrate, sc.prot[1], sc.prot[2], sc.kind == "dabplus")
    return subch_uep(sc.subchid, sc.start_cu, sc.prot[1], sc.kind == "dabplus")


def parse_subch(fibs: np.ndarray, fib_ok: Optional[np.ndarray] = None) -> List[SubCh]:
    fibs = np.ascontiguousarray(fibs, dtype=np.uint8).reshape(-1, 32)
    ok = None if fib_ok is None else np.ascontiguousarray(fib_ok, dtype=np.uint8)
    out = (SubCh * 64)()
    n = lib().b200dab_parse_subch(fibs.ctypes.data, None if ok is None else ok.ctypes.data, fibs.shape[0], out, 64)
    if n < 0:
        raise RuntimeError("parse_subch failed")
    return [out[i] for i in range(n)]


def spectrum_layout() -> np.ndarray:
    """[3584] uint16: slots of the OFDM kernel's spectrum buffer (see b200dab_spectrum_layout)"""
    out = np.zeros(2048 + 1536, np.uint16)
    if lib().b200dab_spectrum_layout(out.ctypes.data, out.size) != out.size:
        raise RuntimeError("spectrum layout failed")
    return out


class PadDecoder:
    """X-PAD decoder of one DAB+ audio service: feed access units in order, get (kind, bytes) per complete data group
    (kind 0 = dynamic label data group, kind >= 4 = MSC data group of that X-PAD application type)."""

    def __init__(self):
        self._h = lib().b200dab_pad_create()
        self.out = []
        self._cb = PAD_CB(lambda user, kind, data, n: self.out.append((kind, bytes(data[:n]))))

    def feed(self, au: bytes) -> bool:
        buf = (ct.c_uint8 * len(au)).from_buffer_copy(au)
        return bool(lib().b200dab_pad_feed_au(self._h, buf, len(au), self._cb, None))

    def feed_mp2(self, frame: bytes) -> bool:
        buf = (ct.c_uint8 * len(frame)).from_buffer_copy(frame)
        return bool(lib().b200dab_pad_feed_mp2(self._h, buf, len(frame), self._cb, None))

    def feed_packets(self, addr: int, frame: bytes) -> bool:
        """one logical frame of a packet-mode sub-channel; complete data groups of `addr` come out with kind -1"""
        buf = (ct.c_uint8 * len(frame)).from_buffer_copy(frame)
        return bool(lib().b200dab_pad_feed_packets(self._h, addr, buf, len(frame), self._cb, None))

    def close(self):
        if self._h:
            lib().b200dab_pad_destroy(self._h)
            self._h = None


def dynamic_label_text(groups) -> list:
    """Assemble dynamic-label data groups into messages the way DLDecoder does (src/data/dldecoder.cpp:139-186)."""
    msgs, cur, seg = [], None, 0
    for g in groups:
        if g[0] & 0x10:
            continue
        n = (g[0] & 0x0F) + 1
        if g[0] & 0x40:
            cur, seg = b"", 0
        elif cur is None or seg != ((g[1] >> 4) & 7):
            cur = None
            continue
        cur += g[2:2 + n]
        seg += 1
        if g[0] & 0x20:
            msgs.append(cur.decode("latin1"))
            cur = None
    return msgs


def parse_ensemble(fibs: np.ndarray, fib_ok: Optional[np.ndarray] = None) -> dict:
    """Ensemble and service summary from decoded FIBs (what RadioControl gets as dabsdrNtfEnsemble_t + the service list)."""
    fibs = np.ascontiguousarray(fibs, dtype=np.uint8).reshape(-1, 32)
    ok = None if fib_ok is None else np.ascontiguousarray(fib_ok, dtype=np.uint8)
    e = EnsembleInfo()
    if lib().b200dab_parse_ensemble(fibs.ctypes.data, None if ok is None else ok.ctypes.data, fibs.shape[0], ct.byref(e)) != 0:
        raise RuntimeError("parse_ensemble failed")
    svcs = [dict(sid=s.sid, label=s.label.decode("latin1"), components=[(int(s.subch_id[i]), int(s.ascty[i])) for i in range(s.n_components)])
            for s in (e.services[i] for i in range(e.n_services))]
    return dict(ueid=e.ueid, eid=e.ueid & 0xFFFF, ecc=(e.ueid >> 16) & 0xFF, lto=e.lto, int_table=e.int_table, alarm=e.alarm,
                label=e.label.decode("latin1"), first_cif_count=e.first_cif_count, services=svcs)


@dataclass
class AccessUnit:
    subch_index: int
    superframe: int
    cif: int
    crc_ok: bool
    header: int
    data: bytes


@dataclass
class TiiWindow:
    ids: list            # [(main, sub, level)]
    n_tii_frames: int
    spectrum: np.ndarray
    norm: float
    noise: float


@dataclass
class Decoded:
    n_candidates: int
    n_frames: int
    n_slots: int
    t_f: np.ndarray
    cfo_hz: np.ndarray
    cir_ratio: np.ndarray
    fibs: np.ndarray          # [n_slots*12, 32]
    fib_ok: np.ndarray        # [n_slots*12]
    subch: List[SubCh]
    frames: List[np.ndarray]  # per sub-channel [n_rows, 24*bitrate/8]
    aus: List[AccessUnit]
    rs_corrected: int
    rs_bit_errors: int
    rs_uncorrectable: int
    firecode_fail: int
    superframes: int
    tii: List[TiiWindow]
    snr_db: float
    gpu_launches: int
    # DAB+ superframe layer and per-frame TII material (what a chunked decode is merged from, shard.decode_chunked)
    sf: Optional[np.ndarray] = None          # [n_sf, 8] int32: subch_index, row, firecode_ok, rs_corrected, rs_bit_errors, rs_uncorrectable, n_aus, header
    sf_flags: Optional[list] = None          # per sub-channel: uint8 [rows] fire-code hit before RS, or None
    sf_phase: Optional[list] = None          # per sub-channel: row % 5 of the superframe starts, -1 = none
    tii_fold: Optional[np.ndarray] = None    # [n_slots, 384]
    tii_norm: Optional[np.ndarray] = None    # [n_slots]
    tii_sign: Optional[np.ndarray] = None    # [n_slots]


def _ptr(a):
    return a.ctypes.data if isinstance(a, np.ndarray) else int(a)


class Receiver:
    """One CUDA context (stream + buffers) on one GPU."""

    def __init__(self, device: int = 0):
        self._h = ct.c_void_p()
        rc = lib().b200dab_create(ct.byref(self._h), device)
        if rc != 0:
            raise RuntimeError(f"b200dab_create failed ({rc}): a CUDA device is required, there is no CPU fallback")
        self.device = device

    def close(self):
        if self._h:
            lib().b200dab_destroy(self._h)
            self._h = ct.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self, what):
        msg = lib().b200dab_last_error(self._h)
        return RuntimeError(f"{what}: {msg.decode() if msg else '?'}")

    @property
    def stream(self) -> int:
        return int(lib().b200dab_stream(self._h) or 0)

    def set_profiling(self, on: bool) -> None:
        lib().b200dab_set_profiling(self._h, int(on))

    def set_viterbi_split(self, min_steps_four_lane: int) -> None:
        """codewords of at least this many trellis steps use the four-lane kernel mode (0 = all, 1 << 30 = none)"""
        lib().b200dab_set_viterbi_split(self._h, int(min_steps_four_lane))

    def kernel_times(self, back: int = 0):
        """[(kernel name, ms)] of the last decode, or of the decode `back` calls ago (the last 8 are kept); needs
        set_profiling(True) before them."""
        names = ct.create_string_buffer(32768)
        ms = np.zeros(1024, np.float32)
        n = lib().b200dab_get_kernel_times_back(self._h, back, names, 32768, ms.ctypes.data, 1024)
        nm = names.value.decode().split(";") if n > 0 else []
        tot = {}
        for k, v in zip(nm, ms[:n].tolist()):      # a kernel may be launched several times per decode (chunked front end)
            tot[k] = tot.get(k, 0.0) + v
        return list(tot.items())

    def kernel_timeline(self, back: int = 0):
        """[(kernel name, start ms after the first kernel of that decode, duration ms)] in launch order."""
        names = ct.create_string_buffer(32768)
        ms = np.zeros(1024, np.float32)
        st = np.zeros(1024, np.float32)
        n = lib().b200dab_get_kernel_timeline(self._h, back, names, 32768, st.ctypes.data, ms.ctypes.data, 1024)
        nm = names.value.decode().split(";") if n > 0 else []
        return [(k, float(a), float(b)) for k, a, b in zip(nm, st[:n], ms[:n])]

    # ------------------------------------------------------------------ whole path
    def decode_raw(self, ptr: int, n_samples: int, fmt: str, on_device: bool, subch: Optional[Sequence[SubCh]] = None,
                   flags: int = 0) -> Result:
        """Thin call: returns the ctypes Result (pointers into context-owned pinned memory)."""
        res = Result()
        arr = (SubCh * len(subch))(*subch) if subch else None
        rc = lib().b200dab_decode(self._h, ptr, n_samples, FMT[fmt], int(on_device), arr, len(subch) if subch else 0, flags,
                                  ct.byref(res))
        if rc != 0:
            raise self._err("b200dab_decode")
        return res

    def decode(self, raw, fmt: str, subch: Optional[Sequence[SubCh]] = None, tii: bool = False, symstat: bool = False,
               no_msc: bool = False, n_samples: Optional[int] = None, on_device: bool = False,
               sf_phase: Optional[Sequence[int]] = None) -> Decoded:
        """raw: numpy array (host) or an integer device pointer with n_samples given.  sf_phase: forced superframe alignment per
        sub-channel (row % 5, -1 = vote) for this call (chunked decoding)."""
        if sf_phase is not None:
            ph = np.ascontiguousarray(sf_phase, np.int32)
            lib().b200dab_set_superframe_phase(self._h, ph.ctypes.data, ph.size)
        if isinstance(raw, np.ndarray):
            raw = np.ascontiguousarray(raw)
            n_samples = raw.nbytes // FMT_BYTES[fmt]
            ptr = raw.ctypes.data
        else:
            ptr = int(raw)
        flags = (OPT_TII if tii else 0) | (OPT_SYMSTAT if symstat else 0) | (OPT_NO_MSC if no_msc else 0)
        r = self.decode_raw(ptr, n_samples, fmt, on_device, subch, flags)
        ns = r.n_slots
        t_f = np.ctypeslib.as_array(r.t_f, (ns,)).copy() if ns else np.zeros(0, np.int64)
        cfo = np.ctypeslib.as_array(r.cfo_hz, (ns,)).copy() if ns else np.zeros(0, np.float32)
        cir = np.ctypeslib.as_array(r.cir_ratio, (ns,)).copy() if ns else np.zeros(0, np.float32)
        fibs = np.ctypeslib.as_array(r.fibs, (ns * 12, 32)).copy() if ns and r.fibs else np.zeros((0, 32), np.uint8)
        ok = np.ctypeslib.as_array(r.fib_ok, (ns * 12,)).copy() if ns and r.fib_ok else np.zeros(0, np.uint8)
        scs = [r.subch[i] for i in range(r.n_subch)]
        frames, aus = [], []
        if r.frames:
            for i in range(r.n_subch):
                rows, pitch = r.frame_rows[i], scs[i].bitrate * 3
                frames.append(np.ctypeslib.as_array(r.frames[i], (rows, pitch)).copy())
            for k in range(r.n_aus):
                a = r.aus[k]
                buf = frames[a.subch_index].reshape(-1)
                aus.append(AccessUnit(a.subch_index, a.superframe, a.cif, bool(a.crc_ok), a.header,
                                      bytes(buf[a.offset:a.offset + a.length])))
        tiis = []
        for w in range(r.n_tii):
            t = r.tii[w]
            ids = [(int(t.main_id[i]), int(t.sub_id[i]), float(t.level[i])) for i in range(t.n_ids)]
            tiis.append(TiiWindow(ids, t.n_tii_frames, np.array(t.spectrum[:], dtype=np.float32), t.norm, t.noise))
        dec = Decoded(r.n_candidates, r.n_frames, ns, t_f, cfo, cir, fibs, ok, scs, frames, aus, r.rs_corrected,
                      r.rs_bit_errors, r.rs_uncorrectable, r.firecode_fail, r.superframes, tiis, r.snr_db, r.gpu_launches)
        if r.n_sf > 0 and r.sf:
            dec.sf = np.ctypeslib.as_array(ct.cast(r.sf, ct.POINTER(ct.c_int32)), (r.n_sf, 8)).copy()
        else:
            dec.sf = np.zeros((0, 8), np.int32)
        if r.sf_flags and r.frames:
            dec.sf_flags = [np.ctypeslib.as_array(r.sf_flags[i], (r.frame_rows[i],)).copy() if r.sf_flags[i] else None for i in range(r.n_subch)]
            dec.sf_phase = [int(r.sf_phase[i]) for i in range(r.n_subch)]
        if tii and ns and r.tii_fold:
            dec.tii_fold = np.ctypeslib.as_array(r.tii_fold, (ns, TII_BINS)).copy()
            dec.tii_norm = np.ctypeslib.as_array(r.tii_norm, (ns,)).copy()
            dec.tii_sign = np.ctypeslib.as_array(r.tii_sign, (ns,)).copy()
        return dec

    # ------------------------------------------------------------------ stages (parity tests, profiling)
    def stage_coarse(self, raw: np.ndarray, fmt: str):
        n = raw.nbytes // FMT_BYTES[fmt]
        nblk = n // 32
        P = np.zeros(nblk, np.float32)
        S = np.zeros(nblk, np.float32)
        maxc = nblk // 83 + 2
        cand = np.zeros(maxc, np.int64)
        nc = ct.c_int32()
        if lib().b200dab_stage_coarse(self._h, raw.ctypes.data, n, FMT[fmt], P.ctypes.data, S.ctypes.data, cand.ctypes.data,
                                      ct.byref(nc), maxc) != 0:
            raise self._err("stage_coarse")
        return P, S, cand[:nc.value].copy()

    def stage_fine(self, raw: np.ndarray, fmt: str, cand: np.ndarray):
        n = raw.nbytes // FMT_BYTES[fmt]
        cand = np.ascontiguousarray(cand, np.int64)
        t_f = np.zeros(cand.size, np.int64)
        cfo = np.zeros(cand.size, np.float32)
        met = np.zeros((cand.size, 4), np.float32)
        if lib().b200dab_stage_fine(self._h, raw.ctypes.data, n, FMT[fmt], cand.ctypes.data, cand.size, t_f.ctypes.data,
                                    cfo.ctypes.data, met.ctypes.data) != 0:
            raise self._err("stage_fine")
        return t_f, cfo, met

    def stage_demod(self, raw: np.ndarray, fmt: str, t_f, cfo):
        n = raw.nbytes // FMT_BYTES[fmt]
        t_f = np.ascontiguousarray(t_f, np.int64)
        cfo = np.ascontiguousarray(cfo, np.float32)
        soft = np.zeros((t_f.size, FRAME_SOFT), np.int8)
        if lib().b200dab_stage_demod(self._h, raw.ctypes.data, n, FMT[fmt], t_f.ctypes.data, cfo.ctypes.data, t_f.size,
                                     soft.ctypes.data) != 0:
            raise self._err("stage_demod")
        return soft

    def stage_viterbi(self, soft4: np.ndarray) -> np.ndarray:
        soft4 = np.ascontiguousarray(soft4, np.int8)
        n_cw, m = soft4.shape
        steps = m // 4
        bits = np.zeros((n_cw, steps - 6), np.uint8)
        if lib().b200dab_stage_viterbi(self._h, soft4.ctypes.data, n_cw, steps, bits.ctypes.data) != 0:
            raise self._err("stage_viterbi")
        return bits

    def stage_fic(self, soft: np.ndarray):
        soft = np.ascontiguousarray(soft, np.int8).reshape(-1, FRAME_SOFT)
        nf = soft.shape[0]
        fibs = np.zeros((nf * 12, 32), np.uint8)
        ok = np.zeros(nf * 12, np.uint8)
        if lib().b200dab_stage_fic(self._h, soft.ctypes.data, nf, fibs.ctypes.data, ok.ctypes.data) != 0:
            raise self._err("stage_fic")
        return fibs, ok

    def stage_msc(self, soft: np.ndarray, sc: SubCh) -> np.ndarray:
        soft = np.ascontiguousarray(soft, np.int8).reshape(-1, FRAME_SOFT)
        nf = soft.shape[0]
        out = np.zeros((4 * nf - 15, sc.bitrate * 3), np.uint8)
        if lib().b200dab_stage_msc(self._h, soft.ctypes.data, nf, ct.byref(sc), out.ctypes.data) != 0:
            raise self._err("stage_msc")
        return out

    def stage_dabplus(self, frames: np.ndarray, bitrate: int):
        frames = np.ascontiguousarray(frames, np.uint8).copy()
        n_rows = frames.shape[0]
        max_aus = (n_rows // 5 + 1) * 6
        aus = (Au * max_aus)()
        n = ct.c_int32()
        stats = np.zeros(5, np.int64)
        if lib().b200dab_stage_dabplus(self._h, frames.ctypes.data, n_rows, bitrate, aus, ct.byref(n), max_aus,
                                       stats.ctypes.data) != 0:
            raise self._err("stage_dabplus")
        flat = frames.reshape(-1)
        out = [AccessUnit(0, aus[i].superframe, aus[i].cif, bool(aus[i].crc_ok), aus[i].header,
                          bytes(flat[aus[i].offset:aus[i].offset + aus[i].length])) for i in range(n.value)]
        return out, dict(superframes=int(stats[0]), rs_corrected=int(stats[1]), rs_bit_errors=int(stats[2]),
                         rs_uncorrectable=int(stats[3]), firecode_fail=int(stats[4])), frames

    def stage_spectrum(self, raw: np.ndarray, fmt: str, w0) -> np.ndarray:
        """[len(w0), 2048] |FFT|^2 of the raw samples starting at every w0 (the spectrum callback's arithmetic)"""
        n = raw.nbytes // FMT_BYTES[fmt]
        w0 = np.ascontiguousarray(w0, np.int64)
        out = np.zeros((w0.size, 2048), np.float32)
        if lib().b200dab_stage_spectrum(self._h, raw.ctypes.data, n, FMT[fmt], w0.ctypes.data, w0.size, out.ctypes.data) != 0:
            raise self._err("stage_spectrum")
        return out

    def stage_tii_fold(self, nulls: np.ndarray) -> np.ndarray:
        nulls = np.ascontiguousarray(nulls, np.complex64).reshape(-1, 2048)
        fold = np.zeros((nulls.shape[0], TII_BINS), np.float32)
        if lib().b200dab_stage_tii_fold(self._h, nulls.ctypes.data, nulls.shape[0], fold.ctypes.data) != 0:
            raise self._err("stage_tii_fold")
        return fold

    def stage_tii_detect(self, fold: np.ndarray, norm: np.ndarray, sign: np.ndarray, thr_k: float = 6.0) -> TiiWindow:
        fold = np.ascontiguousarray(fold, np.float32)
        norm = np.ascontiguousarray(norm, np.float32)
        sign = np.ascontiguousarray(sign, np.int32)
        t = Tii()
        if lib().b200dab_stage_tii_detect(self._h, fold.ctypes.data, norm.ctypes.data, sign.ctypes.data, fold.shape[0], thr_k,
                                          ct.byref(t)) != 0:
            raise self._err("stage_tii_detect")
        ids = [(int(t.main_id[i]), int(t.sub_id[i]), float(t.level[i])) for i in range(t.n_ids)]
        return TiiWindow(ids, t.n_tii_frames, np.array(t.spectrum[:], dtype=np.float32), t.norm, t.noise)


class ReceiverPool:
    """Several receiver contexts on one GPU, fed from host threads: while one context runs the back end (Viterbi, RS, TII) of
    recording k, the next context is already uploading and demodulating recording k+1, so the PCIe copy and the compute of
    successive, independent recordings overlap.  This is the batch form of a band scan (one recording per channel,
    reference: src/bandscanbackend.cpp:114-139); results come back in input order."""

    def __init__(self, device: int = 0, depth: int = 2):
        self.rx = [Receiver(device) for _ in range(depth)]

    def close(self):
        for r in self.rx:
            r.close()

    def map_raw(self, items, fn=None):
        """items: sequence of (ptr, n_samples, fmt, on_device, subch, flags).  Calls decode_raw for each item on the next free
        context; `fn(index, receiver, result)` runs right after the call, while the result buffers of that context are valid."""
        import threading
        out = [None] * len(items)
        lock = threading.Lock()
        nxt = [0]

        def work(rx):
            while True:
                with lock:
                    i = nxt[0]
                    nxt[0] += 1
                if i >= len(items):
                    return
                ptr, n, fmt, on_dev, subch, flags = items[i]
                res = rx.decode_raw(ptr, n, fmt, on_dev, subch, flags)
                out[i] = fn(i, rx, res) if fn else (res.n_frames, res.n_aus)

        th = [threading.Thread(target=work, args=(r,)) for r in self.rx]
        for t in th:
            t.start()
        for t in th:
            t.join()
        return out
