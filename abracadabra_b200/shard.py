"""Multi-GPU partitioning of the receive path: work shards by recording / channel / time chunk, one process per GPU,
no collective on the data path -- only a host-side gather of the (small) results (SURVEY.md sec. 8 e).

The reference has no counterpart (one dabsdr thread per handle, src/application.cpp:476-558); what this mirrors is its
band scan, which walks DabTables::channelList one channel at a time (src/bandscanbackend.cpp:114-139).
"""
from __future__ import annotations

from typing import Any, Callable, List, Sequence, Tuple

FS = 2_048_000
T_FRAME = 196_608
# warm-up a chunk needs before its first deliverable CIF: 2 frames sync margin + 15 CIFs of time de-interleaver + 5 CIFs superframe
CHUNK_OVERLAP_FRAMES = 8


def assign(n_units: int, rank: int, world: int) -> List[int]:
    """Static round-robin: unit u goes to rank u % world (config 4: 38 channel files over 1/2/4/8 GPUs)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_units, world))


CHUNK_ALIGN = 32768          # chunk starts are multiples of this: every sync kernel sees the same block grid as in a single pass


def chunk_recording(n_samples: int, n_chunks: int, overlap_frames: int = CHUNK_OVERLAP_FRAMES) -> List[Tuple[int, int]]:
    """Cut one long recording into `n_chunks` sample ranges [start, stop) that overlap by `overlap_frames` frames
    (config 5: a 1-hour recording over 8 GPUs).  decode_chunked() puts the results together again."""
    if n_chunks <= 1:
        return [(0, n_samples)]
    body = -(-n_samples // n_chunks)
    ov = overlap_frames * T_FRAME
    out = []
    for c in range(n_chunks):
        a = max(0, c * body - (ov if c else 0)) // CHUNK_ALIGN * CHUNK_ALIGN
        b = min(n_samples, (c + 1) * body + T_FRAME)
        if a < b:
            out.append((a, b))
    return out


def merge_chunks(chunks: Sequence[Sequence[Tuple[int, Any]]]) -> List[Tuple[int, Any]]:
    """chunks[i] = [(absolute CIF number, item), ...] in ascending order; keeps the first occurrence of every CIF number."""
    seen = set()
    out = []
    for ch in chunks:
        for cif, item in ch:
            if cif in seen:
                continue
            seen.add(cif)
            out.append((cif, item))
    out.sort(key=lambda x: x[0])
    return out


# ------------------------------------------------------------------ one recording in time chunks (SURVEY.md sec. 8 e, config 5)
def _subch_key(scs) -> bytes:
    import ctypes as ct
    return b"".join(bytes(ct.string_at(ct.addressof(s), ct.sizeof(s))) for s in scs)


def chain_frames(t_abs_per_chunk) -> Tuple[list, List[Any]]:
    """Chain the frames of overlapping chunks exactly as the front end chains the frames of one pass (b200dab.cu: front_end):
    a frame is accepted when it starts at least 0.9 frame periods after the previous one, and slot numbers advance by the
    rounded distance.  t_abs_per_chunk[c][k] = absolute FFT-window start of the frame in local slot k of chunk c (-1 = lost).
    Returns ([(chunk, local slot, global slot)], [global slot of every chunk's local slot 0, None for a chunk without frames])."""
    chain, slot0 = [], []                  # chain entries: (chunk, local slot, global slot, t)
    for c, t_abs in enumerate(t_abs_per_chunk):
        anchor = None
        recent = {e[3]: e[2] for e in chain[-64:]}
        for k, t in enumerate(t_abs):
            t = int(t)
            if t < 0:
                continue
            if chain and t - chain[-1][3] < T_FRAME * 9 // 10:
                if anchor is None and t in recent:                  # the same frame as one already chained: fixes the numbering
                    anchor = recent[t] - k
                continue
            g = 0 if not chain else chain[-1][2] + int((t - chain[-1][3] + T_FRAME // 2) // T_FRAME)
            if anchor is None:
                anchor = g - k
            if g != anchor + k:
                raise RuntimeError(f"chunk {c}: frame numbering of the chunk (slot {k}) and of the recording (slot {g}) disagree")
            chain.append((c, k, g, t))
        slot0.append(anchor)
    return [(c, k, g) for c, k, g, _ in chain], slot0


def decode_chunked(iq, fmt: str, n_chunks: int, rx=None, device: int = 0, tii: bool = False, rank: int = 0, world: int = 1,
                   dist=None, overlap_frames: int = CHUNK_OVERLAP_FRAMES, stats: dict = None):
    """Decode ONE recording cut into `n_chunks` overlapping time chunks (each a separate b200dab_decode, on this rank's GPU for
    the chunks c % world == rank) and merge the results by absolute frame / CIF position.  The merged result is identical --
    byte for byte -- to decoding the recording in one pass (tests/test_gpu_paths.py::test_chunked_decode_equals_single_pass):

    * frames are chained across chunks with the front end's own rule; a frame found in the overlap by two chunks has the same
      t_f and CFO in both (the sync kernels see the same block grid: chunk starts are multiples of 32768 samples);
    * what a single pass decides globally is decided globally here too: the sub-channel table (FIG 0/1, 0/2 of the whole
      FIC) and the DAB+ superframe alignment (fire-code hits over all logical frames).  Every chunk is
      decoded once with its own view; a chunk whose view differs from the global one (possible only where the signal is too
      noisy to decode anyway) is decoded a second time with the global table / alignment forced;
    * logical frames, superframes, access units and RS statistics are taken from exactly one chunk each (boundaries on
      superframe starts), TII windows are re-detected on the merged per-frame spectra in windows of 40 global frames.

    iq: numpy array (or memmap) of the whole recording, interleaved I/Q in `fmt`.  Returns pipeline.Decoded on rank 0 (None on
    the other ranks).  Mirrors how the reference resumes inside a file (src/input/rawfileinput.cpp:299-343) -- by sample offset,
    with the core re-synchronising -- but without losing the interleaver depth at each cut."""
    import numpy as np
    from . import pipeline
    own_rx = rx is None
    rx = rx or pipeline.Receiver(device)
    bps = pipeline.FMT_BYTES[fmt]
    item = np.dtype({"u8": np.uint8, "s16": np.int16, "f32": np.float32}[fmt]).itemsize
    per = bps // item                                     # array elements per complex sample
    n_samples = iq.size // per
    ranges = chunk_recording(n_samples, n_chunks, overlap_frames)
    mine = [c for c in range(len(ranges)) if c % world == rank]

    def run(c, subch=None, phase=None):
        a, b = ranges[c]
        return rx.decode(np.ascontiguousarray(iq[a * per:b * per]), fmt, subch=subch, tii=tii, sf_phase=phase)

    def gather(obj, everyone=True):
        if world == 1 or dist is None:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    decs = {c: run(c) for c in mine}
    # ---- light summaries to every rank: frame positions, FIC, each chunk's own view of the table and of the alignment
    light = {c: dict(t_abs=np.where(d.t_f >= 0, d.t_f + ranges[c][0], -1), fibs=d.fibs, ok=d.fib_ok, key=_subch_key(d.subch),
                     flags=d.sf_flags, phase=d.sf_phase, rows=[f.shape[0] for f in d.frames]) for c, d in decs.items()}
    every = {}
    for part in gather(light):
        every.update(part)
    order = sorted(every)
    chain, slot0 = chain_frames([every[c]["t_abs"] for c in order])
    n_slots = chain[-1][2] + 1 if chain else 0
    fibs = np.zeros((n_slots * 12, 32), np.uint8)
    ok = np.zeros(n_slots * 12, np.uint8)
    for c, k, g in chain:
        fibs[g * 12:(g + 1) * 12] = every[c]["fibs"][k * 12:(k + 1) * 12]
        ok[g * 12:(g + 1) * 12] = every[c]["ok"][k * 12:(k + 1) * 12]
    # ---- the global sub-channel table: what the whole FIC describes (b200dab_decode's own rule)
    table = pipeline.parse_subch(fibs, ok) if n_slots else []
    gkey = _subch_key(table)
    n_rows = max(4 * n_slots - 15, 0)
    # ---- the global superframe alignment per DAB+ sub-channel: fire-code hits of every logical frame, counted once
    best = [-1] * len(table)
    redo = {}
    for i, sc in enumerate(table):
        if sc.kind != 1 or n_rows <= 0:
            continue
        fl = np.zeros(n_rows, np.uint8)
        for c in order:
            e = every[c]
            if e["key"] != gkey or slot0[order.index(c)] is None or not e["flags"] or e["flags"][i] is None:
                continue
            r0 = 4 * slot0[order.index(c)]
            f = e["flags"][i]
            lo, hi = max(r0, 0), min(r0 + f.size, n_rows)
            if lo < hi:
                fl[lo:hi] = f[lo - r0:hi - r0]
        cnt = [int(fl[k::5].sum()) for k in range(5)]
        if max(cnt) > 0:
            best[i] = int(np.argmax(cnt))                 # first maximum = the strict '>' of k_fire_align
    # a chunk whose own table differs could not contribute its flags above: decode it again first, then settle the alignment
    for c in order:
        if every[c]["key"] != gkey and slot0[order.index(c)] is not None:
            redo[c] = True
    if redo:
        for c in [c for c in redo if c in decs]:
            decs[c] = run(c, subch=table)
        light2 = {c: dict(flags=decs[c].sf_flags, phase=decs[c].sf_phase, rows=[f.shape[0] for f in decs[c].frames]) for c in redo if c in decs}
        for part in gather(light2):
            for c, v in part.items():
                every[c].update(v)
                every[c]["key"] = gkey
        for i, sc in enumerate(table):
            if sc.kind != 1 or n_rows <= 0:
                continue
            fl = np.zeros(n_rows, np.uint8)
            for c in order:
                e = every[c]
                if slot0[order.index(c)] is None or not e["flags"] or e["flags"][i] is None:
                    continue
                r0 = 4 * slot0[order.index(c)]
                f = e["flags"][i]
                lo, hi = max(r0, 0), min(r0 + f.size, n_rows)
                if lo < hi:
                    fl[lo:hi] = f[lo - r0:hi - r0]
            cnt = [int(fl[k::5].sum()) for k in range(5)]
            best[i] = int(np.argmax(cnt)) if max(cnt) > 0 else -1
    n_redo = len(redo)
    for c in mine:
        s0 = slot0[order.index(c)]
        if s0 is None or not table or not decs[c].frames:
            continue
        want = [(best[i] - 4 * s0) % 5 if (table[i].kind == 1 and best[i] >= 0) else -1 for i in range(len(table))]
        have = every[c]["phase"] or [-1] * len(table)
        if any(table[i].kind == 1 and want[i] != have[i] and want[i] >= 0 for i in range(len(table))):
            decs[c] = run(c, subch=table, phase=want)
            n_redo += 1
    if stats is not None:
        stats["chunks"] = len(ranges)
        stats["second_passes_this_rank"] = n_redo
    # ---- heavy results to rank 0
    heavy = {c: decs[c] for c in mine}
    parts = {}
    if world == 1 or dist is None:
        parts = heavy
    else:
        got = [None] * world if rank == 0 else None
        dist.gather_object(_pack_heavy(heavy), got, dst=0)
        if rank == 0:
            for g in got:
                parts.update(_unpack_heavy(g))
    if rank != 0:
        if own_rx:
            rx.close()
        return None
    out = _merge(parts, order, ranges, chain, slot0, n_slots, n_rows, table, best, fibs, ok, tii, rx, pipeline)
    if own_rx:
        rx.close()
    return out


def _pack_heavy(decs: dict) -> dict:
    """pipeline.Decoded holds ctypes structures (not picklable): ship plain arrays"""
    out = {}
    for c, d in decs.items():
        out[c] = dict(t_f=d.t_f, cfo=d.cfo_hz, cir=d.cir_ratio, frames=d.frames, sf=d.sf, n_cand=d.n_candidates,
                      aus=[(a.subch_index, a.superframe, a.cif, a.crc_ok, a.header, a.data) for a in d.aus],
                      tii_fold=d.tii_fold, tii_norm=d.tii_norm, launches=d.gpu_launches)
    return out


def _unpack_heavy(p: dict) -> dict:
    return p


def _merge(parts, order, ranges, chain, slot0, n_slots, n_rows, table, best, fibs, ok, tii, rx, pipeline):
    import numpy as np
    P = {c: (_pack_heavy({c: d})[c] if not isinstance(d, dict) else d) for c, d in parts.items()}
    t_f = np.full(n_slots, -1, np.int64)
    cfo = np.zeros(n_slots, np.float32)
    cir = np.zeros(n_slots, np.float32)
    fold = np.zeros((n_slots, pipeline.TII_BINS), np.float32) if tii else None
    norm = np.zeros(n_slots, np.float32) if tii else None
    for c, k, g in chain:
        p = P[c]
        t_f[g] = p["t_f"][k] + ranges[c][0]
        cfo[g] = p["cfo"][k]
        cir[g] = p["cir"][k]
        if tii and p["tii_fold"] is not None:
            fold[g] = p["tii_fold"][k]
            norm[g] = p["tii_norm"][k]
    frames, aus, sf_rows = [], [], []
    live = [c for c in order if slot0[order.index(c)] is not None and P[c]["frames"]]
    for i, sc in enumerate(table):
        pitch = sc.bitrate * 3
        out = np.zeros((n_rows, pitch), np.uint8)
        dabp = sc.kind == 1 and best[i] >= 0
        # chunk c contributes the rows [B_c, B_{c+1}); boundaries sit on superframe starts of DAB+ sub-channels
        bounds = []
        for c in live:
            r0 = 4 * slot0[order.index(c)]
            b = max(r0, 0)
            if dabp:
                b += (best[i] - b) % 5
            bounds.append(b)
        if bounds:
            bounds[0] = 0
        for n, c in enumerate(live):
            r0 = 4 * slot0[order.index(c)]
            lo = bounds[n]
            hi = bounds[n + 1] if n + 1 < len(live) else n_rows
            f = P[c]["frames"][i]
            hi = min(hi, r0 + f.shape[0], n_rows)
            if lo < hi:
                out[lo:hi] = f[lo - r0:hi - r0]
            if sc.kind == 1:
                for row in P[c]["sf"]:
                    if row[0] == i and lo <= row[1] + r0 < hi and row[1] + r0 + 5 <= n_rows:
                        sf_rows.append((i, row[1] + r0) + tuple(int(x) for x in row[2:]))
                for (si, m, cif, crc_ok, header, data) in P[c]["aus"]:
                    if si == i and lo <= cif + r0 < hi and cif + r0 + 5 <= n_rows:
                        aus.append(pipeline.AccessUnit(i, (cif + r0 - best[i]) // 5, cif + r0, crc_ok, header, data))
        frames.append(out)
    sf = np.array(sf_rows, np.int32).reshape(-1, 8) if sf_rows else np.zeros((0, 8), np.int32)
    tiis = []
    sign = None
    if tii and n_slots:
        sign = np.zeros(n_slots, np.int32)
        for s in range(n_slots):
            if t_f[s] < 0:
                continue
            fib = fibs[s * 12]
            if ok[s * 12] and (fib[0] >> 5) == 0 and (fib[1] & 31) == 0 and (fib[0] & 31) >= 5:
                cifc = (int(fib[4]) & 0x1F) * 250 + int(fib[5])
                sign[s] = 1 if cifc % 8 == 0 else (-1 if cifc % 8 == 4 else 0)
        for s in range(1, n_slots):
            if not sign[s] and sign[s - 1] and t_f[s] >= 0:
                sign[s] = -sign[s - 1]
        for s in range(n_slots - 2, -1, -1):
            if not sign[s] and sign[s + 1] and t_f[s] >= 0:
                sign[s] = -sign[s + 1]
        wins = [(w, w + 40) for w in range(0, n_slots - 39, 40)] or [(0, n_slots & ~1)]
        for a, b in wins:
            if b > a:
                tiis.append(rx.stage_tii_detect(fold[a:b], norm[a:b], sign[a:b]))
    dec = pipeline.Decoded(sum(p["n_cand"] for p in P.values()), len(chain), n_slots, t_f, cfo, cir, fibs, ok, list(table), frames, aus,
                           int(sf[:, 3].sum()), int(sf[:, 4].sum()), int(sf[:, 5].sum()), int((sf[:, 2] == 0).sum()), int(sf.shape[0]),
                           tiis, 0.0, sum(p["launches"] for p in P.values()))
    dec.sf = sf
    dec.tii_fold, dec.tii_norm, dec.tii_sign = fold, norm, sign
    return dec


def run_sharded(units: Sequence[Any], fn: Callable[[Any], Any], rank: int, world: int, dist=None) -> List[Any]:
    """Every rank runs fn on its share; rank 0 receives all results in unit order (host-side gather, no NCCL on the hot path).
    `dist` is torch.distributed (any backend able to all_gather_object: gloo on CPU, nccl on GPUs) or None for world == 1."""
    mine = [(u, fn(units[u])) for u in assign(len(units), rank, world)]
    if world == 1 or dist is None:
        return [r for _, r in mine]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    flat = sorted((x for part in gathered for x in part), key=lambda x: x[0])
    return [r for _, r in flat]


def partition_cpus(local_sets: Sequence[Sequence[int]], device: int) -> List[int]:
    """Cores for the process driving `device` when every visible GPU has one: local_sets[g] = cores NVML reports as local to GPU g.
    GPUs that share one local set (a box that reports `CPU Affinity 0-31` for all eight GPUs) split it into equal contiguous
    slices in device order, so the ranks, their pool threads and the NCCL helper threads never contend for the same cores."""
    mine = sorted(local_sets[device])
    peers = [g for g, cs in enumerate(local_sets) if sorted(cs) == mine]
    k, n = peers.index(device), len(peers)
    if not mine or len(mine) < n:
        return list(mine)
    return mine[k * len(mine) // n:(k + 1) * len(mine) // n]


def bind_to_gpu(device: int, n_devices: int = 1) -> list:
    """Pin the calling process to CPU cores local to `device` (same NUMA node / PCIe root, as NVML reports them), so that pinned
    host buffers are allocated next to the GPU and launches / host reads do not cross the socket interconnect.  One process per
    GPU is the deployment unit of this path: with `n_devices` > 1 processes on the box, the GPUs sharing a local core set get
    disjoint slices of it (partition_cpus).  Returns the CPU list (empty if NVML is unavailable: then nothing is changed)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        n_cpu = os.cpu_count() or 1
        allowed = set(os.sched_getaffinity(0))
        sets = []
        for g in range(max(n_devices, device + 1)):
            h = pynvml.nvmlDeviceGetHandleByIndex(g)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
            cs = [64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1 and 64 * w + b < n_cpu]
            sets.append([c for c in cs if c in allowed])
        cpus = partition_cpus(sets, device) if n_devices > 1 else sets[device]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return []
