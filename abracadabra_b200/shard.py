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


def chunk_recording(n_samples: int, n_chunks: int, overlap_frames: int = CHUNK_OVERLAP_FRAMES) -> List[Tuple[int, int]]:
    """Cut one long recording into `n_chunks` sample ranges [start, stop) that overlap by `overlap_frames` frames
    (config 5: a 1-hour recording over 8 GPUs).  Results in the overlap are de-duplicated by CIF number with merge_chunks()."""
    if n_chunks <= 1:
        return [(0, n_samples)]
    body = -(-n_samples // n_chunks)
    ov = overlap_frames * T_FRAME
    out = []
    for c in range(n_chunks):
        a = max(0, c * body - (ov if c else 0))
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


def bind_to_gpu(device: int) -> list:
    """Pin the calling process to the CPU cores NVML reports as local to `device` (same NUMA node / PCIe root), so that pinned
    host buffers are allocated next to the GPU and launches / host reads do not cross the socket interconnect.  One process per
    GPU is the deployment unit of this path; without this, an 8-rank run loses ~20 % per rank to remote-NUMA latency.
    Returns the CPU list (empty if NVML is unavailable: then nothing is changed)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1 and 64 * w + b < n_cpu]
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return []
