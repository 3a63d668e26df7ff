#!/usr/bin/env python
"""bench.py -- throughput of the DAB mode I receive-and-decode hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torch.distributed.run, one rank per GPU)
    python bench.py --impl reference ...                     (the reference's own CPU core on the box's host cores)
    python bench.py --impl dropin ...                        (the drop-in libdabsdr.so.4 behind the reference's C API, one stream)
    python bench.py --config {1,2,3,4,5} ...                 (BASELINE.json configs[0..4]; default 2 = the metric's configuration)

One step (default, configs[1]) = one pass of the whole hot path (sync -> FFT -> DQPSK -> de-interleave -> Viterbi -> FIB CRC ->
FIG 0/1 discovery -> MSC Viterbi of every sub-channel -> RS/AU -> TII) over one synthetic 60 s full-ensemble recording per GPU
(245.76 MB of u8 IQ, larger than the 126 MB L2, so every step re-reads its input from HBM).
`value`  : recording resident in HBM when the timed region starts.
`e2e`    : same call with the recording in pinned HOST memory: H2D copy of the IQ and D2H of all results inside the region.
Work shards by recording: each rank decodes its own recording, no collective on the data path (scaling = weak).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

FS = 2_048_000
T_FRAME = 196_608
METRIC = "Msamples/s of DAB mode I IQ decoded"
WORKLOADS = {
    1: "10 s synthetic DAB mode I u8 IQ (2.048 Msps), FIC + one DAB+ 96 kbps EEP-3A sub-channel + TII, one recording per GPU",
    2: "60 s synthetic full ensemble (FIC + 12 sub-channels, 854 CU: 8 DAB+ EEP, 4 MP2 UEP), u8 IQ @ 2.048 Msps, one recording per GPU",
    3: "TII batch: 500 recordings x 40 transmission frames (10 000 TII-bearing null symbols), 1-8 transmitters, multipath, CFO +-5 kHz, sharded over the GPUs",
    4: "Band III scan batch: 38 channels (5A-13F) x 10 s u8 recordings, ensemble + TII per channel, sharded round-robin over the GPUs",
    5: "one long s16 full-ensemble recording with CFO drift and an SNR staircase 0-20 dB, cut into overlapping time chunks over the GPUs",
}


def make_recording(seconds: float, seed: int):
    from abracadabra_b200 import synth
    ens = synth.full_ensemble(seed=2000 + 17 * seed)
    n_frames = int(round(seconds * FS / T_FRAME))
    imp = synth.Impairments(snr_db=25.0, cfo_hz=1234.5 + 10 * seed, cfo_drift_hz_per_s=3.0, seed=seed + 1, lead_in=49152)
    raw, res = synth.synthesize(ens, n_frames, "u8", imp, seed=seed + 1, tail=65536)
    return raw, res, n_frames


# ------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        self.proc = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


# ------------------------------------------------------------------ reference / CPU baseline
def _run_pool(cmds, cores):
    """run the command lines `cores` at a time; returns (wall seconds, [last stdout line parsed as JSON])"""
    t0 = time.time()
    pending, running, out = list(cmds), [], []
    while pending or running:
        while pending and len(running) < cores:
            running.append(subprocess.Popen(pending.pop(0), stdout=subprocess.PIPE, text=True))
        for p in list(running):
            if p.poll() is not None:
                lines = p.stdout.read().strip().splitlines()
                try:
                    out.append(json.loads(lines[-1]))
                except Exception:
                    pass
                running.remove(p)
        time.sleep(0.002)
    return time.time() - t0, out


def run_reference_cpu(jobs, cores, keep_streams=False):
    """The reference's own core (oracle/_ref/libdabsdr.so.4.0.1 behind oracle/_ref/ref_harness, the RadioControl-like driver):
    jobs = [(path, fmt, n_services, tii)].  The core decodes one audio service per decoder instance, so a file with n services is
    decoded by n processes (one service each); a second arrangement puts two services on one handle (AUDIO_PRIMARY + AUDIO_SECONDARY,
    dabsdr.h:52-53) and needs half the processes.  Both are timed with `cores` processes at a time; the faster one is reported.
    keep_streams: one more, untimed pass with event logs; returns {(path, sid): (ascty, [bytes of every audio callback])}."""
    import orc
    exe = str(orc.REF_DIR / "ref_harness")
    n_samples = sum(os.path.getsize(p) // orc.FMT_BYTES[fmt] for p, fmt, _, _ in jobs)

    def cmds(per_handle, log=None):
        out = []
        for p, fmt, ns, tii in jobs:
            base = [exe, p, fmt]
            for s in range(0, max(ns, 1), per_handle):
                c = base + [log(p, s) if log else "-"] + (["--service", str(s)] if ns else [])
                if per_handle == 2 and s + 1 < ns:
                    c += ["--second", str(s + 1)]
                c += ["--tii", "1" if tii else "0", "--tail-frames", "2"] + ([] if log else ["--free-run"])
                out.append(c)
        return out

    res = {}
    for per_handle in (1, 2):
        if per_handle == 2 and not any(ns > 1 for _, _, ns, _ in jobs):
            continue
        c = cmds(per_handle)
        wall, per = _run_pool(c, cores)
        res[per_handle] = dict(wall=wall, procs=len(c), aus=sum(x.get("aus", 0) for x in per),
                               inst=statistics.mean(x["msps"] for x in per) if per else 0.0)
    best = min(res, key=lambda k: res[k]["wall"])
    r = res[best]
    used = min(cores, r["procs"])
    note = "; ".join(f"{k} service(s) per handle: {v['procs']} processes, wall {v['wall']:.2f} s, per-process {v['inst']:.1f} Msps, {v['aus']} AUs"
                     for k, v in res.items())
    out = dict(value=n_samples / r["wall"] / 1e6, unit="Msamples/s", cores=used, kind="reference",
               sample=f"{n_samples / FS:.1f} s of signal in {len(jobs)} file(s), every audio service decoded (libdabsdr 4.0.1, {used} processes at a time; reported: "
                      f"{best} service(s) per handle).  {note}", aus=r["aus"])
    streams = {}
    if keep_streams:
        td = tempfile.mkdtemp(prefix="dab_ref_log_")
        _run_pool(cmds(1, log=lambda p, s: os.path.join(td, f"{os.path.basename(p)}.{s}.jsonl")), cores)
        for p, fmt, ns, _ in jobs:
            for s in range(ns):
                lp = os.path.join(td, f"{os.path.basename(p)}.{s}.jsonl")
                sid, ascty, data = None, None, []
                try:
                    for line in open(lp):
                        if '"ev":"audio"' in line:
                            e = json.loads(line)
                            ascty = e["ascty"]
                            data.append(bytes.fromhex(e["data"]))
                        elif '"ev":"selection"' in line:
                            sid = json.loads(line)["sid"]
                    os.unlink(lp)
                except Exception:
                    pass
                if sid is not None:
                    streams[(p, sid)] = (ascty, data)
        try:
            os.rmdir(td)
        except Exception:
            pass
    return out, streams


def run_port_cpu(raw, seconds):
    """fallback when the reference binary did not travel: the C checker (oracle/daboracle.c), one thread, short sample"""
    import orc
    from abracadabra_b200 import synth
    n = min(raw.size // 2, int(min(seconds, 4.0) * FS))
    scs = [orc.subch_struct(s) for s in synth.full_ensemble().subchannels]
    t0 = time.time()
    orc.decode(raw[:2 * n], "u8", scs)
    wall = time.time() - t0
    return dict(value=n / wall / 1e6, unit="Msamples/s", cores=1, kind="port",
                sample=f"first {n / FS:.1f} s of the workload through oracle/daboracle.c (scalar C, 1 thread)")


def compare_with_reference(dec, res, streams, path):
    """GPU result against what the reference core delivered for the same file: per service, the reference's access units (DAB+)
    / audio frames (MP2) must appear in the GPU's stream byte for byte, in order, as one contiguous run."""
    by_sub = {s.subchid: s.sid for s in res.ens.services}
    ok, n_ref, detail = True, 0, []
    for i, sc in enumerate(dec.subch):
        sid = by_sub.get(sc.subch_id)
        got = streams.get((path, sid))
        if got is None or not got[1]:
            ok = False
            detail.append(f"subch {sc.subch_id}: no reference stream")
            continue
        ref = got[1]
        n_ref += len(ref)
        # The reference flushes its time de-interleaver with the zero-filled tail of the file and so delivers up to 15 logical frames
        # (three superframes) more than the 4 n - 15 rows this path decodes from n whole frames: its stream may run past the END of ours.
        if sc.kind == 1:
            mine = [a.data for a in dec.aus if a.subch_index == i and a.crc_ok]
            try:
                i0 = mine.index(ref[0])
            except ValueError:
                i0 = -1
            k = min(len(ref), len(mine) - i0) if i0 >= 0 else 0
            good = i0 >= 0 and mine[i0:i0 + k] == ref[:k] and len(ref) - k <= 18 and k > 0
        else:
            mine = dec.frames[i].tobytes()
            flen = sc.bitrate * 3
            head = b"".join(ref[:max(len(ref) - 15, 1)])
            at = mine.find(head)
            rest = b"".join(ref)[len(head):]
            have = mine[at + len(head):at + len(head) + len(rest)] if at >= 0 else b""
            good = at >= 0 and at % flen == 0 and rest[:len(have)] == have and (len(have) == len(rest) or at + len(head) + len(have) == len(mine))
        if not good:
            ok = False
            detail.append(f"subch {sc.subch_id}: streams differ")
    return ok, n_ref, detail


# ------------------------------------------------------------------ algorithmic bytes per kernel
def algorithmic_bytes(name: str, n_frames: int, subch, n_samples: int) -> float:
    n_cif = 4 * n_frames
    rows = max(n_cif - 15, 0)
    punct = sum(s.size_cu * 64 for s in subch)
    info = sum(s.bitrate * 24 for s in subch)
    steps = sum(s.bitrate * 24 + 6 for s in subch)
    if name == "k_block_power":
        return 2.0 * n_samples + 4.0 * n_samples / 32
    if name == "k_coarse":
        return 2 * 4.0 * n_samples / 32
    if name == "k_fine_sync":
        return n_frames * (76 * 184 * 2 * 2 + 2048 * 2) * 1.0
    if name == "k_ofdm_demod":
        return n_frames * (76 * 2048 * 2 + 4 * 2048 * 2 + 75 * 3072) * 1.0   # u8 IQ of every useful part (+1 reference symbol per run) + soft bits
    if name == "k_tii_fold":
        return n_frames * (2 * 2048 * 2 + 384 * 4) * 1.0
    if name.endswith("(early fic)") or name == "k_fib_crc(early)":      # the same FIC kernels on the first 16 frames only
        base = name.replace("(early fic)", "(fic)")
        return algorithmic_bytes(base, min(n_frames, 16), subch, n_samples) if base != name else 0.0
    if name == "k_gather(fic)":
        return n_cif * (2304 + 774 * 4) * 1.0
    if name == "k_viterbi(fic)":
        return n_cif * (774 * 4 + 774 * 8 + 2 * 64 * 2) * 1.0
    if name == "k_vit_trace(fic)":
        return n_cif * (774 * 8 + 2 * 64 * 2 + 96) * 1.0
    if name == "k_gather(msc)":
        return rows * (punct + steps * 4) * 1.0
    if name == "k_viterbi(msc)":
        return rows * (steps * 4 + steps * 8 + len(subch) * 2 * 64 * 2) * 1.0  # soft words in, decisions + meeting metrics out
    if name == "k_vit_trace(msc)":
        return rows * (steps * 8 + len(subch) * 2 * 64 * 2 + info / 8) * 1.0  # decisions + meeting metrics in, bytes out
    if name == "k_sf_decode":        # every byte of the DAB+ logical frames read once
        return rows * info / 8 * 1.0
    return 0.0


def viterbi_issue_model():
    """warp instructions per trellis step of one k_viterbi warp, counted in the SASS of the binary that runs (tools/sass_loops.py
    writes abracadabra_b200/lib/vit_instr.json at build time): {"two_lane": per 32 codewords, "four_lane": per 16 codewords}"""
    try:
        return json.load(open(ROOT / "abracadabra_b200" / "lib" / "vit_instr.json"))
    except Exception:
        return None


# ------------------------------------------------------------------ torch / distributed plumbing of the GPU arms
class Gpu:
    def __init__(self, args):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        self.dev = self.local_rank if self.world > 1 else 0
        torch.cuda.set_device(self.dev)
        from abracadabra_b200 import shard
        self.affinity0 = os.sched_getaffinity(0)
        # before any pinned allocation: NUMA-local host buffers and launch threads; ranks sharing a local core set get disjoint slices
        self.cpus = shard.bind_to_gpu(self.dev, self.world)

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, *vals):
        if self.dist is None:
            return list(vals), [[v] for v in vals]
        t = self.torch.tensor(list(vals), device=f"cuda:{self.dev}", dtype=self.torch.float64)
        every = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(every, t)
        per = [[float(e[i]) for e in every] for i in range(len(vals))]
        return [max(p) for p in per], per

    def sum_over_ranks(self, v):
        if self.dist is None:
            return v
        t = self.torch.tensor([float(v)], device=f"cuda:{self.dev}", dtype=self.torch.float64)
        self.dist.all_reduce(t)
        return float(t[0])

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()

    def timed(self, fn, k):
        """device time (CUDA events on the default stream around host-synchronous calls) and wall time of k calls, max over ranks"""
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        t0 = time.perf_counter()
        e0.record()
        last = None
        for _ in range(k):
            last = fn()
        torch.cuda.synchronize(self.dev)
        e1.record()
        e1.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        ms = e0.elapsed_time(e1)
        self.barrier()
        (ms_max, wall_max), per = self.max_over_ranks(ms, wall)
        return ms_max, wall_max, [round(x / k, 4) for x in per[0]], last

    def h2d_ceiling(self, host_tensor, dev_tensor, k):
        """raw pinned host-to-device copy of the step's input on every rank at the same time: the floor of an end-to-end step"""
        torch = self.torch
        for _ in range(2):
            dev_tensor.copy_(host_tensor, non_blocking=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for _ in range(k):
            dev_tensor.copy_(host_tensor, non_blocking=True)
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / k
        self.barrier()
        (m,), _ = self.max_over_ranks(ms)
        return m


def peak_hbm():
    try:
        return float(json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def base_line(args, g, config, value, ms_per_step, scaling="weak"):
    return {"metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": g.world if g else max(args.gpus, 1), "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "int16 path metrics / f32 FFT",
            "data": "synthetic", "config": config, "x_realtime": value * 1e6 / FS}


# ================================================================== config 2 (default): 60 s full ensemble per GPU
def run_config2(args):
    from abracadabra_b200 import pipeline
    g = Gpu(args)
    torch, dev, rank, world = g.torch, g.dev, g.rank, g.world
    config = {"workload": WORKLOADS[2].replace("60 s", f"{args.seconds:.0f} s"), "recording_bytes": None,
              "l2_policy": "input (245.76 MB for 60 s) larger than the 126 MB L2; re-read from HBM every step",
              "format": "u8", "tii": True,
              "subchannel_discovery": "FIG 0/1 + 0/2 parsed from the decoded FIC inside the step (first from the FIC of the first 16 frames, decoded under the demodulation of the rest; the whole FIC is decoded next to the MSC chain and confirms the table)",
              "cpu_affinity": f"{len(g.cpus)} cores of the GPU's NUMA-local set, disjoint per rank (NVML)" if g.cpus else "unchanged"}
    raw, res, n_frames = make_recording(args.seconds, rank)
    n_samples = raw.size // 2
    config["recording_bytes"] = int(raw.nbytes)
    rx = pipeline.Receiver(dev)
    d_raw = torch.from_numpy(raw).to(f"cuda:{dev}")
    h_raw = torch.from_numpy(raw).pin_memory()
    flags = pipeline.OPT_TII

    def step_device():
        return rx.decode_raw(d_raw.data_ptr(), n_samples, "u8", True, None, flags)

    def step_host():
        return rx.decode_raw(h_raw.data_ptr(), n_samples, "u8", False, None, flags)

    # correctness gate on this very workload: every FIB passes CRC, every sub-channel equals the modulator's payload, and (below,
    # rank 0 at N = 1) every access unit / audio frame the reference core delivers for this file is in the GPU's stream
    chk = rx.decode(d_raw.data_ptr(), "u8", tii=True, n_samples=n_samples, on_device=True)
    ok_fib = bool(chk.fib_ok.all()) and chk.n_frames == n_frames
    ok_msc = len(chk.subch) == len(res.ens.subchannels) and all(
        np.array_equal(chk.frames[i], res.logical[chk.subch[i].subch_id][:chk.frames[i].shape[0]]) for i in range(len(chk.subch)))
    n_aus, n_aus_ok = len(chk.aus), sum(a.crc_ok for a in chk.aus)
    tii_ids = [(m, s) for m, s, _ in chk.tii[0].ids] if chk.tii else []
    ok_tii = tii_ids == [(t.main, t.sub) for t in res.ens.tii]
    subch = chk.subch

    for _ in range(args.warmup):
        step_device()
    clocks = ClockSampler(dev)
    rx.set_profiling(False)
    ms_dev, wall_dev, per_rank_dev, last = g.timed(step_device, args.steps)
    launches = last.gpu_launches
    d2h = int(last.n_slots * 12 * 33 + sum(last.frame_rows[i] * subch[i].bitrate * 3 for i in range(last.n_subch))
              + last.n_tii * ct_sizeof_tii() + last.n_slots * 32)
    # per-kernel times: the same steps again with the profiling events on (they cost ~2 % of a step, so they stay out of `value`)
    # In the timed steps k_vit_trace(msc) runs BESIDE k_viterbi(msc) as its programmatic dependent launch (no event can be recorded
    # between the two without serialising them); for the stage table a second receiver launches it behind k_viterbi (B200DAB_TRACE_PDL=0),
    # so that the dominant kernel is timed on its own.  `value` above comes from the default receiver.
    os.environ["B200DAB_TRACE_PDL"] = "0"
    rx_p = pipeline.Receiver(dev)
    os.environ.pop("B200DAB_TRACE_PDL", None)
    rx_p.set_profiling(True)
    ktimes = {}
    for _ in range(min(args.steps, 8) + 2):
        rx_p.decode_raw(d_raw.data_ptr(), n_samples, "u8", True, None, flags)
    for back in range(min(args.steps, 8)):
        for nm, v in rx_p.kernel_times(back):
            ktimes.setdefault(nm, []).append(v)
    rx_p.close()
    # two receiver contexts on the GPU, recordings resident in HBM: the front end of recording k + 1 fills the tail of recording k
    pool = pipeline.ReceiverPool(dev, depth=2)
    summ = lambda i, r, res_: (res_.n_frames, res_.n_aus, int(res_.fib_ok[0]))
    item_dev = (d_raw.data_ptr(), n_samples, "u8", True, None, flags)
    pool.map_raw([item_dev] * max(2, args.warmup), summ)
    ms_pool, _, _, got = g.timed(lambda: pool.map_raw([item_dev] * args.steps, summ), 1)
    assert all(x[0] == n_frames for x in got)
    # end to end: host buffers in, host results out
    for _ in range(max(1, args.warmup // 2)):
        step_host()
    ms_seq, wall_seq, _, _ = g.timed(step_host, args.steps)
    item = (h_raw.data_ptr(), n_samples, "u8", False, None, flags)
    pool.map_raw([item] * max(2, args.warmup), summ)
    ms_e2e, _, per_rank_e2e, got = g.timed(lambda: pool.map_raw([item] * args.steps, summ), 1)
    assert all(x[0] == n_frames for x in got)
    pool.close()
    ms_h2d = g.h2d_ceiling(h_raw, d_raw, max(args.steps, 5))
    clk = clocks.stop()

    total_samples = n_samples * world
    value = total_samples * args.steps / (ms_dev * 1e-3) / 1e6
    value_pool = total_samples * args.steps / (ms_pool * 1e-3) / 1e6
    e2e_v = total_samples * args.steps / (ms_e2e * 1e-3) / 1e6

    peak, peak_src = peak_hbm()
    stages = []
    for nm, v in ktimes.items():
        avg = statistics.mean(v)
        ab = algorithmic_bytes(nm, n_frames, subch, n_samples)
        stages.append({"kernel": nm, "ms": round(avg, 4), "algorithmic_bytes": ab, "gbs": round(ab / (avg * 1e-3) / 1e9, 2) if avg > 0 else None,
                       "frac_hbm": round(ab / (avg * 1e-3) / 1e9 / peak, 4) if avg > 0 else None})
    stages.sort(key=lambda s: -s["ms"])
    traffic, traffic_src = {}, None
    try:   # DRAM bytes per launch from the committed ncu --set full captures of this round's binary (profiles/)
        tj = json.load(open(ROOT / "profiles" / "r02_traffic.json"))   # written by tools/ncu_summary.py traffic from this round's capture
        traffic, traffic_src = tj["bytes_per_launch"], tj.get("source")
    except Exception:
        pass
    for st in stages:
        st["dram_traffic_ncu"] = traffic.get(st["kernel"])
    dom = stages[0] if stages else None
    roof = None
    sm_mhz = clk.get("sm_mhz") or 1965.0
    if dom and "viterbi" in dom["kernel"]:
        # the dominant kernel is bound by instruction issue / the ALU pipe, not by HBM: achieved = warp
        # instructions per second, instruction count = trellis steps x the per-step count of the SASS that ran, peak = one warp
        # instruction per cycle on each of the 148 x 4 sub-partitions at the SM clock sampled during the run
        rows = max(4 * n_frames - 15, 0)
        vm = viterbi_issue_model() or {}
        per_cw_step = (vm.get("four_lane", 111.0) / 16.0)
        n_instr = rows * sum(s.bitrate * 24 + 6 for s in subch) * per_cw_step
        peak_issue = 148 * 4 * sm_mhz * 1e6
        ach = n_instr / (dom["ms"] * 1e-3)
        roof = {"kernel": dom["kernel"], "bound": "issue", "achieved": round(ach / 1e9, 2), "peak": round(peak_issue / 1e9, 2), "unit": "Gwarp-instr/s",
                "frac": round(ach / peak_issue, 4), "traffic": traffic.get(dom["kernel"]) if abs(args.seconds - 60.0) < 1e-6 else None,
                "traffic_source": traffic_src, "ms": dom["ms"], "warp_instructions": n_instr, "instr_per_16_codeword_step": vm.get("four_lane"),
                "acs_per_s": rows * sum(s.bitrate * 24 + 6 for s in subch) * 64.0 / (dom["ms"] * 1e-3),
                "hbm": {"achieved": dom["gbs"], "peak": peak, "unit": "GB/s", "frac": dom["frac_hbm"], "algorithmic_bytes": dom["algorithmic_bytes"]},
                "note": "peak = 1 warp instruction per cycle per sub-partition; this instruction mix tops out near 0.75 of it (tools/ubench), the ALU pipe alone bounds it at 0.85"}
    elif dom:
        roof = {"kernel": dom["kernel"], "bound": "hbm", "achieved": dom["gbs"], "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": dom["frac_hbm"],
                "traffic": traffic.get(dom["kernel"]), "ms": dom["ms"], "algorithmic_bytes": dom["algorithmic_bytes"]}
    hbm_dom = next((s for s in stages if s["kernel"] in ("k_ofdm_demod", "k_vit_trace(msc)", "k_gather(msc)", "k_block_power")), None)

    line = base_line(args, g, config, value, ms_dev / args.steps)
    line.update({"x_realtime_per_gpu": value * 1e6 / FS / world,
                 "value_pool": value_pool, "value_pool_api": "ReceiverPool(depth=2) on device-resident recordings: two contexts per GPU, the next recording's front end runs in the tail of the previous one",
                 "e2e": {"value": e2e_v, "unit": "Msamples/s", "h2d_bytes_per_step": int(raw.nbytes), "d2h_bytes_per_step": d2h,
                         "ms_per_step": ms_e2e / args.steps, "x_realtime": e2e_v * 1e6 / FS,
                         "api": "ReceiverPool(depth=2).map_raw: two receiver contexts per GPU, steps pipelined (upload of step k+1 under the back end of step k)",
                         "sequential_value": total_samples * args.steps / (ms_seq * 1e-3) / 1e6, "sequential_ms_per_step": ms_seq / args.steps,
                         "h2d_ceiling_ms": round(ms_h2d, 4), "h2d_ceiling_gbs_per_gpu": round(raw.nbytes / (ms_h2d * 1e-3) / 1e9, 2),
                         "frac_of_h2d_ceiling": round(ms_h2d / (ms_e2e / args.steps), 4),
                         "h2d_ceiling_note": f"the same {raw.nbytes / 1e6:.0f} MB pinned-host-to-device copy alone, on all {world} ranks at the same time (max over ranks)"},
                 "gpu_launches": int(launches) * args.steps, "launches_per_step": int(launches), "clocks": clk,
                 "wall_ms_per_step": wall_dev / args.steps, "per_rank_ms_per_step": per_rank_dev, "per_rank_e2e_ms_per_step": [round(x / args.steps, 4) for x in per_rank_e2e] if per_rank_e2e else None,
                 "roofline": roof,
                 "roofline_hbm": ({"kernel": hbm_dom["kernel"], "bound": "hbm", "achieved": hbm_dom["gbs"], "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                                   "frac": hbm_dom["frac_hbm"], "traffic": hbm_dom["dram_traffic_ncu"], "ms": hbm_dom["ms"]} if hbm_dom else None),
                 "stages": stages,
                 "stages_note": "per-kernel CUDA-event times from a second receiver that launches k_vit_trace(msc) behind k_viterbi(msc) (B200DAB_TRACE_PDL=0): in the timed steps the trace runs beside the recursions as a programmatic dependent launch; k_vit_trace(msc) here = sum of its two launches (DAB/data and DAB+ halves), which run side by side on two streams; k_ofdm_demod = sum of its two launches (first 16 frames on the FIC stream, the rest on the main stream, side by side)",
                 "check": {"frames": n_frames, "all_fib_crc_ok": ok_fib, "all_subchannels_equal_payload": ok_msc, "aus": n_aus, "aus_crc_ok": int(n_aus_ok),
                           "tii_ids": tii_ids, "tii_ids_as_transmitted": ok_tii}})
    failed = not (ok_fib and ok_msc and ok_tii and n_aus == n_aus_ok and n_aus > 0)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, g.affinity0)            # the reference gets every host core again
        import orc
        cores = os.cpu_count() or 1
        if orc.have_ref():
            path = tempfile.mktemp(prefix="dab_ref_", suffix=".u8")
            n = n_samples if args.cpu_sample_seconds >= args.seconds else min(n_samples, int(args.cpu_sample_seconds * FS))
            raw[:2 * n].tofile(path)
            base, streams = run_reference_cpu([(path, "u8", len(res.ens.services), True)], cores, keep_streams=True)
            os.unlink(path)
            same, n_ref, detail = compare_with_reference(chk, res, streams, path)
            line["check"].update({"reference_aus_equal": same, "reference_aus": n_ref,
                                  "reference_aus_note": "every access unit (DAB+) / audio frame (MP2) the reference core delivered for this file, per service, found byte for byte and in order in the GPU result" + ("" if same else ": " + "; ".join(detail))})
            failed = failed or not same
            base.pop("aus", None)
        else:
            base = run_port_cpu(raw, args.cpu_sample_seconds)
        line["cpu_baseline"] = base
    if rank == 0:
        print(json.dumps(line))
    g.close()
    rx.close()
    if failed:
        print("bench.py: the correctness checks of the workload failed -- the numbers above are not valid", file=sys.stderr)
        raise SystemExit(1)


def ct_sizeof_tii():
    import ctypes as ct
    from abracadabra_b200 import pipeline
    return ct.sizeof(pipeline.Tii)


# ================================================================== the other configurations (bench_configs.py keeps main() short)
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "dropin"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5])
    ap.add_argument("--seconds", type=float, default=None, help="length of the synthetic recording (config 2: per GPU, default 60; config 5: total, default 3600)")
    ap.add_argument("--cpu-sample-seconds", type=float, default=60.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chunks", type=int, default=0, help="config 5: number of time chunks (default max(8, gpus))")
    args = ap.parse_args()
    if args.impl == "b200":
        args.warmup = max(args.warmup, 3)
    if args.seconds is None:
        args.seconds = {1: 10.0, 2: 60.0, 3: 0.0, 4: 10.0, 5: 3600.0}[args.config]
    import bench_configs
    if args.impl == "reference":
        return bench_configs.run_reference_arm(args)
    if args.impl == "dropin":
        return bench_configs.run_dropin_arm(args)
    if args.config == 2:
        return run_config2(args)
    return {1: bench_configs.run_config1, 3: bench_configs.run_config3, 4: bench_configs.run_config4, 5: bench_configs.run_config5}[args.config](args)


if __name__ == "__main__":
    main()
