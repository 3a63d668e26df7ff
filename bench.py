#!/usr/bin/env python
"""bench.py -- throughput of the DAB mode I receive-and-decode hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torch.distributed.run, one rank per GPU)
    python bench.py --impl reference ...                     (the reference's own CPU core on the box's host cores)

One step = one pass of the whole hot path (sync -> FFT -> DQPSK -> de-interleave -> Viterbi -> FIB CRC -> FIG 0/1 discovery
-> MSC Viterbi of every sub-channel -> RS/AU -> TII) over one synthetic 60 s full-ensemble recording per GPU
(BASELINE.json configs[1]; 245.76 MB of u8 IQ, larger than the 126 MB L2, so every step re-reads its input from HBM).
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
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
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
def run_reference_cpu(raw: np.ndarray, n_services: int, sample_seconds: float, cores: int):
    """Full-ensemble decode of the first `sample_seconds` of the recording by the reference's own core
    (oracle/_ref/libdabsdr.so.4.0.1 through oracle/_ref/ref_harness): the core decodes one audio service per instance,
    so one instance per service, `cores` instances at a time.  Returns dict(value Msps, kind, cores, sample)."""
    import orc
    n = min(raw.size // 2, int(sample_seconds * FS))
    if orc.have_ref():
        path = tempfile.mktemp(prefix="dab_ref_", suffix=".u8")
        raw[:2 * n].tofile(path)
        t0 = time.time()
        procs, done, per = [], 0, []
        pending = list(range(n_services))
        running = []
        exe = str(orc.REF_DIR / "ref_harness")
        while pending or running:
            while pending and len(running) < cores:
                s = pending.pop(0)
                running.append(subprocess.Popen([exe, path, "u8", "-", "--service", str(s), "--tii", "1", "--tail-frames", "2", "--free-run"],
                                                stdout=subprocess.PIPE, text=True))
            for p in list(running):
                if p.poll() is not None:
                    out = p.stdout.read().strip().splitlines()
                    try:
                        per.append(json.loads(out[-1]))
                    except Exception:
                        pass
                    running.remove(p)
            time.sleep(0.002)
        wall = time.time() - t0
        os.unlink(path)
        aus = sum(x.get("aus", 0) for x in per)
        inst = statistics.mean(x["msps"] for x in per) if per else 0.0
        return dict(value=n / wall / 1e6, unit="Msamples/s", cores=min(cores, n_services), kind="reference",
                    sample=f"first {n / FS:.1f} s of the workload, all {n_services} services: one libdabsdr 4.0.1 instance per service "
                           f"(API limit), {min(cores, n_services)} at a time; per-instance {inst:.1f} Msps; {aus} AUs delivered")
    # checker port, single thread, smaller sample
    from abracadabra_b200 import synth
    n = min(n, 4 * FS)
    ens = synth.full_ensemble()
    scs = [orc.subch_struct(s) for s in ens.subchannels]
    t0 = time.time()
    orc.decode(raw[:2 * n], "u8", scs)
    wall = time.time() - t0
    return dict(value=n / wall / 1e6, unit="Msamples/s", cores=1, kind="port",
                sample=f"first {n / FS:.1f} s of the workload through oracle/daboracle.c (scalar C, 1 thread)")


# ------------------------------------------------------------------ algorithmic bytes per kernel
def algorithmic_bytes(name: str, n_frames: int, subch, n_samples: int) -> float:
    n_cif = 4 * n_frames
    rows = max(n_cif - 15, 0)
    punct = sum(s.size_cu * 64 for s in subch)
    info = sum(s.bitrate * 24 for s in subch)
    steps = sum(s.bitrate * 24 + 6 for s in subch)
    if name == "k_block_power":
        return 2.0 * n_samples + 4.0 * n_samples / 32
    if name.startswith("k_window_sum"):
        return 3 * 4.0 * n_samples / 32
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
    if name == "k_rs_decode":
        return rows * info / 8 * 1.0
    return 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--seconds", type=float, default=60.0, help="length of the synthetic recording per GPU")
    ap.add_argument("--cpu-sample-seconds", type=float, default=30.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = max(args.gpus, 1)
    cores = os.cpu_count() or 1

    workload = f"{args.seconds:.0f} s synthetic full ensemble (FIC + 12 sub-channels, 854 CU: 8 DAB+ EEP, 4 MP2 UEP), u8 IQ @ 2.048 Msps, one recording per GPU"
    config = {"workload": workload, "recording_bytes": None, "l2_policy": "input (245.76 MB for 60 s) larger than the 126 MB L2; re-read from HBM every step",
              "format": "u8", "tii": True, "subchannel_discovery": "FIG 0/1 + 0/2 parsed from the decoded FIC inside the step (first from the FIC of the first 16 frames, decoded under the demodulation of the rest; the whole FIC is decoded next to the MSC chain)"}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        raw, res, n_frames = make_recording(min(args.seconds, args.cpu_sample_seconds + 1.0), 0)
        n_serv = len(res.ens.services)
        vals = []
        base = None
        for i in range(args.warmup + args.steps):
            base = run_reference_cpu(raw, n_serv, args.cpu_sample_seconds, cores)
            if i >= args.warmup:
                vals.append(base["value"])
        v = statistics.mean(vals) if vals else base["value"]
        n = min(raw.size // 2, int(args.cpu_sample_seconds * FS))
        base["value"] = v
        config["recording_bytes"] = int(2 * n)
        line = {"metric": METRIC, "value": v, "unit": "Msamples/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * n / (v * 1e6), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "impl": "reference", "x_realtime": v * 1e6 / FS, "cpu_baseline": base,
                "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    from abracadabra_b200 import pipeline
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)
    from abracadabra_b200 import shard
    affinity0 = os.sched_getaffinity(0)
    cpus = shard.bind_to_gpu(dev)                 # before any pinned allocation: NUMA-local host buffers and launch threads
    config["cpu_affinity"] = f"{len(cpus)} cores local to the GPU (NVML)" if cpus else "unchanged"

    raw, res, n_frames = make_recording(args.seconds, rank)
    n_samples = raw.size // 2
    config["recording_bytes"] = int(raw.nbytes)
    rx = pipeline.Receiver(dev)
    rx.set_profiling(True)
    stream = torch.cuda.ExternalStream(rx.stream, device=dev)
    d_raw = torch.from_numpy(raw).to(f"cuda:{dev}")
    h_raw = torch.from_numpy(raw).pin_memory()
    flags = pipeline.OPT_TII

    def step_device():
        return rx.decode_raw(d_raw.data_ptr(), n_samples, "u8", True, None, flags)

    def step_host():
        return rx.decode_raw(h_raw.data_ptr(), n_samples, "u8", False, None, flags)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    per_rank = []          # per-rank ms per step of every timed run (device-resident, sequential host), for the scaling picture

    def timed(fn, k, collect):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        ktimes = {}
        last = None
        for _ in range(k):
            last = fn()
        e1.record(stream)
        e1.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        ms = e0.elapsed_time(e1)
        if collect:                          # per-kernel CUDA-event times of the timed steps, read back after the timed region
            for back in range(min(k, 8)):
                for nm, v in rx.kernel_times(back):
                    ktimes.setdefault(nm, []).append(v)
        barrier()
        if dist is not None:
            t = torch.tensor([ms, wall], device=f"cuda:{dev}", dtype=torch.float64)
            every = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(every, t)
            per_rank.append([round(float(x[0]) / k, 4) for x in every])
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1])
        return ms, wall, ktimes, last

    # correctness gate on this very workload: every FIB passes CRC and every sub-channel equals the modulator's payload
    chk = rx.decode(d_raw.data_ptr(), "u8", tii=True, n_samples=n_samples, on_device=True)
    ok_fib = bool(chk.fib_ok.all()) and chk.n_frames == n_frames
    ok_msc = all(np.array_equal(chk.frames[i], res.logical[chk.subch[i].subch_id][:chk.frames[i].shape[0]]) for i in range(len(chk.subch)))
    n_aus, n_aus_ok = len(chk.aus), sum(a.crc_ok for a in chk.aus)
    tii_ids = [(m, s) for m, s, _ in chk.tii[0].ids] if chk.tii else []
    subch = chk.subch
    del chk

    for _ in range(args.warmup):
        step_device()
    clocks = ClockSampler(dev)
    ms_dev, wall_dev, ktimes, last = timed(step_device, args.steps, True)
    launches = last.gpu_launches
    d2h = int(last.n_slots * 12 * 33 + sum(last.frame_rows[i] * subch[i].bitrate * 3 for i in range(last.n_subch))
              + last.n_tii * ct_sizeof_tii() + last.n_slots * 32)
    for _ in range(max(1, args.warmup // 2)):
        step_host()
    rx.set_profiling(False)
    ms_seq, wall_seq, _, _ = timed(step_host, args.steps, False)
    # the call a batch user makes: a pool of two receiver contexts, so that the upload + front end of step k+1 overlaps the
    # back end of step k.  Every step still copies its own IQ from pinned host memory and reads its results back.
    pool = pipeline.ReceiverPool(dev, depth=2)
    item = (h_raw.data_ptr(), n_samples, "u8", False, None, flags)
    summ = lambda i, r, res: (res.n_frames, res.n_aus, int(res.fib_ok[0]))
    pool.map_raw([item] * max(2, args.warmup), summ)

    def timed_pool(k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        got = pool.map_raw([item] * k, summ)
        torch.cuda.synchronize(dev)
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        barrier()
        if dist is not None:
            t = torch.tensor([ms], device=f"cuda:{dev}", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        assert all(g[0] == n_frames for g in got)
        return ms
    ms_e2e = timed_pool(args.steps)
    pool.close()
    clk = clocks.stop()

    total_samples = n_samples * world
    value = total_samples * args.steps / (ms_dev * 1e-3) / 1e6
    e2e_v = total_samples * args.steps / (ms_e2e * 1e-3) / 1e6

    # per-kernel roofline (HBM): algorithmic bytes / measured launch time; peak from MEASURED_PEAKS.json (fallback 6650 GB/s)
    peak, peak_src = 6650.0, "fallback"
    try:
        mp = json.load(open(ROOT / "MEASURED_PEAKS.json"))
        peak, peak_src = float(mp["hbm_gbs"]), "measured"
    except Exception:
        pass
    stages = []
    for nm, v in ktimes.items():
        avg = statistics.mean(v)
        ab = algorithmic_bytes(nm, n_frames, subch, n_samples)
        stages.append({"kernel": nm, "ms": round(avg, 4), "algorithmic_bytes": ab, "gbs": round(ab / (avg * 1e-3) / 1e9, 2) if avg > 0 else None,
                       "frac_hbm": round(ab / (avg * 1e-3) / 1e9 / peak, 4) if avg > 0 else None})
    stages.sort(key=lambda s: -s["ms"])
    dom = stages[0] if stages else None
    roof = None
    traffic = {}
    try:   # DRAM bytes per launch from the committed ncu --set full capture of this very command (profiles/)
        traffic = json.load(open(ROOT / "profiles" / "r01_traffic.json"))["bytes_per_launch"]
    except Exception:
        pass
    for st in stages:
        st["dram_traffic_ncu"] = traffic.get(st["kernel"])
    if dom:
        trellis_steps = max(4 * n_frames - 15, 0) * sum(s.bitrate * 24 + 6 for s in subch)
        roof = {"kernel": dom["kernel"], "bound": "hbm", "achieved": dom["gbs"], "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": dom["frac_hbm"], "traffic": traffic.get(dom["kernel"]) if abs(args.seconds - 60.0) < 1e-6 else None, "ms": dom["ms"],
                "algorithmic_bytes": dom["algorithmic_bytes"],
                "note": "Viterbi ACS is issue-bound, not HBM-bound: see issue{}" if "viterbi" in dom["kernel"] else ""}
        if "viterbi(msc)" in dom["kernel"]:
            acs = trellis_steps * 64.0
            roof["issue"] = {"acs_per_s": acs / (dom["ms"] * 1e-3), "trellis_steps": trellis_steps,
                             "warp_instr_per_step_per_32cw": 209, "est_issue_slots_used": round(trellis_steps / 32 * 209 / (dom["ms"] * 1e-3)
                             / (148 * 4 * (clk.get("sm_mhz") or 1500) * 1e6), 4)}

    line = {"metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16 path metrics / f32 FFT",
            "data": "synthetic", "config": config, "x_realtime": value * 1e6 / FS, "x_realtime_per_gpu": value * 1e6 / FS / world,
            "e2e": {"value": e2e_v, "unit": "Msamples/s", "h2d_bytes_per_step": int(raw.nbytes), "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps, "x_realtime": e2e_v * 1e6 / FS,
                    "api": "ReceiverPool(depth=2).map_raw: two receiver contexts per GPU, steps pipelined (upload of step k+1 under the back end of step k)",
                    "sequential_value": total_samples * args.steps / (ms_seq * 1e-3) / 1e6, "sequential_ms_per_step": ms_seq / args.steps},
            "gpu_launches": int(launches) * args.steps, "launches_per_step": int(launches), "clocks": clk,
            "wall_ms_per_step": wall_dev / args.steps, "per_rank_ms_per_step": per_rank[0] if per_rank else None, "roofline": roof, "stages": stages,
            "check": {"frames": n_frames, "all_fib_crc_ok": ok_fib, "all_subchannels_equal_payload": ok_msc, "aus": n_aus, "aus_crc_ok": int(n_aus_ok),
                      "tii_ids": tii_ids}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, affinity0)            # the reference gets every host core again
        line["cpu_baseline"] = run_reference_cpu(raw, len(res.ens.services), args.cpu_sample_seconds, cores)
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    rx.close()


def ct_sizeof_tii():
    import ctypes as ct
    from abracadabra_b200 import pipeline
    return ct.sizeof(pipeline.Tii)


if __name__ == "__main__":
    main()
