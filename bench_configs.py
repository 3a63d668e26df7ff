"""The other legs of bench.py: BASELINE.json configs[0], [2], [3], [4] on the GPU (python bench.py --config {1,3,4,5}), the reference
arm (--impl reference: the reference's own CPU core on the host cores, any config) and the drop-in arm (--impl dropin: one stream
through the reference-facing dabsdr C API of libdabsdr.so.4).  Every line carries the keys of the default line (bench.py)."""
from __future__ import annotations

import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

import bench
from bench import FS, T_FRAME, METRIC, WORKLOADS, ClockSampler, Gpu, base_line, peak_hbm, run_reference_cpu


def _config(args, k, **extra):
    c = {"workload": WORKLOADS[k], "format": "u8", "tii": True}
    c.update(extra)
    return c


def _flush_l2(torch, dev, buf=[None]):
    if buf[0] is None:
        buf[0] = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{dev}")
    buf[0].zero_()


# ================================================================== reference arm
def _reference_jobs(args, td):
    """files of the config's workload for the reference core (bounded sample where the whole workload would take minutes of CPU)"""
    from abracadabra_b200 import synth, workloads as W
    k = args.config
    if k == 2:
        raw, res, _ = bench.make_recording(args.seconds, 0)
        n = raw.size // 2 if args.cpu_sample_seconds >= args.seconds else min(raw.size // 2, int(args.cpu_sample_seconds * FS))   # the same file, lead-in and tail included
        p = os.path.join(td, "full.u8")
        raw[:2 * n].tofile(p)
        return [(p, "u8", len(res.ens.services), True)], int(2 * n), "the whole recording, all 12 audio services + TII"
    if k == 1:
        raw, res = synth.synthesize(synth.config1_ensemble(), int(round(args.seconds * FS / T_FRAME)), "u8",
                                    synth.Impairments(snr_db=40.0, cfo_hz=777.0, seed=1001, lead_in=49152), seed=1001, tail=65536)
        p = os.path.join(td, "c1.u8")
        raw.tofile(p)
        return [(p, "u8", 1, True)], int(raw.nbytes), "the whole recording, the DAB+ service + TII"
    if k == 3:
        jobs, nb = [], 0
        for i in range(24):
            raw, _, _ = W.tii_recording(i)
            p = os.path.join(td, f"tii{i}.u8")
            raw.tofile(p)
            jobs.append((p, "u8", 0, True)); nb += raw.nbytes
        return jobs, nb, "24 of the 500 recordings (TII only, no service selected)"
    if k == 4:
        jobs, nb = [], 0
        for ch in range(0, 38, 3):
            raw, _, name, khz = W.band_scan_recording(ch, args.seconds)
            p = os.path.join(td, f"scan_{name}.u8")
            raw.tofile(p)
            jobs.append((p, "u8", 0, True)); nb += raw.nbytes
        return jobs, nb, "13 of the 38 channel recordings (every third channel), ensemble + TII, no service selected (what the scanner does)"
    raw, _ = W.staircase_recording(int(round(min(args.seconds, 120.0) * FS / T_FRAME)), "s16")
    p = os.path.join(td, "long.s16")
    raw.tofile(p)
    return [(p, "s16", 12, False)], int(raw.nbytes), f"the first {min(args.seconds, 120.0):.0f} s of the recording (0-20 dB staircase compressed into it), all 12 audio services"


def run_reference_arm(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import orc
    cores = os.cpu_count() or 1
    config = _config(args, args.config)
    if args.config == 2:
        config = {"workload": WORKLOADS[2].replace("60 s", f"{args.seconds:.0f} s"), "recording_bytes": None,
                  "l2_policy": "input (245.76 MB for 60 s) larger than the 126 MB L2; re-read from HBM every step", "format": "u8", "tii": True}
    with tempfile.TemporaryDirectory() as td:
        if not orc.have_ref():
            raw, _, _ = bench.make_recording(min(args.seconds, 4.0), 0)
            base = bench.run_port_cpu(raw, 4.0)
            vals = [base["value"]]
            nbytes, n_samples = raw.nbytes, raw.size // 2
        else:
            jobs, nbytes, what = _reference_jobs(args, td)
            n_samples = sum(os.path.getsize(p) // orc.FMT_BYTES[f] for p, f, _, _ in jobs)
            vals, base = [], None
            for i in range(args.warmup + args.steps):
                base, _ = run_reference_cpu(jobs, cores)
                if i >= args.warmup:
                    vals.append(base["value"])
            base["sample"] = what + ".  " + base["sample"]
            base.pop("aus", None)
    v = statistics.mean(vals)
    base["value"] = v
    config["recording_bytes"] = int(nbytes)
    line = {"metric": METRIC, "value": v, "unit": "Msamples/s", "n_gpus": max(args.gpus, 1), "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * n_samples / (v * 1e6), "higher_is_better": True, "scaling": "weak" if args.config != 5 else "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config, "impl": "reference", "x_realtime": v * 1e6 / FS, "cpu_baseline": base,
            "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ================================================================== drop-in arm: one stream through the dabsdr C API
def run_dropin_arm(args):
    """`shim_harness --free-run` (the RadioControl-like driver of oracle/ref_harness.c linked against OUR libdabsdr.so.4, which runs
    the CUDA path) against `ref_harness --free-run` (same driver, the reference core) on the same file: one audio service + TII,
    float IQ pulled through the registered input function as fast as the core asks for it.  This is the reference-facing call of a
    single live stream; the batch C ABI (default arm) is the throughput interface."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import orc
    raw, res, n_frames = bench.make_recording(60.0, 0)
    out = {}
    with tempfile.TemporaryDirectory() as td:
        files = {}
        for secs in (20.0, 60.0):
            files[secs] = os.path.join(td, f"s{int(secs)}.u8")
            raw[:2 * int(secs * FS)].tofile(files[secs])
        for name, exe in (("dropin", orc.REF_DIR / "shim_harness"), ("reference", orc.REF_DIR / "ref_harness")):
            per = {}
            for secs, p in files.items():
                runs = []
                for i in range(args.warmup + args.steps):
                    r = subprocess.run([str(exe), p, "u8", "-", "--service", "0", "--tii", "1", "--tail-frames", "2", "--free-run"],
                                       capture_output=True, text=True, timeout=600)
                    try:
                        runs.append(json.loads(r.stdout.strip().splitlines()[-1]))
                    except Exception:
                        runs.append({"msps": 0.0, "seconds": 1e9, "samples": 0, "aus": 0, "error": (r.stderr or "")[-300:]})
                keep = runs[args.warmup:]
                per[secs] = {"seconds": statistics.mean(x["seconds"] for x in keep), "samples": keep[-1]["samples"], "aus": keep[-1].get("aus"),
                             "crc_ok": keep[-1].get("crc_ok"), "fib_err": keep[-1].get("fib_err")}
            a, b = per[20.0], per[60.0]
            steady = (b["samples"] - a["samples"]) / max(b["seconds"] - a["seconds"], 1e-9) / 1e6
            out[name] = {"msps": b["samples"] / b["seconds"] / 1e6, "msps_steady": steady,
                         "startup_s": max(a["seconds"] - a["samples"] / (steady * 1e6), 0.0), "aus": b["aus"], "crc_ok": b["crc_ok"], "fib_err": b["fib_err"],
                         "seconds_20s_file": a["seconds"], "seconds_60s_file": b["seconds"]}
    secs = 60.0
    v = out["dropin"]["msps"]
    line = {"metric": METRIC, "value": v, "unit": "Msamples/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * (raw.size // 2) / max(v, 1e-9) / 1e6, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16 path metrics / f32 FFT", "data": "synthetic", "impl": "dropin",
            "config": {"workload": f"one live stream: the {secs:.0f} s full ensemble, ONE audio service + TII, through the dabsdr C API (dabsdr.h:397-429) with the input function free-running",
                       "recording_bytes": int(raw.nbytes), "format": "u8 file -> float IQ through dabsdrRegisterInputFcn"},
            "x_realtime": v * 1e6 / FS, "dropin": out["dropin"], "reference_same_call": out["reference"],
            "ratio_vs_reference_same_call": v / max(out["reference"]["msps"], 1e-9),
            "ratio_steady_state": out["dropin"]["msps_steady"] / max(out["reference"]["msps_steady"], 1e-9),
            "note": "msps = samples of the 60 s file / wall time from the Tune request to the end of the file; msps_steady = (60 s file - 20 s file) samples / time difference, i.e. without what a run spends before its first result (first-call allocations on the GPU; the CUDA context itself is created in dabsdr(), before Tune)",
            "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": int(raw.nbytes * 4), "d2h_bytes_per_step": None,
                    "note": "the call is end to end by construction: host float IQ in through the input function, callbacks out"}}
    print(json.dumps(line))


# ================================================================== config 1: 10 s, one DAB+ service
def run_config1(args):
    from abracadabra_b200 import pipeline, synth
    g = Gpu(args)
    torch, dev = g.torch, g.dev
    n_frames = int(round(args.seconds * FS / T_FRAME))
    raw, res = synth.synthesize(synth.config1_ensemble(), n_frames, "u8", synth.Impairments(snr_db=40.0, cfo_hz=777.0 + g.rank, seed=1001 + g.rank, lead_in=49152),
                                seed=1001 + g.rank, tail=65536)
    n_samples = raw.size // 2
    rx = pipeline.Receiver(dev)
    d_raw = torch.from_numpy(raw).to(f"cuda:{dev}")
    h_raw = torch.from_numpy(raw).pin_memory()
    chk = rx.decode(d_raw.data_ptr(), "u8", tii=True, n_samples=n_samples, on_device=True)
    sc = res.ens.subchannels[0]
    ok = bool(chk.fib_ok.all()) and chk.n_frames == n_frames and len(chk.subch) == 1 and \
        np.array_equal(chk.frames[0], res.logical[sc.subchid][:chk.frames[0].shape[0]]) and all(a.crc_ok for a in chk.aus)
    tii_ids = [(m, s) for m, s, _ in chk.tii[0].ids] if chk.tii else []
    clocks = ClockSampler(dev)

    def each(fn):   # the 41 MB input fits the 126 MB L2: flush it before every timed step (256 MB written), time the steps one by one
        tot = 0.0
        last = None
        for i in range(args.warmup + args.steps):
            _flush_l2(torch, dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            e0.record()
            last = fn()
            torch.cuda.synchronize(dev)
            e1.record(); e1.synchronize()
            if i >= args.warmup:
                tot += e0.elapsed_time(e1)
        g.barrier()
        (m,), _ = g.max_over_ranks(tot)
        return m, last
    ms_dev, last = each(lambda: rx.decode_raw(d_raw.data_ptr(), n_samples, "u8", True, None, pipeline.OPT_TII))
    ms_e2e, _ = each(lambda: rx.decode_raw(h_raw.data_ptr(), n_samples, "u8", False, None, pipeline.OPT_TII))
    clk = clocks.stop()
    tot = n_samples * g.world
    value, e2e_v = tot * args.steps / (ms_dev * 1e-3) / 1e6, tot * args.steps / (ms_e2e * 1e-3) / 1e6
    line = base_line(args, g, _config(args, 1, recording_bytes=int(raw.nbytes), l2_policy="L2 flushed (256 MB written) before every timed step; steps timed one by one"),
                     value, ms_dev / args.steps)
    line.update({"e2e": {"value": e2e_v, "unit": "Msamples/s", "h2d_bytes_per_step": int(raw.nbytes),
                         "d2h_bytes_per_step": int(last.n_slots * 12 * 33 + last.frame_rows[0] * sc.bitrate * 3), "ms_per_step": ms_e2e / args.steps},
                 "gpu_launches": int(last.gpu_launches) * args.steps, "clocks": clk,
                 "check": {"frames": n_frames, "fibs_frames_aus_equal_payload": bool(ok), "aus": len(chk.aus), "tii_ids": tii_ids}})
    if g.rank == 0 and g.world == 1 and not args.no_cpu_baseline:
        import orc
        os.sched_setaffinity(0, g.affinity0)
        if orc.have_ref():
            with tempfile.TemporaryDirectory() as td:
                p = os.path.join(td, "c1.u8")
                raw.tofile(p)
                base, streams = run_reference_cpu([(p, "u8", 1, True)], os.cpu_count() or 1, keep_streams=True)
                same, n_ref, _ = bench.compare_with_reference(chk, res, streams, p)
            base.pop("aus", None)
            line["cpu_baseline"] = base
            line["check"].update({"reference_aus_equal": same, "reference_aus": n_ref})
            ok = ok and same
    if g.rank == 0:
        print(json.dumps(line))
    g.close(); rx.close()
    if not ok:
        raise SystemExit(1)


# ================================================================== config 3: TII batch
def run_config3(args):
    from abracadabra_b200 import pipeline, workloads as W
    g = Gpu(args)
    torch, dev = g.torch, g.dev
    n_rec, n_unique = 500, 24
    mine = [r for r in range(n_rec) if r % g.world == g.rank]
    uniq = sorted({r % n_unique for r in mine})
    recs = {u: W.tii_recording(u) for u in uniq}
    host = {u: torch.from_numpy(recs[u][0]).pin_memory() for u in uniq}
    devb = {u: host[u].to(f"cuda:{dev}") for u in uniq}
    flags = pipeline.OPT_TII | pipeline.OPT_NO_MSC
    rx = pipeline.Receiver(dev)
    ok, n_ids = True, 0
    for u in uniq:                      # ids of every unique recording = the transmitted ones that are strong enough to detect, in sub-id order
        d = rx.decode(recs[u][0], "u8", tii=True, no_msc=True)
        win = [w for w in d.tii if w.ids]
        ids = [(m, s) for m, s, _ in win[0].ids] if win else []
        sent = {(t.main, t.sub): t.amp for t in recs[u][1]}
        top = max(sent.values())
        must = sorted(k for k, a in sent.items() if a >= 0.35 * top)
        ok = ok and all(i in sent for i in ids) and all(m in ids for m in must) and ids == sorted(ids, key=lambda x: x[1])
        n_ids += len(ids)
    pool = pipeline.ReceiverPool(dev, depth=4)          # small recordings: four contexts keep the copy engine and the SMs busy
    n_s = {u: recs[u][0].size // 2 for u in uniq}
    items_d = [(devb[r % n_unique].data_ptr(), n_s[r % n_unique], "u8", True, None, flags) for r in mine]
    items_h = [(host[r % n_unique].data_ptr(), n_s[r % n_unique], "u8", False, None, flags) for r in mine]
    summ = lambda i, r, res_: (res_.n_frames, res_.n_tii, res_.gpu_launches)
    pool.map_raw(items_d[:8], summ)
    clocks = ClockSampler(dev)
    ms_dev, _, _, got = g.timed(lambda: pool.map_raw(items_d, summ), 1)
    ms_e2e, _, _, _ = g.timed(lambda: pool.map_raw(items_h, summ), 1)
    clk = clocks.stop()
    tot = g.sum_over_ranks(sum(n_s[r % n_unique] for r in mine))
    nulls = g.sum_over_ranks(sum(x[0] for x in got)) / 2.0
    value, e2e_v = tot / (ms_dev * 1e-3) / 1e6, tot / (ms_e2e * 1e-3) / 1e6
    args.steps, args.warmup = 1, 3
    line = base_line(args, g, _config(args, 3, recordings=n_rec, unique_recordings=n_unique, recording_bytes=int(sum(recs[u][0].nbytes for u in uniq) / len(uniq)),
                                      l2_policy=f"{len(uniq)} distinct recordings of 17 MB per rank cycled ({len(uniq) * 17} MB > 126 MB L2 where 8 or more)",
                                      note="one step = the whole batch of 500 recordings (sync + FIC + TII, no MSC); 24 distinct recordings are cycled"),
                     value, ms_dev)
    line.update({"tii_nulls_per_s": nulls / (ms_dev * 1e-3), "recordings_per_s": n_rec / (ms_dev * 1e-3),
                 "e2e": {"value": e2e_v, "unit": "Msamples/s", "h2d_bytes_per_step": int(sum(2 * n_s[r % n_unique] for r in mine)),
                         "d2h_bytes_per_step": int(sum(x[0] for x in got) * (12 * 33 + 384 * 4 + 8) + sum(x[1] for x in got) * bench.ct_sizeof_tii()), "ms_per_step": ms_e2e},
                 "gpu_launches": int(sum(x[2] for x in got)), "clocks": clk,
                 "check": {"tii_ids_consistent_with_transmitters": bool(ok), "ids_detected_in_unique_recordings": n_ids}})
    if g.rank == 0:
        print(json.dumps(line))
    pool.close(); rx.close(); g.close()
    if not ok:
        raise SystemExit(1)


# ================================================================== config 4: band scan
def run_config4(args):
    from abracadabra_b200 import pipeline, scanner, workloads as W
    g = Gpu(args)
    torch, dev = g.torch, g.dev
    mine = [c for c in range(38) if c % g.world == g.rank]
    recs = {c: W.band_scan_recording(c, args.seconds) for c in mine}
    host = {c: torch.from_numpy(recs[c][0]).pin_memory() for c in mine}
    devb = {c: host[c].to(f"cuda:{dev}") for c in mine}
    flags = pipeline.OPT_TII | pipeline.OPT_NO_MSC | pipeline.OPT_SYMSTAT
    rx = pipeline.Receiver(dev)
    ok = True
    for c in mine:                      # scanner rows of every channel: ensemble id / label / services and the transmitted TII ids
        rows = scanner.scan_recording(rx, recs[c][0], "u8", recs[c][3])
        ens = W.band_scan_ensemble(c)
        want = sorted(((t.main, t.sub) for t in ens.tii), key=lambda x: x[1])
        ok = ok and [(r.main, r.sub) for r in rows] == want and all(r.ueid == (0xE1 << 16 | ens.eid) and r.channel == recs[c][2] and r.n_services == 1 for r in rows)
    pool = pipeline.ReceiverPool(dev, depth=4)
    n_s = {c: recs[c][0].size // 2 for c in mine}
    summ = lambda i, r, res_: (res_.n_frames, res_.n_tii, res_.gpu_launches)
    items_d = [(devb[c].data_ptr(), n_s[c], "u8", True, None, flags) for c in mine]
    items_h = [(host[c].data_ptr(), n_s[c], "u8", False, None, flags) for c in mine]
    for _ in range(args.warmup):
        pool.map_raw(items_d, summ)
    clocks = ClockSampler(dev)
    ms_dev, _, _, got = g.timed(lambda: [pool.map_raw(items_d, summ) for _ in range(args.steps)][-1], 1)
    ms_e2e, _, _, _ = g.timed(lambda: [pool.map_raw(items_h, summ) for _ in range(args.steps)][-1], 1)
    clk = clocks.stop()
    tot = g.sum_over_ranks(sum(n_s.values()))
    value, e2e_v = tot * args.steps / (ms_dev * 1e-3) / 1e6, tot * args.steps / (ms_e2e * 1e-3) / 1e6
    line = base_line(args, g, _config(args, 4, recording_bytes=int(recs[mine[0]][0].nbytes), channels=38,
                                      l2_policy=f"{len(mine)} recordings of 41 MB per rank in turn ({len(mine) * 41} MB; at N = 8 a rank's 5 recordings exceed the 126 MB L2 as well)",
                                      note="one step = the whole band (38 recordings): sync + FIC + TII + SNR per channel, no MSC -- what ScannerBackend collects (src/tii/scannerbackend.cpp:725-956)"),
                     value, ms_dev / args.steps, scaling="strong")
    line.update({"channels_per_s": 38 * args.steps / (ms_dev * 1e-3),
                 "e2e": {"value": e2e_v, "unit": "Msamples/s", "h2d_bytes_per_step": int(sum(2 * v for v in n_s.values())),
                         "d2h_bytes_per_step": int(sum(x[0] for x in got) * (12 * 33 + 384 * 4 + 8 + 600) + sum(x[1] for x in got) * bench.ct_sizeof_tii()),
                         "ms_per_step": ms_e2e / args.steps},
                 "gpu_launches": int(sum(x[2] for x in got)) * args.steps, "clocks": clk,
                 "check": {"scanner_rows_equal_transmitted": bool(ok)}})
    if g.rank == 0:
        print(json.dumps(line))
    pool.close(); rx.close(); g.close()
    if not ok:
        raise SystemExit(1)


# ================================================================== config 5: one long recording in time chunks
def run_config5(args):
    from abracadabra_b200 import pipeline, shard, workloads as W
    g = Gpu(args)
    torch, dev = g.torch, g.dev
    fmt = "s16"
    n_chunks = args.chunks or max(8, g.world)
    t0 = time.time()
    gen = W.TiledLongRecording(args.seconds, device=f"cuda:{dev}")
    ranges = shard.chunk_recording(gen.n_samples, n_chunks)
    mine = [c for c in range(len(ranges)) if c % g.world == g.rank]
    host = {c: torch.from_numpy(gen.chunk(*ranges[c], fmt)).pin_memory() for c in mine}
    t_gen = time.time() - t0
    del gen.clean
    torch.cuda.empty_cache()
    rx = pipeline.Receiver(dev)
    n_s = {c: host[c].numel() // 2 for c in mine}
    flags = 0
    clocks = ClockSampler(dev)
    # device-resident: one chunk at a time on the device (a chunk of an hour-long s16 recording is 3.7 GB)
    ms_dev = 0.0
    launches = 0
    for c in mine:
        d = host[c].to(f"cuda:{dev}")
        for i in range(2):              # first pass warms the context up for this chunk size
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            e0.record()
            r = rx.decode_raw(d.data_ptr(), n_s[c], fmt, True, None, flags)
            torch.cuda.synchronize(dev)
            e1.record(); e1.synchronize()
        ms_dev += e0.elapsed_time(e1)
        launches += r.gpu_launches
        del d
    g.barrier()
    (ms_dev,), per_dev = g.max_over_ranks(ms_dev)
    # end to end: the chunks from pinned host memory through shard.decode_chunked's own calls, results merged on rank 0
    g.barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    decs = {c: rx.decode_raw(host[c].data_ptr(), n_s[c], fmt, False, None, flags).n_frames for c in mine}
    torch.cuda.synchronize(dev)
    e1.record(); e1.synchronize()
    (ms_e2e,), _ = g.max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.stop()
    # the product call: chunked decode with the global decisions and the merge (host side), timed as a whole
    st = {}
    big = None
    t0 = time.perf_counter()

    class _Rec:          # presents this rank's chunks as slices of the whole recording without materialising it
        size = gen.n_samples * 2

        def __getitem__(self, sl):
            a = sl.start // 2
            for c in mine:
                if ranges[c][0] == a:
                    return host[c].numpy()
            raise KeyError(sl)
    merged = shard.decode_chunked(_Rec(), fmt, n_chunks, rx=rx, tii=False, rank=g.rank, world=g.world, dist=g.dist, stats=st)
    t_merge = time.perf_counter() - t0
    tot = float(gen.n_samples)
    value, e2e_v = tot / (ms_dev * 1e-3) / 1e6, tot / (ms_e2e * 1e-3) / 1e6
    args.steps, args.warmup = 1, 3
    line = base_line(args, g, _config(args, 5, format=fmt, tii=False, recording_seconds=args.seconds, recording_bytes=int(tot * 4), chunks=len(ranges),
                                      chunk_overlap_frames=shard.CHUNK_OVERLAP_FRAMES,
                                      l2_policy="every chunk (hundreds of MB to GB) exceeds the 126 MB L2",
                                      note="one step = the whole recording once; chunk c on rank c % N; value / e2e = device time of the slowest rank for its chunks; the tiled test signal restarts its interleaver every 57.6 s"),
                     value, ms_dev, scaling="strong")
    line.update({"e2e": {"value": e2e_v, "unit": "Msamples/s", "h2d_bytes_per_step": int(sum(4 * v for v in n_s.values())), "d2h_bytes_per_step": None, "ms_per_step": ms_e2e},
                 "gpu_launches": int(launches), "clocks": clk, "per_rank_ms": [round(x, 2) for x in per_dev[0]],
                 "decode_chunked": {"wall_s": round(t_merge, 2), "second_passes_this_rank": st.get("second_passes_this_rank"),
                                    "merged_frames": merged.n_frames if merged is not None else None, "merged_aus": len(merged.aus) if merged is not None else None,
                                    "merged_aus_crc_ok": int(sum(a.crc_ok for a in merged.aus)) if merged is not None else None,
                                    "note": "shard.decode_chunked: every chunk decoded from host memory, global sub-channel table / superframe alignment, host-side gather and merge (Python)"},
                 "generation_s": round(t_gen, 1),
                 "check": {"frames_found": int(sum(decs.values()))}})
    if g.rank == 0:
        print(json.dumps(line))
    rx.close(); g.close()
