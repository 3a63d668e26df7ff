#!/usr/bin/env python
"""Raw pinned host-to-device copy ceiling of the box: every visible GPU copies `--mb` MB from its own pinned buffer at the same
time (one process, one stream per GPU; the copy engines run concurrently), for 1, 2, 4, ... GPUs.  Nothing of the receive path
runs here: this is the denominator for bench.py's e2e.frac_of_h2d_ceiling (the end-to-end step is bound by this copy).

    python tools/h2d_ceiling.py [--mb 246] [--reps 20] > profiles/rNN_h2d_ceiling.json
"""
import argparse
import json
import os
import subprocess
import time

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=245.76)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    n = int(args.mb * 1e6)
    ng = torch.cuda.device_count()
    host = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(ng)]
    dev = [torch.empty(n, dtype=torch.uint8, device=f"cuda:{g}") for g in range(ng)]
    streams = [torch.cuda.Stream(device=g) for g in range(ng)]
    out = {"mb_per_copy": args.mb, "reps": args.reps, "cpu_count": os.cpu_count(), "ways": {}}
    ways = [w for w in (1, 2, 4, 8) if w <= ng]
    for w in ways:
        for rep in range(2):                       # first pass warms up
            for g in range(w):
                torch.cuda.synchronize(g)
            t0 = time.perf_counter()
            for _ in range(args.reps):
                for g in range(w):
                    with torch.cuda.stream(streams[g]):
                        dev[g].copy_(host[g], non_blocking=True)
            for g in range(w):
                streams[g].synchronize()
            dt = time.perf_counter() - t0
        out["ways"][str(w)] = {"aggregate_gbs": round(w * args.reps * n / dt / 1e9, 2), "per_gpu_gbs": round(args.reps * n / dt / 1e9, 2),
                               "ms_per_copy": round(dt / args.reps * 1e3, 3)}
    for name, cmd in (("topo", ["nvidia-smi", "topo", "-m"]), ("numa", ["numactl", "-H"]), ("lscpu", ["lscpu"])):
        try:
            out[name] = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout.splitlines()[:60]
        except Exception as e:
            out[name] = [f"unavailable: {e}"]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
