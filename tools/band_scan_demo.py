#!/usr/bin/env python
"""Band scan over synthetic channel recordings (BASELINE.json configs[3] in small), sharded over the ranks of a torchrun job:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/band_scan_demo.py --channels 8
Every rank decodes its share on its own GPU; rank 0 prints the ScannerBackend-style CSV and the aggregate rate."""
import argparse
import os
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from abracadabra_b200 import dabconst, rawfile, scanner, shard, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--channels", type=int, default=8)
ap.add_argument("--frames", type=int, default=100)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
shard.bind_to_gpu(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tmp = Path(tempfile.gettempdir()) / "b200dab_scan_demo"
tmp.mkdir(exist_ok=True)
names = list(dabconst.BAND3)[:a.channels]
paths = [str(tmp / f"scan_{n}.uff") for n in names]
for k in shard.assign(len(names), rank, world):          # every rank writes the recordings it will not necessarily read: shared /tmp
    ens = synth.config1_ensemble(seed=500 + k)
    ens.eid, ens.label = 0x6000 + k, f"CHANNEL {names[k]}"
    ens.tii = [synth.TiiTx(10 + k, 1 + k % 23, 1.0)]
    raw, _ = synth.synthesize(ens, a.frames, "s16", synth.Impairments(snr_db=25.0, cfo_hz=100.0 * (k - 3), seed=k, rms=7000.0), seed=k, tail=20000)
    rawfile.write_uff(paths[k], raw, "s16", dabconst.BAND3[names[k]])
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
lines = scanner.scan_files(paths, device=local, rank=rank, world=world, dist=dist if world > 1 else None)
dt = time.perf_counter() - t0
if rank == 0:
    print("\n".join(lines))
    n = a.channels * a.frames * 196608
    print(f"# {a.channels} channels x {a.frames * 0.096:.1f} s on {world} GPU(s): {dt:.2f} s incl. file reads = {n / dt / 1e6:.0f} Msamples/s", file=sys.stderr)
if world > 1:
    dist.destroy_process_group()
