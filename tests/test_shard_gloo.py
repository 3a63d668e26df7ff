"""N > 1 host logic on CPU: two gloo ranks shard a list of recordings round-robin, decode their share with the CPU checker
(standing in for the per-GPU receiver) and gather the results on the host -- the same code path bench.py / a band scan
uses under torchrun with one rank per GPU."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from abracadabra_b200 import shard, synth
    import orc
    dist.init_process_group("gloo", rank=rank, world_size=world)
    units = [dict(seed=s, cfo=100.0 * s) for s in range(5)]             # five "channels"

    def decode(u):
        ens = synth.config1_ensemble(seed=u["seed"])
        raw, res = synth.synthesize(ens, 6, "u8", synth.Impairments(snr_db=25, cfo_hz=u["cfo"], seed=u["seed"]), seed=u["seed"], tail=150000)
        d = orc.decode(raw, "u8", [orc.subch_struct(ens.subchannels[0])])
        ok = bool(np.array_equal(d["frames"][0], res.logical[3][:d["frames"][0].shape[0]]))
        return dict(seed=u["seed"], rank=rank, nf=int(d["nf"]), ok=ok, cfo=float(np.mean(d["cfo"])))

    out = shard.run_sharded(units, decode, rank, world, dist)
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_round_robin_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [o["seed"] for o in out] == [0, 1, 2, 3, 4]
    assert [o["rank"] for o in out] == [0, 1, 0, 1, 0]
    assert all(o["ok"] and o["nf"] == 6 for o in out)
    assert all(abs(o["cfo"] - 100.0 * o["seed"]) < 2.0 for o in out)


def test_assign_and_chunking():
    from abracadabra_b200 import shard
    assert shard.assign(38, 3, 8) == [3, 11, 19, 27, 35]
    assert sorted(sum((shard.assign(38, r, 8) for r in range(8)), [])) == list(range(38))
    with pytest.raises(ValueError):
        shard.assign(4, 4, 4)
    n = 3600 * shard.FS
    ch = shard.chunk_recording(n, 8)
    assert ch[0][0] == 0 and ch[-1][1] == n and len(ch) == 8
    for (a0, b0), (a1, b1) in zip(ch, ch[1:]):
        assert b0 - a1 >= shard.CHUNK_OVERLAP_FRAMES * shard.T_FRAME     # >= 0.7 s of overlap (SURVEY.md sec. 8 e)
    merged = shard.merge_chunks([[(0, "a"), (1, "b"), (2, "c")], [(1, "B"), (2, "C"), (3, "d")]])
    assert merged == [(0, "a"), (1, "b"), (2, "c"), (3, "d")]
