"""world_size-2 gloo test of the N>1 path's host logic (no GPU): row sharding tiles the batch, the timing
protocol takes the max over ranks, and both ranks compile the identical program (same step list) so that a
row gives the same answer whichever GPU it lands on."""
import hashlib
import json
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hbayes_b200.dist import max_over_ranks, shard_rows, whole_job_rate


def test_shard_rows_tile_the_batch():
    for n in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            cuts = [shard_rows(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hbayes_b200 as hb
    from hbayes_b200 import synth

    net, observed = synth.config_c2()
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    prog = hb.JunctionTree(pnet, device=None).program(observed)
    digest = hashlib.sha256(json.dumps(prog["physical"], sort_keys=True).encode()).hexdigest()
    lo, hi = shard_rows(1000, rank, world)
    ms = max_over_ranks(10.0 * (rank + 1))  # rank 1 is the slow one
    gathered = [None] * world
    dist.all_gather_object(gathered, (digest, lo, hi))
    dist.barrier()
    q.put((rank, ms, gathered))
    dist.destroy_process_group()


def test_two_ranks_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ms, gathered in res:
        assert ms == 20.0  # max over ranks
        assert gathered[0][0] == gathered[1][0]  # identical compiled programs
        assert (gathered[0][1], gathered[0][2], gathered[1][1], gathered[1][2]) == (0, 500, 500, 1000)
    assert whole_job_rate([500, 500], 4, 20.0) == 1000 * 4 / 0.02
