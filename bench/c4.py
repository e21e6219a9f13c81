#!/usr/bin/env python
"""BASELINE config C4: one ~2^30-entry clique table (8 GiB fp64) split along three binary variables across the
GPUs of one box; the sum over the split variables is an NCCL all-reduce of the per-slice partial posteriors.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P bench/c4.py

Each rank generates only its own 2^27-entry slice of the procedural CPT on its GPU (1 GiB) and evaluates the
junction-tree query restricted to it; ranks then all-reduce a [rows][60] matrix.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import hbayes_b200 as hb
from hbayes_b200 import synth
from hbayes_b200.dist import max_over_ranks
from hbayes_b200.sharded import SplitJunctionTree

ap = argparse.ArgumentParser()
ap.add_argument("--parents", type=int, default=29)
ap.add_argument("--rows", type=int, default=4)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
net, split = synth.config_c4(a.parents)
pnet = hb.BayesianNetwork.from_tables_procedural(net.dims, net.scopes, net.values, net.proc_seed)
t0 = time.perf_counter()
sjt = SplitJunctionTree(pnet, split, rank=rank, world=world, device=local)
sjt.specialise([net.n - 1])
rng = np.random.RandomState(5)
ev = np.full((a.rows, net.n), -1, np.int8)
ev[:, net.n - 1] = rng.randint(0, 2, a.rows)  # the child is observed: every table of the query is per-row
out = sjt.posteriors_batch(ev)  # compiles, generates the slice, first pass
torch.cuda.synchronize()
setup_s = time.perf_counter() - t0
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(a.steps):
    out = sjt.posteriors_batch(ev)
e1.record()
torch.cuda.synchronize()
ms = max_over_ranks(e0.elapsed_time(e1) / a.steps, device="cuda")
st = sjt.slices[0].last_stats()
res = out.cpu().numpy()
off = np.concatenate([[0], np.cumsum(net.dims)])
norm_err = max(abs(res[:, off[v]:off[v + 1]].sum(1) - 1).max() for v in range(net.n))
gathered = [None] * world
if world > 1:
    dist.all_gather_object(gathered, res.tobytes())
else:
    gathered = [res.tobytes()]
if rank == 0:
    slices_per_rank = len(sjt.slices)
    slice_entries = 2 ** (a.parents + 1) // (8 // 1) if True else 0
    bytes_per_pass = 8.0 * (st["row_read_per_row"] + st["row_write_per_row"]) * a.rows * slices_per_rank
    print(json.dumps({
        "config": "C4: 1 child + %d binary parents, CPT 2^%d entries (%.1f GiB), split along variables %s into 8 slices"
                  % (a.parents, a.parents + 1, 8.0 * 2 ** (a.parents + 1) / 2 ** 30, split),
        "n_gpus": world, "slices_per_gpu": slices_per_rank, "rows": a.rows, "ms_per_pass": ms, "queries_per_s": a.rows / ms * 1e3,
        "setup_s_incl_slice_generation": setup_s, "allreduce_doubles": int(a.rows * pnet.width),
        "per_gpu_algorithmic_GBps": bytes_per_pass / (ms / 1e3) / 1e9, "per_row_per_slice": st,
        "max_abs_deviation_of_marginal_sums_from_1": float(norm_err), "all_ranks_identical": len(set(gathered)) == 1,
        "posterior_parent_3_row0": res[0, off[3]:off[4]].tolist(),
    }))
if world > 1:
    dist.destroy_process_group()
