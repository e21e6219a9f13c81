"""Split clique (BASELINE config C4): a clique/CPT table too large for one GPU is split along `split_vars`.

Two forms, same numbers (1e-12: the sum over the split variables is reassociated, so not bit-identical to the unsplit engine):

* **exchange inside the library** (`nccl_id` given): one rank per state combination of the split variables; every rank's
  handle (`hb_jt_compile_split_nccl`) owns its slice and the library sums over the split variables where the reference's
  elimination does -- ncclAllReduce(ncclDouble, ncclSum) of |separator| doubles per row for every message that leaves the
  split clique, of |query| doubles for every posterior inside it -- on its own stream.  Work downstream of such a message
  is not conditioned on the slice.  Batch calls are collective.
* **cutset conditioning** (no `nccl_id`): every rank owns the slices rank, rank + world, ... and evaluates the whole
  program conditioned on each; the unnormalised result matrices are summed locally, then across ranks with one
  torch.distributed all-reduce, then normalised.  Works for any number of slices per rank and on one GPU."""
from __future__ import annotations

import itertools

import numpy as np

from .api import BayesianNetwork, JunctionTree, nccl_unique_id, nodeComparisonForTriangulation


def broadcast_unique_id(rank: int, src: int = 0) -> bytes:
    """rank `src` creates the NCCL id, every rank of the default torch.distributed group receives it"""
    import torch.distributed as dist

    box = [nccl_unique_id() if rank == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


class SplitJunctionTree:
    def __init__(self, net: BayesianNetwork, split_vars, rank=0, world=1, device=0, cmp=nodeComparisonForTriangulation, nccl_id=None):
        self.net, self.split_vars, self.rank, self.world = net, list(split_vars), rank, world
        combos = list(itertools.product(*[range(net.dims[v]) for v in self.split_vars]))
        self.nccl = nccl_id is not None
        if self.nccl:
            if len(combos) != world:
                raise ValueError("the library-side exchange deals one slice per rank: %d slices, %d ranks" % (len(combos), world))
            self.combos = [combos[rank]]
            self.slices = [JunctionTree(net, cmp=cmp, device=device, split=list(zip(self.split_vars, combos[rank])), nccl=(rank, world, nccl_id))]
            return
        self.combos = combos[rank::world]
        self.slices = [JunctionTree(net, cmp=cmp, device=device, split=list(zip(self.split_vars, c))) for c in self.combos]

    def specialise(self, observed) -> int:
        """build the generated kernels of every slice for rows that observe `observed` (the split variables are added)"""
        obs = sorted(set(observed) | set(self.split_vars))
        return sum(jt.specialise(obs) for jt in self.slices)

    def partial_batch(self, evidence: np.ndarray):
        """cutset form: sum over this rank's slices of the unnormalised posteriors; a CUDA tensor [rows][sum dims]"""
        import torch

        assert not self.nccl
        evidence = np.ascontiguousarray(evidence, np.int8)
        total = torch.zeros((evidence.shape[0], self.net.width), dtype=torch.float64, device="cuda")
        for combo, jt in zip(self.combos, self.slices):
            ev = evidence.copy()
            keep = np.ones(ev.shape[0], bool)
            for v, s in zip(self.split_vars, combo):  # a row that observes a split variable only has that slice
                keep &= (ev[:, v] < 0) | (ev[:, v] == s)
                ev[:, v] = s
            part = jt.posteriors_batch(torch.from_numpy(ev).cuda())
            part[torch.from_numpy(~keep).cuda()] = 0.0
            total += part
        return total

    def posteriors_batch(self, evidence, out=None):
        """posterior of every variable for every row, replicated on all ranks (collective)"""
        if self.nccl:  # the library exchanges and normalises; rows leave the split variables unobserved
            import torch

            ev = torch.from_numpy(np.ascontiguousarray(evidence, np.int8)).cuda() if isinstance(evidence, np.ndarray) else evidence
            return self.slices[0].posteriors_batch(ev, out)
        import torch.distributed as dist

        total = self.partial_batch(evidence)
        if self.world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.SUM)  # the sum-out over the split variable
        return self.slices[0].normalise_batch(total)
