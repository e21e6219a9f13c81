"""Split clique (BASELINE config C4): a clique/CPT table too large for one GPU is split along
`split_vars`; every rank owns the slices rank, rank + world, ... of the state combinations of those variables,
evaluates the junction-tree query restricted to each of its slices (libhbayes_cuda stores only that slice of
the big table), and the unnormalised results are summed: locally over the rank's slices, then across ranks
with ONE all-reduce (NCCL over NVLink through torch.distributed) -- the sum over the split variable is the only
exchange step of the path.  A final kernel normalises.

Not bit-identical to the unsplit engine: the sum over the split variable is taken last instead of at its place
in the elimination order (agreement ~1e-15 relative; tests use 1e-12)."""
from __future__ import annotations

import itertools

import numpy as np

from .api import BayesianNetwork, JunctionTree, nodeComparisonForTriangulation


class SplitJunctionTree:
    def __init__(self, net: BayesianNetwork, split_vars, rank=0, world=1, device=0, cmp=nodeComparisonForTriangulation):
        self.net, self.split_vars, self.rank, self.world = net, list(split_vars), rank, world
        combos = list(itertools.product(*[range(net.dims[v]) for v in self.split_vars]))
        self.combos = combos[rank::world]
        self.slices = [JunctionTree(net, cmp=cmp, device=device, split=list(zip(self.split_vars, c))) for c in self.combos]

    def partial_batch(self, evidence: np.ndarray):
        """sum over this rank's slices of the unnormalised posteriors; a CUDA tensor [rows][sum dims]"""
        import torch

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

    def posteriors_batch(self, evidence: np.ndarray):
        """posterior of every variable for every row, replicated on all ranks"""
        import torch.distributed as dist

        total = self.partial_batch(evidence)
        if self.world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.SUM)  # the sum-out over the split variable
        return self.slices[0].normalise_batch(total)
