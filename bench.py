#!/usr/bin/env python
"""bench.py -- junction-tree evidence queries/sec on B200 (BASELINE.json metric), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W]                (N>1: launched under torch.distributed.run)
  python bench.py --impl reference ...                               (the reference arm: CPU, see below)

Workload (BASELINE.json configs[1], "C2"): synthetic Alarm-sized random DAG, 37 nodes with 2-4 states,
2^20 evidence rows per GPU (9 observed variables, ancestral samples so P(e) > 0).  One *step* = one pass of
the hot path over the batch: for every row changeEvidence(row) and posterior [v] for all 37 variables
(107 doubles out per row).  Rows are independent, so N GPUs each take their own 2^20 rows with no
communication (weak scaling); value = all rows of all ranks / max-over-ranks device time.

  value      evidence and output already resident in HBM (HB_DEVICE_PTRS), CUDA events on the launching stream
  e2e        the same call with pinned HOST buffers (HB_HOST_PTRS): H2D of the evidence matrix and D2H of the
             posterior matrix inside the timed region
  roofline   the fused product/sum-out kernels of the pass (prodsum_kernel<K>, the dominant kernel family):
             algorithmic bytes = 8 * (per-row table entries read + written) * rows, time = CUDA events
             around those launches on the same stream
  cpu_baseline / --impl reference
             the reference is Haskell and no GHC exists on this image: the CPU arm is the C++
             restatement of the reference algorithm (oracle/, kind "port"), rows spread over all host
             cores, on a bounded row sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# stdout carries exactly one JSON line: NCCL's own messages go to stderr (NCCL_DEBUG itself is left as the caller set it)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS_PER_GPU = 1 << 20
METRIC = "junction-tree evidence queries/sec"
UNIT = "queries/s"


def workload():
    from hbayes_b200 import synth

    net, observed = synth.config_c2()
    return net, observed


def config_dict(n_gpus, rows_per_gpu):
    return {
        "workload": "C2: synthetic Alarm-sized random DAG (37 nodes, 2-4 states, 50 edges, net seed 37001), "
                    "%d evidence rows per GPU, 9 observed variables (seed 37002/37003); per row changeEvidence + "
                    "posterior of all 37 variables (107 doubles)" % rows_per_gpu,
        "rows_per_gpu": rows_per_gpu,
        "n_vars": 37,
        "out_doubles_per_row": 107,
        "sharding": "rows x%d, no collective" % n_gpus,
        "l2_policy": "inputs larger than L2: every pass reads 38.8 MB of evidence and writes 0.9 GB of posteriors per GPU (L2 = 126 MB); "
                     "nothing else goes through HBM (per-row tables live in shared memory): no flush needed",
    }


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe): NVML polled every few
    milliseconds (the same counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints); falls back
    to spawning nvidia-smi when the NVML binding is not importable."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.mhz, self.max_mhz, self.reasons, self.stop_flag, self.thread = index, [], None, set(), False, None
        self.nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _poll_nvml(self):
        n = self.nvml
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self.stop_flag:
            try:
                self.mhz.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                try:
                    mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for name, bit in bits.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def _poll_smi(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                if len(f) >= 6:
                    self.mhz.append(float(f[0]))
                    self.max_mhz = float(f[1])
                    for i, name in enumerate(self.NAMES):
                        if f[2 + i].lower().startswith("active"):
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def start(self):
        self.thread = threading.Thread(target=self._poll_nvml if self.nvml else self._poll_smi, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        return {"sm_mhz": float(np.median(self.mhz)) if self.mhz else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.mhz), "source": "nvml" if self.nvml else "nvidia-smi"}


def cpu_port_rate(n_rows, threads, want_posteriors=False):
    """queries/s of the oracle (restatement of the reference algorithm) on `threads` host threads."""
    from oracle import OracleJT, OracleNet
    from hbayes_b200 import synth

    net, observed = workload()
    onet = OracleNet.from_arrays(net.dims, net.scopes, net.values)
    ojt = OracleJT(onet)
    ev = synth.config_c2_evidence(net, observed, n_rows)
    ojt.posteriors_batch(ev[: max(threads, 8)], n_threads=threads)  # touch code and memory once
    t = time.perf_counter()
    post = ojt.posteriors_batch(ev, n_threads=threads)
    rate = n_rows / (time.perf_counter() - t)
    return (rate, post) if want_posteriors else rate


def run_reference(args):
    """The reference arm: the reference's algorithm on the box's host cores (oracle port; GHC is absent)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rows = int(min(1 << 16, max(1024, threads * 330 * 4)))  # ~4 s per step at ~330 queries/s/core
    rates = []
    for i in range(args.warmup + args.steps):
        r = cpu_port_rate(rows, threads)
        if i >= args.warmup:
            rates.append(r)
    value = len(rates) / sum(1.0 / r for r in rates)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * rows / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": dict(config_dict(args.gpus, ROWS_PER_GPU), sample_rows_per_step=rows),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d rows of the C2 evidence matrix per step, rows split over %d threads; C++ restatement of "
                                   "the reference algorithm (oracle/), not GHC" % (rows, threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def pcie_peaks(pinned_out, d_out, pinned_ev, d_ev, stream):
    """pinned host<->device copy bandwidth of the very buffers the e2e step moves (GB/s, best of 3, CUDA events)"""
    import torch

    def best(fn, nbytes):
        ms = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return nbytes / (min(ms) / 1e3) / 1e9

    return {"d2h_GBps": best(lambda: pinned_out.copy_(d_out, non_blocking=True), pinned_out.numel() * 8),
            "h2d_GBps": best(lambda: d_ev.copy_(pinned_ev, non_blocking=True), pinned_ev.numel())}


def other_configs():
    """short runs of the other BASELINE.json configs on this GPU (bench/configs.py): C1 latency and batch, C3, C5"""
    sys.path.insert(0, os.path.join(ROOT, "bench"))
    import configs

    res = {}
    for name, fn in (("C1", configs.c1), ("C3", configs.c3), ("C5", lambda: configs.c5(64))):
        t = time.perf_counter()
        try:
            res[name] = fn()
        except Exception as e:  # a secondary measurement must never cost the headline line
            res[name] = {"error": repr(e)[:300]}
        res[name]["wall_s"] = time.perf_counter() - t
    return res


def c4_leg(rank, world, local):
    """BASELINE config C4 through the C ABI (collective: every rank calls it): one 2^30-entry clique table (8 GiB) split along
    log2(world) binary parents, one slice per GPU, the sum over the split variables taken by ncclAllReduce inside the
    library (hb_jt_compile_split_nccl); compared with the cutset form (whole-program conditioning + one torch all-reduce)."""
    import torch
    import torch.distributed as dist

    import hbayes_b200 as hb
    from hbayes_b200 import synth
    from hbayes_b200.dist import max_over_ranks
    from hbayes_b200.sharded import SplitJunctionTree, broadcast_unique_id

    k = world.bit_length() - 1
    if world < 2 or (1 << k) != world or k > 3:
        return None
    net, split = synth.config_c4(29)
    split = split[:k]
    pnet = hb.BayesianNetwork.from_tables_procedural(net.dims, net.scopes, net.values, net.proc_seed)
    rows = 4
    rng = np.random.RandomState(5)
    ev = np.full((rows, net.n), -1, np.int8)
    ev[:, net.n - 1] = rng.randint(0, 2, rows)  # the child is observed: every table of the query is per-row
    d_ev = torch.from_numpy(ev).cuda()
    t0 = time.perf_counter()
    sjt = SplitJunctionTree(pnet, split, rank=rank, world=world, device=local, nccl_id=broadcast_unique_id(rank))
    n_spec = sjt.specialise([net.n - 1])  # generated step kernels up front (NVRTC or the cubin cache), not in the middle of the timed passes
    out = sjt.posteriors_batch(d_ev)  # compiles, generates this GPU's slice of the table on the device, first pass
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    dist.barrier()
    steps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = sjt.posteriors_batch(d_ev, out)
    e1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1) / steps, device="cuda")
    st = sjt.slices[0].last_stats()
    res = out.cpu().numpy()
    del sjt
    torch.cuda.empty_cache()
    cut = SplitJunctionTree(pnet, split, rank=rank, world=world, device=local)  # the cutset form on the same slices
    ref = cut.posteriors_batch(ev).cpu().numpy()
    off = np.concatenate([[0], np.cumsum(net.dims)])
    gathered = [None] * world
    dist.all_gather_object(gathered, res.tobytes())
    return {
        "config": "C4: 1 child + 29 binary parents, CPT 2^30 entries (8 GiB), split along parents %s into %d slices, one per GPU; "
                  "exchange = ncclAllReduce inside libhbayes_cuda (hb_jt_compile_split_nccl)" % (split, world),
        "n_gpus": world, "rows": rows, "ms_per_pass": ms, "queries_per_s": rows / ms * 1e3, "setup_s_incl_slice_generation": setup_s,
        "launches_per_pass": st["launches"], "specialised_kernels": n_spec, "per_row": {kk: st[kk] for kk in ("row_entries", "prod_per_row", "row_read_per_row", "row_write_per_row")},
        "per_gpu_algorithmic_GBps": 8.0 * (st["row_read_per_row"] + st["row_write_per_row"]) * rows / (ms / 1e3) / 1e9,
        "max_abs_deviation_of_marginal_sums_from_1": float(max(abs(res[:, off[v]:off[v + 1]].sum(1) - 1).max() for v in range(net.n))),
        "all_ranks_identical": len(set(gathered)) == 1,
        "max_rel_err_vs_cutset_form": float((np.abs(res - ref) / np.maximum(np.abs(ref), 1e-15)).max()),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=ROWS_PER_GPU, help="evidence rows per GPU (default: the named workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import hbayes_b200 as hb
    from hbayes_b200 import synth
    from hbayes_b200.dist import max_over_ranks, shard_rows, whole_job_rate

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    # stdout carries exactly one JSON line: whatever a library prints there meanwhile (NCCL's version banner at
    # NCCL_DEBUG=VERSION, for one) is sent to stderr at the file-descriptor level; the environment is left alone
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W = max(args.warmup, 3)
    rows = args.rows

    net, observed = workload()
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    jt = hb.createJunctionTree(hb.nodeComparisonForTriangulation, pnet, device=local)
    stream = torch.cuda.current_stream()
    jt.set_stream(stream.cuda_stream)
    jt.set_timing(True)
    # The kernels generated for this (network, observed set) are built here, next to createJunctionTree, not inside a timed
    # step: NVRTC the first time on a box, the cubin cache (~/.cache/hbayes_b200) afterwards.
    t0 = time.perf_counter()
    n_spec = jt.specialise(observed)
    specialise_s = time.perf_counter() - t0
    info = jt.kernel_info()
    # every rank draws its own rows from the same stream of evidence (seed offset by rank)
    ev_host = synth.evidence_matrix(net, rows, observed, 37003 + rank)
    ev_pinned = torch.from_numpy(ev_host).pin_memory()
    out_pinned = torch.empty((rows, pnet.width), dtype=torch.float64, pin_memory=True)
    d_ev = ev_pinned.cuda(non_blocking=True)
    d_out = torch.empty((rows, pnet.width), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """exactly `steps` calls bracketed by barrier + synchronize; CUDA events on the launching stream; max over ranks"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), device="cuda")

    dev_step = lambda: jt.posteriors_batch(d_ev, d_out)
    e2e_step = lambda: jt.posteriors_batch(ev_pinned, out_pinned)

    for _ in range(W):
        dev_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(dev_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    stats = jt.last_stats()
    # kernel timing of one more (untimed) pass, CUDA events inside the library on the same stream
    dev_step()
    tm = jt.last_timing()
    # strong scaling: the SAME 2^20 rows sharded over the ranks (BASELINE config 2 as written); weak scaling is `value`
    lo, hi = shard_rows(rows, rank, world)
    d_ev_s, d_out_s = d_ev[: hi - lo], d_out[: hi - lo]
    strong_step = lambda: jt.posteriors_batch(d_ev_s, d_out_s)
    for _ in range(W):
        strong_step()
    ms_strong = timed(strong_step, args.steps)
    dev_step()  # d_out holds the full batch again for the parity check below
    for _ in range(2):
        e2e_step()
    e2e_steps = max(2, min(args.steps, 5))
    ms_e2e = timed(e2e_step, e2e_steps)
    barrier()  # every rank measures its copies at the same time: the denominator is the rate the box sustains for N concurrent transfers
    pcie = pcie_peaks(out_pinned, d_out, ev_pinned, d_ev, stream)
    jt.posteriors_batch(ev_pinned, out_pinned)  # out_pinned holds this rank's answers again (the copy test overwrote it)
    torch.cuda.synchronize()
    # the same call from ordinary (pageable) host memory: what a caller that did not pin its buffers gets
    out_pageable = np.empty((rows, pnet.width), np.float64)
    jt.posteriors_batch(ev_host, out_pageable)
    ms_pageable = timed(lambda: jt.posteriors_batch(ev_host, out_pageable), 2)
    c4 = None
    if world > 1 and not args.no_other_configs:
        keep = (d_out[: 1 << 16].clone(), out_pinned[: 1 << 16].clone())  # what the parity check below needs
        full_equal = bool(torch.equal(out_pinned, d_out.cpu()))
        del d_out, d_ev, d_ev_s, d_out_s
        torch.cuda.empty_cache()
        try:
            c4 = c4_leg(rank, world, local)
        except Exception as e:  # a secondary measurement must never cost the headline line
            c4 = {"error": repr(e)[:400]}
        d_out, out_pinned_head = keep

    if rank == 0:
        value = whole_job_rate([rows] * world, args.steps, ms)
        e2e_value = whole_job_rate([rows] * world, e2e_steps, ms_e2e)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        kernel_ms = tm["prodsum_ms"]
        on_chip = bool(info["on_chip"])
        if on_chip:  # one launch per pass; HBM sees the evidence rows and the posteriors only
            algo_bytes = float(rows) * (len(net.dims) + 8.0 * pnet.width)
            kernel_name = ("hb_tree (generated with NVRTC, csrc/jit.cpp): changeEvidence + collect + distribute + posterior of all 37 variables "
                           "for tiles of %d rows, per-row tables in shared memory; 1 launch per pass" % info["rows_per_cta"])
        else:
            algo_bytes = 8.0 * (stats["row_read_per_row"] + stats["row_write_per_row"]) * rows
            kernel_name = "the product/sum-out launches of one pass (per-step kernels)"
        achieved = algo_bytes / (kernel_ms / 1e3) / 1e9
        traffic, traffic_note = None, "no ncu capture for this row count"
        try:  # dram__bytes_read.sum + dram__bytes_write.sum of the same launch (ncu --set full, profiles/r2_traffic.json)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
            if rows == tj["rows"] and on_chip == bool(tj["on_chip"]):
                traffic, traffic_note = tj["traffic_bytes"], tj["source"]
        except Exception:
            pass
        sm_mhz = float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))
        n_sm = torch.cuda.get_device_properties(local).multi_processor_count
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(world, rows),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(ev_host.nbytes), "d2h_bytes_per_step": int(out_pinned.nbytes),
                    "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps,
                    "roofline": {"bound": "pcie d2h", "achieved": out_pinned.nbytes / (ms_e2e / e2e_steps / 1e3) / 1e9, "peak": pcie["d2h_GBps"],
                                 "unit": "GB/s", "frac": out_pinned.nbytes / (ms_e2e / e2e_steps / 1e3) / 1e9 / pcie["d2h_GBps"],
                                 "peak_source": "pinned cudaMemcpyAsync of the same %d-byte result matrix in this run (rank 0, all %d ranks copying at the "
                                                "same time), best of 3; H2D %.1f GB/s" % (out_pinned.nbytes, world, pcie["h2d_GBps"])},
                    "pageable": {"value": whole_job_rate([rows] * world, 2, ms_pageable), "ms_per_step": ms_pageable / 2,
                                 "note": "same call, caller buffers NOT pinned (numpy arrays)"}},
            "gpu_launches": int(stats["launches"]) * args.steps,
            "specialised_kernels": n_spec, "specialise_s": specialise_s,
            "kernel_origin": {-1: "none", 0: "process", 1: "disk cache", 2: "nvrtc"}.get(info["origin"], "?"),
            "scaling_strong": {"value": float(rows) * args.steps / (ms_strong / 1e3), "unit": UNIT, "rows_total": rows, "rows_per_gpu": hi - lo,
                               "ms_per_step": ms_strong / args.steps,
                               "note": "BASELINE config 2 as written: the same 2^20 rows sharded over the ranks (total work fixed); `value` is the weak-scaling figure"},
            "clocks": clocks,
            "roofline": {
                "bound": "hbm", "kernel": kernel_name,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "algorithmic_bytes_per_pass": algo_bytes, "kernel_ms_per_pass": kernel_ms, "traffic": traffic, "traffic_note": traffic_note,
                "pass_ms": tm,
            },
        }
        if on_chip:
            # HBM is no longer what binds this path: per row the kernel moves smem_loads + smem_stores table entries through
            # shared memory and does fp64_mul + fp64_add operations; both against this GPU's nominal rates at the SM clock
            # sampled during the timed region
            smem_bytes = 8.0 * (info["smem_loads_per_row"] + info["smem_stores_per_row"]) * rows
            smem_peak = 128.0 * n_sm * sm_mhz * 1e6 / 1e9
            flops = float(info["fp64_mul_per_row"] + info["fp64_add_per_row"]) * rows
            fp64_peak = 64.0 * n_sm * sm_mhz * 1e6 / 1e12
            line["roofline"]["note"] = ("this kernel keeps every per-row table on chip: its HBM fraction is small by design (round 1 moved "
                                        "61.7 KB per query through HBM, this one %.2f KB); the binding resources are below" % (algo_bytes / rows / 1e3))
            line["roofline"]["binding"] = {
                "shared_memory": {"achieved": smem_bytes / (kernel_ms / 1e3) / 1e9, "peak": smem_peak, "unit": "GB/s",
                                  "frac": smem_bytes / (kernel_ms / 1e3) / 1e9 / smem_peak,
                                  "per_row": {"loads": info["smem_loads_per_row"], "stores": info["smem_stores_per_row"]},
                                  "peak_source": "128 B/clk/SM x %d SMs x %.0f MHz (sampled)" % (n_sm, sm_mhz)},
                "fp64": {"achieved": flops / (kernel_ms / 1e3) / 1e12, "peak": fp64_peak, "unit": "Tflop/s (separate multiply and add, no FMA)",
                         "frac": flops / (kernel_ms / 1e3) / 1e12 / fp64_peak,
                         "per_row": {"mul": info["fp64_mul_per_row"], "add": info["fp64_add_per_row"]},
                         "note": "operations the kernel executes: the `0 +` each reference sum starts with is left out while no "
                                 "potential is negative, -0.0 or NaN (0 + x == x bit for bit; csrc/jit.cpp jit_zero_add_elidable)",
                         "peak_source": "64 fp64 lanes/clk/SM x %d SMs x %.0f MHz (sampled)" % (n_sm, sm_mhz)},
                "kernel": {k: info[k] for k in ("rows_per_cta", "threads_per_cta", "levels", "smem_entries_per_row", "smem_bytes")},
            }
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sample = int(min(1 << 16, max(1024, threads * 330 * 12)))
            rate, want = cpu_port_rate(sample, threads, want_posteriors=True)
            # the rows the CPU port just answered double as a full-size parity check of BOTH timed paths (same engine state as
            # the timed steps): the device-resident output and the host buffer the e2e call filled, bit for bit, NaN == NaN

            def same(got):
                return bool(np.array_equal(np.isnan(got), np.isnan(want)) and got[~np.isnan(got)].tobytes() == want[~np.isnan(want)].tobytes())

            line["parity"] = {"rows_checked": sample, "bit_identical_to_cpu_port": same(d_out[:sample].cpu().numpy()),
                              "e2e_host_buffer_bit_identical_to_cpu_port": same(out_pinned[:sample].numpy()),
                              "e2e_pageable_buffer_bit_identical_to_cpu_port": same(out_pageable[:sample]),
                              "e2e_equals_device_path_all_rows": full_equal if c4 is not None else bool(torch.equal(out_pinned, d_out.cpu()))}
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "first %d rows of the same evidence matrix, rows split over %d host threads; C++ "
                                              "restatement of the reference algorithm (oracle/), not GHC" % (sample, threads)}
        if c4 is not None:
            line["other_configs"] = {"C4": c4}
        if world == 1 and not args.no_other_configs:
            del d_out, d_ev, out_pinned, ev_pinned
            torch.cuda.empty_cache()
            line["other_configs"] = other_configs()
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
