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

# stdout carries exactly one JSON line.  NCCL prints its banner ("NCCL version ...") to stdout at NCCL_DEBUG=VERSION, and
# only honours NCCL_DEBUG_FILE above that level: raise VERSION to WARN (same banner, no more) and send it to stderr.
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
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
        "l2_policy": "working set per step >> L2 (row arena 21 GB, output 0.9 GB per GPU): no flush needed",
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
        "dtype": "f64", "data": "synthetic", "config": config_dict(args.gpus, ROWS_PER_GPU),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d rows of the C2 evidence matrix per step, rows split over %d threads; C++ restatement of "
                                   "the reference algorithm (oracle/), not GHC" % (rows, threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=ROWS_PER_GPU, help="evidence rows per GPU (default: the named workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import hbayes_b200 as hb
    from hbayes_b200 import synth
    from hbayes_b200.dist import max_over_ranks, whole_job_rate

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
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
    # kernel-family timing of one more (untimed) pass, CUDA events inside the library on the same stream
    dev_step()
    tm = jt.last_timing()
    for _ in range(2):
        e2e_step()
    e2e_steps = max(2, min(args.steps, 5))
    ms_e2e = timed(e2e_step, e2e_steps)

    if rank == 0:
        value = whole_job_rate([rows] * world, args.steps, ms)
        e2e_value = whole_job_rate([rows] * world, e2e_steps, ms_e2e)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        algo_bytes = 8.0 * (stats["row_read_per_row"] + stats["row_write_per_row"]) * rows
        achieved = algo_bytes / (tm["prodsum_ms"] / 1e3) / 1e9
        n_prodsum = stats["launches"] - 2 * stats["tiles"]
        traffic, traffic_note = None, "no ncu capture for this row count"
        try:  # dram__bytes_read.sum + dram__bytes_write.sum summed over the same launches (ncu, profiles/r1_traffic.json)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
            if rows == ROWS_PER_GPU and abs(tj["algorithmic_bytes"] - algo_bytes) < 1e-6 * algo_bytes:
                traffic, traffic_note = tj["traffic_bytes"], tj["source"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(world, rows),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(ev_host.nbytes), "d2h_bytes_per_step": int(out_pinned.nbytes),
                    "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps},
            "gpu_launches": int(stats["launches"]) * args.steps,
            "specialised_kernels": jt.specialised_kernels(),
            "clocks": clocks,
            "roofline": {
                "bound": "hbm",
                "kernel": "the %d product/sum-out launches of one pass (per-step specialised kernels hb_step_<i> compiled with NVRTC on "
                          "the first warm-up pass, prodsum_kernel<K,V,D> / fused_kernel<KO,KI,V,DI> for the rest; 93%% of the pass; "
                          "aggregated: one launch = one schedule step over all rows)" % n_prodsum,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "algorithmic_bytes_per_pass": algo_bytes, "prodsum_ms_per_pass": tm["prodsum_ms"], "traffic": traffic,
                "traffic_note": traffic_note,
                "per_row": {"entries_read": stats["row_read_per_row"], "entries_written": stats["row_write_per_row"],
                            "product_entries": stats["prod_per_row"], "arena_entries": stats["row_entries"]},
                "pass_ms": tm,
            },
        }
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sample = int(min(1 << 16, max(1024, threads * 330 * 12)))
            rate, want = cpu_port_rate(sample, threads, want_posteriors=True)
            # the rows the CPU port just answered double as a full-size parity check of the timed GPU path (same engine
            # state as the timed steps: specialised kernels, CUDA graph): bit for bit, NaN == NaN
            got = d_out[:sample].cpu().numpy()
            same = bool(np.array_equal(np.isnan(got), np.isnan(want)) and got[~np.isnan(got)].tobytes() == want[~np.isnan(want)].tobytes())
            line["parity"] = {"rows_checked": sample, "bit_identical_to_cpu_port": same}
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "first %d rows of the same evidence matrix, rows split over %d host threads; C++ "
                                              "restatement of the reference algorithm (oracle/), not GHC" % (sample, threads)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
