#!/usr/bin/env python
"""bench.py — headline benchmark of the FastWindingNumber hot path (BASELINE.json: winding queries/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): is_inside classification of the 512^3 cell-centred lattice over
[-1.1, 1.1]^3 (134 217 728 queries) against an icosahedron midpoint-subdivided 8x (1 310 720 triangles), beta = 2,
order 2. One "step" = one pass of the query path over the whole lattice (the tree is built once, before the timed
region; its build time is reported beside the throughput). With N GPUs the lattice is cut into N z-slabs, the tree is
built on rank 0 and broadcast once with NCCL; there is no collective on the query path (weak/strong: total work fixed).

The JSON line carries: value (device-resident, CUDA-event timed, max over ranks), e2e (through the public API with a
host output buffer, D2H inside the timed region), roofline (FP32 FMA pipe: algorithmic flops from the traversal's own
counters / measured FMA peak; plus HBM GB/s), cpu_baseline (the oracle restatement on the host cores, bounded sample),
clocks, gpu_launches.  `--impl reference` times the oracle restatement (the reference binary cannot be built here, see
DESIGN.md) on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "winding queries/sec"
UNIT = "Gqueries/s"
# host threads of the CPU arm: every core this process may run on, whatever the launcher exported (torchrun sets
# OMP_NUM_THREADS=1 for its workers, which made the round-1 reference arm single-threaded at N > 1)
HOST_THREADS = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def config_dict(args, name, n_total):
    """The same `config` for both arms (the driver compares them key by key)."""
    return {"workload": name, "queries_per_step": n_total,
            "sharding": f"tile layers (8 z-planes) round-robin over {args.gpus} rank(s), tree built on rank 0 and broadcast; CPU arm: rank 0 only",
            "l2": "GPU arm: flushed between timed steps (256 MiB fill); CPU arm: not applicable",
            "gpu_arm": {"leaf_size": args.leaf_size, "hierarchy": args.hierarchy, "tiled": os.environ.get("WN_TILE", "1") != "0"},
            "cpu_arm": "oracle restatement of the reference algorithm (reference binary unbuildable here: Eigen/TBB/WindingNumber sources absent)"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, help="BASELINE config (2 = headline)")
    ap.add_argument("--grid", type=int, default=0, help="override lattice resolution (debug)")
    ap.add_argument("--subdiv", type=int, default=-1, help="override sphere subdivision level (debug)")
    ap.add_argument("--leaf-size", type=int, default=int(os.environ.get("WN_BENCH_LEAF", "4")), help="max triangles per leaf (ignored by the reference hierarchy)")
    ap.add_argument("--hierarchy", default=os.environ.get("WN_BENCH_HIERARCHY", "reference"), choices=["lbvh", "kd", "kd_sah", "reference"],
                    help="reference (default): the reference builder's own tree built on the GPU, results match the reference algorithm to float "
                         "rounding; kd_sah: k-d hierarchy with SAH-guided cuts and 4-triangle leaves (~2 %% faster queries, results ~1e-3 off); "
                         "kd: balanced k-d; lbvh: Morton/Karras (1.9 ms build)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target duration of the CPU baseline sample")
    return ap.parse_args()


def workload(args):
    from lagrange_b200 import primitive as prim

    level = args.subdiv if args.subdiv >= 0 else 8
    n = args.grid or 512
    V, F = prim.generate_subdivided_sphere("icosahedron", level)
    origin = np.full(3, -1.1, dtype=np.float32)
    spacing = np.full(3, 2.2 / n, dtype=np.float32)
    dims = np.array([n, n, n], dtype=np.int64)
    name = f"cfg2: icosphere L{level} ({len(F)} tris), {n}^3 cell-centred lattice over [-1.1,1.1]^3, is_inside, beta=2, order 2"
    return V, F, (origin, spacing, dims), name


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                smax.append(float(p[2]))
                power.append(float(p[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)), "samples": len(sm),
                "reasons": sorted(reasons)}


def cpu_baseline(V, F, lattice, seconds, want_build=True):
    """The oracle restatement (oracle/wn_oracle.cpp) on the host cores, bounded strided sample of the same lattice."""
    import oracle

    origin, spacing, dims = lattice
    total = int(dims[0] * dims[1] * dims[2])
    t0 = time.perf_counter()
    ref = oracle.RefEngine(V, F)
    build_s = time.perf_counter() - t0
    cores = HOST_THREADS
    # pilot to size the sample
    stride0 = max(1, total // 200_000) | 1
    t0 = time.perf_counter()
    ref.grid(origin, spacing, dims, first=0, stride=stride0, nthreads=cores)
    pilot = time.perf_counter() - t0
    pilot_n = (total + stride0 - 1) // stride0
    rate = pilot_n / max(pilot, 1e-6)
    n = int(min(total, max(pilot_n, rate * seconds)))
    stride = max(1, total // n) | 1
    t0 = time.perf_counter()
    out = ref.grid(origin, spacing, dims, first=0, stride=stride, nthreads=cores)
    dt = time.perf_counter() - t0
    n = len(out)
    return ref, {"value": n / dt / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
                 "sample": f"{n} of {total} lattice points (every {stride}-th in x-fastest order) in {dt:.2f} s; oracle restatement of the "
                           f"reference algorithm (4-ary SAH BVH, float32), OpenMP over queries; tree build {build_s:.2f} s on 1 thread",
                 "build_s": build_s, "seconds": dt}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle restatement; the reference binary is unbuildable here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    V, F, lattice, name = workload(args)
    import oracle

    origin, spacing, dims = lattice
    total = int(dims[0] * dims[1] * dims[2])
    t0 = time.perf_counter()
    ref = oracle.RefEngine(V, F)
    build_s = time.perf_counter() - t0
    cores = HOST_THREADS
    stride0 = max(1, total // 100_000) | 1
    t0 = time.perf_counter()
    ref.grid(origin, spacing, dims, stride=stride0, nthreads=cores)
    rate = ((total + stride0 - 1) // stride0) / (time.perf_counter() - t0)
    per_step = max(50_000, int(rate * 4.0))  # ~4 s per step
    stride = max(1, total // per_step) | 1
    times = []
    n = 0
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        out = ref.grid(origin, spacing, dims, first=it % stride, stride=stride, nthreads=cores)
        dt = time.perf_counter() - t0
        n = len(out)
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = n / (ms * 1e-3) / 1e9
    sample = f"{n} of {total} lattice points per step (stride {stride}), {cores} OpenMP threads; tree build {build_s:.2f} s"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, name, total),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def profile_traffic():
    """DRAM bytes (read + write) per launch of the dominant kernel, k_tile_query, from the committed `ncu --set full` summary
    (profiles/r1q_tile_plan_query_ncu_full.txt: one launch = 131072 tiles = 67.1 M queries). The tree (158 MB with 4-triangle leaves) is part of it:
    every launch streams the records it touches from HBM once; the algorithmic output is 1 byte per query."""
    path = os.path.join(ROOT, "profiles", "r1q_tile_plan_query_ncu_full.txt")
    try:
        rd = wr = None
        in_query = False
        for ln in open(path):
            if ln.startswith("## "):
                in_query = "k_tile_query" in ln
            elif in_query and ln.startswith("dram__bytes_read.sum "):
                rd = float(ln.split()[-1]) * 1e6
            elif in_query and ln.startswith("dram__bytes_write.sum "):
                wr = float(ln.split()[-1]) * 1e6
        return None if rd is None or wr is None else {"bytes_per_launch": rd + wr, "queries_per_launch": 131072 * 512,
                                                      "source": "profiles/r1q_tile_plan_query_ncu_full.txt (k_tile_query)"}
    except Exception:
        return None



def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch

    import lagrange_b200 as lb
    from lagrange_b200.distributed import interleaved_layers, replicate_engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    V = F = None
    origin = np.full(3, -1.1, dtype=np.float32)
    n = args.grid or 512
    spacing = np.full(3, 2.2 / n, dtype=np.float32)
    dims = np.array([n, n, n], dtype=np.int64)
    name = None
    eng = None
    build_info = {}
    if rank == 0:
        V, F, (origin, spacing, dims), name = workload(args)
        # warm-up build on a small mesh: loads the build kernels (CUDA lazy module loading) so that build_ms is kernel time
        lb.FastWindingNumber(*lb.primitive.generate_subdivided_sphere("icosahedron", 4), leaf_size=args.leaf_size, hierarchy=args.hierarchy).close()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng = lb.FastWindingNumber(V, F, leaf_size=args.leaf_size, hierarchy=args.hierarchy)
        torch.cuda.synchronize()
        build_wall_ms = 1e3 * (time.perf_counter() - t0)
        build_info = eng.info
        build_info["build_wall_ms"] = build_wall_ms
    bcast_ms = 0.0
    if world > 1:
        # warm the communicator first (NCCL sets up its channels lazily on the first collective: ~40-70 ms that are not the
        # cost of moving a tree), then time the replication itself
        warm = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
        dist.broadcast(warm, src=0)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        eng = replicate_engine(eng, src=0)
        torch.cuda.synchronize()
        dist.barrier()
        bcast_ms = 1e3 * (time.perf_counter() - t0)

    nz = int(dims[2])
    # rank r classifies the tile layers r, r+N, r+2N, ... (8 z-planes each) in ONE strided call: same mix of work on every
    # rank (contiguous z-slabs of a sphere load-imbalance: measured 6.5x at 8 GPUs), results compact in the rank's buffer
    ranges = interleaved_layers(nz, rank, world, depth=8)
    per_layer = int(dims[0] * dims[1])
    n_local = per_layer * sum(b - a for a, b in ranges)
    layers = (rank, world) if world > 1 else None
    n_total = int(dims[0] * dims[1] * dims[2])
    out_dev = torch.empty(max(n_local, 1), dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step_device():
        eng.query_grid(origin, spacing, dims, want_inside=True, out_inside=out_dev, layers=layers)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in ev:
        flush.fill_(1)  # evict L2 between timed iterations
        a.record()
        step_device()
        b.record()
    barrier()
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    clocks = sampler.stop() if rank == 0 else None
    ms_local = float(np.mean(ms_steps))
    t = torch.tensor([ms_local], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item())
    value = n_total / (ms_step * 1e-3) / 1e9

    # ---- e2e: the public API with a HOST output buffer (pinned), D2H inside the timed region --------------------------------
    # (bit-packed: 1 bit per query, WN_QUERY_OUT_BITS; the copies of finished batches overlap the next batch)
    out_host = torch.empty(max((n_local + 7) // 8, 1), dtype=torch.uint8).pin_memory().numpy()
    def step_host():
        eng.query_grid(origin, spacing, dims, want_inside=True, out_inside=out_host, layers=layers, bits=True)

    for _ in range(2):
        step_host()
    barrier()
    e2e_times = []
    for _ in range(max(3, min(args.steps, 5))):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step_host()
        checksum = int(out_host[::509].sum())  # the caller reads the result
        e2e_times.append(time.perf_counter() - t0)
    t = torch.tensor([float(np.mean(e2e_times)) * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = n_total / (e2e_ms * 1e-3) / 1e9
    inside_local = torch.tensor([int(out_dev[:n_local].sum().item())], dtype=torch.int64, device="cuda")
    if dist is not None:
        dist.all_reduce(inside_local)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline (rank 0): flops EXECUTED by the timed kernels (their own counters), FMA peak measured live -----------------
    tiled = os.environ.get("WN_TILE", "1") != "0"
    def summed_stats(tiling):
        tot = {}
        for za, zb in ranges:
            st = eng.query_stats_grid(origin, spacing, dims, z_range=(za, zb), tiling=tiling)
            for k, v in st.items():
                tot[k] = tot.get(k, 0) + v
        return tot

    executed = summed_stats(tiled)
    per_point = summed_stats(False)
    nq = max(1, executed["queries"])
    # SURVEY.md 8(d) units (10 / 83 / 75 flop) + the far-field interpolation of the tiled path: 64 FMA + ~60 for the weights
    interp_flops = (2 * 64 + 60) if tiled else 0
    flops_per_query = executed["algorithmic_flops"] / nq + interp_flops
    import ctypes

    from lagrange_b200 import _capi

    tf, pms = ctypes.c_float(), ctypes.c_float()
    _capi.check(_capi.lib().wn_debug_fma_peak(local_rank, 1 << 14, ctypes.byref(tf), ctypes.byref(pms)))
    fma_peak = float(tf.value)
    achieved_tflops = flops_per_query * n_local / (ms_local * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    ncu_traffic = profile_traffic()
    algo_bytes = n_local * 1  # implicit lattice in, 1 byte out per query (12 B in only for explicit points)
    roofline = {
        "bound": "fp32_fma", "achieved": achieved_tflops, "peak": fma_peak, "unit": "TFLOP/s", "frac": achieved_tflops / fma_peak,
        "peak_source": "live FMA microbenchmark k_fma_peak (MEASURED_PEAKS.json has no FP32 CUDA-core figure)",
        "note": "achieved = flops executed by the timed kernels (SURVEY 8(d) units from the kernels' own counters) / event time; "
                "the per-point counts of the reference algorithm on the same tree are under reference_algorithm",
        "flops_per_query": flops_per_query, "tests_per_query": executed["node_tests"] / nq, "evals_per_query": executed["far_field_evals"] / nq,
        "exact_tris_per_query": executed["exact_triangles"] / nq, "lane_utilisation": executed["node_tests"] / max(1, executed["lane_slots"]),
        "reference_algorithm": {"flops_per_query": per_point["algorithmic_flops"] / nq, "tests_per_query": per_point["node_tests"] / nq,
                                "evals_per_query": per_point["far_field_evals"] / nq, "exact_tris_per_query": per_point["exact_triangles"] / nq,
                                "equivalent_tflops": per_point["algorithmic_flops"] / nq * n_local / (ms_local * 1e-3) / 1e12},
        "traffic": None if ncu_traffic is None else ncu_traffic["bytes_per_launch"], "traffic_detail": ncu_traffic,
        "hbm": {"achieved_gbs": algo_bytes / (ms_local * 1e-3) / 1e9, "peak_gbs": hbm_peak, "peak_source": hbm_src,
                "frac": algo_bytes / (ms_local * 1e-3) / 1e9 / hbm_peak},
    }

    cpu = None
    if not args.no_cpu_baseline:
        _, cpu = cpu_baseline(V, F, (origin, spacing, dims), args.cpu_seconds if world == 1 else min(args.cpu_seconds, 6.0))
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, name, n_total),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 60, "d2h_bytes_per_step": (n_total + 7) // 8, "ms_per_step": e2e_ms,
                "api": "FastWindingNumber.query_grid(bits=True) -> wn_query_grid[_strided](WN_QUERY_OUT_BITS) with a pinned HOST output buffer, 1 bit per "
                       "query (lattice is implicit: 60-byte descriptor in)"},
        # per step and rank: tiled = (k_tile_plan + k_tile_query) per batch of <= 131072 tiles (the probe that picks the path runs
        # once, in the warm-up, and is remembered per lattice); generic = one k_query
        "gpu_launches": args.steps * (2 * max(1, -(-(-(-n_local // 512)) // 131072)) if tiled else 1),
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "build": {k: build_info.get(k) for k in ("build_ms", "build_ms_morton", "build_ms_sort", "build_ms_hierarchy", "build_ms_moments",
                                                 "build_ms_pack", "build_wall_ms", "num_entries", "tree_bytes", "max_depth")},
        "tree_broadcast_ms": bcast_ms, "inside_count": int(inside_local.item()), "checksum": checksum,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
