#!/usr/bin/env python
"""bench.py — headline benchmark of the FastWindingNumber hot path (BASELINE.json: winding queries/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config {1..5}] [--mode {tree,exact}]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default workload (BASELINE.json configs[1]): is_inside classification of the 512^3 cell-centred lattice over
[-1.1, 1.1]^3 (134 217 728 queries) against an icosahedron midpoint-subdivided 8x (1 310 720 triangles), beta = 2, order 2,
on the reference builder's own hierarchy built on the GPU (WN_HIERARCHY_REFERENCE: results match the reference algorithm).
One "step" = one pass of the query path over the whole query set (the tree is built once, before the timed region; its build
time is reported beside the throughput). With N GPUs the queries are sharded (lattices: 8-plane tile layers dealt round-robin,
one call per rank; point sets: index ranges), the tree is built on rank 0 and broadcast once with NCCL; there is no collective
on the query path.
--config 1/3/4/5 run the other BASELINE configs through the same flow (cfg4/cfg5 are point sets: device-resident points for
`value`, pinned host points + H2D inside the timed region for `e2e`); --mode exact times the brute-force all-pairs mode (cfg5's
mesh, a bounded number of queries per step).

The JSON line carries: value (device-resident, CUDA-event timed, max over ranks), e2e (through the public API with HOST
buffers, copies inside the timed region, bit-packed result), roofline (FP32 FMA pipe: flops executed by the timed kernels from
their own counters / FMA peak measured live; plus HBM GB/s), cpu_baseline (the oracle restatement on the host cores, bounded
sample), clocks, gpu_launches, parity (strict-band check of the timed configuration against the restatement on a sample).
`--impl reference` times the oracle restatement (the reference binary cannot be built here) on all host threads.
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
# the CPU arm is compiled for the host it is timed on (-march=native, oracle/_native/; falls back to the shipped x86-64-v3 build)
os.environ.setdefault("WN_ORACLE_NATIVE", "1")

METRIC = "winding queries/sec"
UNIT = "Gqueries/s"
# host threads of the CPU arm: every core this process may run on, whatever the launcher exported (torchrun sets
# OMP_NUM_THREADS=1 for its workers, which made the round-1 reference arm single-threaded at N > 1)
HOST_THREADS = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
FOUR_PI = 4.0 * np.pi


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5], help="BASELINE config (2 = headline)")
    ap.add_argument("--mode", default="tree", choices=["tree", "exact"], help="exact: brute-force all-pairs mode (cfg5 mesh)")
    ap.add_argument("--grid", type=int, default=0, help="override lattice resolution (debug)")
    ap.add_argument("--subdiv", type=int, default=-1, help="override sphere subdivision level (debug)")
    ap.add_argument("--points", type=int, default=0, help="override the number of query points of cfg4 / cfg5 / exact mode (debug)")
    ap.add_argument("--leaf-size", type=int, default=int(os.environ.get("WN_BENCH_LEAF", "4")), help="max triangles per leaf (ignored by the reference hierarchy)")
    ap.add_argument("--hierarchy", default=os.environ.get("WN_BENCH_HIERARCHY", "reference"), choices=["lbvh", "kd", "kd_sah", "reference"],
                    help="reference (default): the reference builder's own tree built on the GPU, results match the reference algorithm to float "
                         "rounding; kd_sah: k-d hierarchy with SAH-guided cuts and 4-triangle leaves (~2 %% faster queries, results ~1e-3 off); "
                         "kd: balanced k-d; lbvh: Morton/Karras (1.9 ms build)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target duration of the CPU baseline sample")
    args = ap.parse_args()
    if args.mode == "exact":
        args.config = 5
    return args


def workload(args):
    """{name, V, F, kind: 'grid'|'points', lattice | points, n_total} of the selected BASELINE config."""
    from lagrange_b200 import primitive as prim

    cfg = args.config
    if cfg == 2:
        level = args.subdiv if args.subdiv >= 0 else 8
        n = args.grid or 512
        V, F = prim.generate_subdivided_sphere("icosahedron", level)
        lattice = (np.full(3, -1.1, dtype=np.float32), np.full(3, 2.2 / n, dtype=np.float32), np.array([n, n, n], dtype=np.int64))
        name = f"cfg2: icosphere L{level} ({len(F)} tris), {n}^3 cell-centred lattice over [-1.1,1.1]^3, is_inside, beta=2, order 2"
        return {"name": name, "V": V, "F": F, "kind": "grid", "lattice": lattice, "n_total": n ** 3}
    V, F = prim.config_mesh(cfg)
    if cfg in (1, 3):
        n = args.grid or (100 if cfg == 1 else 256)
        lattice = prim.lattice_for_bbox(*prim.mesh_bbox(V), n)
        what = "torus 100x50 centroid-fan" if cfg == 1 else "open non-manifold soup (torus 250x200 with holes, duplicates, flips)"
        name = f"cfg{cfg}: {what} ({len(F)} tris), {n}^3 cell-centred lattice over the bbox + 5 %, is_inside, beta=2, order 2"
        return {"name": name, "V": V, "F": F, "kind": "grid", "lattice": lattice, "n_total": n ** 3}
    if args.mode == "exact":
        n = args.points or (1 << 20)
        pts = prim.uniform_points_in_bbox(*prim.mesh_bbox(V), n, seed=0xC0FFEE05)
        name = f"cfg5 exact mode: torus 250x100 ({len(F)} tris) x {n} uniform points per step, brute-force all-pairs solid angle (75 flop per pair)"
        return {"name": name, "V": V, "F": F, "kind": "points", "points": pts, "n_total": n}
    if cfg == 4:
        n = args.points or (64 << 20)
        pts = prim.near_surface_points(V, F, n, seed=0xC0FFEE04)
        name = f"cfg4: octasphere L10 ({len(F)} tris), {n} near-surface jittered points in random order, is_inside, beta=2, order 2"
    else:
        n = args.points or (1 << 24)
        pts = prim.uniform_points_in_bbox(*prim.mesh_bbox(V), n, seed=0xC0FFEE05)
        name = f"cfg5: torus 250x100 ({len(F)} tris), {n} uniform points in the bbox + 5 %, is_inside, beta=2, order 2"
    return {"name": name, "V": V, "F": F, "kind": "points", "points": pts, "n_total": n}


def config_dict(args, name, n_total):
    """The same `config` for both arms (the driver compares them key by key)."""
    return {"workload": name, "queries_per_step": n_total, "mode": args.mode,
            "sharding": f"lattices: 8-plane tile layers dealt round-robin over {args.gpus} rank(s) (wn_query_grid_layers, one call per rank); "
                        "point sets: index ranges; tree built on rank 0 and broadcast; CPU arm: rank 0 only",
            "l2": "GPU arm: flushed between timed steps (256 MiB fill); CPU arm: not applicable",
            "gpu_arm": {"hierarchy": args.hierarchy, "leaf_size": 1 if args.hierarchy == "reference" else args.leaf_size,
                        "tiled": os.environ.get("WN_TILE", "1") != "0"},
            "cpu_arm": "oracle restatement of the reference algorithm (reference binary unbuildable here: Eigen/TBB/WindingNumber sources absent)"}


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


# ---- the CPU arm: the oracle restatement on the host cores --------------------------------------------------------------------
class CpuArm:
    """Times `count` queries of the workload (every stride-th one) with the restatement, all host threads."""

    def __init__(self, wl, mode):
        import oracle

        self.oracle = oracle
        self.wl = wl
        self.mode = mode
        t0 = time.perf_counter()
        self.ref = oracle.RefEngine(wl["V"], wl["F"]) if mode == "tree" else None
        self.build_s = time.perf_counter() - t0
        self.cores = HOST_THREADS

    def run(self, first, stride):
        wl = self.wl
        t0 = time.perf_counter()
        if self.mode == "exact":
            q = wl["points"][first::stride]
            self.oracle.exact32(wl["V"], wl["F"], q, nthreads=self.cores)
            n = len(q)
        elif wl["kind"] == "grid":
            n = len(self.ref.grid(*wl["lattice"], first=first, stride=stride, nthreads=self.cores))
        else:
            q = wl["points"][first::stride]
            n = len(self.ref.is_inside(q, nthreads=self.cores))
        return n, time.perf_counter() - t0

    def sized_stride(self, seconds):
        total = self.wl["n_total"]
        pilot_n = 2_000 if self.mode == "exact" else 200_000
        stride0 = max(1, total // pilot_n) | 1
        n, dt = self.run(0, stride0)
        rate = n / max(dt, 1e-6)
        want = int(min(total, max(n, rate * seconds)))
        stride = max(1, total // max(1, want))
        return stride | 1 if stride > 1 else 1

    def describe(self):
        kind = "float32 brute force with the reference's triangle formula (exact32)" if self.mode == "exact" else \
            "restatement of the reference algorithm (4-ary SAH BVH, float32, beta = 2, the 4 child lanes of a node evaluated 4-wide with SSE)"
        kind += ", built -march=native on this host" if "_native" in str(self.oracle.loaded_path) else ", shipped x86-64-v3 build"
        return f"{kind}, OpenMP over queries on {self.cores} threads; tree build {self.build_s:.2f} s on 1 thread"


def cpu_baseline(wl, mode, seconds):
    arm = CpuArm(wl, mode)
    stride = arm.sized_stride(seconds)
    n, dt = arm.run(0, stride)
    return {"value": n / dt / 1e9, "unit": UNIT, "cores": arm.cores, "kind": "port",
            "sample": f"{n} of {wl['n_total']} queries (every {stride}-th) in {dt:.2f} s; {arm.describe()}"}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle restatement; the reference binary is unbuildable here)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    wl = workload(args)
    arm = CpuArm(wl, args.mode)
    stride = arm.sized_stride(4.0)  # ~4 s per step
    times, n = [], 0
    for it in range(args.warmup + args.steps):
        n, dt = arm.run(it % stride, stride)
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = n / (ms * 1e-3) / 1e9
    sample = f"{n} of {wl['n_total']} queries per step (stride {stride}); {arm.describe()}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, wl["name"], wl["n_total"]),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def profile_traffic():
    """DRAM bytes (read + write) per launch of the dominant kernel, k_tile_query, from the newest committed `ncu --set full` summary
    under profiles/ (STATIC: captured with tools/final_capture.sh on the final code of the round, not in this run; one launch = the
    whole cfg2 lattice, 262144 tiles = 134.2 M queries). The tree is part of it: every launch streams the records it touches from HBM
    once; the algorithmic output is 1 byte per query."""
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for fn, tiles in (("r2_tile_plan_query_ncu_full.txt", 262144), ("r1q_tile_plan_query_ncu_full.txt", 131072)):
        path = os.path.join(ROOT, "profiles", fn)
        try:
            rd = wr = None
            in_query = False
            for ln in open(path):
                if ln.startswith("## "):
                    in_query = "k_tile_query" in ln
                elif in_query and (ln.startswith("dram__bytes_read.sum ") or ln.startswith("dram__bytes_write.sum ")):
                    f = ln.split()
                    v = float(f[-1]) * scale[f[-2]]
                    if ln.startswith("dram__bytes_read"):
                        rd = v
                    else:
                        wr = v
            if rd is not None and wr is not None:
                return {"bytes_per_launch": rd + wr, "queries_per_launch": tiles * 512, "source": f"profiles/{fn} (k_tile_query)",
                        "static": True}
        except Exception:
            continue
    return None


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch

    import lagrange_b200 as lb
    from lagrange_b200.distributed import replicate_engine, shard_range

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

    wl = workload(args)  # every rank generates the (deterministic) workload; only rank 0 builds the tree
    exact = args.mode == "exact"
    grid = wl["kind"] == "grid"
    n_total = wl["n_total"]
    eng = None
    build_info = {}
    ekw = dict(hierarchy=args.hierarchy, leaf_size=args.leaf_size)
    if rank == 0:
        # warm-up build on a small mesh: loads the build kernels (CUDA lazy module loading) so that build_ms is kernel time
        lb.FastWindingNumber(*lb.primitive.generate_subdivided_sphere("icosahedron", 4), **ekw).close()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng = lb.FastWindingNumber(wl["V"], wl["F"], **ekw)
        torch.cuda.synchronize()
        build_info = eng.info
        build_info["build_wall_ms"] = 1e3 * (time.perf_counter() - t0)
    bcast_ms = bcast_steady_ms = 0.0
    if world > 1:
        # warm the communicator first (NCCL sets up its channels lazily on the first collective: ~40-70 ms that are not the
        # cost of moving a tree), then time the replication itself
        warm = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
        dist.broadcast(warm, src=0)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        eng0 = eng
        eng = replicate_engine(eng0, src=0)
        torch.cuda.synchronize()
        dist.barrier()
        bcast_ms = 1e3 * (time.perf_counter() - t0)
        # steady state: the same replication once more (NCCL has set up its large-message channels, the allocator holds the blocks)
        t0 = time.perf_counter()
        again = replicate_engine(eng0, src=0)
        torch.cuda.synchronize()
        dist.barrier()
        bcast_steady_ms = 1e3 * (time.perf_counter() - t0)
        if rank != 0:
            again.close()
        del again

    # ---- this rank's share of the queries ---------------------------------------------------------------------------------------
    lo = 0
    if grid:
        origin, spacing, dims = wl["lattice"]
        # tile layers (8 z-planes) dealt round-robin: every rank sees every height, ONE call per rank, results compact in the rank's
        # buffer. Measured on one GPU rank by rank (tools/shard_balance.py): the 8 shares differ by 2-3 %; what is lost against
        # 1/8 of the full lattice is per-call (smaller launches: tails and the planning levels). Contiguous z-slabs: 0.81 at 8 ranks.
        # The diagonal (layer x y-part) sharding of wn_query_grid_sharded measured the same within 1 % and is not used here.
        layers = (rank, world) if world > 1 else None
        nx_, ny_, nz_ = int(dims[0]), int(dims[1]), int(dims[2])
        unit_list = [(8 * l, min(nz_, 8 * l + 8), 0, ny_) for l in range(rank, (nz_ + 7) // 8, world)]
        layout = {"units": unit_list, "n_points": sum((z1 - z0) * (y1 - y0) * nx_ for z0, z1, y0, y1 in unit_list)}
        n_local = layout["n_points"]
        h2d_bytes = 60
    else:
        lo, hi = shard_range(n_total, rank, world)
        pts_host = torch.from_numpy(wl["points"][lo:hi]).pin_memory()
        pts_dev = pts_host.cuda()
        n_local = hi - lo
        h2d_bytes = 12 * n_total
    out_dev = torch.empty(max(n_local, 1), dtype=torch.float32 if exact else torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step_device():
        if exact:
            eng.exact_solid_angle(pts_dev, out=out_dev)
        elif grid:
            eng.query_grid(origin, spacing, dims, want_inside=True, out_inside=out_dev, layers=layers)
        else:
            eng.is_inside(pts_dev, out=out_dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    # cold first call (fresh engine, fresh query set: the tiling probe and every scratch allocation are inside)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step_device()
    torch.cuda.synchronize()
    cold_ms = 1e3 * (time.perf_counter() - t0)
    for _ in range(max(0, args.warmup - 1)):
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

    # ---- e2e: the public API with HOST buffers (pinned), copies inside the timed region ---------------------------------------------
    # lattices: 60-byte descriptor in, bit-packed result out (WN_QUERY_OUT_BITS; copies of finished batches overlap the next batch);
    # point sets: 12 B per query in (H2D), bit-packed result out; exact mode: points in, float32 solid angles out
    if exact:
        out_host = torch.empty(max(n_local, 1), dtype=torch.float32).pin_memory().numpy()
        d2h_bytes = 4 * n_total
    else:
        out_host = torch.empty(max((n_local + 7) // 8, 1), dtype=torch.uint8).pin_memory().numpy()
        d2h_bytes = (n_total + 7) // 8
    pts_np = None if grid else pts_host.numpy()

    def step_host():
        if exact:
            eng.exact_solid_angle(pts_np, out=out_host)
        elif grid:
            eng.query_grid(origin, spacing, dims, want_inside=True, out_inside=out_host, layers=layers, bits=True)
        else:
            eng.is_inside(pts_np, out=out_host, bits=True)

    for _ in range(2):
        step_host()
    barrier()
    e2e_times = []
    checksum = 0
    for _ in range(max(3, min(args.steps, 5))):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step_host()
        checksum = float(out_host[::509].sum()) if exact else int(out_host[::509].sum())  # the caller reads the result
        e2e_times.append(time.perf_counter() - t0)
    t = torch.tensor([float(np.mean(e2e_times)) * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = n_total / (e2e_ms * 1e-3) / 1e9
    if exact:
        inside_local = torch.tensor([int((out_dev[:n_local] >= 6.2831854820251465).sum().item())], dtype=torch.int64, device="cuda")
    else:
        inside_local = torch.tensor([int(out_dev[:n_local].sum().item())], dtype=torch.int64, device="cuda")
    if dist is not None:
        dist.all_reduce(inside_local)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline (rank 0): flops EXECUTED by the timed kernels (their own counters), FMA peak measured live -----------------
    import ctypes

    from lagrange_b200 import _capi

    tiled = os.environ.get("WN_TILE", "1") != "0"
    nT = len(wl["F"])
    extra = {}
    if exact:
        flops_per_query = 75.0 * nT  # 75 flop per point-triangle pair, exactly
        extra = {"pairs_per_query": nT, "gpairs_per_s": n_local * nT / (ms_local * 1e-3) / 1e9}
    else:
        def stats(tiling):
            if grid:
                return eng.query_stats_grid(origin, spacing, dims, tiling=tiling)  # whole lattice: per-query averages are what is used
            return eng.query_stats(pts_dev, tiling=tiling)

        executed = stats(tiled)
        per_point = executed if not tiled else stats(False)
        nq = max(1, executed["queries"])
        # did the batch actually take the tiled path? (the probe decides per batch; the generic path executes the per-point counts)
        took_tiles = tiled and executed["node_tests"] != per_point["node_tests"]
        # SURVEY.md 8(d) units (10 / 83 / 75 flop) + the far-field interpolation of the tiled path: 64 FMA + ~60 for the weights
        interp_flops = (2 * 64 + 60) if took_tiles else 0
        flops_per_query = executed["algorithmic_flops"] / nq + interp_flops
        extra = {"tests_per_query": executed["node_tests"] / nq, "evals_per_query": executed["far_field_evals"] / nq,
                 "exact_tris_per_query": executed["exact_triangles"] / nq, "lane_utilisation": executed["node_tests"] / max(1, executed["lane_slots"]),
                 "path": "tiled (k_tile_plan + k_tile_query)" if took_tiles else "generic (k_query)",
                 "reference_algorithm": {"flops_per_query": per_point["algorithmic_flops"] / nq, "tests_per_query": per_point["node_tests"] / nq,
                                         "evals_per_query": per_point["far_field_evals"] / nq, "exact_tris_per_query": per_point["exact_triangles"] / nq,
                                         "equivalent_tflops": per_point["algorithmic_flops"] / nq * n_local / (ms_local * 1e-3) / 1e12,
                                         "note": "SURVEY 8(d) counts T, A, E of the reference algorithm on the same tree (what the CPU executes per "
                                                 "query); `achieved` / `frac` above use the smaller number of operations this engine executes"}}
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
    ncu_traffic = profile_traffic() if (args.config == 2 and not exact) else None
    # 12 B in per explicit point (0 for an implicit lattice) + 1 B (is_inside) or 4 B (solid angle) out per query
    algo_bytes = n_local * ((0 if grid else 12) + (4 if exact else 1))
    roofline = {
        "bound": "fp32_fma", "achieved": achieved_tflops, "peak": fma_peak, "unit": "TFLOP/s", "frac": achieved_tflops / fma_peak,
        "peak_source": "live FMA microbenchmark k_fma_peak (MEASURED_PEAKS.json has no FP32 CUDA-core figure)",
        "note": "achieved = flops executed by the timed kernels (SURVEY 8(d) units from the kernels' own counters) / event time; "
                "the per-point counts of the reference algorithm on the same tree are under reference_algorithm",
        "flops_per_query": flops_per_query, **extra,
        "traffic": None if ncu_traffic is None else ncu_traffic["bytes_per_launch"], "traffic_detail": ncu_traffic,
        "hbm": {"achieved_gbs": algo_bytes / (ms_local * 1e-3) / 1e9, "peak_gbs": hbm_peak, "peak_source": hbm_src,
                "frac": algo_bytes / (ms_local * 1e-3) / 1e9 / hbm_peak},
    }

    # ---- parity of the timed configuration against the restatement, strict band, on a bounded sample (rank 0) -----------------------
    parity = None
    if not exact and not args.no_cpu_baseline:
        import oracle

        m = min(n_local, 1 << 18)
        sel = np.linspace(0, n_local - 1, m).astype(np.int64)
        if grid:
            # lattice indices of the sampled output positions, from the rank's unit layout (z0, z1, y0, y1 per unit, x fastest)
            nx = int(dims[0])
            starts = np.cumsum([0] + [(z1 - z0) * (y1 - y0) * nx for z0, z1, y0, y1 in layout["units"]])
            u = np.searchsorted(starts, sel, side="right") - 1
            units = np.asarray(layout["units"], dtype=np.int64)
            rem = sel - starts[u]
            rows = units[u, 3] - units[u, 2]
            ix = rem % nx
            iy = units[u, 2] + (rem // nx) % rows
            iz = units[u, 0] + rem // (nx * rows)
            half = np.float32(0.5)
            P = np.stack([origin[0] + spacing[0] * (ix.astype(np.float32) + half), origin[1] + spacing[1] * (iy.astype(np.float32) + half),
                          origin[2] + spacing[2] * (iz.astype(np.float32) + half)], axis=1).astype(np.float32)
        else:
            P = wl["points"][lo:lo + n_local][sel]
        refe = oracle.RefEngine(wl["V"], wl["F"])
        w_ref = refe.solid_angle(P, nthreads=HOST_THREADS) / FOUR_PI
        ins_ref = refe.is_inside(P, nthreads=HOST_THREADS)
        ins_gpu = out_dev[:n_local].cpu().numpy()[sel]
        strict = np.abs(w_ref.astype(np.float64) - 0.5) > 1e-3
        parity = {"sample": int(m), "strict_band_mismatches": int((ins_gpu[strict] != ins_ref[strict]).sum()),
                  "mismatches_anywhere": int((ins_gpu != ins_ref).sum()), "points_inside_band": int((~strict).sum()),
                  "against": "oracle restatement of the reference algorithm (parity unpinned: no reference-held vectors exist)"}

    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_baseline(wl, args.mode, args.cpu_seconds if world == 1 else min(args.cpu_seconds, 6.0))

    # launches of OUR kernels per timed (device-output) step on this rank. Tiled lattice: per batch of <= 262144 tiles two k_plan_block
    # levels + k_tile_plan + k_tile_query (the probe that picks the path runs once, in the cold call, and is remembered per lattice);
    # tiled point set: k_tile_plan + k_tile_query per batch; generic = one k_query; point sets add the Morton sort of the queries
    # (2 + 4 passes x 5 launches); exact = k_exact + k_exact_reduce. profiles/r2_launches_tiled.csv is the ncu list of the same command.
    tiles = -(-n_local // 512)
    if exact:
        launches = 2
    elif extra.get("path", "").startswith("tiled") and grid:
        per_layer = -(-int(dims[0]) // 8) * -(-int(dims[1]) // 8)
        layers_local = len(layout["units"])
        per_launch = max(1, 262144 // per_layer)
        if world == 1 and per_launch >= 4:
            per_launch -= per_launch % 4  # whole 4x4x4 planning blocks per batch
        launches = 4 * max(1, -(-layers_local // per_launch))
    elif extra.get("path", "").startswith("tiled"):
        launches = 2 * max(1, -(-tiles // 262144)) + 22
    else:
        launches = 1 + (0 if grid else 22)
    api = ("FastWindingNumber.exact_solid_angle(host points) -> wn_exact: pinned HOST points in, float32 solid angles out" if exact else
           "FastWindingNumber.query_grid(bits=True) -> wn_query_grid[_strided](WN_QUERY_OUT_BITS): implicit lattice (60-byte descriptor in), "
           "pinned HOST output, 1 bit per query" if grid else
           "FastWindingNumber.is_inside(host points, bits=True) -> wn_is_inside(WN_QUERY_OUT_BITS): pinned HOST points in, 1 bit per query out")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, wl["name"], n_total),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms, "api": api},
        "gpu_launches": args.steps * launches,
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "parity": parity,
        "value_cold": {"ms_first_call": cold_ms, "value": n_local / (cold_ms * 1e-3) / 1e9,
                       "note": "rank 0's first call on a fresh engine: tiling probe, scratch allocation and lazy kernel loading included"},
        "build": {k: build_info.get(k) for k in ("build_ms", "build_ms_morton", "build_ms_sort", "build_ms_hierarchy", "build_ms_moments",
                                                 "build_ms_pack", "build_wall_ms", "num_entries", "tree_bytes", "max_depth")},
        "tree_broadcast_ms": bcast_ms, "tree_broadcast_ms_steady": bcast_steady_ms, "inside_count": int(inside_local.item()), "checksum": checksum,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
