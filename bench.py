#!/usr/bin/env python3
"""Benchmark of the PS-BAX sample-path hot path (BASELINE.json metric:
GP sample-path evals/sec = candidates x samples / time(evaluate + reduce)).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): level-set estimation on a dense 2-D grid,
4096 x 4096 = 16,777,216 candidates per GPU, S = 64 Matheron sample paths, L = 1000 RFF
features, n = 64 training points (standardized Himmelblau), Matern-5/2 ARD,
threshold = 55th percentile.  One "step" = one pass of the fused evaluate + level-set
reduction over the rank's candidate shard (weak scaling: every GPU holds its own 4096 x 4096
slab of a (gpus*4096) x 4096 grid; no data-path collective for the level set).
Inputs are larger than L2 (134 MB of X* + 134 MB of mask per pass), so no L2 flush is needed.

Printed JSON keys follow the driver contract; see DESIGN.md section "Measurement".
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

GRID = 4096
S, L, NTRAIN, D = 64, 1000, 64, 2
KERNEL = "matern52"
NOISE = 1e-4
TAU_Q = 0.55
METRIC = "gp_sample_path_evals_per_sec"
UNIT = "candidate*sample evals/s"


# ------------------------------------------------------------------------------------------
# synthetic config-2 problem (same on every rank: seed 0 for the model, grid is analytic)
# ------------------------------------------------------------------------------------------
def himmelblau(X):
    Z = 12.0 * X - 6.0
    a, b = Z[:, 0], Z[:, 1]
    return -((a ** 2 + b - 11) ** 2 + (a + b ** 2 - 7) ** 2)


def problem():
    gen = torch.Generator().manual_seed(0)
    Xt = torch.rand(NTRAIN, D, dtype=torch.float64, generator=gen)
    y = himmelblau(Xt)
    ls = 0.1 + 0.4 * torch.rand(D, dtype=torch.float64, generator=gen)
    g2 = torch.Generator().manual_seed(1234)
    ft = himmelblau(torch.rand(10000, D, dtype=torch.float64, generator=g2))
    tau = float(torch.sort(ft).values[int(TAU_Q * 10000)])  # demos/level-set.py:42-48
    return Xt, y, ls, tau


def grid_slab(rank, world, device):
    """Rows [rank*GRID^2, (rank+1)*GRID^2) of reshape_mesh(get_mesh(2, .)) for a
    (world*GRID) x GRID grid ('ij' order: first coordinate slowest), float32 on `device`."""
    a0 = torch.linspace(0, 1, world * GRID, dtype=torch.float64)[rank * GRID:(rank + 1) * GRID]
    a1 = torch.linspace(0, 1, GRID, dtype=torch.float64)
    X = torch.stack(torch.meshgrid(a0, a1, indexing="ij"), dim=-1).reshape(-1, 2)
    return X.to(torch.float32).to(device)


# ------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# CPU reference path: the oracle's restatement of the reference flow (fp64, one fresh basis
# per sample, eval_in_batch chunks of 100, samples in series) on the host cores
# ------------------------------------------------------------------------------------------
def cpu_reference_pass(n_rows, n_samples, seed=0):
    """One bounded sample of the workload through the reference-style CPU path.
    Returns (evals, seconds) of evaluate + reduce; parameter draws are not timed."""
    from oracle import psbax_oracle as O

    Xt, y, ls, tau = problem()
    gp = O.GPState(train_X=Xt, train_Y=y, lengthscale=ls, outputscale=1.0, noise=NOISE, kernel=KERNEL)
    gen = torch.Generator().manual_seed(seed)
    p = O.draw_reference_paths(gp, n_samples, L=L, generator=gen)
    steps = int(math.ceil(math.sqrt(n_rows)))
    X = O.reshape_mesh(O.get_mesh(2, steps))[:n_rows]
    t0 = time.perf_counter()
    for s in range(n_samples):
        f = O.make_callable(p, s)
        fx = O.eval_in_batch(f, X.unsqueeze(1)).detach().numpy().flatten()  # levelset.py:26-29
        O.levelset_reduce(fx, tau)
    return n_rows * n_samples, time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_rows, n_samples = 200_000, 4
    for _ in range(args.warmup):
        cpu_reference_pass(5_000, 1)
    evals = secs = 0.0
    for i in range(args.steps):
        e, t = cpu_reference_pass(n_rows, n_samples, seed=i)
        evals += e
        secs += t
    value = evals / secs
    sample = (f"{n_rows} grid rows x {n_samples} paths per step, fp64, fresh RFF basis per path, "
              f"chunks of 100 rows, {torch.get_num_threads()} torch threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(gpus):
    return {"workload": "level-set on a dense 2-D grid (BASELINE configs[1])",
            "candidates_per_gpu": GRID * GRID, "grid": f"{GRID}x{GRID} per GPU", "d": D, "samples": S,
            "rff_features": L, "train_points": NTRAIN, "kernel": KERNEL, "threshold_quantile": TAU_Q,
            "sampler": "matheron (shared basis + exact-kernel update)",
            "sharding": f"candidate rows over {gpus} GPU(s), no collective",
            "l2": "inputs larger than L2 (134 MB X* + 134 MB mask per pass); no flush"}


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    from psbax_b200.models import SingleTaskGP
    from psbax_b200.sample_paths import draw_matheron_paths

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    N = GRID * GRID
    K, Wm = args.steps, max(args.warmup, 3)

    Xt, y, ls, tau = problem()
    model = SingleTaskGP(Xt, y, kernel=KERNEL, lengthscale=ls, outputscale=1.0, noise=NOISE)
    gen = torch.Generator(device=dev).manual_seed(0)  # identical parameters on every rank
    t_draw0 = time.perf_counter()
    paths = draw_matheron_paths(model, S, L, generator=gen, device=dev, engine=args.engine)
    torch.cuda.synchronize()
    t_draw = time.perf_counter() - t_draw0
    X = grid_slab(rank, world, dev)
    row_offset = rank * N

    def step():
        return paths.levelset(X, tau, want_mask=True, row_offset=row_offset)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(Wm):
        res = step()
    sync_all()
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    # ---- device-resident timing: K steps, per-step CUDA events on the launching stream ----
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sync_all()
    w0 = time.time()
    t0 = time.perf_counter()
    for a, b in evs:
        a.record()
        res = step()
        b.record()
    sync_all()
    t_wall = time.perf_counter() - t0
    w1 = time.time()
    kern_ms = [a.elapsed_time(b) for a, b in evs]
    span_ms = evs[0][0].elapsed_time(evs[-1][1])
    # ---- end to end: host buffers through the public streaming API.  Every step uploads the
    # rank's X* from pinned host memory and downloads the mask + counts; uploads, kernels and
    # downloads of successive 2M-row chunks overlap on three streams. ----
    Xh = X.cpu().pin_memory()
    words = (N + 31) // 32

    def e2e_step():
        return paths.levelset_streamed(Xh, tau, chunk_rows=int(os.environ.get("PSBAX_E2E_CHUNK", 1 << 21)), row_offset=row_offset)

    e2e_step()
    sync_all()
    e0 = time.perf_counter()
    for _ in range(K):
        masks_h, _b, cnt_h, _bv, _bi = e2e_step()
    sync_all()
    t_e2e = time.perf_counter() - e0
    w1 = time.time()  # clocks are sampled over both timed loops (device-resident and end-to-end)
    assert int(cnt_h[0]) == int(res.count[0]), "streamed and resident level sets disagree"
    clk = clocks.stop(w0, w1) if rank == 0 else None

    # max over ranks
    tt = torch.tensor([span_ms / 1e3, t_wall, t_e2e, sum(kern_ms) / 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev, t_wall_max, t_e2e_max, t_kern = tt.tolist()
    counts = res.count.cpu().tolist()

    if rank == 0:
        evals_step = N * S * world
        value = evals_step * K / t_dev
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        eng = paths_engine(paths)
        bf16 = peaks.get("bf16_tflops", 1590.0)
        which = "measured (MEASURED_PEAKS.json bf16_tflops burst)" if peaks else "fallback 1590 TF bf16"
        # three MMAs per algorithmic FMA.  BF16x3: peak = bf16 / 3;  3xTF32: TF32 dense = bf16 / 2 -> bf16 / 6
        peak_tf = bf16 / (3.0 if eng == "tensor_bf16x3" else 6.0)
        flops_launch = 2.0 * (L + NTRAIN) * S * N  # algorithmic, per launch (per GPU)
        kern_s = (t_kern / K)
        achieved = flops_launch / kern_s / 1e12
        cpu = None
        if not args.no_cpu_baseline and world == 1:  # reported on rank 0 at N = 1 only
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            cpu_reference_pass(5_000, 1)
            ev = se = 0.0
            n_rows, n_samples = 200_000, 4
            while se < 12.0:
                e, t = cpu_reference_pass(n_rows, n_samples, seed=int(ev) % 97)
                ev += e
                se += t
            cpu = {"value": ev / se, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": (f"{int(ev)} evals in {se:.1f} s: passes of {n_rows} grid rows x {n_samples} paths, "
                              f"oracle restatement of the reference flow (fp64, fresh RFF basis per path, "
                              f"chunks of 100 rows, samples in series), {torch.get_num_threads()} torch threads")}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": 1e3 * t_dev / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"simt": "f32", "tensor_tf32x3": "f32 (3xTF32 split, fp32 accumulate)",
                                           "tensor_bf16x3": "f32 (BF16x3 split, fp32 accumulate)"}[eng],
            "data": "synthetic", "config": dict(workload_config(world), engine=eng,
                                                **({"numa_node_rank0": numa} if numa is not None else {})),
            "e2e": {"value": evals_step * K / t_e2e_max, "unit": UNIT,
                    "h2d_bytes_per_step": N * D * 4, "d2h_bytes_per_step": S * words * 4 + S * 16,
                    "ms_per_step": 1e3 * t_e2e_max / K},
            "gpu_launches": 2 * K,  # per step: fused evaluate+level-set kernel + finalize (device-resident loop)
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf,
                         # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this workload, one
                         # ncu --set full capture (profiles/r01_ncu_bench_kernel_bf16x3_details.csv);
                         # algorithmic bytes = 134.2 MB X* + 134.2 MB mask
                         "traffic": 224.2e6 if eng == "tensor_bf16x3" else None,
                         "kernel": "fused evaluate+level-set kernel (+3 us finalize)",
                         "kernel_ms": 1e3 * kern_s,
                         "note": (f"algorithmic flops = 2*(L+n)*S per candidate = {2 * (L + NTRAIN) * S}; "
                                  f"peak = bf16 {bf16} TF/s / {3 if eng == 'tensor_bf16x3' else 6} "
                                  f"(three split-product MMAs per algorithmic FMA; TF32 runs at bf16/2), {which}")},
            "cpu_baseline": cpu,
            "clocks": clk,
            "wall_ms_per_step": 1e3 * t_wall_max / K,
            "param_draw_ms": 1e3 * t_draw,
            "check": {"count_s0": counts[0], "frac_above_s0": counts[0] / N},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def paths_engine(paths):
    """Which kernel family the library resolves to for this batch (mirrors make_layout)."""
    shared = paths.G == 1 and paths.S >= 16 and paths.L > 0
    if paths.engine == "simt" or not shared:
        return "simt"
    if paths.engine == "tensor":
        return "tensor_tf32x3"
    return "tensor_bf16x3"


def bind_to_gpu_numa_node(index: int):
    """Multi-GPU runs: keep this rank's threads (and therefore its pinned staging buffers, placed by
    first touch) on the NUMA node its GPU hangs off, so that eight ranks streaming 2 x 134 MB per step
    do not all cross the socket interconnect.  Best effort: returns the node or None."""
    try:
        import pynvml

        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:  # nvml prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default="auto", choices=["auto", "simt", "tensor", "tensor_bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
