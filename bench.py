#!/usr/bin/env python
"""
bench.py -- free-energy + gradient evaluations per second (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one ``E_cost(x)`` with gradient (free_energy.py:651-751) on the
north-star configuration: Lorenz '96, D=500, d=125 partial observations,
M=16 observation times (examples/example_500D.ipynb), synthetic state
(SURVEY.md 8d).  Rank 0 prints ONE JSON line:

  value      evals/s with x resident in HBM (CUDA events, L2 flushed between steps)
  e2e        evals/s through FreeEnergy.E_cost on HOST numpy buffers (pinned x;
             H2D of x and D2H of F+gradient inside the timed region)
  roofline   the dominant kernel (esde_l96_kernel) against the FP64 pipe
             (DFMA micro-benchmark measured in this run) and HBM
  cpu_baseline  the numpy/scipy oracle port on the host cores (bounded sample)
  c_abi_back_to_back  mfvgp_ecost on device pointers in a loop, no flush (what SCG sees)
  scg        device SCG iterations per second on the same configuration
  batched    evals/s when B independent states of the same problem are evaluated per step
             (mfvgp_ecost_batch; what check_gradients=True uses) -- the size at which the
             north-star problem fills the GPU
  large      Lorenz '96 D=8192, M=4097 (L=4096 intervals) long horizon: at
             N ranks the intervals are sharded (strong scaling); this is the
             configuration on which the roofline fraction is meaningful.
  next_posterior / next_trajectory  the rows of SURVEY.md 8(f) built so far:
             posterior reconstruction (HBM roofline) and the bit-identical
             Lorenz 96 sample path (latency bound), each with its CPU port timed.

With N > 1 (torchrun, one rank per GPU) the D=500 evaluation is far below one
wave of one GPU, so it does not shard: ``value`` is N independent replicas
("replicas only"); the sharded multi-GPU path is the ``large`` entry.

``--impl reference`` times the CPU implementation of the same path (the oracle
port: the reference itself is Python under /root/reference and cannot travel
to the GPU box) with all host threads it can use.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

FLOP_PER_CELL = 4620.0      # 110 FLOP/(dim,node) x 42 nodes
BYTE_PER_CELL = 96.0        # 40 B read + 56 B written per cell
# What the kernels actually execute and move per cell, from the committed ncu captures
# (profiles/r1_kernel_counts.json: FP64 thread instructions, DFMA share, DRAM bytes):
COUNTS = json.loads((Path(__file__).resolve().parent / "profiles" / "r1_kernel_counts.json")
                    .read_text())
METRIC = "free-energy+gradient evals/sec (L96 D=500)"
WORKLOAD = "L96 D=500 d=125 M=16 (example_500D.ipynb), synthetic state seed 8192"


def north_star_problem():
    from helpers import synthetic_l96
    return synthetic_l96(500, 16, seed=8192)


class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [v.strip() for v in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                smax.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)),
                   "reasons": sorted(reasons), "samples": len(sm)}
        return out


def time_steps(torch, fn, steps, warmup, flush_buf):
    """CUDA-event time of ``steps`` calls of fn, L2 flushed before each."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(steps)]
    for e0, e1 in ev:
        if flush_buf is not None:
            flush_buf.zero_()
        e0.record()
        fn()
        e1.record()
    torch.cuda.synchronize()
    return sum(e0.elapsed_time(e1) for e0, e1 in ev)        # ms


def kernel_time(fe, lib, fn, reps):
    """Average duration of the Esde kernel (events recorded inside the library)."""
    tot, cnt = ctypes.c_double(), ctypes.c_int64()
    lib.mfvgp_timing(fe._h, 1, ctypes.byref(tot), ctypes.byref(cnt))
    for _ in range(reps):
        fn()
    lib.mfvgp_timing(fe._h, 0, ctypes.byref(tot), ctypes.byref(cnt))
    return tot.value / max(cnt.value, 1), cnt.value


def roofline(D, L, kernel_ms, fp64_peak_tf, hbm_peak_gbs, kernel):
    """
    achieved = ALGORITHMIC FLOPs of SURVEY 8(d) (4 620 per (dimension, interval) cell,
    the 42-node formulation) / measured kernel time; peak = DFMA rate measured in this
    run.  The kernels use the split rule, which needs fewer FLOPs for the same result
    (DESIGN.md 4), so `executed` states what the pipe really did (ncu counts).
    """
    cells = float(D) * L
    sec = kernel_ms * 1e-3
    tf = FLOP_PER_CELL * cells / sec * 1e-12
    gbs = BYTE_PER_CELL * cells / sec * 1e-9
    c = COUNTS.get(kernel, {})
    out = {"bound": "fp64", "achieved": tf, "peak": fp64_peak_tf, "unit": "TFLOP/s",
           "frac": tf / fp64_peak_tf,
           "traffic": c["dram_bytes_per_cell"] * cells if "dram_bytes_per_cell" in c else None,
           "peak_source": "DFMA micro-benchmark measured in this run (mfvgp_fp64_peak); "
                          "MEASURED_PEAKS.json has no FP64 entry",
           "kernel": kernel, "kernel_ms": kernel_ms,
           "algorithmic_flop_per_launch": FLOP_PER_CELL * cells,
           "algorithmic_bytes_per_launch": BYTE_PER_CELL * cells,
           "hbm": {"achieved": gbs, "peak": hbm_peak_gbs, "unit": "GB/s",
                   "frac": gbs / hbm_peak_gbs, "peak_source": "MEASURED_PEAKS.json hbm_gbs"}}
    if "fp64_inst_per_cell" in c:
        inst = c["fp64_inst_per_cell"] * cells
        flop = (c["fp64_inst_per_cell"] + c["dfma_per_cell"]) * cells       # FMA = 2
        out["executed"] = {
            "source": c.get("source"),
            "fp64_inst_per_cell": c["fp64_inst_per_cell"],
            "tflops": flop / sec * 1e-12,
            # executed FP64 instructions / (time x measured DFMA issue rate): utilisation of
            # the pipe that bounds the kernel (an FMA, a multiply and an add each take one slot)
            "fp64_pipe_frac": inst / sec / (fp64_peak_tf * 0.5e12)}
    return out


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.is_file():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def cpu_baseline(g, budget_s=12.0, n_jobs=1):
    """Oracle port (numpy + scipy.quad_vec, adaptive like the reference)."""
    from helpers import make_oracle
    o = make_oracle(g, adaptive=True)
    o.n_jobs = n_jobs
    o.E_cost(g["x0"])                     # warm-up (worker spawn)
    t0 = time.perf_counter()
    n = 0
    while True:
        o.E_cost(g["x0"])
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 50:
            break
    wall = time.perf_counter() - t0
    return {"value": n / wall, "unit": "evals/s", "cores": n_jobs, "kind": "port",
            "sample": f"{n} E_cost evaluations of the same workload "
                      f"(oracle/mfvgp_oracle.py, scipy.quad_vec adaptive GK21, "
                      f"n_jobs={n_jobs}) in {wall:.1f} s"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    g = north_star_problem()
    cores = os.cpu_count() or 1
    from helpers import make_oracle
    o = make_oracle(g, adaptive=True)
    o.n_jobs = cores
    for _ in range(max(args.warmup, 1)):
        o.E_cost(g["x0"])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.E_cost(g["x0"])
    wall = time.perf_counter() - t0
    val = args.steps / wall
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": val, "unit": "evals/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} E_cost evaluations per run, joblib over "
                                       f"the 15 intervals, n_jobs={cores}"},
            "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "note": "oracle port of the reference's CPU path (numpy + scipy.quad_vec); the "
                    "reference itself (Python under /root/reference) measured 5.59 s/eval "
                    "with n_jobs=8 on the build container (BASELINE.md)"}
    print(json.dumps(line), flush=True)


def run_large(torch, dist, rank, world, steps, warmup, fp64_peak_tf, hbm_gbs, D=8192, M=4097):
    """L96 D=8192 long horizon; intervals sharded over the ranks."""
    from meanfieldvargp_b200 import _lib
    from meanfieldvargp_b200.parallel import ShardedFreeEnergy
    from helpers import synthetic_l96_device
    g = synthetic_l96_device(torch, D, M, seed=8192)
    fe = ShardedFreeEnergy(g, rank, world)
    x = g["x0"]
    L = M - 1

    def step():
        fe.E_cost(x)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()          # working set (1.3 GB x + 1.3 GB grad + 1.9 GB workspace) >> L2
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = ms.item() / steps
    lib = _lib.load()
    k_ms, _ = kernel_time(fe.local, lib, step, 3)
    Lloc = fe.local.intervals[1] - fe.local.intervals[0]
    desc = fe.local.describe()
    out = {"workload": f"L96 D={D} d={D // 4} M={M} (L={L} intervals), synthetic seed 8192",
           "value": 1e3 / ms_step, "unit": "evals/s", "ms_per_eval": ms_step,
           "scaling": "strong", "parallelism": f"time-intervals sharded x{world}",
           "l2": "inputs larger than L2",
           "path": desc,
           "roofline": roofline(D, Lloc, k_ms, fp64_peak_tf, hbm_gbs, desc["kernel"]),
           "job_tflops_algorithmic": FLOP_PER_CELL * D * L / (ms_step * 1e-3) * 1e-12}
    del fe, g, x
    torch.cuda.empty_cache()
    return out


def run_batched(torch, fe, g, fp64_peak_tf, hbm_gbs, steps=20, warmup=3):
    """
    The north-star problem evaluated for B independent states per step (mfvgp_ecost_batch: two
    launches per batch): what fills 148 SMs at this size, where ONE evaluation is 51 cells per
    SM.  The use in the reference's API is find_minimum(check_gradients=True) -- len(x) + 1
    evaluations -- and multi-start runs.  CUDA events around `steps` batches; X, the gradients
    and the per-entry workspaces of B >= 256 are larger than L2.
    """
    n = g["x0"].size
    D, L = 500, 15
    out = {"workload": "B independent states of L96 D=500 d=125 M=16 per step (x0 + 0.05 N(0,1))",
           "unit": "evals/s", "sizes": {}}
    rng = np.random.default_rng(11)
    best = None
    for B in (16, 64, 256, 1024, 4096):
        X = torch.as_tensor(g["x0"][None, :] + 0.05 * rng.standard_normal((B, n)), device="cuda")
        for _ in range(warmup):
            fe.E_cost_batch(X, chunk=B)
        torch.cuda.synchronize()
        l0 = fe.kernel_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            F, G = fe.E_cost_batch(X, chunk=B)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        # mfvgp_ecost_batch's rule: the one-thread-per-cell split kernel once the batch alone
        # fills the GPU, the cooperative kernel of the single evaluation below that
        kern = "esde_l96_split_kernel" if B * D * L >= 148 * 2048 else "esde_l96_coop_kernel"
        roof = roofline(D, L * B, ms, fp64_peak_tf, hbm_gbs, kern)
        rec = {"value": B * 1e3 / ms, "ms_per_batch": ms, "main_kernel": kern,
               "launches_per_batch": (fe.kernel_launches - l0) / steps,
               "roofline_frac_whole_step": roof["frac"],
               "fp64_pipe_frac_whole_step": roof.get("executed", {}).get("fp64_pipe_frac"),
               "l2": "larger than L2" if B >= 256 else "fits L2"}
        if B >= 4096:
            rec["note"] = "1.4 GB of states, 1.4 GB of gradients per step"
        out["sizes"][str(B)] = rec
        if best is None or rec["value"] > best[1]["value"]:
            best = (B, rec)
        del X, F, G
    # the call in the reference's API that needs a batch: scipy.optimize.check_grad behind
    # find_minimum(check_gradients=True) -- len(x) + 1 evaluations (free_energy.py:805)
    fe.check_grad_batched(g["x0"])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    err = fe.check_grad_batched(g["x0"])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["check_grad"] = {"evaluations": int(n + 1), "wall_ms": dt * 1e3, "evals_per_s": (n + 1) / dt,
                         "error_norm": err,
                         "note": "FreeEnergy.check_grad_batched(x0): forward differences of all "
                                 "len(x) coordinates, F only, 512 states per launch pair"}
    out["value"] = best[1]["value"]
    out["B"] = best[0]
    out["note"] = ("whole step (main + tail kernel, status read-back, host wrapper) against the FP64 "
                   "roofline with the per-cell counts of the single-state kernel; rows are "
                   "bit-identical to single E_cost calls (tests/test_gpu_batch.py)")
    return out


def run_posterior(torch, hbm_gbs):
    """posterior reconstruction on the fine grid (csrc/posterior.cuh)."""
    import meanfieldvargp_b200 as mf
    out = {}
    rng = np.random.default_rng(7)
    for tag, D, M, num in (("L96-500", 500, 16, 100), ("L96-8192-long", 8192, 4097, 20)):
        obs_times = 0.25 * np.arange(1, M + 1)
        mp = torch.randn((D, 3 * M + 4), dtype=torch.float64, device="cuda")
        ls = torch.log(4.0 + 2.0 * torch.rand((D, 2 * M + 3), dtype=torch.float64, device="cuda"))
        # the C-ABI entry on preallocated device buffers (what is timed); the Python wrapper
        # mf.reconstruct adds the output allocations and the upload of Tx
        from meanfieldvargp_b200 import _lib
        lib = _lib.load()
        P = (M + 1) * num + 1
        Tx = torch.tensor(mf.time_knots(0.0, obs_times, obs_times[-1] + 0.25), device="cuda")
        tt = torch.empty(P, dtype=torch.float64, device="cuda")
        mt = torch.empty((D, P), dtype=torch.float64, device="cuda")
        st = torch.empty((D, P), dtype=torch.float64, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream
        cargs = (D, M, num, Tx.data_ptr(), mp.data_ptr(), mp.shape[1], ls.data_ptr(), ls.shape[1],
                 1, tt.data_ptr(), mt.data_ptr(), st.data_ptr(), stream)
        for _ in range(3):
            _lib.check(lib.mfvgp_reconstruct(*cargs), "mfvgp_reconstruct")
        torch.cuda.synchronize()
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            lib.mfvgp_reconstruct(*cargs)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        P = (M + 1) * num + 1
        nbytes = 8.0 * (2.0 * D * P + P + D * (5 * M + 7))          # written + read once
        out[tag] = {"D": D, "M": M, "num": num, "samples": P, "ms": ms,
                    "roofline": {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) * 1e-9,
                                 "peak": hbm_gbs, "unit": "GB/s",
                                 "frac": nbytes / (ms * 1e-3) * 1e-9 / hbm_gbs,
                                 "algorithmic_bytes_per_launch": nbytes}}
        if tag == "L96-500":
            from oracle import posterior_oracle as O
            mph, sph = mp.cpu().numpy(), np.exp(ls.cpu().numpy())
            t0 = time.perf_counter()
            O.reconstruct(mph, sph, 0.0, obs_times, obs_times[-1] + 0.25, num)
            out[tag]["cpu_baseline"] = {"value": time.perf_counter() - t0, "unit": "s", "cores": 1,
                                        "kind": "port",
                                        "sample": "one call of oracle/posterior_oracle.py (numpy, "
                                                  "vectorised over D and t; the reference's own "
                                                  "triple Python loop is ~1.7 M lambda calls)"}
        del tt, mt, st, mp, ls
        torch.cuda.empty_cache()
    return out


def run_trajectory(torch):
    """Lorenz 96 sample path on the device (csrc/trajectory.cuh)."""
    import meanfieldvargp_b200 as mf
    out = {}
    for tag, D, tf, dt in (("L96-500 (example_500D.ipynb: T=[0,4], dt=1e-3)", 500, 4.0, 1.0e-3),
                           ("L96-8192", 8192, 1.0, 1.0e-3)):
        rng = np.random.default_rng(575)
        dim_t = mf.trajectory.time_window(0.0, tf, dt).size
        noise = torch.as_tensor(rng.standard_normal((D, dim_t)), device="cuda")
        mf.simulate_l96(40.0, 8.0, D, 0.0, tf, dt, noise=noise)            # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tk, x = mf.simulate_l96(40.0, 8.0, D, 0.0, tf, dt, noise=noise)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        steps = 125 * D + dim_t - 1
        out[tag] = {"D": D, "burn_in_steps": 125 * D, "path_steps": dim_t - 1, "wall_s": wall,
                    "steps_per_s": steps / wall, "us_per_step": 1e6 * wall / steps,
                    "bound": "latency (time is sequential: one CTA, one barrier per step)"}
        if D == 500:
            from oracle import trajectory_oracle as O
            t0 = time.perf_counter()
            _, xo = O.make_trajectory(40.0, 8.0, D, 0.0, tf, dt, noise=noise.cpu().numpy())
            cw = time.perf_counter() - t0
            out[tag]["cpu_baseline"] = {"value": steps / cw, "unit": "steps/s", "cores": 1,
                                        "kind": "port", "sample": f"the same path, {cw:.2f} s "
                                        "(oracle/trajectory_oracle.py = the reference's numpy loop)"}
            out[tag]["bit_identical_to_cpu"] = bool(np.array_equal(x.cpu().numpy(), xo))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-large", action="store_true", help="skip the D=8192 entry")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--large-steps", type=int, default=10)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import meanfieldvargp_b200 as mf
    from meanfieldvargp_b200 import _lib
    from helpers import make_free_energy
    lib = _lib.require_device()

    tf = ctypes.c_double()
    _lib.check(lib.mfvgp_fp64_peak(local_rank, ctypes.byref(tf)), "mfvgp_fp64_peak")
    fp64_peak_tf = tf.value
    hbm_gbs, hbm_src = hbm_peak()

    g = north_star_problem()
    D, M = 500, 16
    L = M - 1
    fe = make_free_energy(g, device=local_rank)
    n = g["x0"].size
    x_dev = torch.tensor(g["x0"], dtype=torch.float64, device="cuda")
    x_pin = torch.tensor(g["x0"], dtype=torch.float64).pin_memory()
    x_host = x_pin.numpy()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")

    # ---- value: device-resident evaluations ---------------------------------
    def step_dev():
        fe.E_cost(x_dev)

    clocks = ClockSampler(local_rank)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = fe.kernel_launches
    for _ in range(args.warmup):
        step_dev()
    launches_w = fe.kernel_launches
    clocks.start()
    ms_total = time_steps(torch, step_dev, args.steps, 0, flush)
    launches = fe.kernel_launches - launches_w
    tms = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = tms.item() / args.steps
    value = world * 1e3 / ms_step          # N independent replicas

    # ---- extra: the bare C-ABI call, back to back, no flush (what SCG sees) -------------
    out_d = torch.empty(8, dtype=torch.float64, device="cuda")
    grad_d = torch.empty(n, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    cargs = (fe._h, x_dev.data_ptr(), 1, out_d.data_ptr(), grad_d.data_ptr(),
             out_d.data_ptr() + 32, stream)
    for _ in range(args.warmup):
        lib.mfvgp_ecost(*cargs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        lib.mfvgp_ecost(*cargs)
    e1.record()
    torch.cuda.synchronize()
    c_abi = {"value": world * args.steps / (e0.elapsed_time(e1) * 1e-3), "unit": "evals/s",
             "note": "mfvgp_ecost on device pointers, back to back, x in L2 (no flush)"}

    # ---- e2e: host buffers through the public API ----------------------------
    for _ in range(args.warmup):
        fe.E_cost(x_host)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # K steps are ~10 ms of wall clock at this size, which a single scheduler hiccup on the
    # host distorts: nine blocks of K steps, the median block is reported (spread alongside)
    walls = []
    for _ in range(9):
        t0 = time.perf_counter()
        for _ in range(args.steps):
            F_h, grad_h = fe.E_cost(x_host)
        walls.append(time.perf_counter() - t0)
    walls.sort()
    wall = walls[len(walls) // 2]
    twall = torch.tensor([wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(twall, op=dist.ReduceOp.MAX)
    clock_info = clocks.stop()          # sampled over the device-timed and the e2e regions
    e2e = {"value": world * args.steps / twall.item(), "unit": "evals/s",
           "h2d_bytes_per_step": int(n * 8), "d2h_bytes_per_step": int(n * 8 + 40),
           "api": "FreeEnergy.E_cost(numpy x in pinned memory) -> (float, numpy grad); wall "
                  "clock; one H2D and one D2H DMA per call, stream synchronised inside",
           "blocks": {"n": len(walls), "steps_each": args.steps, "reported": "median",
                      "fastest": world * args.steps / walls[0],
                      "slowest": world * args.steps / walls[-1]}}

    # ---- roofline of the dominant kernel ------------------------------------
    k_ms, k_cnt = kernel_time(fe, lib, step_dev, 50)
    desc = fe.describe()
    roof = roofline(D, L, k_ms, fp64_peak_tf, hbm_gbs, desc["kernel"])
    roof["hbm"]["peak_source"] = f"MEASURED_PEAKS.json hbm_gbs ({hbm_src})"
    roof["note"] = ("7 500 cells = 51 per SM: this launch is latency-bound (DESIGN.md 5); the "
                    "roofline fraction that characterises the kernels is large.roofline")

    # ---- SCG iterations per second -------------------------------------------
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        fe.find_minimum(x_dev, max_iter=10)
        torch.cuda.synchronize()
        scg_walls = []
        for _ in range(5):                  # a run is ~2.5 ms of wall clock: median of five
            t0 = time.perf_counter()
            _, fx, stats = fe.find_minimum(x_dev, max_iter=50)
            torch.cuda.synchronize()
            scg_walls.append(time.perf_counter() - t0)
        scg_walls.sort()
        scg_wall = scg_walls[2]
    scg = {"iters_per_s": stats["nit"] / scg_wall, "nit": int(stats["nit"]),
           "func_eval": int(stats["func_eval"]), "fx_final": float(fx),
           "fastest": stats["nit"] / scg_walls[0], "slowest": stats["nit"] / scg_walls[-1],
           "note": "mfvgp_scg on device-resident x, whole find_minimum(max_iter=50) calls incl. "
                   "their start-up iterations, wall clock, median of five runs"}

    large = None
    if not args.no_large:
        try:
            large = run_large(torch, dist, rank, world, args.large_steps, 3, fp64_peak_tf,
                              hbm_gbs)
        except Exception as ex:                     # reported, never hidden
            large = {"error": f"{type(ex).__name__}: {ex}"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(g, n_jobs=1)
    batched = None
    if rank == 0 and not args.no_large:
        try:
            batched = run_batched(torch, fe, g, fp64_peak_tf, hbm_gbs)
        except Exception as ex:
            batched = {"error": f"{type(ex).__name__}: {ex}"}
    posterior = None
    if rank == 0 and not args.no_large:
        try:
            posterior = run_posterior(torch, hbm_gbs)
        except Exception as ex:
            posterior = {"error": f"{type(ex).__name__}: {ex}"}
    trajectory = None
    if rank == 0 and not args.no_large:
        try:
            trajectory = run_trajectory(torch)
        except Exception as ex:
            trajectory = {"error": f"{type(ex).__name__}: {ex}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "l2": "flushed (256 MiB write) between steps",
                           "path": desc,
                           "parallelism": "single GPU" if world == 1 else
                           f"replicas only x{world} (evaluation is < 1 wave of one GPU)"},
                "clocks": clock_info, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roof, "cpu_baseline": cpu, "c_abi_back_to_back": c_abi,
                "scg": scg, "batched": batched, "large": large, "next_posterior": posterior,
                "next_trajectory": trajectory,
                "fp64_peak_tflops_measured": fp64_peak_tf}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
