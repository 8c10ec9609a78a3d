#!/usr/bin/env python
"""bench.py -- MC steps/s of the mcdiff Monte-Carlo likelihood loop on B200.

Metric (BASELINE.json): MC steps/sec summed over all chains at 100 bins x 4 lag times,
plus the fraction of the FP64 roofline.  One bench "step" = ONE launch of the fused chain
kernel that advances every chain of this rank by MC_PER_LAUNCH Monte-Carlo steps (one MC
step = proposal + rate matrix + 4 propagators expm(lag R) + log L + Metropolis,
i.e. one iteration of lib/mc.py:27 of the reference).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Under torchrun (N > 1) every rank owns CHAINS_PER_GPU chains (weak scaling, no data-path
collective; one NCCL all_gather of the posterior summaries at the end).
"""
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

N_BINS = 100
LAGS = (10.0, 20.0, 30.0, 40.0)
CHAINS_PER_GPU = 1024
MC_PER_LAUNCH = 200
NCOS_F, NCOS_D = 10, 6           # the reference's own example model (examples/example-HexdWat-run1.sh)
DV = DW = 0.05
NMC_UPDATE = 1000.0
NCU_DRAM_BYTES_PER_LAUNCH = 3954944 + 56320
METRIC = "MC steps/sec (all chains, 100 bins x 4 lags)"
UNIT = "MC steps/s"


def flops_per_mc_step(n, L):
    """F_step = 9 n^3 + L (2 n^3 + 3 n^2)  (17.1 MFLOP at 100 x 4)."""
    return 9.0 * n ** 3 + L * (2.0 * n ** 3 + 3.0 * n ** 2)


def workload():
    from oracle import mcdiff_oracle as O   # synthetic-input generator only
    counts, lags, edges, F, w_true = O.synthetic_counts(N_BINS, LAGS)
    vb = O.cos_basis_center(N_BINS, NCOS_F)
    wb = O.cos_basis_border(N_BINS, N_BINS, NCOS_D)
    dz = edges[1] - edges[0]
    wc0 = np.zeros(NCOS_D)
    wc0[0] = np.log(1.0) - np.log(dz * dz)       # D0 = 1 A^2/ps
    return dict(counts=counts.astype(np.int32), lags=lags, vb=vb, wb=wb, vc0=np.zeros(NCOS_F), wc0=wc0,
                v0=np.zeros(N_BINS), w0=wb @ wc0)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([x.strip() for x in out.stdout.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for k, nm in enumerate(names) if any(r[3 + k] == "Active" for r in self.rows)]
        power = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "power_w_max": max(power) if power else None, "samples": len(self.rows), "reasons": reasons}


def cpu_port_steps_per_s(wl, nsteps, seed=1):
    """Oracle port of the reference loop (scipy expm), one chain, one core."""
    from oracle import mcdiff_oracle as O
    cfg = O.ChainConfig(wl["counts"], wl["lags"], True, v_basis=wl["vb"], w_basis=wl["wb"], dv=DV, dw=DW,
                        temp=1.0, nmc=nsteps, num_mc_update=NMC_UPDATE)
    tu, ti = O.make_tape(np.random.default_rng(seed), nsteps, cfg)
    t0 = time.perf_counter()
    O.replay(cfg, wl["v0"], wl["w0"], tu, ti, v_coeff0=wl["vc0"], w_coeff0=wl["wc0"])
    return nsteps / (time.perf_counter() - t0)


_WL_CACHE = None


def _ref_worker(args):
    global _WL_CACHE
    nsteps, seed = args
    if _WL_CACHE is None:              # once per worker process, outside the timed steps (warm-up)
        _WL_CACHE = workload()
    wl = _WL_CACHE
    from oracle import mcdiff_oracle as O
    cfg = O.ChainConfig(wl["counts"], wl["lags"], True, v_basis=wl["vb"], w_basis=wl["wb"], dv=DV, dw=DW,
                        temp=1.0, nmc=nsteps, num_mc_update=NMC_UPDATE)
    tu, ti = O.make_tape(np.random.default_rng(seed), nsteps, cfg)
    O.replay(cfg, wl["v0"], wl["w0"], tu, ti, v_coeff0=wl["vc0"], w_coeff0=wl["wc0"])
    return nsteps


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port: scipy.linalg.expm per lag
    per step, lib/utils.py:307-368) on all host cores, one independent chain per core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"      # BLAS threads hurt at n = 100; one chain per core instead
    cores = os.cpu_count() or 1
    per_step = 150           # MC steps per worker per bench step (bounded sample, ~0.3 s per step)
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        for w in range(max(1, args.warmup)):
            pool.map(_ref_worker, [(8, 1000 + w * cores + c) for c in range(cores)], chunksize=1)
        t0 = time.perf_counter()
        done = 0
        for k in range(args.steps):
            done += sum(pool.map(_ref_worker, [(per_step, k * cores + c) for c in range(cores)], chunksize=1))
        el = time.perf_counter() - t0
    val = done / el
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic pbc 100 bins x 4 lags {10,20,30,40}, CosinusModel ncosF=10 ncosD=6, "
                               "one chain per host core", "mc_steps_per_step": per_step * cores},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d MC steps per core per bench step, %d cores, BLAS threads=1" % (per_step, cores)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    from mcdiff_b200 import _capi as K
    from mcdiff_b200.engine import Engine, DeviceProblem, ChainBatch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    eng = Engine(local)
    wl = workload()
    nch = CHAINS_PER_GPU
    L = len(LAGS)

    def params(seed=1):
        mp = K.McParams()
        mp.num_mc_update = NMC_UPDATE
        mp.dtemp = 0.0
        mp.min_lt = float(min(LAGS))
        mp.move_timezero = 0
        mp.sample_every = 100
        mp.refresh_every = 512
        mp.use_tape = 0
        mp.seed = seed
        mp.chain_offset = rank * nch
        mp.tape_stride = 0
        mp.method = {"auto": 0, "eigen": 1, "squaring": 2, "squaring_dfma": 3}[args.method]
        return mp

    # ---------------- device-resident arm: inputs already in HBM ----------------
    prob = DeviceProblem(eng, wl["counts"], wl["lags"], True, v_basis=wl["vb"], w_basis=wl["wb"])
    chains = ChainBatch(prob, nch)
    chains.set_state(wl["v0"], wl["w0"], wl["vc0"], wl["wc0"], dv=DV, dw=DW)
    chains.init_log_like()
    mp = params()
    nsamp = MC_PER_LAUNCH // 100
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=eng.device)   # > 126 MB L2
    for _ in range(args.warmup):
        chains.run(MC_PER_LAUNCH, mp, nsamples_cap=nsamp)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev0.record()
    for k in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations
        kev[k][0].record()
        chains.run(MC_PER_LAUNCH, mp, nsamples_cap=nsamp)
        kev[k][1].record()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches = eng.launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    sampler.stop_flag = True
    sampler.join(timeout=2)
    t = torch.tensor([ms], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    total_mc = float(world) * nch * MC_PER_LAUNCH * args.steps
    value = total_mc / (ms_max * 1e-3)

    # posterior summaries: the one collective of the design (per-chain acceptance + log L)
    st = chains.host_state()
    summary = torch.tensor(np.concatenate([st["scal"][:, :6], st["counters"][:, :3].astype(np.float64)], axis=1),
                           dtype=torch.float64, device=eng.device)
    if world > 1:
        gathered = torch.empty((world,) + tuple(summary.shape), dtype=torch.float64, device=eng.device)
        dist.all_gather_into_tensor(gathered, summary)
        summary_all = gathered.reshape(-1, summary.shape[1]).cpu().numpy()
    else:
        summary_all = summary.cpu().numpy()
    bad = int(st["counters"][:, 6].astype(bool).sum())

    # ---------------- end-to-end arm: host buffers in, host results out, every step ----------------
    h_counts = torch.from_numpy(wl["counts"]).pin_memory()
    st_h = dict(v=torch.from_numpy(st["v"]).pin_memory(), w=torch.from_numpy(st["w"]).pin_memory(),
                vc=torch.from_numpy(st["vcoef"]).pin_memory(), wc=torch.from_numpy(st["wcoef"]).pin_memory(),
                scal=torch.from_numpy(st["scal"]).pin_memory(),
                counters=torch.from_numpy(st["counters"]).pin_memory())
    e2e_steps = max(2, min(args.steps, 5))
    h2d = d2h = 0

    def e2e_step():
        nonlocal h2d, d2h
        p2 = DeviceProblem(eng, h_counts.numpy(), wl["lags"], True, v_basis=wl["vb"], w_basis=wl["wb"])
        c2 = ChainBatch(p2, nch, keep_basis=True)
        c2.v.copy_(st_h["v"], non_blocking=True)
        c2.w.copy_(st_h["w"], non_blocking=True)
        c2.vcoef.copy_(st_h["vc"], non_blocking=True)
        c2.wcoef.copy_(st_h["wc"], non_blocking=True)
        c2.scal.copy_(st_h["scal"], non_blocking=True)
        c2.counters.copy_(st_h["counters"], non_blocking=True)
        _, _, smp = c2.run(MC_PER_LAUNCH, mp, nsamples_cap=nsamp)
        out = c2.host_state()
        smp_h = smp.cpu()
        h2d = (h_counts.numel() * 4 + wl["vb"].nbytes + wl["wb"].nbytes + wl["lags"].nbytes
               + sum(x.numel() * x.element_size() for x in st_h.values()))
        d2h = sum(v_.nbytes for v_ in out.values() if v_ is not None) + smp_h.numel() * 8
        p2.close()
        return out

    e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = float(world) * nch * MC_PER_LAUNCH * e2e_steps / float(t.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline + CPU baseline (rank 0) ----------------
    peak_tf = eng.fp64_peak_tflops()
    fl = flops_per_mc_step(N_BINS, L)
    achieved_tf = nch * MC_PER_LAUNCH * fl / (kernel_ms * 1e-3) / 1e12
    cpu = None
    if world == 1:
        for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
            os.environ[var] = "1"
        nref = 2500
        cpu_val = cpu_port_steps_per_s(wl, nref)
        cpu = {"value": cpu_val, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "%d MC steps of one chain, oracle port of lib/mc.py loop (scipy expm), same 100x4 workload" % nref}
    sweeps = float(st["counters"][:, 7].sum()) / max(1.0, float(st["counters"][:, 5].sum()))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": "synthetic pbc 100 bins x 4 lags {10,20,30,40} (SURVEY 8d, seed 20240607), CosinusModel "
                        "ncosF=10 ncosD=6, F=0 D=1 start, dv=dw=0.05, T=1",
            "chains_per_gpu": nch, "mc_steps_per_launch": MC_PER_LAUNCH, "mc_steps_per_bench_step": nch * MC_PER_LAUNCH,
            "l2": "256 MiB buffer rewritten between timed launches; working set (160 KB counts + per-chain state) "
                  "is smem/L2 resident by design, the kernel is FP64/shared-memory bound, not HBM bound",
            "philox_seed": 1, "propagator_engine": ("eigen" if (args.method == "eigen" or not prob.squaring_ok) else ("squaring_dfma" if args.method == "squaring_dfma" else "squaring_dmma")),
            "jacobi_sweeps_or_squarings_per_mc_step": sweeps, "chains_with_status": bad,
            "mean_acceptance": float((summary_all[:, 6] + summary_all[:, 7]).sum() / max(1.0, float(world) * st["counters"][:, 5].sum())),
        },
        "roofline": {"bound": "tensor", "bound_detail": "FP64 tensor pipe (DMMA m8n8k4); same nominal peak as the FP64 vector pipe",
                     "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / peak_tf if peak_tf else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this configuration
                     # (ncu --set full, profiles/r01_ncu_full_mc_run_final.csv): 4.0 MB
                     "traffic": NCU_DRAM_BYTES_PER_LAUNCH if (nch, MC_PER_LAUNCH) == (1024, 200) else None,
                     "traffic_unit": "bytes of DRAM traffic per launch (ncu); the kernel is not HBM bound",
                     "peak_source": "measured here by mcd_fp64_peak (dense DFMA loop); MEASURED_PEAKS.json has no FP64 "
                                    "figure; nominal 148 SM x 64 DFMA x 2 x 1.965 GHz = 37.2",
                     "algorithmic_flops_per_mc_step": fl, "kernel": "mcd_mc_run_kernel<true>",
                     "kernel_ms_per_launch": kernel_ms},
        "cpu_baseline": cpu,
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps},
        "gpu_launches": int(launches),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--method", default="auto", choices=["auto", "eigen", "squaring", "squaring_dfma"],
                    help="propagator engine (auto = symmetric scaling-and-squaring for this workload)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
