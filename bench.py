#!/usr/bin/env python
"""bench.py -- MC steps/s of the mcdiff Monte-Carlo likelihood loop on B200.

Metric (BASELINE.json): MC steps/sec summed over all chains at 100 bins x 4 lag times, plus the fraction of the FP64
roofline.  One bench "step" = ONE launch of the fused chain kernel that advances every chain of this rank by
MC_PER_LAUNCH Monte-Carlo steps (one MC step = proposal + rate matrix + 4 propagators expm(lag R) + log L +
Metropolis, i.e. one iteration of lib/mc.py:27 of the reference).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Legs of the b200 arm (one JSON line, rank 0):
  value      headline: BASELINE shape, 1024 chains per GPU resident in HBM, K timed launches (weak scaling)
  e2e        the same metric through the USER API: find_parameters(files..., nchains=) = run-mcdiff: transition
             files in, <out>.dat / .pic / .pic.ave.dat / .chains.npz out, uploads, the final NCCL gathers and the
             device->host reads inside the timed region
  strong     e2e with a FIXED total of 4096 chains over the N GPUs (strong-scaling point)
  secondary  config 3 (200 bins x 8 lags, Fourier 12/8, blocked engine) and config 5 (radial model), 512 chains per
             GPU each: MC steps/s, algorithmic flops, fraction of the FP64 peak, kernel name
  cpu_baseline  (N = 1) the UNMODIFIED reference's do_mc_cycles (baseline/_ref, scripts/install_reference.sh) on one
             host core; the oracle port only when the reference is not installed
`--impl reference` times the reference's own loop on all host cores (one chain per core) + the two single-process
variants of BASELINE.md 4.3.
"""
import argparse
import contextlib
import io
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
REF_DIR = os.path.join(ROOT, "baseline", "_ref")

N_BINS = 100
LAGS = (10.0, 20.0, 30.0, 40.0)
CHAINS_PER_GPU = 1024
MC_PER_LAUNCH = 200
NCOS_F, NCOS_D = 10, 6           # the reference's own example model (examples/example-HexdWat-run1.sh)
DV = DW = 0.05
NMC_UPDATE = 1000.0
STRONG_CHAINS = 4096
METRIC = "MC steps/sec (all chains, 100 bins x 4 lags)"
UNIT = "MC steps/s"
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (1024 chains x 200 steps), ncu --set full, see profiles/
NCU_DRAM_BYTES_PER_LAUNCH = {"mcd_mc_run_kernel<true>": 3954944 + 56320, "mcd_mc_run_sq_kernel": 3990528 + 33280}


def flops_per_mc_step(n, L):
    """F_step = 9 n^3 + L (2 n^3 + 3 n^2)  (17.1 MFLOP at 100 x 4)."""
    return 9.0 * n ** 3 + L * (2.0 * n ** 3 + 3.0 * n ** 2)


def flops_per_radial_step(n, L, lmax, dim_rad):
    """SURVEY.md 8(d), radial: lmax 9 n^3 + L lmax 2 n^3 + L (2 dim_rad lmax n^2 + 3 dim_rad n^2)."""
    return lmax * 9.0 * n ** 3 + L * lmax * 2.0 * n ** 3 + L * (2.0 * dim_rad * lmax * n ** 2 + 3.0 * dim_rad * n ** 2)


# ---------------------------------------------------------------------------------------------------------------
# synthetic inputs.  Input GENERATION only: a truth profile, its propagator and multinomial counts.
# Neither the product nor the oracle is involved, so both arms read byte-identical files.
# ---------------------------------------------------------------------------------------------------------------
def truth_rate_matrix(F, w):
    """Periodic rate matrix of lib/utils.py:77-103 for the truth profile (generator side only)."""
    n = len(F)
    R = np.zeros((n, n))
    for i in range(n):
        j = (i + 1) % n
        R[j, i] = np.exp(w[i] - 0.5 * (F[j] - F[i]))
        R[i, j] = np.exp(w[i] + 0.5 * (F[j] - F[i]))
    R[np.arange(n), np.arange(n)] = -R.sum(axis=0)
    return R


def synthetic_counts(n, lags, seed=20240607, total=1.0e6, dz=0.5):
    import scipy.linalg
    i = np.arange(n)
    F = 2.0 * np.cos(2 * np.pi * (i + 0.5) / n) + 0.5 * np.cos(4 * np.pi * (i + 0.5) / n)
    D = 0.5 + 0.3 * np.cos(2 * np.pi * (i + 1.0) / n)
    w = np.log(D) - np.log(dz * dz / 1.0)
    rng = np.random.default_rng(seed)
    pi = np.exp(-F)
    pi /= pi.sum()
    M = np.round(total * pi).astype(np.int64)
    R = truth_rate_matrix(F, w)
    out = np.zeros((len(lags), n, n), dtype=np.int64)
    for il, lag in enumerate(lags):
        P = np.clip(scipy.linalg.expm(lag * R), 0.0, None)
        P /= P.sum(axis=0, keepdims=True)
        for j in range(n):
            out[il, :, j] = rng.multinomial(M[j], P[:, j])
    return out, np.linspace(-n * dz / 2, n * dz / 2, n + 1)


def write_square_files(outdir, counts, lags, edges, tag="pbc"):
    """Reference text format (lib/tools/extract.py:19-36); plain writer so that the reference arm needs no product code."""
    files = []
    n = counts.shape[-1]
    for l, lag in enumerate(lags):
        fn = os.path.join(outdir, "transitions.nbins%d.%d.%s.dat" % (n, int(lag), tag))
        with open(fn, "w") as f:
            f.write("#lt    {}\n#count {}\n#dt    {}\n#dn    {}\n".format(float(lag), tag, 1.0, int(lag)))
            f.write("#edges  " + " ".join(str(b) for b in edges) + "\n")
            for row in counts[l]:
                f.write(" ".join(str(x) for x in row) + "\n")
        files.append(fn)
    return files


def headline_files(outdir):
    counts, edges = synthetic_counts(N_BINS, LAGS)
    return write_square_files(outdir, counts, LAGS, edges)


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


# ---------------------------------------------------------------------------------------------------------------
# CPU side: the unmodified reference (baseline/_ref) or, when it is not installed, the oracle port
# ---------------------------------------------------------------------------------------------------------------
def reference_available():
    return os.path.isdir(os.path.join(REF_DIR, "mcdiff"))


def cpu_model_string():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def _reference_chain(files, nmc, seed):
    """`nmc` steps of ONE chain of the unmodified reference: mcdiff.mc.do_mc_cycles (lib/mc.py:20-51) on the MCState
    that lib/mc.py:62-86 builds.  Returns (seconds inside do_mc_cycles, final log-likelihood)."""
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import mcdiff.mc as rmc
    from mcdiff.MCState import MCState
    from mcdiff.log import Logger
    from mcdiff.transitions import Transitions
    import warnings
    warnings.filterwarnings("ignore", category=RuntimeWarning)    # the reference overflows exp() in its accept test
    np.random.seed(seed)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        MC = MCState(True, -1)
        MC.set_MC_params(DV, DW, 0.5, 1.0, 0.1, 1.0, nmc, NMC_UPDATE, False, -1, temp_end=1.0)
        MC.set_model("CosinusModel", Transitions(files), NCOS_F, NCOS_D, 0, None)
        logger = Logger(MC)
        t0 = time.perf_counter()
        rmc.do_mc_cycles(MC, logger)
        el = time.perf_counter() - t0
    return el, float(MC.log_like)


def _port_chain(files, nmc, seed):
    """Fallback when baseline/_ref is absent: the oracle's restatement of the same loop."""
    from oracle import mcdiff_oracle as O
    counts = np.array([O.read_counts(fn)[1] for fn in files])
    vb, wb = O.cos_basis_center(N_BINS, NCOS_F), O.cos_basis_border(N_BINS, N_BINS, NCOS_D)
    wc0 = np.zeros(NCOS_D)
    wc0[0] = -np.log(0.25)
    cfg = O.ChainConfig(counts, np.array(LAGS), True, v_basis=vb, w_basis=wb, dv=DV, dw=DW, temp=1.0, nmc=nmc,
                        num_mc_update=NMC_UPDATE)
    tu, ti = O.make_tape(np.random.default_rng(seed), nmc, cfg)
    t0 = time.perf_counter()
    out = O.replay(cfg, np.zeros(N_BINS), wb @ wc0, tu, ti, v_coeff0=np.zeros(NCOS_F), w_coeff0=wc0)
    return time.perf_counter() - t0, float(out["ll"])


def cpu_chain(files, nmc, seed):
    return (_reference_chain if reference_available() else _port_chain)(files, nmc, seed)


def _ref_worker(args):
    files, nmc, seed = args
    el, _ = cpu_chain(files, nmc, seed)
    return nmc, el


def _single_process_rate(files, nmc, threads):
    """One chain in a fresh process with `threads` BLAS threads (BASELINE.md 4.3 variants i and ii)."""
    env = dict(os.environ)
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        env[var] = str(threads)
    code = ("import sys, json; sys.path.insert(0, %r); import bench; "
            "el, ll = bench.cpu_chain(%r, %d, 1); print(json.dumps([el, ll]))" % (ROOT, list(files), nmc))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    if out.returncode != 0:
        return None
    el, _ = json.loads(out.stdout.strip().splitlines()[-1])
    return nmc / el


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores, one independent chain
    per core (BLAS threads = 1: they hurt at n = 100), plus the single-process variants."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"
    cores = os.cpu_count() or 1
    per_step = 150           # MC steps per worker per bench step (bounded sample, ~0.4 s per step)
    tmp = tempfile.mkdtemp(prefix="mcd_bench_ref_")
    try:
        files = headline_files(tmp)
        kind = "reference" if reference_available() else "port"
        ctx = mp.get_context("spawn")
        with ctx.Pool(cores) as pool:
            for w in range(max(1, args.warmup)):
                pool.map(_ref_worker, [(files, 8, 1000 + w * cores + c) for c in range(cores)], chunksize=1)
            t0 = time.perf_counter()
            done = 0
            for k in range(args.steps):
                done += sum(r[0] for r in pool.map(_ref_worker, [(files, per_step, k * cores + c) for c in range(cores)],
                                                   chunksize=1))
            el = time.perf_counter() - t0
        val = done / el
        one = _single_process_rate(files, 600, 1)
        allthreads = _single_process_rate(files, 600, cores)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    import scipy
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic pbc 100 bins x 4 lags {10,20,30,40} (SURVEY 8d, seed 20240607), CosinusModel "
                               "ncosF=10 ncosD=6, F=0 D=1 start, dv=dw=0.05, T=1; one chain per host core",
                   "mc_steps_per_step": per_step * cores},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%d MC steps per core per bench step, %d cores, BLAS threads=1; %s" % (
                             per_step, cores, "mcdiff.mc.do_mc_cycles of the unmodified reference (baseline/_ref)"
                             if kind == "reference" else "oracle port (baseline/_ref not installed)"),
                         "variants": {"one_process_1_thread": one, "one_process_all_threads": allthreads,
                                      "one_process_per_core_aggregate": val},
                         "cpu_model": cpu_model_string(), "numpy": np.__version__, "scipy": scipy.__version__},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def _mc_params(K, rank_offset, lags, method=0, seed=1):
    mp = K.McParams()
    mp.num_mc_update = NMC_UPDATE
    mp.dtemp = 0.0
    mp.min_lt = float(min(lags))
    mp.move_timezero = 0
    mp.sample_every = 100
    mp.refresh_every = 512
    mp.use_tape = 0
    mp.seed = seed
    mp.chain_offset = rank_offset
    mp.tape_stride = 0
    mp.method = method
    return mp


def _time_launches(torch, dist, world, flush, launch, steps, warmup):
    """W untimed launches, then K timed ones bracketed by barrier + synchronize; returns (total ms max over ranks,
    mean kernel ms of this rank)."""
    for _ in range(warmup):
        launch()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ev0.record()
    for k in range(steps):
        flush.zero_()                       # L2 flush between timed iterations
        kev[k][0].record()
        launch()
        kev[k][1].record()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), kernel_ms


def _bcast_path(dist, world, rank, make):
    """rank 0 creates a scratch directory (and its files), every rank learns the path."""
    obj = [make() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(obj, src=0)
    return obj[0]


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    from mcdiff_b200 import _capi as K
    from mcdiff_b200 import engine as E
    from mcdiff_b200.engine import ChainBatch, DeviceProblem, Engine, RadialChains, RadialProblem
    from mcdiff_b200.mc import find_parameters
    from mcdiff_b200.model import CosinusModel, RadCosinusModel
    from mcdiff_b200.transitions import RadTransitions, Transitions, write_Tmat_cube

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
    method = {"auto": 0, "eigen": 1, "squaring": 2, "squaring_dfma": 3}[args.method]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=eng.device)   # > 126 MB L2
    peak_tf = eng.fp64_peak_tflops()

    def make_inputs():
        d = tempfile.mkdtemp(prefix="mcd_bench_")
        headline_files(d)
        c3, e3 = synthetic_counts(200, C3_LAGS, seed=3)
        os.makedirs(os.path.join(d, "c3"))
        write_square_files(os.path.join(d, "c3"), c3, C3_LAGS, e3)
        c4, e4 = synthetic_counts(N_BINS, C4_LAGS, seed=4)
        os.makedirs(os.path.join(d, "c4"))
        write_square_files(os.path.join(d, "c4"), c4, C4_LAGS, e4)
        return d

    C3_LAGS = (5.0, 10.0, 15.0, 20.0, 25.0, 30.0, 35.0, 40.0)
    # config 4: 16 independent single-lag problems, 256 chains each
    C4_LAGS = (1.0, 2.0, 3.0, 4.0, 5.0, 8.0, 10.0, 12.0, 15.0, 20.0, 25.0, 30.0, 35.0, 40.0, 45.0, 50.0)
    tmp = _bcast_path(dist, world, rank, make_inputs)
    files = [os.path.join(tmp, "transitions.nbins%d.%d.pbc.dat" % (N_BINS, int(l))) for l in LAGS]
    sink = io.StringIO()

    # ---------------- headline: device-resident arm, inputs already in HBM ----------------
    data = Transitions(files)
    model = CosinusModel(data, 1.0, NCOS_F, NCOS_D)           # the product's own bases and start (F = 0, D = 1)
    prob = DeviceProblem(eng, np.asarray(data.list_trans).astype(np.int64), data.list_lt, True,
                         v_basis=model.v_basis, w_basis=model.w_basis)
    nch = CHAINS_PER_GPU
    L = len(LAGS)
    nsamp = MC_PER_LAUNCH // 100

    def headline(nchains):
        chains = ChainBatch(prob, nchains)
        chains.set_state(model.v, model.w, model.v_coeff, model.w_coeff, dv=DV, dw=DW)
        chains.init_log_like()
        mp = _mc_params(K, rank * nchains, LAGS, method)
        return chains, (lambda: chains.run(MC_PER_LAUNCH, mp, nsamples_cap=nsamp))

    chains, launch = headline(nch)
    for _ in range(args.warmup):
        launch()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = eng.launch_count()
    ms_max, kernel_ms = _time_launches(torch, dist, world, flush, launch, args.steps, 0)
    launches = eng.launch_count() - launches0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    value = float(world) * nch * MC_PER_LAUNCH * args.steps / (ms_max * 1e-3)
    st = chains.host_state()
    bad = int(st["counters"][:, 6].astype(bool).sum())
    squarings = float(st["counters"][:, 7].sum()) / max(1.0, float(st["counters"][:, 5].sum()))
    acc = float((st["counters"][:, 0] + st["counters"][:, 1]).sum() / max(1.0, float(st["counters"][:, 5].sum())))
    # a chain count that fills whole waves of 148 CTAs (7 x 148): what the wave quantisation of 1024 costs
    nch148 = 7 * 148
    chains148, launch148 = headline(nch148)
    ms148, _ = _time_launches(torch, dist, world, flush, launch148, max(2, min(args.steps, 4)), 3)
    value148 = float(world) * nch148 * MC_PER_LAUNCH * max(2, min(args.steps, 4)) / (ms148 * 1e-3)
    del chains148

    # ---------------- secondary: config 3 (200 x 8, blocked engine) and config 5 (radial) ----------------
    secondary = {}
    sec_steps = max(2, min(args.steps, 3))
    # config 3
    files3 = [os.path.join(tmp, "c3", "transitions.nbins200.%d.pbc.dat" % int(l)) for l in C3_LAGS]
    data3 = Transitions(files3)
    model3 = CosinusModel(data3, 1.0, 12, 8)
    prob3 = DeviceProblem(eng, np.asarray(data3.list_trans).astype(np.int64), data3.list_lt, True,
                          v_basis=model3.v_basis, w_basis=model3.w_basis)
    ch3 = ChainBatch(prob3, 512)
    ch3.set_state(model3.v, model3.w, model3.v_coeff, model3.w_coeff, dv=DV, dw=DW)
    ch3.init_log_like()
    mp3 = _mc_params(K, rank * 512, C3_LAGS, 0)
    C3_MC = 20
    ms3, k3 = _time_launches(torch, dist, world, flush, lambda: ch3.run(C3_MC, mp3), sec_steps, 3)
    fl3 = flops_per_mc_step(200, len(C3_LAGS))
    st3 = ch3.host_state()
    secondary["config3_200bins_8lags"] = {
        "value": float(world) * 512 * C3_MC * sec_steps / (ms3 * 1e-3), "unit": UNIT, "chains_per_gpu": 512,
        "mc_steps_per_launch": C3_MC, "launches": sec_steps,
        "kernel": "mcd_mc_run_blk_kernel (444 chains = 3 full waves) + mcd_mc_run_blk2_kernel (68 chains on two-CTA clusters)",
        "kernel_ms_per_launch": k3, "algorithmic_flops_per_mc_step": fl3,
        "achieved_tflops": 512 * C3_MC * fl3 / (k3 * 1e-3) / 1e12,
        "frac": 512 * C3_MC * fl3 / (k3 * 1e-3) / 1e12 / peak_tf,
        "squarings_per_mc_step": float(st3["counters"][:, 7].sum()) / max(1.0, float(st3["counters"][:, 5].sum())),
        "chains_with_status": int(st3["counters"][:, 6].astype(bool).sum()),
        "workload": "synthetic pbc 200 bins x 8 lags {5..40} (seed 3), CosinusModel ncosF=12 ncosD=8"}
    del ch3
    # config 5: radial cube drawn from the product's own radial propagator at a truth profile (generation only)
    RN, RDIM, RLMAX, RLAGS = 30, 41, 20, (5.0, 10.0)

    def make_radial():
        from scipy import special
        # v, w are frozen in radial runs: the truth uses the flat profiles the model starts from (F = 0, D = 1)
        v = np.zeros(RN)
        w = np.full(RN, -np.log(0.25))
        wrad = np.log((0.4 + 0.15 * np.cos(2 * np.pi * (np.arange(RN) + 0.5) / RN)) / 0.25)
        zeros = special.jn_zeros(0, RLMAX)
        r = np.arange(RDIM, dtype=np.float64) + 0.5
        bess = np.array([2 * r * special.j0(r / RDIM * z) / special.j1(z) ** 2 / RDIM ** 2 for z in zeros])
        p0 = RadialProblem(eng, v, w, True, RLAGS, np.zeros((len(RLAGS), RDIM, RN, RN)), bess, zeros)
        _, P, _ = p0.loglike_device(eng.to_device(wrad[None, :], torch.float64), want_P=True)
        P = np.clip(E.to_host(P)[0], 0.0, None)
        p0.close()
        cube = np.random.default_rng(0).poisson(2000.0 * P * np.exp(-v)[None, None, None, :])
        fns = []
        for l, lag in enumerate(RLAGS):
            fn = os.path.join(tmp, "rad.%d.dat" % int(lag))
            write_Tmat_cube(cube[l].astype(int), fn, float(lag), "pbc", edges=np.linspace(-7.5, 7.5, RN + 1),
                            redges=np.arange(RDIM) * 0.5, dt=1.0, dn=int(lag))
            fns.append(fn)
        return fns

    if rank == 0:
        make_radial()
    if world > 1:
        dist.barrier()
    rfiles = [os.path.join(tmp, "rad.%d.dat" % int(l)) for l in RLAGS]
    rdata = RadTransitions(rfiles)
    rmodel = RadCosinusModel(rdata, 1.0, 0, 0, 0, 4)
    from scipy import special
    zeros = special.jn_zeros(0, RLMAX)
    rr = np.arange(rmodel.dim_rad, dtype=np.float64) + 0.5
    bess = np.array([2 * rr * special.j0(rr / rmodel.dim_rad * z) / special.j1(z) ** 2 / rmodel.dim_rad ** 2 for z in zeros])
    rch = RadialChains(eng, rmodel.v, rmodel.w, True, rdata.list_lt, rdata.list_trans, bess, zeros, 512,
                       wrad_basis=rmodel.wrad_basis)
    rch.set_state(rmodel.wrad, rmodel.wrad_coeff, dwrad=0.05)
    rch.init_log_like()
    R_MC = 20
    msr, kr = _time_launches(torch, dist, world, flush,
                             lambda: rch.run(R_MC, num_mc_update=NMC_UPDATE, seed=1, chain_offset=rank * 512),
                             sec_steps, 3)
    flr = flops_per_radial_step(RN, len(RLAGS), RLMAX, rmodel.dim_rad)
    rst = rch.host_state()
    secondary["config5_radial"] = {
        "value": float(world) * 512 * R_MC * sec_steps / (msr * 1e-3), "unit": UNIT, "chains_per_gpu": 512,
        "mc_steps_per_launch": R_MC, "launches": sec_steps, "kernel": "mcd_rad_mc_run_kernel",
        "kernel_ms_per_launch": kr, "algorithmic_flops_per_mc_step": flr,
        "achieved_tflops": 512 * R_MC * flr / (kr * 1e-3) / 1e12,
        "frac": 512 * R_MC * flr / (kr * 1e-3) / 1e12 / peak_tf,
        "chains_with_status": int(rst["counters"][:, 3].astype(bool).sum()),
        "mean_acceptance": float(rst["counters"][:, 0].sum() / max(1.0, float(rst["counters"][:, 2].sum()))),
        "workload": "radial model n=30 bins, dim_rad=%d, lmax=20, lags {5,10}, RadCosinusModel ncosDrad=4 "
                    "(SURVEY 8d recipe, Poisson(2000 P e^-v), seed 0)" % rmodel.dim_rad}
    del rch

    # ---------------- config 4: lag-time scan + extrapolation, through files (run_lag_scan shards the pairs) ----------
    from mcdiff_b200.infty import extrapolate, run_lag_scan
    by_lag = {l: os.path.join(tmp, "c4", "transitions.nbins%d.%d.pbc.dat" % (N_BINS, int(l))) for l in C4_LAGS}
    C4_CHAINS, C4_NMC = 256, 400
    c4_best = []
    for r in range(3):                          # the first call is the warm-up
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(sink):
            outs4 = run_lag_scan(by_lag, os.path.join(tmp, "c4out"), nmc=C4_NMC, nmc_update=NMC_UPDATE, nchains=C4_CHAINS,
                                 device=local, dv=DV, dw=DW)
            if rank == 0:
                lts4, Dreal4, _ = extrapolate(outs4, os.path.join(tmp, "c4out", "scan."))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        if r > 0:
            c4_best.append(time.perf_counter() - t0)
        sink.seek(0)
        sink.truncate(0)
    t4 = torch.tensor([float(np.mean(c4_best))], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t4, op=dist.ReduceOp.MAX)
    secondary["config4_lag_scan"] = {
        "value": len(C4_LAGS) * C4_CHAINS * C4_NMC / float(t4.item()), "unit": UNIT, "problems": len(C4_LAGS),
        "chains_per_problem": C4_CHAINS, "mc_steps": C4_NMC, "seconds": float(t4.item()),
        "extrapolated_D_finite": bool(np.all(np.isfinite(Dreal4))) if rank == 0 else None,
        "api": "mcdiff_b200.infty.run_lag_scan + extrapolate (= mcdiff-lagscan): 16 transition files in, 16 .dat + "
               "Dreal/t0 out, wall clock incl. file reading, uploads, first likelihoods, gathers, output; (problem, "
               "chain) pairs block-partitioned over %d GPU(s), one CUDA stream per problem" % world,
        "workload": "16 single-lag problems, 100 bins, lags {1,2,3,4,5,8,10,12,15,20,25,30,35,40,45,50} (seed 4), "
                    "CosinusModel 10/6"}

    # ---------------- trajectory -> counts (HBM-bound integer work) ----------------
    # 2^28 coordinates ([frames][particles], 2 GiB of fp64 -- far larger than L2), 100 bins, the headline's 4 shifts.
    # Algorithmic bytes: every coordinate is read once = 8 B per element; the count matrices are 333 KB.
    from mcdiff_b200 import extract as X
    TR_T, TR_M, TR_BINS, TR_ZPBC = 1 << 19, 512, 100, 50.0
    gen = torch.Generator(device=eng.device).manual_seed(11 + rank)
    ztr = torch.cumsum(torch.randn((TR_T, TR_M), generator=gen, device=eng.device, dtype=torch.float64) * 0.35, dim=0)
    tr_edges = np.arange(-TR_BINS * (TR_ZPBC / TR_BINS) / 2, (TR_BINS + 1) * (TR_ZPBC / TR_BINS) / 2., TR_ZPBC / TR_BINS)
    tr_shifts = [int(l) for l in LAGS]
    tr_out = torch.zeros((len(tr_shifts), TR_BINS + 2, TR_BINS + 2), dtype=torch.int64, device=eng.device)
    tr_edges_dev = eng.to_device(tr_edges, torch.float64)
    tr_launch = lambda: X.transition_counts_device(ztr, tr_edges_dev, tr_shifts, period=TR_ZPBC, engine=eng, out=tr_out)
    mstr, ktr = _time_launches(torch, dist, world, flush, tr_launch, sec_steps, 3)
    tr_total = int(tr_out.sum().item())
    tr_expect = (3 + sec_steps) * sum((TR_T - s) * TR_M for s in tr_shifts)
    tr_bytes = 8.0 * TR_T * TR_M
    hbm_peak = None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_peak = float(json.load(f)["hbm_gbs"])
    except Exception:
        hbm_peak = None
    hbm_src = "MEASURED_PEAKS.json hbm_gbs" if hbm_peak else "fallback of B200_PROFILING.md (copy bandwidth)"
    hbm_peak = hbm_peak or 6550.0
    secondary["trajectory_counts"] = {
        "value": float(world) * TR_T * TR_M * sec_steps / (mstr * 1e-3), "unit": "coordinates/s",
        "frames": TR_T, "particles": TR_M, "bins": TR_BINS, "shifts": tr_shifts, "launches": sec_steps,
        "kernels": ["mcd_traj_digitize_kernel", "mcd_traj_count_kernel"], "ms_per_call": ktr,
        "roofline": {"bound": "hbm", "achieved": tr_bytes / (ktr * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": tr_bytes / (ktr * 1e-3) / 1e9 / hbm_peak,
                     # dram__bytes_read + write of the two kernels of one call (ncu --set full, profiles/r02_ncu_full_traj_*.csv)
                     "traffic": 2147514000 + 214921000 + 268601000 + 4628000, "peak_source": hbm_src,
                     "algorithmic_bytes_per_coordinate": 8},
        "pairs_counted_ok": bool(tr_total == tr_expect),
        "workload": "synthetic random walks, periodic fold at zpbc=50, 100 bins (+2 outside), shifts = the headline's "
                    "lags; inputs resident in HBM, counts accumulated on device"}
    del ztr

    # ---------------- end to end through the user API: files in, result files out ----------------
    def e2e_leg(nchains_total, nmc, reps, tag):
        out = os.path.join(tmp, "%s.r%d.dat" % (tag, rank)) if rank != 0 else os.path.join(tmp, tag + ".dat")
        best = []
        h2d = d2h = 0
        for r in range(reps + 1):               # the first call is the warm-up
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            E.TRAFFIC["h2d"] = E.TRAFFIC["d2h"] = 0
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(sink):
                find_parameters(files, True, "CosinusModel", DV, DW, 0.5, 1.0, 0.1, 1.0, 1.0, nmc, NMC_UPDATE, 1, out,
                                NCOS_F, NCOS_D, 0, False, None, -1, -1, False, None, nchains=nchains_total, device=local)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            el = time.perf_counter() - t0
            sink.seek(0)
            sink.truncate(0)
            if r > 0:
                best.append(el)
                h2d, d2h = E.TRAFFIC["h2d"], E.TRAFFIC["d2h"]
        t = torch.tensor([float(np.mean(best))], dtype=torch.float64, device=eng.device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return nchains_total * nmc / float(t.item()), float(t.item()), int(h2d), int(d2h)

    e2e_reps = max(2, min(args.steps, 3))
    E2E_NMC = 1000
    e2e_value, e2e_s, h2d, d2h = e2e_leg(world * nch, E2E_NMC, e2e_reps, "e2e")
    strong_value, strong_s, _, _ = e2e_leg(STRONG_CHAINS, E2E_NMC, 2, "strong")

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline + CPU baseline (rank 0) ----------------
    fl = flops_per_mc_step(N_BINS, L)
    achieved_tf = nch * MC_PER_LAUNCH * fl / (kernel_ms * 1e-3) / 1e12
    engine_name = ("eigen" if (args.method == "eigen" or not prob.squaring_ok)
                   else ("squaring_dfma" if args.method == "squaring_dfma" else "squaring_dmma"))
    # executed tensor-core flops: (squarings + L - 1) symmetric products of 12 warp tasks x 8 tiles x (np / 4) m8n8k4
    # (512 flop each); squarings per step is counted by the kernel.  ncu's DMMA instruction count agrees (profiles/).
    executed = (squarings + L - 1) * 96 * ((N_BINS + 3) // 4) * 512.0 if engine_name == "squaring_dmma" else None
    cpu = None
    if world == 1:
        for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
            os.environ[var] = "1"
        nref = 4000
        rate = _single_process_rate(files, nref, 1)
        kind = "reference" if reference_available() else "port"
        cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": "%d MC steps of one chain, %s, same 100x4 files, BLAS threads=1" % (
                   nref, "mcdiff.mc.do_mc_cycles of the unmodified reference (baseline/_ref)" if kind == "reference"
                   else "oracle port of lib/mc.py (baseline/_ref not installed)"),
               "cpu_model": cpu_model_string(), "host_cores": os.cpu_count()}
    # the tensor-core squaring engine has its own (leaner) kernel; the other engines share mcd_mc_run_kernel<true>
    kname = "mcd_mc_run_sq_kernel" if engine_name == "squaring_dmma" else "mcd_mc_run_kernel<true>"
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
            "philox_seed": 1, "propagator_engine": engine_name,
            "jacobi_sweeps_or_squarings_per_mc_step": squarings, "chains_with_status": bad,
            "mean_acceptance": acc,
            "value_at_1036_chains_per_gpu": value148,
        },
        "roofline": {"bound": "tensor", "bound_detail": "FP64 tensor pipe (DMMA m8n8k4); same nominal peak as the FP64 vector pipe",
                     "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / peak_tf if peak_tf else None,
                     "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(kname) if (nch, MC_PER_LAUNCH) == (1024, 200) else None,
                     "traffic_unit": "bytes of DRAM traffic per launch (ncu); the kernel is not HBM bound",
                     "peak_source": "measured here by mcd_fp64_peak (dense DFMA loop); MEASURED_PEAKS.json has no FP64 "
                                    "figure; nominal 148 SM x 64 DFMA x 2 x 1.965 GHz = 37.2",
                     "algorithmic_flops_per_mc_step": fl, "executed_flops_per_mc_step": executed,
                     "executed_frac": (nch * MC_PER_LAUNCH * executed / (kernel_ms * 1e-3) / 1e12 / peak_tf) if executed else None,
                     "kernel": kname, "kernel_ms_per_launch": kernel_ms},
        "cpu_baseline": cpu,
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_reps, "seconds_per_step": e2e_s,
                "api": "mcdiff_b200.mc.find_parameters (= run-mcdiff): %d transition files in, .dat/.pic/.pic.ave.dat/"
                       ".chains.npz out, %d chains x %d MC steps, gathers included" % (L, world * nch, E2E_NMC)},
        "strong": {"value": strong_value, "unit": UNIT, "chains_total": STRONG_CHAINS, "mc_steps": E2E_NMC,
                   "seconds": strong_s, "api": "find_parameters, fixed total chains over %d GPU(s)" % world},
        "secondary": secondary,
        "gpu_launches": int(launches),
    }
    print(json.dumps(line), flush=True)
    shutil.rmtree(tmp, ignore_errors=True)
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
