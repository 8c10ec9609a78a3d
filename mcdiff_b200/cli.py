"""Command line of run-mcdiff.  Every option of the reference (scripts/run-mcdiff:18-96) is kept
with its name, type and default -- including --nmc_update being parsed as a float and the
annealing default --Tend 0.01 -- so existing recipes (examples/*.sh) run unchanged."""
from __future__ import annotations

from optparse import OptionParser


def build_parser():
    p = OptionParser(usage="usage: %prog [options] file1 [file2 ...]", version="%prog b200")
    p.add_option("-o", "--outfile", dest="outfile", default=None, metavar="FILE",
                 help="filename FILE where F and D results will be stored")
    p.add_option("--nopbc", dest="pbc", default=True, action="store_false",
                 help="when no periodic boundary conditions should be used")
    p.add_option("--initf", dest="initfile", default=None, metavar="FILE",
                 help="filename FILE with initial guess for F, D, dv, dw")
    p.add_option("-n", "--nmc", dest="nmc", default=1000, type="int", help="number of Monte Carlo cycles")
    p.add_option("-T", dest="temp", default=1.0, type="float", metavar="TEMP",
                 help="temperature in parameter space in Monte Carlo run")
    p.add_option("--Tend", dest="temp_end", default=0.01, type="float",
                 help="final temperature, attained by updating every nmc_update steps")
    p.add_option("--nbins", dest="nbins", default=None, type="int", help="number of bins (unused)")
    p.add_option("--dv", dest="dv", default=0.5, type="float", help="potential Monte Carlo move width")
    p.add_option("--dw", dest="dw", default=0.5, type="float", help="log(diffusion) Monte Carlo move width")
    p.add_option("--dwrad", dest="dwrad", default=0.5, type="float",
                 help="log(radial diffusion) Monte Carlo move width")
    p.add_option("--D0", dest="D0", default=1.0, type="float", help="first guess D0")
    p.add_option("--model", dest="model", default="Model", help="set model, default is individual points")
    p.add_option("--nmc_update", dest="nmc_update", default=100, type="float", metavar="nmc_update",
                 help="number of moves between Monte Carlo step width/temp adjustments (0 if no adjustment)")
    p.add_option("-s", "--seed", dest="seed", default=None, type="int", metavar="SEED",
                 help="SEED of the random generator (Philox key)")
    p.add_option("--ncosF", dest="ncosF", default=0, type="int", metavar="NCOSF",
                 help="number of cosine basis functions of the free energy profile")
    p.add_option("--ncosD", dest="ncosD", default=0, type="int", metavar="NCOSD",
                 help="number of cosine basis functions of the diffusion profile")
    p.add_option("--ncosDrad", dest="ncosDrad", default=0, type="int", metavar="NCOSDRAD",
                 help="number of cosine basis functions of the radial diffusion profile")
    p.add_option("-k", dest="k", default=-1, type="float", metavar="K",
                 help="spring constant to keep D profile smoother")
    p.add_option("--t0", dest="move_timezero", default=False, action="store_true",
                 help="Also make the time offset a variable.")
    p.add_option("--dt0", dest="dtimezero", default=0.1, type="float",
                 help="time offset (timezero t0) Monte Carlo move width")
    p.add_option("--rad", dest="lmax", default=-1, type="int", metavar="LMAX",
                 help="Solve radial diffusion equation with LMAX Bessel functions")
    p.add_option("--reduction", dest="reduction", default=False, action="store_true",
                 help="delete zero rows/columns of the transition matrix before the Monte Carlo")
    p.add_option("--pull", dest="pull", default=None, type="float", metavar="PULL",
                 help="Constant external force in kBT/Angstrom (periodic profiles; general squaring engine)")
    # ---- additions ----
    p.add_option("--nchains", dest="nchains", default=1, type="int",
                 help="number of independent chains (all GPUs together)")
    p.add_option("--device", dest="device", default=None, type="int", help="CUDA device index (default: current)")
    p.add_option("--method", dest="method", default="auto", help="propagator engine: auto | eigen | squaring")
    p.add_option("--jitter", dest="jitter", default=0.0, type="float",
                 help="Gaussian spread of the starting parameters of chains 1.. (0: all start alike)")
    p.add_option("--resume", dest="resume", default=None, metavar="FILE",
                 help="continue every chain exactly from FILE = <outfile>.chains.npz of an earlier run "
                      "(same files, model, --nchains; the seed and step counters are taken from FILE)")
    return p


def main(argv=None):
    parser = build_parser()
    options, args = parser.parse_args(argv)
    print("options:")
    print(options.__dict__)
    print("args:")
    print(args)
    if len(args) < 1:
        parser.error("at least one transition matrix file is needed")
    print("=" * 20)
    import os
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        options.device = local
    from .mc import find_parameters
    find_parameters(args, options.pbc, options.model, options.dv, options.dw, options.dwrad, options.D0,
                    options.dtimezero, options.temp, options.temp_end, options.nmc, options.nmc_update,
                    options.seed, options.outfile, options.ncosF, options.ncosD, options.ncosDrad,
                    options.move_timezero, options.initfile, options.k, options.lmax, options.reduction,
                    options.pull, nchains=options.nchains, device=options.device, method=options.method,
                    jitter=options.jitter, resume=options.resume)
