"""Where the end-to-end leg of bench.py spends its host time: cProfile of find_parameters on the headline workload
(1024 chains x 1000 MC steps, files in, result files out), second call (warm)."""
import cProfile, contextlib, io, os, pstats, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mcdiff_b200.mc import find_parameters

NCH = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
NMC = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
tmp = tempfile.mkdtemp(prefix="mcd_e2e_")
bench.headline_files(tmp)
files = [os.path.join(tmp, "transitions.nbins%d.%d.pbc.dat" % (bench.N_BINS, int(l))) for l in bench.LAGS]
sink = io.StringIO()


def call():
    with contextlib.redirect_stdout(sink):
        find_parameters(files, True, "CosinusModel", bench.DV, bench.DW, 0.5, 1.0, 0.1, 1.0, 1.0, NMC, bench.NMC_UPDATE, 1,
                        os.path.join(tmp, "e2e.dat"), bench.NCOS_F, bench.NCOS_D, 0, False, None, -1, -1, False, None,
                        nchains=NCH, device=0)
    torch.cuda.synchronize()


call()
t0 = time.perf_counter(); call(); print("warm call: %.3f s" % (time.perf_counter() - t0))
pr = cProfile.Profile()
pr.enable(); call(); pr.disable()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(35)
