// CPU emulation of the CUDA chain body -- TEST INFRASTRUCTURE ONLY.
//
// Compiles mcdiff_b200/csrc/mcd_chain.cuh with g++ under the HostExec policy: the same
// source the GPU runs as one CTA is executed by looping the thread index between barriers.
// This validates tile maps, the Jacobi schedule, the Metropolis logic and the state layout
// in the GPU-less build container.  It is never part of libmcdiff_b200.so and never a
// fallback for it.
//
// Build: g++ -O2 -std=c++17 -shared -fPIC tests/emul/emul.cpp -o tests/emul/libmcd_emul.so
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/mcdiff_b200.h"
#include "../../mcdiff_b200/csrc/mcd_chain.cuh"

using namespace mcd;

namespace {
// emulator-only method code: the packed symmetric engine, which the library selects by itself (mcd_mc_run, 57 <= np <= 104)
constexpr int kEmulPacked = 4;
struct HostProblem {
  Problem P;
  std::vector<int> counts, nsym;
  std::vector<double> rvec;
};

void make_problem(const mcd_problem_desc* d, HostProblem& hp) {
  Problem& P = hp.P;
  const int n = d->n;
  set_layout(P, n);
  const int np = P.np;
  P.pbc = d->pbc ? 1 : 0; P.dim_w = d->pbc ? n : n - 1;
  P.nlag = d->nlag; P.ncosF = d->ncosF; P.ncosD = d->ncosD;
  hp.counts.assign((size_t)d->nlag * np * np, 0);
  for (int l = 0; l < d->nlag; ++l)
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) hp.counts[((size_t)l * np + i) * np + j] = d->counts[((size_t)l * n + i) * n + j];
  hp.nsym.assign((size_t)d->nlag * np * np, 0);
  hp.rvec.assign(np, 0.0);
  for (int l = 0; l < d->nlag; ++l) {
    const int* c = hp.counts.data() + (size_t)l * np * np;
    for (int i = 0; i < np; ++i)
      for (int j = 0; j < np; ++j) {
        hp.nsym[((size_t)l * np + i) * np + j] = (i == j) ? c[i * np + j] : c[i * np + j] + c[j * np + i];
        hp.rvec[j] += c[i * np + j];
        hp.rvec[i] -= c[i * np + j];
      }
  }
  P.lagtimes = d->lagtimes;
  P.counts = hp.counts.data();
  P.nsym = hp.nsym.data();
  P.rvec = hp.rvec.data();
  P.v_basis = d->v_basis;
  P.w_basis = d->w_basis;
  P.svecs = d->string_vecs;
  P.spring_k = d->spring_k;
  P.clip = d->clip;
  P.use_pull = d->use_pull ? 1 : 0;
  P.pull = d->use_pull ? d->pull : 0.0;
  P.reserved = 0;
  plan_lags(d->lagtimes, d->nlag, false, &P.plan[0]);
  plan_lags(d->lagtimes, d->nlag, true, &P.plan[1]);
}
}  // namespace

// emulator-only method code: Chain<Exec, 4>, the tensor-only build of the full-storage engine
constexpr int kEmulTensorOnly = 5;

// absorbing rates of the next emul_loglike_batch call ([B][n], as mcd_propagator_batch's sink), or null
static const double* g_sink = nullptr;

template <int kMode>
static int loglike_full(const Problem& P, int method, int B, const double* v, const double* w, const double* lagtimes,
                        double* out_ll, double* out_P, int* status, int* sweeps, std::vector<double>& mat,
                        std::vector<unsigned char>& vec, std::vector<double>& store) {
  Smem s = carve(P, mat.data(), vec.data());
  HostExec ex(store.data());
  Chain<HostExec, kMode> ch(ex, P, s);
  ch.build_tables();
  for (int b = 0; b < B; ++b) {
    for (int i = 0; i < P.np; ++i) {
      s.v[i] = i < P.n ? v[(size_t)b * P.n + i] : 0.0;
      s.w[i] = i < P.dim_w ? w[(size_t)b * P.dim_w + i] : 0.0;
    }
    for (int l = 0; l < P.nlag; ++l) s.lag[l] = lagtimes ? lagtimes[(size_t)b * P.nlag + l] : P.lagtimes[l];
    s.sc->status = 0;
    ch.sink_ = g_sink ? g_sink + (size_t)b * P.n : nullptr;
    int sw = 0;
    double* Pb = out_P ? out_P + (size_t)b * P.nlag * P.n * P.n : nullptr;
    std::vector<double> slots((size_t)kMaxSlots * P.np * P.ld);
    if (method >= MCD_METHOD_SQUARING) {   // as mcd_loglike_kernel: shifted lag times walk plan[1], anything else is flagged
      if (lagtimes && !shifted_lags_ok(P, s.lag)) s.sc->status |= ST_BAD_LAGS;
      out_ll[b] = (s.sc->status & ST_BAD_LAGS) ? nan("")
                                               : ch.evaluate_sq(s.v, s.w, s.lag, P.plan[lagtimes ? 1 : 0], slots.data(), 0,
                                                                method == MCD_METHOD_SQUARING, Pb, &sw);
    } else {
      out_ll[b] = ch.evaluate(s.v, s.w, s.lag, 1, Pb, &sw);
    }
    if (status) status[b] = (int)s.sc->status;
    if (sweeps) sweeps[b] = sw;
  }
  return 0;
}


extern "C" {

void emul_set_sink(const double* sink) { g_sink = sink; }

int emul_loglike_batch(const mcd_problem_desc* d, int method, int B, const double* v, const double* w,
                       const double* lagtimes, double* out_ll, double* out_P, int* status, int* sweeps) {
  HostProblem hp;
  make_problem(d, hp);
  const Problem& P = hp.P;
  std::vector<double> mat(smem_matrix_bytes(P.np, P.ld) / 8 + 2);
  std::vector<unsigned char> vec(smem_vectors_bytes(P.np) + 64);
  std::vector<double> store((size_t)kThreads * kMaxRegs);
  if (P.use_pull) {   // general engine (--pull)
    Smem s = carve(P, mat.data(), vec.data());
    HostExec ex(store.data());
    Chain<HostExec, 2> ch(ex, P, s);
    ch.build_tables();
    std::vector<double> slots((size_t)kMaxSlots * P.np * P.ld);
    for (int b = 0; b < B; ++b) {
      for (int i = 0; i < P.np; ++i) {
        s.v[i] = i < P.n ? v[(size_t)b * P.n + i] : 0.0;
        s.w[i] = i < P.dim_w ? w[(size_t)b * P.dim_w + i] : 0.0;
      }
      for (int l = 0; l < P.nlag; ++l) s.lag[l] = lagtimes ? lagtimes[(size_t)b * P.nlag + l] : P.lagtimes[l];
      s.sc->status = 0;
      if (lagtimes && !shifted_lags_ok(P, s.lag)) s.sc->status |= ST_BAD_LAGS;   // as mcd_loglike_gen_kernel
      int sw = 0;
      double* Pb = out_P ? out_P + (size_t)b * P.nlag * P.n * P.n : nullptr;
      out_ll[b] = (s.sc->status & ST_BAD_LAGS) ? nan("")
                                               : ch.evaluate_gen(s.v, s.w, s.lag, P.plan[lagtimes ? 1 : 0], slots.data(), 0, Pb, &sw);
      if (status) status[b] = (int)s.sc->status;
      if (sweeps) sweeps[b] = sw;
    }
    return 0;
  }
  if (P.blocked && method == MCD_METHOD_SQUARING) {   // blocked squaring engine (panels + global workspace)
    Smem s = carve(P, mat.data(), vec.data(), true);
    HostExec ex(store.data());
    Chain<HostExec, 1> ch(ex, P, s);
    ch.build_tables();
    std::vector<double> ws(blocked_ws_doubles(P, plan_slots(P)));
    for (int b = 0; b < B; ++b) {
      for (int i = 0; i < P.np; ++i) {
        s.v[i] = i < P.n ? v[(size_t)b * P.n + i] : 0.0;
        s.w[i] = i < P.dim_w ? w[(size_t)b * P.dim_w + i] : 0.0;
      }
      for (int l = 0; l < P.nlag; ++l) s.lag[l] = lagtimes ? lagtimes[(size_t)b * P.nlag + l] : P.lagtimes[l];
      s.sc->status = 0;
      ch.sink_ = g_sink ? g_sink + (size_t)b * P.n : nullptr;
      if (lagtimes && !shifted_lags_ok(P, s.lag)) s.sc->status |= ST_BAD_LAGS;   // as mcd_loglike_blk_kernel
      int sw = 0;
      double* Pb = out_P ? out_P + (size_t)b * P.nlag * P.n * P.n : nullptr;
      out_ll[b] = (s.sc->status & ST_BAD_LAGS) ? nan("")
                                               : ch.evaluate_blk(s.v, s.w, s.lag, P.plan[lagtimes ? 1 : 0], ws.data(), 0, Pb, &sw);
      if (status) status[b] = (int)s.sc->status;
      if (sweeps) sweeps[b] = sw;
    }
    return 0;
  }
  if (method == kEmulPacked) {   // packed symmetric engine (mcd_pk.cu): slim vector layout, packed work matrices
    Smem s = carve(P, mat.data(), vec.data(), true, true);
    HostExec ex(store.data());
    Chain<HostExec, 3> ch(ex, P, s);
    ch.build_tables();
    std::vector<double> slots((size_t)(kMaxSlots + 1) * pk_elems(P.np));
    for (int b = 0; b < B; ++b) {
      for (int i = 0; i < P.np; ++i) {
        s.v[i] = i < P.n ? v[(size_t)b * P.n + i] : 0.0;
        s.w[i] = i < P.dim_w ? w[(size_t)b * P.dim_w + i] : 0.0;
      }
      for (int l = 0; l < P.nlag; ++l) s.lag[l] = lagtimes ? lagtimes[(size_t)b * P.nlag + l] : P.lagtimes[l];
      s.sc->status = 0;
      if (lagtimes && !shifted_lags_ok(P, s.lag)) s.sc->status |= ST_BAD_LAGS;
      int sw = 0;
      double* Pb = out_P ? out_P + (size_t)b * P.nlag * P.n * P.n : nullptr;
      out_ll[b] = (s.sc->status & ST_BAD_LAGS) ? nan("")
                                               : ch.evaluate_pk(s.v, s.w, s.lag, P.plan[lagtimes ? 1 : 0], slots.data(), 0, Pb, &sw);
      if (status) status[b] = (int)s.sc->status;
      if (sweeps) sweeps[b] = sw;
    }
    return 0;
  }
  if (method == kEmulTensorOnly)   // the production kernel of the in-shared-memory engine (mcd_mc_run_sq_kernel, mcd_small.cu)
    return loglike_full<4>(P, MCD_METHOD_SQUARING, B, v, w, lagtimes, out_ll, out_P, status, sweeps, mat, vec, store);
  return loglike_full<0>(P, method, B, v, w, lagtimes, out_ll, out_P, status, sweeps, mat, vec, store);
}

int emul_mc_run(const mcd_problem_desc* d, const mcd_mc_params* params, const mcd_chain_io* io, int nchains,
                int nsteps) {
  HostProblem hp;
  make_problem(d, hp);
  const Problem& P = hp.P;
  McParams mp;
  ChainIO cio;
  static_assert(sizeof(mp) == sizeof(*params), "params layout");
  static_assert(sizeof(cio) == sizeof(*io), "io layout");
  std::memcpy(&mp, params, sizeof(mp));
  std::memcpy(&cio, io, sizeof(cio));
  std::vector<double> mat(smem_matrix_bytes(P.np, P.ld) / 8 + 2);
  std::vector<unsigned char> vec(smem_vectors_bytes(P.np) + 64);
  std::vector<double> store((size_t)kThreads * kMaxRegs);
  if (P.use_pull) {
    std::vector<double> slots((size_t)kMaxSlots * P.np * P.ld);
    for (int c = 0; c < nchains; ++c) {
      Smem s = carve(P, mat.data(), vec.data());
      HostExec ex(store.data());
      Chain<HostExec, 2> ch(ex, P, s);
      ch.run(mp, cio, c, nsteps, slots.data());
    }
    return 0;
  }
  if (P.blocked && mp.method == MCD_METHOD_SQUARING) {
    std::vector<double> ws(blocked_ws_doubles(P, plan_slots(P)));
    for (int c = 0; c < nchains; ++c) {
      Smem s = carve(P, mat.data(), vec.data(), true);
      HostExec ex(store.data());
      Chain<HostExec, 1> ch(ex, P, s);
      ch.run(mp, cio, c, nsteps, ws.data());
    }
    return 0;
  }
  if (mp.method == kEmulPacked) {
    mp.method = MCD_METHOD_SQUARING;
    std::vector<double> slots((size_t)(kMaxSlots + 1) * pk_elems(P.np));
    for (int c = 0; c < nchains; ++c) {
      Smem s = carve(P, mat.data(), vec.data(), true, true);
      HostExec ex(store.data());
      Chain<HostExec, 3> ch(ex, P, s);
      ch.run(mp, cio, c, nsteps, slots.data());
    }
    return 0;
  }
  if (mp.method == kEmulTensorOnly) {
    mp.method = MCD_METHOD_SQUARING;
    for (int c = 0; c < nchains; ++c) {
      Smem s = carve(P, mat.data(), vec.data());
      HostExec ex(store.data());
      Chain<HostExec, 4> ch(ex, P, s);
      std::vector<double> slots((size_t)kMaxSlots * P.np * P.ld);
      ch.run(mp, cio, c, nsteps, slots.data());
    }
    return 0;
  }
  for (int c = 0; c < nchains; ++c) {
    Smem s = carve(P, mat.data(), vec.data());
    HostExec ex(store.data());
    Chain<HostExec> ch(ex, P, s);
    std::vector<double> slots((size_t)kMaxSlots * P.np * P.ld);
    ch.run(mp, cio, c, nsteps, slots.data());
  }
  return 0;
}

// fast_log of the likelihood epilogue, for the accuracy test
void emul_fast_log(const double* x, double* y, int n) {
  Problem P;
  std::memset(&P, 0, sizeof(P));
  set_layout(P, 8); P.nlag = 1; P.dim_w = 8; P.pbc = 1;
  std::vector<double> mat(smem_matrix_bytes(P.np, P.ld) / 8 + 2);
  std::vector<unsigned char> vec(smem_vectors_bytes(P.np) + 64);
  std::vector<double> store((size_t)kThreads * kMaxRegs);
  Smem s = carve(P, mat.data(), vec.data());
  HostExec ex(store.data());
  Chain<HostExec> ch(ex, P, s);
  ch.build_tables();
  for (int i = 0; i < n; ++i) y[i] = ch.fast_log(x[i]);
}

int emul_plan(const double* lag, int L, int shifted, int* out) {
  LagPlan pl;
  plan_lags(lag, L, shifted != 0, &pl);
  out[0] = pl.ok; out[1] = pl.fast; out[2] = pl.nchain; out[3] = pl.nslots;
  return 0;
}

void emul_philox(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned* out) {
  philox4x32_10(c0, c1, c2, c3, k0, k1, out);
}

}  // extern "C"
