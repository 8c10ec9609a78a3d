// Radial (2-D) likelihood, lib/twod.py:13-32 and 89-130 of the reference.
//
//   P_l[k] = sum_m bessels[m,k] expm(lag_l (R - diag(exp(wrad) b_m^2 / rmax^2))),  rmax = dim_rad
//   log L  = sum_l sum_{k,i,j} N[l,k,i,j] log(max(P_l[k,i,j], 1e-32))
//
// R - diag(sink_m) is symmetrised by the same Pi = diag(exp(-v)) as the 1-D generator, so every
// Bessel term has a symmetric kernel G_m(t) = exp(t S_m) >= 0.  One CTA per radial profile.
// Default route: the lmax kernels of every lag time are formed one after the other in shared
// memory by the tensor-core scaling-and-squaring engine of the 1-D path (Chain::kernels_sq, same
// LagPlan) and parked in a per-CTA global (L2 resident) workspace; the likelihood pass then takes
// the Bessel-weighted sums per (i, j) pair.  Lag sets without a LagPlan fall back to lmax
// symmetric eigenproblems (parallel Jacobi) contracted per pair.
#pragma once
#include "mcd_chain.cuh"

namespace mcd {

struct RadProblem {
  int n, np, ld, pbc, dim_w, nlag, dim_rad, lmax;
  const double* v;         // [n]
  const double* w;         // [dim_w]
  const double* lagtimes;  // [nlag]
  const double* counts;    // [nlag][dim_rad][n][n]
  const double* bessels;   // [lmax][dim_rad]
  const double* b0zeros;   // [lmax]
  double clip;
  LagPlan plan;            // squaring-engine route to the lag times (plan.ok == 0: eigen fallback)
};

constexpr int kMaxBessel = 64;

MCD_HOSTDEV size_t rad_workspace_doubles(const RadProblem& R) {
  const size_t eig = (size_t)R.lmax * R.np * R.ld + (size_t)R.lmax * R.np;
  const size_t sq = ((size_t)R.nlag * R.lmax + kMaxSlots) * R.np * R.ld;
  return eig > sq ? eig : sq;
}
MCD_HOSTDEV size_t rad_smem_bytes(const RadProblem& R) {
  return smem_matrix_bytes(R.np, R.ld) + smem_vectors_bytes(R.np) + ((size_t)R.np + (size_t)R.lmax * R.np) * 8 +
         (((size_t)R.n * (R.n + 1) + 15) & ~(size_t)15);   // + pair table
}

// Monte-Carlo schedule and per-chain state of the radial loop (lib/mc.py:29-30,
// lib/MCState.py:306-340, 383-386): only wrad is sampled, v and w are frozen.
struct RadMcParams {
  double num_mc_update;
  double dtemp;
  int sample_every;
  int use_tape;
  int ncosDrad;          // 0: per-bin moves of wrad, >0: wrad = wrad_basis . wrad_coeff
  int reserved;
  unsigned long long seed;
  long long chain_offset;
  long long tape_stride;
  long long nsamples_cap;
};

enum { RS_LL = 0, RS_DWRAD = 1, RS_TEMP = 2, RS_NSCAL = 8 };
enum { RC_NACC = 0, RC_NACC_UPD = 1, RC_STEP = 2, RC_STATUS = 3, RC_NCOUNT = 8 };

struct RadChainIO {
  double* wrad;           // [C][n]
  double* wcoef;          // [C][ncosDrad]
  double* scal;           // [C][RS_NSCAL]  {log_like, dwrad, temp}
  long long* counters;    // [C][RC_NCOUNT] {naccwrad, naccwrad_update, step, status}
  long long* nacc_coeff;  // [C][max(ncosDrad,1)]
  const double* wrad_basis;  // [n][ncosDrad]
  const double* tape_u;   // [steps][5]: column 1 = delta uniform, column 2 = accept uniform
  const int* tape_idx;    // [steps]
  unsigned char* decisions;  // [C][nsteps]
  double* ll_try;            // [C][nsteps]
  double* samples;           // [C][cap][3 + (ncosDrad ? ncosDrad : n)]  {log_like, dwrad, temp, wrad|coeff...}
};

#if defined(__CUDACC__)
// Shared set-up + likelihood of one radial profile for the CTA.
struct RadCtx {
  const RadProblem& R;
  DeviceExec& ex;
  Chain<DeviceExec>& ch;
  Smem& s;
  double* a0;      // [np] diagonal of the 1-D generator
  double* elm;     // [lmax][np]
  double* ws_x;    // squaring: [lmax][nlag][np][ld] kernels; eigen: [lmax][np][ld] eigenvectors (global, L2 resident)
  double* ws_lam;  // [lmax][np]
  const unsigned char* ptab;  // (i, j) of the pairs i <= j, row-major

  // log(max(p, clip)), lib/twod.py:29; the table-based logarithm needs a positive normal argument
  __device__ __forceinline__ double clipped_log(double p) const {
    const double x = fmax(p, R.clip);
    return (x >= 1.0e-300) ? ch.fast_log(x) : log(x);
  }

  // Squaring route: kernels of all lag times per Bessel term -> workspace [nlag][lmax][np][ld]
  // (+ the plan's slots behind it), then the Bessel sums and the logarithms per pair.
  __device__ __forceinline__ double loglike_sq(const double* wrad_vec, double* Pb, int* bad_out) {
    const int n = R.n, np = R.np, ld = R.ld, lmax = R.lmax, dim_rad = R.dim_rad;
    const double rmax = (double)dim_rad;
    const size_t me = (size_t)np * ld;
    double* slots = ws_x + (size_t)R.nlag * lmax * me;
    const bool use_mma = ch.dmma_ok();
    int bad = 0;
    for (int m = 0; m < lmax && !bad; ++m) {
      const double z = R.b0zeros[m];
      ex.par([&](int tid) {
        for (int i = tid; i < np; i += kThreads)
          s.a[i] = (i < n) ? a0[i] - exp(wrad_vec[i]) * z * z / (rmax * rmax) : 0.0;  // lib/twod.py:113-117
      });
      const double gersh = ex.max([&](int tid) {
        double mx = 0.0;
        for (int i = tid; i < n; i += kThreads) {
          const int im = (i == 0) ? n - 1 : i - 1;
          const double r = fabs(s.a[i]) + s.b[im] + s.b[i];
          mx = (r > mx || r != r) ? r : mx;
        }
        return mx;
      });
      // workspace layout [lmax][nlag][np][ld]: kernels_sq writes the nlag kernels of term m contiguously
      if (!ch.kernels_sq(R.lagtimes, R.plan, slots, ws_x + (size_t)m * R.nlag * me, gersh, use_mma)) bad = 1;
    }
    if (bad) {
      if (bad_out) *bad_out = 1;
      return nan("");
    }
    // ---- Bessel sums + logarithms.  The pairs i <= j are taken in chunks of CH: the lmax kernel
    //      entries of every pair of the chunk are staged in shared memory (the work matrices are free
    //      now), then thread (p, kg) handles the radial bins k = kg, kg + KG, ... of pair p ----
    ex.par([&](int tid) { ex.regs(tid)[0] = 0.0; });
    const int npairs = n * (n + 1) / 2;
    int CH = (int)((2 * me) / (size_t)lmax);
    CH = CH >= 128 ? 128 : (CH >= 64 ? 64 : (CH >= 32 ? 32 : CH));
    const int KG = kThreads / CH;
    double* gs = s.XT;   // [lmax][CH]
    for (int l = 0; l < R.nlag; ++l) {
      for (int c0 = 0; c0 < npairs; c0 += CH) {
        ex.par([&](int tid) {
          for (int e = tid; e < CH * lmax; e += kThreads) {
            const int m = e / CH, p = e - m * CH;
            if (c0 + p < npairs) {
              const int i = ptab[2 * (c0 + p)], j = ptab[2 * (c0 + p) + 1];
              gs[e] = ws_x[((size_t)m * R.nlag + l) * me + (size_t)i * ld + j];
            }
          }
        });
        ex.par([&](int tid) {
          const int kg = tid / CH, p = tid - kg * CH;
          if (kg < KG && c0 + p < npairs) {
            const int i = ptab[2 * (c0 + p)], j = ptab[2 * (c0 + p) + 1];
            const double sij = s.ea[i] * s.eb[j], sji = s.ea[j] * s.eb[i];
            double part = 0.0;
            for (int k = kg; k < dim_rad; k += KG) {
              double pk = 0.0;
              for (int m = 0; m < lmax; ++m) pk = fma(R.bessels[(size_t)m * dim_rad + k], gs[m * CH + p], pk);  // lib/twod.py:125-126
              const size_t base = (((size_t)l * dim_rad + k) * n) * n;
              const double pij = sij * pk;
              const double nij = R.counts[base + (size_t)i * n + j];
              if (nij != 0.0) part += nij * clipped_log(pij);
              if (Pb) Pb[base + (size_t)i * n + j] = pij;
              if (i != j) {
                const double pji = sji * pk;
                const double nji = R.counts[base + (size_t)j * n + i];
                if (nji != 0.0) part += nji * clipped_log(pji);
                if (Pb) Pb[base + (size_t)j * n + i] = pji;
              }
            }
            ex.regs(tid)[0] += part;
          }
        });
      }
    }
    const double ll = ex.sum([&](int tid) { return ex.regs(tid)[0]; });
    if (bad_out) *bad_out = 0;
    return ll;
  }

  __device__ __forceinline__ double loglike(const double* wrad_vec, double* Pb, int* bad_out) {
    if (R.plan.ok) return loglike_sq(wrad_vec, Pb, bad_out);
    return loglike_eig(wrad_vec, Pb, bad_out);
  }

  __device__ __forceinline__ double loglike_eig(const double* wrad_vec, double* Pb, int* bad_out) {
    const int n = R.n, np = R.np, ld = R.ld, lmax = R.lmax, dim_rad = R.dim_rad;
    const double rmax = (double)dim_rad;
    int bad = 0;
    for (int m = 0; m < lmax; ++m) {
      const double z = R.b0zeros[m];
      ex.par([&](int tid) {
        for (int i = tid; i < np; i += kThreads)
          s.a[i] = (i < n) ? a0[i] - exp(wrad_vec[i]) * z * z / (rmax * rmax) : 0.0;  // lib/twod.py:113-117
      });
      const double scl = ex.max([&](int tid) {
        double mx = 0.0;
        for (int i = tid; i < n; i += kThreads) mx = nanmax(mx, fabs(s.a[i]));
        return mx;
      });
      ex.single([&]() { s.sc->scale = scl; });
      ch.cold_init();
      const int sw = ch.jacobi();
      if (sw < 0) bad = 1;
      ex.par([&](int tid) {
        double* dst = ws_x + (size_t)m * np * ld;
        for (int e = tid; e < np * ld; e += kThreads) dst[e] = s.XT[e];
        for (int i = tid; i < np; i += kThreads) ws_lam[(size_t)m * np + i] = s.lam[i];
      });
    }
    ex.par([&](int tid) { ex.regs(tid)[0] = 0.0; });
    for (int l = 0; l < R.nlag; ++l) {
      const double t = R.lagtimes[l];
      ex.par([&](int tid) {
        for (int e = tid; e < lmax * np; e += kThreads) elm[e] = exp(t * ws_lam[e]);
      });
      ex.par([&](int tid) {
        double part = 0.0;
        double g[kMaxBessel];
        const int npairs = n * (n + 1) / 2;
        for (int e = tid; e < npairs; e += kThreads) {
          int i = 0, rem = e;   // decode e -> (i <= j)
          while (rem >= n - i) { rem -= n - i; ++i; }
          const int j = i + rem;
          for (int m = 0; m < lmax; ++m) {
            const double* x = ws_x + (size_t)m * np * ld;
            const double* el = elm + (size_t)m * np;
            double acc = 0.0;
            for (int k = 0; k < np; ++k) acc = fma(x[k * ld + i] * x[k * ld + j], el[k], acc);
            g[m] = acc;
          }
          const double sij = s.ea[i] * s.eb[j], sji = s.ea[j] * s.eb[i];
          for (int k = 0; k < dim_rad; ++k) {
            double pk = 0.0;
            for (int m = 0; m < lmax; ++m) pk = fma(R.bessels[(size_t)m * dim_rad + k], g[m], pk);  // lib/twod.py:125-126
            const size_t base = (((size_t)l * dim_rad + k) * n) * n;
            const double pij = sij * pk;
            const double nij = R.counts[base + (size_t)i * n + j];
            if (nij != 0.0) part += nij * log(fmax(pij, R.clip));
            if (Pb) Pb[base + (size_t)i * n + j] = pij;
            if (i != j) {
              const double pji = sji * pk;
              const double nji = R.counts[base + (size_t)j * n + i];
              if (nji != 0.0) part += nji * log(fmax(pji, R.clip));
              if (Pb) Pb[base + (size_t)j * n + i] = pji;
            }
          }
        }
        ex.regs(tid)[0] += part;
      });
    }
    const double ll = ex.sum([&](int tid) { return ex.regs(tid)[0]; });
    if (bad_out) *bad_out = bad;
    return bad ? nan("") : ll;
  }
};

#define MCD_RAD_SETUP()                                                                                         \
  extern __shared__ __align__(16) unsigned char smem_raw[];                                                     \
  Problem P;                                                                                                    \
  P.n = R.n; P.np = R.np; P.ld = R.ld; P.dim_w = R.dim_w; P.pbc = R.pbc; P.nlag = R.nlag;                        \
  P.ncosF = 0; P.ncosD = 0; P.lagtimes = R.lagtimes; P.counts = nullptr; P.nsym = nullptr; P.rvec = nullptr;     \
  P.v_basis = nullptr; P.w_basis = nullptr; P.svecs = nullptr; P.spring_k = 0.0; P.clip = R.clip;                \
  P.nbk = R.np; P.lds = R.ld; P.blocked = 0;                                                                    \
  const size_t mat_bytes = smem_matrix_bytes(R.np, R.ld);                                                       \
  Smem s = carve(P, reinterpret_cast<double*>(smem_raw), smem_raw + mat_bytes);                                 \
  double* a0 = reinterpret_cast<double*>(smem_raw + mat_bytes + smem_vectors_bytes(R.np));                       \
  double* elm = a0 + R.np;                                                                                      \
  unsigned char* ptab = reinterpret_cast<unsigned char*>(elm + (size_t)R.lmax * R.np);                          \
  DeviceExec ex(s.red);                                                                                         \
  Chain<DeviceExec> ch(ex, P, s);                                                                               \
  ch.build_tables();                                                                                            \
  const int n = R.n, np = R.np;                                                                                 \
  double* ws_x = gws + (size_t)blockIdx.x * rad_workspace_doubles(R);                                           \
  double* ws_lam = ws_x + (size_t)R.lmax * np * R.ld;                                                           \
  ex.par([&](int tid) {                                                                                         \
    for (int i = tid; i < np; i += kThreads) {                                                                  \
      s.v[i] = (i < n) ? R.v[i] : 0.0;                                                                          \
      s.w[i] = (i < R.dim_w) ? R.w[i] : 0.0;                                                                    \
    }                                                                                                           \
    if (tid == 0) s.sc->status = 0;                                                                             \
  });                                                                                                           \
  ch.build_S(s.v, s.w);                                                                                         \
  ex.par([&](int tid) {                                                                                         \
    for (int i = tid; i < np; i += kThreads) a0[i] = s.a[i];                                                    \
    for (int i = tid; i < n; i += kThreads) {     /* row i of the pair table starts at i n - i (i - 1) / 2 */    \
      const int row0 = i * n - i * (i - 1) / 2;                                                                 \
      for (int j = i; j < n; ++j) { ptab[2 * (row0 + j - i)] = (unsigned char)i; ptab[2 * (row0 + j - i) + 1] = (unsigned char)j; } \
    }                                                                                                           \
  });                                                                                                           \
  RadCtx ctx{R, ex, ch, s, a0, elm, ws_x, ws_lam, ptab};

__global__ void __launch_bounds__(kThreads, 1)
mcd_rad_loglike_kernel(RadProblem R, int B, const double* __restrict__ wrad, double* __restrict__ out_ll,
                       double* __restrict__ out_P, int* __restrict__ status, double* gws) {
  MCD_RAD_SETUP();
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    ex.par([&](int tid) {
      for (int i = tid; i < np; i += kThreads) s.vt[i] = (i < n) ? wrad[(size_t)b * n + i] : 0.0;
    });
    int bad = 0;
    double* Pb = out_P ? out_P + (size_t)b * R.nlag * R.dim_rad * n * n : nullptr;
    const double ll = ctx.loglike(s.vt, Pb, &bad);
    ex.par([&](int tid) {
      if (tid == 0) {
        out_ll[b] = ll;
        if (status) status[b] = bad ? ST_JACOBI_NOCONV : 0;
      }
    });
  }
}

// Radial Monte-Carlo loop: one CTA per chain (lib/mc.py:29-30, lib/MCState.py:306-340).
__global__ void __launch_bounds__(kThreads, 1)
mcd_rad_mc_run_kernel(RadProblem R, RadMcParams mp, RadChainIO io, int nsteps, double* gws) {
  MCD_RAD_SETUP();
  const long long chain = blockIdx.x;
  const int ncos = mp.ncosDrad;
  const int sdim = 3 + (ncos > 0 ? ncos : n);
  Scalars* sc = s.sc;
  double* wcur = s.vt;   // current wrad
  double* wtry = s.wt;   // trial wrad
  double* ccur = s.wc;   // current coefficients
  double* ctry = s.ct;   // trial coefficients
  ex.par([&](int tid) {
    for (int i = tid; i < np; i += kThreads) wcur[i] = (i < n) ? io.wrad[chain * n + i] : 0.0;
    for (int c = tid; c < ncos; c += kThreads) ccur[c] = io.wcoef[chain * ncos + c];
    if (tid == 0) {
      const double* q = io.scal + chain * RS_NSCAL;
      sc->ll_cur = q[RS_LL]; sc->dw = q[RS_DWRAD]; sc->temp = q[RS_TEMP];
      const long long* c = io.counters + chain * RC_NCOUNT;
      sc->naccw = c[RC_NACC]; sc->naccw_upd = c[RC_NACC_UPD]; sc->step = c[RC_STEP]; sc->status = c[RC_STATUS];
      sc->nsamp = 0;
    }
  });
  for (int it = 0; it < nsteps; ++it) {
    ex.single([&]() {   // proposal: lib/MCState.py:308-317
      const long long imc = sc->step;
      const long long ts = (long long)it + mp.tape_stride * chain;
      const int range = ncos > 0 ? ncos : n;
      double u_delta;
      if (mp.use_tape) {
        u_delta = io.tape_u[ts * 5 + 1];
        sc->idx = io.tape_idx[ts];
      } else {
        uint32_t r0[4], r1[4];
        const uint64_t cid = (uint64_t)(chain + mp.chain_offset);
        philox4x32_10((uint32_t)imc, (uint32_t)((uint64_t)imc >> 32), (uint32_t)cid, 0u, (uint32_t)mp.seed,
                      (uint32_t)(mp.seed >> 32), r0);
        philox4x32_10((uint32_t)imc, (uint32_t)((uint64_t)imc >> 32), (uint32_t)cid, 1u, (uint32_t)mp.seed,
                      (uint32_t)(mp.seed >> 32), r1);
        u_delta = u53(r0[2], r0[3]);
        sc->idx = (int)(((uint64_t)r1[0] * (uint64_t)range) >> 32);
      }
      sc->delta = sc->dw * (u_delta - 0.5);
      sc->accepted = 0;
    });
    const int idx = sc->idx;
    const double delta = sc->delta;
    ex.par([&](int tid) {
      if (ncos <= 0) {
        for (int i = tid; i < np; i += kThreads) wtry[i] = wcur[i] + (i == idx ? delta : 0.0);
      } else {
        for (int c = tid; c < ncos; c += kThreads) ctry[c] = ccur[c] + (c == idx ? delta : 0.0);
      }
    });
    if (ncos > 0) {
      ex.par([&](int tid) {
        for (int i = tid; i < np; i += kThreads) {
          double acc = 0.0;
          if (i < n)
            for (int c = 0; c < ncos; ++c) acc += ctry[c] * io.wrad_basis[(size_t)i * ncos + c];
          wtry[i] = acc;
        }
      });
    }
    int bad = 0;
    const double ll_new = ctx.loglike(wtry, nullptr, &bad);
    ex.single([&]() {   // Metropolis: lib/MCState.py:324-339
      if (ll_new == ll_new) {
        const long long imc = sc->step;
        const long long ts = (long long)it + mp.tape_stride * chain;
        double u_acc;
        if (mp.use_tape) {
          u_acc = io.tape_u[ts * 5 + 2];
        } else {
          uint32_t r1[4];
          const uint64_t cid = (uint64_t)(chain + mp.chain_offset);
          philox4x32_10((uint32_t)imc, (uint32_t)((uint64_t)imc >> 32), (uint32_t)cid, 1u, (uint32_t)mp.seed,
                        (uint32_t)(mp.seed >> 32), r1);
          u_acc = u53(r1[2], r1[3]);
        }
        if (u_acc < exp((ll_new - sc->ll_cur) / sc->temp)) {
          sc->accepted = 1;
          sc->ll_cur = ll_new;
          sc->naccw += 1;
          sc->naccw_upd += 1;
        }
      } else {
        sc->status |= ST_NAN_TRIAL;
      }
      sc->ll_try = ll_new;
    });
    if (sc->accepted) {
      ex.par([&](int tid) {
        for (int i = tid; i < np; i += kThreads) wcur[i] = wtry[i];
        for (int c = tid; c < ncos; c += kThreads) ccur[c] = ctry[c];
        if (tid == 0 && ncos > 0) io.nacc_coeff[chain * ncos + idx] += 1;
      });
    }
    ex.par([&](int tid) {
      const long long j = sc->step + 1;
      if (tid == 0) {
        if (io.decisions) io.decisions[(size_t)chain * nsteps + it] = (unsigned char)(sc->accepted ? 1 : 0);
        if (io.ll_try) io.ll_try[(size_t)chain * nsteps + it] = sc->ll_try;
      }
      if (io.samples && mp.sample_every > 0 && (j % mp.sample_every) == 0 && sc->nsamp < mp.nsamples_cap) {
        double* row = io.samples + ((size_t)chain * mp.nsamples_cap + sc->nsamp) * sdim;
        if (tid == 0) { row[0] = sc->ll_cur; row[1] = sc->dw; row[2] = sc->temp; }
        const int nv = ncos > 0 ? ncos : n;
        const double* pv = ncos > 0 ? ccur : wcur;
        for (int i = tid; i < nv; i += kThreads) row[3 + i] = pv[i];
      }
    });
    ex.single([&]() {   // lib/MCState.py:383-386, 371-377
      const long long j = sc->step + 1;
      if (io.samples && mp.sample_every > 0 && (j % mp.sample_every) == 0 && sc->nsamp < mp.nsamples_cap)
        sc->nsamp += 1;
      if (mp.num_mc_update > 0.0 && fmod((double)j, mp.num_mc_update) == 0.0) {
        sc->dw *= exp(0.1 * ((double)sc->naccw_upd / mp.num_mc_update - 0.3));
        sc->naccw_upd = 0;
        if (mp.dtemp != 0.0) sc->temp += mp.dtemp;
      }
      sc->step = j;
    });
  }
  ex.par([&](int tid) {
    for (int i = tid; i < n; i += kThreads) io.wrad[chain * n + i] = wcur[i];
    for (int c = tid; c < ncos; c += kThreads) io.wcoef[chain * ncos + c] = ccur[c];
    if (tid == 0) {
      double* q = io.scal + chain * RS_NSCAL;
      q[RS_LL] = sc->ll_cur; q[RS_DWRAD] = sc->dw; q[RS_TEMP] = sc->temp;
      long long* c = io.counters + chain * RC_NCOUNT;
      c[RC_NACC] = sc->naccw; c[RC_NACC_UPD] = sc->naccw_upd; c[RC_STEP] = sc->step; c[RC_STATUS] = sc->status;
    }
  });
}
#endif

}  // namespace mcd
