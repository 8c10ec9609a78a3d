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

namespace MCD_NS {

struct RadProblem {
  int n, np, ld, pbc, dim_w, nlag, dim_rad, lmax;
  const double* v;         // [n]
  const double* w;         // [dim_w]
  const double* lagtimes;  // [nlag]
  const double* counts;    // [nlag][dim_rad][n][n]
  const double* pcounts;   // [nlag][dim_rad][rad_pairs_padded(n)][2] = {N_ij, N_ji (0 when i == j)} per pair i <= j (library scratch)
  const double* bessels;   // [lmax][dim_rad]
  const double* b0zeros;   // [lmax]
  double clip;
  int nbatch;              // Bessel terms whose work matrices fit in shared memory side by side (>= 1)
  int reserved;
  LagPlan plan;            // squaring-engine route to the lag times (plan.ok == 0: eigen fallback)
};

constexpr int kMaxBessel = 64;
constexpr int kMaxRadBatch = 6;   // Bessel terms advanced in lockstep by the squaring route

// Squaring route: the kernels G_m(t_l) wait for the Bessel sums in a per-CTA global (L2) workspace.  Only the pairs
// i <= j are kept, pair-major [nlag][lmax][rad_kernel_stride(n)]: 3.7 KB per kernel at n = 30 instead of the 9.2 KB of a
// padded matrix, so that the workspaces of all resident CTAs (592 x 150 KB at config 5) stay inside L2 (the full
// matrices did not: ncu r02d, 7.4 GB of DRAM traffic per launch, long_scoreboard the top stall), and the staging
// loads of the Bessel sums are contiguous instead of one sector per element.
MCD_HOSTDEV size_t rad_kernel_stride(int n) { return (((size_t)n * (n + 1) / 2) + 3) & ~(size_t)3; }
MCD_HOSTDEV size_t rad_workspace_doubles(const RadProblem& R) {
  const size_t eig = (size_t)R.lmax * R.np * R.ld + (size_t)R.lmax * R.np;
  const size_t sq = (size_t)R.nlag * R.lmax * rad_kernel_stride(R.n) + (size_t)kMaxSlots * kMaxRadBatch * R.np * R.ld;
  return eig > sq ? eig : sq;
}
// shared memory: nbatch pairs of work matrices, the chain vectors, a0[np], elm[lmax][np] (only the eigen
// fallback needs it: lag sets without a LagPlan), exp(wrad)[np], the batch diagonals [nbatch][np], the pair table
MCD_HOSTDEV size_t rad_elm_doubles(const RadProblem& R) { return R.plan.ok ? 0 : (size_t)R.lmax * R.np; }
MCD_HOSTDEV size_t rad_smem_bytes(const RadProblem& R, int nbatch) {
  return (size_t)nbatch * smem_matrix_bytes(R.np, R.ld) + smem_vectors_bytes(R.np) +
         ((size_t)2 * R.np + rad_elm_doubles(R) + (size_t)nbatch * R.np) * 8 +
         (((size_t)R.n * (R.n + 1) + 15) & ~(size_t)15);
}
// pairs i <= j, padded so that the two-pair (32-byte) loads of the Bessel-sum epilogue stay in bounds
MCD_HOSTDEV size_t rad_pairs_padded(int n) { return (((size_t)n * (n + 1) / 2 + 1) & ~(size_t)1) + 2; }
MCD_HOSTDEV size_t rad_pcounts_doubles(const RadProblem& R) {
  return (size_t)R.nlag * R.dim_rad * rad_pairs_padded(R.n) * 2;
}
MCD_HOSTDEV int rad_pick_batch(const RadProblem& R, size_t smem_cap) {
  int nb = 1;
  while (nb < kMaxRadBatch && nb < R.lmax && rad_smem_bytes(R, nb + 1) <= smem_cap) ++nb;
  return nb;
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
  double* ws_x;    // squaring: [nlag][lmax][rad_kernel_stride] packed kernels + slots; eigen: [lmax][np][ld] eigenvectors (global, L2 resident)
  double* ws_lam;  // [lmax][np]
  const unsigned char* ptab;  // (i, j) of the pairs i <= j, row-major
  double* ewr;     // [np] exp(wrad)
  double* ab;      // [nbatch][np] diagonals of the Bessel terms of the current batch

  // log(max(p, clip)), lib/twod.py:29; the table-based logarithm needs a positive normal argument
  __device__ __forceinline__ double clipped_log(double p) const {
    const double x = fmax(p, R.clip);
    return (x >= 1.0e-300) ? ch.fast_log(x) : log(x);
  }

  // ---- batched squaring route ------------------------------------------------------------
  // The work matrices of a radial problem are small (n ~ 30-60: one product keeps only 2-4 of the 12
  // warps busy and every phase costs a barrier), so NB Bessel terms advance in LOCKSTEP: their work
  // matrices sit side by side in two shared-memory banks, a Horner pass / a product handles all of
  // them at once (warp w -> term w / T1, warp task w % T1), and the squaring count is the batch
  // maximum (a smaller step h only makes the Taylor part more accurate).  S_m differs from S only in
  // its diagonal (ab).  Workspace layout [nlag][lmax][np][ld], the plan's slots (NB matrices each) behind.

  // D_b <- X_b Y_b for the nb terms of the batch (banks X, Y, D: nb matrices each, stride me)
  __device__ __forceinline__ void batch_product(const double* X, const double* Y, double* D, int nb, int T1, int store) {
    const int np = R.np, ld = R.ld;
    const size_t me = (size_t)np * ld;
    using CH = Chain<DeviceExec>;
    ex.par([&](int tid) {
      const int warp = tid >> 5, lane = tid & 31;
      const int b = warp / T1, tsk = warp - b * T1;
      const bool active = b < nb;
      const CH::WarpTask task = ch.load_task(tsk, active ? T1 : 0);
      double c[CH::kTilesPerWarp][2];
#pragma unroll
      for (int t = 0; t < CH::kTilesPerWarp; ++t) { c[t][0] = 0.0; c[t][1] = 0.0; }
      const int fr = lane >> 2, fk = lane & 3;
#if defined(__CUDA_ARCH__)
      if (active) {
        int offB[CH::kTilesPerWarp];
#pragma unroll
        for (int t = 0; t < CH::kTilesPerWarp; ++t) offB[t] = task.jb[t] * 64;
        const unsigned ax = (unsigned)__cvta_generic_to_shared(X + b * me) + (unsigned)(fk * ld + fr) * 8u;
        const unsigned ay = (unsigned)__cvta_generic_to_shared(Y + b * me) + (unsigned)(fk * ld + fr) * 8u;
        const int pitch4 = 4 * ld * 8, nk4 = np / 4;
        switch (task.na) {
          case 1: CH::mma_loop<1>(ax, ay, pitch4, nk4, task.rowA * 64, task.rowB * 64, offB, c); break;
          case 2: CH::mma_loop<2>(ax, ay, pitch4, nk4, task.rowA * 64, task.rowB * 64, offB, c); break;
          case 3: CH::mma_loop<3>(ax, ay, pitch4, nk4, task.rowA * 64, task.rowB * 64, offB, c); break;
          case 4: CH::mma_loop<4>(ax, ay, pitch4, nk4, task.rowA * 64, task.rowB * 64, offB, c); break;
          case 5: CH::mma_loop<5>(ax, ay, pitch4, nk4, task.rowA * 64, task.rowB * 64, offB, c); break;
          case 6: CH::mma_loop<6>(ax, ay, pitch4, nk4, task.rowA * 64, task.rowB * 64, offB, c); break;
          case 7: CH::mma_loop<7>(ax, ay, pitch4, nk4, task.rowA * 64, task.rowB * 64, offB, c); break;
          default: CH::mma_loop<8>(ax, ay, pitch4, nk4, task.rowA * 64, task.rowB * 64, offB, c); break;
        }
      }
#endif
      double* r = ex.regs(tid);
#pragma unroll
      for (int t = 0; t < CH::kTilesPerWarp; ++t) { r[2 * t] = c[t][0]; r[2 * t + 1] = c[t][1]; }
      if (store == 1 && active) ch.dmma_store(D + b * me, task, lane, r);
    });
    if (store == 2) {
      ex.par([&](int tid) {
        const int warp = tid >> 5;
        const int b = warp / T1, tsk = warp - b * T1;
        if (b < nb) ch.dmma_store(D + b * me, ch.load_task(tsk, T1), tid & 31, ex.regs(tid));
      });
    }
  }

  __device__ __forceinline__ void batch_copy(double* dst, const double* src, int nb) {
    const int total = nb * R.np * R.ld;
    ex.par([&](int tid) {
      for (int e = tid * 2; e < total; e += 2 * kThreads) {
        const d2 x = ld2(src + e);
        st2(dst + e, x.x, x.y);
      }
    });
  }

  // kernels of the nb terms of a batch -> packed pair-major rows of the workspace (dst: row of the first term)
  __device__ __forceinline__ void batch_pack(double* dst, const double* src, int nb) {
    const int npairs = R.n * (R.n + 1) / 2, ld = R.ld;
    const size_t me = (size_t)R.np * ld, kst = rad_kernel_stride(R.n);
    ex.par([&](int tid) {
      for (int p = tid; p < npairs; p += kThreads) {
        const int off = (int)ptab[2 * p] * ld + (int)ptab[2 * p + 1];
        for (int b = 0; b < nb; ++b) dst[(size_t)b * kst + p] = src[(size_t)b * me + off];
      }
    });
  }

  // exp(t S_b) for the nb terms of the batch: Taylor/Horner + squarings.  bank0 / bank1: NB matrices
  // each.  Returns the bank holding the results (*other: the second one), nullptr for an absurd scale.
  __device__ __forceinline__ double* batch_chain(double t, double gersh, int nb, int NB, int T1, double* bank0,
                                                 double* bank1, double** other) {
    const int n = R.n, np = R.np, ld = R.ld;
    const size_t me = (size_t)np * ld;
    using CH = Chain<DeviceExec>;
    // one plan for the batch: gersh is the batch maximum, so the shift c = gersh / 2 and the step h suit every
    // term (the spectrum of S_b lies inside [-gersh, 0])
    CH::TaylorPlan tp;
    if (!ch.taylor_plan(t, gersh, 0, tp)) return nullptr;
    const int nsq = tp.nsq;
    // T_b = e^{-hc} T_m(h(S_b + cI)) for the nb terms, one thread per (term, column) -- Chain::taylor_column.
    // The coefficient arrays of all terms sit in bank1 (free until the first product) when they fit there,
    // otherwise the terms take turns with the chain's own array.
    const size_t te = CH::taylor_ext_doubles(n);
    const bool together = (size_t)nb * te <= (size_t)NB * me;
    ex.par([&](int tid) {
      for (int e = tid * 2; e < (int)(NB * me); e += 2 * kThreads) st2(bank0 + e, 0.0, 0.0);
      if (together)
        for (int b = 0; b < nb; ++b) ch.taylor_coefs(tid, bank1 + b * te, ab + b * np, s.b, s.b, tp);
    });
    if (together) {
      ex.par([&](int tid) {
        const int total = nb * n;
        for (int e0 = 0; e0 < total; e0 += kThreads / 2) {
          if (e0 + ((tid >> 5) << 4) >= total) break;
          const int e = e0 + CH::pair_col(tid);
          const bool valid = e < total;
          const int ec = valid ? e : total - 1;
          const int b = ec / n, j = ec - b * n;
          ch.taylor_pair<true>(tid, valid, j, bank1 + b * te, tp, bank0 + b * me, ld);
        }
      });
    } else {
      for (int b = 0; b < nb; ++b) {
        ex.par([&](int tid) { ch.taylor_coefs(tid, s.text, ab + b * np, s.b, s.b, tp); });
        ex.par([&](int tid) {
          for (int j0 = 0; j0 < n; j0 += kThreads / 2) {
            if (j0 + ((tid >> 5) << 4) >= n) break;
            const int j = j0 + CH::pair_col(tid);
            ch.taylor_pair<true>(tid, j < n, j < n ? j : n - 1, s.text, tp, bank0 + b * me, ld);
          }
        });
      }
    }
    MCD_CLK(25);   // Taylor polynomials of the batch (23: the squarings that follow)
    double* src = bank0;
    double* dst = bank1;
    double* G = src;
    double* O = dst;
    for (int q = 0; q < nsq; ++q) {
      batch_product(G, G, O, nb, T1, 1);
      double* tmp = G; G = O; O = tmp;
    }
    *other = O;
    return G;
  }

  // Bessel sums on the FP64 tensor cores.  For a chunk of 96 pairs (one 8-pair tile per warp)
  //   P[k][pair] = sum_m bessels[m][k] G_m[pair]          (lib/twod.py:125-126)
  // is a (dim_rad x lmax) x (lmax x 96) product: A = bessels^T (fragments straight from L1), B = the
  // kernel entries staged in shared memory (pitch 100: conflict-free fragment loads), one m8n8k4 per
  // (8 radial bins, 8 pairs, 4 terms).  A lane ends up with 2 pairs x (dim_rad / 8) bins; their count
  // loads are issued before the MMAs and the logarithms are branch-free (one per pair and bin:
  // log P_ij = log p + (v_j - v_i)/2, log P_ji = log p - (v_j - v_i)/2, clip as lib/twod.py:29).
  // Returns false (nothing done) when the shapes do not fit; the vector-pipe version below then runs.
  // radial-bin tiles (8 bins each) accumulated at a time: 3 x 2 x kBinGroup doubles per lane.  The 128-thread build has 128
  // registers per thread (four chains per SM) and spilled most of a group of 6 (ptxas: 960 bytes of stack)
  static constexpr int kBinGroup = (kThreads <= 128) ? 3 : 6;
  __device__ __forceinline__ bool bessel_sums_mma(double* Pb, int NB) {
    const int n = R.n, np = R.np, ld = R.ld, lmax = R.lmax, dim_rad = R.dim_rad;
    const size_t me = (size_t)np * ld;
    constexpr int nwarp = kThreads / 32, CHT = 8 * nwarp, CHP = CHT + 4;
    const int lpad = (lmax + 3) & ~3, mt = (dim_rad + 7) / 8;
    if ((size_t)lpad * CHP > 2 * (size_t)NB * me) return false;
    const int npairs = n * (n + 1) / 2;
    const size_t kst = rad_kernel_stride(n);
    double* gs = s.XT;   // [lpad][CHP]
    const double clip = R.clip, logclip = log(R.clip);
    for (int l = 0; l < R.nlag; ++l) {
      for (int c0 = 0; c0 < npairs; c0 += CHT) {
        ex.par([&](int tid) {   // staging: eight independent L2 loads in flight per thread
          for (int e0 = tid; e0 < CHT * lpad; e0 += 8 * kThreads) {
            double g[8];
            int dst[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int e = e0 + q * kThreads;
              const int m = e / CHT, p = e - m * CHT;
              const bool inr = e < CHT * lpad;
              const bool live = inr && m < lmax && c0 + p < npairs;
              const int pr = live ? c0 + p : 0;
              const double x = ws_x[((size_t)l * lmax + (live ? m : 0)) * kst + pr];
              g[q] = live ? x : 0.0;
              dst[q] = inr ? m * CHP + p : -1;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (dst[q] >= 0) gs[dst[q]] = g[q];
          }
        });
        ex.par([&](int tid) {
          const int warp = tid >> 5, lane = tid & 31;
          const int fr = lane >> 2, fk = lane & 3;
          // this lane's two pairs and their count rows
          int pi[2], pj[2];
          double wv[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int pr = c0 + 8 * warp + 2 * fk + e;
            const bool okp = pr < npairs;
            wv[e] = okp ? 1.0 : 0.0;
            pi[e] = okp ? ptab[2 * pr] : 0;
            pj[e] = okp ? ptab[2 * pr + 1] : 0;
          }
          double sij[2], sji[2], dh[2], wji[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            sij[e] = s.ea[pi[e]] * s.eb[pj[e]];
            sji[e] = s.ea[pj[e]] * s.eb[pi[e]];
            dh[e] = s.hv[pj[e]] - s.hv[pi[e]];
            wji[e] = (pi[e] != pj[e]) ? wv[e] : 0.0;
          }
          double part = 0.0;
          // bin tiles in groups of kBinGroup (registers: 3 x 2 x kBinGroup doubles per lane)
          for (int t0 = 0; t0 < mt; t0 += kBinGroup) {
            double nij[kBinGroup][2], nji[kBinGroup][2], c[kBinGroup][2];
            // counts of this lane's two pairs: one 32-byte piece of the pair-major copy per radial bin
            const size_t npp = rad_pairs_padded(n);
            const size_t pr0 = (size_t)(c0 + 8 * warp + 2 * fk);
            const size_t prc = (pr0 + 2 <= npp) ? pr0 : npp - 2;
#pragma unroll
            for (int t = 0; t < kBinGroup; ++t) {
              c[t][0] = 0.0; c[t][1] = 0.0;
              const int kb = 8 * (t0 + t) + fr;
              const int kc = (kb < dim_rad) ? kb : dim_rad - 1;
              const double* q = R.pcounts + (((size_t)l * dim_rad + kc) * npp + prc) * 2;
              const d2 x0 = ld2(q), x1 = ld2(q + 2);
              nij[t][0] = x0.x; nji[t][0] = x0.y; nij[t][1] = x1.x; nji[t][1] = x1.y;
            }
#if defined(__CUDA_ARCH__)
            const unsigned gb = (unsigned)__cvta_generic_to_shared(gs) + (unsigned)(fk * CHP + 8 * warp + fr) * 8u;
            for (int k4 = 0; k4 < lpad / 4; ++k4) {
              const double bfrag = Chain<DeviceExec>::lds64(gb + (unsigned)(k4 * 4 * CHP) * 8u);
              const int term = 4 * k4 + fk;
              const double* brow = R.bessels + (size_t)((term < lmax) ? term : lmax - 1) * dim_rad;
#pragma unroll
              for (int t = 0; t < kBinGroup; ++t) {
                const int kb = 8 * (t0 + t) + fr;
                const double afrag = __ldg(brow + ((kb < dim_rad) ? kb : dim_rad - 1));
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[t][0]), "+d"(c[t][1])
                             : "d"(afrag), "d"(bfrag));
              }
            }
#endif
#pragma unroll
            for (int t = 0; t < kBinGroup; ++t) {
              const double wk = (8 * (t0 + t) + fr < dim_rad) ? 1.0 : 0.0;
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const double pk = c[t][e];
                const double pij = sij[e] * pk, pji = sji[e] * pk;
                const double lp = ch.fast_log(fmax(pk, 1.0e-300));
                const double lij = (pij >= clip) ? lp + dh[e] : logclip;
                const double lji = (pji >= clip) ? lp - dh[e] : logclip;
                part += wk * (wv[e] * (nij[t][e] * lij) + wji[e] * (nji[t][e] * lji));
              }
            }
            if (Pb) {
#pragma unroll
              for (int t = 0; t < kBinGroup; ++t) {
                const int kb = 8 * (t0 + t) + fr;
                if (kb < dim_rad) {
                  const size_t base = (((size_t)l * dim_rad + kb) * n) * n;
#pragma unroll
                  for (int e = 0; e < 2; ++e)
                    if (wv[e] != 0.0) {
                      Pb[base + (size_t)pi[e] * n + pj[e]] = sij[e] * c[t][e];
                      Pb[base + (size_t)pj[e] * n + pi[e]] = sji[e] * c[t][e];
                    }
                }
              }
            }
          }
          ex.regs(tid)[0] += part;
        });
      }
    }
    return true;
  }

  // Squaring route: kernels of all lag times per Bessel term -> workspace, then the Bessel sums and the
  // logarithms per pair.
  __device__ __forceinline__ double loglike_sq(const double* wrad_vec, double* Pb, int* bad_out) {
    const int n = R.n, np = R.np, ld = R.ld, lmax = R.lmax, dim_rad = R.dim_rad;
    const double rmax = (double)dim_rad;
    const size_t me = (size_t)np * ld;
    const LagPlan& pl = R.plan;
    const int T1 = s.sc->mma_tasks;
    int NB = R.nbatch;
    if (NB * T1 > kThreads / 32) NB = (kThreads / 32) / T1;
    double* bank0 = s.XT;
    double* bank1 = s.XT + (size_t)NB * me;
    const size_t kst = rad_kernel_stride(n);
    double* slots = ws_x + (size_t)R.nlag * lmax * kst;
    const bool fast = pl.fast != 0;
    int bad = 0;
    ex.par([&](int tid) {
      for (int i = tid; i < np; i += kThreads) ewr[i] = (i < n) ? exp(wrad_vec[i]) : 0.0;
    });
    for (int m0 = 0; m0 < lmax && !bad; m0 += NB) {
      const int nb = (lmax - m0 < NB) ? lmax - m0 : NB;
      ex.par([&](int tid) {
        for (int e = tid; e < nb * np; e += kThreads) {
          const int b = e / np, i = e - b * np;
          const double z = R.b0zeros[m0 + b];
          ab[e] = (i < n) ? a0[i] - ewr[i] * z * z / (rmax * rmax) : 0.0;  // lib/twod.py:113-117
        }
      });
      const double gersh = ex.max([&](int tid) {
        double mx = 0.0;
        for (int e = tid; e < nb * n; e += kThreads) {
          const int b = e / n, i = e - b * n;
          const int im = (i == 0) ? n - 1 : i - 1;
          const double r = fabs(ab[b * np + i]) + s.b[im] + s.b[i];
          mx = (r > mx || r != r) ? r : mx;
        }
        return mx;
      });
      // ---- LagPlan walk (same as Chain::kernels_sq) on the whole batch ----
      double* Rk = nullptr;
      double* O = nullptr;
      for (int c = 0; c < pl.nchain && !bad; ++c) {
        const double t = R.lagtimes[pl.chain_a[c]] - (pl.chain_b[c] >= 0 ? R.lagtimes[pl.chain_b[c]] : 0.0);
        double* other = nullptr;
        MCD_CLK(22);
        double* G = batch_chain(t, gersh, nb, NB, T1, bank0, bank1, &other);
        MCD_CLK(23);
        if (!G) { bad = 1; break; }
        if (!fast && pl.chain_slot[c] >= 0) batch_copy(slots + (size_t)pl.chain_slot[c] * NB * me, G, nb);
        Rk = G;
        O = other;
      }
      if (bad) break;
      batch_pack(ws_x + ((size_t)pl.order[0] * lmax + m0) * kst, Rk, nb);
      double* G0 = Rk;
      int in_O = -1;
      for (int i = 1; i < R.nlag; ++i) {
        const int l = pl.order[i];
        const double* res = Rk;
        if (fast || pl.op_slot[i] >= 0) {
          if (fast) {
            batch_product((i == 1) ? G0 : O, G0, O, nb, T1, (i == 1) ? 1 : 2);
            res = O;
          } else {
            if (in_O != pl.op_slot[i]) {
              batch_copy(O, slots + (size_t)pl.op_slot[i] * NB * me, nb);
              in_O = pl.op_slot[i];
            }
            batch_product(Rk, O, Rk, nb, T1, 2);
          }
        }
        batch_pack(ws_x + ((size_t)l * lmax + m0) * kst, res, nb);
        if (!fast && pl.park[i] >= 0) batch_copy(slots + (size_t)pl.park[i] * NB * me, Rk, nb);
      }
    }
    if (bad) {
      if (bad_out) *bad_out = 1;
      return nan("");
    }
    MCD_CLK(20);
    // ---- Bessel sums + logarithms.  The pairs i <= j are taken in chunks of CH: the lmax kernel
    //      entries of every pair of the chunk are staged in shared memory (the work matrices are free
    //      now), then thread (p, kg) handles the radial bins k = kg, kg + KG, ... of pair p ----
    ex.par([&](int tid) { ex.regs(tid)[0] = 0.0; });
    const int npairs = n * (n + 1) / 2;
    if (bessel_sums_mma(Pb, NB)) {
      const double llm = ex.sum([&](int tid) { return ex.regs(tid)[0]; });
      MCD_CLK(21);
      if (bad_out) *bad_out = 0;
      return llm;
    }
    int CH = (int)((2 * me) / (size_t)lmax);
    CH = CH >= 128 ? 128 : (CH >= 64 ? 64 : (CH >= 32 ? 32 : CH));
    const int KG = kThreads / CH;
    double* gs = s.XT;   // [lmax][CH]
    for (int l = 0; l < R.nlag; ++l) {
      for (int c0 = 0; c0 < npairs; c0 += CH) {
        ex.par([&](int tid) {
          for (int e = tid; e < CH * lmax; e += kThreads) {
            const int m = e / CH, p = e - m * CH;
            if (c0 + p < npairs) gs[e] = ws_x[((size_t)l * lmax + m) * kst + c0 + p];
          }
        });
        // thread (p, kg): four radial bins at a time -- four independent Bessel sums, their count loads
        // issued up front, and branch-free logarithms (this pass is latency bound like every scalar phase).
        // P_ij = s_ij p_k and P_ji = s_ji p_k share ONE logarithm: log P_ij = log p_k + (v_j - v_i)/2,
        // log P_ji = log p_k - (v_j - v_i)/2, each replaced by log(clip) when the entry is below the clip
        // (lib/twod.py:29).
        ex.par([&](int tid) {
          const int kg = tid / CH, p = tid - kg * CH;
          if (kg < KG && c0 + p < npairs) {
            const int i = ptab[2 * (c0 + p)], j = ptab[2 * (c0 + p) + 1];
            const double sij = s.ea[i] * s.eb[j], sji = s.ea[j] * s.eb[i];
            const double dh = s.hv[j] - s.hv[i];
            const double wji = (i != j) ? 1.0 : 0.0;
            const double clip = R.clip, logclip = log(R.clip);
            const size_t oij = (size_t)i * n + j, oji = (size_t)j * n + i;
            double part = 0.0;
            for (int kb = kg; kb < dim_rad; kb += 4 * KG) {
              int kk[4];
              double wq[4], nij[4], nji[4], pk[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int k = kb + q * KG;
                wq[q] = (k < dim_rad) ? 1.0 : 0.0;
                kk[q] = (k < dim_rad) ? k : kg;
                const size_t base = (((size_t)l * dim_rad + kk[q]) * n) * n;
                nij[q] = R.counts[base + oij];
                nji[q] = R.counts[base + oji];
                pk[q] = 0.0;
              }
              for (int m = 0; m < lmax; ++m) {
                const double g = gs[m * CH + p];
                const double* brow = R.bessels + (size_t)m * dim_rad;
#pragma unroll
                for (int q = 0; q < 4; ++q) pk[q] = fma(brow[kk[q]], g, pk[q]);  // lib/twod.py:125-126
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const double pij = sij * pk[q], pji = sji * pk[q];
                const double lp = ch.fast_log(fmax(pk[q], 1.0e-300));
                const double lij = (pij >= clip) ? lp + dh : logclip;
                const double lji = (pji >= clip) ? lp - dh : logclip;
                part += wq[q] * (nij[q] * lij + wji * (nji[q] * lji));
              }
              if (Pb) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  if (wq[q] != 0.0) {
                    const size_t base = (((size_t)l * dim_rad + kk[q]) * n) * n;
                    Pb[base + oij] = sij * pk[q];
                    Pb[base + oji] = sji * pk[q];
                  }
              }
            }
            ex.regs(tid)[0] += part;
          }
        });
      }
    }
    const double ll = ex.sum([&](int tid) { return ex.regs(tid)[0]; });
    MCD_CLK(21);
    if (bad_out) *bad_out = 0;
    return ll;
  }

  __device__ __forceinline__ double loglike(const double* wrad_vec, double* Pb, int* bad_out) {
    if (R.plan.ok) {          // elm (eigen fallback) is not allocated then: never fall through
      if (!ch.dmma_ok()) {
        if (bad_out) *bad_out = 1;
        return nan("");
      }
      return loglike_sq(wrad_vec, Pb, bad_out);
    }
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

// pair-major copy of the radial counts (see RadProblem::pcounts)
__global__ void mcd_rad_pair_counts_kernel(const double* __restrict__ counts, double* __restrict__ out, int L, int dim_rad,
                                           int n) {
  const size_t npp = rad_pairs_padded(n);
  const size_t planes = (size_t)L * dim_rad;
  for (size_t pl = blockIdx.x; pl < planes; pl += gridDim.x) {
    const double* src = counts + pl * n * n;
    double* dst = out + pl * npp * 2;
    const int npairs = n * (n + 1) / 2;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
      const int i = e / n, j = e - i * n;
      if (j < i) continue;
      const int p = i * n - i * (i - 1) / 2 + (j - i);
      dst[2 * p] = src[(size_t)i * n + j];
      dst[2 * p + 1] = (i != j) ? src[(size_t)j * n + i] : 0.0;
    }
    for (size_t e = (size_t)2 * npairs + threadIdx.x; e < 2 * npp; e += blockDim.x) dst[e] = 0.0;
  }
}

#define MCD_RAD_SETUP()                                                                                         \
  extern __shared__ __align__(16) unsigned char smem_raw[];                                                     \
  Problem P;                                                                                                    \
  P.n = R.n; P.np = R.np; P.ld = R.ld; P.dim_w = R.dim_w; P.pbc = R.pbc; P.nlag = R.nlag;                        \
  P.ncosF = 0; P.ncosD = 0; P.lagtimes = R.lagtimes; P.counts = nullptr; P.nsym = nullptr; P.rvec = nullptr;     \
  P.v_basis = nullptr; P.w_basis = nullptr; P.svecs = nullptr; P.spring_k = 0.0; P.clip = R.clip;                \
  P.nbk = R.np; P.lds = R.ld; P.blocked = 0;                                                                    \
  const size_t mat_bytes = (size_t)R.nbatch * smem_matrix_bytes(R.np, R.ld);                                    \
  Smem s = carve(P, reinterpret_cast<double*>(smem_raw), smem_raw + mat_bytes);                                 \
  double* a0 = reinterpret_cast<double*>(smem_raw + mat_bytes + smem_vectors_bytes(R.np));                       \
  double* elm = a0 + R.np;                                                                                      \
  double* ewr = elm + rad_elm_doubles(R);                                                                       \
  double* ab = ewr + R.np;                                                                                      \
  unsigned char* ptab = reinterpret_cast<unsigned char*>(ab + (size_t)R.nbatch * R.np);                         \
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
  RadCtx ctx{R, ex, ch, s, a0, elm, ws_x, ws_lam, ptab, ewr, ab};

#ifndef MCD_RAD_MIN_BLOCKS
#define MCD_RAD_MIN_BLOCKS 1
#endif
__global__ void __launch_bounds__(kThreads, MCD_RAD_MIN_BLOCKS)
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
        if (status) status[b] = bad ? ST_NONFINITE : 0;
      }
    });
  }
}

// Radial Monte-Carlo loop: one CTA per chain (lib/mc.py:29-30, lib/MCState.py:306-340).
__global__ void __launch_bounds__(kThreads, MCD_RAD_MIN_BLOCKS)
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
  MCD_CLK_BEGIN();
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
    MCD_CLK(24);
  }
  MCD_CLK_END();
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

}  // namespace MCD_NS
