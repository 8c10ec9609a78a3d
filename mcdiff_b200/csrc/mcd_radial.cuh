// Radial (2-D) likelihood, lib/twod.py:13-32 and 89-130 of the reference.
//
//   P_l[k] = sum_m bessels[m,k] expm(lag_l (R - diag(exp(wrad) b_m^2 / rmax^2))),  rmax = dim_rad
//   log L  = sum_l sum_{k,i,j} N[l,k,i,j] log(max(P_l[k,i,j], 1e-32))
//
// R - diag(sink_m) is symmetrised by the same Pi = diag(exp(-v)) as the 1-D generator, so every
// Bessel term is one symmetric eigenproblem shared by all lag times.  One CTA per radial
// profile: the lmax eigensystems are solved one after the other in shared memory (same
// parallel Jacobi as the 1-D path) and parked in a per-CTA global (L2 resident) workspace;
// the likelihood pass then contracts them per (i, j) pair.
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
};

constexpr int kMaxBessel = 64;

MCD_HOSTDEV size_t rad_workspace_doubles(const RadProblem& R) {
  return (size_t)R.lmax * R.np * R.ld + (size_t)R.lmax * R.np;
}
MCD_HOSTDEV size_t rad_smem_bytes(const RadProblem& R) {
  return smem_matrix_bytes(R.np, R.ld) + smem_vectors_bytes(R.np) + ((size_t)R.np + (size_t)R.lmax * R.np) * 8;
}

#if defined(__CUDACC__)
__global__ void __launch_bounds__(kThreads, 1)
mcd_rad_loglike_kernel(RadProblem R, int B, const double* __restrict__ wrad, double* __restrict__ out_ll,
                       double* __restrict__ out_P, int* __restrict__ status, double* gws) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Problem P;
  P.n = R.n; P.np = R.np; P.ld = R.ld; P.dim_w = R.dim_w; P.pbc = R.pbc; P.nlag = R.nlag;
  P.ncosF = 0; P.ncosD = 0; P.lagtimes = R.lagtimes; P.counts = nullptr; P.nsym = nullptr; P.rvec = nullptr;
  P.v_basis = nullptr;
  P.w_basis = nullptr; P.svecs = nullptr; P.spring_k = 0.0; P.clip = R.clip;
  const size_t mat_bytes = smem_matrix_bytes(R.np, R.ld);
  Smem s = carve(P, reinterpret_cast<double*>(smem_raw), smem_raw + mat_bytes);
  double* a0 = reinterpret_cast<double*>(smem_raw + mat_bytes + smem_vectors_bytes(R.np));
  double* elm = a0 + R.np;  // [lmax][np] exp(lag * lambda)
  DeviceExec ex(s.red);
  Chain<DeviceExec> ch(ex, P, s);
  ch.build_tables();
  const int n = R.n, np = R.np, ld = R.ld, lmax = R.lmax, dim_rad = R.dim_rad;
  const double rmax = (double)dim_rad;
  double* ws_x = gws + (size_t)blockIdx.x * rad_workspace_doubles(R);
  double* ws_lam = ws_x + (size_t)lmax * np * ld;

  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    ex.par([&](int tid) {
      for (int i = tid; i < np; i += kThreads) {
        s.v[i] = (i < n) ? R.v[i] : 0.0;
        s.w[i] = (i < R.dim_w) ? R.w[i] : 0.0;
      }
      if (tid == 0) s.sc->status = 0;
    });
    ch.build_S(s.v, s.w);
    ex.par([&](int tid) {
      for (int i = tid; i < np; i += kThreads) a0[i] = s.a[i];
    });
    int bad = 0;
    for (int m = 0; m < lmax; ++m) {
      const double z = R.b0zeros[m];
      ex.par([&](int tid) {
        for (int i = tid; i < np; i += kThreads)
          s.a[i] = (i < n) ? a0[i] - exp(wrad[(size_t)b * n + i]) * z * z / (rmax * rmax) : 0.0;  // lib/twod.py:113-117
      });
      const double scl = ex.max([&](int tid) {
        double mx = 0.0;
        for (int i = tid; i < n; i += kThreads) mx = fmax(mx, fabs(s.a[i]));
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
          // decode e -> (i <= j)
          int i = 0, rem = e;
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
            if (out_P) out_P[(size_t)b * R.nlag * dim_rad * n * n + base + (size_t)i * n + j] = pij;
            if (i != j) {
              const double pji = sji * pk;
              const double nji = R.counts[base + (size_t)j * n + i];
              if (nji != 0.0) part += nji * log(fmax(pji, R.clip));
              if (out_P) out_P[(size_t)b * R.nlag * dim_rad * n * n + base + (size_t)j * n + i] = pji;
            }
          }
        }
        ex.regs(tid)[0] += part;
      });
    }
    const double ll = ex.sum([&](int tid) { return ex.regs(tid)[0]; });
    ex.par([&](int tid) {
      if (tid == 0) {
        out_ll[b] = bad ? nan("") : ll;
        if (status) status[b] = bad ? ST_JACOBI_NOCONV : 0;
      }
    });
  }
}
#endif

}  // namespace mcd
