// Microbenchmark of the Taylor/Horner passes of the squaring engine (n = 100, pitch 108, 384 threads, m = 12).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 horner.cu -o horner && ./horner
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 100, NP = 104, LD = 108, T = 384, M = 12;

template <int V>
__global__ void __launch_bounds__(T, 1) bench(double* out, long long* cyc, int reps) {
  extern __shared__ __align__(16) double sm[];
  double* A = sm;
  double* B = sm + NP * LD;
  double* a = B + NP * LD;
  double* b = a + NP;
  double* fac = b + NP;
  unsigned char* hrow = reinterpret_cast<unsigned char*>(fac + 32);
  unsigned char* hrem = hrow + T;
  const int tid = threadIdx.x, n = N, ld = LD;
  const int tpr = T / N;
  for (int i = tid; i < NP; i += T) { a[i] = -2.0 - 1e-3 * i; b[i] = 1.0 + 1e-3 * i; }
  hrow[tid] = (unsigned char)(tid / tpr < n ? tid / tpr : 255);
  hrem[tid] = (unsigned char)(tid % tpr);
  if (tid < 32) fac[tid] = 0.01 / (tid + 1);
  __syncthreads();
  long long tot = 0;
  for (int rep = 0; rep < reps; ++rep) {
    for (int e = tid; e < NP * LD; e += T) { A[e] = 0.0; B[e] = 0.0; }
    __syncthreads();
    for (int i = tid; i < n; i += T) A[i * ld + i] = 1.0;
    __syncthreads();
    double* src = A;
    double* dst = B;
    const long long t0 = clock64();
    for (int tt = 1; tt <= M; ++tt) {
      const int width = 2 * tt + 1;
      if (V == 0) {   // library version
        const int i = hrow[tid], r = hrem[tid];
        if (i < n) {
          const double f = fac[tt];
          const int im = (i == 0) ? n - 1 : i - 1, ip = (i == n - 1) ? 0 : i + 1;
          const double ai = f * a[i], bm = f * b[im], bp = f * b[i];
          const double* __restrict__ s0 = src + i * ld;
          const double* __restrict__ sm_ = src + im * ld;
          const double* __restrict__ sp = src + ip * ld;
          double* __restrict__ d0 = dst + i * ld;
          for (int dd0 = r; dd0 < width; dd0 += 4 * tpr) {
            int jq[4]; bool ok[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int dd = dd0 + q * tpr;
              bool o = dd < width;
              int j = i - tt + dd;
              if (j < 0) j += n; else if (j >= n) j -= n;
              ok[q] = o; jq[q] = o ? j : i;
            }
            double x0[4], xm[4], xp[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { x0[q] = s0[jq[q]]; xm[q] = sm_[jq[q]]; xp[q] = sp[jq[q]]; }
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (ok[q]) d0[jq[q]] = ((i == jq[q]) ? 1.0 : 0.0) + (ai * x0[q] + bm * xm[q] + bp * xp[q]);
          }
        }
      } else if (V == 1) {   // element-flat: e -> (row, band slot) with a constant divisor per pass
        const double f = fac[tt];
        for (int e = tid; e < n * width; e += T) {
          const int i = e / width, d = e - i * width;
          const int im = (i == 0) ? n - 1 : i - 1, ip = (i == n - 1) ? 0 : i + 1;
          int j = i - tt + d;
          if (j < 0) j += n; else if (j >= n) j -= n;
          dst[i * ld + j] = ((i == j) ? 1.0 : 0.0) + f * (a[i] * src[i * ld + j] + b[im] * src[im * ld + j] + b[i] * src[ip * ld + j]);
        }
      } else if (V == 2) {   // empty pass (barrier only)
      } else if (V == 3) {   // two Horner steps per pass: H <- I + f1 S + f1 f2 S^2 H  (pentadiagonal), every second tt
        if (tt & 1) {
          const int i = hrow[tid], r = hrem[tid];
          const int w2 = (tt + 1 <= M) ? tt + 1 : tt;      // half-width after this pass
          const int width2 = 2 * w2 + 1;
          if (i < n) {
            const double f1 = fac[tt], f2 = fac[tt + 1];
            const int im = (i == 0) ? n - 1 : i - 1, ip = (i == n - 1) ? 0 : i + 1;
            const int im2 = (im == 0) ? n - 1 : im - 1, ip2 = (ip == n - 1) ? 0 : ip + 1;
            const double ai = a[i], am = a[im], ap = a[ip], bi = b[i], bm = b[im], bm2 = b[im2], bp = b[ip];
            const double g = f1 * f2;
            const double cm2 = g * bm2 * bm, cm1 = g * bm * (am + ai), c0 = g * (ai * ai + bm * bm + bi * bi),
                         cp1 = g * bi * (ai + ap), cp2 = g * bi * bp;
            const double sa = f1 * ai, sbm = f1 * bm, sbp = f1 * bi;
            const double* __restrict__ r0 = src + i * ld;
            const double* __restrict__ r1 = src + im * ld;
            const double* __restrict__ r2 = src + ip * ld;
            const double* __restrict__ r3 = src + im2 * ld;
            const double* __restrict__ r4 = src + ip2 * ld;
            double* __restrict__ d0 = dst + i * ld;
            for (int dd0 = r; dd0 < width2; dd0 += 4 * tpr) {
              int jq[4]; bool ok[4]; double lin[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int dd = dd0 + q * tpr;
                const bool o = dd < width2;
                const int off = dd - w2;               // column offset j - i
                int j = i + off;
                if (j < 0) j += n; else if (j >= n) j -= n;
                ok[q] = o; jq[q] = o ? j : i;
                lin[q] = (off == 0) ? 1.0 + sa : ((off == -1) ? sbm : ((off == 1) ? sbp : 0.0));
              }
              double x0[4], x1[4], x2[4], x3[4], x4[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) { x0[q] = r0[jq[q]]; x1[q] = r1[jq[q]]; x2[q] = r2[jq[q]]; x3[q] = r3[jq[q]]; x4[q] = r4[jq[q]]; }
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (ok[q]) d0[jq[q]] = lin[q] + (c0 * x0[q] + cm1 * x1[q] + cp1 * x2[q] + cm2 * x3[q] + cp2 * x4[q]);
            }
          }
        }
      }
      if (V != 3 || (tt & 1)) {
        __syncthreads();
        double* tmp = src; src = dst; dst = tmp;
      }
    }
    tot += clock64() - t0;
  }
  out[blockIdx.x * T + tid] = A[tid] + B[tid];
  if (tid == 0) cyc[blockIdx.x] = tot;
}

template <int V>
void run(const char* name, double* out, long long* cyc) {
  const int reps = 200;
  const size_t smem = (2 * NP * LD + 2 * NP + 32) * sizeof(double) + 2 * T;
  cudaFuncSetAttribute(bench<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  bench<V><<<148, T, smem>>>(out, cyc, 5);
  bench<V><<<148, T, smem>>>(out, cyc, reps);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < 148; ++i) mean += (double)h[i];
  mean /= 148.0 * reps;
  printf("%-44s %8.0f cycles per Horner (m = %d) = %6.0f per pass  (%s)\n", name, mean, M, mean / M,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * T * sizeof(double));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  run<2>("barrier only", out, cyc);
  run<0>("library walk (row slots, chunks of 4)", out, cyc);
  run<1>("flat elements, e / width", out, cyc);
  run<3>("two steps per pass (pentadiagonal)", out, cyc);
  return 0;
}
