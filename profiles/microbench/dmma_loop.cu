// Microbenchmark of the m8n8k4 k loop of the squaring engine (profiles/, not part of the library).
// One CTA per SM, 12 warps, every warp owns 8 tiles (NA of them on tile row A) of a 104 x 104
// symmetric product with operands in shared memory (pitch 108 doubles), nk4 = 26 k-steps.
// Reports cycles per product and the DMMA-pipe utilisation that implies
// (8 tiles x 26 steps x 16 cycles per warp, 3 warps per SM sub-partition).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 dmma_loop.cu -o dmma_loop && ./dmma_loop
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NP = 104, LD = 108, NK4 = NP / 4, TPW = 8;

__device__ __forceinline__ double lds64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void lds128(unsigned addr, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// V0: the library's loop (loads interleaved by the compiler, unroll 2)
template <int NA>
__device__ __forceinline__ void loop_v0(unsigned ax, unsigned ay, int offA0, int offA1, const int (&offB)[TPW],
                                        double (&c)[TPW][2]) {
  const int pitch4 = 4 * LD * 8;
#pragma unroll 2
  for (int it = 0; it < NK4; ++it) {
    const double a0 = lds64(ax + offA0);
    const double a1 = (NA < TPW) ? lds64(ax + offA1) : 0.0;
#pragma unroll
    for (int t = 0; t < TPW; ++t) {
      const double b = lds64(ay + offB[t]);
      dmma(c[t][0], c[t][1], (t < NA) ? a0 : a1, b);
    }
    ax += pitch4;
    ay += pitch4;
  }
}

// V1: explicit software pipeline -- operands of step it+1 are loaded before the MMAs of step it
template <int NA>
__device__ __forceinline__ void loop_v1(unsigned ax, unsigned ay, int offA0, int offA1, const int (&offB)[TPW],
                                        double (&c)[TPW][2]) {
  const int pitch4 = 4 * LD * 8;
  double a0 = lds64(ax + offA0), a1 = lds64(ax + offA1), b[TPW];
#pragma unroll
  for (int t = 0; t < TPW; ++t) b[t] = lds64(ay + offB[t]);
#pragma unroll 1
  for (int it = 0; it < NK4; ++it) {
    ax += pitch4;
    ay += pitch4;
    double na0 = 0, na1 = 0, nb[TPW];
    const bool more = it + 1 < NK4;
    if (more) {
      na0 = lds64(ax + offA0);
      na1 = lds64(ax + offA1);
#pragma unroll
      for (int t = 0; t < TPW; ++t) nb[t] = lds64(ay + offB[t]);
    }
#pragma unroll
    for (int t = 0; t < TPW; ++t) dmma(c[t][0], c[t][1], (t < NA) ? a0 : a1, b[t]);
    a0 = na0; a1 = na1;
#pragma unroll
    for (int t = 0; t < TPW; ++t) b[t] = nb[t];
  }
}

// V2: adjacent tiles paired, one 128-bit load feeds the B fragments of both (columns interleaved)
template <int NA>
__device__ __forceinline__ void loop_v2(unsigned ax, unsigned ay2, int offA0, int offA1, const int (&offB2)[TPW / 2],
                                        double (&c)[TPW][2]) {
  const int pitch4 = 4 * LD * 8;
#pragma unroll 2
  for (int it = 0; it < NK4; ++it) {
    const double a0 = lds64(ax + offA0);
    const double a1 = (NA < TPW) ? lds64(ax + offA1) : 0.0;
#pragma unroll
    for (int t = 0; t < TPW / 2; ++t) {
      double b0, b1;
      lds128(ay2 + offB2[t], b0, b1);
      dmma(c[2 * t][0], c[2 * t][1], (2 * t < NA) ? a0 : a1, b0);
      dmma(c[2 * t + 1][0], c[2 * t + 1][1], (2 * t + 1 < NA) ? a0 : a1, b1);
    }
    ax += pitch4;
    ay2 += pitch4;
  }
}

// V3: pure issue rate, operands in registers
__device__ __forceinline__ void loop_v3(double (&c)[TPW][2], double a, double b) {
#pragma unroll 2
  for (int it = 0; it < NK4; ++it) {
#pragma unroll
    for (int t = 0; t < TPW; ++t) dmma(c[t][0], c[t][1], a, b);
  }
}

template <int V, int NA>
__global__ void __launch_bounds__(384, 1) bench(double* out, long long* cyc, int reps) {
  extern __shared__ __align__(16) double sm[];
  double* X = sm;
  double* Y = sm + NP * LD;
  for (int e = threadIdx.x; e < 2 * NP * LD; e += blockDim.x) sm[e] = 1e-3 * ((e * 7) % 13);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int fr = lane >> 2, fk = lane & 3;
  const int rowA = warp % 13, rowB = (warp + 5) % 13;
  int offB[TPW], offB2[TPW / 2];
  for (int t = 0; t < TPW; ++t) offB[t] = ((warp + t) % 13) * 64;
  for (int t = 0; t < TPW / 2; ++t) offB2[t] = ((warp + 2 * t) % 12) * 64;
  const unsigned ax = (unsigned)__cvta_generic_to_shared(X) + (unsigned)(fk * LD + fr) * 8u;
  const unsigned ay = (unsigned)__cvta_generic_to_shared(Y) + (unsigned)(fk * LD + fr) * 8u;
  const unsigned ay2 = (unsigned)__cvta_generic_to_shared(Y) + (unsigned)(fk * LD + 2 * fr) * 8u;
  double acc = 0.0;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    double c[TPW][2];
#pragma unroll
    for (int t = 0; t < TPW; ++t) { c[t][0] = 0.0; c[t][1] = 0.0; }
    if (V == 0) loop_v0<NA>(ax, ay, rowA * 64, rowB * 64, offB, c);
    if (V == 1) loop_v1<NA>(ax, ay, rowA * 64, rowB * 64, offB, c);
    if (V == 2) loop_v2<NA>(ax, ay2, rowA * 64, rowB * 64, offB2, c);
    if (V == 3) loop_v3(c, X[lane], Y[lane]);
#pragma unroll
    for (int t = 0; t < TPW; ++t) acc += c[t][0] + c[t][1];
    __syncthreads();
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int V, int NA>
void run(const char* name, double* out, long long* cyc, int nsm) {
  const int reps = 200;
  const size_t smem = 2 * NP * LD * sizeof(double);
  cudaFuncSetAttribute(bench<V, NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  bench<V, NA><<<nsm, 384, smem>>>(out, cyc, 10);
  bench<V, NA><<<nsm, 384, smem>>>(out, cyc, reps);
  cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < nsm; ++i) mean += (double)h[i];
  mean /= nsm * (double)reps;
  const double ideal = 3.0 * TPW * NK4 * 16.0;
  printf("%-44s NA=%d  %8.0f cycles/product  DMMA pipe %.1f %%  (%s)\n", name, NA, mean, 100.0 * ideal / mean,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  int nsm = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  double* out;
  long long* cyc;
  cudaMalloc(&out, (size_t)nsm * 384 * sizeof(double));
  cudaMalloc(&cyc, nsm * sizeof(long long));
  run<3, 8>("V3 registers only (issue rate)", out, cyc, nsm);
  run<0, 8>("V0 library loop", out, cyc, nsm);
  run<0, 5>("V0 library loop", out, cyc, nsm);
  run<1, 8>("V1 explicit software pipeline", out, cyc, nsm);
  run<1, 5>("V1 explicit software pipeline", out, cyc, nsm);
  run<2, 8>("V2 paired tiles, 128-bit B loads", out, cyc, nsm);
  run<2, 5>("V2 paired tiles, 128-bit B loads", out, cyc, nsm);
  return 0;
}
