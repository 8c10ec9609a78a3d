// Dependent-issue latencies that shape the scalar phases of the MC kernel (one warp, one CTA).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 latency.cu -o latency && ./latency
#include <cstdio>
#include <cuda_runtime.h>

__global__ void lat(double* out, long long* cyc, int n, int d) {
  double x = out[threadIdx.x], a = 1.0000001, b = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < 1024; ++i) x = fma(x, a, b);          // dependent DFMA chain
  long long t1 = clock64();
  double y = x;
  for (int i = 0; i < 1024; ++i) y = y * a;                  // dependent DMUL chain
  long long t2 = clock64();
  double z = y;
  for (int i = 0; i < 256; ++i) z = z / (a + z * 1e-30);     // dependent double divisions
  long long t3 = clock64();
  int q = n;
  for (int i = 0; i < 256; ++i) q = q / d + n;               // dependent 32-bit integer divisions (runtime divisor)
  long long t4 = clock64();
  double l = z;
  for (int i = 0; i < 256; ++i) l = log(l + 2.0);            // dependent libm log
  long long t5 = clock64();
  double e = l;
  for (int i = 0; i < 256; ++i) e = exp(e * 1e-3);           // dependent libm exp
  long long t6 = clock64();
  out[threadIdx.x] = x + y + z + q + l + e;
  if (threadIdx.x == 0) {
    cyc[0] = (t1 - t0) / 1024; cyc[1] = (t2 - t1) / 1024; cyc[2] = (t3 - t2) / 256; cyc[3] = (t4 - t3) / 256;
    cyc[4] = (t5 - t4) / 256; cyc[5] = (t6 - t5) / 256;
  }
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 32 * sizeof(double)); cudaMemset(out, 0, 32 * sizeof(double));
  cudaMalloc(&cyc, 8 * sizeof(long long));
  lat<<<1, 32>>>(out, cyc, 1000, 3);
  lat<<<1, 32>>>(out, cyc, 1000, 3);
  long long h[8];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("dependent DFMA %lld cycles, DMUL %lld, double division %lld, int division %lld, log %lld, exp %lld  (%s)\n",
         h[0], h[1], h[2], h[3], h[4], h[5], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
