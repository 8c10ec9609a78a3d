// How fast can one CTA per SM pull an 86.5 KB operand panel (169 tiles of 8 x 8 doubles, L2 resident) into shared
// memory?  Variants of the blocked engine's panel staging, all 148 SMs at once, clock64() around the copy + wait:
//   0  cp.async 16 B, one warp instruction per tile (512 contiguous bytes)          -- the product kernel today
//   1  cp.async.bulk (TMA) 64 B per tile row into the pitched panel (8 per tile)    -- same smem layout
//   2  cp.async.bulk 512 B per tile, tile-contiguous smem layout
//   3  cp.async.bulk 6656 B per tile row of 13 tiles, tile-contiguous smem layout
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o panel_copy panel_copy.cu ; run: ./panel_copy
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

constexpr int kT = 13, kTiles = kT * kT, kLds = 108, kThreads = 384;
constexpr int kPanelDoubles = 104 * kLds;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(
          smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_copy(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(b))
               : "memory");
}

template <int kVariant>
__global__ void __launch_bounds__(kThreads, 1) copy_kernel(const double* __restrict__ G, long long* cycles, double* sink, int reps) {
  extern __shared__ __align__(128) unsigned char raw[];
  double* panel = reinterpret_cast<double*>(raw);
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(raw + (size_t)kPanelDoubles * 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const double* src0 = G + (size_t)blockIdx.x * kTiles * 64;
  if (tid == 0) mbar_init(mbar, 1);
  __syncthreads();
  long long total = 0;
  double acc = 0.0;
  for (int rep = 0; rep < reps; ++rep) {
    __syncthreads();
    const long long t0 = clock64();
    if (kVariant == 0) {
      const int r = lane >> 2, cs = lane & 3;
      for (int tile = warp; tile < kTiles; tile += kThreads / 32) {
        const int ta = tile / kT, tb = tile - ta * kT;
        const double* s = src0 + (size_t)tile * 64 + r * 8 + 2 * cs;
        double* d = panel + (size_t)(ta * 8 + r) * kLds + tb * 8 + 2 * cs;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d)), "l"(s) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
    } else {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (tid == 0) mbar_expect(mbar, kTiles * 512);
      if (kVariant == 1) {
        for (int e = tid; e < kTiles * 8; e += kThreads) {
          const int tile = e >> 3, r = e & 7, ta = tile / kT, tb = tile - ta * kT;
          bulk_copy(panel + (size_t)(ta * 8 + r) * kLds + tb * 8, src0 + (size_t)tile * 64 + r * 8, 64, mbar);
        }
      } else if (kVariant == 2) {
        for (int tile = tid; tile < kTiles; tile += kThreads) bulk_copy(panel + (size_t)tile * 64, src0 + (size_t)tile * 64, 512, mbar);
      } else {
        if (tid < kT) bulk_copy(panel + (size_t)tid * kT * 64, src0 + (size_t)tid * kT * 64, kT * 512, mbar);
      }
      mbar_wait(mbar, rep & 1);
    }
    const long long t1 = clock64();
    total += t1 - t0;
    acc += panel[(tid * 7) % (kTiles * 64)];
  }
  if (tid == 0) cycles[blockIdx.x] = total / reps;
  sink[blockIdx.x * kThreads + tid] = acc;
}

template <int kVariant>
void run(const double* G, long long* cyc, double* sink, int nsm, const char* what) {
  const size_t smem = (size_t)kPanelDoubles * 8 + 64;
  cudaFuncSetAttribute(copy_kernel<kVariant>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  copy_kernel<kVariant><<<nsm, kThreads, smem>>>(G, cyc, sink, 4);
  copy_kernel<kVariant><<<nsm, kThreads, smem>>>(G, cyc, sink, 50);
  cudaError_t e = cudaDeviceSynchronize();
  long long* h = new long long[nsm];
  cudaMemcpy(h, cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0, sum = 0;
  for (int i = 0; i < nsm; ++i) { sum += h[i]; if (h[i] > mx) mx = h[i]; }
  printf("variant %d (%s): %s  mean %lld cycles per panel (max %lld) = %.1f B/clk/SM\n", kVariant, what, cudaGetErrorString(e),
         sum / nsm, mx, kTiles * 512.0 / (double)(sum / nsm));
  delete[] h;
}

int main() {
  int nsm = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  double* G; long long* cyc; double* sink;
  cudaMalloc(&G, (size_t)nsm * kTiles * 64 * 8);
  cudaMemset(G, 0, (size_t)nsm * kTiles * 64 * 8);
  cudaMalloc(&cyc, nsm * sizeof(long long));
  cudaMalloc(&sink, (size_t)nsm * kThreads * 8);
  run<0>(G, cyc, sink, nsm, "cp.async 16 B, warp per tile");
  run<1>(G, cyc, sink, nsm, "bulk 64 B rows, pitched panel");
  run<2>(G, cyc, sink, nsm, "bulk 512 B tiles, packed panel");
  run<3>(G, cyc, sink, nsm, "bulk 6.5 KB tile rows, packed panel");
  return 0;
}
