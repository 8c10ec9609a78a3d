// 192-thread build of the chain body with PACKED symmetric work matrices (Chain<Exec, 3>), 57 <= np <= 104.
//
// The 384-thread kernel of mcd_lib.cu keeps two full n x n matrices in shared memory (180 KB at 104 bins): one
// chain per SM, and the tensor pipe idles during every scalar phase of that chain (ncu r02: DMMA pipe 49 %).
// Packed tile storage (mcd_chain.cuh: pk_tile / pk_off) halves the matrices, so TWO chains fit in one SM: the Taylor
// passes, epilogues, Metropolis steps and barriers of one overlap the m8n8k4 loops of the other.
// 192 threads = 6 warps: the 12 warp tasks of a product at 104 bins run in two full rounds, and two CTAs x 192 threads
// leave 170 registers per thread (256 threads would cap them at 128: spills in the Taylor lanes and the k loop).
// Separate translation unit, separate namespace (mcdpk): nothing here changes the code of the other kernels.
#define MCD_THREADS 256
#define MCD_NS mcdpk
#include <cuda_runtime.h>

#include <cstring>

#include "mcd_chain.cuh"

using namespace mcdpk;

__global__ void __launch_bounds__(kThreads, 2)
mcdpk_mc_run_kernel(Problem P, McParams mp, ChainIO io, int nsteps, double* gws) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const size_t mat_bytes = smem_packed_bytes(P.np);
  Smem s = carve(P, reinterpret_cast<double*>(smem_raw), smem_raw + mat_bytes, true, true);
  DeviceExec ex(s.red);
  Chain<DeviceExec, 3> ch(ex, P, s);
  ch.run(mp, io, (long long)blockIdx.x, nsteps, gws + (size_t)blockIdx.x * (plan_slots(P) + 1) * pk_elems(P.np));
}

// C++ entry points used by mcd_lib.cu (structs passed as bytes: the translation units do not share types)
size_t mcdpk_problem_bytes() { return sizeof(Problem); }
size_t mcdpk_smem_bytes(int np) { return smem_packed_bytes(np) + smem_vectors_bytes(np, true); }
size_t mcdpk_scratch_doubles(const void* problem) {
  Problem P;
  std::memcpy(&P, problem, sizeof(P));
  return (size_t)(plan_slots(P) + 1) * pk_elems(P.np);
}

int mcdpk_launch_mc_run(const void* problem, const void* params, const void* io, int nchains, int nsteps, size_t smem,
                         double* gws, cudaStream_t st) {
  Problem P;
  McParams mp;
  ChainIO cio;
  std::memcpy(&P, problem, sizeof(P));
  std::memcpy(&mp, params, sizeof(mp));
  std::memcpy(&cio, io, sizeof(cio));
  cudaError_t e = cudaFuncSetAttribute(mcdpk_mc_run_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  mcdpk_mc_run_kernel<<<nchains, kThreads, smem, st>>>(P, mp, cio, nsteps, gws);
  return (int)cudaGetLastError();
}
