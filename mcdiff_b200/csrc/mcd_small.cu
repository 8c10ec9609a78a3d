// 128-thread build of the chain body for small profiles (np <= 56).
//
// A symmetric product of a matrix with at most 7 x 7 tiles has at most 4 warp tasks, so the 384-thread
// kernel of mcd_lib.cu keeps 4 of its 12 warps busy and every scalar phase costs a CTA barrier.  The same
// source compiled for 4 warps needs ~70 KB of shared memory and <= 168 registers, i.e. THREE CTAs per SM:
// one chain's Horner passes / epilogues / Metropolis overlap the other chains' MMAs.
// Separate translation unit, separate namespace (mcd128): nothing here changes the code of the main kernels.
#define MCD_THREADS 128
#define MCD_NS mcd128
#include <cuda_runtime.h>

#include <cstring>

#define MCD_RAD_MIN_BLOCKS 4
#include "mcd_chain.cuh"
#include "mcd_radial.cuh"

using namespace mcd128;

__global__ void __launch_bounds__(kThreads, 3)
mcd128_mc_run_kernel(Problem P, McParams mp, ChainIO io, int nsteps, double* gws) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const size_t mat_bytes = smem_matrix_bytes(P.np, P.ld);
  Smem s = carve(P, reinterpret_cast<double*>(smem_raw), smem_raw + mat_bytes);
  DeviceExec ex(s.red);
  Chain<DeviceExec, 4> ch(ex, P, s);   // tensor-core squaring only (mcd_mc_run launches this kernel for MCD_METHOD_SQUARING)
  ch.run(mp, io, (long long)blockIdx.x, nsteps, gws + (size_t)blockIdx.x * plan_slots(P) * P.np * P.ld);
}

// C++ entry points used by mcd_lib.cu (the structs have the same layout in both namespaces; they are passed
// as bytes so that the two translation units do not share types)
size_t mcd128_problem_bytes() { return sizeof(Problem); }
size_t mcd128_smem_bytes(int np, int ld) { return smem_matrix_bytes(np, ld) + smem_vectors_bytes(np); }

int mcd128_launch_mc_run(const void* problem, const void* params, const void* io, int nchains, int nsteps, size_t smem,
                         double* gws, cudaStream_t st) {
  Problem P;
  McParams mp;
  ChainIO cio;
  std::memcpy(&P, problem, sizeof(P));
  std::memcpy(&mp, params, sizeof(mp));
  std::memcpy(&cio, io, sizeof(cio));
  cudaError_t e = cudaFuncSetAttribute(mcd128_mc_run_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  mcd128_mc_run_kernel<<<nchains, kThreads, smem, st>>>(P, mp, cio, nsteps, gws);
  return (int)cudaGetLastError();
}

// ---- radial model, small profiles (np <= 56): the same radial kernels with 4 warps per chain and three chains per
// SM.  A radial product keeps 2-4 warps busy whatever the CTA size, so the 384-thread kernel of mcd_lib.cu spends most
// of a step in barrier-closed scalar phases with one chain per SM (ncu r02: DMMA pipe 18 %); three independent
// 128-thread CTAs overlap those phases with each other's MMAs.
size_t mcd128_rad_problem_bytes() { return sizeof(RadProblem); }

// Picks the lockstep batch for a per-CTA shared-memory budget; returns the dynamic shared memory, 0 if it does not fit.
size_t mcd128_rad_configure(void* rproblem, size_t smem_cap) {
  RadProblem R;
  std::memcpy(&R, rproblem, sizeof(R));
  int nb = rad_pick_batch(R, smem_cap);
  if (nb > kThreads / 32) nb = kThreads / 32;
  if (rad_smem_bytes(R, nb) > smem_cap) return 0;
  R.nbatch = nb;
  std::memcpy(rproblem, &R, sizeof(R));
  return rad_smem_bytes(R, nb);
}

int mcd128_launch_rad_mc_run(const void* rproblem, const void* params, const void* io, int nchains, int nsteps, size_t smem,
                             double* gws, cudaStream_t st) {
  RadProblem R;
  RadMcParams mp;
  RadChainIO cio;
  std::memcpy(&R, rproblem, sizeof(R));
  std::memcpy(&mp, params, sizeof(mp));
  std::memcpy(&cio, io, sizeof(cio));
  cudaError_t e = cudaFuncSetAttribute(mcd_rad_mc_run_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  mcd_rad_mc_run_kernel<<<nchains, kThreads, smem, st>>>(R, mp, cio, nsteps, gws);
  return (int)cudaGetLastError();
}

int mcd128_launch_rad_loglike(const void* rproblem, int grid, int B, const double* wrad, double* out_ll, double* out_P,
                              int* status, size_t smem, double* gws, cudaStream_t st) {
  RadProblem R;
  std::memcpy(&R, rproblem, sizeof(R));
  cudaError_t e = cudaFuncSetAttribute(mcd_rad_loglike_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  mcd_rad_loglike_kernel<<<grid, kThreads, smem, st>>>(R, B, wrad, out_ll, out_P, status, gws);
  return (int)cudaGetLastError();
}

#if defined(MCD_PHASE_CLOCK)
// debug build only: the phase clocks of this translation unit's kernels (read and reset), added to mcd_debug_phase_clocks
int mcd128_debug_phase_clocks(long long* out32) {
  long long z[32] = {0};
  cudaMemcpyFromSymbol(out32, mcd128::g_phase_clk, sizeof(z));
  cudaMemcpyToSymbol(mcd128::g_phase_clk, z, sizeof(z));
  return (int)cudaGetLastError();
}
#endif
