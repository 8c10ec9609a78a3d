// Execution policies for the chain body.
//
// The Monte-Carlo chain body (mcd_chain.cuh) is written once as a template over an
// execution policy.  DeviceExec runs it as ONE CTA (one chain per CTA): every `par`
// section is executed by all threads of the CTA and ends in __syncthreads().  HostExec
// (tests only, see tests/emul/) runs the very same source on the CPU by looping the
// thread index, so the control flow, the tile maps and the Jacobi schedule can be
// validated without a GPU.  HostExec is never linked into the product library.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define MCD_HD __device__ __forceinline__
#define MCD_HOSTDEV __host__ __device__ inline
#define MCD_NOINLINE __device__ __noinline__
#else
#define MCD_HD inline
#define MCD_HOSTDEV inline
#define MCD_NOINLINE inline
#endif

namespace mcd {

#ifndef MCD_THREADS
#define MCD_THREADS 384
#endif
constexpr int kThreads = MCD_THREADS;   // 12 warps, 3 per SM sub-partition (512 / 6 tiles per warp measured slower)
constexpr int kMaxRegs = 16;    // per-thread doubles that must survive a barrier

struct d2 {
  double x, y;
};
struct i4 {
  int x, y, z, w;
};

// max that propagates NaN (fmax would drop it): non-finite profiles must surface as NaN
MCD_HD double nanmax(double a, double b) { return (a != a || a > b) ? a : ((b != b) ? b : (a > b ? a : b)); }

MCD_HD d2 ld2(const double* p) {
#if defined(__CUDA_ARCH__)
  double2 t = *reinterpret_cast<const double2*>(p);
  return d2{t.x, t.y};
#else
  return d2{p[0], p[1]};
#endif
}
MCD_HD void st2(double* p, double a, double b) {
#if defined(__CUDA_ARCH__)
  *reinterpret_cast<double2*>(p) = make_double2(a, b);
#else
  p[0] = a;
  p[1] = b;
#endif
}
MCD_HD i4 ld4i(const int* p) {
#if defined(__CUDA_ARCH__)
  int4 t = __ldg(reinterpret_cast<const int4*>(p));
  return i4{t.x, t.y, t.z, t.w};
#else
  return i4{p[0], p[1], p[2], p[3]};
#endif
}

#if defined(__CUDACC__)
struct DeviceExec {
  int tid;
  double* red;  // shared scratch, >= 32 doubles
  double regs_[kMaxRegs];

  __device__ DeviceExec(double* red_scratch) : tid(threadIdx.x), red(red_scratch) {}
  static __device__ __forceinline__ int nthr() { return kThreads; }

  template <class F>
  __device__ __forceinline__ void par(F f) {
    f(tid);
    __syncthreads();
  }
  // thread 0 only, followed by a barrier
  template <class F>
  __device__ __forceinline__ void single(F f) {
    if (tid == 0) f();
    __syncthreads();
  }
  __device__ __forceinline__ double* regs(int) { return regs_; }

  template <class F>
  __device__ __forceinline__ double sum(F f) {
    double v = f(tid);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += red[w];
    __syncthreads();
    return t;
  }
  template <class F>
  __device__ __forceinline__ double max(F f) {
    double v = f(tid);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = nanmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double t = red[0];
#pragma unroll
    for (int w = 1; w < kThreads / 32; ++w) t = nanmax(t, red[w]);
    __syncthreads();
    return t;
  }
};
#endif

#if !defined(__CUDACC__)
struct HostExec {
  double* store;  // kThreads * kMaxRegs doubles
  explicit HostExec(double* s) : store(s) {}
  static int nthr() { return kThreads; }
  template <class F>
  void par(F f) {
    for (int t = 0; t < kThreads; ++t) f(t);
  }
  template <class F>
  void single(F f) {
    f();
  }
  double* regs(int tid) { return store + (size_t)tid * kMaxRegs; }
  template <class F>
  double sum(F f) {
    // same shape as the device reduction: per-warp sums, then warp order
    double tot = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) {
      double v[32];
      for (int l = 0; l < 32; ++l) v[l] = f(w * 32 + l);
      for (int o = 16; o > 0; o >>= 1)
        for (int l = 0; l < 32; ++l)
          if ((l & o) == 0) {
            double s = v[l] + v[l | o];
            v[l] = s;
            v[l | o] = s;
          }
      tot += v[0];
    }
    return tot;
  }
  template <class F>
  double max(F f) {
    double m = f(0);
    for (int t = 1; t < kThreads; ++t) m = nanmax(m, f(t));
    return m;
  }
};
#endif

}  // namespace mcd
