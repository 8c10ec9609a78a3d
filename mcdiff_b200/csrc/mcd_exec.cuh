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
#define MCD_RESTRICT __restrict__
#else
#define MCD_HD inline
#define MCD_HOSTDEV inline
#define MCD_NOINLINE inline
#define MCD_RESTRICT __restrict__
#endif

// Optional phase clocks (-DMCD_PHASE_CLOCK, profiles/prof_phases.py): thread 0 of CTA 0 adds the cycles
// since its previous stamp to g_phase_clk[id].  Compiled out of the product library.
// All device code lives in namespace MCD_NS (default `mcd`): a second translation unit compiles the same
// chain body with another CTA size (-DMCD_THREADS=128 -DMCD_NS=mcd128, mcd_small.cu).
#ifndef MCD_NS
#define MCD_NS mcd
#endif

#if defined(MCD_PHASE_CLOCK) && defined(__CUDACC__)
namespace MCD_NS {
__device__ long long g_phase_clk[32];
__shared__ long long s_phase_clk[33];   // [32] = previous stamp; shared memory keeps a stamp at ~2 x 30 cycles
}
#endif
#if defined(MCD_PHASE_CLOCK) && defined(__CUDA_ARCH__)
#define MCD_CLK(id)                                                             \
  do {                                                                          \
    if (threadIdx.x == 0 && blockIdx.x == 0) {                                  \
      const long long t_ = clock64();                                           \
      MCD_NS::s_phase_clk[id] += t_ - MCD_NS::s_phase_clk[32];                        \
      MCD_NS::s_phase_clk[32] = clock64();                                         \
    }                                                                           \
  } while (0)
#define MCD_CLK_BEGIN()                                                         \
  do {                                                                          \
    if (threadIdx.x == 0 && blockIdx.x == 0) {                                  \
      for (int q_ = 0; q_ < 32; ++q_) MCD_NS::s_phase_clk[q_] = 0;                 \
      MCD_NS::s_phase_clk[32] = clock64();                                         \
    }                                                                           \
  } while (0)
#define MCD_CLK_END()                                                           \
  do {                                                                          \
    if (threadIdx.x == 0 && blockIdx.x == 0)                                    \
      for (int q_ = 0; q_ < 32; ++q_) MCD_NS::g_phase_clk[q_] += MCD_NS::s_phase_clk[q_]; \
  } while (0)
#else
#define MCD_CLK(id) do { } while (0)
#define MCD_CLK_BEGIN() do { } while (0)
#define MCD_CLK_END() do { } while (0)
#endif

namespace MCD_NS {

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
struct i2 {
  int x, y;
};
// four maxima (NaN-propagating) and one sum reduced over the CTA in one go (Exec::max4sum)
struct r5 {
  double m0, m1, m2, m3, sum;
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
MCD_HD i2 ld2i(const int* p) {   // 8-byte aligned pair of ints
#if defined(__CUDA_ARCH__)
  int2 t = *reinterpret_cast<const int2*>(p);
  return i2{t.x, t.y};
#else
  return i2{p[0], p[1]};
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

// 16-byte aligned quadruple of ints from shared (or any non-constant) memory
MCD_HD i4 lds4i(const int* p) {
#if defined(__CUDA_ARCH__)
  int4 t = *reinterpret_cast<const int4*>(p);
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
  // A chain may be shared by the CTAs of a thread-block cluster (blocked engine: two SMs per chain).  Every CTA runs
  // the same control flow on the same data; crank / csize say which share of the parallel work is this CTA's.
  int crank = 0, csize = 1;

  __device__ DeviceExec(double* red_scratch) : tid(threadIdx.x), red(red_scratch) {}
  static __device__ __forceinline__ int nthr() { return kThreads; }
  // Barrier over all threads of the cluster; global-memory writes before it are visible to every CTA behind it.
  __device__ __forceinline__ void cluster_sync() {
    if (csize > 1) {
      __threadfence();
      asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
  }

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
  // Four maxima and a sum with ONE pair of barriers (a reduction costs two barriers and a 5-step shuffle
  // tree whatever it reduces; the five trees interleave).  scratch: 5 * kThreads / 32 doubles.
  // The sum has the same shape as sum(): shuffle tree per warp, then warp order.
  template <class F>
  __device__ __forceinline__ r5 max4sum(double* scratch, F f) {
    r5 v = f(tid);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v.m0 = nanmax(v.m0, __shfl_xor_sync(0xffffffffu, v.m0, o));
      v.m1 = nanmax(v.m1, __shfl_xor_sync(0xffffffffu, v.m1, o));
      v.m2 = nanmax(v.m2, __shfl_xor_sync(0xffffffffu, v.m2, o));
      v.m3 = nanmax(v.m3, __shfl_xor_sync(0xffffffffu, v.m3, o));
      v.sum += __shfl_xor_sync(0xffffffffu, v.sum, o);
    }
    constexpr int nw = kThreads / 32;
    if ((tid & 31) == 0) {
      const int w = tid >> 5;
      scratch[w] = v.m0; scratch[nw + w] = v.m1; scratch[2 * nw + w] = v.m2; scratch[3 * nw + w] = v.m3;
      scratch[4 * nw + w] = v.sum;
    }
    __syncthreads();
    r5 t{scratch[0], scratch[nw], scratch[2 * nw], scratch[3 * nw], 0.0};
#pragma unroll
    for (int w = 1; w < nw; ++w) {
      t.m0 = nanmax(t.m0, scratch[w]);
      t.m1 = nanmax(t.m1, scratch[nw + w]);
      t.m2 = nanmax(t.m2, scratch[2 * nw + w]);
      t.m3 = nanmax(t.m3, scratch[3 * nw + w]);
    }
#pragma unroll
    for (int w = 0; w < nw; ++w) t.sum += scratch[4 * nw + w];
    __syncthreads();
    return t;
  }
};
#endif

#if !defined(__CUDACC__)
struct HostExec {
  double* store;  // kThreads * kMaxRegs doubles
  int crank = 0, csize = 1;   // the emulation runs a chain as ONE CTA
  explicit HostExec(double* s) : store(s) {}
  static int nthr() { return kThreads; }
  void cluster_sync() {}
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
  template <class F>
  r5 max4sum(double*, F f) {
    r5 out = f(0);
    for (int t = 1; t < kThreads; ++t) {
      const r5 x = f(t);
      out.m0 = nanmax(out.m0, x.m0); out.m1 = nanmax(out.m1, x.m1); out.m2 = nanmax(out.m2, x.m2);
      out.m3 = nanmax(out.m3, x.m3);
    }
    out.sum = sum([&](int t) { return f(t).sum; });
    return out;
  }
};
#endif

}  // namespace MCD_NS
