// C ABI + kernels of libmcdiff_b200.so (see include/mcdiff_b200.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/mcdiff_b200.h"
#include "mcd_handle.h"
#include "mcd_chain.cuh"
#include "mcd_radial.cuh"

using namespace mcd;

// mcd_small.cu: the same chain body compiled for 128-thread CTAs (3 per SM), used for np <= kSmallBins
size_t mcd128_problem_bytes();
size_t mcd128_smem_bytes(int np, int ld);
int mcd128_launch_mc_run(const void* problem, const void* params, const void* io, int nchains, int nsteps, size_t smem,
                         double* gws, cudaStream_t st);
#if defined(MCD_PHASE_CLOCK)
int mcd128_debug_phase_clocks(long long* out32);
#endif
constexpr int kSmallBins = 56;   // 7 x 7 tiles: at most 4 symmetric warp tasks
// mcd_pk.cu: 256-thread CTAs with packed symmetric work matrices, two chains per SM (kSmallBins < np <= 104)
size_t mcdpk_problem_bytes();
size_t mcdpk_smem_bytes(int np);
size_t mcdpk_scratch_doubles(const void* problem);
int mcdpk_launch_mc_run(const void* problem, const void* params, const void* io, int nchains, int nsteps, size_t smem,
                         double* gws, cudaStream_t st);
// the radial kernels in the same 128-thread build (three chains per SM)
size_t mcd128_rad_problem_bytes();
size_t mcd128_rad_configure(void* rproblem, size_t smem_cap);
int mcd128_launch_rad_mc_run(const void* rproblem, const void* params, const void* io, int nchains, int nsteps, size_t smem,
                             double* gws, cudaStream_t st);
int mcd128_launch_rad_loglike(const void* rproblem, int grid, int B, const double* wrad, double* out_ll, double* out_P,
                              int* status, size_t smem, double* gws, cudaStream_t st);

static_assert(sizeof(mcd_mc_params) == sizeof(McParams), "mcd_mc_params layout");
static_assert(sizeof(mcd_chain_io) == sizeof(ChainIO), "mcd_chain_io layout");
static_assert(MCD_NSCAL == S_NSCAL && MCD_NCOUNT == C_NCOUNT, "slot counts");
static_assert((int)MCD_ST_NONFINITE == (int)ST_NONFINITE && (int)MCD_ST_NAN_TRIAL == (int)ST_NAN_TRIAL &&
              (int)MCD_ST_BAD_LAGS == (int)ST_BAD_LAGS,
              "status bits");


// Per-call scratch (LagPlan slots, blocked-engine work matrices, eigen work matrices, radial kernels) is a
// STREAM-ORDERED allocation: cudaMallocAsync on the caller's stream before the launch, cudaFreeAsync behind it.
// Two calls on different streams therefore never share scratch, nothing synchronises, and the pool (release
// threshold = unlimited, set in mcd_create) hands the same blocks back to the next call on that stream.
struct Scratch {
  double* p = nullptr;
  cudaStream_t st = nullptr;
  int alloc(size_t bytes, cudaStream_t stream) {
    st = stream;
    if (bytes == 0) return MCD_OK;
    cudaError_t e = cudaMallocAsync((void**)&p, bytes, stream);
    if (e != cudaSuccess) { p = nullptr; return e == cudaErrorMemoryAllocation ? MCD_E_NOMEM : (int)e; }
    return MCD_OK;
  }
  ~Scratch() { if (p) cudaFreeAsync(p, st); }
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct mcd_problem {
  mcd_handle* h;
  Problem P;       // device pointers below are owned by this object
  void* blob;      // one allocation holding all private copies
  int in_smem;
  int squaring_ok;  // a LagPlan exists and the work matrices (or their panels, `blocked`) fit in shared memory
  int blocked;      // squaring engine runs blocked (kMaxSmemBins < n <= 2 kMaxSmemBins)
  size_t smem_bytes;
  size_t blk_smem_bytes;
  mutable int max_cl2_plus1;   // 1 + co-resident two-CTA clusters of the blocked MC kernel (0: not asked yet)
};

#define CUDA_TRY(x)                          \
  do {                                       \
    cudaError_t e_ = (x);                    \
    if (e_ != cudaSuccess) return (int)e_;   \
  } while (0)

// ---------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------
template <bool kSmemMat>
__global__ void __launch_bounds__(kThreads, 1)
mcd_mc_run_kernel(Problem P, McParams mp, ChainIO io, int nsteps, double* gws) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const size_t mat_bytes = smem_matrix_bytes(P.np, P.ld);
  double* mat = kSmemMat ? reinterpret_cast<double*>(smem_raw)
                         : gws + (size_t)blockIdx.x * (mat_bytes / sizeof(double));
  unsigned char* vec = kSmemMat ? smem_raw + mat_bytes : smem_raw;
  Smem s = carve(P, mat, vec);
  DeviceExec ex(s.red);
  Chain<DeviceExec> ch(ex, P, s);
  double* slots = kSmemMat ? gws + (size_t)blockIdx.x * plan_slots(P) * P.np * P.ld : nullptr;
  ch.run(mp, io, (long long)blockIdx.x, nsteps, slots);
}

// The production kernel of the in-shared-memory engine: tensor-core squaring only (Chain<Exec, 4>), i.e. the body of
// mcd_mc_run_kernel<true> without the eigen engine and the DFMA products it selects at run time -- a third less
// code in a kernel whose phases evict each other from the instruction cache.
__global__ void __launch_bounds__(kThreads, 1)
mcd_mc_run_sq_kernel(Problem P, McParams mp, ChainIO io, int nsteps, double* gws) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const size_t mat_bytes = smem_matrix_bytes(P.np, P.ld);
  Smem s = carve(P, reinterpret_cast<double*>(smem_raw), smem_raw + mat_bytes);
  DeviceExec ex(s.red);
  Chain<DeviceExec, 4> ch(ex, P, s);
  ch.run(mp, io, (long long)blockIdx.x, nsteps, gws + (size_t)blockIdx.x * plan_slots(P) * P.np * P.ld);
}

// Thread 0, inside ex.single (see shifted_lags_ok).
__device__ __forceinline__ void check_shifted_lags(const Problem& P, const double* lag, Scalars* sc) {
  if (!shifted_lags_ok(P, lag)) sc->status |= MCD_ST_BAD_LAGS;
}

template <bool kSmemMat>
__global__ void __launch_bounds__(kThreads, 1)
mcd_loglike_kernel(Problem P, int method, int B, const double* __restrict__ v, const double* __restrict__ w,
                   const double* __restrict__ lagtimes, const double* __restrict__ sink, double* __restrict__ out_ll,
                   double* __restrict__ out_P, int* __restrict__ status, double* gws) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const size_t mat_bytes = smem_matrix_bytes(P.np, P.ld);
  double* mat = kSmemMat ? reinterpret_cast<double*>(smem_raw)
                         : gws + (size_t)blockIdx.x * (mat_bytes / sizeof(double));
  double* slots = gws;   // squaring engine (kSmemMat): gws holds the parked operands of LagPlan
  unsigned char* vec = kSmemMat ? smem_raw + mat_bytes : smem_raw;
  Smem s = carve(P, mat, vec);
  DeviceExec ex(s.red);
  Chain<DeviceExec> ch(ex, P, s);
  ch.build_tables();
  const int n = P.n, np = P.np, dim_w = P.dim_w;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    ch.sink_ = sink ? sink + (size_t)b * n : nullptr;
    ex.par([&](int tid) {
      for (int i = tid; i < np; i += kThreads) {
        s.v[i] = (i < n) ? v[(size_t)b * n + i] : 0.0;
        s.w[i] = (i < dim_w) ? w[(size_t)b * dim_w + i] : 0.0;
      }
      for (int l = tid; l < P.nlag; l += kThreads)
        s.lag[l] = lagtimes ? lagtimes[(size_t)b * P.nlag + l] : P.lagtimes[l];
      if (tid == 0) s.sc->status = 0;
    });
    int sw = 0;
    double* Pb = out_P ? out_P + (size_t)b * P.nlag * n * n : nullptr;
    double ll;
    if (kSmemMat && method >= MCD_METHOD_SQUARING) {
      const LagPlan& pl = lagtimes ? P.plan[1] : P.plan[0];
      const int nsl = plan_slots(P);
      if (lagtimes) ex.single([&]() { check_shifted_lags(P, s.lag, s.sc); });
      ll = (s.sc->status & MCD_ST_BAD_LAGS)
               ? nan("")
               : ch.evaluate_sq(s.v, s.w, s.lag, pl, slots + (size_t)blockIdx.x * nsl * P.np * P.ld, 0,
                                method == MCD_METHOD_SQUARING, Pb, &sw);
    } else {
      ll = ch.evaluate(s.v, s.w, s.lag, 1, Pb, &sw);
    }
    ex.par([&](int tid) {
      if (tid == 0) {
        out_ll[b] = ll;
        if (status) status[b] = (int)s.sc->status;
      }
    });
  }
}

// Blocked squaring engine: operand panels in shared memory, matrices in the per-CTA slice of gws.
__global__ void __launch_bounds__(kThreads, 1)
mcd_mc_run_blk_kernel(Problem P, McParams mp, ChainIO io, int nsteps, double* gws) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const size_t mat_bytes = smem_matrix_bytes(P.nbk, P.lds);
  Smem s = carve(P, reinterpret_cast<double*>(smem_raw), smem_raw + mat_bytes, true);
  DeviceExec ex(s.red);
  Chain<DeviceExec, 1> ch(ex, P, s);
  ch.run(mp, io, (long long)blockIdx.x, nsteps, gws + (size_t)blockIdx.x * blocked_ws_doubles(P, plan_slots(P)));
}

// The same engine with TWO SMs per chain (2 x 2 blocks, i.e. 105 <= n <= 208): a thread-block cluster of two CTAs
// shares the chain's global workspace; CTA r owns the diagonal block C_rr and half of the off-diagonal block of every
// product, half of the Taylor columns, and both keep identical copies of the chain state (Chain::run, ex.crank).
// 74 chains are resident instead of 148: their workspaces (40 MB) sit comfortably in L2, a product is two rounds of
// warp tasks instead of four, and 512 chains are 6.9 waves of 74 instead of 3.5 waves of 148.  Opt-in (mcd_mc_run).
constexpr size_t kClusterTail = 8;   // doubles behind a chain's workspace: the partial sums the two CTAs exchange
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
mcd_mc_run_blk2_kernel(Problem P, McParams mp, ChainIO io, int nsteps, double* gws, int chain0) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const size_t mat_bytes = smem_matrix_bytes(P.nbk, P.lds);
  Smem s = carve(P, reinterpret_cast<double*>(smem_raw), smem_raw + mat_bytes, true);
  DeviceExec ex(s.red);
  ex.crank = (int)(blockIdx.x & 1);
  ex.csize = 2;
  Chain<DeviceExec, 1, true> ch(ex, P, s);
  const long long local = (long long)(blockIdx.x >> 1);   // chain0: first chain of this launch (tail of a split launch)
  ch.run(mp, io, chain0 + local, nsteps, gws + (size_t)local * (blocked_ws_doubles(P, plan_slots(P)) + kClusterTail));
}

__global__ void __launch_bounds__(kThreads, 1)
mcd_loglike_blk_kernel(Problem P, int B, const double* __restrict__ v, const double* __restrict__ w,
                       const double* __restrict__ lagtimes, const double* __restrict__ sink, double* __restrict__ out_ll,
                       double* __restrict__ out_P, int* __restrict__ status, double* gws) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const size_t mat_bytes = smem_matrix_bytes(P.nbk, P.lds);
  Smem s = carve(P, reinterpret_cast<double*>(smem_raw), smem_raw + mat_bytes, true);
  DeviceExec ex(s.red);
  Chain<DeviceExec, 1> ch(ex, P, s);
  ch.build_tables();
  const int n = P.n, np = P.np, dim_w = P.dim_w;
  const LagPlan& pl = lagtimes ? P.plan[1] : P.plan[0];
  const int nsl = plan_slots(P);
  double* ws = gws + (size_t)blockIdx.x * blocked_ws_doubles(P, nsl);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    ch.sink_ = sink ? sink + (size_t)b * n : nullptr;
    ex.par([&](int tid) {
      for (int i = tid; i < np; i += kThreads) {
        s.v[i] = (i < n) ? v[(size_t)b * n + i] : 0.0;
        s.w[i] = (i < dim_w) ? w[(size_t)b * dim_w + i] : 0.0;
      }
      for (int l = tid; l < P.nlag; l += kThreads) s.lag[l] = lagtimes ? lagtimes[(size_t)b * P.nlag + l] : P.lagtimes[l];
      if (tid == 0) s.sc->status = 0;
    });
    if (lagtimes) ex.single([&]() { check_shifted_lags(P, s.lag, s.sc); });
    int sw = 0;
    double* Pb = out_P ? out_P + (size_t)b * P.nlag * n * n : nullptr;
    const double ll = (s.sc->status & MCD_ST_BAD_LAGS) ? nan("") : ch.evaluate_blk(s.v, s.w, s.lag, pl, ws, 0, Pb, &sw);
    ex.par([&](int tid) {
      if (tid == 0) {
        out_ll[b] = ll;
        if (status) status[b] = (int)s.sc->status;
      }
    });
  }
}

// General engine (--pull): the generator is not symmetrisable, exp(tR) by squaring with full products.
__global__ void __launch_bounds__(kThreads, 1)
mcd_mc_run_gen_kernel(Problem P, McParams mp, ChainIO io, int nsteps, double* gws) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const size_t mat_bytes = smem_matrix_bytes(P.np, P.ld);
  Smem s = carve(P, reinterpret_cast<double*>(smem_raw), smem_raw + mat_bytes);
  DeviceExec ex(s.red);
  Chain<DeviceExec, 2> ch(ex, P, s);
  ch.run(mp, io, (long long)blockIdx.x, nsteps, gws + (size_t)blockIdx.x * plan_slots(P) * P.np * P.ld);
}

__global__ void __launch_bounds__(kThreads, 1)
mcd_loglike_gen_kernel(Problem P, int B, const double* __restrict__ v, const double* __restrict__ w,
                       const double* __restrict__ lagtimes, double* __restrict__ out_ll, double* __restrict__ out_P,
                       int* __restrict__ status, double* gws) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const size_t mat_bytes = smem_matrix_bytes(P.np, P.ld);
  Smem s = carve(P, reinterpret_cast<double*>(smem_raw), smem_raw + mat_bytes);
  DeviceExec ex(s.red);
  Chain<DeviceExec, 2> ch(ex, P, s);
  ch.build_tables();
  const int n = P.n, np = P.np, dim_w = P.dim_w;
  // per-profile lag times are accepted when they are the problem's lag times plus one offset (the time-offset
  // trial of lib/MCState.py:207-210): the LagPlan of shifted lag times then applies
  const LagPlan& pl = lagtimes ? P.plan[1] : P.plan[0];
  const int nslots = plan_slots(P);
  double* slots = gws + (size_t)blockIdx.x * nslots * P.np * P.ld;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    ex.par([&](int tid) {
      for (int i = tid; i < np; i += kThreads) {
        s.v[i] = (i < n) ? v[(size_t)b * n + i] : 0.0;
        s.w[i] = (i < dim_w) ? w[(size_t)b * dim_w + i] : 0.0;
      }
      for (int l = tid; l < P.nlag; l += kThreads) s.lag[l] = lagtimes ? lagtimes[(size_t)b * P.nlag + l] : P.lagtimes[l];
      if (tid == 0) s.sc->status = 0;
    });
    if (lagtimes) ex.single([&]() { check_shifted_lags(P, s.lag, s.sc); });
    int sw = 0;
    double* Pb = out_P ? out_P + (size_t)b * P.nlag * n * n : nullptr;
    const double ll = (s.sc->status & MCD_ST_BAD_LAGS) ? nan("") : ch.evaluate_gen(s.v, s.w, s.lag, pl, slots, 0, Pb, &sw);
    ex.par([&](int tid) {
      if (tid == 0) {
        out_ll[b] = ll;
        if (status) status[b] = (int)s.sc->status;
      }
    });
  }
}

// counts [L][n][n] -> padded [L][np][np]
__global__ void mcd_pack_counts_kernel(const int* __restrict__ src, int* __restrict__ dst, int L, int n, int np) {
  const size_t total = (size_t)L * np * np;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(e % np);
    const int i = (int)((e / np) % np);
    const int l = (int)(e / ((size_t)np * np));
    dst[e] = (i < n && j < n) ? src[((size_t)l * n + i) * n + j] : 0;
  }
}

// nsym[l][i][j] = N_ij + N_ji (i != j), N_ii on the diagonal   (padded layout)
__global__ void mcd_sym_counts_kernel(const int* __restrict__ cnt, int* __restrict__ nsym, int L, int np) {
  const size_t total = (size_t)L * np * np;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(e % np);
    const int i = (int)((e / np) % np);
    const size_t base = (e / ((size_t)np * np)) * (size_t)np * np;
    nsym[e] = (i == j) ? cnt[e] : cnt[base + (size_t)i * np + j] + cnt[base + (size_t)j * np + i];
  }
}

// rvec[k] = sum_l (sum_i N[l][i][k] - sum_j N[l][k][j])
__global__ void mcd_rvec_kernel(const int* __restrict__ cnt, double* __restrict__ rvec, int L, int np) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= np) return;
  long long acc = 0;
  for (int l = 0; l < L; ++l) {
    const int* c = cnt + (size_t)l * np * np;
    for (int i = 0; i < np; ++i) acc += (long long)c[(size_t)i * np + k] - (long long)c[(size_t)k * np + i];
  }
  rvec[k] = (double)acc;
}

// dense DFMA throughput probe: 8 independent chains per thread
__global__ void __launch_bounds__(256) mcd_dfma_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double m = 0.999999, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" {

int mcd_version(void) { return MCD_VERSION; }

const char* mcd_error_string(int code) {
  switch (code) {
    case MCD_OK: return "ok";
    case MCD_E_ARG: return "invalid argument";
    case MCD_E_UNSUPPORTED: return "unsupported configuration";
    case MCD_E_NOMEM: return "out of memory";
    case MCD_E_NODEVICE: return "no CUDA device";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
  }
}

int mcd_create(int device, mcd_handle** out) {
  if (!out) return MCD_E_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return MCD_E_NODEVICE;
  if (device < 0 || device >= ndev) return MCD_E_ARG;
  CUDA_TRY(cudaSetDevice(device));
  mcd_handle* h = new (std::nothrow) mcd_handle();
  if (!h) return MCD_E_NOMEM;
  h->device = device;
  h->launches = 0;
  {   // keep freed scratch blocks in the device's default pool instead of returning them to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  *out = h;
  return MCD_OK;
}

int mcd_destroy(mcd_handle* h) {
  if (!h) return MCD_OK;
  cudaSetDevice(h->device);
  delete h;
  return MCD_OK;
}

int64_t mcd_launch_count(const mcd_handle* h) { return h ? h->launches : 0; }

int mcd_problem_create(mcd_handle* h, const mcd_problem_desc* d, void* stream, mcd_problem** out) {
  if (!h || !d || !out) return MCD_E_ARG;
  if (d->n < 3 || d->n > 508 || d->nlag < 1 || d->nlag > kMaxLag) return MCD_E_ARG;
  if (d->ncosF < 0 || d->ncosF > kMaxCoef || d->ncosD < 0 || d->ncosD > kMaxCoef) return MCD_E_ARG;
  if (!d->lagtimes || !d->counts) return MCD_E_ARG;
  if ((d->ncosF > 0 && !d->v_basis) || (d->ncosD > 0 && !d->w_basis)) return MCD_E_ARG;
  if (d->spring_k > 0.0 && !d->string_vecs) return MCD_E_ARG;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int n = d->n, dim_w = d->pbc ? n : n - 1;

  mcd_problem* p = new (std::nothrow) mcd_problem();
  if (!p) return MCD_E_NOMEM;
  p->h = h;
  set_layout(p->P, n);
  const int np = p->P.np, ld = p->P.ld;
  const size_t b_lag = align256((size_t)d->nlag * 8);
  const size_t b_cnt = align256((size_t)d->nlag * np * np * 4);
  const size_t b_rv = align256((size_t)np * 8);
  const size_t b_vb = align256((size_t)n * d->ncosF * 8);
  const size_t b_wb = align256((size_t)dim_w * d->ncosD * 8);
  const size_t b_sv = d->spring_k > 0.0 ? align256((size_t)dim_w * dim_w * 8) : 0;
  unsigned char* blob = nullptr;
  // stream-ordered allocation from the device pool: cudaMalloc / cudaFree cost tens of milliseconds per problem here
  cudaError_t e = cudaMallocAsync((void**)&blob, b_lag + 2 * b_cnt + b_rv + b_vb + b_wb + b_sv + 256, st);
  if (e != cudaSuccess) { delete p; return (int)e; }
  p->blob = blob;
  unsigned char* q = blob;
  double* lag = (double*)q; q += b_lag;
  int* cnt = (int*)q; q += b_cnt;
  int* nsym = (int*)q; q += b_cnt;
  double* rvec = (double*)q; q += b_rv;
  double* vb = (double*)q; q += b_vb;
  double* wb = (double*)q; q += b_wb;
  double* sv = (double*)q;
  cudaMemcpyAsync(lag, d->lagtimes, (size_t)d->nlag * 8, cudaMemcpyDeviceToDevice, st);
  mcd_pack_counts_kernel<<<h->sm_count, 256, 0, st>>>(d->counts, cnt, d->nlag, n, np);
  mcd_sym_counts_kernel<<<h->sm_count, 256, 0, st>>>(cnt, nsym, d->nlag, np);
  mcd_rvec_kernel<<<(np + 127) / 128, 128, 0, st>>>(cnt, rvec, d->nlag, np);
  h->launches += 3;
  if (d->ncosF > 0) cudaMemcpyAsync(vb, d->v_basis, (size_t)n * d->ncosF * 8, cudaMemcpyDeviceToDevice, st);
  if (d->ncosD > 0) cudaMemcpyAsync(wb, d->w_basis, (size_t)dim_w * d->ncosD * 8, cudaMemcpyDeviceToDevice, st);
  if (d->spring_k > 0.0)
    cudaMemcpyAsync(sv, d->string_vecs, (size_t)dim_w * dim_w * 8, cudaMemcpyDeviceToDevice, st);
  e = cudaGetLastError();
  if (e != cudaSuccess) { cudaFreeAsync(blob, st); delete p; return (int)e; }

  Problem& P = p->P;
  P.dim_w = dim_w; P.pbc = d->pbc ? 1 : 0; P.nlag = d->nlag;
  P.ncosF = d->ncosF; P.ncosD = d->ncosD;
  P.lagtimes = lag; P.counts = cnt; P.nsym = nsym; P.rvec = rvec;
  P.v_basis = d->ncosF > 0 ? vb : nullptr;
  P.w_basis = d->ncosD > 0 ? wb : nullptr;
  P.svecs = d->spring_k > 0.0 ? sv : nullptr;
  P.spring_k = d->spring_k;
  P.clip = d->clip;
  P.use_pull = d->use_pull ? 1 : 0;
  P.pull = d->use_pull ? d->pull : 0.0;
  P.reserved = 0;
  const size_t full = smem_matrix_bytes(np, ld) + smem_vectors_bytes(np);
  // the warm-start tile GEMM keeps one 4x4 tile per thread
  const int ntiles = (np / 4) * (np / 4 + 1) / 2;
  p->in_smem = (full <= (size_t)h->max_smem_optin && ntiles <= kThreads) ? 1 : 0;
  {
    double hl[kMaxLag];
    e = cudaMemcpyAsync(hl, d->lagtimes, (size_t)d->nlag * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { cudaFreeAsync(blob, st); delete p; return (int)e; }
    plan_lags(hl, d->nlag, false, &P.plan[0]);
    plan_lags(hl, d->nlag, true, &P.plan[1]);
    p->blk_smem_bytes = smem_matrix_bytes(P.nbk, P.lds) + smem_vectors_bytes(np, true);
    p->blocked = (P.blocked && p->blk_smem_bytes <= (size_t)h->max_smem_optin) ? 1 : 0;
    p->squaring_ok = (p->in_smem || p->blocked) && P.plan[0].ok && P.plan[1].ok;
    // --pull: only the general squaring engine applies (periodic, work matrices in shared memory, a LagPlan)
    if (P.use_pull && !(P.pbc && p->in_smem && P.plan[0].ok && P.plan[1].ok)) {
      cudaFreeAsync(blob, st);
      delete p;
      return MCD_E_UNSUPPORTED;
    }
  }
  p->smem_bytes = p->in_smem ? full : smem_vectors_bytes(np);
  if (!p->in_smem && ntiles > kThreads) {
    // the generic (global work matrices) path loops over tiles in the likelihood but the
    // warm-start product needs one tile per thread: those problems always solve cold.
  }
  *out = p;
  return MCD_OK;
}

int mcd_problem_destroy(mcd_problem* p) {
  if (!p) return MCD_OK;
  cudaSetDevice(p->h->device);
  cudaDeviceSynchronize();          // as cudaFree would: no kernel of any stream still reads the problem
  cudaFreeAsync(p->blob, nullptr);  // back to the pool (no driver unmap)
  delete p;
  return MCD_OK;
}

int mcd_problem_info(const mcd_problem* p, int32_t* np, int64_t* basis_elems, int32_t* sdim, int32_t* in_smem,
                     int32_t* squaring_ok) {
  if (!p) return MCD_E_ARG;
  if (np) *np = p->P.np;
  if (basis_elems) *basis_elems = (int64_t)p->P.np * p->P.ld;
  if (sdim) *sdim = sample_dim(p->P);
  if (in_smem) *in_smem = p->in_smem;
  if (squaring_ok) *squaring_ok = p->squaring_ok;
  return MCD_OK;
}

// mcd_loglike_batch and mcd_propagator_batch: the same kernels, the latter with absorbing rates on the diagonal.
static int loglike_impl(mcd_handle* h, const mcd_problem* p, int32_t method, int32_t B, const double* v, const double* w,
                        const double* lagtimes, const double* sink, double* out_ll, double* out_P, int32_t* status,
                        void* stream) {
  if (!h || !p || B < 0 || !v || !w || !out_ll) return MCD_E_ARG;
  if (method == MCD_METHOD_AUTO)
    method = (p->squaring_ok && (!lagtimes || p->P.use_pull)) ? MCD_METHOD_SQUARING : MCD_METHOD_EIGEN;
  if (method < MCD_METHOD_EIGEN || method > MCD_METHOD_SQUARING_DFMA) return MCD_E_ARG;
  if (method >= MCD_METHOD_SQUARING && !p->squaring_ok) return MCD_E_UNSUPPORTED;
  if (B == 0) return MCD_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const Problem& P = p->P;
  Scratch ws;
  if (P.use_pull) {
    if (method == MCD_METHOD_EIGEN || method == MCD_METHOD_SQUARING_DFMA || sink) return MCD_E_UNSUPPORTED;
    const int grid = B < 8 * h->sm_count ? B : 8 * h->sm_count;
    int rc = ws.alloc((size_t)grid * plan_slots(P) * P.np * P.ld * sizeof(double), st);
    if (rc != MCD_OK) return rc;
    CUDA_TRY(cudaFuncSetAttribute(mcd_loglike_gen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)p->smem_bytes));
    mcd_loglike_gen_kernel<<<grid, kThreads, p->smem_bytes, st>>>(P, B, v, w, lagtimes, out_ll, out_P, status, ws.p);
  } else if (p->blocked && method >= MCD_METHOD_SQUARING) {
    if (method != MCD_METHOD_SQUARING) return MCD_E_UNSUPPORTED;   // tensor-core products only
    const int grid = B < 2 * h->sm_count ? B : 2 * h->sm_count;
    int rc = ws.alloc((size_t)grid * blocked_ws_doubles(P, plan_slots(P)) * sizeof(double), st);
    if (rc != MCD_OK) return rc;
    CUDA_TRY(cudaFuncSetAttribute(mcd_loglike_blk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)p->blk_smem_bytes));
    mcd_loglike_blk_kernel<<<grid, kThreads, p->blk_smem_bytes, st>>>(P, B, v, w, lagtimes, sink, out_ll, out_P, status,
                                                                      ws.p);
  } else if (p->in_smem) {
    const int grid = B < 8 * h->sm_count ? B : 8 * h->sm_count;
    if (method >= MCD_METHOD_SQUARING) {
      int rc = ws.alloc((size_t)grid * plan_slots(P) * P.np * P.ld * sizeof(double), st);
      if (rc != MCD_OK) return rc;
    }
    CUDA_TRY(cudaFuncSetAttribute(mcd_loglike_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)p->smem_bytes));
    mcd_loglike_kernel<true><<<grid, kThreads, p->smem_bytes, st>>>(P, method, B, v, w, lagtimes, sink, out_ll, out_P,
                                                                     status, ws.p);
  } else {
    const int grid = B < 2 * h->sm_count ? B : 2 * h->sm_count;
    int rc = ws.alloc((size_t)grid * smem_matrix_bytes(P.np, P.ld), st);
    if (rc != MCD_OK) return rc;
    CUDA_TRY(cudaFuncSetAttribute(mcd_loglike_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)p->smem_bytes));
    mcd_loglike_kernel<false><<<grid, kThreads, p->smem_bytes, st>>>(P, method, B, v, w, lagtimes, sink, out_ll, out_P,
                                                                      status, ws.p);
  }
  h->launches++;
  return (int)cudaGetLastError();
}

int mcd_loglike_batch(mcd_handle* h, const mcd_problem* p, int32_t method, int32_t B, const double* v,
                      const double* w, const double* lagtimes, double* out_ll, double* out_P, int32_t* status,
                      void* stream) {
  return loglike_impl(h, p, method, B, v, w, lagtimes, nullptr, out_ll, out_P, status, stream);
}

int mcd_propagator_batch(mcd_handle* h, const mcd_problem* p, int32_t B, const double* v, const double* w,
                         const double* sink, const double* times, double* out_P, int32_t* status, void* stream) {
  if (!h || !p || B < 0 || !v || !w || !times || !out_P) return MCD_E_ARG;
  if (p->P.nlag != 1) return MCD_E_ARG;                        // one time per item: a problem created with one lag
  if (p->P.use_pull || !p->squaring_ok) return MCD_E_UNSUPPORTED;
  if (B == 0) return MCD_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  Scratch ll;   // the kernels also form sum N log P of the problem's counts (zeros for a propagator problem): discarded
  int rc = ll.alloc((size_t)B * sizeof(double), (cudaStream_t)stream);
  if (rc != MCD_OK) return rc;
  return loglike_impl(h, p, MCD_METHOD_SQUARING, B, v, w, times, sink, ll.p, out_P, status, stream);
}

// How many two-CTA clusters of the blocked MC kernel the device hosts at once (GPC sizes decide; <= sm_count / 2).
static int max_blk2_clusters(mcd_handle* h, const mcd_problem* p) {
  if (p->max_cl2_plus1 > 0) return p->max_cl2_plus1 - 1;
  p->max_cl2_plus1 = 1;
  if (cudaFuncSetAttribute(mcd_mc_run_blk2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->blk_smem_bytes) !=
      cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (unsigned)h->sm_count, 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = p->blk_smem_bytes;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  int ncl = 0;
  if (cudaOccupancyMaxActiveClusters(&ncl, mcd_mc_run_blk2_kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  p->max_cl2_plus1 = ncl + 1;
  return ncl;
}

int mcd_mc_run(mcd_handle* h, const mcd_problem* p, const mcd_mc_params* params, const mcd_chain_io* io,
               int32_t nchains, int32_t nsteps, void* stream) {
  if (!h || !p || !params || !io || nchains < 0 || nsteps < 0) return MCD_E_ARG;
  if (!io->v || !io->w || !io->scal || !io->counters) return MCD_E_ARG;
  const Problem& P = p->P;
  if (P.ncosF > 0 && (!io->vcoef || !io->naccv_coeff)) return MCD_E_ARG;
  if (P.ncosD > 0 && (!io->wcoef || !io->naccw_coeff)) return MCD_E_ARG;
  if (params->use_tape && (!io->tape_u || !io->tape_idx)) return MCD_E_ARG;
  if (io->basis && !io->basis_valid) return MCD_E_ARG;
  if (nchains == 0 || nsteps == 0) return MCD_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  McParams mp;
  ChainIO cio;
  std::memcpy(&mp, params, sizeof(mp));
  std::memcpy(&cio, io, sizeof(cio));
  const bool sq_ok = p->squaring_ok;
  if (mp.method == MCD_METHOD_AUTO) mp.method = sq_ok ? MCD_METHOD_SQUARING : MCD_METHOD_EIGEN;
  if (mp.method < MCD_METHOD_EIGEN || mp.method > MCD_METHOD_SQUARING_DFMA) return MCD_E_ARG;
  if (mp.method >= MCD_METHOD_SQUARING && !sq_ok) return MCD_E_UNSUPPORTED;
  Scratch ws;
  if (P.use_pull) {
    if (mp.method != MCD_METHOD_SQUARING) return MCD_E_UNSUPPORTED;   // general engine: tensor-core squaring only
    int rc = ws.alloc((size_t)nchains * plan_slots(P) * P.np * P.ld * sizeof(double), st);
    if (rc != MCD_OK) return rc;
    CUDA_TRY(cudaFuncSetAttribute(mcd_mc_run_gen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)p->smem_bytes));
    mcd_mc_run_gen_kernel<<<nchains, kThreads, p->smem_bytes, st>>>(P, mp, cio, nsteps, ws.p);
  } else if (p->blocked && mp.method >= MCD_METHOD_SQUARING) {
    if (mp.method != MCD_METHOD_SQUARING) return MCD_E_UNSUPPORTED;   // tensor-core products only
    // two SMs per chain (thread-block cluster): OPT-IN, MCDIFF_B200_CLUSTER=1.  Measured at 200 bins x 8 lags
    // (profiles/prof_c3_cluster.py): a chain runs 1.6x faster on two SMs, not 2x -- each CTA still stages every operand
    // panel of its blocks, so the copies weigh twice as much -- 1.05e5 against 1.26e5 MC steps/s for one SM per chain.
    //
    // What the cluster kernel IS used for by default: the last, partly filled wave.  One CTA per SM means nchains
    // chains run in ceil(nchains / 148) waves of equal length; when the last wave holds no more chains than the
    // device can host as two-CTA clusters, those chains go to the cluster kernel, launched behind the full waves on
    // the same stream: 512 chains = 3 waves + 68 chains on 136 SMs at 1.6x = 3.6 wave times instead of 4.  Both
    // kernels produce bit-identical chains (test_cluster_kernel_equals_single), so the split is invisible to callers.
    // MCDIFF_B200_CLUSTER=1 forces every chain onto clusters, =0 disables the split.
    const char* want_cl = getenv("MCDIFF_B200_CLUSTER");
    const bool can_cl = P.nblk == 2;
    int ncl = 0;   // chains [nchains - ncl, nchains) run on two SMs each
    if (can_cl && want_cl && want_cl[0] == '1') {
      ncl = nchains;
    } else if (can_cl && !(want_cl && want_cl[0] == '0')) {
      const int tail = nchains % h->sm_count;
      if (tail > 0 && tail <= max_blk2_clusters(h, p)) ncl = tail;
    }
    const int nsingle = nchains - ncl;
    const size_t ws1 = blocked_ws_doubles(P, plan_slots(P));
    int rc = ws.alloc(((size_t)nsingle * ws1 + (size_t)ncl * (ws1 + kClusterTail)) * sizeof(double), st);
    if (rc != MCD_OK) return rc;
    if (nsingle > 0) {
      CUDA_TRY(cudaFuncSetAttribute(mcd_mc_run_blk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)p->blk_smem_bytes));
      mcd_mc_run_blk_kernel<<<nsingle, kThreads, p->blk_smem_bytes, st>>>(P, mp, cio, nsteps, ws.p);
      CUDA_TRY(cudaGetLastError());
    }
    if (ncl > 0) {
      CUDA_TRY(cudaFuncSetAttribute(mcd_mc_run_blk2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)p->blk_smem_bytes));
      mcd_mc_run_blk2_kernel<<<2 * ncl, kThreads, p->blk_smem_bytes, st>>>(P, mp, cio, nsteps,
                                                                           ws.p + (size_t)nsingle * ws1, nsingle);
      if (nsingle > 0) h->launches++;
    }
  } else if (p->in_smem) {
    // Packed symmetric engine (mcd_pk.cu), two chains per SM: OPT-IN (MCDIFF_B200_PACKED=1).  Measured on B200 at
    // 100 bins x 4 lags (profiles/prof_pair.py): a packed chain alone costs 138 us per MC step against 94 us for the
    // full-storage kernel (8 warps and two product rounds instead of 12 warps and one, per-tile address selects in the k
    // loop); two resident chains overlap by 1.36x -> 1.46e6 MC steps/s at 296 chains, 1.33e6 at 1024, against 1.58e6.
    const size_t smem_pk = mcdpk_smem_bytes(P.np);
    const char* want_pk = getenv("MCDIFF_B200_PACKED");
    if (want_pk && want_pk[0] == '1' && mp.method == MCD_METHOD_SQUARING && P.np > kSmallBins &&
        mcdpk_problem_bytes() == sizeof(Problem) && smem_pk <= (size_t)(227 * 1024) / 2 - 1024) {
      int rc = ws.alloc((size_t)nchains * mcdpk_scratch_doubles(&P) * sizeof(double), st);
      if (rc != MCD_OK) return rc;
      rc = mcdpk_launch_mc_run(&P, &mp, &cio, nchains, nsteps, smem_pk, ws.p, st);
      if (rc != MCD_OK) return rc;
      h->launches++;
      return MCD_OK;
    }
    if (mp.method >= MCD_METHOD_SQUARING) {
      int rc = ws.alloc((size_t)nchains * plan_slots(P) * P.np * P.ld * sizeof(double), st);
      if (rc != MCD_OK) return rc;
    }
    if (mp.method == MCD_METHOD_SQUARING && P.np <= kSmallBins && mcd128_problem_bytes() == sizeof(Problem)) {
      // small profiles: 128-thread CTAs, three per SM (mcd_small.cu)
      int rc = mcd128_launch_mc_run(&P, &mp, &cio, nchains, nsteps, mcd128_smem_bytes(P.np, P.ld), ws.p, st);
      if (rc != MCD_OK) return rc;
      h->launches++;
      return MCD_OK;
    }
    if (mp.method == MCD_METHOD_SQUARING) {
      CUDA_TRY(cudaFuncSetAttribute(mcd_mc_run_sq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)p->smem_bytes));
      mcd_mc_run_sq_kernel<<<nchains, kThreads, p->smem_bytes, st>>>(P, mp, cio, nsteps, ws.p);
    } else {
      CUDA_TRY(cudaFuncSetAttribute(mcd_mc_run_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)p->smem_bytes));
      mcd_mc_run_kernel<true><<<nchains, kThreads, p->smem_bytes, st>>>(P, mp, cio, nsteps, ws.p);
    }
  } else {
    const int ntiles = (P.np / 4) * (P.np / 4 + 1) / 2;
    if (ntiles > kThreads) mp.refresh_every = 1;  // always cold: warm product needs one tile per thread
    int rc = ws.alloc((size_t)nchains * smem_matrix_bytes(P.np, P.ld), st);
    if (rc != MCD_OK) return rc;
    CUDA_TRY(cudaFuncSetAttribute(mcd_mc_run_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)p->smem_bytes));
    mcd_mc_run_kernel<false><<<nchains, kThreads, p->smem_bytes, st>>>(P, mp, cio, nsteps, ws.p);
  }
  h->launches++;
  return (int)cudaGetLastError();
}

// ---- radial problem object: private copies of everything that does not change between calls, the LagPlan and the
//      pair-major counts are made ONCE here; the two entry points below launch one kernel and never synchronise ----
struct mcd_rad_problem {
  mcd_handle* h;
  RadProblem R;      // device pointers below are owned by this object
  void* blob;
  double* wrad_basis;   // [n][ncosDrad] or nullptr
  int ncosDrad;
  size_t smem_bytes;
  RadProblem R128;         // the same problem configured for the 128-thread kernels (small profiles, squaring route)
  size_t smem128;          // 0: not used
};

int mcd_rad_problem_create(mcd_handle* h, const mcd_rad_problem_desc* d, void* stream, mcd_rad_problem** out) {
  if (!h || !d || !out) return MCD_E_ARG;
  if (!d->v || !d->w || !d->lagtimes || !d->counts || !d->bessels || !d->b0zeros) return MCD_E_ARG;
  if (d->n < 3 || d->n > 256 || d->nlag < 1 || d->nlag > kMaxLag || d->dim_rad < 1 || d->lmax < 1 ||
      d->lmax > kMaxBessel || d->ncosDrad < 0 || d->ncosDrad > kMaxCoef || (d->ncosDrad > 0 && !d->wrad_basis))
    return MCD_E_ARG;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  mcd_rad_problem* p = new (std::nothrow) mcd_rad_problem();
  if (!p) return MCD_E_NOMEM;
  p->h = h;
  RadProblem& R = p->R;
  const int n = d->n;
  R.n = n; R.np = padded_bins(n); R.ld = leading_dim(R.np); R.pbc = d->pbc ? 1 : 0; R.dim_w = d->pbc ? n : n - 1;
  R.nlag = d->nlag; R.dim_rad = d->dim_rad; R.lmax = d->lmax; R.clip = d->clip; R.reserved = 0;
  const size_t b_v = align256((size_t)n * 8), b_w = align256((size_t)R.dim_w * 8), b_lag = align256((size_t)R.nlag * 8);
  const size_t b_cnt = align256((size_t)R.nlag * R.dim_rad * n * n * 8);
  const size_t b_pc = align256(rad_pcounts_doubles(R) * 8);
  const size_t b_be = align256((size_t)R.lmax * R.dim_rad * 8), b_z = align256((size_t)R.lmax * 8);
  const size_t b_wb = align256((size_t)n * d->ncosDrad * 8);
  unsigned char* blob = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&blob, b_v + b_w + b_lag + b_cnt + b_pc + b_be + b_z + b_wb + 256, st);
  if (e != cudaSuccess) { delete p; return e == cudaErrorMemoryAllocation ? MCD_E_NOMEM : (int)e; }
  p->blob = blob;
  unsigned char* q = blob;
  double* dv = (double*)q; q += b_v;
  double* dw = (double*)q; q += b_w;
  double* dl = (double*)q; q += b_lag;
  double* dc = (double*)q; q += b_cnt;
  double* dp = (double*)q; q += b_pc;
  double* db = (double*)q; q += b_be;
  double* dz = (double*)q; q += b_z;
  double* dwb = (double*)q;
  cudaMemcpyAsync(dv, d->v, (size_t)n * 8, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(dw, d->w, (size_t)R.dim_w * 8, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(dl, d->lagtimes, (size_t)R.nlag * 8, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(dc, d->counts, (size_t)R.nlag * R.dim_rad * n * n * 8, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(db, d->bessels, (size_t)R.lmax * R.dim_rad * 8, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(dz, d->b0zeros, (size_t)R.lmax * 8, cudaMemcpyDeviceToDevice, st);
  if (d->ncosDrad > 0) cudaMemcpyAsync(dwb, d->wrad_basis, (size_t)n * d->ncosDrad * 8, cudaMemcpyDeviceToDevice, st);
  mcd_rad_pair_counts_kernel<<<2 * h->sm_count, 256, 0, st>>>(dc, dp, R.nlag, R.dim_rad, n);
  h->launches++;
  double hl[kMaxLag];
  e = cudaMemcpyAsync(hl, dl, (size_t)R.nlag * 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { cudaFreeAsync(blob, st); delete p; return (int)e; }
  R.v = dv; R.w = dw; R.lagtimes = dl; R.counts = dc; R.pcounts = dp; R.bessels = db; R.b0zeros = dz;
  p->wrad_basis = d->ncosDrad > 0 ? dwb : nullptr;
  p->ncosDrad = d->ncosDrad;
  plan_lags(hl, R.nlag, false, &R.plan);
  if (rad_smem_bytes(R, 1) > (size_t)h->max_smem_optin) { cudaFreeAsync(blob, st); delete p; return MCD_E_UNSUPPORTED; }
  R.nbatch = R.plan.ok ? rad_pick_batch(R, (size_t)h->max_smem_optin) : 1;
  p->smem_bytes = rad_smem_bytes(R, R.nbatch);
  // small profiles on the squaring route: 128-thread CTAs, three per SM (a third of the SM's shared memory each,
  // minus the 1 KB the hardware reserves per CTA)
  p->smem128 = 0;
  p->R128 = R;
  if (R.plan.ok && R.np <= kSmallBins && mcd128_rad_problem_bytes() == sizeof(RadProblem))
    p->smem128 = mcd128_rad_configure(&p->R128, (size_t)(227 * 1024) / 4 - 1024);
  *out = p;
  return MCD_OK;
}

int mcd_rad_problem_destroy(mcd_rad_problem* p) {
  if (!p) return MCD_OK;
  cudaSetDevice(p->h->device);
  cudaDeviceSynchronize();          // as cudaFree would: no kernel of any stream still reads the problem
  cudaFreeAsync(p->blob, nullptr);  // back to the pool (no driver unmap)
  delete p;
  return MCD_OK;
}

int mcd_rad_problem_info(const mcd_rad_problem* p, int32_t* np, int32_t* sample_dim, int32_t* nbatch, int32_t* squaring_ok) {
  if (!p) return MCD_E_ARG;
  if (np) *np = p->R.np;
  if (sample_dim) *sample_dim = 3 + (p->ncosDrad > 0 ? p->ncosDrad : p->R.n);
  if (nbatch) *nbatch = p->R.nbatch;
  if (squaring_ok) *squaring_ok = p->R.plan.ok;
  return MCD_OK;
}

int mcd_rad_loglike_batch(mcd_handle* h, const mcd_rad_problem* p, int32_t B, const double* wrad, double* out_ll,
                          double* out_P, int32_t* status, void* stream) {
  if (!h || !p || !wrad || !out_ll || B < 0) return MCD_E_ARG;
  if (B == 0) return MCD_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const RadProblem& R = p->R;
  const int per_sm = p->smem128 ? 6 : 2;     // two rounds of resident CTAs
  const int grid = B < per_sm * h->sm_count ? B : per_sm * h->sm_count;
  Scratch ws;
  int rc = ws.alloc((size_t)grid * rad_workspace_doubles(R) * sizeof(double), st);
  if (rc != MCD_OK) return rc;
  h->launches++;
  if (p->smem128) {
    rc = mcd128_launch_rad_loglike(&p->R128, grid, B, wrad, out_ll, out_P, status, p->smem128, ws.p, st);
    return rc == 0 ? MCD_OK : rc;
  }
  CUDA_TRY(cudaFuncSetAttribute(mcd_rad_loglike_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bytes));
  mcd_rad_loglike_kernel<<<grid, kThreads, p->smem_bytes, st>>>(R, B, wrad, out_ll, out_P, status, ws.p);
  return (int)cudaGetLastError();
}

static_assert(sizeof(mcd_rad_mc_params) == sizeof(RadMcParams), "mcd_rad_mc_params layout");
static_assert(sizeof(mcd_rad_chain_io) == sizeof(RadChainIO), "mcd_rad_chain_io layout");

int mcd_rad_mc_run(mcd_handle* h, const mcd_rad_problem* p, const mcd_rad_mc_params* params,
                   const mcd_rad_chain_io* io, int32_t nchains, int32_t nsteps, void* stream) {
  if (!h || !p || !params || !io) return MCD_E_ARG;
  if (!io->wrad || !io->scal || !io->counters) return MCD_E_ARG;
  if (params->ncosDrad != p->ncosDrad) return MCD_E_ARG;
  if (params->ncosDrad > 0 && (!io->wcoef || !io->nacc_coeff)) return MCD_E_ARG;
  if (params->use_tape && (!io->tape_u || !io->tape_idx)) return MCD_E_ARG;
  if (nchains <= 0 || nsteps <= 0) return nchains < 0 || nsteps < 0 ? MCD_E_ARG : MCD_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const RadProblem& R = p->R;
  Scratch ws;
  int rc = ws.alloc((size_t)nchains * rad_workspace_doubles(R) * sizeof(double), st);
  if (rc != MCD_OK) return rc;
  RadMcParams mp;
  RadChainIO cio;
  std::memcpy(&mp, params, sizeof(mp));
  std::memcpy(&cio, io, sizeof(cio));
  cio.wrad_basis = p->wrad_basis;   // the library's private copy
  h->launches++;
  if (p->smem128) {
    rc = mcd128_launch_rad_mc_run(&p->R128, &mp, &cio, nchains, nsteps, p->smem128, ws.p, st);
    return rc == 0 ? MCD_OK : rc;
  }
  CUDA_TRY(cudaFuncSetAttribute(mcd_rad_mc_run_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bytes));
  mcd_rad_mc_run_kernel<<<nchains, kThreads, p->smem_bytes, st>>>(R, mp, cio, nsteps, ws.p);
  return (int)cudaGetLastError();
}

#if defined(MCD_PHASE_CLOCK)
// debug build only (profiles/prof_phases.py): read and reset the phase clocks of CTA 0
int mcd_debug_phase_clocks(long long* out32) {
  long long z[32] = {0}, small[32] = {0};
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out32, mcd::g_phase_clk, sizeof(z));
  cudaMemcpyToSymbol(mcd::g_phase_clk, z, sizeof(z));
  mcd128_debug_phase_clocks(small);   // the 128-thread kernels of mcd_small.cu keep their own clocks
  for (int i = 0; i < 32; ++i) out32[i] += small[i];
  return (int)cudaGetLastError();
}
#endif

int mcd_fp64_peak(mcd_handle* h, double* tflops) {
  if (!h || !tflops) return MCD_E_ARG;
  CUDA_TRY(cudaSetDevice(h->device));
  const int blocks = h->sm_count * 8, threads = 256, iters = 4096;
  double* out = nullptr;
  CUDA_TRY(cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  mcd_dfma_kernel<<<blocks, threads>>>(out, 256);  // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    mcd_dfma_kernel<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * 64.0 * iters * (double)blocks * threads;
    const double tf = flop / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  h->launches += 6;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return (int)cudaGetLastError();
}

}  // extern "C"
