// Trajectory -> transition counts on the device (the step BEFORE the likelihood path).
//
//   mcd_traj_counts      <->  transition_matrix_add1   lib/tools/extract.py:71-90, looped over particles and shifts as
//                             examples/create-transition-matrix/extract_Tmat.py:36-50 does (incl. its periodic fold)
//   mcd_traj_counts_rad  <->  count_2D                 lib/tools/extract.py:127-156
//
// HBM-bound integer work, two kernels:
//   1. digitize: every coordinate is read ONCE (8 B), folded into the periodic box and searched in the bin edges
//      (np.digitize semantics: number of edges <= x, NaN -> nedges); the bin index is written as one byte (two when
//      there are more than 255 edges).  Vectorised: 4 coordinates per thread and iteration, one 4/8-byte store.
//   2. count: the pairs (bin[t], bin[t + shift]) of ALL shifts are formed from the byte array (1 B per element and
//      shift group, mostly L2 hits) and accumulated in per-CTA shared-memory histograms -- (nedges+1)^2 32-bit cells
//      per shift, as many shifts per CTA as fit in 227 KB (5 at 100 bins) -- which are flushed once with 64-bit global
//      atomics.  The trajectory array is [frame][particle]: consecutive lanes hold different particles of one frame,
//      so lanes of a warp rarely hit the same cell.  Histograms that do not fit (more than 237 edges) and the radial
//      cube are accumulated with 64-bit atomics in L2 directly.
// Counts are exact integers: results are bit-identical to the reference's loops (tests/test_gpu_traj.py).
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/mcdiff_b200.h"
#include "mcd_handle.h"

namespace mcdtraj {

constexpr int kMaxShifts = 64;
constexpr int kMaxEdges = 4096;   // edges staged in shared memory (32 KB)
constexpr int kDigThreads = 256;
constexpr int kCntThreads = 1024;

struct Shifts {
  long long off[kMaxShifts];   // shift * ncols: distance of the partner element in the flattened [frame][particle] array
};

#define TRAJ_TRY(x)                          \
  do {                                       \
    cudaError_t e_ = (x);                    \
    if (e_ != cudaSuccess) return (int)e_;   \
  } while (0)

// np.digitize(x, edges) for increasing edges (right=False): the number of edges <= x; NaN sorts behind every edge.
// guess: index from the uniform-grid formula (exact for most inputs), verified against the edges and corrected by a
// binary search when it is off -- the result never depends on the guess.
__device__ __forceinline__ int digitize(double x, const double* __restrict__ e, int ne, double e0, double inv_dx) {
  // the conversion saturates (and maps NaN to 0), so the clamp is integer work
  const int k = min(max(__double2int_rz((x - e0) * inv_dx + 1.0), 0), ne);
  if ((k == 0 || e[k - 1] <= x) && (k == ne || e[k] > x)) return k;
  if (x != x) return ne;
  int lo = 0, hi = ne;   // e[j] <= x for j < lo, e[j] > x for j >= hi
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (e[mid] <= x) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// traj = z - zpbc * floor(z / zpbc + 0.5)   (extract_Tmat.py:45), rounded operation by operation like numpy does
// The division is the expensive part (a ~25-instruction FP64 sequence): q' = x * (1/period) differs from the correctly
// rounded quotient by a few ulp, which changes floor(q + 0.5) only when q + 0.5 sits within ~1e-12 (relative) of an
// integer -- those rare coordinates take the exact division, everything else the multiplication.
__device__ __forceinline__ double fold(double x, double period, double inv_period) {
  const double t = __dadd_rn(__dmul_rn(x, inv_period), 0.5);
  double f = floor(t);
  const double frac = t - f;                       // exact for finite t, in [0, 1)
  // safe <=> 1e-9 < frac < 1 - 2^-20 and |t| < 2^20 (the two quotients differ by < |t| 5e-16): compared on the high
  // words (monotonic for non-negative doubles) so that the test runs on the integer pipe, not the FP64 pipe
  const unsigned hf = (unsigned)__double2hiint(frac), ht = (unsigned)__double2hiint(t) & 0x7fffffffu;
  const bool safe = hf > 0x3e112e0bu && hf < 0x3feffffeu && ht < 0x41300000u;
  if (!safe) f = floor(__dadd_rn(__ddiv_rn(x, period), 0.5));   // also NaN / inf (hf or ht out of range)
  return __dsub_rn(x, __dmul_rn(period, f));
}

template <typename BinT>
struct Pack4;
template <>
struct Pack4<uint8_t> {
  using word = uint32_t;
  static __device__ __forceinline__ word make(int a, int b, int c, int d) {
    return (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16) | ((uint32_t)d << 24);
  }
  static __device__ __forceinline__ int get(word w, int e) { return (int)__byte_perm(w, 0u, 0x4440u + (unsigned)e); }
  // four bins starting at element j (any alignment) of an array padded by one word
  static __device__ __forceinline__ word load(const uint8_t* bins, long long j) {
    const word* w = reinterpret_cast<const word*>(bins) + (j >> 2);
    return __funnelshift_r(__ldg(w), __ldg(w + 1), 8 * (int)(j & 3));
  }
  static __device__ __forceinline__ word load_aligned(const uint8_t* bins, long long pack) {
    return __ldg(reinterpret_cast<const word*>(bins) + pack);
  }
};
template <>
struct Pack4<uint16_t> {
  using word = unsigned long long;
  static __device__ __forceinline__ word make(int a, int b, int c, int d) {
    return (word)a | ((word)b << 16) | ((word)c << 32) | ((word)d << 48);
  }
  static __device__ __forceinline__ int get(word w, int e) { return (int)((w >> (16 * e)) & 0xffffull); }
  static __device__ __forceinline__ word load(const uint16_t* bins, long long j) {
    const word* w = reinterpret_cast<const word*>(bins) + (j >> 2);
    const int r = 16 * (int)(j & 3);
    const word w0 = __ldg(w), w1 = __ldg(w + 1);
    return r ? ((w0 >> r) | (w1 << (64 - r))) : w0;
  }
  static __device__ __forceinline__ word load_aligned(const uint16_t* bins, long long pack) {
    return __ldg(reinterpret_cast<const word*>(bins) + pack);
  }
};

// ---- kernel 1: coordinates -> bin indices -------------------------------------------------------------
// Four coordinates per thread and iteration on a STRAIGHT-LINE fast path: the four folds and the four guesses are
// computed without a branch (edges staged with sentinels -inf / +inf around them, so the verification is the same two
// comparisons for every bin), and only if one of them is unsafe does the thread redo that coordinate exactly.  The
// branchy per-coordinate version ran 52 instructions per coordinate and was issue bound (ncu: 68 % issue slots, 4.1
// TB/s); this one is bound by the 9 B per coordinate it moves.
template <typename BinT, bool kVec>
__global__ void __launch_bounds__(kDigThreads)
mcd_traj_digitize_kernel(const double* __restrict__ x, long long N, double period, const double* __restrict__ edges,
                         int ne, BinT* __restrict__ bins) {
  extern __shared__ double se_raw[];   // [ne + 2]: -inf, edges, +inf
  double* se = se_raw + 1;
  for (int k = threadIdx.x; k < ne; k += blockDim.x) se[k] = edges[k];
  if (threadIdx.x == 0) {
    se[-1] = __longlong_as_double((long long)0xfff0000000000000ull);      // -inf
    se[ne] = __longlong_as_double(0x7ff0000000000000ll);                  // +inf
  }
  __syncthreads();
  const double e0 = se[0];
  const double inv_dx = (ne > 1 && se[ne - 1] > e0) ? (double)(ne - 1) / (se[ne - 1] - e0) : 0.0;
  const bool do_fold = period > 0.0;
  const double inv_period = do_fold ? 1.0 / period : 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (kVec) {
    using P = Pack4<BinT>;
    const long long npack = N >> 2;
    typename P::word* out = reinterpret_cast<typename P::word*>(bins);
    // the 32 bytes of the next pack are requested before this one is digitized (ncu: long_scoreboard was the top stall)
    double2 a = make_double2(0.0, 0.0), b = a;
    if (t0 < npack) {
      a = __ldcs(reinterpret_cast<const double2*>(x) + 2 * t0);
      b = __ldcs(reinterpret_cast<const double2*>(x) + 2 * t0 + 1);
    }
    for (long long p = t0; p < npack; p += stride) {
      double v[4] = {a.x, a.y, b.x, b.y};
      if (p + stride < npack) {
        a = __ldcs(reinterpret_cast<const double2*>(x) + 2 * (p + stride));
        b = __ldcs(reinterpret_cast<const double2*>(x) + 2 * (p + stride) + 1);
      }
      if (do_fold) {
        double f[4];
        unsigned unsafe = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const double t = __dadd_rn(__dmul_rn(v[e], inv_period), 0.5);
          f[e] = floor(t);
          const unsigned hf = (unsigned)__double2hiint(t - f[e]), ht = (unsigned)__double2hiint(t) & 0x7fffffffu;
          unsafe |= (hf > 0x3e112e0bu && hf < 0x3feffffeu && ht < 0x41300000u) ? 0u : (1u << e);   // see fold()
        }
        if (unsafe) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (unsafe & (1u << e)) f[e] = floor(__dadd_rn(__ddiv_rn(v[e], period), 0.5));
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = __dsub_rn(v[e], __dmul_rn(period, f[e]));
      }
      int k[4];
      unsigned miss = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        k[e] = min(max(__double2int_rz((v[e] - e0) * inv_dx + 1.0), 0), ne);
        miss |= (se[k[e] - 1] <= v[e] && v[e] < se[k[e]]) ? 0u : (1u << e);
      }
      if (miss) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (miss & (1u << e)) k[e] = digitize(v[e], se, ne, e0, inv_dx);
      }
      out[p] = P::make(k[0], k[1], k[2], k[3]);
    }
    for (long long i = (npack << 2) + t0; i < N; i += stride)
      bins[i] = (BinT)digitize(do_fold ? fold(x[i], period, inv_period) : x[i], se, ne, e0, inv_dx);
  } else {
    for (long long i = t0; i < N; i += stride)
      bins[i] = (BinT)digitize(do_fold ? fold(x[i], period, inv_period) : x[i], se, ne, e0, inv_dx);
  }
}

// ---- kernel 2: bin pairs -> counts[l][end][start] ------------------------------------------------------
// grid (slices, shift groups).  kSmem: the group's histograms live in shared memory (32-bit cells) and are flushed
// once; otherwise every pair is a 64-bit atomic on the output (L2).
template <typename BinT, bool kSmem>
__global__ void __launch_bounds__(kCntThreads, 1)
mcd_traj_count_kernel(const BinT* __restrict__ bins, long long N, Shifts sh, int L, int gl, int H,
                      unsigned long long* __restrict__ out) {
  using P = Pack4<BinT>;
  using word = typename P::word;
  extern __shared__ unsigned hist[];
  const int l0 = blockIdx.y * gl;
  const int nl = min(gl, L - l0);
  const int HH = H * H;
  if (kSmem) {
    for (int c = threadIdx.x; c < nl * HH; c += blockDim.x) hist[c] = 0u;
    __syncthreads();
  }
  long long maxoff = 0;
  for (int k = 0; k < nl; ++k) maxoff = max(maxoff, sh.off[l0 + k]);
  unsigned long long* outg = out + (size_t)l0 * HH;
  const long long npack = (N + 3) >> 2;
  const long long nfast = (N - 3 - maxoff) >> 2;   // packs p < nfast: all four partners of every shift exist
  const long long stride = (long long)gridDim.x * blockDim.x;
  // fast packs (every partner exists): at most four shifts per CTA (mcd_traj_counts caps the group), software
  // pipelined -- the nine word loads of the next pack are in flight while the 16 increments of this one issue
  long long off[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) off[q] = sh.off[l0 + (q < nl ? q : 0)];
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  word s4 = 0, e4[4] = {0, 0, 0, 0};
  if (p < nfast) {
    s4 = P::load_aligned(bins, p);
#pragma unroll
    for (int q = 0; q < 4; ++q) e4[q] = (q < nl) ? P::load(bins, (p << 2) + off[q]) : 0;
  }
  while (p < nfast) {
    const long long pn = p + stride;
    word s4n = 0, e4n[4] = {0, 0, 0, 0};
    if (pn < nfast) {
      s4n = P::load_aligned(bins, pn);
#pragma unroll
      for (int q = 0; q < 4; ++q) e4n[q] = (q < nl) ? P::load(bins, (pn << 2) + off[q]) : 0;
    }
    const int s0 = P::get(s4, 0), s1 = P::get(s4, 1), s2 = P::get(s4, 2), s3 = P::get(s4, 3);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (q < nl) {
        if (kSmem) {
          // byte offsets end * 4H + 4 start against the (warp-uniform) base of this shift's histogram: three
          // instructions per increment (byte extract, multiply-add, ATOMS [reg + uniform]); A[end, start] += 1
          char* hk = reinterpret_cast<char*>(hist + q * HH);
          const int H4 = 4 * H;
          atomicAdd(reinterpret_cast<unsigned*>(hk + (P::get(e4[q], 0) * H4 + 4 * s0)), 1u);
          atomicAdd(reinterpret_cast<unsigned*>(hk + (P::get(e4[q], 1) * H4 + 4 * s1)), 1u);
          atomicAdd(reinterpret_cast<unsigned*>(hk + (P::get(e4[q], 2) * H4 + 4 * s2)), 1u);
          atomicAdd(reinterpret_cast<unsigned*>(hk + (P::get(e4[q], 3) * H4 + 4 * s3)), 1u);
        } else {
          unsigned long long* ok = outg + (size_t)q * HH;
          atomicAdd(ok + (P::get(e4[q], 0) * H + s0), 1ull);
          atomicAdd(ok + (P::get(e4[q], 1) * H + s1), 1ull);
          atomicAdd(ok + (P::get(e4[q], 2) * H + s2), 1ull);
          atomicAdd(ok + (P::get(e4[q], 3) * H + s3), 1ull);
        }
      }
    }
    s4 = s4n;
#pragma unroll
    for (int q = 0; q < 4; ++q) e4[q] = e4n[q];
    p = pn;
  }
  // the last max(shift) frames: partner by partner
  for (; p < npack; p += stride) {
    const long long i = p << 2;
    const word t4 = P::load_aligned(bins, p);
    const int sv[4] = {P::get(t4, 0), P::get(t4, 1), P::get(t4, 2), P::get(t4, 3)};
    for (int k = 0; k < nl; ++k) {
      const long long j = i + sh.off[l0 + k];
      if (j >= N) continue;
      const word w4 = P::load(bins, j);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (j + e < N) {
          const int cell = P::get(w4, e) * H + sv[e];
          if (kSmem) atomicAdd(&hist[k * HH + cell], 1u);
          else atomicAdd(&outg[(size_t)k * HH + cell], 1ull);
        }
      }
    }
  }
  if (kSmem) {
    __syncthreads();
    for (int c = threadIdx.x; c < nl * HH; c += blockDim.x) {
      const unsigned v = hist[c];
      if (v) atomicAdd(&outg[c], (unsigned long long)v);
    }
  }
}

// ---- radial cube: B[l][r-1][end][start]  (extract.py:139-154) --------------------------------------------
template <typename BinT>
__global__ void __launch_bounds__(256)
mcd_traj_count_rad_kernel(const BinT* __restrict__ zb, const double* __restrict__ X, const double* __restrict__ Y,
                          long long N, Shifts sh, int L, const double* __restrict__ redges, int nre, int H,
                          unsigned long long* __restrict__ out) {
  extern __shared__ double sr[];
  for (int k = threadIdx.x; k < nre; k += blockDim.x) sr[k] = redges[k];
  __syncthreads();
  const double r0 = sr[0];
  const double inv_dr = (nre > 1 && sr[nre - 1] > r0) ? (double)(nre - 1) / (sr[nre - 1] - r0) : 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    const int s = (int)zb[i];
    const double x0 = X[i], y0 = Y[i];
    for (int l = 0; l < L; ++l) {
      const long long j = i + sh.off[l];
      if (j >= N) continue;
      const int e = (int)zb[j];
      const double dx = __dsub_rn(X[j], x0), dy = __dsub_rn(Y[j], y0);
      const double dr = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
      int r = digitize(dr, sr, nre, r0, inv_dr) - 1;
      if (r < 0) r += nre;   // the reference indexes B[r-1]: r = 0 wraps to the last radial bin
      atomicAdd(&out[(((size_t)l * nre + r) * H + e) * H + s], 1ull);
    }
  }
}

struct TrajScratch {
  void* p = nullptr;
  cudaStream_t st = nullptr;
  int alloc(size_t bytes, cudaStream_t stream) {
    st = stream;
    cudaError_t e = cudaMallocAsync(&p, bytes, stream);
    if (e != cudaSuccess) { p = nullptr; return e == cudaErrorMemoryAllocation ? MCD_E_NOMEM : (int)e; }
    return MCD_OK;
  }
  ~TrajScratch() { if (p) cudaFreeAsync(p, st); }
};

int check_shifts(const int32_t* shifts_host, int nshifts, long long nframes, long long ncols, Shifts* sh) {
  if (!shifts_host || nshifts < 1 || nshifts > kMaxShifts) return MCD_E_ARG;
  for (int l = 0; l < nshifts; ++l) {
    if (shifts_host[l] < 1 || (long long)shifts_host[l] >= nframes) return MCD_E_ARG;   // extract.py:73 asserts shift < len(x)
    sh->off[l] = (long long)shifts_host[l] * ncols;
  }
  return MCD_OK;
}

template <typename BinT>
int digitize_launch(mcd_handle* h, const double* x, long long N, double period, const double* edges, int ne, BinT* bins,
                    cudaStream_t st) {
  const bool vec = (reinterpret_cast<uintptr_t>(x) & 15u) == 0;
  long long want = (N / 4 + kDigThreads - 1) / kDigThreads;
  const int grid = (int)(want < 1 ? 1 : (want > (long long)h->sm_count * 32 ? (long long)h->sm_count * 32 : want));
  const size_t smem = (size_t)(ne + 2) * sizeof(double);
  if (vec) mcd_traj_digitize_kernel<BinT, true><<<grid, kDigThreads, smem, st>>>(x, N, period, edges, ne, bins);
  else mcd_traj_digitize_kernel<BinT, false><<<grid, kDigThreads, smem, st>>>(x, N, period, edges, ne, bins);
  TRAJ_TRY(cudaGetLastError());
  h->launches++;
  return MCD_OK;
}

template <typename BinT>
int counts_impl(mcd_handle* h, const double* x, long long nframes, long long ncols, double period, const double* edges,
                int ne, const Shifts& sh, int L, unsigned long long* counts, cudaStream_t st) {
  const long long N = nframes * ncols;
  const int H = ne + 1;
  TrajScratch bins;
  int rc = bins.alloc((size_t)N * sizeof(BinT) + 64, st);   // + one pack of padding for the unaligned pack loads
  if (rc != MCD_OK) return rc;
  TRAJ_TRY(cudaMemsetAsync((char*)bins.p + (size_t)N * sizeof(BinT), 0, 64, st));
  rc = digitize_launch<BinT>(h, x, N, period, edges, ne, (BinT*)bins.p, st);
  if (rc != MCD_OK) return rc;
  const size_t hist = (size_t)H * H * sizeof(unsigned);
  int gl = (int)((size_t)h->max_smem_optin / hist);
  const long long npack = (N + 3) / 4;
  if (gl >= 1) {
    if (gl > L) gl = L;
    if (gl > 4) gl = 4;   // the count kernel pipelines one group of at most four shifts
    const int G = (L + gl - 1) / gl;
    gl = (L + G - 1) / G;   // balanced groups
    long long S = h->sm_count / G;
    if (S < 1) S = 1;
    const long long need = (npack + kCntThreads - 1) / kCntThreads;
    if (S > need) S = need < 1 ? 1 : need;
    TRAJ_TRY(cudaFuncSetAttribute(mcd_traj_count_kernel<BinT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(gl * hist)));
    mcd_traj_count_kernel<BinT, true><<<dim3((unsigned)S, (unsigned)G), kCntThreads, gl * hist, st>>>(
        (const BinT*)bins.p, N, sh, L, gl, H, counts);
  } else {
    long long S = (npack + kCntThreads - 1) / kCntThreads;
    if (S > 2LL * h->sm_count) S = 2LL * h->sm_count;
    if (S < 1) S = 1;
    mcd_traj_count_kernel<BinT, false><<<dim3((unsigned)S, (unsigned)((L + 3) / 4)), kCntThreads, 0, st>>>(
        (const BinT*)bins.p, N, sh, L, 4, H, counts);
  }
  TRAJ_TRY(cudaGetLastError());
  h->launches++;
  return MCD_OK;
}

template <typename BinT>
int counts_rad_impl(mcd_handle* h, const double* x, const double* y, const double* z, long long nframes, long long ncols,
                    double period, const double* edges, int ne, const double* redges, int nre, const Shifts& sh, int L,
                    unsigned long long* counts, cudaStream_t st) {
  const long long N = nframes * ncols;
  TrajScratch bins;
  int rc = bins.alloc((size_t)N * sizeof(BinT) + 64, st);
  if (rc != MCD_OK) return rc;
  rc = digitize_launch<BinT>(h, z, N, period, edges, ne, (BinT*)bins.p, st);
  if (rc != MCD_OK) return rc;
  long long S = (N + 255) / 256;
  if (S > 8LL * h->sm_count) S = 8LL * h->sm_count;
  if (S < 1) S = 1;
  mcd_traj_count_rad_kernel<BinT><<<(unsigned)S, 256, (size_t)nre * sizeof(double), st>>>(
      (const BinT*)bins.p, x, y, N, sh, L, redges, nre, ne + 1, counts);
  TRAJ_TRY(cudaGetLastError());
  h->launches++;
  return MCD_OK;
}

}  // namespace mcdtraj

using namespace mcdtraj;

extern "C" {

int mcd_traj_counts(mcd_handle* h, const double* x, int64_t nframes, int64_t ncols, double period, const double* edges,
                    int32_t nedges, const int32_t* shifts_host, int32_t nshifts, int64_t* counts, void* stream) {
  if (!h || !x || !edges || !counts) return MCD_E_ARG;
  if (nframes < 2 || ncols < 1 || nedges < 1 || nedges > kMaxEdges || !(period >= 0.0)) return MCD_E_ARG;
  Shifts sh;
  int rc = check_shifts(shifts_host, nshifts, nframes, ncols, &sh);
  if (rc != MCD_OK) return rc;
  TRAJ_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* out = reinterpret_cast<unsigned long long*>(counts);
  if (nedges <= 255) return counts_impl<uint8_t>(h, x, nframes, ncols, period, edges, nedges, sh, nshifts, out, st);
  return counts_impl<uint16_t>(h, x, nframes, ncols, period, edges, nedges, sh, nshifts, out, st);
}

int mcd_traj_counts_rad(mcd_handle* h, const double* x, const double* y, const double* z, int64_t nframes, int64_t ncols,
                        double period, const double* edges, int32_t nedges, const double* redges, int32_t nredges,
                        const int32_t* shifts_host, int32_t nshifts, int64_t* counts, void* stream) {
  if (!h || !x || !y || !z || !edges || !redges || !counts) return MCD_E_ARG;
  if (nframes < 2 || ncols < 1 || nedges < 1 || nedges > kMaxEdges || nredges < 1 || nredges > kMaxEdges ||
      !(period >= 0.0))
    return MCD_E_ARG;
  Shifts sh;
  int rc = check_shifts(shifts_host, nshifts, nframes, ncols, &sh);
  if (rc != MCD_OK) return rc;
  TRAJ_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* out = reinterpret_cast<unsigned long long*>(counts);
  if (nedges <= 255)
    return counts_rad_impl<uint8_t>(h, x, y, z, nframes, ncols, period, edges, nedges, redges, nredges, sh, nshifts, out, st);
  return counts_rad_impl<uint16_t>(h, x, y, z, nframes, ncols, period, edges, nedges, redges, nredges, sh, nshifts, out, st);
}

}  // extern "C"
