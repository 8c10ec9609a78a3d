// One Monte-Carlo chain = one CTA.  Everything a chain touches per step lives in shared
// memory: the eigenvector matrix XT (row k = k-th eigenvector of the symmetrised rate
// matrix), a second n x n work matrix B2 (S.X / rotated matrix / scaled eigenvectors), and
// the profile vectors.  Count matrices are identical for all chains and stay L2 resident.
//
// Path restated here (reference file:line, /root/reference):
//   rate matrix                 lib/utils.py:77-131     -> build_S (symmetrised form)
//   P = expm(lag R), log L      lib/utils.py:307-368    -> jacobi + loglike
//   moves / Metropolis          lib/MCState.py:205-304  -> mc_step
//   width / temperature update  lib/MCState.py:371-393  -> mc_step (tail)
//   step order, sampling        lib/mc.py:27-51, lib/log.py:35-58
//
// Math: with Pi = diag(exp(-v)), S = Pi^-1/2 R Pi^1/2 is symmetric, S[i,i+1] = exp(w_i),
// S[i,i] = R[i,i]; P = Pi^1/2 V exp(t Lambda) V^T Pi^-1/2, one eigendecomposition serves
// every lag time.
#pragma once
#include <cstdint>
#include <cstring>

#include "mcd_exec.cuh"

namespace MCD_NS {

constexpr int kMaxLag = 64;   // LagPlan arrays are sized for this
constexpr int kMaxCoef = 64;
constexpr int kMaxSweeps = 60;
constexpr int kTaylorMargin = 33;
constexpr int kMaxTasks = 16;       // warp tasks of a symmetric product (more tasks than warps: the product runs in rounds)   // wrap margin of the Taylor coefficient arrays (Chain::taylor_coefs)

// How the squaring engine reaches every lag time (built on the host, mcd_lib.cu::plan_lags).
// Lags are visited in ascending order; G(t_i) = G(t_{i-1}) G(delta_i), delta_i = t_i - t_{i-1}.  Each
// distinct operand G(delta) comes from its own squaring chain or is an earlier G(t_j) parked in a
// per-CTA global (L2 resident) slot.  `fast`: arithmetic progression t_i = (i+1) t_0, no slots.
constexpr int kMaxChains = 8;
constexpr int kMaxSlots = 8;
struct LagPlan {
  int ok = 0, fast = 0, nchain = 0, nslots = 0;
  signed char order[64] = {};             // lag index visited i-th (ascending lag time)
  signed char chain_a[kMaxChains] = {};   // chain time = lagt[a] - (b >= 0 ? lagt[b] : 0); the LAST chain is G(t_0)
  signed char chain_b[kMaxChains] = {};
  signed char chain_slot[kMaxChains] = {};  // slot the chain result is parked in, -1: none
  signed char op_slot[64] = {};           // i >= 1: slot holding G(delta_i); -1: delta_i = 0 (same kernel as i-1)
  signed char park[64] = {};              // slot G(t_i) is parked in after lag i, -1: none
};

struct Problem {
  int n, np, ld, dim_w, pbc, nlag;
  int ncosF, ncosD;
  int nbk, lds, blocked;   // shared-memory panel size and pitch (= np, ld unless blocked), see set_layout
  int nblk, reserved2;     // blocks per dimension of the blocked engine (np = nblk * nbk), 1 otherwise
  const double* lagtimes;  // [nlag]
  const int* counts;       // [nlag][np][np]  (padded, counts[l][i][j] = # transitions j -> i)
  const int* nsym;         // [nlag][np][np]  N_ij + N_ji (i != j), N_ii on the diagonal
  const double* rvec;      // [np]  sum_lag (column sum - row sum) of the counts: the part of log L linear in v
  const double* v_basis;   // [n][ncosF]      (null when ncosF == 0)
  const double* w_basis;   // [dim_w][ncosD]  (null when ncosD == 0)
  const double* svecs;     // [dim_w][dim_w]  Laplacian eigenvectors (null when spring_k <= 0)
  double spring_k;
  double clip;
  LagPlan plan[2];         // [0]: lag times as given, [1]: lag times shifted by a time offset (--t0)
  double pull;             // dF per bin of a constant pulling force (lib/utils.py:133-183), used when use_pull
  int use_pull;            // 1: non-reversible generator -> general (non-symmetric) squaring engine
  int reserved;
};

MCD_HOSTDEV bool plan_close(double a, double b) { return fabs(a - b) <= 1e-12 * fabs(b); }

// Build the LagPlan of the squaring engine (see mcd_chain.cuh).  shifted: a time offset is added to
// every lag at run time, so only DIFFERENCES of lag times are known to coincide.
MCD_HOSTDEV void plan_lags(const double* lag, int L, bool shifted, LagPlan* pl) {
  *pl = LagPlan();
  int ord[kMaxLag];
  for (int i = 0; i < L; ++i) ord[i] = i;
  for (int i = 1; i < L; ++i)
    for (int j = i; j > 0 && lag[ord[j]] < lag[ord[j - 1]]; --j) { int t = ord[j]; ord[j] = ord[j - 1]; ord[j - 1] = t; }
  for (int i = 0; i < L; ++i) { pl->order[i] = (signed char)ord[i]; pl->op_slot[i] = -1; pl->park[i] = -1; }
  if (!(lag[ord[0]] > 0.0)) return;

  // operand descriptors: either "G(t_j)" (j = sorted index) or a fresh chain for a difference
  int nslots = 0, nchain = 0;
  int slot_of_R[kMaxLag];
  for (int i = 0; i < L; ++i) slot_of_R[i] = -1;
  int chain_first[kMaxChains];          // sorted index i whose difference the chain computes
  bool fast = !shifted;
  for (int i = 1; i < L; ++i) {
    const double ti = lag[ord[i]], tp = lag[ord[i - 1]];
    const double delta = ti - tp;
    if (!(delta > 1e-12 * ti)) { fast = false; continue; }      // duplicate lag time: op_slot = -1
    int slot = -1;
    if (!shifted) {
      for (int j = 0; j < i && slot < 0; ++j)
        if (plan_close(lag[ord[j]], delta)) {
          if (slot_of_R[j] < 0) { if (nslots >= kMaxSlots) return; slot_of_R[j] = nslots++; }
          slot = slot_of_R[j];
          if (j != 0) fast = false;
        }
    }
    for (int c = 0; c < nchain && slot < 0; ++c) {
      const int k = chain_first[c];
      if (plan_close(lag[ord[k]] - lag[ord[k - 1]], delta)) slot = pl->chain_slot[c];
    }
    if (slot < 0) {
      if (nchain >= kMaxChains - 1 || nslots >= kMaxSlots) return;
      fast = false;
      chain_first[nchain] = i;
      pl->chain_a[nchain] = (signed char)ord[i];
      pl->chain_b[nchain] = (signed char)ord[i - 1];
      pl->chain_slot[nchain] = (signed char)nslots;
      slot = nslots++;
      ++nchain;
    }
    pl->op_slot[i] = (signed char)slot;
  }
  // the G(t_0) chain comes last
  pl->chain_a[nchain] = (signed char)ord[0];
  pl->chain_b[nchain] = -1;
  pl->chain_slot[nchain] = (signed char)slot_of_R[0];
  ++nchain;
  for (int i = 1; i < L; ++i) pl->park[i] = (signed char)slot_of_R[i];
  pl->nchain = nchain;
  pl->nslots = fast ? 0 : nslots;
  pl->fast = fast ? 1 : 0;
  pl->ok = 1;
}

// Slots a launch must provide per chain: either plan may be walked (a chain whose time offset is non-zero uses the
// shifted plan even when no time-offset moves are made, see Chain::run).
MCD_HOSTDEV int plan_slots(const Problem& P) {
  return P.plan[0].nslots > P.plan[1].nslots ? P.plan[0].nslots : P.plan[1].nslots;
}

// Lag times handed to a squaring engine must be the problem's lag times plus ONE offset (the time-offset trial of
// lib/MCState.py:207-210): only that form has a LagPlan (plan[1]).
MCD_HOSTDEV bool shifted_lags_ok(const Problem& P, const double* lag) {
  const double shift = lag[0] - P.lagtimes[0];
  for (int l = 1; l < P.nlag; ++l)
    if (!(fabs((lag[l] - P.lagtimes[l]) - shift) <= 1e-9 * fabs(P.lagtimes[l]))) return false;
  return true;
}

struct McParams {
  double num_mc_update;  // lib/MCState.py:382 (parsed as float by the CLI)
  double dtemp;          // lib/MCState.py:46-51
  double min_lt;         // lib/MCState.py:207
  int move_timezero;
  int sample_every;      // lib/log.py:9  (100)
  int refresh_every;     // cold re-solve period (bounds orthogonality drift of the warm start)
  int use_tape;          // 1: proposals/uniforms from the host tape, 0: Philox4x32-10
  int method;            // 1: Jacobi eigen, 2: scaling-and-squaring on FP64 tensor cores, 3: same with DFMA
  int taylor_m;          // Taylor order of the squaring engine (0 = default 20)
  unsigned long long seed;
  long long chain_offset;  // global id of chain 0 of this launch (multi-GPU sharding)
  long long tape_stride;   // 0: all chains share one tape, else per-chain stride in steps
  long long nsamples_cap;  // rows available per chain in the sample buffer
};

// scal[] slots
enum { S_LL = 0, S_T0, S_DV, S_DW, S_DT0, S_TEMP, S_NSCAL = 8 };
// counters[] slots
enum { C_NACCV = 0, C_NACCW, C_NACCT0, C_NACCV_UPD, C_NACCW_UPD, C_STEP, C_STATUS, C_SWEEPS, C_NCOUNT = 8 };
// status bits
// ST_NONFINITE: the propagator engine rejected the profile (eigensolver did not converge, or the squaring engines met
// a non-finite / absurd norm lag * ||S||); ST_BAD_LAGS: see shifted_lags_ok
enum { ST_NONFINITE = 1, ST_NAN_TRIAL = 2, ST_BAD_LAGS = 4 };

struct ChainIO {
  double* v;        // [C][n]
  double* w;        // [C][dim_w]
  double* vcoef;    // [C][ncosF]
  double* wcoef;    // [C][ncosD]
  double* scal;     // [C][S_NSCAL]
  long long* counters;     // [C][C_NCOUNT]
  long long* naccv_coeff;  // [C][max(ncosF,1)]
  long long* naccw_coeff;  // [C][max(ncosD,1)]
  double* basis;    // [C][np*ld] warm-start eigenvectors (may be null)
  int* basis_valid; // [C]
  // optional per-step records
  const double* tape_u;   // [steps][5]
  const int* tape_idx;    // [steps]
  unsigned char* decisions;  // [C][nsteps]  bit0 accepted, bit1 t0 accepted, bits 2-3 move type
  double* ll_try;            // [C][nsteps]
  double* samples;           // [C][nsamples_cap][sample_dim]
};

// Padded bin count (multiple of 8: 8x8 tensor tiles, 4x4 register tiles) and leading dimension.
// ld = 4 mod 8 doubles: a 64-bit shared load is served per half-warp, and the half-warp of an
// m8n8k4 fragment load touches 4 consecutive k-rows x 4 doubles; with a row pitch of 32 mod 64
// bytes those four 32-byte pieces fall into four different bank groups (conflict-free).  Rows stay
// 16-byte aligned for the 128-bit loads of the vector-pipe kernels.
MCD_HOSTDEV int padded_bins(int n) { return (n + 7) & ~7; }
MCD_HOSTDEV int leading_dim(int np) { return np + 4; }

// Largest bin count whose two work matrices fit in shared memory; beyond it the squaring engine runs
// blocked: matrices live in a per-CTA global workspace (L2) and are multiplied as nblk x nblk blocks
// (nblk = 2 ... 5) whose nbk x nbk panels (nbk <= 104) are staged through the two shared-memory buffers.
constexpr int kMaxSmemBins = 104;
constexpr int kMaxBlocks = 5;
MCD_HOSTDEV void set_layout(Problem& P, int n) {
  P.n = n;
  P.nblk = (n + kMaxSmemBins - 1) / kMaxSmemBins;
  P.blocked = (P.nblk >= 2 && P.nblk <= kMaxBlocks) ? 1 : 0;
  if (!P.blocked) P.nblk = 1;
  const int q = 8 * P.nblk;
  P.np = ((n + q - 1) / q) * q;
  P.ld = leading_dim(P.np);
  P.nbk = P.np / P.nblk;
  P.lds = leading_dim(P.nbk);
  P.reserved2 = 0;
}

MCD_HOSTDEV int sample_dim(const Problem& P) {
  return 6 + (P.ncosF > 0 ? P.ncosF : P.n) + (P.ncosD > 0 ? P.ncosD : P.dim_w);
}

// ---------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  counter = (step_lo, step_hi, chain, stream)
// ---------------------------------------------------------------------------------------
MCD_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
MCD_HD double u53(uint32_t hi, uint32_t lo) {
  return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

// ---------------------------------------------------------------------------------------
// shared-memory carve-up
// ---------------------------------------------------------------------------------------
struct Scalars {
  double ll_cur, ll_try, timezero, t0_try, dv, dw, dt0, temp, scale, delta, e_spring, gmin, lin, gersh;
  long long naccv, naccw, nacct0, naccv_upd, naccw_upd, step, status, sweeps;
  int move, idx, accepted, accepted_t0, nsamp, basis_is_trial, do_t0, mma_tasks;
};

struct Smem {
  double* XT;
  double* B2;
  double *a, *b, *ea, *eb, *hv, *lam, *el;
  double *lgc, *lgl;   // log tables: c_k ~ 1/(1+(k+1/2)/128), -log(c_k)
  double *v, *w, *vt, *wt;
  double *vc, *wc, *ct;
  double *lag, *lag_try;
  double *rc, *rs;
  double* red;
  double* part;  // per-thread likelihood partial sums (squaring engine)
  double* text;  // coefficient arrays of the Taylor passes (taylor_coefs); null in the panel layout (the panels serve)
  int *rp, *rq;
  int* wti;                // warp tasks as 16 ints per warp: {nt, na, rowA, rowB, jb[8], -, -, -, -} (three 128-bit loads)
  unsigned char* tiles;    // 2 bytes per tile (ib, jb), ib <= jb
  unsigned char* jblocks;  // 2 bytes per Jacobi block (ka, kb), ka < kb
  unsigned char* wtab;     // 16 bytes per warp task of the tensor-core product (see build_tables)
  unsigned char* hrow;     // Horner passes: band row owned by thread tid (tid / tpr) ...
  unsigned char* hrem;     // ... and its slot inside the row (tid % tpr); 255 = no row
  unsigned char* tord;     // Taylor order for s = 0..31 squarings at argument bound kTheta; [32 + s]: at kThetaHi (255: none)
  Scalars* sc;
  int ntiles, nblk;
};

// panels: layout of the blocked engine, without the vectors and tables only the eigen engine uses
// (lam, el, rc, rs, rp, rq, tile and Jacobi-block lists) -- at np = 520 those alone would be 100 KB
MCD_HOSTDEV size_t smem_vectors_bytes(int np, bool panels = false) {
  size_t d = (size_t)np * (panels ? 9 : 11) + 256 + kMaxCoef * 3 + kMaxLag * 2 + (panels ? 0 : (size_t)np) /*rc,rs*/ + 32 + kThreads +
             (panels ? 0 : (size_t)6 * (np + 2 * kTaylorMargin)) /*text*/;
  size_t nb = np / 4, h = np / 2;
  size_t bytes = d * 8 + (panels ? 0 : (size_t)np * 4) /*rp,rq*/ + 64 * kMaxTasks /*wti*/ +
                 (panels ? 0 : nb * (nb + 1) + h * (h - 1)) + 16 * kMaxTasks + sizeof(Scalars) + 64 +
                 2 * kThreads + 64 /*hrow, hrem, tord (two tables of 32)*/;
  return (bytes + 15) & ~(size_t)15;
}
MCD_HOSTDEV size_t smem_matrix_bytes(int np, int ld) { return (size_t)2 * np * ld * 8; }

// Packed symmetric storage (kMode 3): only the 8 x 8 tiles (a, b), a <= b, of a symmetric matrix are kept, 64
// contiguous doubles each, row-major inside the tile with the column index XOR-swizzled by the row ((r >> 1) & 1)
// << 2 -- with that, the m8n8k4 fragment loads of a tile AND of its transpose are both bank-conflict free.
// Diagonal tiles are stored in full.  91 tiles = 46.6 KB at 104 bins instead of 89.9 KB: two work matrices of TWO
// chains fit in one SM.
MCD_HOSTDEV int pk_tile(int a, int b, int nb8) { return a * nb8 - a * (a - 1) / 2 + (b - a); }   // a <= b
MCD_HOSTDEV int pk_off(int r, int c) { return r * 8 + (c ^ (((r >> 1) & 1) << 2)); }
MCD_HOSTDEV size_t pk_elems(int np) { return (size_t)(np / 8) * (np / 8 + 1) / 2 * 64; }
MCD_HOSTDEV size_t pk_index(int i, int j, int nb8) {
  const int a = i >> 3, b = j >> 3;
  return a <= b ? (size_t)pk_tile(a, b, nb8) * 64 + pk_off(i & 7, j & 7) : (size_t)pk_tile(b, a, nb8) * 64 + pk_off(j & 7, i & 7);
}
MCD_HOSTDEV size_t smem_packed_bytes(int np) { return 2 * pk_elems(np) * 8; }
// global workspace of one chain of the blocked squaring engine: 3 packed work matrices + the plan's slots
MCD_HOSTDEV size_t blocked_ws_doubles(const Problem& P, int nslots) { return (size_t)(3 + nslots) * pk_elems(P.np); }

// mat: storage for XT and B2 (2*np*ld doubles; shared or global), vec: shared storage.
// panels: mat holds the two nbk x lds operand panels of the blocked squaring engine; packed: two packed symmetric
// matrices (and the slim vector layout of the panel engine).
MCD_HD Smem carve(const Problem& P, double* mat, unsigned char* vec, bool panels = false, bool packed = false) {
  Smem s;
  const int np = P.np;
  s.XT = mat;
  s.B2 = mat + (packed ? pk_elems(np) : (panels ? (size_t)P.nbk * P.lds : (size_t)np * P.ld));
  double* d = reinterpret_cast<double*>(vec);
  s.a = d; d += np;
  s.b = d; d += np;
  s.ea = d; d += np;
  s.eb = d; d += np;
  s.hv = d; d += np;
  s.lgc = d; d += 128;
  s.lgl = d; d += 128;
  s.lam = d; d += panels ? 0 : np;
  s.el = d; d += panels ? 0 : np;
  s.v = d; d += np;
  s.w = d; d += np;
  s.vt = d; d += np;
  s.wt = d; d += np;
  s.vc = d; d += kMaxCoef;
  s.wc = d; d += kMaxCoef;
  s.ct = d; d += kMaxCoef;
  s.lag = d; d += kMaxLag;
  s.lag_try = d; d += kMaxLag;
  s.rc = d; d += panels ? 0 : np / 2;
  s.rs = d; d += panels ? 0 : np / 2;
  s.red = d; d += 32;
  s.part = d; d += kThreads;
  s.text = panels ? nullptr : d; d += panels ? 0 : 6 * (np + 2 * kTaylorMargin);
  s.sc = reinterpret_cast<Scalars*>(d);
  d += (sizeof(Scalars) + 7) / 8;
  int* ip = reinterpret_cast<int*>((reinterpret_cast<uintptr_t>(d) + 15) & ~(uintptr_t)15);   // wti: 128-bit loads
  s.wti = ip; ip += 16 * kMaxTasks;
  s.rp = ip; ip += panels ? 0 : np / 2;
  s.rq = ip; ip += panels ? 0 : np / 2;
  unsigned char* c = reinterpret_cast<unsigned char*>(ip);
  const int nb = panels ? 0 : np / 4, h = panels ? 0 : np / 2;
  s.ntiles = nb * (nb + 1) / 2;
  s.nblk = h * (h - 1) / 2;
  s.tiles = c; c += 2 * s.ntiles;
  s.jblocks = c; c += 2 * s.nblk;
  s.wtab = c; c += 16 * kMaxTasks;
  s.hrow = c; c += kThreads;
  s.hrem = c; c += kThreads;
  s.tord = c;
  return s;
}

// ---------------------------------------------------------------------------------------
// kMode: 0 = symmetric engines (eigen / squaring in shared memory), 1 = blocked squaring engine
// (104 < n <= 208), 2 = general engine for the non-reversible --pull generator.  The mode is a template
// parameter so that every kernel only carries the code of its own likelihood evaluation.
// kCluster: the chain is shared by the two CTAs of a thread-block cluster (blocked engine only, see blk_product).
template <class Exec, int kMode = 0, bool kCluster = false>
struct Chain {
  static constexpr bool kBlocked = (kMode == 1);
  static constexpr bool kGeneral = (kMode == 2);
  static constexpr bool kPacked = (kMode == 3);
  static constexpr bool kTensorOnly = (kMode == 4);   // mode 0 without the eigen engine and the DFMA products: the production kernel
  static constexpr bool kPackedDst = (kMode == 1 || kMode == 3);   // the Taylor polynomial is written in packed tile storage
  Exec& ex;
  const Problem& P;
  // optional absorbing rates subtracted from the diagonal of the generator (mcd_propagator_batch: the cut rate matrices
  // of lib/utils.py:56-74); null everywhere on the likelihood / MC path
  const double* sink_ = nullptr;
  Smem s;

  MCD_HD Chain(Exec& e, const Problem& p, const Smem& sm) : ex(e), P(p), s(sm) {}

  // Tile list: the upper triangle of the (np/4)^2 tile grid, enumerated patch by patch
  // (4 x 8 tiles) so that the 32 lanes of a warp share operand rows (shared-memory
  // broadcast).  Jacobi block list: all pair-slot pairs ka < kb.
  MCD_HD void build_tables() {
    ex.par([&](int tid) {
      if (tid < 128) {
        const double c = 1.0 / (1.0 + ((double)tid + 0.5) / 128.0);
        s.lgc[tid] = c;
        s.lgl[tid] = -log(c);
      }
      // Horner passes of the squaring engine: thread -> (band row, slot); runtime integer divisions cost
      // ~75 dependent cycles each, so they are done once here
      const int tpr = (Exec::nthr() / P.n > 1) ? Exec::nthr() / P.n : 1;
      const int row = tid / tpr;
      s.hrow[tid] = (unsigned char)((row < P.n && row < 255) ? row : 255);
      s.hrem[tid] = (unsigned char)(tid - row * tpr);
      if (tid < 32) {
        s.tord[tid] = (unsigned char)taylor_order(tid, 0);
        s.tord[32 + tid] = (unsigned char)taylor_order_hi(tid);
      }
    });
    ex.single([&]() {
      // 4x4 tile table: DFMA products, and buffer_loglike of every full-storage engine (the tensor-only kernel needs it for
      // duplicate lag times and for kernels that need no squaring)
      const int nb = (kBlocked || kPacked) ? 0 : P.np / 4;
      int t = 0;
      for (int pi = 0; pi < nb; pi += 4)
        for (int pj = (pi / 8) * 8; pj < nb; pj += 8)
          for (int ib = pi; ib < pi + 4 && ib < nb; ++ib)
            for (int jb = pj; jb < pj + 8 && jb < nb; ++jb)
              if (ib <= jb) {
                s.tiles[2 * t] = (unsigned char)ib;
                s.tiles[2 * t + 1] = (unsigned char)jb;
                ++t;
              }
      // Tensor-core warp tasks: the upper triangle of the (np/8)^2 grid of 8x8 tiles is cut into
      // row segments of at most 8 tiles; a warp owns one full segment or two partial ones
      // (first-fit decreasing), i.e. at most 8 tiles from at most 2 tile rows, so that one or
      // two A fragments serve all of its MMAs.  wtab[w] = {nt, na, rowA, rowB, jb[8]}: tiles
      // t < na belong to rowA, na <= t < nt to rowB, t >= nt are padding (computed, not used).
      {
        const int nb8 = P.nbk / 8, cap = (kThreads >= 512) ? 6 : 8, maxw = kMaxTasks;
        unsigned char sib[64], sjb[64], sln[64], used[64];
        int nseg = 0;
        for (int ib = 0; ib < nb8; ++ib)
          for (int jb = ib; jb < nb8; jb += cap) {
            sib[nseg] = (unsigned char)ib; sjb[nseg] = (unsigned char)jb;
            sln[nseg] = (unsigned char)((nb8 - jb < cap) ? nb8 - jb : cap);
            used[nseg] = 0; ++nseg;
          }
        int nw = 0;
        bool ok = true;
        auto put = [&](int w, int sg, bool first) {
          unsigned char* wt = s.wtab + 16 * w;
          if (first) { wt[0] = 0; wt[1] = sln[sg]; wt[2] = sib[sg]; wt[3] = sib[sg]; }
          else wt[3] = sib[sg];
          for (int t = 0; t < sln[sg]; ++t) wt[4 + wt[0] + t] = (unsigned char)(sjb[sg] + t);
          wt[0] = (unsigned char)(wt[0] + sln[sg]);
        };
        for (int sg = 0; sg < nseg; ++sg)
          if (sln[sg] == cap) {
            if (nw < maxw) put(nw, sg, true); else ok = false;
            used[sg] = 1; ++nw;
          }
        const int first_partial = nw;
        int nsegs_in[kMaxTasks];
        for (int w = 0; w < maxw; ++w) nsegs_in[w] = 0;
        for (;;) {
          int best = -1;
          for (int sg = 0; sg < nseg; ++sg)
            if (!used[sg] && (best < 0 || sln[sg] > sln[best])) best = sg;
          if (best < 0) break;
          used[best] = 1;
          int w = -1;
          for (int c = first_partial; c < nw && c < maxw; ++c)
            if (nsegs_in[c] == 1 && s.wtab[16 * c] + sln[best] <= cap) { w = c; break; }
          if (w < 0) {
            w = nw++;
            if (w < maxw) { put(w, best, true); nsegs_in[w] = 1; } else ok = false;
          } else { put(w, best, false); nsegs_in[w] = 2; }
        }
        for (int w = 0; w < nw && w < maxw; ++w) {
          unsigned char* wt = s.wtab + 16 * w;
          for (int t = wt[0]; t < cap; ++t) wt[4 + t] = wt[4];
        }
        s.sc->mma_tasks = ok ? nw : 0;
        for (int w = 0; w < maxw; ++w) {
          const unsigned char* wt = s.wtab + 16 * w;
          int* q = s.wti + 16 * w;
          const bool have = ok && w < nw;
          for (int k = 0; k < 16; ++k) q[k] = (have && k < 12) ? wt[k] : 0;
        }
      }
      const int h = (kBlocked || kPacked || kTensorOnly) ? 0 : P.np / 2;
      int q = 0;
      for (int ka = 0; ka < h; ++ka)
        for (int kb = ka + 1; kb < h; ++kb) {
          s.jblocks[2 * q] = (unsigned char)ka;
          s.jblocks[2 * q + 1] = (unsigned char)kb;
          ++q;
        }
    });
  }

  // Symmetrised generator of lib/utils.py:77-131: diagonal a (= R[i,i]), coupling
  // b[i] = exp(w_i) between bin i and i+1 (b[n-1] = periodic corner), and the similarity
  // scalings ea = exp(-v/2), eb = exp(+v/2).
  MCD_HD void build_S(const double* v, const double* w) {
    const int n = P.n, np = P.np, pbc = P.pbc, dim_w = P.dim_w;
    ex.par([&](int tid) {
      for (int i = tid; i < np; i += Exec::nthr()) {
        if (i < n) {
          const double vi = v[i];
          double ru = 0.0, rd = 0.0;
          if (i < n - 1) ru = exp(w[i] - 0.5 * (v[i + 1] - vi));
          else if (pbc) ru = exp(w[n - 1] - 0.5 * (v[0] - vi));
          if (i > 0) rd = exp(w[i - 1] + 0.5 * (vi - v[i - 1]));
          else if (pbc) rd = exp(w[n - 1] + 0.5 * (vi - v[n - 1]));
          s.a[i] = sink_ ? -rd - ru - sink_[i] : -rd - ru;
          s.b[i] = (i < dim_w) ? exp(w[i]) : 0.0;
          s.ea[i] = exp(-0.5 * vi);
          s.eb[i] = exp(0.5 * vi);
          s.hv[i] = 0.5 * vi;
        } else {
          s.a[i] = 0.0; s.b[i] = 0.0; s.ea[i] = 0.0; s.eb[i] = 0.0; s.hv[i] = 0.0;
        }
      }
    });
    // one fused reduction (two barriers instead of ten): ||diag||_max (Jacobi scale), max / min of v/2
    // (g >= gmin guarantees P_ij = exp(-v_i/2) g exp(v_j/2) >= clip for every pair), the Gershgorin bound
    // of ||S|| (squaring engines), and sum_ij N_ij (v_j - v_i)/2 = sum_k (v_k/2) (colsum_k - rowsum_k).
    // s.part is free here (the likelihood partial sums are zeroed after build_S).
    const r5 red = ex.max4sum(s.part, [&](int tid) {
      r5 x{0.0, -1.0e308, -1.0e308, 0.0, 0.0};
      for (int i = tid; i < n; i += Exec::nthr()) {
        x.m0 = fmax(x.m0, fabs(s.a[i]));
        x.m1 = fmax(x.m1, s.hv[i]);
        x.m2 = fmax(x.m2, -s.hv[i]);
        const int im = (i == 0) ? n - 1 : i - 1;
        const double r = fabs(s.a[i]) + s.b[im] + s.b[i];
        x.m3 = (r > x.m3 || r != r) ? r : x.m3;
        if (P.rvec) x.sum += s.hv[i] * P.rvec[i];
      }
      return x;
    });
    const double sc = red.m0, hmax = red.m1, hmin = -red.m2, lin = P.rvec ? red.sum : 0.0;
    ex.single([&]() {
      s.sc->gersh = red.m3;
      s.sc->scale = sc;
      s.sc->gmin = P.clip * exp(hmax - hmin);
      s.sc->lin = lin;
    });
  }

  // Cold start: XT = identity, B2 = dense S.
  MCD_HD void cold_init() {
    const int n = P.n, np = P.np, ld = P.ld;
    ex.par([&](int tid) {
      for (int e = tid; e < np * np; e += Exec::nthr()) {
        const int i = e / np, j = e - i * np;
        s.XT[i * ld + j] = (i == j) ? 1.0 : 0.0;
        double val = 0.0;
        if (i < n && j < n) {
          if (i == j) val = s.a[i];
          else if (j == i + 1) val = s.b[i];
          else if (i == j + 1) val = s.b[j];
          else if (P.pbc && ((i == 0 && j == n - 1) || (j == 0 && i == n - 1))) val = s.b[n - 1];
        }
        s.B2[i * ld + j] = val;
      }
    });
  }

  // Warm start: B2 <- XT S XT^T with the eigenvectors of the previous trial matrix.
  MCD_HD void warm_form() {
    const int n = P.n, np = P.np, ld = P.ld, nb = np / 4;
    // (1) B2[j][r] = (S x_j)[r]
    ex.par([&](int tid) {
      for (int e = tid; e < np * np; e += Exec::nthr()) {
        const int j = e / np, r = e - j * np;
        double val = 0.0;
        if (r < n) {
          const double* x = s.XT + j * ld;
          const int rm = (r == 0) ? n - 1 : r - 1;
          const int rp = (r == n - 1) ? 0 : r + 1;
          val = s.a[r] * x[r] + s.b[rm] * x[rm] + s.b[r] * x[rp];
        }
        s.B2[j * ld + r] = val;
      }
    });
    // (2) A'_ij = sum_r XT[i][r] B2[j][r] for tile rows {ib + nb a} x {jb + nb b}
    ex.par([&](int tid) {
      double* acc = ex.regs(tid);
      if (tid < s.ntiles) {
        const int ib = s.tiles[2 * tid], jb = s.tiles[2 * tid + 1];
        double c[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) c[a][b] = 0.0;
        const double* x0 = s.XT + ib * ld;
        const double* y0 = s.B2 + jb * ld;
        const int sx = nb * ld;
        for (int r = 0; r < np; r += 2) {
          d2 xa[4], yb[4];
#pragma unroll
          for (int a = 0; a < 4; ++a) xa[a] = ld2(x0 + a * sx + r);
#pragma unroll
          for (int b = 0; b < 4; ++b) yb[b] = ld2(y0 + b * sx + r);
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              c[a][b] = fma(xa[a].x, yb[b].x, c[a][b]);
              c[a][b] = fma(xa[a].y, yb[b].y, c[a][b]);
            }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a * 4 + b] = c[a][b];
      }
    });
    // (3) mirror-store (symmetric by construction)
    ex.par([&](int tid) {
      const double* acc = ex.regs(tid);
      if (tid < s.ntiles) {
        const int ib = s.tiles[2 * tid], jb = s.tiles[2 * tid + 1];
        for (int a = 0; a < 4; ++a)
          for (int b = 0; b < 4; ++b) {
            const int i = ib + nb * a, j = jb + nb * b;
            if (ib == jb && a > b) continue;  // keep one representative inside diagonal tiles
            const double val = acc[a * 4 + b];
            s.B2[i * ld + j] = val;
            s.B2[j * ld + i] = val;
          }
      }
    });
  }

  // Parallel cyclic Jacobi (round-robin ordering) on the symmetric matrix in B2,
  // accumulating the rotations into the rows of XT.  Returns the number of sweeps, or -1
  // when the off-diagonal part did not vanish (NaN input or kMaxSweeps exceeded).
  MCD_HD int jacobi() {
    const int np = P.np, ld = P.ld, h = np / 2, m = np - 1;
    const double scale = s.sc->scale;
    const double tol = 2.0e-15 * scale;
    const double skip = 1.0e-18 * scale;
    int sweeps = 0;
    for (;;) {
      if (!(scale < 1.0e300)) { sweeps = -1; break; }  // non-finite profile (exp overflow / NaN)
      double off = ex.max([&](int tid) {
        double mx = 0.0;
        for (int e = tid; e < np * np; e += Exec::nthr()) {
          const int i = e / np, j = e - i * np;
          if (j > i) {
            const double t = fabs(s.B2[i * ld + j]);
            mx = (t > mx || t != t) ? t : mx;  // propagate NaN
          }
        }
        return mx;
      });
      if (off != off) { sweeps = -1; break; }
      if (off <= tol) break;
      if (sweeps >= kMaxSweeps) { sweeps = -1; break; }
      for (int t = 0; t < m; ++t) {
        // J1: rotation of every pair of this round, from the 2x2 diagonal blocks
        ex.par([&](int tid) {
          if (tid < h) {
            int p, q;
            if (tid == 0) { p = m; q = t; }
            else { p = (t + tid) % m; q = (t - tid + m) % m; }
            const double apq = s.B2[p * ld + q];
            double c = 1.0, sn = 0.0;
            if (fabs(apq) > skip) {
              const double app = s.B2[p * ld + p], aqq = s.B2[q * ld + q];
              const double theta = (aqq - app) / (2.0 * apq);
              const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
              c = 1.0 / sqrt(tt * tt + 1.0);
              sn = tt * c;
              s.B2[p * ld + p] = app - tt * apq;
              s.B2[q * ld + q] = aqq + tt * apq;
              s.B2[p * ld + q] = 0.0;
              s.B2[q * ld + p] = 0.0;
            }
            s.rc[tid] = c; s.rs[tid] = sn; s.rp[tid] = p; s.rq[tid] = q;
          }
        });
        // J2: two-sided update of every off-diagonal 2x2 block + rotation of the XT rows
        ex.par([&](int tid) {
          for (int e = tid; e < s.nblk; e += Exec::nthr()) {
            const int ka = s.jblocks[2 * e], kb = s.jblocks[2 * e + 1];
            const double sa = s.rs[ka], sb = s.rs[kb];
            if (sa == 0.0 && sb == 0.0) continue;
            const double ca = s.rc[ka], cb = s.rc[kb];
            const int pa = s.rp[ka], qa = s.rq[ka], pb = s.rp[kb], qb = s.rq[kb];
            const double m00 = s.B2[pa * ld + pb], m01 = s.B2[pa * ld + qb];
            const double m10 = s.B2[qa * ld + pb], m11 = s.B2[qa * ld + qb];
            const double r00 = ca * m00 - sa * m10, r01 = ca * m01 - sa * m11;
            const double r10 = sa * m00 + ca * m10, r11 = sa * m01 + ca * m11;
            const double n00 = cb * r00 - sb * r01, n01 = sb * r00 + cb * r01;
            const double n10 = cb * r10 - sb * r11, n11 = sb * r10 + cb * r11;
            s.B2[pa * ld + pb] = n00; s.B2[pb * ld + pa] = n00;
            s.B2[pa * ld + qb] = n01; s.B2[qb * ld + pa] = n01;
            s.B2[qa * ld + pb] = n10; s.B2[pb * ld + qa] = n10;
            s.B2[qa * ld + qb] = n11; s.B2[qb * ld + qa] = n11;
          }
          const int half = np / 2;  // XT rows in chunks of 2 columns
          for (int e = tid; e < h * half; e += Exec::nthr()) {
            const int k = e / half, r = (e - k * half) * 2;
            const double sn = s.rs[k];
            if (sn == 0.0) continue;
            const double c = s.rc[k];
            double* xp = s.XT + s.rp[k] * ld + r;
            double* xq = s.XT + s.rq[k] * ld + r;
            const d2 x = ld2(xp), y = ld2(xq);
            st2(xp, c * x.x - sn * y.x, c * x.y - sn * y.y);
            st2(xq, sn * x.x + c * y.x, sn * x.y + c * y.y);
          }
        });
      }
      ++sweeps;
    }
    ex.par([&](int tid) {
      for (int i = tid; i < np; i += Exec::nthr()) s.lam[i] = s.B2[i * ld + i];
    });
    return sweeps;
  }

  // c = tile (ib, jb) of X^T Y for row-major n x n operands (leading dimension ld); with
  // symmetric X this is the tile of X Y.  Register-blocked 4x4, operands read as 128-bit rows.
  MCD_HD void tile_product(const double* X, const double* Y, int ib, int jb, double (&c)[4][4]) {
    const int np = P.np, ld = P.ld;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) c[a][b] = 0.0;
    const double* xi = X + ib * 4;
    const double* yj = Y + jb * 4;
#pragma unroll 2
    for (int k = 0; k < np; ++k) {
      const d2 x01 = ld2(xi + k * ld), x23 = ld2(xi + k * ld + 2);
      const d2 y01 = ld2(yj + k * ld), y23 = ld2(yj + k * ld + 2);
      const double xa[4] = {x01.x, x01.y, x23.x, x23.y};
      const double yb[4] = {y01.x, y01.y, y23.x, y23.y};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) c[a][b] = fma(xa[a], yb[b], c[a][b]);
    }
  }

  // Natural logarithm for positive, finite, normal arguments (the likelihood epilogue only calls it
  // with g >= gmin >= clip > 0).  Branch-free: g = 2^e m, m in [1,2); table of 128 reciprocals c_k
  // brings r = m c_k - 1 into |r| < 2^-8; log g = e ln2 - log c_k + (r - r^2/2 + ... - r^6/6).
  // Within 4e-16 + 1 ulp of libm (checked in tests/test_emulation.py).
  MCD_HD double fast_log(double g) const {
#if defined(__CUDA_ARCH__)
    const long long bits = __double_as_longlong(g);
#else
    long long bits;
    std::memcpy(&bits, &g, 8);
#endif
    const int e = (int)((bits >> 52) & 0x7ff) - 1023;
    const int k = (int)((bits >> 45) & 127);
    const long long mb = (bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL;
#if defined(__CUDA_ARCH__)
    const double m = __longlong_as_double(mb);
#else
    double m;
    std::memcpy(&m, &mb, 8);
#endif
    const double r = fma(m, s.lgc[k], -1.0);
    double p = fma(r, -1.0 / 6.0, 0.2);
    p = fma(r, p, -0.25);
    p = fma(r, p, 1.0 / 3.0);
    p = fma(r, p, -0.5);
    p = fma(r * r, p, r);
    return fma((double)e, 0.6931471805599453094, s.lgl[k] + p);
  }

  // Likelihood terms of the unordered pair {i, j} of the symmetric kernel entry g = G_ij:
  //   N_ij log P_ij + N_ji log P_ji,   P_ij = exp(-v_i/2) g exp(+v_j/2),  P_ji = exp(-v_j/2) g exp(+v_i/2)
  // = (N_ij + N_ji) log g + (N_ij - N_ji)(v_j - v_i)/2.
  // The second part is linear in v and summed once per evaluation from the precomputed row/column
  // sums (build_S: sc->lin), so the pair costs ONE logarithm weighted by nsym = N_ij + N_ji.
  // Entries that fall below the clip of lib/utils.py:317 take the exact per-entry route.
  MCD_HD double pair_term(double g, int ns, int i, int j, const int* cnt) const {
    if (ns == 0) return 0.0;
    if (g >= s.sc->gmin) return (double)ns * fast_log(g);
    return pair_term_clipped(g, ns, i, j, cnt, P.np, P.clip, s.ea, s.eb, s.hv);
  }
  // Rare branch (an entry below the clip), kept out of line and free of `this`: the epilogues are
  // unrolled 16-32 times per thread and the kernel is instruction-cache sensitive.
  static MCD_NOINLINE double pair_term_clipped(double g, int ns, int i, int j, const int* cnt, int np, double clip,
                                               const double* ea, const double* eb, const double* hv) {
    if (i == j) return (double)ns * log(fmax(g, clip));
    const int nij = cnt[(size_t)i * np + j], nji = cnt[(size_t)j * np + i];
    const double pij = ea[i] * g * eb[j], pji = ea[j] * g * eb[i];
    double part = -(double)(nij - nji) * (hv[j] - hv[i]);   // undo this pair's share of sc->lin
    if (nij != 0) part += (double)nij * log(fmax(pij, clip));
    if (nji != 0) part += (double)nji * log(fmax(pji, clip));
    return part;
  }

  // Likelihood contribution of one 4x4 tile of the symmetric kernel G (upper triangle of the tile
  // grid; inside diagonal tiles only i <= j).
  MCD_HD double tile_loglike(const double (&c)[4][4], int ib, int jb, const i4 (&ns)[4], const int* cnt, double* Pl) {
    const int n = P.n;
    const int i0 = ib * 4, j0 = jb * 4;
    double part = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = i0 + a;
      const int nrow[4] = {ns[a].x, ns[a].y, ns[a].z, ns[a].w};
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int j = j0 + b;
        if (ib == jb && j < i) continue;
        part += pair_term(c[a][b], nrow[b], i, j, cnt);
        if (Pl && i < n && j < n) {
          Pl[(size_t)i * n + j] = s.ea[i] * c[a][b] * s.eb[j];
          Pl[(size_t)j * n + i] = s.ea[j] * c[a][b] * s.eb[i];
        }
      }
    }
    return part;
  }

  MCD_HD void load_counts(const int* nsym, int ib, int jb, i4 (&ns)[4]) {
    const int np = P.np;
#pragma unroll
    for (int a = 0; a < 4; ++a) ns[a] = ld4i(nsym + (size_t)(ib * 4 + a) * np + jb * 4);
  }

  // log L = sum_lag sum_ij N[lag,i,j] log(max(P_ij, clip)),  P = Pi^1/2 XT^T e^{t lam} XT Pi^-1/2
  // (lib/utils.py:307-368).  B2 is overwritten with the scaled eigenvectors of each lag.
  // When Pout != nullptr the propagators are also written ([nlag][n][n], row-major).
  MCD_HD double loglike(const double* lagt, double* Pout) {
    const int n = P.n, np = P.np, ld = P.ld;
    ex.par([&](int tid) { ex.regs(tid)[0] = 0.0; });
    for (int l = 0; l < P.nlag; ++l) {
      const double t = lagt[l];
      ex.par([&](int tid) {
        for (int k = tid; k < np; k += Exec::nthr()) s.el[k] = exp(t * s.lam[k]);
      });
      ex.par([&](int tid) {
        const int half = np / 2;
        for (int e = tid; e < np * half; e += Exec::nthr()) {
          const int k = e / half, r = (e - k * half) * 2;
          const d2 x = ld2(s.XT + k * ld + r);
          const double f = s.el[k];
          st2(s.B2 + k * ld + r, x.x * f, x.y * f);
        }
      });
      const int* cnt = P.counts + (size_t)l * np * np;
      const int* nsy = P.nsym + (size_t)l * np * np;
      double* Pl = Pout ? Pout + (size_t)l * n * n : nullptr;
      ex.par([&](int tid) {
        double part = 0.0;
        for (int tt = tid; tt < s.ntiles; tt += Exec::nthr()) {
          const int ib = s.tiles[2 * tt], jb = s.tiles[2 * tt + 1];
          i4 ns[4];
          load_counts(nsy, ib, jb, ns);  // before the k loop: hides the L2 latency
          double c[4][4];
          tile_product(s.XT, s.B2, ib, jb, c);
          part += tile_loglike(c, ib, jb, ns, cnt, Pl);
        }
        ex.regs(tid)[0] += part;
      });
    }
    return s.sc->lin + ex.sum([&](int tid) { return ex.regs(tid)[0]; });
  }

  // ------------------------------------------------------------------------------------
  // Propagators without an eigensolver: symmetric scaling-and-squaring.
  //   G(t0) = exp(t0 S) = (T_m(h S))^(2^s),  h = t0 / 2^s,  ||h S||_inf <= 1,
  //   G((l+1) t0) = G(l t0) G(t0)            (lag times in arithmetic progression)
  // exp(hS) >= 0 entrywise (S is a symmetrised generator), so every product is a sum of
  // non-negative terms and keeps componentwise relative accuracy.  All products are
  // symmetric: only the upper-triangular tiles are computed, register-staged and written
  // back in place (with their mirror) after a barrier -- two n x n buffers suffice.
  // Same quantity as lib/utils.py:313 (scipy.linalg.expm is also scaling-and-squaring).
  // ------------------------------------------------------------------------------------
  MCD_HD void store_tile_mirror(double* D, int ib, int jb, const double* r) {
    const int ld = P.ld;
    const int i0 = ib * 4, j0 = jb * 4;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      st2(D + (i0 + a) * ld + j0, r[a * 4 + 0], r[a * 4 + 1]);
      st2(D + (i0 + a) * ld + j0 + 2, r[a * 4 + 2], r[a * 4 + 3]);
    }
    if (ib != jb) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        st2(D + (j0 + b) * ld + i0, r[0 * 4 + b], r[1 * 4 + b]);
        st2(D + (j0 + b) * ld + i0 + 2, r[2 * 4 + b], r[3 * 4 + b]);
      }
    }
  }

  // D <- X Y (symmetric, commuting operands; D may alias X or Y).  If cnt != nullptr the
  // product is the kernel of a lag time and its likelihood terms are added to regs[0]... the
  // partial sum lives in `part_acc` (per-thread, caller-owned).
  // store: 0 = result not kept, 1 = D is a third buffer (direct store, no extra barrier phase),
  //        2 = D aliases an operand (register-staged, written after the barrier)
  MCD_HD void sym_product(const double* X, const double* Y, double* D, const int* cnt, const int* nsy, double* Pl,
                          int store) {
    ex.par([&](int tid) {
      double* r = ex.regs(tid);
      if (tid < s.ntiles) {
        const int ib = s.tiles[2 * tid], jb = s.tiles[2 * tid + 1];
        i4 ns[4];
        if (cnt) load_counts(nsy, ib, jb, ns);
        double c[4][4];
        tile_product(X, Y, ib, jb, c);
        if (cnt) s.part[tid] += tile_loglike(c, ib, jb, ns, cnt, Pl);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) r[a * 4 + b] = c[a][b];
        if (store == 1) store_tile_mirror(D, ib, jb, r);
      }
    });
    if (store == 2) {
      ex.par([&](int tid) {
        if (tid < s.ntiles) store_tile_mirror(D, s.tiles[2 * tid], s.tiles[2 * tid + 1], ex.regs(tid));
      });
    }
  }

  // ---- FP64 tensor-core variant of sym_product (mma.sync.m8n8k4.f64) ----------------------
  // ncu on the DFMA version: LSU shared wavefronts 74 % busy, FP64 pipe 29 % -- a 4x4 register
  // tile needs 0.5 operand doubles per FMA from shared memory.  An m8n8k4 fragment shares its
  // operands across the 32 lanes: 2 doubles per lane feed 8 FMAs per lane.  Each warp owns up to
  // 8 consecutive upper-triangular 8x8 tiles (row-major, so the A fragment is mostly reused) and
  // keeps their 2-doubles-per-lane accumulators in registers; results are register-staged and
  // written back in place after the barrier, like the DFMA version.
  static constexpr int kTilesPerWarp = (kThreads >= 512) ? 6 : 8;
  // the engines that do a product in ONE round need a warp per task (the packed engine runs in rounds)
  MCD_HD bool dmma_ok() const { return s.sc->mma_tasks > 0 && s.sc->mma_tasks <= Exec::nthr() / 32; }

#if defined(__CUDA_ARCH__)
  static __device__ __forceinline__ double lds64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
  }
  // k loop of one warp task: NA tiles use the A fragment of rowA, the others that of rowB.
  template <int NA>
  static __device__ __forceinline__ void mma_loop(unsigned ax, unsigned ay, int pitch4, int nk4, int offA0, int offA1,
                                                  const int (&offB)[kTilesPerWarp], double (&c)[kTilesPerWarp][2]) {
#pragma unroll 2
    for (int it = 0; it < nk4; ++it) {
      const double a0 = lds64(ax + offA0);
      const double a1 = (NA < kTilesPerWarp) ? lds64(ax + offA1) : 0.0;
#pragma unroll
      for (int t = 0; t < kTilesPerWarp; ++t) {
        const double b = lds64(ay + offB[t]);
        const double a = (t < NA) ? a0 : a1;
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[t][0]), "+d"(c[t][1])
                     : "d"(a), "d"(b));
      }
      ax += pitch4;
      ay += pitch4;
    }
  }
#endif

  // decoded warp task (registers): tiles t < na on tile row rowA, na <= t < nt on rowB, jb[t] = tile column
  struct WarpTask {
    int nt, na, rowA, rowB, jb[8];
  };
  MCD_HD WarpTask load_task(int warp, int ntasks) const {
    WarpTask k;
    const i4 h0 = lds4i(s.wti + 16 * warp), h1 = lds4i(s.wti + 16 * warp + 4), h2 = lds4i(s.wti + 16 * warp + 8);
    k.nt = (warp < ntasks) ? h0.x : 0; k.na = h0.y; k.rowA = h0.z; k.rowB = h0.w;
    k.jb[0] = h1.x; k.jb[1] = h1.y; k.jb[2] = h1.z; k.jb[3] = h1.w;
    k.jb[4] = h2.x; k.jb[5] = h2.y; k.jb[6] = h2.z; k.jb[7] = h2.w;
    return k;
  }

  MCD_HD void dmma_store(double* D, const WarpTask& k, int lane, const double* r) {
    const int ld = P.ld;
    const int fr = lane >> 2, fk = lane & 3;
#pragma unroll
    for (int t = 0; t < kTilesPerWarp; ++t) {
      if (t < k.nt) {
        const int ib = (t < k.na) ? k.rowA : k.rowB, jb = k.jb[t];
        const int i = ib * 8 + fr, j = jb * 8 + 2 * fk;
        st2(D + i * ld + j, r[2 * t], r[2 * t + 1]);
        if (ib != jb) {
          D[j * ld + i] = r[2 * t];
          D[(j + 1) * ld + i] = r[2 * t + 1];
        }
      }
    }
  }

  MCD_HD void dmma_product(const double* X, const double* Y, double* D, const int* cnt, const int* nsy, double* Pl,
                           int store) {
    const int n = P.n, np = P.np, ld = P.ld;
    const int ntasks = s.sc->mma_tasks;
    MCD_CLK(17);
    ex.par([&](int tid) {
      const int warp = tid >> 5, lane = tid & 31;
      const WarpTask task = load_task(warp, ntasks);
      const bool active = warp < ntasks;
      const int nt = task.nt, na = task.na, rowA = task.rowA, rowB = task.rowB;
      MCD_CLK(18);
      double c[kTilesPerWarp][2];
#pragma unroll
      for (int t = 0; t < kTilesPerWarp; ++t) { c[t][0] = 0.0; c[t][1] = 0.0; }
      const int fr = lane >> 2, fk = lane & 3;   // fragment row (m or n) and k index of this lane
      // symmetric counts of the two entries per tile this lane owns, fetched before the k loop
      // (padding tiles t >= nt repeat a valid tile of the task: loaded without a select so that the
      // loads stay in flight across the k loop instead of being consumed right here)
      int nsv[kTilesPerWarp][2];
      if (cnt && active) {
#pragma unroll
        for (int t = 0; t < kTilesPerWarp; ++t) {
          const int ib = (t < na) ? rowA : rowB, jb = task.jb[t];
          const i2 q = ld2i(nsy + (size_t)(ib * 8 + fr) * np + jb * 8 + 2 * fk);
          nsv[t][0] = q.x;
          nsv[t][1] = q.y;
        }
      }
      MCD_CLK(5);
      if (active) {
#if defined(__CUDA_ARCH__)
        int offB[kTilesPerWarp];
#pragma unroll
        for (int t = 0; t < kTilesPerWarp; ++t) offB[t] = task.jb[t] * 64;
        const unsigned ax = (unsigned)__cvta_generic_to_shared(X) + (unsigned)(fk * ld + fr) * 8u;
        const unsigned ay = (unsigned)__cvta_generic_to_shared(Y) + (unsigned)(fk * ld + fr) * 8u;
        // rows k >= n of both operands are padding (zero): the k loop stops at the last group of four that holds a
        // real row -- 25 instead of 26 steps at 100 bins
        const int pitch4 = 4 * ld * 8, nk4 = (n + 3) >> 2;
        switch (na) {
          case 1: mma_loop<1>(ax, ay, pitch4, nk4, rowA * 64, rowB * 64, offB, c); break;
          case 2: mma_loop<2>(ax, ay, pitch4, nk4, rowA * 64, rowB * 64, offB, c); break;
          case 3: mma_loop<3>(ax, ay, pitch4, nk4, rowA * 64, rowB * 64, offB, c); break;
          case 4: mma_loop<4>(ax, ay, pitch4, nk4, rowA * 64, rowB * 64, offB, c); break;
          case 5: mma_loop<5>(ax, ay, pitch4, nk4, rowA * 64, rowB * 64, offB, c); break;
          case 6: mma_loop<6>(ax, ay, pitch4, nk4, rowA * 64, rowB * 64, offB, c); break;
          case 7: mma_loop<7>(ax, ay, pitch4, nk4, rowA * 64, rowB * 64, offB, c); break;
          default: mma_loop<8>(ax, ay, pitch4, nk4, rowA * 64, rowB * 64, offB, c); break;
        }
#else
        for (int t = 0; t < nt; ++t) {
          const int ib = (t < na) ? rowA : rowB, jb = task.jb[t];
          const int i = ib * 8 + fr;
          for (int e = 0; e < 2; ++e) {
            const int j = jb * 8 + 2 * fk + e;
            double acc = 0.0;
            for (int k = 0; k < np; ++k) acc = fma(X[k * ld + i], Y[k * ld + j], acc);
            c[t][e] = acc;
          }
        }
#endif
      }
      MCD_CLK(6);
      // likelihood epilogue: this lane owns G[i][j], G[i][j+1]; inside diagonal tiles only i <= j.
      // Branch-free over the 16 entries (weight 0 for padding / lower-triangle / empty pairs, argument 1
      // for entries below the clip) so that the 16 logarithms interleave instead of running as 16
      // dependent chains; the rare clipped entries are redone exactly afterwards.
      if (cnt && nt > 0) {
        const double gmin = s.sc->gmin;
        double part = 0.0;
        bool rare = false;
#pragma unroll
        for (int t = 0; t < kTilesPerWarp; ++t) {
          const int ib = (t < na) ? rowA : rowB, jb = task.jb[t];
          const int i = ib * 8 + fr, j = jb * 8 + 2 * fk;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const bool valid = (t < nt) && !(ib == jb && j + e < i);
            const double g = c[t][e];
            const bool above = g >= gmin;
            const int ns = valid ? nsv[t][e] : 0;
            rare |= (ns != 0) && !above;
            part += (double)(above ? ns : 0) * fast_log(above ? g : 1.0);
          }
        }
        if (rare) {
#pragma unroll
          for (int t = 0; t < kTilesPerWarp; ++t) {
            const int ib = (t < na) ? rowA : rowB, jb = task.jb[t];
            const int i = ib * 8 + fr, j = jb * 8 + 2 * fk;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const bool valid = (t < nt) && !(ib == jb && j + e < i);
              if (valid && nsv[t][e] != 0 && !(c[t][e] >= gmin))
                part += pair_term_clipped(c[t][e], nsv[t][e], i, j + e, cnt, P.np, P.clip, s.ea, s.eb, s.hv);
            }
          }
        }
        s.part[tid] += part;
        if (Pl) {
#pragma unroll
          for (int t = 0; t < kTilesPerWarp; ++t) {
            const int ib = (t < na) ? rowA : rowB, jb = task.jb[t];
            const int i = ib * 8 + fr, j = jb * 8 + 2 * fk;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int jj = j + e;
              if (t < nt && !(ib == jb && jj < i) && i < n && jj < n) {
                Pl[(size_t)i * n + jj] = s.ea[i] * c[t][e] * s.eb[jj];
                Pl[(size_t)jj * n + i] = s.ea[jj] * c[t][e] * s.eb[i];
              }
            }
          }
        }
      }
      MCD_CLK(7);
      double* r = ex.regs(tid);
#pragma unroll
      for (int t = 0; t < kTilesPerWarp; ++t) { r[2 * t] = c[t][0]; r[2 * t + 1] = c[t][1]; }
      if (store == 1) dmma_store(D, task, lane, r);
      MCD_CLK(8);
    });
    MCD_CLK(9);
    if (store == 2) {
      ex.par([&](int tid) { dmma_store(D, load_task(tid >> 5, ntasks), tid & 31, ex.regs(tid)); });
    }
    MCD_CLK(10);
  }

  // Taylor order of the squaring engines.  T_m(x) = e^x (1 - x^(m+1)/(m+1)! ...) per eigenmode, and the
  // error is amplified 2^s times by the squarings -- but only modes with t |lambda| <= 45 survive in
  // G(t) at all (e^-45 = 3e-20), i.e. |x| <= x* = min(1, 45 / 2^s).  Smallest m with
  // 2^s x*^(m+1)/(m+1)! <= 1e-17.
  //
  // Spectral shift.  The spectrum of the (symmetrised) generator lies in [-rho, 0], rho = Gershgorin bound, so
  // S + cI with c = rho/2 has spectral radius rho/2: exp(hS) = e^{-hc} exp(h(S + cI)) needs ONE squaring less
  // for the same argument bound, and the diagonal of h(S + cI) is mostly non-negative (less cancellation in
  // the polynomial).  The modes that survive in G(t) now sit at the TOP of the shifted spectrum, x ~ +theta,
  // where every Taylor term is positive: relative error theta^(m+1)/(m+1)! per application, amplified 2^s
  // times by the squarings.  Argument bound kTheta = 2 (one more squaring saved; a Taylor pass on the band costs
  // ~1 % of a dense product, see taylor_column).  Order: smallest m with 2^s theta^(m+1)/(m+1)! <= 1e-17,
  // capped at 30 (the register window of taylor_lane; reached only for s > 25, where the bound is still < 2e-16).
  static constexpr int kThetaLog2 = 1;
  static constexpr double kTheta = 2.0;
  static MCD_HD int taylor_order(int nsq, int taylor_m) {
    if (taylor_m > 0) return taylor_m < 30 ? taylor_m : 30;
    double term = ldexp(kTheta, nsq);   // 2^s theta^k / k!  for k = 1
    int m = 1;
    while (m < 30 && !(term * kTheta / (double)(m + 1) <= 1.0e-17)) {
      term = term * kTheta / (double)(m + 1);
      ++m;
    }
    return m < 6 ? 6 : m;
  }

  // One squaring less when the argument allows it: with s - 1 squarings the argument is twice as large, in (2, 4]; up to
  // kThetaHi = 2.8 the Taylor order that keeps 2^s theta^(m+1)/(m+1)! <= 1e-17 still fits the register window (m <= 30
  // for s <= 10), i.e. four more Taylor passes (~1 % of a step) instead of a dense product (~7 %) for the 48 % of the
  // arguments that fall in (2, 2.8].  255: the bound is not reached with m <= 30 (no stretch).
  static constexpr double kThetaHi = 2.8;
  static MCD_HD int taylor_order_hi(int nsq) {
    double term = ldexp(kThetaHi, nsq);
    int m = 1;
    while (m < 30 && !(term * kThetaHi / (double)(m + 1) <= 1.0e-17)) {
      term = term * kThetaHi / (double)(m + 1);
      ++m;
    }
    return (term * kThetaHi / (double)(m + 1) <= 1.0e-17) ? (m < 6 ? 6 : m) : 255;
  }

  // How exp(tS) is reached: T = e^{-hc} T_m(h(S + cI)), then nsq squarings.
  struct TaylorPlan {
    int nsq, m, ring;
    double h, shift, scale;
  };
  // gersh: bound of the spectral interval [-gersh, 0].  false: non-finite / absurd scale.
  MCD_HD bool taylor_plan(double t, double gersh, int taylor_m, TaylorPlan& tp) const {
    const double rho = t * gersh;
    if (!(rho < 1.0e9) || !(t > 0.0)) return false;
    const double y = ldexp(rho, -1 - kThetaLog2);      // (rho/2) / theta
    int nsq = 0;
    if (y > 1.0) {
      int ex2;
      const double fr2 = frexp(y, &ex2);               // y = fr2 2^ex2, fr2 in [0.5, 1)
      nsq = (fr2 == 0.5) ? ex2 - 1 : ex2;
    }
    int m_auto = s.tord[nsq];
    if (taylor_m <= 0 && nsq >= 1 && ldexp(y, 1 - nsq) * kTheta <= kThetaHi && s.tord[32 + nsq - 1] != 255) {
      --nsq;                                            // argument (rho/2) / 2^nsq in (kTheta, kThetaHi]
      m_auto = s.tord[32 + nsq];
    }
    tp.nsq = nsq;
    tp.h = ldexp(t, -nsq);
    tp.m = (taylor_m > 0) ? (taylor_m < 30 ? taylor_m : 30) : m_auto;
    tp.m += tp.m & 1;                                   // the passes are done in pairs (taylor_lane): a higher order only helps
    tp.shift = 0.5 * gersh;
    tp.scale = exp(-tp.h * tp.shift);
    tp.ring = (P.pbc && 2 * tp.m + 1 > P.n) ? 1 : 0;    // the band wraps onto itself: whole columns (n <= 62)
    return true;
  }

  // ---- Taylor polynomial on the band, column by column in registers ------------------------------
  // H <- I + (hS'/k) H, k = m..1 (Horner), S' = S + cI tridiagonal (+ periodic corner).  Column j of H only
  // couples to itself:  H'[i][j] = delta_ij + (a_i H[i][j] + lo_i H[i-1][j] + up_i H[i+1][j]) / k,  and after m
  // passes it is non-zero on rows j-m .. j+m.  A PAIR of adjacent lanes keeps the window j-31 .. j+32 in
  // registers -- the lower lane rows j, j-1, .. j-31, the upper lane rows j+1 .. j+32, both indexed by the
  // distance from the centre so that the two lanes run the same code -- and does ALL passes with one shuffle
  // per pass (the two centre elements) and no barrier or shared-memory store.  The pre-scaled coefficients
  // come from arrays extended by a wrap margin of kTM rows on each side (and a reversed copy for the lower
  // lane), so every load is [thread base + compile-time offset].  Chunks of 4 elements the band has not
  // reached yet are skipped (uniform branch).  (The first version kept H in shared memory and did one
  // barrier-closed pass per Taylor term: 12 % of the MC step at 100 bins.)
  // Periodic profiles whose band is wider than the ring (2m+1 > n, n <= 62): the same window is computed on
  // the periodic LINE (the covering space: the extended coefficient arrays are periodic) and the entries are
  // then summed over the rows that coincide on the ring -- in a fixed order, lower lane first.
  static constexpr int kTH = 32;              // window elements per lane
  static constexpr int kTChunk = 4;           // elements per skip-able chunk of a sweep (8: same time, more spills)
  static constexpr int kTM = kTaylorMargin;   // >= kTH + 1
  MCD_HOSTDEV static size_t taylor_ext_doubles(int n) { return (size_t)6 * (n + 2 * kTM); }

  // pre-scaled coefficients (all threads; the caller closes the phase).  Row i: diagonal h (diag[i] + c),
  // coupling to row i+1: h up[i], to row i-1: h lo[i-1].  ext: taylor_ext_doubles(n) doubles = the arrays
  // A, U, L over rows -kTM .. n+kTM-1 and their reversed copies.
  // Horner without the division per pass: with K_tt = m (m-1) .. (m-tt+1), G_tt = K_tt H_tt obeys
  // G_tt = K_tt I + (hS') G_{tt-1} -- three FP64 operations per element; s.red[tt] = K_tt and the final 1 / K_m goes
  // into the scale.  The passes are done two at a time (taylor_lane): m is even (taylor_plan).
  MCD_HD void taylor_coefs(int tid, double* ext, const double* diag, const double* lo, const double* up,
                           const TaylorPlan& tp) {
    const int n = P.n, ne = n + 2 * kTM;
    for (int e = tid; e < ne; e += Exec::nthr()) {
      int i = e - kTM;
      bool ok = true;
      if (P.pbc) { i %= n; if (i < 0) i += n; }
      else ok = (i >= 0 && i < n);
      const double va = ok ? tp.h * (diag[i] + tp.shift) : 0.0;
      const double vu = ok ? tp.h * up[i] : 0.0;
      const double vl = ok ? tp.h * lo[i] : 0.0;
      ext[e] = va; ext[ne + e] = vu; ext[2 * ne + e] = vl;
      const int er = ne - 1 - e;
      ext[3 * ne + er] = va; ext[4 * ne + er] = vu; ext[5 * ne + er] = vl;
    }
    if (tid >= 1 && tid <= tp.m) {          // no reduction runs before the passes end
      double K = 1.0;
      for (int q = 0; q < tid; ++q) K *= (double)(tp.m - q);
      s.red[tid] = K;
    }
  }

  // The two lanes of a pair are lanes l and l + 16 of a warp: each half-warp then reads 16 consecutive doubles
  // of a coefficient array per load (one conflict-free wavefront).
  static MCD_HD int pair_col(int tid) { return ((tid >> 5) << 4) + (tid & 15); }   // column index within the phase
  static MCD_HD int pair_hf(int tid) { return (tid >> 4) & 1; }

  // one lane of the pair of column j: hf = 0 rows j - k, hf = 1 rows j + 1 + k (k = 0 .. kTH-1).
  // swap(v): the partner lane's v (device: shuffle).  kSym: the coupling of rows (i, i+1) is the same in both
  // directions (symmetrised generator), so the inner coefficient of element k is the outer one of element k-1.
  // A sweep does TWO passes (tt, tt+1) with one set of coefficient loads: at element k it forms the pass-tt value
  // of element k+1 and, from the pass-tt values of k-1, k, k+1, the pass-(tt+1) value of element k.
  template <bool kSym, class Swap>
  MCD_HD void taylor_lane(int j, int hf, const double* ext, const TaylorPlan& tp, double (&x)[kTH], Swap swap) {
    const int n = P.n, m = tp.m, ne = n + 2 * kTM;
    // element k: own row i, inner neighbour (towards the centre), outer neighbour
    //   lower lane: i = j - k,     inner = row i+1 (coefficient U[i]),   outer = row i-1 (L[i-1])  -> reversed arrays
    //   upper lane: i = j + 1 + k, inner = row i-1 (coefficient L[i-1]), outer = row i+1 (U[i])
    // (selects, not branches: the two lanes of a pair sit in the same warp and must stay converged for the shuffles)
    const int br = ne - 1 - (j + kTM);     // reversed index of row j
    const int bu = j + kTM + 1;            // index of row j + 1
    const double* MCD_RESTRICT pa = ext + (hf ? bu : 3 * ne + br);
    const double* MCD_RESTRICT pin = ext + (hf ? 2 * ne + bu - 1 : 4 * ne + br);
    const double* MCD_RESTRICT pout = ext + (hf ? ne + bu : 5 * ne + br + 1);
    const double cen = hf ? 0.0 : 1.0;
#pragma unroll
    for (int k = 0; k < kTH; ++k) x[k] = (k == 0) ? cen : 0.0;
    // The coefficient loads of a chunk of kTChunk elements are issued one chunk AHEAD of their use (pinned in place on
    // the device: the compiler otherwise sinks each load next to its first use and every element waits ~30 cycles
    // for shared memory -- measured 1 500 cycles per sweep instead of ~400).  Chunks the band has not reached are
    // skipped (uniform branch).
    constexpr int CH = kTChunk;
    for (int tt = 1; tt < m; tt += 2) {
      const double K1 = cen * s.red[tt], K2 = cen * s.red[tt + 1];
      double ca = ldc(pa), cin = ldc(pin), cout = ldc(pout);    // coefficients of element k (start: k = 0)
      double fa[CH], fo[CH], fi[CH];                            // coefficients of elements k0+1 .. k0+CH (next in use)
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        fa[q] = ldc(pa + 1 + q); fo[q] = ldc(pout + 1 + q);
        fi[q] = kSym ? 0.0 : ldc(pin + 1 + q);
      }
      // pass tt of element 0, then the partner's
      double ym = 0.0;                                          // pass-tt value of element k-1
      double yc = fma(ca, x[0], K1);                            // ... of element k
      yc = fma(cin, swap(x[0]), yc);
      yc = fma(cout, x[1], yc);
      const double yp0 = swap(yc);
      MCD_CLK(19);
#pragma unroll
      for (int k0 = 0; k0 < kTH; k0 += CH) {
        if (tt + 1 >= k0) {                                     // the band has reached this chunk (uniform)
          double ga[CH], go[CH], gi[CH];                        // one chunk ahead (clamped to the window)
#pragma unroll
          for (int q = 0; q < CH; ++q) {
            const int e = (k0 + CH + 1 + q < kTH) ? k0 + CH + 1 + q : kTH - 1;
            ga[q] = ldc(pa + e); go[q] = ldc(pout + e);
            gi[q] = kSym ? 0.0 : ldc(pin + e);
          }
#pragma unroll
          for (int q = 0; q < CH; ++q) {
            const int k = k0 + q;
            // pass tt of element k+1
            double na = 0.0, nin = 0.0, nout = 0.0, yn = 0.0;
            if (k + 1 < kTH) {
              na = fa[q]; nout = fo[q];
              nin = kSym ? cout : fi[q];
              yn = na * x[k + 1];
              yn = fma(nin, x[k], yn);
              yn = fma(nout, (k + 2 < kTH) ? x[k + 2] : 0.0, yn);
            }
            // pass tt+1 of element k
            double z = (k == 0) ? fma(ca, yc, K2) : ca * yc;
            z = fma(cin, (k == 0) ? yp0 : ym, z);
            z = fma(cout, yn, z);
            x[k] = z;
            ym = yc; yc = yn; ca = na; cin = nin; cout = nout;
          }
#pragma unroll
          for (int q = 0; q < CH; ++q) { fa[q] = ga[q]; fo[q] = go[q]; fi[q] = gi[q]; }
        }
      }
      MCD_CLK(20);
    }
  }

  // coefficient load of the Taylor lanes (the arrays are in shared memory)
  static MCD_HD double ldc(const double* p) {
#if defined(__CUDA_ARCH__)
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
#else
    return *p;
#endif
  }

  // rows of this lane -> dst[i * ld + j]; add: accumulate (ring: window entries that coincide on the ring are
  // summed; the row of element k walks around the ring one step per element), else plain stores
  MCD_HD void taylor_lane_store(int j, int hf, const TaylorPlan& tp, const double (&x)[kTH], double* dst, int ld, bool add) {
    const int n = P.n, m = tp.m;
    const double sc = tp.scale / s.red[m];
    const int kmax = hf ? m - 1 : m;                               // last element inside the band
    if (add || kPackedDst) {
      // the row of element k walks around the ring one step per element.  Packed storage keeps T[i][j] only for
      // tile rows <= tile columns (the mirrored entry is written by the thread of column i).
      int i = hf ? j + 1 : j;                                      // row of element 0
      if (i >= n) i -= n;
      const int nb8 = P.np >> 3;
#pragma unroll
      for (int k = 0; k < kTH; ++k) {
        if (k > kmax) break;
        const bool inside = P.pbc || (hf ? j + 1 + k < n : k <= j);
        if (inside) {
          if (kPackedDst) {
            if ((i >> 3) <= (j >> 3)) {
              double* q = dst + pk_index(i, j, nb8);
              if (add) *q += x[k] * sc; else *q = x[k] * sc;
            }
          } else {
            dst[(size_t)i * ld + j] += x[k] * sc;
          }
        }
        if (hf) { if (++i >= n) i = 0; }
        else { if (--i < 0) i = n - 1; }
      }
      return;
    }
    // row of element k: i = j - k (lower lane) or j + 1 + k (upper lane), wrapped once for periodic profiles
    const int step = hf ? ld : -ld;
    double* p0 = dst + (size_t)(j + hf) * ld + j;
    const int wrap = P.pbc ? (hf ? -n * ld : n * ld) : 0;
    const int kin = hf ? n - 2 - j : j;                            // last element before the edge of the profile
#pragma unroll
    for (int k = 0; k < kTH; ++k) {
      if (k > kmax) break;
      if (k > kin && !P.pbc) break;
      double* p = p0 + k * step + ((k > kin) ? wrap : 0);
      *p = x[k] * sc;
    }
  }

  // One lane of the pair of column j (all threads of a warp call it together; col_valid == 0: nothing is stored).
  template <bool kSym>
  MCD_HD void taylor_pair(int tid, int col_valid, int j, const double* ext, const TaylorPlan& tp, double* dst, int ld) {
#if defined(__CUDA_ARCH__)
    const int hf = pair_hf(tid);
    double x[kTH];
    taylor_lane<kSym>(j, hf, ext, tp, x, [](double v) { return __shfl_xor_sync(0xffffffffu, v, 16); });
    MCD_CLK(14);
    if (tp.ring) {
      if (col_valid && hf == 0) taylor_lane_store(j, 0, tp, x, dst, ld, true);
      __syncwarp();
      if (col_valid && hf == 1) taylor_lane_store(j, 1, tp, x, dst, ld, true);
    } else if (col_valid) {
      taylor_lane_store(j, hf, tp, x, dst, ld, false);
    }
    MCD_CLK(15);
#else
    // host: the lower thread of the pair runs both lanes in lockstep (same arithmetic as the device lanes)
    if (pair_hf(tid) || !col_valid) return;
    double xs[2][kTH];
    // the lanes exchange (1) their centre elements before a sweep and (2) the centres after the first pass of the sweep
    const int n = P.n, m = tp.m, ne = n + 2 * kTM;
    const int br = ne - 1 - (j + kTM), bu = j + kTM + 1;
    const double* PA[2] = {ext + 3 * ne + br, ext + bu};
    const double* PIN[2] = {ext + 4 * ne + br, ext + 2 * ne + bu - 1};
    const double* POUT[2] = {ext + 5 * ne + br + 1, ext + ne + bu};
    for (int h = 0; h < 2; ++h)
      for (int k = 0; k < kTH; ++k) xs[h][k] = (h == 0 && k == 0) ? 1.0 : 0.0;
    for (int tt = 1; tt < m; tt += 2) {
      double y0[2];
      const double x00[2] = {xs[0][0], xs[1][0]};
      for (int h = 0; h < 2; ++h) {
        const double K1 = (h == 0 ? 1.0 : 0.0) * s.red[tt];
        double yc = fma(PA[h][0], xs[h][0], K1);
        yc = fma(PIN[h][0], x00[1 - h], yc);
        yc = fma(POUT[h][0], xs[h][1], yc);
        y0[h] = yc;
      }
      for (int h = 0; h < 2; ++h) {
        double* x = xs[h];
        const double K2 = (h == 0 ? 1.0 : 0.0) * s.red[tt + 1];
        double ca = PA[h][0], cin = PIN[h][0], cout = POUT[h][0];
        double ym = 0.0, yc = y0[h];
        const double yp0 = y0[1 - h];
        for (int k = 0; k < kTH; ++k) {
          double na = 0.0, nin = 0.0, nout = 0.0, yn = 0.0;
          if (k + 1 < kTH) {
            na = PA[h][k + 1]; nout = POUT[h][k + 1];
            nin = kSym ? cout : PIN[h][k + 1];
            yn = na * x[k + 1];
            yn = fma(nin, x[k], yn);
            yn = fma(nout, (k + 2 < kTH) ? x[k + 2] : 0.0, yn);
          }
          double z = (k == 0) ? fma(ca, yc, K2) : ca * yc;
          z = fma(cin, (k == 0) ? yp0 : ym, z);
          z = fma(cout, yn, z);
          x[k] = z;
          ym = yc; yc = yn; ca = na; cin = nin; cout = nout;
        }
      }
    }
    taylor_lane_store(j, 0, tp, xs[0], dst, ld, tp.ring != 0);
    taylor_lane_store(j, 1, tp, xs[1], dst, ld, tp.ring != 0);
#endif
  }

  // T -> dst for the generator (diag, lo, up) currently in shared memory; ext: scratch for the coefficient
  // arrays (taylor_ext_doubles(n) doubles, not overlapping dst).  dst (np x ld doubles) is cleared here.
  template <bool kSym>
  MCD_HD void taylor_band(const double* diag, const double* lo, const double* up, const TaylorPlan& tp, double* ext,
                          double* dst, int ld) {
    const int n = P.n;
    const int nzero = kPackedDst ? (int)pk_elems(P.np) : P.np * ld;
    // a cluster shares dst (global): every CTA clears and fills its share, with a cluster barrier in between
    const int cs = ex.csize, cr = ex.crank;
    const int zper = ((nzero / 2 + cs - 1) / cs) * 2;
    const int z_lo = cr * zper, z_hi = (z_lo + zper < nzero) ? z_lo + zper : nzero;
    const int cper = (n + cs - 1) / cs;
    const int c_lo = cr * cper, c_hi = (c_lo + cper < n) ? c_lo + cper : n;
    ex.par([&](int tid) {
      for (int e = z_lo + tid * 2; e < z_hi; e += 2 * Exec::nthr()) st2(dst + e, 0.0, 0.0);
      taylor_coefs(tid, ext, diag, lo, up, tp);
    });
    ex.cluster_sync();
    MCD_CLK(3);
    ex.par([&](int tid) {
      const int per = Exec::nthr() / 2;
      for (int j0 = c_lo; j0 < c_hi; j0 += per) {
        if (j0 + ((tid >> 5) << 4) >= c_hi) break;       // the whole warp is past the last column
        const int j = j0 + pair_col(tid);
        taylor_pair<kSym>(tid, j < c_hi, j < c_hi ? j : c_hi - 1, ext, tp, dst, ld);
      }
    });
    ex.cluster_sync();
  }

  // lagt must be an arithmetic progression lagt[l] = (l+1) lagt[0] (checked on the host).
  // exp(t S) by Taylor + squarings.  Uses both work buffers; returns the one holding the result
  // (*other receives the second one).  epi >= 0: the likelihood terms of lag index epi are added
  // (fused into the last squaring).  Returns nullptr for a non-finite / absurd scale.
  MCD_HD double* sq_chain(double t, double gersh, int taylor_m, bool use_mma, int epi, double* Pout, double** other,
                          int* nsq_out) {
    const int np = P.np, ld = P.ld;
    TaylorPlan tp;
    if (!taylor_plan(t, gersh, taylor_m, tp)) return nullptr;
    const int nsq = tp.nsq;
    MCD_CLK(16);
    // ---- T = e^{-hc} T_m(h(S + cI)) into XT ----
    taylor_band<true>(s.a, s.b, s.b, tp, s.text, s.XT, ld);
    double* src = s.XT;
    double* dst = s.B2;
    // ---- G(t) = T^(2^s): out-of-place squarings (direct stores), the last one carries the
    //      likelihood epilogue of lag `epi` ----
    MCD_CLK(4);
    double* G = src;      // current power
    double* O = dst;      // the other buffer
    const int* cnt_e = (epi >= 0) ? P.counts + (size_t)epi * np * np : nullptr;
    const int* nsy_e = (epi >= 0) ? P.nsym + (size_t)epi * np * np : nullptr;
    for (int q = 0; q < nsq; ++q) {
      const bool last = (q + 1 == nsq);
      const int* cnt = last ? cnt_e : nullptr;
      const int* nsy = last ? nsy_e : nullptr;
      if (use_mma) dmma_product(G, G, O, cnt, nsy, last ? Pout : nullptr, 1);
      else sym_product(G, G, O, cnt, nsy, last ? Pout : nullptr, 1);
      double* tmp = G; G = O; O = tmp;
    }
    if (nsq == 0 && epi >= 0) buffer_loglike(G, cnt_e, nsy_e, Pout);
    *other = O;
    if (nsq_out) *nsq_out += nsq;
    return G;
  }

  // likelihood terms of a kernel that already sits in a buffer (no product needed)
  MCD_HD void buffer_loglike(const double* G, const int* cnt, const int* nsy, double* Pl) {
    const int ld = P.ld;
    ex.par([&](int tid) {
      for (int tt = tid; tt < s.ntiles; tt += Exec::nthr()) {
        const int ib = s.tiles[2 * tt], jb = s.tiles[2 * tt + 1];
        i4 ns[4];
        load_counts(nsy, ib, jb, ns);
        double c[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const d2 g01 = ld2(G + (ib * 4 + a) * ld + jb * 4), g23 = ld2(G + (ib * 4 + a) * ld + jb * 4 + 2);
          c[a][0] = g01.x; c[a][1] = g01.y; c[a][2] = g23.x; c[a][3] = g23.y;
        }
        s.part[tid] += tile_loglike(c, ib, jb, ns, cnt, Pl);
      }
    });
  }

  MCD_HD void copy_buffer(double* dst, const double* src) {
    const int total = P.np * P.ld;
    ex.par([&](int tid) {
      for (int e = tid * 2; e < total; e += 2 * Exec::nthr()) {
        const d2 x = ld2(src + e);
        st2(dst + e, x.x, x.y);
      }
    });
  }

  // log L of (v, w) at the lag times lagt with the squaring engine, following LagPlan `pl`.
  // slots: per-CTA global scratch of pl.nslots work matrices (only read/written when pl.fast == 0).
  MCD_HD double evaluate_sq(const double* v, const double* w, const double* lagt, const LagPlan& pl, double* slots,
                            int taylor_m, bool tensor, double* Pout, int* squarings_out) {
    const bool use_mma = kTensorOnly || (tensor && dmma_ok());
    const int n = P.n, np = P.np;
    const size_t nn = (size_t)n * n, slot_elems = (size_t)np * P.ld;
    build_S(v, w);
    const double gersh = s.sc->gersh;   // Gershgorin bound of ||S||
    ex.par([&](int tid) { s.part[tid] = 0.0; });
    MCD_CLK(2);
    int nsq = 0;
    bool bad = false;
    double* R = nullptr;   // running kernel G(t_i)
    double* O = nullptr;   // the other shared buffer
    const int l0 = pl.order[0];
    const bool fast = pl.fast != 0;
    // operand chains first (their results are parked in the slots), the G(t_0) chain last.
    // ONE inlined call site for the chain and ONE for the products below: the kernel is large and
    // duplicated bodies cost instruction-cache misses (measured: -12 % with two copies).
    for (int c = 0; c < pl.nchain && !bad; ++c) {
      const bool is_first_lag = (c == pl.nchain - 1);
      const double t = lagt[pl.chain_a[c]] - (pl.chain_b[c] >= 0 ? lagt[pl.chain_b[c]] : 0.0);
      double* other = nullptr;
      double* G = sq_chain(t, gersh, taylor_m, use_mma, is_first_lag ? l0 : -1,
                           (is_first_lag && Pout) ? Pout + (size_t)l0 * nn : nullptr, &other, &nsq);
      if (!G) { bad = true; break; }
      if (!fast && pl.chain_slot[c] >= 0) copy_buffer(slots + (size_t)pl.chain_slot[c] * slot_elems, G);
      R = G;
      O = other;
    }
    // fast (arithmetic progression): G((i+1) t0) = G(i t0) G(t0), first product out of place into O,
    // then in place on O with G(t0) staying in R.  general: R <- R G(delta_i) in place, operand in O.
    double* G0 = R;
    int in_O = -1;   // slot currently loaded in O (general plans)
    for (int i = 1; i < P.nlag && !bad; ++i) {
      const int l = pl.order[i];
      const int* cnt = P.counts + (size_t)l * np * np;
      const int* nsy = P.nsym + (size_t)l * np * np;
      double* Pl = Pout ? Pout + (size_t)l * nn : nullptr;
      if (!fast && pl.op_slot[i] < 0) {
        buffer_loglike(R, cnt, nsy, Pl);          // same lag time as the previous one
      } else {
        const double* X;
        const double* Y;
        double* D;
        int store;
        if (fast && P.nlag <= 4) {
          // t, 2t, 3t, 4t: O = G(2t) is stored once, G(3t) = G(2t) G(t) and G(4t) = G(2t)^2 are needed for their
          // likelihood terms only -- no in-place product (register staging, a second barrier phase) at all
          X = (i == 1) ? G0 : O; Y = (i == 3) ? O : G0; D = O;
          store = (i == 1 && P.nlag > 2) ? 1 : 0;
        } else if (fast) {
          X = (i == 1) ? G0 : O; Y = G0; D = O;
          store = (i + 1 < P.nlag) ? (i == 1 ? 1 : 2) : 0;
        } else {
          if (in_O != pl.op_slot[i]) {
            copy_buffer(O, slots + (size_t)pl.op_slot[i] * slot_elems);
            in_O = pl.op_slot[i];
          }
          X = R; Y = O; D = R; store = 2;
        }
        if (use_mma) dmma_product(X, Y, D, cnt, nsy, Pl, store);
        else sym_product(X, Y, D, cnt, nsy, Pl, store);
      }
      if (!fast && pl.park[i] >= 0) copy_buffer(slots + (size_t)pl.park[i] * slot_elems, R);
    }
    if (squarings_out) *squarings_out = bad ? -1 : nsq;
    if (bad) {
      ex.single([&]() { s.sc->status |= ST_NONFINITE; });
      return nan("");
    }
    const double ll_total = s.sc->lin + ex.sum([&](int tid) { return s.part[tid]; });
    MCD_CLK(11);
    return ll_total;
  }

  // Kernels G(lagt[l]) = exp(lagt[l] S) of the generator currently in s.a / s.b for every lag time,
  // without likelihood terms: each one is copied to dump + l * np * ld (global).  The radial model
  // needs a Bessel-weighted sum of such kernels before the logarithm (lib/twod.py:118-126).
  // Same LagPlan walk as evaluate_sq.  Returns false for a non-finite / absurd scale.
  MCD_HD bool kernels_sq(const double* lagt, const LagPlan& pl, double* slots, double* dump, double gersh, bool use_mma) {
    const size_t me = (size_t)P.np * P.ld;
    const bool fast = pl.fast != 0;
    double* R = nullptr;
    double* O = nullptr;
    int nsq = 0;
    for (int c = 0; c < pl.nchain; ++c) {
      const double t = lagt[pl.chain_a[c]] - (pl.chain_b[c] >= 0 ? lagt[pl.chain_b[c]] : 0.0);
      double* other = nullptr;
      double* G = sq_chain(t, gersh, 0, use_mma, -1, nullptr, &other, &nsq);
      if (!G) return false;
      if (!fast && pl.chain_slot[c] >= 0) copy_buffer(slots + (size_t)pl.chain_slot[c] * me, G);
      R = G;
      O = other;
    }
    copy_buffer(dump + (size_t)pl.order[0] * me, R);
    double* G0 = R;
    int in_O = -1;
    for (int i = 1; i < P.nlag; ++i) {
      const int l = pl.order[i];
      const double* res = R;
      if (fast || pl.op_slot[i] >= 0) {
        const double* X;
        const double* Y;
        double* D;
        int store;
        if (fast) {
          X = (i == 1) ? G0 : O; Y = G0; D = O; store = (i == 1) ? 1 : 2;
          res = O;
        } else {
          if (in_O != pl.op_slot[i]) {
            copy_buffer(O, slots + (size_t)pl.op_slot[i] * me);
            in_O = pl.op_slot[i];
          }
          X = R; Y = O; D = R; store = 2;
        }
        if (use_mma) dmma_product(X, Y, D, nullptr, nullptr, nullptr, store);
        else sym_product(X, Y, D, nullptr, nullptr, nullptr, store);
      }
      copy_buffer(dump + (size_t)l * me, res);
      if (!fast && pl.park[i] >= 0) copy_buffer(slots + (size_t)pl.park[i] * me, R);
    }
    return true;
  }

  // ------------------------------------------------------------------------------------
  // General (non-symmetric) squaring engine: --pull, lib/utils.py:133-183.
  // A constant force on a periodic profile breaks detailed balance around the ring, so R cannot be
  // symmetrised.  The engine works on R itself: P(t) = exp(tR) = (T_m(hR))^(2^s), all entries
  // still non-negative (R is a generator), products are full n x n m8n8k4 products (A fragments
  // read rows of X), and the likelihood is sum N_ij log max(P_ij, clip) entry by entry.
  //   s.a = diagonal, s.b[k] = R[k+1][k] (rate k -> k+1), s.eb[k] = R[k][k+1] (rate k+1 -> k).
  // Same LagPlan walk as evaluate_sq.  Warp task q: tile row q / ngrp, tile columns 8 (q % ngrp) ...;
  // a round of 12 tasks completes whole rows, so an in-place product X <- X Y (register-staged,
  // stored after the round's barrier) never overwrites rows a later round still reads.
  // ------------------------------------------------------------------------------------
  MCD_HD void build_R_pull(const double* v, const double* w) {
    const int n = P.n, np = P.np;
    const double dF = P.pull;
    ex.par([&](int tid) {
      for (int i = tid; i < np; i += Exec::nthr()) {
        double up = 0.0, dn = 0.0;
        if (i < n - 1) {
          const double dv = v[i + 1] - v[i] + dF;
          up = exp(w[i] - 0.5 * dv);                    // R[i+1][i]
          dn = exp(w[i] + 0.5 * dv);                    // R[i][i+1]
        } else if (i == n - 1) {                        // the corner pair, exactly as lib/utils.py:172-173
          up = exp(w[n - 1] - 0.5 * (v[0] - v[n - 1] + dF));   // R[0][n-1]
          dn = exp(w[n - 1] - 0.5 * (v[0] - v[n - 1] - dF));   // R[n-1][0]
        }
        s.b[i] = up;
        s.eb[i] = dn;
        s.ea[i] = 1.0; s.hv[i] = 0.0;
      }
    });
    ex.par([&](int tid) {
      for (int i = tid; i < np; i += Exec::nthr()) {
        // column i: R[i-1][i] = eb[i-1] (down), R[i+1][i] = b[i] (up); lib/utils.py:175-181
        const int im = (i == 0) ? n - 1 : i - 1;
        s.a[i] = (i < n) ? -s.eb[im] - s.b[i] : 0.0;
      }
    });
    const double amax = ex.max([&](int tid) {
      double m = 0.0;
      for (int i = tid; i < n; i += Exec::nthr()) {
        const double r = fabs(s.a[i]);
        m = (r > m || r != r) ? r : m;
      }
      return m;
    });
    ex.single([&]() {
      s.sc->gersh = 2.0 * amax;      // ||R||_1: every column sums to zero
      s.sc->scale = amax;
      s.sc->gmin = P.clip;
      s.sc->lin = 0.0;
    });
  }

  // D <- X Y, full product.  cnt != nullptr: D is the propagator of a lag time, its likelihood terms are
  // added to s.part.  store: 1 = D is a third buffer (direct stores), 2 = D aliases X (staged per round).
  MCD_HD void gen_product(const double* X, const double* Y, double* D, const int* cnt, double* Pl, int store) {
    const int n = P.n, np = P.np, ld = P.ld, nb8 = np / 8;
    const int ngrp = (nb8 + kTilesPerWarp - 1) / kTilesPerWarp;
    const int nwarp = Exec::nthr() / 32;
    const int per_round = (nwarp / ngrp) * ngrp;       // whole rows per round
    const int ntask = nb8 * ngrp;
    const double clip = P.clip;
    for (int base = 0; base < ntask; base += per_round) {
      ex.par([&](int tid) {
        const int warp = tid >> 5, lane = tid & 31;
        const int q = base + warp;
        const bool active = warp < per_round && q < ntask;
        const int row = active ? q / ngrp : 0, grp = active ? q - row * ngrp : 0;
        const int jb0 = grp * kTilesPerWarp;
        const int nt = active ? ((nb8 - jb0 < kTilesPerWarp) ? nb8 - jb0 : kTilesPerWarp) : 0;
        const int fr = lane >> 2, fk = lane & 3;
        double c[kTilesPerWarp][2];
        int nv[kTilesPerWarp][2];
#pragma unroll
        for (int t = 0; t < kTilesPerWarp; ++t) {
          c[t][0] = 0.0; c[t][1] = 0.0;
          nv[t][0] = 0; nv[t][1] = 0;
          if (cnt && active) {
            const int jb = jb0 + ((t < nt) ? t : 0);
            const i2 x = ld2i(cnt + (size_t)(row * 8 + fr) * np + jb * 8 + 2 * fk);
            nv[t][0] = x.x; nv[t][1] = x.y;
          }
        }
        if (active) {
#if defined(__CUDA_ARCH__)
          // A fragment: X[8 row + fr][k0 + fk] (rows of X), B fragments: Y[k0 + fk][8 jb + fr]
          unsigned ax = (unsigned)__cvta_generic_to_shared(X) + (unsigned)((row * 8 + fr) * ld + fk) * 8u;
          unsigned ay = (unsigned)__cvta_generic_to_shared(Y) + (unsigned)(fk * ld + fr) * 8u;
          int offB[kTilesPerWarp];
#pragma unroll
          for (int t = 0; t < kTilesPerWarp; ++t) offB[t] = (jb0 + ((t < nt) ? t : 0)) * 64;
          const int pitch4 = 4 * ld * 8, nk4 = (P.n + 3) >> 2;   // columns k >= n of X / rows k >= n of Y are zero padding
#pragma unroll 2
          for (int it = 0; it < nk4; ++it) {
            const double a0 = lds64(ax);
#pragma unroll
            for (int t = 0; t < kTilesPerWarp; ++t) {
              const double b = lds64(ay + offB[t]);
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                           : "+d"(c[t][0]), "+d"(c[t][1])
                           : "d"(a0), "d"(b));
            }
            ax += 32;
            ay += pitch4;
          }
#else
          for (int t = 0; t < nt; ++t) {
            const int i = row * 8 + fr;
            for (int e = 0; e < 2; ++e) {
              const int j = (jb0 + t) * 8 + 2 * fk + e;
              double acc = 0.0;
              for (int k = 0; k < np; ++k) acc = fma(X[i * ld + k], Y[k * ld + j], acc);
              c[t][e] = acc;
            }
          }
#endif
        }
        if (cnt && nt > 0) {
          double part = 0.0;
#pragma unroll
          for (int t = 0; t < kTilesPerWarp; ++t) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int ns = (t < nt) ? nv[t][e] : 0;
              part += (double)ns * fast_log(fmax(c[t][e], clip));     // lib/utils.py:317-318
            }
          }
          s.part[tid] += part;
          if (Pl) {
#pragma unroll
            for (int t = 0; t < kTilesPerWarp; ++t) {
              const int i = row * 8 + fr, j = (jb0 + t) * 8 + 2 * fk;
              if (t < nt && i < n) {
                if (j < n) Pl[(size_t)i * n + j] = c[t][0];
                if (j + 1 < n) Pl[(size_t)i * n + j + 1] = c[t][1];
              }
            }
          }
        }
        double* r = ex.regs(tid);
#pragma unroll
        for (int t = 0; t < kTilesPerWarp; ++t) { r[2 * t] = c[t][0]; r[2 * t + 1] = c[t][1]; }
        if (store == 1) {
#pragma unroll
          for (int t = 0; t < kTilesPerWarp; ++t)
            if (t < nt) st2(D + (row * 8 + fr) * ld + (jb0 + t) * 8 + 2 * fk, r[2 * t], r[2 * t + 1]);
        }
      });
      if (store == 2) {
        ex.par([&](int tid) {
          const int warp = tid >> 5, lane = tid & 31;
          const int q = base + warp;
          if (warp < per_round && q < ntask) {
            const int row = q / ngrp, grp = q - row * ngrp;
            const int jb0 = grp * kTilesPerWarp;
            const int nt = (nb8 - jb0 < kTilesPerWarp) ? nb8 - jb0 : kTilesPerWarp;
            const int fr = lane >> 2, fk = lane & 3;
            const double* r = ex.regs(tid);
            for (int t = 0; t < nt; ++t) st2(D + (row * 8 + fr) * ld + (jb0 + t) * 8 + 2 * fk, r[2 * t], r[2 * t + 1]);
          }
        });
      }
    }
  }

  // likelihood terms of a propagator that already sits in a buffer (general engine)
  MCD_HD void gen_buffer_loglike(const double* G, const int* cnt, double* Pl) {
    const int n = P.n, np = P.np, ld = P.ld;
    const double clip = P.clip;
    ex.par([&](int tid) {
      double part = 0.0;
      for (int e = tid; e < n * n; e += Exec::nthr()) {
        const int i = e / n, j = e - i * n;
        const double g = G[i * ld + j];
        part += (double)cnt[(size_t)i * np + j] * fast_log(fmax(g, clip));
        if (Pl) Pl[e] = g;
      }
      s.part[tid] += part;
    });
  }

  // exp(t R) by Taylor + squarings (general engine); same contract as sq_chain
  MCD_HD double* gen_chain(double t, double gersh, int taylor_m, int epi, double* Pout, double** other, int* nsq_out) {
    const int np = P.np, ld = P.ld;
    // uniformisation: R + cI, c = max |R_jj| = gersh / 2, is c times a column-stochastic matrix -- every entry
    // non-negative, 1-norm exactly c (taylor_plan's shift)
    TaylorPlan tp;
    if (!taylor_plan(t, gersh, taylor_m, tp)) return nullptr;
    const int nsq = tp.nsq;
    // (R H)[i][j] = a_i H[i][j] + R[i][i-1] H[i-1][j] + R[i][i+1] H[i+1][j]:  lower coupling s.b[i-1], upper s.eb[i]
    taylor_band<false>(s.a, s.b, s.eb, tp, s.text, s.XT, ld);
    double* G = s.XT;
    double* O = s.B2;
    const int* cnt_e = (epi >= 0) ? P.counts + (size_t)epi * np * np : nullptr;
    for (int q = 0; q < nsq; ++q) {
      const bool last = (q + 1 == nsq);
      gen_product(G, G, O, last ? cnt_e : nullptr, last ? Pout : nullptr, 1);
      double* tmp = G; G = O; O = tmp;
    }
    if (nsq == 0 && epi >= 0) gen_buffer_loglike(G, cnt_e, Pout);
    *other = O;
    if (nsq_out) *nsq_out += nsq;
    return G;
  }

  MCD_HD double evaluate_gen(const double* v, const double* w, const double* lagt, const LagPlan& pl, double* slots,
                             int taylor_m, double* Pout, int* squarings_out) {
    const int n = P.n, np = P.np;
    const size_t nn = (size_t)n * n, slot_elems = (size_t)np * P.ld;
    build_R_pull(v, w);
    const double gersh = s.sc->gersh;
    ex.par([&](int tid) { s.part[tid] = 0.0; });
    int nsq = 0;
    bool bad = false;
    double* R = nullptr;
    double* O = nullptr;
    const int l0 = pl.order[0];
    const bool fast = pl.fast != 0;
    for (int c = 0; c < pl.nchain && !bad; ++c) {
      const bool is_first_lag = (c == pl.nchain - 1);
      const double t = lagt[pl.chain_a[c]] - (pl.chain_b[c] >= 0 ? lagt[pl.chain_b[c]] : 0.0);
      double* other = nullptr;
      double* G = gen_chain(t, gersh, taylor_m, is_first_lag ? l0 : -1,
                            (is_first_lag && Pout) ? Pout + (size_t)l0 * nn : nullptr, &other, &nsq);
      if (!G) { bad = true; break; }
      if (!fast && pl.chain_slot[c] >= 0) copy_buffer(slots + (size_t)pl.chain_slot[c] * slot_elems, G);
      R = G;
      O = other;
    }
    double* G0 = R;
    int in_O = -1;
    for (int i = 1; i < P.nlag && !bad; ++i) {
      const int l = pl.order[i];
      const int* cnt = P.counts + (size_t)l * np * np;
      double* Pl = Pout ? Pout + (size_t)l * nn : nullptr;
      if (!fast && pl.op_slot[i] < 0) {
        gen_buffer_loglike(R, cnt, Pl);
      } else if (fast) {
        gen_product((i == 1) ? G0 : O, G0, O, cnt, Pl, (i == 1) ? 1 : 2);   // G((i+1) t0) = G(i t0) G(t0)
      } else {
        if (in_O != pl.op_slot[i]) {
          copy_buffer(O, slots + (size_t)pl.op_slot[i] * slot_elems);
          in_O = pl.op_slot[i];
        }
        gen_product(R, O, R, cnt, Pl, 2);
      }
      if (!fast && pl.park[i] >= 0) copy_buffer(slots + (size_t)pl.park[i] * slot_elems, R);
    }
    if (squarings_out) *squarings_out = bad ? -1 : nsq;
    if (bad) {
      ex.single([&]() { s.sc->status |= ST_NONFINITE; });
      return nan("");
    }
    return ex.sum([&](int tid) { return s.part[tid]; });
  }

  // ------------------------------------------------------------------------------------
  // Packed symmetric squaring engine (kPacked, np <= 104): same mathematics as evaluate_sq, the work matrices
  // in packed tile storage (pk_tile / pk_off) so that TWO chains fit in the shared memory of one SM -- the
  // scalar phases of one chain (Taylor, epilogues, Metropolis, barriers) overlap the m8n8k4 loops of the other.
  // 192 threads per chain (two CTAs x 170 registers per thread fill the register file): a product has more warp
  // tasks (12 at 104 bins) than warps (6) and runs in two full rounds.
  //   C_IJ = sum_K X_KI^T Y_KJ with 8 x 8 tiles: the A fragment of tile row I at k-block K comes from the stored tile
  //   (K, I) when K <= I and from the TRANSPOSE of the stored tile (I, K) when K > I (likewise B with J); every
  //   operand tile keeps a running byte offset (+ 512 (nb8 - K - 1) per step while K < I, + 512 afterwards).
  // Products never run in place: a result that must replace one of its operands is written to a per-chain
  // global scratch matrix (L2) and copied back.
  // ------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
  // k loop of one warp task on packed operands, software pipelined: the 10 fragment loads of step (K, h) are issued
  // BEFORE the 8 MMAs of the previous step, so that a warp alone on its sub-partition (the other warps of both
  // resident CTAs may be in scalar phases) still keeps the tensor pipe fed -- the volatile asm statements keep
  // exactly this order.
  template <int NA>
  static __device__ __forceinline__ void pk_mma_loop(unsigned xb, unsigned yb, int nb8, int rowA, int rowB,
                                                     const int (&jb)[kTilesPerWarp], int fr, int fk,
                                                     double (&c)[kTilesPerWarp][2]) {
    // lane offsets (bytes) inside a stored tile: direct (rows = k) and transposed (rows = i) reading, k-halves 0 / 1
    const int sw_k = ((fk >> 1) & 1) << 2, sw_r = ((fr >> 1) & 1) << 2;
    const int d0 = (fk * 8 + (fr ^ sw_k)) * 8;                 // direct, h = 0 (h = 1: + 256)
    const int t0 = (fr * 8 + fk + sw_r) * 8;                   // transposed, h = 0
    const int th = (4 - 2 * sw_r) * 8;                         // transposed, h = 1 relative to h = 0
    unsigned pA0 = xb + 512u * (unsigned)rowA, pA1 = xb + 512u * (unsigned)rowB;   // tile (0, Q) has index Q
    unsigned pB[kTilesPerWarp];
#pragma unroll
    for (int t = 0; t < kTilesPerWarp; ++t) pB[t] = yb + 512u * (unsigned)jb[t];
    double xa0, xa1, xb_[kTilesPerWarp];                      // operands of (K, 0)
    double ya0, ya1, yb_[kTilesPerWarp];                      // operands of (K, 1)
    xa0 = lds64(pA0 + d0);
    xa1 = (NA < kTilesPerWarp) ? lds64(pA1 + d0) : 0.0;
#pragma unroll
    for (int t = 0; t < kTilesPerWarp; ++t) xb_[t] = lds64(pB[t] + d0);     // K = 0 <= every block index: direct
#pragma unroll 1
    for (int K = 0; K < nb8; ++K) {
      // ---- loads of (K, 1), MMAs of (K, 0) ----
      ya0 = lds64(pA0 + ((K <= rowA) ? d0 + 256 : t0 + th));
      if (NA < kTilesPerWarp) ya1 = lds64(pA1 + ((K <= rowB) ? d0 + 256 : t0 + th));
#pragma unroll
      for (int t = 0; t < kTilesPerWarp; ++t) yb_[t] = lds64(pB[t] + ((K <= jb[t]) ? d0 + 256 : t0 + th));
#pragma unroll
      for (int t = 0; t < kTilesPerWarp; ++t) {
        const double a = (t < NA) ? xa0 : xa1;
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[t][0]), "+d"(c[t][1])
                     : "d"(a), "d"(xb_[t]));
      }
      // ---- advance to K + 1 (the last step stays where it is: its loads are not used), loads of (K + 1, 0), MMAs of (K, 1) ----
      const int K1 = K + 1;
      if (K1 < nb8) {
        const int inc = 512 * (nb8 - K1);
        pA0 += (K < rowA) ? inc : 512;
        pA1 += (K < rowB) ? inc : 512;
#pragma unroll
        for (int t = 0; t < kTilesPerWarp; ++t) pB[t] += (K < jb[t]) ? inc : 512;
      }
      xa0 = lds64(pA0 + ((K1 <= rowA) ? d0 : t0));
      if (NA < kTilesPerWarp) xa1 = lds64(pA1 + ((K1 <= rowB) ? d0 : t0));
#pragma unroll
      for (int t = 0; t < kTilesPerWarp; ++t) xb_[t] = lds64(pB[t] + ((K1 <= jb[t]) ? d0 : t0));
#pragma unroll
      for (int t = 0; t < kTilesPerWarp; ++t) {
        const double a = (t < NA) ? ya0 : ya1;
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[t][0]), "+d"(c[t][1])
                     : "d"(a), "d"(yb_[t]));
      }
    }
  }
#endif

  // element (i, j) of a packed symmetric matrix (global workspace matrices are read past L1: the other CTA of a
  // cluster may have written them)
  MCD_HD double pk_get(const double* G, int i, int j) const {
#if defined(__CUDA_ARCH__)
    if (kBlocked && kCluster) return __ldcg(G + pk_index(i, j, P.np >> 3));
#endif
    return G[pk_index(i, j, P.np >> 3)];
  }

  // D <- X Y for packed symmetric commuting operands; D is another packed buffer (shared memory or the global
  // scratch), never an operand.  cnt != nullptr: likelihood terms of a lag time (as dmma_product).  store: 0 / 1.
  MCD_HD void pk_product(const double* X, const double* Y, double* D, const int* cnt, const int* nsy, double* Pl,
                         int store) {
    const int n = P.n, np = P.np, nb8 = np >> 3;
    const int ntasks = s.sc->mma_tasks;
    const int nwarp = Exec::nthr() / 32;
    for (int base = 0; base < ntasks; base += nwarp) {
      ex.par([&](int tid) {
        const int warp = tid >> 5, lane = tid & 31;
        const WarpTask task = load_task(base + warp, ntasks);
        const bool active = base + warp < ntasks;
        const int nt = task.nt, na = task.na, rowA = task.rowA, rowB = task.rowB;
        double c[kTilesPerWarp][2];
#pragma unroll
        for (int t = 0; t < kTilesPerWarp; ++t) { c[t][0] = 0.0; c[t][1] = 0.0; }
        const int fr = lane >> 2, fk = lane & 3;
        int nsv[kTilesPerWarp][2];
        if (cnt && active) {
#pragma unroll
          for (int t = 0; t < kTilesPerWarp; ++t) {
            const int ib = (t < na) ? rowA : rowB, jb = task.jb[t];
            const i2 q = ld2i(nsy + (size_t)(ib * 8 + fr) * np + jb * 8 + 2 * fk);
            nsv[t][0] = q.x;
            nsv[t][1] = q.y;
          }
        }
        if (active) {
#if defined(__CUDA_ARCH__)
          const unsigned xb = (unsigned)__cvta_generic_to_shared(X), yb = (unsigned)__cvta_generic_to_shared(Y);
          switch (na) {
            case 1: pk_mma_loop<1>(xb, yb, nb8, rowA, rowB, task.jb, fr, fk, c); break;
            case 2: pk_mma_loop<2>(xb, yb, nb8, rowA, rowB, task.jb, fr, fk, c); break;
            case 3: pk_mma_loop<3>(xb, yb, nb8, rowA, rowB, task.jb, fr, fk, c); break;
            case 4: pk_mma_loop<4>(xb, yb, nb8, rowA, rowB, task.jb, fr, fk, c); break;
            case 5: pk_mma_loop<5>(xb, yb, nb8, rowA, rowB, task.jb, fr, fk, c); break;
            case 6: pk_mma_loop<6>(xb, yb, nb8, rowA, rowB, task.jb, fr, fk, c); break;
            case 7: pk_mma_loop<7>(xb, yb, nb8, rowA, rowB, task.jb, fr, fk, c); break;
            default: pk_mma_loop<8>(xb, yb, nb8, rowA, rowB, task.jb, fr, fk, c); break;
          }
#else
          for (int t = 0; t < nt; ++t) {
            const int ib = (t < na) ? rowA : rowB, jb = task.jb[t];
            const int i = ib * 8 + fr;
            for (int e = 0; e < 2; ++e) {
              const int j = jb * 8 + 2 * fk + e;
              double acc = 0.0;
              for (int k = 0; k < np; ++k) acc = fma(pk_get(X, k, i), pk_get(Y, k, j), acc);
              c[t][e] = acc;
            }
          }
#endif
        }
        // likelihood epilogue: as dmma_product (branch-free over the 16 entries, clipped entries redone exactly)
        if (cnt && nt > 0) {
          const double gmin = s.sc->gmin;
          double part = 0.0;
          bool rare = false;
#pragma unroll
          for (int t = 0; t < kTilesPerWarp; ++t) {
            const int ib = (t < na) ? rowA : rowB, jb = task.jb[t];
            const int i = ib * 8 + fr, j = jb * 8 + 2 * fk;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const bool valid = (t < nt) && !(ib == jb && j + e < i);
              const double g = c[t][e];
              const bool above = g >= gmin;
              const int ns = valid ? nsv[t][e] : 0;
              rare |= (ns != 0) && !above;
              part += (double)(above ? ns : 0) * fast_log(above ? g : 1.0);
            }
          }
          if (rare || Pl) {
#pragma unroll
            for (int t = 0; t < kTilesPerWarp; ++t) {
              const int ib = (t < na) ? rowA : rowB, jb = task.jb[t];
              const int i = ib * 8 + fr, j = jb * 8 + 2 * fk;
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int jj = j + e;
                if (!(t < nt) || (ib == jb && jj < i)) continue;
                if (nsv[t][e] != 0 && !(c[t][e] >= gmin))
                  part += pair_term_clipped(c[t][e], nsv[t][e], i, jj, cnt, P.np, P.clip, s.ea, s.eb, s.hv);
                if (Pl && i < n && jj < n) {
                  Pl[(size_t)i * n + jj] = s.ea[i] * c[t][e] * s.eb[jj];
                  Pl[(size_t)jj * n + i] = s.ea[jj] * c[t][e] * s.eb[i];
                }
              }
            }
          }
          s.part[tid] += part;
        }
        if (store) {
          const int swz = ((fr >> 1) & 1) << 2;
#pragma unroll
          for (int t = 0; t < kTilesPerWarp; ++t) {
            if (t < nt) {
              const int ib = (t < na) ? rowA : rowB, jb = task.jb[t];
              st2(D + (size_t)pk_tile(ib, jb, nb8) * 64 + fr * 8 + ((2 * fk) ^ swz), c[t][0], c[t][1]);
            }
          }
        }
      });
    }
  }

  // likelihood terms of a packed kernel that already sits in a buffer
  MCD_HD void pk_buffer_loglike(const double* G, const int* cnt, const int* nsy, double* Pl) {
    const int n = P.n, np = P.np;
    ex.par([&](int tid) {
      double part = 0.0;
      for (int e = tid; e < n * n; e += Exec::nthr()) {
        const int i = e / n, j = e - i * n;
        if (j < i) continue;
        const double g = pk_get(G, i, j);
        part += pair_term(g, nsy[(size_t)i * np + j], i, j, cnt);
        if (Pl) {
          Pl[(size_t)i * n + j] = s.ea[i] * g * s.eb[j];
          Pl[(size_t)j * n + i] = s.ea[j] * g * s.eb[i];
        }
      }
      s.part[tid] += part;
    });
  }

  MCD_HD void pk_copy(double* dst, const double* src) {
    const int total = (int)pk_elems(P.np);
    ex.par([&](int tid) {
      for (int e = tid * 2; e < total; e += 2 * Exec::nthr()) {
        const d2 x = ld2(src + e);
        st2(dst + e, x.x, x.y);
      }
    });
  }

  // exp(t S) by Taylor + squarings in packed storage; same contract as sq_chain
  MCD_HD double* pk_chain(double t, double gersh, int taylor_m, int epi, double* Pout, double** other, int* nsq_out) {
    const int np = P.np;
    TaylorPlan tp;
    if (!taylor_plan(t, gersh, taylor_m, tp)) return nullptr;
    const int nsq = tp.nsq;
    taylor_band<true>(s.a, s.b, s.b, tp, s.B2, s.XT, 0);       // the coefficient arrays live in B2 (free until the first squaring)
    double* G = s.XT;
    double* O = s.B2;
    const int* cnt_e = (epi >= 0) ? P.counts + (size_t)epi * np * np : nullptr;
    const int* nsy_e = (epi >= 0) ? P.nsym + (size_t)epi * np * np : nullptr;
    for (int q = 0; q < nsq; ++q) {
      const bool last = (q + 1 == nsq);
      pk_product(G, G, O, last ? cnt_e : nullptr, last ? nsy_e : nullptr, last ? Pout : nullptr, 1);
      double* tmp = G; G = O; O = tmp;
    }
    if (nsq == 0 && epi >= 0) pk_buffer_loglike(G, cnt_e, nsy_e, Pout);
    *other = O;
    if (nsq_out) *nsq_out += nsq;
    return G;
  }

  // log L following LagPlan `pl`.  slots: pl.nslots + 1 packed matrices of this chain in global memory (the last
  // one is the scratch of products that replace an operand).
  MCD_HD double evaluate_pk(const double* v, const double* w, const double* lagt, const LagPlan& pl, double* slots,
                            int taylor_m, double* Pout, int* squarings_out) {
    const int n = P.n, np = P.np;
    const size_t nn = (size_t)n * n, slot_elems = pk_elems(np);
    double* scratch = slots + (size_t)plan_slots(P) * slot_elems;
    build_S(v, w);
    const double gersh = s.sc->gersh;
    ex.par([&](int tid) { s.part[tid] = 0.0; });
    int nsq = 0;
    bool bad = false;
    double* R = nullptr;
    double* O = nullptr;
    const int l0 = pl.order[0];
    const bool fast = pl.fast != 0;
    for (int c = 0; c < pl.nchain && !bad; ++c) {
      const bool is_first_lag = (c == pl.nchain - 1);
      const double t = lagt[pl.chain_a[c]] - (pl.chain_b[c] >= 0 ? lagt[pl.chain_b[c]] : 0.0);
      double* other = nullptr;
      double* G = pk_chain(t, gersh, taylor_m, is_first_lag ? l0 : -1,
                           (is_first_lag && Pout) ? Pout + (size_t)l0 * nn : nullptr, &other, &nsq);
      if (!G) { bad = true; break; }
      if (!fast && pl.chain_slot[c] >= 0) pk_copy(slots + (size_t)pl.chain_slot[c] * slot_elems, G);
      R = G;
      O = other;
    }
    // fast plan (t_i = (i+1) t_0): G0 stays in R, O receives G(2 t_0); with three or four lags G(3 t_0) = G(2 t_0) G(t_0)
    // and G(4 t_0) = G(2 t_0)^2 are needed for their likelihood terms only.  Longer progressions and general plans
    // replace an operand: product into the scratch, copy back.
    double* G0 = R;
    int in_O = -1;
    for (int i = 1; i < P.nlag && !bad; ++i) {
      const int l = pl.order[i];
      const int* cnt = P.counts + (size_t)l * np * np;
      const int* nsy = P.nsym + (size_t)l * np * np;
      double* Pl = Pout ? Pout + (size_t)l * nn : nullptr;
      if (!fast && pl.op_slot[i] < 0) {
        pk_buffer_loglike(R, cnt, nsy, Pl);
      } else if (fast) {
        const bool last = (i + 1 == P.nlag);
        if (i == 1) pk_product(G0, G0, O, cnt, nsy, Pl, last ? 0 : 1);
        else if (i == 2 && P.nlag <= 4) pk_product(O, G0, nullptr, cnt, nsy, Pl, 0);
        else if (i == 3 && P.nlag == 4) pk_product(O, O, nullptr, cnt, nsy, Pl, 0);
        else {
          pk_product(O, G0, scratch, cnt, nsy, Pl, last ? 0 : 1);
          if (!last) pk_copy(O, scratch);
        }
      } else {
        if (in_O != pl.op_slot[i]) {
          pk_copy(O, slots + (size_t)pl.op_slot[i] * slot_elems);
          in_O = pl.op_slot[i];
        }
        pk_product(R, O, scratch, cnt, nsy, Pl, 1);
        pk_copy(R, scratch);
      }
      if (!fast && pl.park[i] >= 0) pk_copy(slots + (size_t)pl.park[i] * slot_elems, R);
    }
    if (squarings_out) *squarings_out = bad ? -1 : nsq;
    if (bad) {
      ex.single([&]() { s.sc->status |= ST_NONFINITE; });
      return nan("");
    }
    return s.sc->lin + ex.sum([&](int tid) { return s.part[tid]; });
  }

  // ------------------------------------------------------------------------------------
  // Blocked squaring engine (kBlocked, kMaxSmemBins < n <= kMaxBlocks kMaxSmemBins).
  // The np x np matrices (np = nblk nbk) live in a per-CTA global workspace that stays in L2; a product
  //   C_IJ = sum_K X_KI^T Y_KJ      (I <= J; X, Y symmetric and commuting)
  // is done block by block with the nbk x nbk operand panels staged in the two shared-memory buffers
  // and an m8n8k4 loop like that of the in-smem engine.  Diagonal blocks use the symmetric warp tasks of
  // build_tables, an off-diagonal block is cut into row-major runs of 8 tiles.  The partial sums over K
  // are parked in the destination and picked up again by the same lane at the next K; the last K carries
  // the likelihood epilogue.
  // ------------------------------------------------------------------------------------
  // Workspace matrices are symmetric and live in PACKED tile storage in global memory (pk_tile / pk_off over the
  // whole np x np matrix): 180 KB instead of 353 KB per matrix at 208 bins, so that the three work matrices of all
  // resident chains (148 x 540 KB = 80 MB) stay inside the 126 MB L2 -- with full squares (157 MB) ncu showed 9 MB of
  // DRAM traffic per MC step and chain.  The shared-memory panels keep the plain row-major layout of the m8n8k4 loop:
  //   block (K, I), K < I : every tile is stored -> 16-byte cp.async copies (un-swizzling is a chunk permutation);
  //   block (K, K)        : stored tiles a <= b as above, their mirror images through registers;
  //   block (K, I), K > I : the stored block (I, K) is staged as it is and the panel is flagged TRANSPOSED: the
  //                         fragment loads then walk rows instead of columns (both patterns are conflict free at
  //                         pitch lds = 4 mod 8).
  // Returns the transposed flag.
  // global workspace data another CTA of the cluster may have written is read past L1
#define MCD_LD_PEER(p) (kCluster ? __ldcg(p) : *(p))
  MCD_HD bool blk_stage(double* dst, const double* G, int K, int I) {
    ex.par([&](int tid) {
      blk_stage_issue(tid, dst, G, K, I);
      blk_stage_wait();
    });
    MCD_CLK(25);
    return K > I;
  }
  // the copies of one thread, asynchronous on the device (cp.async); blk_stage_wait + a barrier complete them
  MCD_HD void blk_stage_issue(int tid, double* dst, const double* G, int K, int I) {
    const int nbk = P.nbk, lds = P.lds, t13 = nbk >> 3, nb8 = P.np >> 3;
    const bool tr = K > I;
    const int bi = tr ? I : K, bj = tr ? K : I;
    const bool diag = (bi == bj);
    // One warp instruction = one tile: the 32 lanes copy the 32 16-byte chunks of the 512 contiguous bytes of a stored
    // tile (16 full sectors per request).  The first version gave a lane a tile ROW and issued its four chunks one after
    // the other: every request touched 32 half-used sectors spread over four tiles, and issuing a panel took ~9 k cycles
    // (phase clocks) -- the copies were bound by L2 -> SM transactions, not bytes.
    const int warp = tid >> 5, lane = tid & 31, nwarp = Exec::nthr() / 32;
    const int r = lane >> 2, cs = lane & 3;
    const int sw = ((r >> 1) & 1) << 1;                         // stored chunk cs holds logical chunk cs ^ sw
    const int c = (cs ^ sw) * 2;                                // first of the two logical columns of this lane's chunk
    // tiles warp, warp + nwarp, ... in row-major order; (ta, tb) advance without a division
    int ta = warp / t13, tb = warp - ta * t13;
    for (int tile = warp; tile < t13 * t13; tile += nwarp, tb += nwarp) {
      while (tb >= t13) { tb -= t13; ++ta; }
      {
        if (diag && ta > tb) continue;
        const double* src = G + (size_t)pk_tile(bi * t13 + ta, bj * t13 + tb, nb8) * 64 + r * 8 + 2 * cs;
        double* d = dst + (size_t)(ta * 8 + r) * lds + tb * 8 + c;
#if defined(__CUDA_ARCH__)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(d)), "l"(src) : "memory");
#else
        d[0] = src[0]; d[1] = src[1];
#endif
        if (diag && ta < tb) {                                   // mirror image: columns c, c + 1 of row r -> tile (tb, ta)
#if defined(__CUDA_ARCH__)
          const double2 xg = MCD_LD_PEER(reinterpret_cast<const double2*>(src));
          const d2 x = d2{xg.x, xg.y};
#else
          const d2 x = ld2(src);
#endif
          dst[(size_t)(tb * 8 + c) * lds + ta * 8 + r] = x.x;
          dst[(size_t)(tb * 8 + c + 1) * lds + ta * 8 + r] = x.y;
        }
      }
    }
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
  }
  MCD_HD void blk_stage_wait() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
  }

#if defined(__CUDA_ARCH__)
  // k loop of one warp task with separate strides for the two operands (a transposed panel is walked along its rows)
  template <int NA>
  static __device__ __forceinline__ void mma_loop2(unsigned pa, unsigned pb, int stepA, int stepB, int nk4, int offA0,
                                                   int offA1, const int (&offB)[kTilesPerWarp], double (&c)[kTilesPerWarp][2]) {
#pragma unroll 2
    for (int it = 0; it < nk4; ++it) {
      const double a0 = lds64(pa + offA0);
      const double a1 = (NA < kTilesPerWarp) ? lds64(pa + offA1) : 0.0;
#pragma unroll
      for (int t = 0; t < kTilesPerWarp; ++t) {
        const double b = lds64(pb + offB[t]);
        const double a = (t < NA) ? a0 : a1;
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[t][0]), "+d"(c[t][1])
                     : "d"(a), "d"(b));
      }
      pa += stepA;
      pb += stepB;
    }
  }
#endif

  // One block: D_IJ (+)= A^T B, where the panel A holds block (K, I) of X (or, trA, its transpose (I, K)) and B block
  // (K, J) of Y (trB: (J, K)).  sym: diagonal block (upper tiles only).  D: packed global matrix, (bI, bJ) the block.
  // The partial sums over K are parked in D and picked up again by the same lane; the last K carries the epilogue.
  // pf_dst != nullptr: the copies of block (pfK, pfI) of pf_G into the panel pf_dst (not an operand of this call) are
  // issued before the k loops and completed behind them -- staging overlaps the MMAs.
  MCD_HD void blk_mma(const double* A, bool trA, const double* B, bool trB, bool sym, double* D, int bI, int bJ, bool first,
                      bool fin, const int* cnt, const int* nsy, double* Pl, double* pf_dst = nullptr,
                      const double* pf_G = nullptr, int pfK = 0, int pfI = 0, int share = 0, int nshare = 1) {
    const int n = P.n, np = P.np, lds = P.lds, nbk = P.nbk, nb8 = nbk / 8, gnb8 = np >> 3;
    const int i_off = bI * nbk, j_off = bJ * nbk;
    // full block: the nb8 x nb8 tiles in row-major order, cut into runs of kTilesPerWarp -- a run touches at most two
    // tile rows (nb8 >= kTilesPerWarp), which is what the two A fragments of the k loop serve; 22 tasks = two rounds
    // of 12 warps at 13 x 13 tiles (2-row x 4-column patches needed 28 tasks = three rounds)
    const int ntask_all = sym ? s.sc->mma_tasks : (nb8 * nb8 + kTilesPerWarp - 1) / kTilesPerWarp;
    // share / nshare: this call does the tasks [task_lo, ntask) only (a CTA's part of a block split over a cluster)
    const int tper = (ntask_all + nshare - 1) / nshare;
    const int task_lo = share * tper;
    const int ntask = (task_lo + tper < ntask_all) ? task_lo + tper : ntask_all;
    const int nwarp = Exec::nthr() / 32;
    const bool epi = fin && cnt != nullptr;
    for (int base = task_lo; base < ntask; base += nwarp) {
      ex.par([&](int tid) {
        const int warp = tid >> 5, lane = tid & 31;
        const int q = base + warp;
        const bool active = q < ntask;
        if (pf_dst && base == task_lo) blk_stage_issue(tid, pf_dst, pf_G, pfK, pfI);
        int nt = 0, na = 1, rowA = 0, rowB = 0;
        int jbs[kTilesPerWarp];
        if (sym) {
          const unsigned char* wt = s.wtab + 16 * (active ? q : 0);
          if (active) { nt = wt[0]; na = wt[1]; rowA = wt[2]; rowB = wt[3]; }
#pragma unroll
          for (int t = 0; t < kTilesPerWarp; ++t) jbs[t] = active ? wt[4 + t] : 0;
        } else {
          const int t0 = active ? q * kTilesPerWarp : 0;
          const int left = nb8 * nb8 - t0;
          nt = active ? (left < kTilesPerWarp ? left : kTilesPerWarp) : 0;
          rowA = t0 / nb8;
          const int c0 = t0 - rowA * nb8;
          na = (nb8 - c0 < nt) ? nb8 - c0 : nt;              // tiles left in the first row
          rowB = (na < nt) ? rowA + 1 : rowA;
#pragma unroll
          for (int t = 0; t < kTilesPerWarp; ++t) jbs[t] = (t < na) ? c0 + t : ((t < nt) ? t - na : 0);
          if (na == 0) na = 1;
        }
        const int fr = lane >> 2, fk = lane & 3;
        const int swz = ((fr >> 1) & 1) << 2;
        double c[kTilesPerWarp][2];
        int nsv[kTilesPerWarp][2];
#pragma unroll
        for (int t = 0; t < kTilesPerWarp; ++t) {
          const int ib = (t < na) ? rowA : rowB;
          c[t][0] = 0.0; c[t][1] = 0.0;
          if (!first && t < nt) {
            const d2 x = ld2(D + (size_t)pk_tile(bI * nb8 + ib, bJ * nb8 + jbs[t], gnb8) * 64 + fr * 8 + ((2 * fk) ^ swz));
            c[t][0] = x.x; c[t][1] = x.y;
          }
          nsv[t][0] = 0; nsv[t][1] = 0;
          if (epi && t < nt) {
            const int* qn = nsy + (size_t)(i_off + ib * 8 + fr) * np + j_off + jbs[t] * 8 + 2 * fk;
            nsv[t][0] = qn[0]; nsv[t][1] = qn[1];
          }
        }
        MCD_CLK(5);
        if (active) {
#if defined(__CUDA_ARCH__)
          // normal panel: element (k, i) at k * lds + i -> lane base fk * lds + fr, 4 lds doubles per k step, tile row/column
          // offset 8 doubles; transposed panel: element (k, i) at i * lds + k -> lane base fr * lds + fk, 4 doubles per
          // k step, tile offset 8 lds doubles
          const unsigned pa = (unsigned)__cvta_generic_to_shared(A) + (unsigned)(trA ? fr * lds + fk : fk * lds + fr) * 8u;
          const unsigned pb = (unsigned)__cvta_generic_to_shared(B) + (unsigned)(trB ? fr * lds + fk : fk * lds + fr) * 8u;
          const int stepA = trA ? 32 : 4 * lds * 8, stepB = trB ? 32 : 4 * lds * 8;
          const int tileA = trA ? 8 * lds * 8 : 64, tileB = trB ? 8 * lds * 8 : 64;
          int offB[kTilesPerWarp];
#pragma unroll
          for (int t = 0; t < kTilesPerWarp; ++t) offB[t] = jbs[t] * tileB;
          // the last K block ends with padding rows (zero in both operands): stop at the last group of four real rows
          const int nk4 = fin ? ((n - (P.nblk - 1) * nbk + 3) >> 2) : nbk / 4;
          switch (na) {
            case 1: mma_loop2<1>(pa, pb, stepA, stepB, nk4, rowA * tileA, rowB * tileA, offB, c); break;
            case 2: mma_loop2<2>(pa, pb, stepA, stepB, nk4, rowA * tileA, rowB * tileA, offB, c); break;
            case 3: mma_loop2<3>(pa, pb, stepA, stepB, nk4, rowA * tileA, rowB * tileA, offB, c); break;
            case 4: mma_loop2<4>(pa, pb, stepA, stepB, nk4, rowA * tileA, rowB * tileA, offB, c); break;
            case 5: mma_loop2<5>(pa, pb, stepA, stepB, nk4, rowA * tileA, rowB * tileA, offB, c); break;
            case 6: mma_loop2<6>(pa, pb, stepA, stepB, nk4, rowA * tileA, rowB * tileA, offB, c); break;
            case 7: mma_loop2<7>(pa, pb, stepA, stepB, nk4, rowA * tileA, rowB * tileA, offB, c); break;
            default: mma_loop2<8>(pa, pb, stepA, stepB, nk4, rowA * tileA, rowB * tileA, offB, c); break;
          }
#else
          for (int t = 0; t < nt; ++t) {
            const int ib = (t < na) ? rowA : rowB;
            const int i = ib * 8 + fr;
            for (int e = 0; e < 2; ++e) {
              const int j = jbs[t] * 8 + 2 * fk + e;
              double acc = c[t][e];
              for (int k = 0; k < nbk; ++k)
                acc = fma(trA ? A[i * lds + k] : A[k * lds + i], trB ? B[j * lds + k] : B[k * lds + j], acc);
              c[t][e] = acc;
            }
          }
#endif
        }
        MCD_CLK(6);
        // likelihood epilogue, branch-free over the 16 entries like dmma_product (clipped entries redone after)
        if (epi) {
          const double gmin = s.sc->gmin;
          double part = 0.0;
          bool rare = false;
#pragma unroll
          for (int t = 0; t < kTilesPerWarp; ++t) {
            const int ib = (t < na) ? rowA : rowB, jb = jbs[t];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int gi = i_off + ib * 8 + fr, gj = j_off + jb * 8 + 2 * fk + e;
              const bool valid = (t < nt) && !(sym && ib == jb && gj < gi);
              const double g = c[t][e];
              const bool above = g >= gmin;
              const int ns = valid ? nsv[t][e] : 0;
              rare |= (ns != 0) && !above;
              part += (double)(above ? ns : 0) * fast_log(above ? g : 1.0);
            }
          }
          if (rare || Pl) {
#pragma unroll
            for (int t = 0; t < kTilesPerWarp; ++t) {
              const int ib = (t < na) ? rowA : rowB, jb = jbs[t];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int gi = i_off + ib * 8 + fr, gj = j_off + jb * 8 + 2 * fk + e;
                if (!(t < nt) || (sym && ib == jb && gj < gi)) continue;
                if (nsv[t][e] != 0 && !(c[t][e] >= gmin))
                  part += pair_term_clipped(c[t][e], nsv[t][e], gi, gj, cnt, P.np, P.clip, s.ea, s.eb, s.hv);
                if (Pl && gi < n && gj < n) {
                  Pl[(size_t)gi * n + gj] = s.ea[gi] * c[t][e] * s.eb[gj];
                  Pl[(size_t)gj * n + gi] = s.ea[gj] * c[t][e] * s.eb[gi];
                }
              }
            }
          }
          s.part[tid] += part;
        }
        MCD_CLK(7);
        // one 8 x 8 tile = 512 contiguous bytes per warp and tile; no mirror images in packed storage
#pragma unroll
        for (int t = 0; t < kTilesPerWarp; ++t) {
          if (t < nt) {
            const int ib = (t < na) ? rowA : rowB;
            st2(D + (size_t)pk_tile(bI * nb8 + ib, bJ * nb8 + jbs[t], gnb8) * 64 + fr * 8 + ((2 * fk) ^ swz), c[t][0], c[t][1]);
          }
        }
        if (pf_dst && base + nwarp >= ntask) blk_stage_wait();      // last round: the prefetched panel is complete at the barrier
        MCD_CLK(8);
      });
      MCD_CLK(9);
    }
  }

  // Dg <- Xg Yg for symmetric commuting matrices in the global workspace (Dg distinct from both).
  // cnt != nullptr: Dg is the kernel of a lag time, its likelihood terms are added to s.part.
  // Block schedule per K: for I = 0 .. nblk-1 the panel A = X_KI stays while J sweeps I .. nblk-1.
  // Squaring (Xg == Yg): J runs DOWN to I, J == I needs no load (B = A), and the last off-diagonal panel
  // staged, X_K,I+1, is the next I's A (pointer swap): 1 + nblk (nblk-1)/2 panel loads per K (2 for nblk = 2).
  // General product: J alternates direction from one I to the next, so the B panel at the turn is reused.
  MCD_HD void blk_product(const double* Xg, const double* Yg, double* Dg, const int* cnt, const int* nsy, double* Pl) {
    const int nblk = P.nblk;
    const bool same = (Xg == Yg);
    if constexpr (kCluster) {
      // Two SMs per chain (nblk == 2): CTA r owns the diagonal block C_rr and half of the tasks of C_01; after the
      // product a cluster barrier publishes Dg.  Panels: b[i] holds the column-block i operand.
      const int r = ex.crank, o = 1 - r;
      double* b[2] = {s.XT, s.B2};
      for (int K = 0; K < 2; ++K) {
        const bool first = (K == 0), fin = (K == 1);
        if (same) {
          // b[i] = X_Ki.  C_01 = X_K0^T X_K1 first (both panels), then C_rr on b[r] while X_K+1,o streams into b[o].
          bool t[2];
          if (K == 0) { t[r] = blk_stage(b[r], Xg, 0, r); t[o] = blk_stage(b[o], Xg, 0, o); }
          else { t[o] = (1 > o); t[r] = blk_stage(b[r], Xg, 1, r); }            // X_1o was prefetched during K = 0
          blk_mma(b[0], t[0], b[1], t[1], false, Dg, 0, 1, first, fin, cnt, nsy, Pl, nullptr, nullptr, 0, 0, r, 2);
          if (K == 0) blk_mma(b[r], t[r], b[r], t[r], true, Dg, r, r, first, fin, cnt, nsy, Pl, b[o], Xg, 1, o);
          else blk_mma(b[r], t[r], b[r], t[r], true, Dg, r, r, first, fin, cnt, nsy, Pl);
        } else {
          // C_rr = X_Kr^T Y_Kr, then the CTA's half of C_01 = X_K0^T Y_K1: b[0] <- X panels, b[1] <- Y panels
          bool tx = blk_stage(b[0], Xg, K, r);
          bool ty = blk_stage(b[1], Yg, K, r);
          blk_mma(b[0], tx, b[1], ty, true, Dg, r, r, first, fin, cnt, nsy, Pl);
          if (r == 0) ty = blk_stage(b[1], Yg, K, 1);     // X_K0 stays, Y_K1 comes
          else tx = blk_stage(b[0], Xg, K, 0);            // Y_K1 stays, X_K0 comes
          blk_mma(b[0], tx, b[1], ty, false, Dg, 0, 1, first, fin, cnt, nsy, Pl, nullptr, nullptr, 0, 0, r, 2);
        }
      }
      ex.cluster_sync();
      return;
    }
    // ONE call site of blk_mma (the kernel is instruction-cache sensitive: every site inlines the eight k-loop
    // variants and the epilogue).  A squaring with 2 x 2 blocks takes the blocks in the order (0,0), (0,1), (1,1):
    // C_00 needs X_K0 only, so X_K1 streams into the other panel meanwhile; C_11 needs X_K1 only, so X_K+1,0 streams
    // in meanwhile -- every copy but the first overlaps MMAs.
    const bool pipelined = same && nblk == 2;
    double* A = s.XT;
    double* B = s.B2;
    bool trA = false, trB = false;
    bool have_A = false;       // A already holds X_KI (squarings: it was staged or prefetched as a B panel)
    for (int K = 0; K < nblk; ++K) {
      int b_has = -1;          // general product: column block of Y currently staged in B
      for (int I = 0; I < nblk; ++I) {
        if (!have_A) trA = blk_stage(A, Xg, K, I);
        have_A = false;
        const bool down = !pipelined && (same || (I & 1));
        for (int q = 0; q < nblk - I; ++q) {
          const int J = down ? nblk - 1 - q : I + q;
          const double* Bp = B;
          bool trBp = trB;
          double* pf = nullptr;
          int pfK = 0, pfI = 0;
          if (same && J == I) {
            Bp = A; trBp = trA;
            if (pipelined && I == 0) { pf = B; pfK = K; pfI = 1; }
            else if (pipelined && K + 1 < nblk) { pf = B; pfK = K + 1; pfI = 0; }
          } else if (!(pipelined && b_has == J) && (same || b_has != J)) {
            trB = blk_stage(B, Yg, K, J);
            trBp = trB;
            b_has = J;
          }
          blk_mma(A, trA, Bp, trBp, I == J, Dg, I, J, K == 0, K == nblk - 1, cnt, nsy, Pl, pf, Xg, pfK, pfI);
          if (pf) {
            if (I == 0) { b_has = 1; trB = pfK > 1; }                  // B = X_K1
            else {                                                       // B = X_K+1,0: the next K's first A
              double* tmp = A; A = B; B = tmp;
              trA = true;
              have_A = true;
            }
          }
        }
        if (same && I + 1 < nblk) {   // B holds X_K,I+1
          double* tmp = A; A = B; B = tmp;
          trA = trB;
          have_A = true;
        }
      }
    }
  }

  // likelihood terms of a kernel sitting in the global workspace (pairs i <= j)
  MCD_HD void blk_buffer_loglike(const double* G, const int* cnt, const int* nsy, double* Pl) {
    const int n = P.n, np = P.np;
    ex.par([&](int tid) {
      double part = 0.0;
      for (int e = tid; e < n * n; e += Exec::nthr()) {
        const int i = e / n, j = e - i * n;
        if (j < i) continue;
        const double g = pk_get(G, i, j);
        part += pair_term(g, nsy[(size_t)i * np + j], i, j, cnt);
        if (Pl) {
          Pl[(size_t)i * n + j] = s.ea[i] * g * s.eb[j];
          Pl[(size_t)j * n + i] = s.ea[j] * g * s.eb[i];
        }
      }
      s.part[tid] += part;
    });
  }

  MCD_HD void blk_copy(double* dst, const double* src) { pk_copy(dst, src); }

  // log L with the blocked engine.  ws: blocked_ws_doubles(P, pl.nslots) doubles of this chain.
  MCD_HD double evaluate_blk(const double* v, const double* w, const double* lagt, const LagPlan& pl, double* ws,
                             int taylor_m, double* Pout, int* squarings_out) {
    const int n = P.n, np = P.np, ldg = P.ld;
    const size_t nn = (size_t)n * n, melems = pk_elems(np);
    double* buf[3] = {ws, ws + melems, ws + 2 * melems};
    double* slots = ws + 3 * melems;
    build_S(v, w);
    const double gersh = s.sc->gersh;
    ex.par([&](int tid) { s.part[tid] = 0.0; });
    MCD_CLK(2);
    int nsq_total = 0;
    bool bad = false;
    const int l0 = pl.order[0];
    const bool fast = pl.fast != 0;
    double* R = nullptr;
    // ---- squaring chains (the G(t_0) chain last): Horner in buf[0]/buf[1], squarings rotate
    //      through the three buffers ----
    for (int c = 0; c < pl.nchain && !bad; ++c) {
      const bool is_first_lag = (c == pl.nchain - 1);
      const double t = lagt[pl.chain_a[c]] - (pl.chain_b[c] >= 0 ? lagt[pl.chain_b[c]] : 0.0);
      TaylorPlan tp;
      if (!taylor_plan(t, gersh, taylor_m, tp)) { bad = true; break; }
      const int nsq = tp.nsq;
      double* src = buf[0];
      double* dst = buf[1];
      // T = e^{-hc} T_m(h(S + cI)) straight into the workspace matrix (band rows only, the rest cleared);
      // the coefficient arrays sit in the shared-memory panels, which are free until the first product
      taylor_band<true>(s.a, s.b, s.b, tp, s.XT, src, ldg);
      MCD_CLK(4);
      double* G = src;
      double* O = dst;
      const int* cnt_e = is_first_lag ? P.counts + (size_t)l0 * np * np : nullptr;
      const int* nsy_e = is_first_lag ? P.nsym + (size_t)l0 * np * np : nullptr;
      double* Pl0 = (is_first_lag && Pout) ? Pout + (size_t)l0 * nn : nullptr;
      for (int q = 0; q < nsq; ++q) {
        const bool last = (q + 1 == nsq);
        blk_product(G, G, O, last ? cnt_e : nullptr, last ? nsy_e : nullptr, last ? Pl0 : nullptr);
        double* tmp = G; G = O; O = tmp;
      }
      if (nsq == 0 && is_first_lag && ex.crank == 0) blk_buffer_loglike(G, cnt_e, nsy_e, Pl0);
      nsq_total += nsq;
      if (!fast && pl.chain_slot[c] >= 0) {
        if (ex.crank == 0) blk_copy(slots + (size_t)pl.chain_slot[c] * melems, G);
        ex.cluster_sync();
      }
      R = G;
    }
    // ---- remaining lags: R <- R G(delta_i), out of place into one of the two buffers that do not
    //      hold G(t_0) (the operand of the fast plan) ----
    if (!bad) {
      double* G0 = R;
      double* fr[2];
      int nf = 0;
      for (int b = 0; b < 3; ++b) if (buf[b] != G0) fr[nf++] = buf[b];
      int which = 0;
      for (int i = 1; i < P.nlag; ++i) {
        const int l = pl.order[i];
        const int* cnt = P.counts + (size_t)l * np * np;
        const int* nsy = P.nsym + (size_t)l * np * np;
        double* Pl = Pout ? Pout + (size_t)l * nn : nullptr;
        if (!fast && pl.op_slot[i] < 0) {
          if (ex.crank == 0) blk_buffer_loglike(R, cnt, nsy, Pl);
        } else {
          const double* Y = fast ? G0 : slots + (size_t)pl.op_slot[i] * melems;
          double* D = fr[which];
          if (D == R) D = fr[which ^= 1];
          blk_product(R, Y, D, cnt, nsy, Pl);
          R = D;
          which ^= 1;
        }
        if (!fast && pl.park[i] >= 0) {
          if (ex.crank == 0) blk_copy(slots + (size_t)pl.park[i] * melems, R);
          ex.cluster_sync();
        }
      }
    }
    if (squarings_out) *squarings_out = bad ? -1 : nsq_total;
    if (bad) {
      ex.single([&]() { s.sc->status |= ST_NONFINITE; });
      return nan("");
    }
    const double mine = ex.sum([&](int tid) { return s.part[tid]; });
    if (ex.csize == 1) return s.sc->lin + mine;
    // cluster: the CTAs exchange their partial sums through the tail of the workspace and add them in rank order
    double* xch = ws + blocked_ws_doubles(P, plan_slots(P));
    ex.single([&]() { xch[ex.crank] = mine; });
    ex.cluster_sync();
    double tot = s.sc->lin;
#if defined(__CUDA_ARCH__)
    for (int q = 0; q < ex.csize; ++q) tot += __ldcg(xch + q);
#else
    for (int q = 0; q < ex.csize; ++q) tot += xch[q];
#endif
    return tot;
  }

  // Full likelihood of the profile (v, w) at lag times lagt.
  // mode 0: warm start from the eigenvectors in XT, 1: cold start, 2: reuse XT/lam as they are
  // (same v, w as the last solve; only the lag times changed).
  MCD_HD double evaluate(const double* v, const double* w, const double* lagt, int mode, double* Pout,
                         int* sweeps_out) {
    int sw = 0;
    if (mode != 2) {
      int cold = mode || (s.ntiles > Exec::nthr());  // warm product keeps one tile per thread
      for (int attempt = 0; attempt < 2; ++attempt) {
        build_S(v, w);
        if (cold) cold_init();
        else warm_form();
        sw = jacobi();
        if (sw >= 0 || cold) break;
        cold = 1;  // robust fallback: restart from the identity
      }
    }
    if (sweeps_out) *sweeps_out = sw;
    if (sw < 0) {
      ex.single([&]() { s.sc->status |= ST_NONFINITE; });
      return nan("");
    }
    return loglike(lagt, Pout);
  }

  MCD_HD double string_energy(const double* w) {
    // k/2 sum (dw)^2 (+ periodic wrap)   lib/utils.py:283-292
    const int dim_w = P.dim_w;
    double e = ex.sum([&](int tid) {
      double acc = 0.0;
      for (int i = tid; i < dim_w - 1; i += Exec::nthr()) {
        const double d = w[i + 1] - w[i];
        acc += d * d;
      }
      if (tid == 0 && P.pbc) {
        const double d = w[0] - w[dim_w - 1];
        acc += d * d;
      }
      return acc;
    });
    return 0.5 * P.spring_k * e;
  }

  // ------------------------------------------------------------------------------------
  // The MC loop of lib/mc.py:27-51 for one chain.
  // ------------------------------------------------------------------------------------
  MCD_HD void run(const McParams& mp, const ChainIO& io, long long chain, int nsteps, double* sq_slots) {
    const int n = P.n, np = P.np, ld = P.ld, dim_w = P.dim_w;
    const int ncosF = P.ncosF, ncosD = P.ncosD;
    const int sdim = sample_dim(P);
    Scalars* sc = s.sc;
    build_tables();
    // ---- load chain state ----
    ex.par([&](int tid) {
      for (int i = tid; i < np; i += Exec::nthr()) {
        s.v[i] = (i < n) ? io.v[chain * n + i] : 0.0;
        s.w[i] = (i < dim_w) ? io.w[chain * dim_w + i] : 0.0;
      }
      for (int i = tid; i < ncosF; i += Exec::nthr()) s.vc[i] = io.vcoef[chain * ncosF + i];
      for (int i = tid; i < ncosD; i += Exec::nthr()) s.wc[i] = io.wcoef[chain * ncosD + i];
      if (tid == 0) {
        const double* q = io.scal + chain * S_NSCAL;
        sc->ll_cur = q[S_LL]; sc->timezero = q[S_T0]; sc->dv = q[S_DV]; sc->dw = q[S_DW];
        sc->dt0 = q[S_DT0]; sc->temp = q[S_TEMP];
        const long long* c = io.counters + chain * C_NCOUNT;
        sc->naccv = c[C_NACCV]; sc->naccw = c[C_NACCW]; sc->nacct0 = c[C_NACCT0];
        sc->naccv_upd = c[C_NACCV_UPD]; sc->naccw_upd = c[C_NACCW_UPD];
        sc->step = c[C_STEP]; sc->status = c[C_STATUS]; sc->sweeps = c[C_SWEEPS];
        sc->nsamp = 0; sc->e_spring = 0.0;
      }
    });
    ex.par([&](int tid) {
      for (int l = tid; l < P.nlag; l += Exec::nthr()) s.lag[l] = P.lagtimes[l] + sc->timezero;
    });
    int have_basis = 0;
    if (kMode == 0 && io.basis && io.basis_valid[chain]) {
      const double* src = io.basis + (size_t)chain * np * ld;
      ex.par([&](int tid) {
        for (int e = tid; e < np * ld; e += Exec::nthr()) s.XT[e] = src[e];
      });
      have_basis = 1;
    }
    long long since_refresh = 0;
    const bool shifted_plan = mp.move_timezero || sc->timezero != 0.0;
    MCD_CLK_BEGIN();

    for (int it = 0; it < nsteps; ++it) {
      // Two phases per step with ONE inlined likelihood evaluation site:
      // phase 0 = potential/diffusion move (lib/mc.py:33-39), phase 1 = time-offset move
      // (lib/mc.py:40-42).
      int move = 0;
      for (int phase = 0; phase < 2; ++phase) {
        if (phase == 1 && !(mp.move_timezero && sc->dt0 > 0.0)) break;
        // ---- proposal (thread 0): lib/MCState.py:235-245, 269-280, 206-208 ----
        ex.single([&]() {
          const long long imc = sc->step;
          const long long ts = (long long)it + mp.tape_stride * chain;
          const uint64_t cid = (uint64_t)(chain + mp.chain_offset);
          if (phase == 0) {
            double u_choice, u_delta;
            uint32_t r1[4] = {0, 0, 0, 0};
            if (mp.use_tape) {
              u_choice = io.tape_u[ts * 5 + 0];
              u_delta = io.tape_u[ts * 5 + 1];
            } else {
              uint32_t r0[4];
              philox4x32_10((uint32_t)imc, (uint32_t)((uint64_t)imc >> 32), (uint32_t)cid, 0u, (uint32_t)mp.seed,
                            (uint32_t)(mp.seed >> 32), r0);
              philox4x32_10((uint32_t)imc, (uint32_t)((uint64_t)imc >> 32), (uint32_t)cid, 1u, (uint32_t)mp.seed,
                            (uint32_t)(mp.seed >> 32), r1);
              u_choice = u53(r0[0], r0[1]);
              u_delta = u53(r0[2], r0[3]);
            }
            int mv = 0, lo = 0, range = 1;
            if (u_choice < 0.5 && ncosF != 1 && sc->dv > 0.0) {
              mv = 1;
              if (ncosF <= 0) { lo = 0; range = n; } else { lo = 1; range = ncosF - 1; }
            } else if (sc->dw > 0.0) {
              mv = 2;
              if (ncosD <= 0) { lo = 0; range = dim_w; } else { lo = 0; range = ncosD; }
            }
            sc->move = mv;
            sc->idx = mp.use_tape ? io.tape_idx[ts] : lo + (int)(((uint64_t)r1[0] * (uint64_t)range) >> 32);
            sc->delta = (mv == 1 ? sc->dv : sc->dw) * (u_delta - 0.5);
            sc->accepted = 0;
            sc->accepted_t0 = 0;
            sc->do_t0 = 0;
          } else {
            double u_d;
            if (mp.use_tape) {
              u_d = io.tape_u[ts * 5 + 3];
            } else {
              uint32_t r2[4];
              philox4x32_10((uint32_t)imc, (uint32_t)((uint64_t)imc >> 32), (uint32_t)cid, 2u, (uint32_t)mp.seed,
                            (uint32_t)(mp.seed >> 32), r2);
              u_d = u53(r2[0], r2[1]);
            }
            sc->t0_try = sc->timezero + sc->dt0 * (u_d - 0.5);
            sc->do_t0 = (sc->t0_try > -0.5 * mp.min_lt) ? 1 : 0;  // lib/MCState.py:207
          }
        });
        MCD_CLK(0);
        if (phase == 0) move = sc->move;
        if (phase == 0 && move == 0) continue;
        if (phase == 1 && !sc->do_t0) break;

        // ---- trial profile / trial lag times ----
        const double* vtry = s.v;
        const double* wtry = s.w;
        const double* lagp = s.lag;
        int mode;
        if (phase == 0) {
          const int idx = sc->idx;
          const double delta = sc->delta;
          ex.par([&](int tid) {
            if (move == 1) {
              if (ncosF <= 0) {
                for (int i = tid; i < np; i += Exec::nthr()) s.vt[i] = s.v[i] + (i == idx ? delta : 0.0);
              } else {
                for (int c = tid; c < ncosF; c += Exec::nthr()) s.ct[c] = s.vc[c] + (c == idx ? delta : 0.0);
              }
            } else {
              if (ncosD <= 0) {
                if (P.spring_k > 0.0) {  // move along a Laplacian eigenvector, lib/MCState.py:269-271
                  for (int i = tid; i < np; i += Exec::nthr())
                    s.wt[i] = (i < dim_w) ? s.w[i] + delta * P.svecs[(size_t)i * dim_w + idx] : 0.0;
                } else {
                  for (int i = tid; i < np; i += Exec::nthr()) s.wt[i] = s.w[i] + (i == idx ? delta : 0.0);
                }
              } else {
                for (int c = tid; c < ncosD; c += Exec::nthr()) s.ct[c] = s.wc[c] + (c == idx ? delta : 0.0);
              }
            }
          });
          if ((move == 1 && ncosF > 0) || (move == 2 && ncosD > 0)) {
            ex.par([&](int tid) {  // profile = sum_k coeff_k basis[:,k], lib/model.py:128-130
              const int nc = (move == 1) ? ncosF : ncosD;
              const int lim = (move == 1) ? n : dim_w;
              const double* basis = (move == 1) ? P.v_basis : P.w_basis;
              double* dst = (move == 1) ? s.vt : s.wt;
              for (int i = tid; i < np; i += Exec::nthr()) {
                double acc = 0.0;
                if (i < lim)
                  for (int c = 0; c < nc; ++c) acc += s.ct[c] * basis[(size_t)i * nc + c];
                dst[i] = acc;
              }
            });
          }
          if (move == 1) vtry = s.vt; else wtry = s.wt;
          const int cold = (!have_basis) || (mp.refresh_every > 0 && since_refresh >= mp.refresh_every);
          if (cold) since_refresh = 0;
          ++since_refresh;
          mode = cold ? 1 : 0;
        } else {
          const double t0 = sc->t0_try;
          ex.par([&](int tid) {
            for (int l = tid; l < P.nlag; l += Exec::nthr()) s.lag_try[l] = P.lagtimes[l] + t0;
          });
          lagp = s.lag_try;
          // XT holds the eigenvectors of the last TRIAL matrix: reusable as they are only when
          // that trial was accepted (it then is the current state)
          mode = (move != 0 && sc->accepted && have_basis) ? 2 : (have_basis ? 0 : 1);
        }

        MCD_CLK(1);
        int sw = 0;
        double ll_new;
        // plan[0] assumes t_i - t_{i-1} equals an earlier lag time, which only holds for the lag times as given:
        // a chain that carries a time offset (from --t0 moves now, or from an earlier run / set_state) walks plan[1]
        const LagPlan& plan = P.plan[shifted_plan ? 1 : 0];
        if constexpr (kBlocked) {
          ll_new = evaluate_blk(vtry, wtry, lagp, plan, sq_slots, mp.taylor_m, nullptr, &sw);
        } else if constexpr (kGeneral) {
          ll_new = evaluate_gen(vtry, wtry, lagp, plan, sq_slots, mp.taylor_m, nullptr, &sw);
        } else if constexpr (kPacked) {
          ll_new = evaluate_pk(vtry, wtry, lagp, plan, sq_slots, mp.taylor_m, nullptr, &sw);
        } else if constexpr (kTensorOnly) {
          ll_new = evaluate_sq(vtry, wtry, lagp, plan, sq_slots, mp.taylor_m, true, nullptr, &sw);
        } else if (mp.method >= 2) {
          ll_new = evaluate_sq(vtry, wtry, lagp, plan, sq_slots, mp.taylor_m, mp.method == 2, nullptr, &sw);
        } else {
          ll_new = evaluate(vtry, wtry, lagp, mode, nullptr, &sw);
          have_basis = (sw >= 0);
        }
        double e_try = 0.0;
        if (phase == 0 && move == 2 && P.spring_k > 0.0) e_try = string_energy(s.wt);  // lib/MCState.py:287-289

        // ---- Metropolis (thread 0): lib/MCState.py:251-261, 285-301, 213-221 ----
        ex.single([&]() {
          sc->sweeps += (sw > 0 ? sw : 0);
          double lt = ll_new;
          if (lt == lt) {
            lt -= e_try;
            const long long imc = sc->step;
            const long long ts = (long long)it + mp.tape_stride * chain;
            double u_acc;
            if (mp.use_tape) {
              u_acc = io.tape_u[ts * 5 + (phase == 0 ? 2 : 4)];
            } else {
              uint32_t rr[4];
              const uint64_t cid = (uint64_t)(chain + mp.chain_offset);
              philox4x32_10((uint32_t)imc, (uint32_t)((uint64_t)imc >> 32), (uint32_t)cid, phase == 0 ? 1u : 2u,
                            (uint32_t)mp.seed, (uint32_t)(mp.seed >> 32), rr);
              u_acc = u53(rr[2], rr[3]);
            }
            const double dlog = lt - sc->ll_cur;
            if (u_acc < exp(dlog / sc->temp)) {
              sc->ll_cur = lt;
              if (phase == 0) {
                sc->accepted = 1;
                if (move == 1) { sc->naccv += 1; sc->naccv_upd += 1; }
                else { sc->naccw += 1; sc->naccw_upd += 1; }
              } else {
                sc->accepted_t0 = 1;
                sc->timezero = sc->t0_try;
                sc->nacct0 += 1;
              }
            }
          } else {
            sc->status |= ST_NAN_TRIAL;
          }
          if (phase == 0) sc->ll_try = lt;
        });
        MCD_CLK(12);
        // ---- commit ----
        if (phase == 0 && sc->accepted) {
          const int idx = sc->idx;
          ex.par([&](int tid) {
            if (move == 1) {
              for (int i = tid; i < np; i += Exec::nthr()) s.v[i] = s.vt[i];
              if (ncosF > 0) {
                for (int c = tid; c < ncosF; c += Exec::nthr()) s.vc[c] = s.ct[c];
                if (tid == 0 && ex.crank == 0) io.naccv_coeff[chain * ncosF + idx] += 1;
              }
            } else {
              for (int i = tid; i < np; i += Exec::nthr()) s.w[i] = s.wt[i];
              if (ncosD > 0) {
                for (int c = tid; c < ncosD; c += Exec::nthr()) s.wc[c] = s.ct[c];
                if (tid == 0 && ex.crank == 0) io.naccw_coeff[chain * ncosD + idx] += 1;
              }
            }
          });
        } else if (phase == 1 && sc->accepted_t0) {
          ex.par([&](int tid) {
            for (int l = tid; l < P.nlag; l += Exec::nthr()) s.lag[l] = s.lag_try[l];
          });
        }
      }
      // ---- per-step records, sampling (lib/log.py:35-58), adaptation (lib/MCState.py:371-393) ----
      ex.par([&](int tid) {
        const long long j = sc->step + 1;
        if (ex.crank != 0) return;          // the CTAs of a cluster hold identical copies of the chain: rank 0 reports
        if (tid == 0) {
          if (io.decisions)
            io.decisions[(size_t)chain * nsteps + it] =
                (unsigned char)((sc->accepted ? 1 : 0) | (sc->accepted_t0 ? 2 : 0) | (move << 2));
          if (io.ll_try) io.ll_try[(size_t)chain * nsteps + it] = (move != 0) ? sc->ll_try : nan("");
        }
        if (io.samples && mp.sample_every > 0 && (j % mp.sample_every) == 0 && sc->nsamp < mp.nsamples_cap) {
          double* row = io.samples + ((size_t)chain * mp.nsamples_cap + sc->nsamp) * sdim;
          if (tid == 0) {
            row[0] = sc->ll_cur; row[1] = sc->timezero; row[2] = sc->dv; row[3] = sc->dw;
            row[4] = sc->dt0; row[5] = sc->temp;
          }
          const int nv = ncosF > 0 ? ncosF : n, nw = ncosD > 0 ? ncosD : dim_w;
          const double* pv = ncosF > 0 ? s.vc : s.v;
          const double* pw = ncosD > 0 ? s.wc : s.w;
          for (int i = tid; i < nv; i += Exec::nthr()) row[6 + i] = pv[i];
          for (int i = tid; i < nw; i += Exec::nthr()) row[6 + nv + i] = pw[i];
        }
      });
      ex.single([&]() {
        const long long j = sc->step + 1;
        if (io.samples && mp.sample_every > 0 && (j % mp.sample_every) == 0 && sc->nsamp < mp.nsamples_cap)
          sc->nsamp += 1;
        if (mp.num_mc_update > 0.0 && fmod((double)j, mp.num_mc_update) == 0.0) {
          if (ncosF != 1) sc->dv *= exp(0.1 * ((double)sc->naccv_upd / mp.num_mc_update - 0.3));
          sc->dw *= exp(0.1 * ((double)sc->naccw_upd / mp.num_mc_update - 0.3));
          sc->naccv_upd = 0;
          sc->naccw_upd = 0;
          if (mp.dtemp != 0.0) sc->temp += mp.dtemp;
        }
        sc->step = j;
      });
      MCD_CLK(13);
    }
    MCD_CLK_END();
    // ---- store chain state ----
    ex.par([&](int tid) {
      if (ex.crank != 0) return;
      for (int i = tid; i < n; i += Exec::nthr()) io.v[chain * n + i] = s.v[i];
      for (int i = tid; i < dim_w; i += Exec::nthr()) io.w[chain * dim_w + i] = s.w[i];
      for (int i = tid; i < ncosF; i += Exec::nthr()) io.vcoef[chain * ncosF + i] = s.vc[i];
      for (int i = tid; i < ncosD; i += Exec::nthr()) io.wcoef[chain * ncosD + i] = s.wc[i];
      if (tid == 0) {
        double* q = io.scal + chain * S_NSCAL;
        q[S_LL] = sc->ll_cur; q[S_T0] = sc->timezero; q[S_DV] = sc->dv; q[S_DW] = sc->dw;
        q[S_DT0] = sc->dt0; q[S_TEMP] = sc->temp;
        long long* c = io.counters + chain * C_NCOUNT;
        c[C_NACCV] = sc->naccv; c[C_NACCW] = sc->naccw; c[C_NACCT0] = sc->nacct0;
        c[C_NACCV_UPD] = sc->naccv_upd; c[C_NACCW_UPD] = sc->naccw_upd;
        c[C_STEP] = sc->step; c[C_STATUS] = sc->status; c[C_SWEEPS] = sc->sweeps;
      }
      if (io.basis && have_basis) {
        double* dst = io.basis + (size_t)chain * np * ld;
        for (int e = tid; e < np * ld; e += Exec::nthr()) dst[e] = s.XT[e];
        if (tid == 0) io.basis_valid[chain] = 1;
      }
    });
  }
};

}  // namespace MCD_NS
