/* mcdiff_b200 -- C ABI of the B200 Monte-Carlo likelihood engine.
 *
 * Drop-in boundary for the two seams of annekegh/mcdiff's hot path.  The
 * reference has no FFI: both seams are plain Python calls, so each entry point below names
 * the reference function it replaces.
 *
 *   mcd_loglike_batch  <->  mcdiff.utils.log_like_lag            (lib/utils.py:355-368)
 *                           as called by MCState.init_log_like   (lib/MCState.py:101-137)
 *                           (+ scipy.linalg.expm propagators,     lib/utils.py:313)
 *   mcd_mc_run         <->  mcdiff.mc.do_mc_cycles                (lib/mc.py:20-51)
 *                           = MCState.mcmove_potential / mcmove_diffusion / mcmove_timezero /
 *                             update_movewidth / update_temp      (lib/MCState.py:205-304,371-393)
 *                           + Logger.log sampling                 (lib/log.py:35-58)
 *   mcd_rad_loglike_batch <-> mcdiff.twod.rad_log_like_lag        (lib/twod.py:13-32)
 *   mcd_traj_counts[_rad] <-> mcdiff.tools.extract.transition_matrix_add1 / count_2D  (lib/tools/extract.py:71-156)
 *   mcd_propagator_batch  <-> scipy.linalg.expm(rate * t) of the permeability analyses (lib/permeability/mfpt.py:364-475)
 *
 * Conventions
 *   - plain C, no exceptions, no stdout; every function returns 0 on success, a negative
 *     MCD_E_* code for argument errors, or a positive cudaError_t value.
 *   - every array argument is a DEVICE pointer owned by the caller (row-major, contiguous)
 *     unless the name ends in _host.  The library keeps private copies of the problem
 *     description (mcd_problem) and never frees caller memory.
 *   - all compute calls are asynchronous on the cudaStream_t passed as `stream` (0 = default
 *     stream) and never synchronise; the caller synchronises.  Per-call scratch memory is a stream-ordered
 *     allocation (cudaMallocAsync / cudaFreeAsync on `stream`), so calls on different streams that share a
 *     handle or a problem do not interfere.  The *_create calls synchronise `stream` once.
 *     One handle per device; a handle is not thread-safe (it counts launches), problems are read-only.
 *   - counts[l][i][j] = number of observed transitions j -> i at lag l   (lib/tools/extract.py:86-87)
 *     w[i] lives between bin i and i+1 (pbc: w[n-1] couples bin n-1 and bin 0)  (lib/utils.py:96-97)
 */
#ifndef MCDIFF_B200_H
#define MCDIFF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCD_VERSION 200

enum {
  MCD_OK = 0,
  MCD_E_ARG = -1,        /* bad argument (null pointer, size out of range) */
  MCD_E_UNSUPPORTED = -2,/* e.g. n > 508, --pull without pbc or with n > 104, an engine that does not apply */
  MCD_E_NOMEM = -3,
  MCD_E_NODEVICE = -4
};

/* per-profile / per-chain status bits */
enum {
  MCD_ST_NONFINITE = 1,  /* the propagator engine rejected the profile: non-finite v / w (eigensolver did not
                            converge; squaring engines: lag * ||S|| non-finite or > 1e9); log L is NaN */
  MCD_ST_EIG_NOCONV = 1, /* (older name of the same bit) */
  MCD_ST_NAN_TRIAL = 2,  /* a trial log-likelihood was NaN; the move was skipped like
                            lib/MCState.py:251 does */
  MCD_ST_BAD_LAGS = 4    /* mcd_loglike_batch on a squaring engine: per-profile lag times that are not the
                            problem's lag times plus one offset (only that form has a LagPlan) */
};

/* Propagator engines.  Both produce P_l = expm(lag_l R) (lib/utils.py:313):
 *  EIGEN    : in-CTA parallel Jacobi eigendecomposition of the detailed-balance-symmetrised
 *             generator, P = Pi^1/2 V exp(lag Lambda) V^T Pi^-1/2; any lag times, time-offset moves.
 *  SQUARING : symmetric scaling-and-squaring, all products as FP64 tensor-core (mma.sync m8n8k4)
 *             GEMMs on shared-memory operands.  Lag times are visited in ascending order,
 *             G(t_i) = G(t_{i-1}) G(t_i - t_{i-1}); every distinct operand is an earlier G(t_j) or gets its
 *             own squaring chain (at most 8), so arithmetic progressions (10,20,30,40 / a single lag) need
 *             one chain, time-offset moves (--t0) two.  n <= 104: work matrices in shared memory;
 *             104 < n <= 508: blocked (matrices in a per-chain global workspace, nblk x nblk block products,
 *             nblk = ceil(n / 104) <= 5, through shared-memory operand panels; tensor-core products only,
 *             no SQUARING_DFMA). */
enum { MCD_METHOD_AUTO = 0, MCD_METHOD_EIGEN = 1, MCD_METHOD_SQUARING = 2,
       MCD_METHOD_SQUARING_DFMA = 3 /* same algorithm, products on the FP64 vector pipe (A/B reference) */ };

typedef struct mcd_handle mcd_handle;
typedef struct mcd_problem mcd_problem;

/* Static description of one inference problem: what Transitions + Model hold in the
 * reference (lib/transitions.py:22-70, lib/model.py:29-149). */
typedef struct {
  int32_t n;        /* number of bins = dim_v = dim_trans */
  int32_t pbc;      /* 1: periodic (dim_w = n), 0: reflecting ends (dim_w = n-1) */
  int32_t nlag;     /* number of lag times / count matrices (<= 64) */
  int32_t ncosF;    /* 0: per-bin F moves, >0: v = v_basis . v_coeff  (<= 64) */
  int32_t ncosD;    /* 0: per-bin ln D moves, >0: w = w_basis . w_coeff (<= 64) */
  int32_t use_pull; /* 1: constant pulling force (--pull): generator of init_rate_matrix_pbc_pull,
                       lib/utils.py:133-183 -- not symmetrisable, runs on the general squaring engine;
                       needs pbc, n <= 104 and a lag set with a LagPlan (else MCD_E_UNSUPPORTED) */
  const double* lagtimes;  /* [nlag]                       device */
  const int32_t* counts;   /* [nlag][n][n]                 device */
  const double* v_basis;   /* [n][ncosF] or NULL           device (lib/model.py:93-96) */
  const double* w_basis;   /* [dim_w][ncosD] or NULL       device (lib/model.py:98-102) */
  const double* string_vecs; /* [dim_w][dim_w] or NULL     device (lib/utils.py:294-305) */
  double spring_k;         /* -k option; <= 0 disables     (lib/MCState.py:121-124,287-289) */
  double clip;             /* lower bound of P before log: 1e-10 (lib/utils.py:317) */
  double pull;             /* dF between neighbouring bins = -pull * dz (lib/MCState.py:65-67); read when use_pull */
} mcd_problem_desc;

/* Monte-Carlo schedule: what MCState.set_MC_params holds (lib/MCState.py:29-61). */
typedef struct {
  double num_mc_update;  /* --nmc_update (float in the reference CLI); 0 = never adapt */
  double dtemp;          /* (Tend - T)/(nmc/nmc_update - 1), lib/MCState.py:46-51 */
  double min_lt;         /* smallest lag time, lib/transitions.py:39 */
  int32_t move_timezero; /* --t0 */
  int32_t sample_every;  /* Logger.freq = 100, lib/log.py:9; 0 = no samples */
  int32_t refresh_every; /* cold eigen re-solve period (0 = never); 512 is a good default */
  int32_t use_tape;      /* 1: proposals and uniforms come from tape_u / tape_idx */
  int32_t method;        /* MCD_METHOD_*: propagator engine (AUTO picks SQUARING when it applies) */
  int32_t taylor_m;      /* Taylor order of the squaring engine, 0 = adaptive (smallest order with a
                            truncation error <= 1e-17 after the squarings, 6..24); values > 31 are clamped */
  uint64_t seed;         /* Philox4x32-10 key */
  int64_t chain_offset;  /* global index of chain 0 (chains shard over GPUs) */
  int64_t tape_stride;   /* 0: one tape for all chains, else per-chain stride (in steps) */
  int64_t nsamples_cap;  /* rows per chain in `samples` */
} mcd_mc_params;

#define MCD_NSCAL 8   /* scal[c] = {log_like, timezero, dv, dw, dtimezero, temp, -, -} */
#define MCD_NCOUNT 8  /* counters[c] = {naccv, naccw, nacctimezero, naccv_update, naccw_update,
                                        step (imc), status bits, eigen sweeps} */

/* Per-chain state (structure of arrays, C = number of chains), all device pointers. */
typedef struct {
  double* v;            /* [C][n] */
  double* w;            /* [C][dim_w] */
  double* vcoef;        /* [C][ncosF]  (ignored when ncosF == 0) */
  double* wcoef;        /* [C][ncosD]  (ignored when ncosD == 0) */
  double* scal;         /* [C][MCD_NSCAL] */
  int64_t* counters;    /* [C][MCD_NCOUNT] */
  int64_t* naccv_coeff; /* [C][max(ncosF,1)]  lib/MCState.py:130-133 */
  int64_t* naccw_coeff; /* [C][max(ncosD,1)] */
  double* basis;        /* [C][basis_elems] warm-start eigenvectors kept between launches, or NULL */
  int32_t* basis_valid; /* [C] */
  const double* tape_u;   /* [nsteps][5] = u_choice,u_delta,u_accept,u_t0_delta,u_t0_accept (use_tape) */
  const int32_t* tape_idx;/* [nsteps]    = bin / coefficient index of the move           (use_tape) */
  uint8_t* decisions;   /* [C][nsteps] or NULL: bit0 move accepted, bit1 t0 accepted, bits2-3 move
                           type (1 potential, 2 diffusion) */
  double* ll_try;       /* [C][nsteps] or NULL: trial log-likelihood of the main move */
  double* samples;      /* [C][nsamples_cap][sample_dim] or NULL; row = {log_like, timezero, dv, dw,
                           dtimezero, temp, v|v_coeff..., w|w_coeff...}  (lib/log.py:35-58) */
} mcd_chain_io;

int mcd_version(void);
const char* mcd_error_string(int code);

int mcd_create(int device, mcd_handle** out);
int mcd_destroy(mcd_handle* h);

/* Copies the description into library-owned device memory (counts are re-packed to a padded
 * int32 layout, N + N^T and the row/column sums are formed once).  Synchronises `stream` once (the lag
 * times are read back to build the LagPlans).  desc->use_pull selects the general (non-symmetric) engine. */
int mcd_problem_create(mcd_handle* h, const mcd_problem_desc* desc, void* stream, mcd_problem** out);
int mcd_problem_destroy(mcd_problem* p);
/* np: padded bin count, basis_elems: doubles per chain in mcd_chain_io.basis,
 * sample_dim: doubles per sample row, in_smem: 1 when the n x n work matrices fit in shared memory,
 * squaring_ok: 1 when MCD_METHOD_SQUARING applies to this problem (see above) */
int mcd_problem_info(const mcd_problem* p, int32_t* np, int64_t* basis_elems, int32_t* sample_dim,
                     int32_t* in_smem, int32_t* squaring_ok);

/* log L of B profiles: out_ll[b] = sum_l sum_ij counts[l,i,j] log(max(P_l[i,j], clip)) with
 * P_l = expm(lag_l R(v_b, w_b)).  lagtimes == NULL uses the problem's lag times, otherwise
 * [B][nlag] per-profile lag times (time-offset moves): AUTO = EIGEN for them (any lag times); an explicit
 * MCD_METHOD_SQUARING (and every use_pull problem) accepts the problem's lag times plus one offset per profile
 * and flags anything else with MCD_ST_BAD_LAGS.  out_P (optional) receives
 * [B][nlag][n][n] propagators, status (optional) [B] status bits.  method: MCD_METHOD_AUTO = SQUARING when
 * it applies to the problem and no per-profile lag times are given (componentwise accurate), else EIGEN;
 * an explicit MCD_METHOD_EIGEN is the independent cross-check of the squaring engine.
 * Replaces log_like_lag, lib/utils.py:355-368. */
int mcd_loglike_batch(mcd_handle* h, const mcd_problem* p, int32_t method, int32_t B, const double* v,
                      const double* w, const double* lagtimes, double* out_ll, double* out_P, int32_t* status,
                      void* stream);

/* Advance every chain by nsteps Monte-Carlo steps (one CTA per chain).
 * Replaces the loop body of do_mc_cycles, lib/mc.py:27-51. */
int mcd_mc_run(mcd_handle* h, const mcd_problem* p, const mcd_mc_params* params, const mcd_chain_io* io,
               int32_t nchains, int32_t nsteps, void* stream);

/* Static description of a radial (2-D) problem: what RadTransitions + RadModel + the Bessel table of
 * MCState.set_model hold (lib/transitions.py, lib/model.py:316-365, lib/twod.py:34-63).  v and w are frozen in
 * radial mode (lib/mc.py:29-30). */
typedef struct {
  int32_t n;          /* bins along z */
  int32_t pbc;
  int32_t nlag;       /* <= 64 */
  int32_t dim_rad;    /* radial bins */
  int32_t lmax;       /* Bessel terms, <= 64 */
  int32_t ncosDrad;   /* 0: per-bin moves of wrad, >0: wrad = wrad_basis . wrad_coeff */
  const double* v;         /* [n]                          device */
  const double* w;         /* [dim_w]                      device */
  const double* lagtimes;  /* [nlag]                       device */
  const double* counts;    /* [nlag][dim_rad][n][n] double device (lib/reading.py:129) */
  const double* bessels;   /* [lmax][dim_rad]              device (lib/twod.py:45-59) */
  const double* b0zeros;   /* [lmax]                       device */
  const double* wrad_basis;/* [n][ncosDrad] or NULL        device (lib/model.py:349-352) */
  double clip;             /* 1e-32 (lib/twod.py:16-23) */
} mcd_rad_problem_desc;

typedef struct mcd_rad_problem mcd_rad_problem;

/* Private device copies, the LagPlan and the pair-major copy of the counts are made once here
 * (synchronises `stream` once).  MCD_E_UNSUPPORTED when one Bessel term does not fit in shared memory. */
int mcd_rad_problem_create(mcd_handle* h, const mcd_rad_problem_desc* desc, void* stream, mcd_rad_problem** out);
int mcd_rad_problem_destroy(mcd_rad_problem* p);
/* np: padded bins, sample_dim: doubles per sample row, nbatch: Bessel terms advanced in lockstep,
 * squaring_ok: 1 when the lag set has a LagPlan (else the eigen fallback runs) */
int mcd_rad_problem_info(const mcd_rad_problem* p, int32_t* np, int32_t* sample_dim, int32_t* nbatch,
                         int32_t* squaring_ok);

/* Radial likelihood, lib/twod.py:13-32,89-130: for each of B radial profiles wrad [B][n]
 * out_ll[b] = sum_l sum_{k,i,j} counts[l,k,i,j] log(max(P_l[k,i,j], clip)),
 * P_l[k] = sum_m bessels[m,k] expm(lag_l (R(v,w) - diag(exp(wrad_b) b0zeros[m]^2 / dim_rad^2))).
 * out_P optional [B][nlag][dim_rad][n][n], status optional [B].  One kernel launch, asynchronous.
 * Replaces rad_log_like_lag, lib/twod.py:13. */
int mcd_rad_loglike_batch(mcd_handle* h, const mcd_rad_problem* p, int32_t B, const double* wrad, double* out_ll,
                          double* out_P, int32_t* status, void* stream);

/* Radial Monte-Carlo loop (only wrad is sampled; v and w are frozen), one CTA per chain.
 * Replaces the lmax > 0 branch of do_mc_cycles (lib/mc.py:29-30) = MCState.mcmove_diffusion_radial
 * (lib/MCState.py:306-340) + the dwrad update (lib/MCState.py:383-386). */
typedef struct {
  double num_mc_update;
  double dtemp;
  int32_t sample_every;
  int32_t use_tape;
  int32_t ncosDrad;      /* 0: per-bin moves of wrad, >0: wrad = wrad_basis . wrad_coeff */
  int32_t reserved;
  uint64_t seed;
  int64_t chain_offset;
  int64_t tape_stride;
  int64_t nsamples_cap;
} mcd_rad_mc_params;

typedef struct {
  double* wrad;            /* [C][n] */
  double* wcoef;           /* [C][ncosDrad] */
  double* scal;            /* [C][8]  {log_like, dwrad, temp, ...} */
  int64_t* counters;       /* [C][8]  {naccwrad, naccwrad_update, step, status, ...} */
  int64_t* nacc_coeff;     /* [C][max(ncosDrad,1)] */
  const double* wrad_basis;/* ignored: the problem's private copy is used */
  const double* tape_u;    /* [nsteps][5]: column 1 = delta uniform, column 2 = accept uniform */
  const int32_t* tape_idx; /* [nsteps] */
  uint8_t* decisions;      /* [C][nsteps] or NULL */
  double* ll_try;          /* [C][nsteps] or NULL */
  double* samples;         /* [C][nsamples_cap][3 + (ncosDrad ? ncosDrad : n)] = {log_like, dwrad, temp, wrad|coeff...} */
} mcd_rad_chain_io;

int mcd_rad_mc_run(mcd_handle* h, const mcd_rad_problem* p, const mcd_rad_mc_params* params,
                   const mcd_rad_chain_io* io, int32_t nchains, int32_t nsteps, void* stream);

/* ---- propagators with absorbing boundaries (the permeability analyses) -----------------
 *
 *   mcd_propagator_batch <->  scipy.linalg.expm(rate * t) on the rate matrices of
 *                             mcdiff.utils.init_rate_matrix(n, v, w, pbc, st=, end=, side=)   (lib/utils.py:28-74)
 *                             as lib/permeability/mfpt.py:364,373,475 (exit-time distributions over a time grid),
 *                             lib/permeability/perm.py:391 (transient partitioning) and lib/MCState.py:353 call it.
 *
 * p: a problem created with ONE lag time (its counts are not used; zeros will do).  Item b of the batch is the
 * generator R(v[b], w[b]) - diag(sink[b]) at time times[b]:  out_P[b] = exp(times[b] * (R - diag(sink[b]))), [B][n][n],
 * out_P[b][i][j] = probability of j -> i.  sink[B][n] (or NULL) holds the rates into the absorbing neighbours that the
 * reference obtains by cutting rows/columns st+1..end-1 out of the periodic matrix: sink[0] = rate into bin st,
 * sink[n-1] = rate into bin end, 0 elsewhere (any non-negative vector is accepted).  Tensor-core scaling and squaring
 * on the symmetrised generator, same kernels as mcd_loglike_batch.  status[B] or NULL as in mcd_loglike_batch. */
int mcd_propagator_batch(mcd_handle* h, const mcd_problem* p, int32_t B, const double* v, const double* w,
                         const double* sink, const double* times, double* out_P, int32_t* status, void* stream);

/* ---- trajectory -> transition counts (the step before the likelihood path) ------------------
 *
 *   mcd_traj_counts      <->  mcdiff.tools.extract.transition_matrix_add1   (lib/tools/extract.py:71-90), applied to every
 *                             particle column and every shift as examples/create-transition-matrix/extract_Tmat.py:36-50
 *                             does, including its periodic fold  z - zpbc * floor(z / zpbc + 0.5)  (:45)
 *   mcd_traj_counts_rad  <->  mcdiff.tools.extract.count_2D                 (lib/tools/extract.py:127-156)
 *
 * x (and y, z): coordinates [nframes][ncols] (one column per particle), row-major fp64.  period > 0 folds the binned
 * coordinate into [-period/2, period/2) first, 0 bins it as it is.  edges[nedges] / redges[nredges]: increasing bin
 * edges; bins follow numpy.digitize (bin k = number of edges <= x, so 0 and nedges are the two outside bins, NaN
 * lands in bin nedges).  shifts_host[nshifts]: HOST array of frame shifts, 1 <= shift < nframes, nshifts <= 64.
 * counts is ACCUMULATED INTO (the reference's A += 1; zero it for a fresh matrix), int64:
 *   mcd_traj_counts:      [nshifts][nedges+1][nedges+1],           counts[l][end][start]
 *   mcd_traj_counts_rad:  [nshifts][nredges][nedges+1][nedges+1],  counts[l][r-1][end][start] with
 *                         r = digitize(sqrt(dx^2 + dy^2), redges) and r = 0 wrapping to the last radial bin (B[-1]).
 * Asynchronous on `stream`; the bin-index scratch (1-2 bytes per coordinate) is a stream-ordered allocation. */
int mcd_traj_counts(mcd_handle* h, const double* x, int64_t nframes, int64_t ncols, double period,
                    const double* edges, int32_t nedges, const int32_t* shifts_host, int32_t nshifts,
                    int64_t* counts, void* stream);
int mcd_traj_counts_rad(mcd_handle* h, const double* x, const double* y, const double* z, int64_t nframes,
                        int64_t ncols, double period, const double* edges, int32_t nedges, const double* redges,
                        int32_t nredges, const int32_t* shifts_host, int32_t nshifts, int64_t* counts, void* stream);

/* FP64 FMA micro-benchmark: fills *tflops with the measured dense DFMA rate of the device
 * (the roofline denominator bench.py reports against).  Synchronous. */
int mcd_fp64_peak(mcd_handle* h, double* tflops);

/* Number of kernels this library has launched through the handle (bench.py's gpu_launches). */
int64_t mcd_launch_count(const mcd_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* MCDIFF_B200_H */
