/*
 * pdeode_b200.h -- C ABI of libpdeode_b200.so: the implicit 2-D time step of
 * ericmjalbert/PDE-ODEsolver on NVIDIA B200 (sm_100a), fp64.
 *
 * The reference has no plugin / FFI interface: its Fortran subroutines are
 * called directly (only interface block: PDE-ODEsolver.f90:23-51).  The seam
 * this library replaces is the loop body of solveOrder2
 * (PDE-ODEsolver.f90:700-742) plus the in-loop diagnostics
 * (PDE-ODEsolver.f90:667-668,677).  The ABI is step-granular with
 * device-resident state: the caller uploads M and C once, calls
 * pdeode_step() per time step, and downloads fields only at frame times.
 * Each entry point cites the reference lines it stands in for.
 *
 * "P:" = PDE-ODEsolver.f90, "CG:" = CGdiag.f90 in the reference.
 *
 * Conventions
 *   - plain C types only; the caller owns host arrays, the library owns all
 *     device memory, streams and communicators; no pointer outlives a call.
 *   - fields are row-major, cell p = j + i*col (0-based; P:838 1-based),
 *     row*col doubles (the reference build is -fdefault-real-8).
 *   - every function returns 0 (PDEODE_OK) or an error code below; nothing
 *     aborts inside the library.  pdeode_step may OR PDEODE_WARN_CG_CAP into
 *     an otherwise-OK return (the reference ignores its CG cap, P:714).
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     point returns PDEODE_ERR_NODEVICE.
 *   - one context is not re-entrant; use one context per host thread / rank.
 */
#ifndef PDEODE_B200_H
#define PDEODE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PDEODE_OK              0
#define PDEODE_ERR_CUDA        1   /* a CUDA call failed; see pdeode_last_error */
#define PDEODE_ERR_NCCL        2
#define PDEODE_ERR_ARG         3
#define PDEODE_ERR_PICARD_CAP  4   /* countIters > 10000 (P:732-736): the host
                                      prints the two lines of P:733-734 and stops */
#define PDEODE_ERR_NOMEM       5
#define PDEODE_ERR_NODEVICE    6
#define PDEODE_ERR_MASK        0xff
#define PDEODE_WARN_CG_CAP     0x100 /* some solve hit nit > maxit (CG:88) */

#define PDEODE_NCCL_ID_BYTES   128

typedef struct pdeode_ctx pdeode_ctx;

/* The numerical parameters solveOrder2 receives (P:600-611).  xDel is what the
 * host computed at P:110 (1/row); e1 is carried for completeness (criterion 1
 * is commented out in the reference, CG:89). */
typedef struct {
    double delta, nu, kappa, gama; /* P:61 */
    double tDel, xDel;             /* P:69 */
    double e1, e2, eSoln;          /* P:69 */
    int alpha, beta;               /* P:62, integer exponents of D(M) */
    int dSelect, fSelect, gSelect; /* P:64 */
    int reserved;                  /* must be 0 */
} pdeode_params;

/* ---- lifetime ------------------------------------------------------- */

/* Allocates device state for a row x col grid on CUDA device `device`
 * (P:133, P:623-628).  The CG warm start Mnew is zero.
 * Requires row >= 3, col >= 2. */
int pdeode_create(pdeode_ctx **ctx, int row, int col, const pdeode_params *p,
                  int device);

/* Row-slab member of an nranks-wide solve: rank r owns global rows
 * [r*row/nranks, (r+1)*row/nranks).  id = PDEODE_NCCL_ID_BYTES bytes produced
 * by pdeode_nccl_unique_id on rank 0 and distributed by the caller (the
 * reference has no multi-process mode; this extends P:700-742 across GPUs). */
int pdeode_create_dist(pdeode_ctx **ctx, int row, int col,
                       const pdeode_params *p, int device, int rank,
                       int nranks, const void *id);
int pdeode_nccl_unique_id(void *id_out);

/* The same row-slab solve on ngpus devices of ONE process: the reference's drivers start a
 * single process (`echo FILE | ./a.out`, runAll.sh:7, twoDimension_runall.sh:7), so this is the
 * form the drop-in solver uses for grids larger than one GPU (PDEODE_NGPUS).  The context it
 * returns is used with every other entry point exactly like a single-GPU one -- fields are the
 * whole grid, counts are the step's -- while inside one worker thread per GPU drives a slab
 * context; the slabs see each other's memory through peer access (NVLink) instead of CUDA IPC.
 * devices: ngpus distinct CUDA ordinals, or NULL for 0 .. ngpus-1.  ngpus = 1 is pdeode_create. */
int pdeode_create_group(pdeode_ctx **ctx, int row, int col, const pdeode_params *p,
                        int ngpus, const int *devices);

int pdeode_destroy(pdeode_ctx *ctx);

/* A new parameter set on an existing context of the same grid: what pdeode_destroy +
 * pdeode_create with *p would give, without returning the device memory (the arrays, the
 * launch configuration and, for several GPUs, the communicator and peer mappings stay).  The
 * state is undefined afterwards: follow with pdeode_set_state / pdeode_set_initial_condition.
 * CG and Picard caps return to their defaults.  For sweeps over numbered parameter files
 * (.gitignore:34-35): pdeode_ensemble keeps one context per worker thread this way. */
int pdeode_reconfigure(pdeode_ctx *ctx, const pdeode_params *p);

/* Number of CUDA devices this process can use (0 and PDEODE_ERR_NODEVICE if none). */
int pdeode_device_count(int *count);

/* The row-slab partition itself (no device needed): rank's rows of a `row`-row
 * grid split over nranks. */
int pdeode_slab_rows(int row, int nranks, int rank, int *row0, int *nrows);

/* Global rows owned by this context: [*row0, *row0 + *nrows). */
int pdeode_local_rows(const pdeode_ctx *ctx, int *row0, int *nrows);

/* Launch on this cudaStream_t (as void*) from now on; NULL = the library's
 * own stream. */
int pdeode_set_stream(pdeode_ctx *ctx, void *cuda_stream);

/* ---- state ---------------------------------------------------------- */

/* H2D of this context's slab (nrows*col doubles each; the whole grid for a
 * single-GPU context).  Resets the CG warm start to zero, as a fresh call of
 * solveOrder2 would (P:133-138, P:623). */
int pdeode_set_state(pdeode_ctx *ctx, const double *M, const double *C);

/* setInitialConditions on the device (P:262-439), instead of
 * generating the fields on the host and uploading them.  MinitialCond 1-10.
 * px, py: the npts inoculation points of cases 4, 5, 9, 10 (the caller's random
 * draws xr, yr of P:312-313, 332-334, 409-414, or the spacing of P:393 with
 * py = 0); ignored for 1-3, 6, 8.  Case 7 (P:365-372: one random_number per
 * cell, in cell order) takes npts = 1 and the 64-bit seed of the caller's
 * generator as px[0] = seed mod 2^32, py[0] = seed div 2^32: cell i gets draw
 * number i of splitmix64(seed) -- the host program's stand-in for
 * random_number, whose k-th draw depends on seed + k * constant only.  Same
 * values as the reference's loops, cell by cell; case 5's normalisation sums
 * in parallel.  Resets the CG warm start like pdeode_set_state. */
int pdeode_set_initial_condition(pdeode_ctx *ctx, int MinitialCond, double depth, double height,
                                 int npts, const double *px, const double *py);

/* D2H of this context's slab; needed at frame times only (P:685-687). */
int pdeode_get_state(pdeode_ctx *ctx, double *M, double *C);

/* Frame output without moving whole fields.  The reference writes
 * every filter-th row and column, filter = (row-1)/256 for row > 257, else 1
 * (P:829-832, P:893-896).  *nrows_out x *ncols_out = the points of this
 * context's slab that a frame contains. */
int pdeode_frame_size(const pdeode_ctx *ctx, int filter, int *nrows_out, int *ncols_out);
/* The M and C values printToFile writes (P:834-842), nrows_out*ncols_out each,
 * row-major, decimated on the device. */
int pdeode_get_frame(pdeode_ctx *ctx, int filter, double *M, double *C);
/* The per-row column means printToFile2D writes (P:898-925), nrows_out each,
 * summed left to right like the reference. */
int pdeode_get_frame_rowmeans(pdeode_ctx *ctx, int filter, double *meanM, double *meanC);

/* ---- the path ------------------------------------------------------- */

/* One implicit time step, Picard loop included (P:700-742).  Returns this
 * step's Picard count and CG totals so the host can keep avgIters, maxIters,
 * avgNit, maxNit exactly as P:722-723,739-740 do. */
int pdeode_step(pdeode_ctx *ctx, int *countIters, long long *nitSum,
                long long *nitMax);

/* n time steps back to back: the stretch of the loop P:663-742 between two
 * points where the host looks at the state (diagnostics P:665, frames P:683).
 * What solveOrder2 accumulates over those steps comes back already summed /
 * maximised: countSum for `avgIters = avgIters + countIters` (P:740), countMax
 * for maxIters (P:739), nitSum for avgNit (P:723), nitMax for maxNit (P:722).
 * On small grids (the whole-step engine) this is one launch and one host wait
 * for all n steps; elsewhere it is n calls of pdeode_step.  stepsDone < n only
 * with PDEODE_ERR_PICARD_CAP (the failing step is not counted) or an error. */
int pdeode_step_many(pdeode_ctx *ctx, int n, int *stepsDone, long long *countSum,
                     int *countMax, long long *nitSum, long long *nitMax);

/* calcMass(M), calcMass(C) (P:667-668, P:772-791): grid means, global over
 * all ranks. */
int pdeode_mass(pdeode_ctx *ctx, double *meanM, double *meanC);

/* calcPeakInterface (P:677, P:978-1004).  peak / intfac are left untouched
 * when the reference would not assign them. */
int pdeode_peak(pdeode_ctx *ctx, double *peak, double *height, double *intfac);

/* ---- knobs and introspection (no reference counterpart) -------------- */

/* CG iteration cap; default 100*n (P:707).  < 0 restores the default. */
int pdeode_set_maxit(pdeode_ctx *ctx, long long maxit);

/* Picard iterate cap of a time step; default 10000 (P:732: `if (countIters > 10000)` ends the
 * run).  A step gives up with PDEODE_ERR_PICARD_CAP -- counts filled in, state = the last
 * iterate -- once countIters > cap, so cap = 0 makes a step exactly one Picard iterate (used
 * by bench.py to time a fixed number of CG iterations of a stiff solve).  < 0 restores 10000. */
int pdeode_set_picard_cap(pdeode_ctx *ctx, int cap);

/* Per-Picard-iterate record of the last pdeode_step: CG iterations of each
 * solve and (diffC, diffM) pairs (P:725-726); at most cap entries.  Returns
 * the number of iterates in *count. */
int pdeode_last_step_trace(pdeode_ctx *ctx, long long *nit_per_solve,
                           double *diffs, int cap, int *count);

/* Record (rho, dot, alfa, err2, normx) of every CG iteration of subsequent
 * solves, up to cap iterations per solve (0 disables). */
int pdeode_cg_trace_enable(pdeode_ctx *ctx, int cap);
/* Fetch the trace of the most recent solve: 5 doubles per iteration. */
int pdeode_cg_trace_get(pdeode_ctx *ctx, double *out, int cap, int *count);

/* Device arrays, for parity tests: which = PDEODE_ARR_*; out = nrows*col. */
#define PDEODE_ARR_M     0
#define PDEODE_ARR_C     1
#define PDEODE_ARR_MNEW  2   /* CG solution / next warm start (P:623) */
#define PDEODE_ARR_D     3   /* D(Mprev) (P:544-546) */
#define PDEODE_ARR_AC    4   /* centre diagonal (P:554-585) */
#define PDEODE_ARR_R     5
#define PDEODE_ARR_P     6
#define PDEODE_ARR_Q     7
int pdeode_debug_get_array(pdeode_ctx *ctx, int which, double *out);

/* How this context exchanges ghost rows and dot products per CG iteration:
 * 0 = single GPU (nothing to exchange), 1 = NCCL send/recv + all-reduce,
 * 2 = inside the CG kernels over NVLink peer memory (CUDA-IPC mappings). */
int pdeode_exchange_mode(const pdeode_ctx *ctx, int *mode);

/* One line describing the engine this context chose: SpMV variant, launch
 * shapes, persistent solve, 88-byte iteration, exchange. */
int pdeode_engine_info(const pdeode_ctx *ctx, char *buf, int len);

/* Kernels launched by this context so far. */
int pdeode_launch_count(const pdeode_ctx *ctx, long long *launches);

/* Time one kernel of the path in isolation on the context's current arrays:
 * `reps` back-to-back launches bracketed by CUDA events on the context's
 * stream; *ms_per_launch = elapsed / reps.  The arrays are scratch afterwards:
 * call pdeode_set_state before stepping again. */
#define PDEODE_KERNEL_CG_SPMV    0  /* p = r + beta p; q = A p; p.q, p.p */
#define PDEODE_KERNEL_CG_UPDATE  1  /* x += alfa p; r -= alfa q; r.r, x.x */
#define PDEODE_KERNEL_INIT       2  /* centre diagonal + r = b - A x0 */
#define PDEODE_KERNEL_SOLVEC     3  /* C update + Picard residuals */
#define PDEODE_KERNEL_DIFFCOEF   4  /* D(M) */
#define PDEODE_KERNEL_CG_SPMV_X  5  /* CG_SPMV + the previous iteration's x += alfa p; x.x (88-byte iteration) */
#define PDEODE_KERNEL_CG_UPDATE_R 6 /* r -= alfa q; r.r (88-byte iteration) */
int pdeode_time_kernel(pdeode_ctx *ctx, int kernel, int reps,
                       float *ms_per_launch);

/* Algorithmic bytes per grid cell moved by one launch of `kernel`. */
int pdeode_kernel_bytes_per_cell(int kernel);

const char *pdeode_strerror(int code);
/* Text of the last CUDA / NCCL failure on this context ("" if none). */
const char *pdeode_last_error(const pdeode_ctx *ctx);
const char *pdeode_version(void);

#ifdef __cplusplus
}
#endif
#endif
