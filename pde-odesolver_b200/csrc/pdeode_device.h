// pdeode_device.h -- types shared by the kernels and the engine.
//
// Device layout (all arrays identical in shape):
//   a slab of `rows` grid rows is stored with row pitch P (col rounded up to
//   16 doubles = 128 B), one ghost row above and one below, and a 16-double
//   guard at both ends, so that every 5-point neighbour of every cell --
//   including (row -1, col -1) -- is a legal, finite address.  Pad columns
//   [col, P) and unused ghost rows are kept at zero.  Array pointers handed
//   to kernels point at (row 0, col 0).
#pragma once
#include <cuda_runtime.h>

namespace pdeode {

struct GridDesc {
    int rows;        // local rows of this slab
    int col;         // real columns
    int P;           // row pitch in doubles
    int row0;        // global index of local row 0
    int grow;        // global row count
    double n_glob;   // (double)(grow*col): divisor of CG:57,87 / P:789,961
};

// Numerical parameters in the form the kernels use them.
struct NumParams {
    double delta, nu, kappa, gama, tDel;
    double h;        // (xCof*0.5), xCof = 1/(xDel*xDel)   (P:537,553)
    double inv_dt;   // (1/tDel)                            (P:585)
    double aux;      // tDel*0.5*gama                       (P:499)
    double tol2;     // e2                                  (CG:40)
    int alpha, beta, dSelect, fSelect, gSelect;
};

// Scalars of one CG solve / Picard iterate; lives in device memory.
struct DevScal {
    // CG state (CG:33)
    double rho, rho_old, alfa;
    double pq, pp;       // dotProd(p,q) (CG:76), sum(p*p) (CG:87)
    double xx;           // sum(sol*sol) before this iteration's update (CG:57)
    double err2, normx;
    long long nit;       // CG:55
    long long maxit;     // CG:40
    int stop;            // stopcrit (CG:53,88,90)
    int done;            // stop != 0: the update of this iteration is the last
    int finished;        // that last update has been applied
    int pad0;
    // Picard residuals (P:725-726) and field sums of the new iterate
    double diffC, diffM;
    // diagnostics (P:772-791, P:978-1004)
    double sumM, sumC;
    double peak_val;
    long long peak_idx;  // global cell index of the last maximal cell, -1 none
    int intfac_row;      // last global row with M > 0.1, -1 none
    int pad1;
    // raw (rank-local) sums awaiting a cross-rank reduction
    double part[4];
    // per-solve trace
    int trace_cap, trace_len;
    // tickets for the last-block-done reductions
    unsigned int ticket[8];
    int error;           // 1: a peer exchange timed out (multi-rank peer mode)
    // one rank, kernel-per-phase engine: the update kernel of the LAST iteration does nothing
    // (its r is dead; CG:81) and leaves sol += alfa p (CG:80) to the C-update kernel that reads
    // sol anyway: tail = 1 until k_solvec has applied it
    int tail;
    long long solvec_seq; // StencilArgs::seq of the solve whose C update has run
};

// How a kernel's reduction leaves the grid (StencilArgs / UpdateArgs ::fuse_scalars).
enum {
    SUMS_RAW = 0,    // rank-local sums into DevScal::part; NCCL all-reduce + a scalar kernel follow
    SUMS_FUSED = 1,  // single rank: the last CTA turns the sums into alfa / beta / stop itself
    SUMS_PEER = 2    // multi rank: the last CTA stores the sums into every rank's comm block over
                     // NVLink; the next kernel's CTAs wait until all ranks' words carry the phase
                     // tag and add them
};

// PDEODE_SUM_ORDER=serial (verification only): every reduction the reference does with a
// serial loop -- dotProd (CG:99-114), sum(sol*sol) / sum(p*p) (CG:57,87), calcDiff (P:956-960),
// calcMass (P:783-787) -- is re-done in the reference's order, one term after the other in
// cell order by one thread, so that a run is bit-identical to the serial CPU build.
enum { SER_DOT = 0, SER_ABSDIFF = 1, SER_SUM = 2 };
constexpr int SER_MAXJ = 4;      // reductions per pass
constexpr int SER_CH = 256;      // cells staged per chunk
struct SerialJob {
    const double *a, *b;         // SER_DOT: sum a*b ; SER_ABSDIFF: sum |a - b| ; SER_SUM: sum a
    int kind, pad;
};

constexpr int PEER_MAX_RANKS = 8;
constexpr int PEER_SLOTS = 4;    // phases in flight never span more than two slots
constexpr int PEER_VALS = 6;     // 8-byte words per (slot, rank): three sums as LL words (peer_post)

// Peer-memory exchange of one context (multi rank).  Every rank owns a comm
// block slots[PEER_SLOTS][PEER_MAX_RANKS][PEER_VALS] of 8-byte LL words (32 bits
// of payload + the low 32 bits of the phase number: peer_post / peer_wait_sums in
// pdeode_kernels.cu) that the other ranks write into through CUDA-IPC mappings;
// the flags area behind it is allocated but no longer used.  r_up / r_dn are
// the neighbours' residual arrays, whose ghost rows this rank fills directly.
struct PeerComm {
    int rank, nranks;
    const double *slots;                       // local comm block, written by peers
    const unsigned long long *flags;
    double *peer_slots[PEER_MAX_RANKS];        // every rank's comm block (self included)
    unsigned long long *peer_flags[PEER_MAX_RANKS];
    double *r_up, *r_dn;                       // origin (row 0, col 0) of the neighbours' r, or null
    int rows_up;                               // local rows of rank-1 (its bottom ghost row index)
    long long timeout_clocks;                  // a wait for a peer gives up after this many GPU clocks
};

enum { TICKET_STENCIL = 0, TICKET_UPDATE = 1, TICKET_SOLVEC = 2, TICKET_MASS = 3 };

// Arguments of the 5-point marching kernel.
struct StencilArgs {
    GridDesc g;
    NumParams np;
    DevScal *S;
    double *partials;      // [2][maxblocks]
    double *trace;         // may be null; 5 doubles per CG iteration
    // INIT: x0 -> (ac, r, x);   CG: (r, p) -> (p, q)
    const double *x0, *d, *Cc, *Mprev;
    const double *p_in;    // direction read by CG mode (never the array written: neighbours
                           // recompute r + beta p from it, so p cannot be updated in place)
    double *ac, *r, *x, *p, *q;  // p = direction written
    int TR;                // rows marched per CTA
    int BS;                // consumer threads per CTA (64 | 128 | 256); strip = 2*BS columns
    int DEPTH;             // ring depth in rows of the bulk-async variant
    int variant;           // CG SpMV: 0 register-staged, 1 bulk-async staged (default)
    int defer_x;           // 1: the SpMV kernel also applies the previous iteration's x += alfa p
                           // (x = StencilArgs::x) and sums x.x; the update kernel then only does r
    int fuse_scalars;      // SUMS_RAW | SUMS_FUSED | SUMS_PEER
    int write_ghost_top, write_ghost_bot; // maintain p ghost rows (multi rank)
    long long seq;
    PeerComm pc;           // SUMS_PEER only
    unsigned long long wait_phase, post_phase;
};

// Arguments of the persistent CG solve (k_cg_persist).
struct PersistArgs {
    StencilArgs st;        // grid, scalars, d, ac, r, q, peer block, wait_phase = INIT's phase
    double *x;             // CG iterate (Mnew)
    double *pbuf[2];       // direction, ping-pong; pbuf[ip] is read by the first iteration
    int ip;
    double *slots;         // [2][G][4] partial sums of the grid-wide sums (persist_allsum)
    unsigned long long *counter;   // arrivals at grid-wide sums since the context was created
    unsigned long long gen0;       // number of grid-wide sums of earlier launches
    int nstrips, nchunks;  // G = nstrips * nchunks CTAs, one tile each
    int serial;            // PDEODE_SUM_ORDER=serial: CTA 0 re-does every sum in cell order
    double *ser_out;       // [2][SER_MAXJ] results of those sums
    int prefetch;          // rows of a tile pulled into L2 while waiting for a grid-wide sum (0: none)
};

// What one whole-step launch (k_step_persist) reports back: the quantities
// solveOrder2 keeps per time step (P:702,722-723,730) plus the buffer roles.
constexpr int STEP_TRACE = 64;
struct StepOut {
    // over all time steps of the launch, as solveOrder2 accumulates them (P:722-723, P:739-740)
    long long count_sum, nit_sum_all, nit_max_all;
    int steps_done, count_max;
    // the last time step of the launch
    long long nit_sum, nit_max;          // P:723, P:722
    long long nit[STEP_TRACE];           // CG iterations of each solve of the step
    double diffs[2 * STEP_TRACE];        // (diffC, diffM) after each Picard iterate (P:725-726)
    int count;                           // countIters (P:730)
    int cg_cap;                          // some solve ran into nit > maxit (CG:88)
    int picard_cap;                      // count > cap (P:732)
    int error;                           // a grid barrier timed out
    int iM, iC, iMnew, ip;               // buffer roles after the step
    unsigned long long gen;              // generation of the last grid-wide sum (see step_allsum)
    unsigned long long seq;              // StepArgs::seq, stored last: the report is complete
    double sumM, sumC;                   // field sums of the final state (calcMass, P:783-787)
    int have_sums, pad1;
    int nstamps, pad;                    // PDEODE_STEP_STAMPS=1: phase time stamps of CTA 0
    unsigned long long stamp_ns[48], stamp_clk[48];
};

// Arguments of the whole time step in one cooperative launch (k_step_persist):
// D(Mprev), then per Picard iterate the centre diagonal, the CG solve, the C
// update and the residuals (P:700-742).  One rank only.
struct StepArgs {
    StencilArgs st;        // grid, numbers, scalars, d, ac, r, q
    double *Mb[3], *Cb[3]; // the rotating field buffers
    int iM, iC;            // current iterate at entry; becomes Mprev / Cprev (P:703-704)
    int nsteps;            // time steps in this launch
    int warm;              // 0: first solve ever, x0 = the zeroed buffer iMnew0 (D8)
    int iMnew0;
    double *pbuf[2];
    int ip;
    double *slots;         // [2][G][4] partial sums (step_allsum)
    unsigned long long *counter; // arrivals at grid-wide sums since the context was created
    unsigned long long gen0;     // number of grid-wide sums of earlier launches
    unsigned int *bar;     // bar[2]: error word
    int nstrips, nchunks;  // G = nstrips * nchunks CTAs of blockDim.x threads, one tile each
    int cw;                // threads per row group (strip width / 2); blockDim.x / cw row groups
    double eSoln;          // P:706
    int picard_cap;        // P:732
    StepOut *out;          // mapped host memory
    int stamps;            // 1: CTA 0 records a time stamp after every phase
    unsigned long long seq;
    int serial;            // PDEODE_SUM_ORDER=serial: CTA 0 re-does every sum in cell order
    double *ser_out;       // [2][SER_MAXJ] results of those sums
    int prefetch;          // pull the next phase's rows into L2 while waiting for a grid-wide sum
};

struct UpdateArgs {
    GridDesc g;
    DevScal *S;
    double *partials;
    double *x, *r;
    const double *p, *q;
    long long n2;          // number of double2 elements to stream (rows*P/2)
    int fuse_scalars;
    int reverse;           // 1: sweep high -> low addresses (L2 reuse against the SpMV kernel)
    int r_only;            // 1: x is updated by the next SpMV kernel (88-byte iteration)
    int tail_ok;           // 1: the last iteration's update may be left to k_solvec (DevScal::tail)
    long long seq;
    NumParams np;          // SUMS_PEER: the stop test runs in this kernel's last CTA
    double *trace;
    PeerComm pc;
    unsigned long long wait_phase, post_phase;
};

struct SolveCArgs {
    GridDesc g;
    NumParams np;
    DevScal *S;
    double *partials;      // [4][blocks]
    const double *Mprev, *Cprev, *Ccur, *Mcur;
    double *Mnew;          // CG solution; written here only when DevScal::tail is pending
    double *Cnew;
    const double *pbuf[2]; // the CG directions; the last one is pbuf[ip0 ^ (nit & 1)]
    int ip0;
    int gated;             // 1: return unless the solve has finished and this C update has not run
    long long n2;
    int fuse_scalars;
    long long seq;
};

struct MassArgs {
    GridDesc g;
    DevScal *S;
    double *partials;      // [4][maxblocks] + index arrays behind it
    const double *M, *C;
    long long n2;
};

// launchers (pdeode_kernels.cu)
cudaError_t launch_diffcoef(const GridDesc &g, const NumParams &np, const double *M,
                            double *d, cudaStream_t st);
cudaError_t launch_stencil_init(const StencilArgs &a, cudaStream_t st);
cudaError_t launch_stencil_cg(const StencilArgs &a, cudaStream_t st);
cudaError_t launch_update(const UpdateArgs &a, cudaStream_t st);
cudaError_t launch_update_r(const UpdateArgs &a, cudaStream_t st);
cudaError_t launch_xtail(double *x, const double *p, const DevScal *S, long long n2,
                         cudaStream_t st);
cudaError_t launch_solvec(const SolveCArgs &a, cudaStream_t st);
cudaError_t launch_mass(const MassArgs &a, cudaStream_t st);
// nj sums in the reference's serial order over the cells of g; out[0 .. nj)
cudaError_t launch_serial_reduce(const GridDesc &g, const SerialJob *jobs, int nj, double *out,
                                 cudaStream_t st);
cudaError_t launch_init_cond(const GridDesc &g, int ic, double depth, double height, double xDel,
                             int npts, const double *px, const double *py, double *M, double *C,
                             cudaStream_t st);
cudaError_t launch_ic_count_sum(const GridDesc &g, const double *M, DevScal *S, double *partials,
                                cudaStream_t st);
cudaError_t launch_ic_scale(const GridDesc &g, double *M, const DevScal *S, double height,
                            cudaStream_t st);
cudaError_t launch_decimate(const GridDesc &g, const double *M, const double *C, int filter,
                            int first_row, int nro, int nco, double *out, cudaStream_t st);
cudaError_t launch_rowmeans(const GridDesc &g, const double *M, const double *C, int filter,
                            int first_row, int nro, double *out, cudaStream_t st);
// scalar steps for the multi-rank path (sums already all-reduced into S->part)
cudaError_t launch_scal_after_init(DevScal *S, cudaStream_t st);
cudaError_t launch_scal_after_spmv(DevScal *S, GridDesc g, double tol2, double *trace,
                                   cudaStream_t st);
cudaError_t launch_scal_after_update(DevScal *S, cudaStream_t st);
cudaError_t launch_scal_after_solvec(DevScal *S, GridDesc g, cudaStream_t st);
// Launch shape of the 5-point kernels for one context.
struct LaunchCfg {
    int BS, TR, DEPTH, variant;
    // persistent CG solve: 0 = off, else one cooperative launch per solve with
    // p_nstrips * p_nchunks CTAs of p_BS consumer threads
    int persist, p_BS, p_nstrips, p_nchunks;
    int kb_reverse;        // K_B sweeps against K_A's direction
    int defer_x;           // 88-byte iteration: x += alfa p deferred into the next SpMV kernel
    int persist_defer;     // the same inside the persistent solve (measured slower: default 0)
    int persist_prefetch;  // L2 prefetch across the grid-wide sums, rows per tile: -1 = auto (multi-GPU only)
    int TR_x;              // rows per CTA of the 88-byte SpMV kernel (its optimum is lower)
    int i_BS, i_TR;        // launch shape of the INIT kernel (two IEEE divides per cell: it
                           // wants more warps per SM than the SpMV does)
    // whole time step in one cooperative launch (small grids, one rank): 0 = off, else
    // s_nstrips * s_nchunks CTAs of s_BS threads
    int whole_step, s_BS, s_cw, s_nstrips, s_nchunks;
};
cudaError_t launch_cg_persist(const PersistArgs &a, int BS, cudaStream_t st);
cudaError_t launch_step_persist(const StepArgs &a, int BS, cudaStream_t st);
// Heuristic defaults for a grid, then PDEODE_BS / PDEODE_TR / PDEODE_DEPTH /
// PDEODE_VARIANT (reg|bulk) overrides from the environment.
LaunchCfg launch_config(const GridDesc &g);
int stencil_max_blocks(const GridDesc &g, int TR);
int stream_blocks();

}  // namespace pdeode
