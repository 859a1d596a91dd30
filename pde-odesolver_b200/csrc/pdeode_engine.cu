// pdeode_engine.cu -- context, launch logic and the C ABI of libpdeode_b200.so
// (include/pdeode_b200.h).  Host-driven engine: the Picard loop and the CG
// loop run as kernel sequences on one stream; the device decides convergence
// (stop test of CG:88-90, residuals of P:725-726) and the host reads one
// small scalar record per chunk of CG iterations / per Picard iterate.
//
// There is no CPU fallback: without a CUDA device every entry point that
// would compute returns PDEODE_ERR_NODEVICE.
#include "../../include/pdeode_b200.h"
#include "nccl_dyn.h"
#include "pdeode_device.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <sched.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace pdeode;

namespace {

constexpr int GUARD = 16;          // doubles in front of / behind every array
constexpr int MAX_PICARD_TRACE = 64;
constexpr int PICARD_CAP = 10000;  // P:732

struct Snapshot { DevScal s; };

}  // namespace

struct pdeode_group;
namespace { struct GroupShared; }

struct pdeode_ctx {
    // an in-process group of row-slab contexts (pdeode_create_group): this object is then only
    // the handle; every entry point fans out to the members, one worker thread per GPU
    pdeode_group *group = nullptr;
    GroupShared *gs = nullptr;    // member of an in-process group: where the peers' pointers are
    int device = 0;
    int rank = 0, nranks = 1;
    GridDesc g{};
    NumParams np{};
    pdeode_params params{};
    long long n_local = 0;        // rows * col
    size_t arr_doubles = 0;       // allocation size of one array
    long long maxit = 0, maxit_default = 0;
    int picard_cap = PICARD_CAP;               // P:732; pdeode_set_picard_cap

    // field buffers: three M, three C (prev / current / new rotate), work arrays
    static constexpr int NARR = 12;
    double *base[NARR] = {nullptr};
    double *pbuf[2] = {nullptr, nullptr};  // CG direction, ping-pong (see StencilArgs::p_in)
    int ip = 0;                            // pbuf[ip] holds the current direction
    double *Mb[3] = {nullptr}, *Cb[3] = {nullptr};
    int iM = 0, iMprev = 0, iC = 0, iCprev = 0;
    int iMnew = 1;                // buffer holding the last CG solution
    double *d = nullptr, *ac = nullptr, *r = nullptr, *p = nullptr, *q = nullptr;
    bool warm = false;            // false until the first solve: x0 = 0 (D8)

    DevScal *S = nullptr;
    unsigned int *bar = nullptr;      // bar[2]: error word of the whole-step launch
    StepOut *step_out_h = nullptr;    // whole-step launch: its report, mapped host memory
    StepOut *step_out = nullptr;      // the same, device-side address
    bool step_poll = true;            // wait for the report by watching its sequence word
    bool step_yield = false;          // PDEODE_STEP_YIELD=1: give the core away between looks (ensembles)
    bool step_stamps = false;         // PDEODE_STEP_STAMPS=1: print CTA 0's phase times of every step to stderr
    unsigned long long step_seq = 0;
    double *persist_slots = nullptr;  // persistent solve: partial sums and arrival counter of its grid-wide sums
    unsigned long long *persist_counter = nullptr;
    unsigned long long persist_gen = 0;
    double *step_slots = nullptr;     // whole-step launch: partial sums and arrival counter of its grid-wide sums
    unsigned long long *step_counter = nullptr;
    unsigned long long step_gen = 0;
    double *partials = nullptr;
    int max_blocks = 0;
    double *trace = nullptr;
    int trace_cap = 0;
    double *gather = nullptr;     // multi-rank diagnostics
    double *frame_buf = nullptr;  // decimated frame staging
    size_t frame_cap = 0;
    Snapshot *snap = nullptr;     // pinned, 2 slots
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaStream_t own_stream = nullptr, stream = nullptr;
    ncclComm_t comm = nullptr;

    // peer-memory exchange (multi rank): see PeerComm in pdeode_device.h
    bool peer = false;
    PeerComm pc{};
    void *comm_block = nullptr;                 // this rank's slots + flags
    void *ipc_open[2 * PEER_MAX_RANKS + 2] = {nullptr};
    int n_ipc_open = 0;
    unsigned long long phase = 0;               // id of the last posting kernel launched
    bool solve_defer = false;                   // the solve in flight uses the 88-byte iteration
    bool pending_persist = false;               // a persistent solve is queued, not yet read back
    int pending_ip0 = 0;

    // PDEODE_SUM_ORDER=serial (verification, one rank): every reduction re-done in the
    // reference's cell order by one thread, so that a run equals the serial CPU build bit for bit
    bool serial = false;
    double *ser_out = nullptr;                  // [2][SER_MAXJ], persistent kernels

    LaunchCfg cfg{256, 64, 4, 1};
    int pred[MAX_PICARD_TRACE];
    long long seq = 0;
    long long launches = 0;
    bool mass_valid = false;
    // calcMass of the current state came with the step's last C update (k_solvec / the
    // whole-step launch): pdeode_mass needs no pass of its own
    bool sums_valid = false;
    double sumM = 0.0, sumC = 0.0;
    bool fuse_tail = true;        // PDEODE_FUSE_TAIL=0: the last update kernel of a solve runs in full

    // last step trace
    long long nit_trace[MAX_PICARD_TRACE];
    double diff_trace[2 * MAX_PICARD_TRACE];
    int trace_count = 0;

    std::string err;
};

namespace {

const char *const kVersion = "pdeode_b200 0.1 (sm_100a, fp64)";

int fail_cuda(pdeode_ctx *c, cudaError_t e, const char *what) {
    if (c) {
        char buf[256];
        snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
        c->err = buf;
    }
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? PDEODE_ERR_NODEVICE
                                                                      : PDEODE_ERR_CUDA;
}

int fail_nccl(pdeode_ctx *c, int e, const char *what) {
    char buf[256];
    const char *s = (nccl().GetErrorString) ? nccl().GetErrorString(e) : "?";
    snprintf(buf, sizeof buf, "%s: nccl error %d (%s)", what, e, s);
    c->err = buf;
    return PDEODE_ERR_NCCL;
}

#define CK(call)                                                      \
    do {                                                              \
        cudaError_t e_ = (call);                                      \
        if (e_ != cudaSuccess) return fail_cuda(c, e_, #call);        \
    } while (0)
#define NK(call)                                                      \
    do {                                                              \
        int e_ = (call);                                              \
        if (e_ != NCCL_SUCCESS) return fail_nccl(c, e_, #call);       \
    } while (0)

inline bool multi(const pdeode_ctx *c) { return c->nranks > 1; }
inline int sums_mode(const pdeode_ctx *c);
// 88-byte iteration: bulk-async SpMV, kernel-per-phase engine, one rank or peer exchange
inline bool defer_x(const pdeode_ctx *c) {
    return c->cfg.defer_x && (c->nranks == 1 || c->peer) && c->cfg.variant == 1 && !c->cfg.persist;
}
inline int sums_mode(const pdeode_ctx *c) {
    if (c->serial) return SUMS_RAW;      // the kernels' own sums are discarded; see serial_sums
    return !multi(c) ? SUMS_FUSED : (c->peer ? SUMS_PEER : SUMS_RAW);
}

// PDEODE_SUM_ORDER=serial: up to SER_MAXJ sums in the reference's cell order, into out[]
// (normally DevScal::part, where the scalar kernels of the SUMS_RAW path pick them up).
int serial_sums(pdeode_ctx *c, const SerialJob *jobs, int nj, double *out) {
    CK(launch_serial_reduce(c->g, jobs, nj, out, c->stream));
    c->launches++;
    return PDEODE_OK;
}

// ghost-row exchange with the slabs above and below (one row each way)
int halo_exchange(pdeode_ctx *c, double *X) {
    if (!multi(c)) return PDEODE_OK;
    const size_t P = (size_t)c->g.P;
    NK(nccl().GroupStart());
    if (c->rank > 0) {
        NK(nccl().Send(X, P, NCCL_DOUBLE, c->rank - 1, c->comm, c->stream));
        NK(nccl().Recv(X - P, P, NCCL_DOUBLE, c->rank - 1, c->comm, c->stream));
    }
    if (c->rank < c->nranks - 1) {
        NK(nccl().Send(X + (size_t)(c->g.rows - 1) * P, P, NCCL_DOUBLE, c->rank + 1, c->comm,
                       c->stream));
        NK(nccl().Recv(X + (size_t)c->g.rows * P, P, NCCL_DOUBLE, c->rank + 1, c->comm,
                       c->stream));
    }
    NK(nccl().GroupEnd());
    return PDEODE_OK;
}

int allreduce_part(pdeode_ctx *c, int count) {
    NK(nccl().AllReduce(c->S->part, c->S->part, (size_t)count, NCCL_DOUBLE, NCCL_SUM, c->comm,
                        c->stream));
    return PDEODE_OK;
}

int snapshot(pdeode_ctx *c, int slot) {
    CK(cudaMemcpyAsync(&c->snap[slot].s, c->S, sizeof(DevScal), cudaMemcpyDeviceToHost,
                       c->stream));
    CK(cudaEventRecord(c->ev[slot], c->stream));
    return PDEODE_OK;
}

int wait_snapshot(pdeode_ctx *c, int slot) {
    CK(cudaEventSynchronize(c->ev[slot]));
    return PDEODE_OK;
}

StencilArgs stencil_args(pdeode_ctx *c) {
    StencilArgs a{};
    a.g = c->g; a.np = c->np; a.S = c->S; a.partials = c->partials;
    a.trace = c->trace_cap > 0 ? c->trace : nullptr;
    a.d = c->d; a.ac = c->ac; a.r = c->r; a.q = c->q;
    a.p_in = c->pbuf[c->ip]; a.p = c->pbuf[c->ip ^ 1];
    a.TR = c->cfg.TR; a.BS = c->cfg.BS; a.DEPTH = c->cfg.DEPTH; a.variant = c->cfg.variant;
    a.fuse_scalars = sums_mode(c);
    a.pc = c->pc;
    a.write_ghost_top = multi(c) && c->rank > 0;
    a.write_ghost_bot = multi(c) && c->rank < c->nranks - 1;
    a.seq = c->seq;
    return a;
}

// Rows per tile the persistent solve pulls into L2 while a grid-wide sum is under way.
int persist_prefetch_rows(const pdeode_ctx *c) {
    if (c->cfg.persist_prefetch >= 0) return c->cfg.persist_prefetch;
    if (!multi(c)) return 0;
    const long long G = (long long)c->cfg.p_nstrips * c->cfg.p_nchunks;
    const long long per_row = 3LL * 2 * c->cfg.p_BS * 8;          // three arrays, one strip wide
    const long long rows = (40LL << 20) / std::max<long long>(1, G * per_row);
    return (int)std::min<long long>(rows, 1 << 20);
}

// one CG iteration: CG:55-90
int cg_iteration(pdeode_ctx *c, double *x) {
    const int mode = sums_mode(c);
    const bool defer = c->solve_defer;
    StencilArgs a = stencil_args(c);
    a.wait_phase = c->phase;                 // posted by INIT or the previous K_B
    a.post_phase = ++c->phase;
    a.defer_x = defer ? 1 : 0;
    if (defer) a.TR = c->cfg.TR_x;
    a.x = x;
    CK(launch_stencil_cg(a, c->stream));
    c->launches++;
    c->ip ^= 1;                      // the direction just written is the current one
    c->p = c->pbuf[c->ip];
    if (mode == SUMS_RAW) {
        int rc;
        if (c->serial) {                     // dotProd(p,q) (CG:76), sum(p*p) (CG:87)
            const SerialJob jb[2] = {{c->p, c->q, SER_DOT, 0}, {c->p, c->p, SER_DOT, 0}};
            if ((rc = serial_sums(c, jb, 2, c->S->part))) return rc;
        }
        if (multi(c) && (rc = allreduce_part(c, 2))) return rc;
        CK(launch_scal_after_spmv(c->S, c->g, c->np.tol2, a.trace, c->stream));
        c->launches++;
    }
    UpdateArgs u{};
    u.g = c->g; u.S = c->S; u.partials = c->partials;
    u.x = x; u.r = c->r; u.p = c->p; u.q = c->q;
    u.n2 = (long long)c->g.rows * c->g.P / 2;
    u.fuse_scalars = mode;
    u.reverse = c->cfg.kb_reverse;
    u.seq = c->seq;
    u.np = c->np; u.trace = a.trace; u.pc = c->pc;
    u.wait_phase = c->phase;                 // posted by the K_A just launched
    u.post_phase = ++c->phase;
    u.r_only = defer ? 1 : 0;                        // 88-byte iteration: r only
    u.tail_ok = (mode == SUMS_FUSED && c->fuse_tail) ? 1 : 0;
    if (defer && mode == SUMS_FUSED) CK(launch_update_r(u, c->stream));
    else CK(launch_update(u, c->stream));
    c->launches++;
    if (mode == SUMS_RAW) {
        int rc;
        if (c->serial) {                     // dotProd(r,r) (CG:59), sum(sol*sol) (CG:57)
            const SerialJob jb[2] = {{c->r, c->r, SER_DOT, 0}, {x, x, SER_DOT, 0}};
            if ((rc = serial_sums(c, jb, 2, c->S->part))) return rc;
        }
        if (multi(c) && (rc = allreduce_part(c, 2))) return rc;
        CK(launch_scal_after_update(c->S, c->stream));
        c->launches++;
        rc = halo_exchange(c, c->r);
        if (rc) return rc;
    }
    return PDEODE_OK;
}

// GenMatrixM's centre diagonal + solveLSDIAG (P:712-714): x0 -> x.
// sc: the C update of this Picard iterate (k_solvec, gated on the device: it runs once, as soon
// as the solve has finished) is queued behind every chunk of CG iterations, so that ONE record
// tells the host the iteration count and the Picard residuals; *solvec_done says whether it ran.
int solve(pdeode_ctx *c, const double *x0, double *x, int solve_idx, long long *nit_out,
          int *stop_out, SolveCArgs *sc, bool *solvec_done) {
    c->seq++;
    *solvec_done = false;
    if (multi(c)) {
        int rc = halo_exchange(c, const_cast<double *>(x0));
        if (rc) return rc;
    }
    StencilArgs a = stencil_args(c);
    a.x0 = x0; a.x = x; a.Cc = c->Cb[c->iC]; a.Mprev = c->Mb[c->iMprev];
    a.BS = c->cfg.i_BS; a.TR = c->cfg.i_TR;
    a.post_phase = ++c->phase;
    CK(launch_stencil_init(a, c->stream));
    c->launches++;
    if (sums_mode(c) == SUMS_RAW) {
        int rc;
        if (c->serial) {                     // dotProd(r,r), sum(sol*sol) of the start vector
            const SerialJob jb[2] = {{c->r, c->r, SER_DOT, 0}, {x, x, SER_DOT, 0}};
            if ((rc = serial_sums(c, jb, 2, c->S->part))) return rc;
        }
        if (multi(c) && (rc = allreduce_part(c, 2))) return rc;
        CK(launch_scal_after_init(c->S, c->stream));
        c->launches++;
        rc = halo_exchange(c, c->r);
        if (rc) return rc;
    }

    const int ip0 = c->ip;
    if (c->cfg.persist && (sums_mode(c) != SUMS_RAW || c->serial)) {
        // the whole CG loop in one cooperative launch (k_cg_persist)
        PersistArgs pa{};
        pa.st = stencil_args(c);
        pa.st.wait_phase = c->phase;             // posted by the INIT just launched
        // the 88-byte iteration also exists inside the persistent solve, but measured slower
        // there (its phase B re-reads p and q it has just written, L2-hot): 9.1e10 against
        // 1.10e11 on 2 GPUs at 4096^2, 3.9e9 against 4.5e9 at 256^2.  Off unless asked for.
        pa.st.defer_x = c->cfg.persist_defer ? 1 : 0;
        pa.x = x;
        pa.pbuf[0] = c->pbuf[0]; pa.pbuf[1] = c->pbuf[1]; pa.ip = c->ip;
        pa.slots = c->persist_slots; pa.counter = c->persist_counter; pa.gen0 = c->persist_gen;
        pa.nstrips = c->cfg.p_nstrips; pa.nchunks = c->cfg.p_nchunks;
        pa.serial = c->serial ? 1 : 0; pa.ser_out = c->ser_out;
        pa.prefetch = persist_prefetch_rows(c);
        CK(launch_cg_persist(pa, c->cfg.p_BS, c->stream));
        c->launches++;
        // no wait here: the C update queues right behind, and the caller reads the
        // iteration count out of the same record as the Picard residuals
        c->pending_persist = true;
        c->pending_ip0 = ip0;
        *nit_out = 0;
        *stop_out = 0;
        return PDEODE_OK;
    }

    // The 88-byte iteration saves 8 B per cell and iteration and costs one 24 B
    // pass at the end: worth it from about four iterations on (the previous
    // step's count for this Picard iterate decides; the same on every rank).
    c->solve_defer = defer_x(c) && c->pred[std::min(solve_idx, MAX_PICARD_TRACE - 1)] >= 4;

    // CG iterations in chunks; the device flags convergence, later launches
    // of a finished solve return at once.
    const int pred = std::max(1, c->pred[std::min(solve_idx, MAX_PICARD_TRACE - 1)]);
    long long launched = 0;
    int slot = 0;
    int chunk = (int)std::min<long long>(pred, 2048);
    bool have_pending = false;
    int pending_slot = 0;
    const bool with_solvec = sc && sums_mode(c) == SUMS_FUSED && c->fuse_tail;
    if (with_solvec) {
        sc->seq = c->seq; sc->gated = 1; sc->ip0 = ip0;
        sc->pbuf[0] = c->pbuf[0]; sc->pbuf[1] = c->pbuf[1];
    }
    for (;;) {
        for (int k = 0; k < chunk; ++k) {
            int rc = cg_iteration(c, x);
            if (rc) return rc;
        }
        launched += chunk;
        if (with_solvec) {
            CK(launch_solvec(*sc, c->stream));       // P:717, P:725-726; no-op until the solve is over
            c->launches++;
        }
        int rc = snapshot(c, slot);
        if (rc) return rc;
        if (have_pending) {
            rc = wait_snapshot(c, pending_slot);
            if (rc) return rc;
            if (c->snap[pending_slot].s.finished) {
                // the chunk just queued is all no-ops; its snapshot is the final state
                rc = wait_snapshot(c, slot);
                if (rc) return rc;
                break;
            }
        }
        // run ahead by one more chunk only when a solve is long enough to hide it
        const long long seen = have_pending ? c->snap[pending_slot].s.nit : 0;
        if (pred > 8 || have_pending) {
            have_pending = true;
            pending_slot = slot;
            slot ^= 1;
            chunk = (int)std::min<long long>(std::max<long long>(2, std::max<long long>(seen, launched) / 8), 1024);
            continue;
        }
        rc = wait_snapshot(c, slot);
        if (rc) return rc;
        if (c->snap[slot].s.finished) break;
        chunk = (int)std::min<long long>(std::max<long long>(2, launched / 2), 1024);
    }
    if (slot != 0) c->snap[0] = c->snap[slot];       // callers read the final record in slot 0
    const DevScal &fin = c->snap[0].s;
    if (fin.error) {
        c->err = multi(c) ? "peer exchange timed out (a rank stopped posting)"
                          : "grid barrier of the persistent solve timed out";
        return PDEODE_ERR_CUDA;
    }
    // launches after convergence were no-ops: the direction of the last real
    // iteration sits in the buffer given by the parity of nit
    c->ip = ip0 ^ (int)(fin.nit & 1);
    c->p = c->pbuf[c->ip];
    *solvec_done = with_solvec;
    if (c->solve_defer && !with_solvec) {
        // the x update of the last iteration was deferred to a kernel that never ran
        CK(launch_xtail(x, c->p, c->S, (long long)c->g.rows * c->g.P / 2, c->stream));
        c->launches++;
    }
    *nit_out = fin.nit;
    *stop_out = fin.stop;
    c->pred[std::min(solve_idx, MAX_PICARD_TRACE - 1)] = (int)std::min<long long>(fin.nit, 1 << 20);
    return PDEODE_OK;
}

int upload(pdeode_ctx *c, double *dst, const double *src) {
    CK(cudaMemcpy2DAsync(dst, (size_t)c->g.P * sizeof(double), src,
                         (size_t)c->g.col * sizeof(double), (size_t)c->g.col * sizeof(double),
                         (size_t)c->g.rows, cudaMemcpyHostToDevice, c->stream));
    return PDEODE_OK;
}

int download(pdeode_ctx *c, double *dst, const double *src) {
    CK(cudaMemcpy2DAsync(dst, (size_t)c->g.col * sizeof(double), src,
                         (size_t)c->g.P * sizeof(double), (size_t)c->g.col * sizeof(double),
                         (size_t)c->g.rows, cudaMemcpyDeviceToHost, c->stream));
    return PDEODE_OK;
}

int zero_array(pdeode_ctx *c, int k) {
    CK(cudaMemsetAsync(c->base[k], 0, c->arr_doubles * sizeof(double), c->stream));
    return PDEODE_OK;
}

int run_mass(pdeode_ctx *c) {
    if (c->mass_valid) return PDEODE_OK;
    MassArgs m{};
    m.g = c->g; m.S = c->S; m.partials = c->partials;
    m.M = c->Mb[c->iM]; m.C = c->Cb[c->iC];
    m.n2 = (long long)c->g.rows * c->g.P / 2;
    CK(launch_mass(m, c->stream));
    c->launches++;
    int rc;
    if (c->serial) {                                 // calcMass in cell order (P:783-787)
        const SerialJob jb[2] = {{m.M, nullptr, SER_SUM, 0}, {m.C, nullptr, SER_SUM, 0}};
        if ((rc = serial_sums(c, jb, 2, &c->S->sumM))) return rc;
    }
    rc = snapshot(c, 0);
    if (rc) return rc;
    rc = wait_snapshot(c, 0);
    if (rc) return rc;
    if (multi(c)) {
        // gather (sumM, sumC, peak_val, peak_idx, intfac_row) of every rank
        double loc[8] = {c->snap[0].s.sumM, c->snap[0].s.sumC, c->snap[0].s.peak_val,
                         (double)c->snap[0].s.peak_idx, (double)c->snap[0].s.intfac_row, 0, 0, 0};
        CK(cudaMemcpyAsync(c->gather + 8 * c->rank, loc, sizeof loc, cudaMemcpyHostToDevice,
                           c->stream));
        NK(nccl().AllGather(c->gather + 8 * c->rank, c->gather, 8, NCCL_DOUBLE, c->comm,
                            c->stream));
        std::vector<double> all(8 * (size_t)c->nranks);
        CK(cudaMemcpyAsync(all.data(), c->gather, all.size() * sizeof(double),
                           cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        DevScal &s = c->snap[0].s;
        s.sumM = s.sumC = 0.0; s.peak_val = -1.0; s.peak_idx = -1; s.intfac_row = -1;
        for (int r = 0; r < c->nranks; ++r) {      // rank order = row order
            const double *v = &all[8 * (size_t)r];
            s.sumM += v[0]; s.sumC += v[1];
            if ((long long)v[3] >= 0 && v[2] >= s.peak_val) { s.peak_val = v[2]; s.peak_idx = (long long)v[3]; }
            if ((int)v[4] > s.intfac_row) s.intfac_row = (int)v[4];
        }
    }
    c->mass_valid = true;
    return PDEODE_OK;
}

void set_numparams(pdeode_ctx *c, const pdeode_params *p) {
    NumParams &n = c->np;
    n.delta = p->delta; n.nu = p->nu; n.kappa = p->kappa; n.gama = p->gama; n.tDel = p->tDel;
    const double xCof = 1.0 / (p->xDel * p->xDel);      // P:537
    n.h = xCof * 0.5;                                   // P:553 (left to right)
    n.inv_dt = 1.0 / p->tDel;                           // P:585
    n.aux = p->tDel * 0.5 * p->gama;                    // P:499
    n.tol2 = p->e2;                                     // CG:40
    n.alpha = p->alpha; n.beta = p->beta;
    n.dSelect = p->dSelect; n.fSelect = p->fSelect; n.gSelect = p->gSelect;
}

// Maps every rank's comm block and the two neighbours' residual arrays into
// this process (CUDA IPC over NVLink) so that K_A / K_B exchange their sums and
// ghost rows themselves.  If the mappings cannot be made on every rank the
// context stays on the NCCL path; the decision is taken collectively.
int setup_peer(pdeode_ctx *c) {
    struct Desc {
        cudaIpcMemHandle_t comm, r;
        int rows, ok;
    };
    const int nr = c->nranks;
    const size_t slots_bytes = sizeof(double) * PEER_SLOTS * PEER_MAX_RANKS * PEER_VALS;
    const size_t flags_bytes = sizeof(unsigned long long) * PEER_SLOTS * PEER_MAX_RANKS;
    CK(cudaMalloc(&c->comm_block, slots_bytes + flags_bytes));
    CK(cudaMemset(c->comm_block, 0, slots_bytes + flags_bytes));
    Desc mine{};
    mine.rows = c->g.rows;
    mine.ok = (cudaIpcGetMemHandle(&mine.comm, c->comm_block) == cudaSuccess &&
               cudaIpcGetMemHandle(&mine.r, c->base[8]) == cudaSuccess);
    cudaGetLastError();
    Desc *dall = nullptr;
    CK(cudaMalloc(&dall, sizeof(Desc) * (size_t)nr));
    CK(cudaMemcpy(dall + c->rank, &mine, sizeof(Desc), cudaMemcpyHostToDevice));
    NK(nccl().AllGather(dall + c->rank, dall, sizeof(Desc), NCCL_INT8, c->comm, c->stream));
    std::vector<Desc> all((size_t)nr);
    CK(cudaMemcpyAsync(all.data(), dall, sizeof(Desc) * (size_t)nr, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaFree(dall));

    bool ok = true;
    for (int r = 0; r < nr; ++r) ok = ok && all[(size_t)r].ok;
    PeerComm pc{};
    pc.rank = c->rank; pc.nranks = nr;
    {
        // a wait for a peer gives up after this long (GPU clocks at ~2 GHz); generous by default:
        // a descheduled peer process or a profiler replay must not turn into a failure
        double ms = 10000.0;
        if (const char *e = getenv("PDEODE_PEER_TIMEOUT_MS")) { const double v = atof(e); if (v > 0) ms = v; }
        pc.timeout_clocks = (long long)(ms * 2.0e6);
    }
    pc.slots = (const double *)c->comm_block;
    pc.flags = (const unsigned long long *)((char *)c->comm_block + slots_bytes);
    const size_t origin = (size_t)GUARD + (size_t)c->g.P;        // (row 0, col 0) inside an array
    for (int r = 0; r < nr && ok; ++r) {
        void *cb = c->comm_block;
        if (r != c->rank) {
            if (cudaIpcOpenMemHandle(&cb, all[(size_t)r].comm, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                ok = false; cudaGetLastError(); break;
            }
            c->ipc_open[c->n_ipc_open++] = cb;
        }
        pc.peer_slots[r] = (double *)cb;
        pc.peer_flags[r] = (unsigned long long *)((char *)cb + slots_bytes);
        if (r == c->rank - 1 || r == c->rank + 1) {
            void *rb = nullptr;
            if (cudaIpcOpenMemHandle(&rb, all[(size_t)r].r, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                ok = false; cudaGetLastError(); break;
            }
            c->ipc_open[c->n_ipc_open++] = rb;
            if (r == c->rank - 1) { pc.r_up = (double *)rb + origin; pc.rows_up = all[(size_t)r].rows; }
            else pc.r_dn = (double *)rb + origin;
        }
    }
    // every rank must take the same path: all-reduce the verdict (sum of failures)
    double *flag = c->gather;
    const double bad = ok ? 0.0 : 1.0;
    CK(cudaMemcpy(flag, &bad, sizeof(double), cudaMemcpyHostToDevice));
    NK(nccl().AllReduce(flag, flag, 1, NCCL_DOUBLE, NCCL_SUM, c->comm, c->stream));
    double nbad = 0.0;
    CK(cudaMemcpyAsync(&nbad, flag, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (nbad == 0.0) {
        c->pc = pc;
        c->peer = true;
    } else {
        c->err = "CUDA IPC peer mapping unavailable; using the NCCL exchange";
    }
    return PDEODE_OK;
}

// What the members of an in-process group (pdeode_create_group) tell each other at create
// time, and the barrier they meet at.
struct GroupShared {
    int n = 0;
    std::mutex mu;
    std::condition_variable cv;
    int arrived = 0;
    unsigned long long gen = 0;
    void *comm_block[PEER_MAX_RANKS] = {nullptr};
    double *r_base[PEER_MAX_RANKS] = {nullptr};
    int rows[PEER_MAX_RANKS] = {0};
    int device[PEER_MAX_RANKS] = {0};
    int ok[PEER_MAX_RANKS] = {0};
    bool dead = false;            // a member failed before reaching a barrier: nobody waits any more
    void barrier() {
        std::unique_lock<std::mutex> lk(mu);
        if (dead) return;
        const unsigned long long g0 = gen;
        if (++arrived == n) { arrived = 0; ++gen; cv.notify_all(); return; }
        cv.wait(lk, [&] { return gen != g0 || dead; });
    }
    void abandon() {
        std::lock_guard<std::mutex> lk(mu);
        dead = true;
        cv.notify_all();
    }
};

// setup_peer for ranks living in ONE process (one worker thread per GPU): no IPC handles --
// cudaIpcOpenMemHandle refuses handles of the own process -- but plain peer access to the
// other devices and the members' raw device pointers.
int setup_peer_inproc(pdeode_ctx *c) {
    GroupShared *gs = c->gs;
    const int nr = c->nranks;
    const size_t slots_bytes = sizeof(double) * PEER_SLOTS * PEER_MAX_RANKS * PEER_VALS;
    const size_t flags_bytes = sizeof(unsigned long long) * PEER_SLOTS * PEER_MAX_RANKS;
    cudaError_t e = cudaMalloc(&c->comm_block, slots_bytes + flags_bytes);
    if (e == cudaSuccess) e = cudaMemset(c->comm_block, 0, slots_bytes + flags_bytes);
    if (e != cudaSuccess) { gs->abandon(); return fail_cuda(c, e, "comm block"); }
    gs->comm_block[c->rank] = c->comm_block;
    gs->r_base[c->rank] = c->base[8];
    gs->rows[c->rank] = c->g.rows;
    gs->device[c->rank] = c->device;
    gs->barrier();
    if (gs->dead) { c->err = "another member of the group failed to initialise"; return PDEODE_ERR_CUDA; }
    bool ok = true;
    for (int r = 0; r < nr && ok; ++r) {
        if (r == c->rank) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, c->device, gs->device[r]) != cudaSuccess || !can) { ok = false; break; }
        e = cudaDeviceEnablePeerAccess(gs->device[r], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
        cudaGetLastError();
    }
    gs->ok[c->rank] = ok ? 1 : 0;
    gs->barrier();
    if (gs->dead) { c->err = "another member of the group failed to initialise"; return PDEODE_ERR_CUDA; }
    for (int r = 0; r < nr; ++r) ok = ok && gs->ok[r];
    if (ok) {
        PeerComm pc{};
        pc.rank = c->rank; pc.nranks = nr;
        double ms = 10000.0;
        if (const char *t = getenv("PDEODE_PEER_TIMEOUT_MS")) { const double v = atof(t); if (v > 0) ms = v; }
        pc.timeout_clocks = (long long)(ms * 2.0e6);
        pc.slots = (const double *)c->comm_block;
        pc.flags = (const unsigned long long *)((char *)c->comm_block + slots_bytes);
        const size_t origin = (size_t)GUARD + (size_t)c->g.P;
        for (int r = 0; r < nr; ++r) {
            pc.peer_slots[r] = (double *)gs->comm_block[r];
            pc.peer_flags[r] = (unsigned long long *)((char *)gs->comm_block[r] + slots_bytes);
        }
        if (c->rank > 0) { pc.r_up = gs->r_base[c->rank - 1] + origin; pc.rows_up = gs->rows[c->rank - 1]; }
        if (c->rank < nr - 1) pc.r_dn = gs->r_base[c->rank + 1] + origin;
        c->pc = pc;
        c->peer = true;
    } else {
        c->err = "peer access between the group's devices unavailable; using the NCCL exchange";
    }
    return PDEODE_OK;
}

int create_impl(pdeode_ctx **out, int row, int col, const pdeode_params *p, int device, int rank,
                int nranks, const void *id, GroupShared *gs = nullptr) {
    if (!out) return PDEODE_ERR_ARG;
    *out = nullptr;
    if (!p || row < 3 || col < 2 || nranks < 1 || rank < 0 || rank >= nranks ||
        row / nranks < 2 || (nranks > 1 && !id) || !(p->tDel > 0) || !(p->xDel > 0))
        return PDEODE_ERR_ARG;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) return PDEODE_ERR_NODEVICE;
    if (device < 0 || device >= ndev) return PDEODE_ERR_ARG;
    pdeode_ctx *c = new (std::nothrow) pdeode_ctx();
    if (!c) return PDEODE_ERR_NOMEM;
    *out = c;   // returned even on failure so that pdeode_last_error works; destroy it
    c->device = device; c->rank = rank; c->nranks = nranks; c->params = *p;
    c->gs = gs;
    CK(cudaSetDevice(device));

    int r0 = 0, nrows_local = 0;
    pdeode_slab_rows(row, nranks, rank, &r0, &nrows_local);
    c->g.rows = nrows_local; c->g.col = col; c->g.P = (col + 15) / 16 * 16;
    c->g.row0 = r0; c->g.grow = row;
    c->g.n_glob = (double)((long long)row * col);
    c->n_local = (long long)c->g.rows * col;
    c->maxit_default = 100LL * row * col;               // P:707 (64-bit: SURVEY D9)
    c->maxit = c->maxit_default;
    set_numparams(c, p);
    for (int k = 0; k < MAX_PICARD_TRACE; ++k) c->pred[k] = (k == 0) ? 8 : 1;

    c->arr_doubles = (size_t)(c->g.rows + 2) * c->g.P + 2 * GUARD;
    for (int k = 0; k < pdeode_ctx::NARR; ++k) {
        cudaError_t me = cudaMalloc(&c->base[k], c->arr_doubles * sizeof(double));
        if (me != cudaSuccess) {
            fail_cuda(c, me, "cudaMalloc(field)");
            return me == cudaErrorMemoryAllocation ? PDEODE_ERR_NOMEM : PDEODE_ERR_CUDA;
        }
        CK(cudaMemset(c->base[k], 0, c->arr_doubles * sizeof(double)));
    }
    auto origin = [&](int k) { return c->base[k] + GUARD + c->g.P; };
    for (int k = 0; k < 3; ++k) { c->Mb[k] = origin(k); c->Cb[k] = origin(3 + k); }
    c->d = origin(6); c->ac = origin(7); c->r = origin(8); c->q = origin(10);
    c->pbuf[0] = origin(9); c->pbuf[1] = origin(11); c->ip = 0; c->p = c->pbuf[0];

    c->cfg = launch_config(c->g);
    if (const char *ft = getenv("PDEODE_FUSE_TAIL")) c->fuse_tail = atoi(ft) != 0;
    if (const char *so = getenv("PDEODE_SUM_ORDER")) {
        if (strcmp(so, "serial") == 0) {
            if (nranks > 1) { c->err = "PDEODE_SUM_ORDER=serial is a one-rank verification mode"; return PDEODE_ERR_ARG; }
            c->serial = true;
            c->cfg.defer_x = 0; c->cfg.persist_defer = 0;     // x.x is taken where the reference takes it
            CK(cudaMalloc(&c->ser_out, sizeof(double) * 2 * SER_MAXJ));
            CK(cudaMemset(c->ser_out, 0, sizeof(double) * 2 * SER_MAXJ));
        } else if (strcmp(so, "tree") != 0 && so[0] != '\0') {
            c->err = "PDEODE_SUM_ORDER must be 'tree' or 'serial'"; return PDEODE_ERR_ARG;
        }
    }
    // the marching kernels never launch more CTAs than stencil_max_blocks(g, TR)
    c->max_blocks = std::max(stencil_max_blocks(c->g, std::min(std::min(c->cfg.TR, c->cfg.TR_x), c->cfg.i_TR)),
                             std::max(stream_blocks(), 512));
    CK(cudaMalloc(&c->partials, sizeof(double) * 8 * (size_t)c->max_blocks));
    {
        const size_t G = (size_t)std::max(1, std::max(c->cfg.p_nstrips * c->cfg.p_nchunks,
                                                      c->cfg.s_nstrips * c->cfg.s_nchunks));
        CK(cudaMalloc(&c->bar, 4 * sizeof(unsigned int)));
        CK(cudaMemset(c->bar, 0, 4 * sizeof(unsigned int)));
        CK(cudaMalloc(&c->persist_slots, sizeof(double) * 8 * G));
        CK(cudaMemset(c->persist_slots, 0, sizeof(double) * 8 * G));
        CK(cudaMalloc(&c->persist_counter, sizeof(unsigned long long)));
        CK(cudaMemset(c->persist_counter, 0, sizeof(unsigned long long)));
    }
    if (nranks > 1) c->cfg.whole_step = 0;              // one rank only
    if (c->cfg.whole_step) {
        // the launch reports straight into mapped host memory: no copy behind the kernel
        CK(cudaHostAlloc(&c->step_out_h, sizeof(StepOut), cudaHostAllocMapped));
        memset(c->step_out_h, 0, sizeof(StepOut));
        CK(cudaHostGetDevicePointer(&c->step_out, c->step_out_h, 0));
        const char *pe = getenv("PDEODE_STEP_POLL");
        c->step_poll = !(pe && pe[0] == '0');
        const char *sh = getenv("PDEODE_STEP_YIELD");
        c->step_yield = sh && sh[0] == '1';
        const char *se = getenv("PDEODE_STEP_STAMPS");
        c->step_stamps = se && se[0] == '1';
        const size_t G = (size_t)c->cfg.s_nstrips * c->cfg.s_nchunks;
        CK(cudaMalloc(&c->step_slots, sizeof(double) * 8 * G));
        CK(cudaMemset(c->step_slots, 0, sizeof(double) * 8 * G));
        CK(cudaMalloc(&c->step_counter, sizeof(unsigned long long)));
        CK(cudaMemset(c->step_counter, 0, sizeof(unsigned long long)));
    }
    CK(cudaMalloc(&c->S, sizeof(DevScal)));
    CK(cudaMemset(c->S, 0, sizeof(DevScal)));
    CK(cudaMallocHost(&c->snap, 2 * sizeof(Snapshot)));
    memset(c->snap, 0, 2 * sizeof(Snapshot));
    CK(cudaEventCreateWithFlags(&c->ev[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev[1], cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CK(cudaMemcpy(&c->S->maxit, &c->maxit, sizeof(long long), cudaMemcpyHostToDevice));

    if (nranks > 1) {
        if (!nccl().ok) { c->err = "libnccl.so.2 not found"; return PDEODE_ERR_NCCL; }
        ncclUniqueId uid;
        memcpy(&uid, id, sizeof uid);
        if (gs) {
            // in-process group: nobody enters the (blocking) communicator setup unless every
            // member got this far
            gs->barrier();
            if (gs->dead) { c->err = "another member of the group failed to initialise"; return PDEODE_ERR_CUDA; }
        }
        NK(nccl().CommInitRank(&c->comm, nranks, uid, rank));
        CK(cudaMalloc(&c->gather, sizeof(double) * 8 * (size_t)nranks));
        const char *pe = getenv("PDEODE_PEER");
        if (!(pe && pe[0] == '0') && nranks <= PEER_MAX_RANKS && c->cfg.variant == 1) {
            // ranks of one process see each other's memory through peer access; ranks in
            // separate processes through CUDA-IPC mappings
            int rc = gs ? setup_peer_inproc(c) : setup_peer(c);
            if (rc) return rc;
        }
    }
    return PDEODE_OK;
}

// ---------------------------------------------------------------- in-process multi-GPU group

// One worker thread per member: every entry point of a group context runs on all members at
// once (their kernels wait for each other across GPUs, so they must be launched together).
struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<void()> job;
    bool has = false, quit = false;
    Worker() {
        th = std::thread([this] {
            for (;;) {
                std::function<void()> j;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return has || quit; });
                    if (quit && !has) return;
                    j = std::move(job);
                }
                j();
                {
                    std::lock_guard<std::mutex> lk(mu);
                    has = false;
                }
                cv.notify_all();
            }
        });
    }
    void post(std::function<void()> j) {
        std::lock_guard<std::mutex> lk(mu);
        job = std::move(j);
        has = true;
        cv.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return !has; });
    }
    ~Worker() {
        {
            std::lock_guard<std::mutex> lk(mu);
            quit = true;
        }
        cv.notify_all();
        if (th.joinable()) th.join();
    }
};

}  // namespace

struct pdeode_group {
    int n = 0;
    std::vector<pdeode_ctx *> m;
    std::vector<std::unique_ptr<Worker>> w;
    GroupShared sh;
};

namespace {

// f(k, member) on every member at once; returns the first error, else member 0's code
// (warnings included).
template <class F>
int group_run(pdeode_ctx *front, F f) {
    pdeode_group *G = front->group;
    std::vector<int> rc((size_t)G->n, 0);
    for (int k = 0; k < G->n; ++k) G->w[(size_t)k]->post([&, k] { rc[(size_t)k] = f(k, G->m[(size_t)k]); });
    for (int k = 0; k < G->n; ++k) G->w[(size_t)k]->wait();
    for (int k = 0; k < G->n; ++k)
        if (rc[(size_t)k] & PDEODE_ERR_MASK) {
            front->err = "GPU " + std::to_string(k) + ": " + (G->m[(size_t)k] ? G->m[(size_t)k]->err : std::string("?"));
            return rc[(size_t)k];
        }
    return rc[0];
}

int group_create(pdeode_ctx **out, int row, int col, const pdeode_params *p, int ngpus,
                 const int *devices) {
    if (!out) return PDEODE_ERR_ARG;
    *out = nullptr;
    if (!p || ngpus < 1 || ngpus > PEER_MAX_RANKS || row / ngpus < 2) return PDEODE_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return PDEODE_ERR_NODEVICE; }
    std::vector<int> dev((size_t)ngpus);
    for (int k = 0; k < ngpus; ++k) {
        dev[(size_t)k] = devices ? devices[k] : k;
        if (dev[(size_t)k] < 0 || dev[(size_t)k] >= ndev) return PDEODE_ERR_ARG;
        for (int j = 0; j < k; ++j)
            if (dev[(size_t)j] == dev[(size_t)k]) return PDEODE_ERR_ARG;   // slabs sharing a GPU would wait for each other forever
    }
    if (ngpus == 1) return create_impl(out, row, col, p, dev[0], 0, 1, nullptr);
    if (!nccl().ok) return PDEODE_ERR_NCCL;
    ncclUniqueId uid;
    if (nccl().GetUniqueId(&uid) != NCCL_SUCCESS) return PDEODE_ERR_NCCL;
    pdeode_ctx *front = new (std::nothrow) pdeode_ctx();
    if (!front) return PDEODE_ERR_NOMEM;
    *out = front;
    front->group = new pdeode_group();
    pdeode_group *G = front->group;
    G->n = ngpus;
    G->sh.n = ngpus;
    G->m.assign((size_t)ngpus, nullptr);
    for (int k = 0; k < ngpus; ++k) G->w.emplace_back(new Worker());
    front->device = dev[0]; front->rank = 0; front->nranks = ngpus; front->params = *p;
    front->g.rows = row; front->g.col = col; front->g.P = (col + 15) / 16 * 16;
    front->g.row0 = 0; front->g.grow = row; front->g.n_glob = (double)((long long)row * col);
    front->n_local = (long long)row * col;
    const int rc = group_run(front, [&](int k, pdeode_ctx *) {
        const int r = create_impl(&G->m[(size_t)k], row, col, p, dev[(size_t)k], k, ngpus, &uid, &G->sh);
        if (r) G->sh.abandon();
        return r;
    });
    return rc;
}

int group_destroy(pdeode_ctx *front) {
    pdeode_group *G = front->group;
    // collective inside (no member frees memory a peer may still write into): all at once
    group_run(front, [&](int, pdeode_ctx *m) { return m ? pdeode_destroy(m) : 0; });
    G->w.clear();
    delete G;
    delete front;
    return PDEODE_OK;
}

// rows of the whole grid a frame contains, per member, in member order
void group_frame_rows(pdeode_ctx *front, int filter, std::vector<int> &nro, int *nco) {
    pdeode_group *G = front->group;
    nro.assign((size_t)G->n, 0);
    for (int k = 0; k < G->n; ++k) pdeode_frame_size(G->m[(size_t)k], filter, &nro[(size_t)k], nco);
}

#define GROUP(c, expr) do { if ((c) && (c)->group) return (expr); } while (0)

}  // namespace

// ======================================================================
extern "C" {

const char *pdeode_version(void) { return kVersion; }

const char *pdeode_strerror(int code) {
    switch (code & PDEODE_ERR_MASK) {
    case PDEODE_OK: return (code & PDEODE_WARN_CG_CAP) ? "ok (a CG solve hit its iteration cap)" : "ok";
    case PDEODE_ERR_CUDA: return "CUDA error";
    case PDEODE_ERR_NCCL: return "NCCL error";
    case PDEODE_ERR_ARG: return "invalid argument";
    case PDEODE_ERR_PICARD_CAP: return "over 10000 Picard iterations in one timestep; solutions not converging";
    case PDEODE_ERR_NOMEM: return "out of memory";
    case PDEODE_ERR_NODEVICE: return "no CUDA device (this library has no CPU fallback)";
    default: return "unknown error";
    }
}

const char *pdeode_last_error(const pdeode_ctx *c) { return c ? c->err.c_str() : ""; }

int pdeode_nccl_unique_id(void *id_out) {
    if (!id_out) return PDEODE_ERR_ARG;
    if (!nccl().ok) return PDEODE_ERR_NCCL;
    ncclUniqueId uid;
    if (nccl().GetUniqueId(&uid) != NCCL_SUCCESS) return PDEODE_ERR_NCCL;
    memcpy(id_out, &uid, sizeof uid);
    return PDEODE_OK;
}

int pdeode_create(pdeode_ctx **ctx, int row, int col, const pdeode_params *p, int device) {
    return create_impl(ctx, row, col, p, device, 0, 1, nullptr);
}

int pdeode_create_dist(pdeode_ctx **ctx, int row, int col, const pdeode_params *p, int device,
                       int rank, int nranks, const void *id) {
    return create_impl(ctx, row, col, p, device, rank, nranks, id);
}

int pdeode_create_group(pdeode_ctx **ctx, int row, int col, const pdeode_params *p, int ngpus,
                        const int *devices) {
    return group_create(ctx, row, col, p, ngpus, devices);
}

int pdeode_reconfigure(pdeode_ctx *c, const pdeode_params *p) {
    if (!c || !p) return PDEODE_ERR_ARG;
    GROUP(c, group_run(c, [&](int, pdeode_ctx *m) { return pdeode_reconfigure(m, p); }));
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->params = *p;
    set_numparams(c, p);
    for (int k = 0; k < MAX_PICARD_TRACE; ++k) c->pred[k] = (k == 0) ? 8 : 1;
    c->maxit = c->maxit_default;
    c->picard_cap = PICARD_CAP;
    CK(cudaMemcpy(&c->S->maxit, &c->maxit, sizeof(long long), cudaMemcpyHostToDevice));
    c->warm = false;
    c->trace_count = 0;
    c->mass_valid = false; c->sums_valid = false;
    c->launches = 0;
    return PDEODE_OK;
}

int pdeode_destroy(pdeode_ctx *c) {
    if (!c) return PDEODE_OK;
    GROUP(c, group_destroy(c));
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->peer && c->comm && c->gather) {
        // collective: no rank unmaps or frees memory a peer may still be writing into
        nccl().AllReduce(c->gather, c->gather, 1, NCCL_DOUBLE, NCCL_SUM, c->comm, c->stream);
        cudaStreamSynchronize(c->stream);
    }
    for (int k = 0; k < c->n_ipc_open; ++k) cudaIpcCloseMemHandle(c->ipc_open[k]);
    if (c->comm_block) cudaFree(c->comm_block);
    if (c->comm) nccl().CommDestroy(c->comm);
    for (int k = 0; k < pdeode_ctx::NARR; ++k) if (c->base[k]) cudaFree(c->base[k]);
    if (c->partials) cudaFree(c->partials);
    if (c->frame_buf) cudaFree(c->frame_buf);
    if (c->bar) cudaFree(c->bar);
    if (c->step_out_h) cudaFreeHost(c->step_out_h);
    if (c->step_slots) cudaFree(c->step_slots);
    if (c->step_counter) cudaFree(c->step_counter);
    if (c->persist_slots) cudaFree(c->persist_slots);
    if (c->persist_counter) cudaFree(c->persist_counter);
    if (c->ser_out) cudaFree(c->ser_out);
    if (c->S) cudaFree(c->S);
    if (c->trace) cudaFree(c->trace);
    if (c->gather) cudaFree(c->gather);
    if (c->snap) cudaFreeHost(c->snap);
    if (c->ev[0]) cudaEventDestroy(c->ev[0]);
    if (c->ev[1]) cudaEventDestroy(c->ev[1]);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return PDEODE_OK;
}

int pdeode_device_count(int *count) {
    if (!count) return PDEODE_ERR_ARG;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        *count = 0;
        return PDEODE_ERR_NODEVICE;
    }
    *count = n;
    return PDEODE_OK;
}

int pdeode_slab_rows(int row, int nranks, int rank, int *row0, int *nrows) {
    if (row < 1 || nranks < 1 || rank < 0 || rank >= nranks) return PDEODE_ERR_ARG;
    const int r0 = (int)((long long)rank * row / nranks);
    const int r1 = (int)((long long)(rank + 1) * row / nranks);
    if (row0) *row0 = r0;
    if (nrows) *nrows = r1 - r0;
    return PDEODE_OK;
}

int pdeode_local_rows(const pdeode_ctx *c, int *row0, int *nrows) {
    if (!c) return PDEODE_ERR_ARG;
    // (a group context owns the whole grid: the fields below describe it)
    if (row0) *row0 = c->g.row0;
    if (nrows) *nrows = c->g.rows;
    return PDEODE_OK;
}

int pdeode_set_stream(pdeode_ctx *c, void *cuda_stream) {
    if (!c) return PDEODE_ERR_ARG;
    if (c->group) return PDEODE_ERR_ARG;             // one stream cannot serve several devices
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return PDEODE_OK;
}

int pdeode_set_state(pdeode_ctx *c, const double *M, const double *C) {
    if (!c || !M || !C) return PDEODE_ERR_ARG;
    GROUP(c, group_run(c, [&](int, pdeode_ctx *m) {
        const size_t off = (size_t)m->g.row0 * (size_t)m->g.col;
        return pdeode_set_state(m, M + off, C + off);
    }));
    CK(cudaSetDevice(c->device));
    c->iM = c->iMprev = 0; c->iC = c->iCprev = 0; c->iMnew = 1;
    int rc;
    if ((rc = upload(c, c->Mb[0], M))) return rc;
    if ((rc = upload(c, c->Cb[0], C))) return rc;
    if ((rc = zero_array(c, 1))) return rc;     // Mnew := 0, the first warm start (D8)
    c->warm = false;
    c->mass_valid = false; c->sums_valid = false;
    CK(cudaStreamSynchronize(c->stream));
    return PDEODE_OK;
}

int pdeode_set_initial_condition(pdeode_ctx *c, int ic, double depth, double height, int npts,
                                 const double *px, const double *py) {
    GROUP(c, group_run(c, [&](int, pdeode_ctx *m) {
        return pdeode_set_initial_condition(m, ic, depth, height, npts, px, py);
    }));
    if (!c || ic < 1 || ic > 10) return PDEODE_ERR_ARG;
    if (ic == 7 && (npts != 1 || !px || !py)) return PDEODE_ERR_ARG;   // the generator's seed, see the header
    const bool pts = (ic == 4 || ic == 5 || ic == 7 || ic == 9 || ic == 10);
    if (pts && (npts < 0 || (npts > 0 && (!px || !py)))) return PDEODE_ERR_ARG;
    CK(cudaSetDevice(c->device));
    c->iM = c->iMprev = 0; c->iC = c->iCprev = 0; c->iMnew = 1;
    double *dp = nullptr;
    // the point list is freed on every path out of here
    struct Free { double *&p; ~Free() { if (p) { cudaFree(p); p = nullptr; } } } free_dp{dp};
    if (pts && npts > 0) {
        CK(cudaMalloc(&dp, sizeof(double) * 2 * (size_t)npts));
        CK(cudaMemcpyAsync(dp, px, sizeof(double) * (size_t)npts, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(dp + npts, py, sizeof(double) * (size_t)npts, cudaMemcpyHostToDevice,
                           c->stream));
    }
    CK(launch_init_cond(c->g, ic, depth, height, c->params.xDel, dp ? npts : 0, dp,
                        dp ? dp + npts : nullptr, c->Mb[0], c->Cb[0], c->stream));
    c->launches++;
    if (ic == 5) {                                   // P:343-352
        CK(launch_ic_count_sum(c->g, c->Mb[0], c->S, c->partials, c->stream));
        c->launches++;
        if (multi(c)) {
            int rc = allreduce_part(c, 2);
            if (rc) return rc;
        }
        CK(launch_ic_scale(c->g, c->Mb[0], c->S, height, c->stream));
        c->launches++;
    }
    int rc = zero_array(c, 1);                       // Mnew := 0, the first warm start (D8)
    if (rc) return rc;
    c->warm = false;
    c->mass_valid = false; c->sums_valid = false;
    CK(cudaStreamSynchronize(c->stream));
    return PDEODE_OK;
}

int pdeode_get_state(pdeode_ctx *c, double *M, double *C) {
    if (!c || !M || !C) return PDEODE_ERR_ARG;
    GROUP(c, group_run(c, [&](int, pdeode_ctx *m) {
        const size_t off = (size_t)m->g.row0 * (size_t)m->g.col;
        return pdeode_get_state(m, M + off, C + off);
    }));
    CK(cudaSetDevice(c->device));
    int rc;
    if ((rc = download(c, M, c->Mb[c->iM]))) return rc;
    if ((rc = download(c, C, c->Cb[c->iC]))) return rc;
    CK(cudaStreamSynchronize(c->stream));
    return PDEODE_OK;
}

namespace {
// rows of this slab a frame with this filter contains: global rows gi, gi % filter == 0
void frame_rows(const pdeode_ctx *c, int filter, int *first, int *nro, int *nco) {
    const int r0 = c->g.row0, r1 = c->g.row0 + c->g.rows;
    const int f = ((r0 + filter - 1) / filter) * filter;
    *first = f;
    *nro = f < r1 ? (r1 - 1 - f) / filter + 1 : 0;
    *nco = (c->g.col - 1) / filter + 1;
}
int frame_buffer(pdeode_ctx *c, size_t doubles) {
    if (doubles <= c->frame_cap) return PDEODE_OK;
    if (c->frame_buf) cudaFree(c->frame_buf);
    c->frame_buf = nullptr; c->frame_cap = 0;
    CK(cudaMalloc(&c->frame_buf, doubles * sizeof(double)));
    c->frame_cap = doubles;
    return PDEODE_OK;
}
}  // namespace

int pdeode_frame_size(const pdeode_ctx *c, int filter, int *nrows_out, int *ncols_out) {
    if (!c || filter < 1) return PDEODE_ERR_ARG;
    if (c->group) {
        std::vector<int> nro; int nco = 0, tot = 0;
        group_frame_rows(const_cast<pdeode_ctx *>(c), filter, nro, &nco);
        for (int v : nro) tot += v;
        if (nrows_out) *nrows_out = tot;
        if (ncols_out) *ncols_out = nco;
        return PDEODE_OK;
    }
    int first, nro, nco;
    frame_rows(c, filter, &first, &nro, &nco);
    if (nrows_out) *nrows_out = nro;
    if (ncols_out) *ncols_out = nco;
    return PDEODE_OK;
}

int pdeode_get_frame(pdeode_ctx *c, int filter, double *M, double *C) {
    if (!c || !M || !C || filter < 1) return PDEODE_ERR_ARG;
    if (c->group) {
        std::vector<int> nro; int nco = 0;
        group_frame_rows(c, filter, nro, &nco);
        std::vector<size_t> off(nro.size(), 0);
        for (size_t k = 1; k < nro.size(); ++k) off[k] = off[k - 1] + (size_t)nro[k - 1] * (size_t)nco;
        return group_run(c, [&](int k, pdeode_ctx *m) {
            return pdeode_get_frame(m, filter, M + off[(size_t)k], C + off[(size_t)k]);
        });
    }
    CK(cudaSetDevice(c->device));
    int first, nro, nco;
    frame_rows(c, filter, &first, &nro, &nco);
    const size_t n = (size_t)nro * nco;
    if (n == 0) return PDEODE_OK;
    int rc = frame_buffer(c, 2 * n);
    if (rc) return rc;
    CK(launch_decimate(c->g, c->Mb[c->iM], c->Cb[c->iC], filter, first, nro, nco, c->frame_buf,
                       c->stream));
    c->launches++;
    CK(cudaMemcpyAsync(M, c->frame_buf, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(C, c->frame_buf + n, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return PDEODE_OK;
}

int pdeode_get_frame_rowmeans(pdeode_ctx *c, int filter, double *meanM, double *meanC) {
    if (!c || !meanM || !meanC || filter < 1) return PDEODE_ERR_ARG;
    if (c->group) {
        std::vector<int> nro; int nco = 0;
        group_frame_rows(c, filter, nro, &nco);
        std::vector<size_t> off(nro.size(), 0);
        for (size_t k = 1; k < nro.size(); ++k) off[k] = off[k - 1] + (size_t)nro[k - 1];
        return group_run(c, [&](int k, pdeode_ctx *m) {
            return pdeode_get_frame_rowmeans(m, filter, meanM + off[(size_t)k], meanC + off[(size_t)k]);
        });
    }
    CK(cudaSetDevice(c->device));
    int first, nro, nco;
    frame_rows(c, filter, &first, &nro, &nco);
    if (nro == 0) return PDEODE_OK;
    int rc = frame_buffer(c, 2 * (size_t)nro);
    if (rc) return rc;
    CK(launch_rowmeans(c->g, c->Mb[c->iM], c->Cb[c->iC], filter, first, nro, c->frame_buf, c->stream));
    c->launches++;
    CK(cudaMemcpyAsync(meanM, c->frame_buf, nro * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(meanC, c->frame_buf + nro, nro * sizeof(double), cudaMemcpyDeviceToHost,
                       c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return PDEODE_OK;
}

namespace {

// P:700-742 as one cooperative launch (k_step_persist) and one host wait.
int step_whole(pdeode_ctx *c, int nsteps) {
    c->seq++;
    StepArgs a{};
    a.st = stencil_args(c);
    a.st.trace = nullptr;
    for (int k = 0; k < 3; ++k) { a.Mb[k] = c->Mb[k]; a.Cb[k] = c->Cb[k]; }
    a.iM = c->iM; a.iC = c->iC;
    a.nsteps = nsteps;
    a.warm = c->warm ? 1 : 0;
    a.iMnew0 = c->iMnew;
    a.pbuf[0] = c->pbuf[0]; a.pbuf[1] = c->pbuf[1]; a.ip = c->ip;
    a.slots = c->step_slots; a.counter = c->step_counter; a.gen0 = c->step_gen;
    a.bar = c->bar;
    a.stamps = c->step_stamps ? 1 : 0;
    a.nstrips = c->cfg.s_nstrips; a.nchunks = c->cfg.s_nchunks; a.cw = c->cfg.s_cw;
    a.eSoln = c->params.eSoln;
    a.picard_cap = c->picard_cap;
    a.out = c->step_out;
    a.serial = c->serial ? 1 : 0; a.ser_out = c->ser_out;
    c->iMprev = c->iM; c->iCprev = c->iC;            // P:703-704
    a.seq = ++c->step_seq;
    CK(launch_step_persist(a, c->cfg.s_BS, c->stream));
    c->launches++;
    if (c->step_poll) {
        // the kernel's last act is to store `seq` behind a system-wide fence; watching that
        // word saves the driver's completion path (event, interrupt or its own polling)
        volatile unsigned long long *seq = &c->step_out_h->seq;
        for (unsigned spins = 1; *seq != a.seq; ++spins) {
            if (c->step_yield && (spins & 0x3fu) == 0u) sched_yield();   // members of an ensemble share the cores
            if ((spins & 0xfffu) == 0u) {
                cudaError_t q = cudaStreamQuery(c->stream);
                if (q == cudaSuccess) break;                 // finished: the stores have landed
                if (q != cudaErrorNotReady) return fail_cuda(c, q, "whole-step launch");
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
    } else {
        CK(cudaEventRecord(c->ev[0], c->stream));
        CK(cudaEventSynchronize(c->ev[0]));
    }
    const StepOut &o = *c->step_out_h;
    if (c->step_stamps && o.nstamps > 1) {
        fprintf(stderr, "step stamps (ns | clk):");
        for (int k = 1; k < o.nstamps; ++k)
            fprintf(stderr, " %llu|%llu", o.stamp_ns[k] - o.stamp_ns[k - 1], o.stamp_clk[k] - o.stamp_clk[k - 1]);
        fprintf(stderr, "\n");
    }
    if (o.error) {
        c->err = "whole-step launch: grid barrier timed out";
        return PDEODE_ERR_CUDA;
    }
    c->step_gen = o.gen;
    c->iM = o.iM; c->iC = o.iC; c->iMnew = o.iMnew;
    c->ip = o.ip; c->p = c->pbuf[c->ip];
    c->warm = c->warm || o.count > 0 || o.steps_done > 0;
    c->mass_valid = false;
    c->sums_valid = o.have_sums != 0;
    c->sumM = o.sumM; c->sumC = o.sumC;
    c->trace_count = std::min(o.count, std::min(MAX_PICARD_TRACE, STEP_TRACE));
    for (int k = 0; k < c->trace_count; ++k) {
        c->nit_trace[k] = o.nit[k];
        c->diff_trace[2 * k] = o.diffs[2 * k];
        c->diff_trace[2 * k + 1] = o.diffs[2 * k + 1];
    }
    if (o.picard_cap) return PDEODE_ERR_PICARD_CAP;
    return PDEODE_OK | (o.cg_cap ? PDEODE_WARN_CG_CAP : 0);
}

inline bool whole_step_path(const pdeode_ctx *c) {
    // small grids, one rank, no per-iteration CG trace asked for
    return c->cfg.whole_step && !multi(c) && c->trace_cap == 0;
}

}  // namespace

// P:700-742
int pdeode_step(pdeode_ctx *c, int *countIters, long long *nitSum, long long *nitMax) {
    if (!c) return PDEODE_ERR_ARG;
    if (c->group) {
        // every slab takes the same decisions (same all-rank sums): member 0's counts are the step's
        const int n = c->group->n;
        std::vector<int> cnt((size_t)n, 0); std::vector<long long> ns((size_t)n, 0), nm((size_t)n, 0);
        const int rc = group_run(c, [&](int k, pdeode_ctx *m) {
            return pdeode_step(m, &cnt[(size_t)k], &ns[(size_t)k], &nm[(size_t)k]);
        });
        if (countIters) *countIters = cnt[0];
        if (nitSum) *nitSum = ns[0];
        if (nitMax) *nitMax = nm[0];
        return rc;
    }
    CK(cudaSetDevice(c->device));
    if (whole_step_path(c)) {                        // the whole step in one launch
        const int rc1 = step_whole(c, 1);
        if ((rc1 & PDEODE_ERR_MASK) && (rc1 & PDEODE_ERR_MASK) != PDEODE_ERR_PICARD_CAP) return rc1;
        const StepOut &o = *c->step_out_h;
        if (countIters) *countIters = o.count;
        if (nitSum) *nitSum = o.nit_sum;
        if (nitMax) *nitMax = o.nit_max;
        return rc1;
    }
    int rc;
    int warn = 0;
    // Cprev = C ; Mprev = M (P:703-704): the current buffers become the frozen ones
    c->iMprev = c->iM; c->iCprev = c->iC;
    // d = D(Mprev) (P:544-546); fixed over the Picard loop because GenMatrixM
    // is always handed Mprev (P:712)
    CK(launch_diffcoef(c->g, c->np, c->Mb[c->iMprev], c->d, c->stream));
    c->launches++;
    if ((rc = halo_exchange(c, c->d))) return rc;

    double diffC = 1.0, diffM = 1.0;                 // P:700-701
    int count = 0;                                   // P:702
    long long sum = 0, mx = 0;
    c->trace_count = 0;
    while (diffC + diffM > c->params.eSoln) {        // P:706
        // target buffer of this solve: neither Mprev nor the current iterate
        int inew = 0;
        while (inew == c->iMprev || inew == c->iM) inew++;
        if (!c->warm) inew = c->iMnew;               // the zeroed buffer (first solve ever)
        const double *x0 = c->warm ? c->Mb[c->iM] : c->Mb[inew];
        long long nit = 0; int stop = 0;
        int icnew = 0;
        while (icnew == c->iCprev || icnew == c->iC) icnew++;
        SolveCArgs s{};
        s.g = c->g; s.np = c->np; s.S = c->S; s.partials = c->partials;
        s.Mprev = c->Mb[c->iMprev]; s.Mnew = c->Mb[inew]; s.Cprev = c->Cb[c->iCprev];
        s.Ccur = c->Cb[c->iC]; s.Mcur = c->Mb[c->iM]; s.Cnew = c->Cb[icnew];
        s.n2 = (long long)c->g.rows * c->g.P / 2;
        s.fuse_scalars = (multi(c) || c->serial) ? 0 : 1;
        bool solvec_done = false;
        if ((rc = solve(c, x0, c->Mb[inew], count, &nit, &stop, &s, &solvec_done))) return rc;
        c->warm = true;
        c->iMnew = inew;

        if (!solvec_done) {
            s.seq = c->seq; s.gated = 0;
            CK(launch_solvec(s, c->stream));             // P:717, P:725-726
            c->launches++;
            if (c->serial) {                             // calcDiff x2 in cell order (P:956-960)
                const SerialJob jb[2] = {{s.Ccur, s.Cnew, SER_ABSDIFF, 0}, {s.Mcur, s.Mnew, SER_ABSDIFF, 0}};
                if ((rc = serial_sums(c, jb, 2, c->S->part))) return rc;
            }
            if (multi(c) || c->serial) {
                if (multi(c) && (rc = allreduce_part(c, 4))) return rc;
                CK(launch_scal_after_solvec(c->S, c->g, c->stream));
                c->launches++;
            }
            if ((rc = snapshot(c, 0))) return rc;
            if ((rc = wait_snapshot(c, 0))) return rc;
        }
        diffC = c->snap[0].s.diffC;
        diffM = c->snap[0].s.diffM;
        if (c->pending_persist) {
            // the persistent solve was not waited for on its own: this one record
            // carries its iteration count next to the Picard residuals
            const DevScal &fin = c->snap[0].s;
            c->pending_persist = false;
            if (fin.error) {
                c->err = multi(c) ? "peer exchange timed out (a rank stopped posting)"
                          : "grid barrier of the persistent solve timed out";
                return PDEODE_ERR_CUDA;
            }
            c->phase += 2ULL * (unsigned long long)fin.nit;   // two exchanges per iteration
            // = two grid-wide sums per iteration (serial order: each is followed by a second barrier)
            c->persist_gen += (c->serial ? 4ULL : 2ULL) * (unsigned long long)fin.nit;
            c->ip = c->pending_ip0 ^ (int)(fin.nit & 1);
            c->p = c->pbuf[c->ip];
            nit = fin.nit;
            stop = fin.stop;
        }
        if (stop == -1 || stop == -101) warn = PDEODE_WARN_CG_CAP;

        if (nit > mx) mx = nit;                      // P:722
        sum += nit;                                  // P:723
        c->iC = icnew;                               // C = Cnew (P:728)
        c->iM = inew;                                // M = Mnew (P:729)
        if (count < MAX_PICARD_TRACE) {
            c->nit_trace[count] = nit;
            c->diff_trace[2 * count] = diffC;
            c->diff_trace[2 * count + 1] = diffM;
            c->trace_count = count + 1;
        }
        count++;                                     // P:730
        c->mass_valid = false;
        // calcMass of the new iterate came with the C update (not in serial-order verification,
        // where pdeode_mass adds in cell order itself)
        c->sums_valid = !c->serial;
        c->sumM = c->snap[0].s.sumM; c->sumC = c->snap[0].s.sumC;
        if (count > c->picard_cap) {                 // P:732
            if (countIters) *countIters = count;
            if (nitSum) *nitSum = sum;
            if (nitMax) *nitMax = mx;
            return PDEODE_ERR_PICARD_CAP;
        }
    }
    if (countIters) *countIters = count;
    if (nitSum) *nitSum = sum;
    if (nitMax) *nitMax = mx;
    return PDEODE_OK | warn;
}

// n time steps with what solveOrder2 accumulates over them (P:722-723, P:739-740)
int pdeode_step_many(pdeode_ctx *c, int n, int *stepsDone, long long *countSum, int *countMax,
                     long long *nitSum, long long *nitMax) {
    if (!c || n < 0) return PDEODE_ERR_ARG;
    if (c->group) {
        const int ng = c->group->n;
        std::vector<int> sd((size_t)ng, 0), cm((size_t)ng, 0);
        std::vector<long long> cs((size_t)ng, 0), ns((size_t)ng, 0), nm((size_t)ng, 0);
        const int rc = group_run(c, [&](int k, pdeode_ctx *m) {
            return pdeode_step_many(m, n, &sd[(size_t)k], &cs[(size_t)k], &cm[(size_t)k], &ns[(size_t)k], &nm[(size_t)k]);
        });
        if (stepsDone) *stepsDone = sd[0];
        if (countSum) *countSum = cs[0];
        if (countMax) *countMax = cm[0];
        if (nitSum) *nitSum = ns[0];
        if (nitMax) *nitMax = nm[0];
        return rc;
    }
    CK(cudaSetDevice(c->device));
    int done = 0, cmax = 0, warn = 0, rc = PDEODE_OK;
    long long csum = 0, nsum = 0, nmax = 0;
    if (whole_step_path(c)) {
        // one launch and one host wait per batch; the report holds STEP_TRACE-independent sums
        while (done < n) {
            const int batch = std::min(n - done, 4096);
            rc = step_whole(c, batch);
            const int err = rc & PDEODE_ERR_MASK;
            if (err && err != PDEODE_ERR_PICARD_CAP) break;      // the outputs below are still filled
            const StepOut &o = *c->step_out_h;
            done += o.steps_done;
            csum += o.count_sum; nsum += o.nit_sum_all;
            cmax = std::max(cmax, o.count_max);
            nmax = std::max(nmax, o.nit_max_all);
            warn |= rc & PDEODE_WARN_CG_CAP;
            if (err) break;
            rc = PDEODE_OK;
        }
    } else {
        for (; done < n; ++done) {
            int cnt = 0; long long s1 = 0, m1 = 0;
            rc = pdeode_step(c, &cnt, &s1, &m1);
            if (rc & PDEODE_ERR_MASK) break;
            warn |= rc & PDEODE_WARN_CG_CAP;
            rc = PDEODE_OK;
            csum += cnt; nsum += s1;
            cmax = std::max(cmax, cnt);
            nmax = std::max(nmax, m1);
        }
    }
    if (stepsDone) *stepsDone = done;
    if (countSum) *countSum = csum;
    if (countMax) *countMax = cmax;
    if (nitSum) *nitSum = nsum;
    if (nitMax) *nitMax = nmax;
    return (rc & PDEODE_ERR_MASK) ? rc : (PDEODE_OK | warn);
}

int pdeode_mass(pdeode_ctx *c, double *meanM, double *meanC) {
    if (!c) return PDEODE_ERR_ARG;
    if (c->group) {
        const int n = c->group->n;
        std::vector<double> a((size_t)n, 0.0), b((size_t)n, 0.0);
        const int rc = group_run(c, [&](int k, pdeode_ctx *m) { return pdeode_mass(m, &a[(size_t)k], &b[(size_t)k]); });
        if (meanM) *meanM = a[0];
        if (meanC) *meanC = b[0];
        return rc;
    }
    CK(cudaSetDevice(c->device));
    if (c->sums_valid && !c->mass_valid) {
        // the sums came with the last C update of the step (k_solvec / the whole-step launch)
        if (meanM) *meanM = c->sumM / c->g.n_glob;           // P:789
        if (meanC) *meanC = c->sumC / c->g.n_glob;
        return PDEODE_OK;
    }
    int rc = run_mass(c);
    if (rc) return rc;
    if (meanM) *meanM = c->snap[0].s.sumM / c->g.n_glob;     // P:789
    if (meanC) *meanC = c->snap[0].s.sumC / c->g.n_glob;
    return PDEODE_OK;
}

int pdeode_peak(pdeode_ctx *c, double *peak, double *height, double *intfac) {
    if (!c) return PDEODE_ERR_ARG;
    if (c->group) {
        // outputs the reference would leave unassigned stay untouched: start every member from the caller's values
        const int n = c->group->n;
        std::vector<double> pk((size_t)n, peak ? *peak : 0.0), he((size_t)n, height ? *height : 0.0),
            it((size_t)n, intfac ? *intfac : 0.0);
        const int rc = group_run(c, [&](int k, pdeode_ctx *m) {
            return pdeode_peak(m, &pk[(size_t)k], &he[(size_t)k], &it[(size_t)k]);
        });
        if (peak) *peak = pk[0];
        if (height) *height = he[0];
        if (intfac) *intfac = it[0];
        return rc;
    }
    CK(cudaSetDevice(c->device));
    int rc = run_mass(c);
    if (rc) return rc;
    const DevScal &s = c->snap[0].s;
    const double denom = (double)(c->g.grow - 1);            // P:990
    if (s.peak_idx >= 0) {
        if (peak) *peak = (double)(s.peak_idx / c->g.col) / denom;
        if (height) *height = s.peak_val;
    } else if (height) {
        *height = 0.0;                                       // hei = 0 (P:988,1002)
    }
    if (s.intfac_row >= 0 && intfac) *intfac = (double)s.intfac_row / denom;
    return PDEODE_OK;
}

int pdeode_set_maxit(pdeode_ctx *c, long long maxit) {
    if (!c) return PDEODE_ERR_ARG;
    GROUP(c, group_run(c, [&](int, pdeode_ctx *m) { return pdeode_set_maxit(m, maxit); }));
    CK(cudaSetDevice(c->device));
    c->maxit = maxit >= 0 ? maxit : c->maxit_default;
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(&c->S->maxit, &c->maxit, sizeof(long long), cudaMemcpyHostToDevice));
    return PDEODE_OK;
}

int pdeode_set_picard_cap(pdeode_ctx *c, int cap) {
    if (!c) return PDEODE_ERR_ARG;
    GROUP(c, group_run(c, [&](int, pdeode_ctx *m) { return pdeode_set_picard_cap(m, cap); }));
    c->picard_cap = cap >= 0 ? cap : PICARD_CAP;
    return PDEODE_OK;
}

int pdeode_last_step_trace(pdeode_ctx *c, long long *nit_per_solve, double *diffs, int cap,
                           int *count) {
    if (!c) return PDEODE_ERR_ARG;
    GROUP(c, pdeode_last_step_trace(c->group->m[0], nit_per_solve, diffs, cap, count));
    const int k = std::min(cap, c->trace_count);
    for (int i = 0; i < k; ++i) {
        if (nit_per_solve) nit_per_solve[i] = c->nit_trace[i];
        if (diffs) { diffs[2 * i] = c->diff_trace[2 * i]; diffs[2 * i + 1] = c->diff_trace[2 * i + 1]; }
    }
    if (count) *count = c->trace_count;
    return PDEODE_OK;
}

int pdeode_cg_trace_enable(pdeode_ctx *c, int cap) {
    if (!c || cap < 0) return PDEODE_ERR_ARG;
    GROUP(c, group_run(c, [&](int, pdeode_ctx *m) { return pdeode_cg_trace_enable(m, cap); }));
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    if (c->trace) { cudaFree(c->trace); c->trace = nullptr; }
    c->trace_cap = cap;
    if (cap > 0) CK(cudaMalloc(&c->trace, sizeof(double) * 5 * (size_t)cap));
    CK(cudaMemcpy(&c->S->trace_cap, &cap, sizeof(int), cudaMemcpyHostToDevice));
    return PDEODE_OK;
}

int pdeode_cg_trace_get(pdeode_ctx *c, double *out, int cap, int *count) {
    if (!c || !out) return PDEODE_ERR_ARG;
    GROUP(c, pdeode_cg_trace_get(c->group->m[0], out, cap, count));
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    int len = 0;
    CK(cudaMemcpy(&len, &c->S->trace_len, sizeof(int), cudaMemcpyDeviceToHost));
    const int k = std::min(std::min(len, cap), c->trace_cap);
    if (k > 0) CK(cudaMemcpy(out, c->trace, sizeof(double) * 5 * (size_t)k, cudaMemcpyDeviceToHost));
    if (count) *count = len;
    return PDEODE_OK;
}

int pdeode_debug_get_array(pdeode_ctx *c, int which, double *out) {
    if (!c || !out) return PDEODE_ERR_ARG;
    GROUP(c, group_run(c, [&](int, pdeode_ctx *m) {
        return pdeode_debug_get_array(m, which, out + (size_t)m->g.row0 * (size_t)m->g.col);
    }));
    CK(cudaSetDevice(c->device));
    const double *src = nullptr;
    switch (which) {
    case PDEODE_ARR_M: src = c->Mb[c->iM]; break;
    case PDEODE_ARR_C: src = c->Cb[c->iC]; break;
    case PDEODE_ARR_MNEW: src = c->Mb[c->iMnew]; break;
    case PDEODE_ARR_D: src = c->d; break;
    case PDEODE_ARR_AC: src = c->ac; break;
    case PDEODE_ARR_R: src = c->r; break;
    case PDEODE_ARR_P: src = c->p; break;
    case PDEODE_ARR_Q: src = c->q; break;
    default: return PDEODE_ERR_ARG;
    }
    int rc = download(c, out, src);
    if (rc) return rc;
    CK(cudaStreamSynchronize(c->stream));
    return PDEODE_OK;
}

int pdeode_exchange_mode(const pdeode_ctx *c, int *mode) {
    if (!c || !mode) return PDEODE_ERR_ARG;
    GROUP(c, pdeode_exchange_mode(c->group->m[0], mode));
    *mode = !multi(c) ? 0 : (c->peer ? 2 : 1);
    return PDEODE_OK;
}

int pdeode_engine_info(const pdeode_ctx *c, char *buf, int len) {
    if (!c || !buf || len < 1) return PDEODE_ERR_ARG;
    if (c->group) {
        char one[512];
        const int rc = pdeode_engine_info(c->group->m[0], one, (int)sizeof one);
        snprintf(buf, (size_t)len, "%s gpus=%d(one-process)", one, c->group->n);
        return rc;
    }
    snprintf(buf, (size_t)len,
             "spmv=%s BS=%d TR=%d depth=%d init_BS=%d init_TR=%d persist=%d persist_ctas=%d "
             "defer_x=%d kb_reverse=%d whole_step=%d step_BS=%d step_ctas=%d exchange=%s sum_order=%s "
             "persist_prefetch=%d",
             c->cfg.variant == 1 ? "bulk-async" : "register-staged", c->cfg.BS, c->cfg.TR,
             c->cfg.DEPTH, c->cfg.i_BS, c->cfg.i_TR, c->cfg.persist,
             c->cfg.p_nstrips * c->cfg.p_nchunks,
             (c->cfg.persist && sums_mode(c) != SUMS_RAW ? c->cfg.persist_defer != 0 : defer_x(c)) ? 1 : 0,
             c->cfg.kb_reverse, c->cfg.whole_step, c->cfg.s_BS,
             c->cfg.s_nstrips * c->cfg.s_nchunks,
             !multi(c) ? "none" : (c->peer ? "peer" : "nccl"), c->serial ? "serial" : "tree",
             persist_prefetch_rows(c));
    return PDEODE_OK;
}

int pdeode_launch_count(const pdeode_ctx *c, long long *launches) {
    if (!c || !launches) return PDEODE_ERR_ARG;
    if (c->group) {
        long long tot = 0;
        for (pdeode_ctx *m : c->group->m) tot += m ? m->launches : 0;
        *launches = tot;
        return PDEODE_OK;
    }
    *launches = c->launches;
    return PDEODE_OK;
}

int pdeode_kernel_bytes_per_cell(int kernel) {
    switch (kernel) {
    case PDEODE_KERNEL_CG_SPMV: return 48;    // R r,p,d,ac  W p,q
    case PDEODE_KERNEL_CG_UPDATE: return 48;  // R x,p,q,r   W x,r
    case PDEODE_KERNEL_INIT: return 56;       // R x0,d,C,Mprev  W ac,r,x
    case PDEODE_KERNEL_SOLVEC: return 48;     // R Mprev,Mnew,Cprev,C,M  W Cnew
    case PDEODE_KERNEL_DIFFCOEF: return 16;   // R Mprev  W d
    case PDEODE_KERNEL_CG_SPMV_X: return 64;  // R r,p,d,ac,x  W p,q,x
    case PDEODE_KERNEL_CG_UPDATE_R: return 24;  // R r,q  W r
    default: return 0;
    }
}

int pdeode_time_kernel(pdeode_ctx *c, int kernel, int reps, float *ms_per_launch) {
    if (!c || reps < 1 || !ms_per_launch) return PDEODE_ERR_ARG;
    if (c->group) return PDEODE_ERR_ARG;
    if (multi(c)) return PDEODE_ERR_ARG;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    // put the CG scalars into a neutral mid-solve state: beta = 0.5, alfa = 1e-3
    DevScal s{};
    CK(cudaMemcpy(&s, c->S, sizeof s, cudaMemcpyDeviceToHost));
    s.rho = 1.0; s.rho_old = 2.0; s.alfa = 1e-3; s.nit = 1; s.done = 0; s.finished = 0;
    s.maxit = c->maxit;
    for (int k = 0; k < 8; ++k) s.ticket[k] = 0;
    CK(cudaMemcpy(c->S, &s, sizeof s, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    StencilArgs a = stencil_args(c);
    a.fuse_scalars = 0;
    a.x0 = c->Mb[c->iM]; a.x = c->Mb[(c->iM + 1) % 3]; a.Cc = c->Cb[c->iC]; a.Mprev = c->Mb[c->iM];
    UpdateArgs u{};
    u.g = c->g; u.S = c->S; u.partials = c->partials; u.x = c->Mb[(c->iM + 1) % 3]; u.r = c->r;
    u.p = c->p; u.q = c->q; u.n2 = (long long)c->g.rows * c->g.P / 2; u.fuse_scalars = 0;
    u.reverse = c->cfg.kb_reverse;
    SolveCArgs sc{};
    sc.g = c->g; sc.np = c->np; sc.S = c->S; sc.partials = c->partials;
    sc.Mprev = c->Mb[c->iM]; sc.Mnew = c->Mb[(c->iM + 1) % 3]; sc.Cprev = c->Cb[c->iC];
    sc.Ccur = c->Cb[c->iC]; sc.Mcur = c->Mb[c->iM]; sc.Cnew = c->Cb[(c->iC + 1) % 3];
    sc.n2 = u.n2; sc.fuse_scalars = 0;
    for (int it = -1; it < reps; ++it) {     // one untimed launch first
        if (it == 0) CK(cudaEventRecord(e0, c->stream));
        switch (kernel) {
        case PDEODE_KERNEL_CG_SPMV: CK(launch_stencil_cg(a, c->stream)); break;
        case PDEODE_KERNEL_CG_UPDATE: CK(launch_update(u, c->stream)); break;
        case PDEODE_KERNEL_CG_SPMV_X: {
            StencilArgs ax = a;
            ax.defer_x = 1; ax.x = u.x; ax.variant = 1; ax.TR = c->cfg.TR_x;
            CK(launch_stencil_cg(ax, c->stream));
            break;
        }
        case PDEODE_KERNEL_CG_UPDATE_R: CK(launch_update_r(u, c->stream)); break;
        case PDEODE_KERNEL_INIT: {
            StencilArgs ai = a;
            ai.BS = c->cfg.i_BS; ai.TR = c->cfg.i_TR;
            CK(launch_stencil_init(ai, c->stream));
            break;
        }
        case PDEODE_KERNEL_SOLVEC: CK(launch_solvec(sc, c->stream)); break;
        case PDEODE_KERNEL_DIFFCOEF:
            CK(launch_diffcoef(c->g, c->np, c->Mb[c->iM], c->d, c->stream)); break;
        default: return PDEODE_ERR_ARG;
        }
        c->launches++;
    }
    CK(cudaEventRecord(e1, c->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    *ms_per_launch = ms / (float)reps;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    c->mass_valid = false; c->sums_valid = false;
    return PDEODE_OK;
}

}  // extern "C"
