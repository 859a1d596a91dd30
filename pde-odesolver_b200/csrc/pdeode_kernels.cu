// pdeode_kernels.cu -- fp64 sm_100a kernels of the implicit time step.
//
// Compiled with -fmad=false: the reference build (gfortran -O3 on baseline
// x86-64, runAll.sh:6) has no fused multiply-add, so every product and sum
// below rounds separately, in the order the cited reference line evaluates it.
// Division and sqrt are IEEE-correct in CUDA fp64.  Element-wise results are
// therefore bit-identical to the reference's; only the reductions (parallel
// tree here, serial loop there) differ, at the 1e-16 level.
//
// Kernels (P: = PDE-ODEsolver.f90, CG: = CGdiag.f90 of the reference):
//   k_diffcoef   d = D(Mprev)                                  P:544-546
//   k_stencil<INIT>  centre diagonal, r = b - A x0, x = x0     P:551-586, CG:46-50
//   k_stencil<CG>    p = r + beta p ; q = A p ; p.q ; p.p      CG:61-76,87  P:1087-1102
//   k_update     x += alfa p ; r -= alfa q ; r.r ; x.x         CG:57,59,79-86
//   k_solvec     C update + Picard residuals                   P:504-506, P:948-963
//   k_mass       calcMass + calcPeakInterface                  P:772-791, P:978-1004
// The 5-diagonal matrix is never stored: off-diagonals are rebuilt from
// d = D(Mprev) in registers, the centre diagonal `ac` is stored once per
// Picard iterate.
#include "pdeode_device.h"

#include <atomic>
#include <cstdlib>

namespace pdeode {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int MAX_DEVICES = 64;

int g_stream_ctas_per_sm = 16;   // grid of the streaming kernels, in CTAs per SM
int g_num_sms = 148;

// ---------------------------------------------------------------- helpers

__device__ __forceinline__ double2 ld2(const double *p) {
    return *reinterpret_cast<const double2 *>(p);
}
__device__ __forceinline__ void st2(double *p, double2 v) {
    *reinterpret_cast<double2 *>(p) = v;
}

// libgcc __powidf2: what gfortran emits for real**integer_variable (P:470-471)
__device__ __forceinline__ double powi(double x, int m) {
    unsigned int n = (m < 0) ? (0u - (unsigned int)m) : (unsigned int)m;
    double y = (n & 1u) ? x : 1.0;
    while (n >>= 1) {
        x = x * x;
        if (n & 1u) y *= x;
    }
    return (m < 0) ? 1.0 / y : y;
}

// The same multiplications in the same order with the exponent known at compile time: the
// square-and-multiply loop unrolls into two to four DMULs (a product with the initial 1.0 is
// exact, so dropping it changes no bit).
template <int M_>
__device__ __forceinline__ double powi_c(double x) {
    static_assert(M_ >= 0, "non-negative exponents only");
    unsigned int n = (unsigned int)M_;
    double y = (n & 1u) ? x : 1.0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        n >>= 1;
        if (n == 0u) break;
        x = x * x;
        if (n & 1u) y *= x;
    }
    return y;
}
// Exponents 0..8 (the shipped file has alpha = beta = 4, README.md:42-44 goes to 6) take the
// unrolled form -- one uniform jump instead of a shift / test / branch per bit; anything else
// runs the loop.
__device__ __forceinline__ double powi_fast(double x, int m) {
    switch (m) {
    case 0: return powi_c<0>(x);
    case 1: return powi_c<1>(x);
    case 2: return powi_c<2>(x);
    case 3: return powi_c<3>(x);
    case 4: return powi_c<4>(x);
    case 5: return powi_c<5>(x);
    case 6: return powi_c<6>(x);
    case 7: return powi_c<7>(x);
    case 8: return powi_c<8>(x);
    default: return powi(x, m);
    }
}

// P:464-472
__device__ __forceinline__ double dfunc(double M, const NumParams &np) {
    double d = 0.0;
    if (np.dSelect == 1) d = np.delta;
    else if (np.dSelect == 2) d = np.delta * powi_fast(M, np.alpha);
    else if (np.dSelect == 3) d = np.delta * powi_fast(M, np.alpha) / powi_fast(1.0 - M, np.beta);
    return d;
}

// P:445-458
__device__ __forceinline__ double ffunc(double M, double C, const NumParams &np) {
    const double eps = C * 0.00000001;
    switch (np.fSelect) {
    case 1: return C / (np.kappa + C) - np.nu;
    case 2: return C / (np.kappa + C);
    case 3: return C / (np.kappa * M + C + eps) - np.nu;
    case 4: return C / (np.kappa * M + C + eps);
    case 5: return 1.0;
    case 6: return C / (np.kappa + C) * (1.0 - M / (C + eps)) - np.nu;
    default: return 0.0;
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// Sum of (a, b) over the block; result valid in thread 0.  sm: 64 doubles.
__device__ __forceinline__ void block_sum2(double &a, double &b, double *sm) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) { sm[w] = a; sm[32 + w] = b; }
    __syncthreads();
    if (w == 0) {
        a = (lane < nw) ? sm[lane] : 0.0;
        b = (lane < nw) ? sm[32 + lane] : 0.0;
        a = warp_sum(a);
        b = warp_sum(b);
    }
    __syncthreads();
}

// Two-stage deterministic reduction: every CTA stores its partial pair, the
// CTA that draws the last ticket adds all partials in a fixed order.  Returns
// true in thread 0 of that last CTA with (a, b) = grid totals.
__device__ __forceinline__ bool grid_sum2(double &a, double &b, double *partials, int nb,
                                          int bid, unsigned int *ticket, double *sm,
                                          bool sysfence = false) {
    __shared__ int s_last;
    // peer mode: this CTA may have stored ghost rows into a neighbour GPU; they
    // must be visible system-wide before the last CTA posts the sums
    if (sysfence) __threadfence_system();
    block_sum2(a, b, sm);
    if (threadIdx.x == 0) {
        partials[bid] = a;
        partials[nb + bid] = b;
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == (unsigned int)(nb - 1));
    }
    __syncthreads();
    if (!s_last) return false;
    if (sysfence) __threadfence_system();   // a peer_post follows: acquire here, release to all GPUs
    else __threadfence();
    double x = 0.0, y = 0.0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
        x += __ldcg(partials + i);
        y += __ldcg(partials + nb + i);
    }
    block_sum2(x, y, sm);
    a = x; b = y;
    if (threadIdx.x == 0) *ticket = 0u;
    return threadIdx.x == 0;
}

// The same for NV values (partials: [NV][nb]); sm: 64 doubles.
template <int NV>
__device__ __forceinline__ bool grid_sum(double (&v)[NV], double *partials, int nb, int bid,
                                         unsigned int *ticket, double *sm) {
    __shared__ int s_lastN;
    double z = 0.0;
#pragma unroll
    for (int k = 0; k + 1 < NV; k += 2) block_sum2(v[k], v[k + 1], sm);
    if (NV & 1) block_sum2(v[NV - 1], z, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) partials[k * nb + bid] = v[k];
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        s_lastN = (t == (unsigned int)(nb - 1));
    }
    __syncthreads();
    if (!s_lastN) return false;
    __threadfence();
    double x[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) x[k] = 0.0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
#pragma unroll
        for (int k = 0; k < NV; ++k) x[k] += __ldcg(partials + k * nb + i);
    }
#pragma unroll
    for (int k = 0; k + 1 < NV; k += 2) block_sum2(x[k], x[k + 1], sm);
    if (NV & 1) block_sum2(x[NV - 1], z, sm);
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = x[k];
    if (threadIdx.x == 0) *ticket = 0u;
    return threadIdx.x == 0;
}

// Same for a triple (partials: [3][nb]).
__device__ __forceinline__ bool grid_sum3(double &a, double &b, double &c, double *partials,
                                          int nb, int bid, unsigned int *ticket, double *sm,
                                          bool sysfence = false) {
    __shared__ int s_last3;
    if (sysfence) __threadfence_system();
    double z0 = 0.0;
    block_sum2(a, b, sm);
    block_sum2(c, z0, sm);
    if (threadIdx.x == 0) {
        partials[bid] = a;
        partials[nb + bid] = b;
        partials[2 * nb + bid] = c;
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        s_last3 = (t == (unsigned int)(nb - 1));
    }
    __syncthreads();
    if (!s_last3) return false;
    if (sysfence) __threadfence_system();
    else __threadfence();
    double x = 0.0, y = 0.0, z = 0.0, z1 = 0.0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
        x += __ldcg(partials + i);
        y += __ldcg(partials + nb + i);
        z += __ldcg(partials + 2 * nb + i);
    }
    block_sum2(x, y, sm);
    block_sum2(z, z1, sm);
    a = x; b = y; c = z;
    if (threadIdx.x == 0) *ticket = 0u;
    return threadIdx.x == 0;
}

// ---------------------------------------------------------------- scalars

// CG:53 (nit = 0, stopcrit = 0) with rho / sum(sol*sol) of the start vector
__device__ __forceinline__ void scal_after_init(DevScal *S, double rr, double xx) {
    S->rho = rr;          // dotProd(r,r) seen by iteration 1 (CG:59)
    S->rho_old = 0.0;     // CG:44
    S->xx = xx;           // sum(sol*sol) seen by iteration 1 (CG:57)
    S->nit = 0; S->stop = 0; S->done = 0; S->finished = 0; S->tail = 0;
    S->trace_len = 0;
}

// CG:55,77,87-90
__device__ __forceinline__ void scal_after_spmv(DevScal *S, double pq, double pp,
                                                double n_glob, double tol2, double *trace) {
    const long long nit = S->nit + 1;                      // CG:55
    const double alfa = S->rho / pq;                       // CG:77
    const double err2 = fabs(alfa) * sqrt(pp) / n_glob;    // CG:87
    const double normx = sqrt(S->xx) / n_glob;             // CG:57
    int stop = 0;
    if (nit > S->maxit) stop = -1;                         // CG:88
    if (err2 < tol2 * normx) stop = stop - 100;            // CG:90
    if (trace && nit <= (long long)S->trace_cap) {
        double *t = trace + 5 * (nit - 1);
        t[0] = S->rho; t[1] = pq; t[2] = alfa; t[3] = err2; t[4] = normx;
        S->trace_len = (int)nit;
    }
    S->pq = pq; S->pp = pp; S->alfa = alfa; S->err2 = err2; S->normx = normx;
    S->nit = nit; S->stop = stop; S->done = (stop != 0);
}

__device__ __forceinline__ void scal_after_update(DevScal *S, double rr, double xx) {
    S->rho_old = S->rho;  // CG:56
    S->rho = rr;          // CG:59 of the next iteration
    S->xx = xx;           // CG:57 of the next iteration
    S->finished = S->done;
}

// P:961
__device__ __forceinline__ void scal_after_solvec(DevScal *S, double dC, double dM,
                                                  double n_glob) {
    S->diffC = dC / n_glob;
    S->diffM = dM / n_glob;
}
// sums of the new iterate, for calcMass (P:783-787) without another pass over the fields
__device__ __forceinline__ void scal_mass_sums(DevScal *S, double sM, double sC) {
    S->sumM = sM;
    S->sumC = sC;
}

__global__ void k_scal_after_init(DevScal *S) { scal_after_init(S, S->part[0], S->part[1]); }
__global__ void k_scal_after_spmv(DevScal *S, GridDesc g, double tol2, double *trace) {
    if (S->done) return;
    scal_after_spmv(S, S->part[0], S->part[1], g.n_glob, tol2, trace);
}
__global__ void k_scal_after_update(DevScal *S) {
    if (S->finished) return;
    scal_after_update(S, S->part[0], S->part[1]);
}
__global__ void k_scal_after_solvec(DevScal *S, GridDesc g) {
    scal_after_solvec(S, S->part[0], S->part[1], g.n_glob);
    scal_mass_sums(S, S->part[2], S->part[3]);
}

// ---------------------------------------------------------------- serial-order sums (verification)

// PDEODE_SUM_ORDER=serial.  nj reductions over the cells of the grid in the
// reference's order: acc = 0; acc = acc + term(cell) for cell = 1 .. n (row
// major), exactly the loops of dotProd (CG:108-112), sum(sol*sol) (CG:57),
// calcDiff (P:956-960) and calcMass (P:783-787).  The terms -- u*v, |a-b|, a --
// are the same bits whoever computes them, so all threads of the CTA stage a
// chunk of terms in shared memory (stage: [2][SER_MAXJ][SER_CH] doubles) while
// thread 0 adds the previous chunk one term after the other.  All threads of the
// CTA call this; out[0 .. nj) is written by thread 0.  Loads bypass L1: inside
// the persistent kernels the arrays were written by other CTAs of the same launch.
__device__ __forceinline__ void serial_reduce(const GridDesc &g, const SerialJob *jobs, const int nj,
                                              double *stage, double *out) {
    const long long n = (long long)g.rows * g.col;
    const int nt = blockDim.x, tid = threadIdx.x;
    auto fill = [&](long long c0, double *buf) {
        for (int e = tid; e < SER_CH; e += nt) {
            const long long k = c0 + e;
            if (k >= n) break;
            const long long i = k / g.col;
            const long long off = i * g.P + (k - i * g.col);
            for (int q = 0; q < nj; ++q) {
                const double a = __ldcg(jobs[q].a + off);
                double t = a;
                if (jobs[q].kind == SER_DOT) t = a * __ldcg(jobs[q].b + off);
                else if (jobs[q].kind == SER_ABSDIFF) t = fabs(a - __ldcg(jobs[q].b + off));
                buf[q * SER_CH + e] = t;
            }
        }
    };
    double acc[SER_MAXJ];
#pragma unroll
    for (int q = 0; q < SER_MAXJ; ++q) acc[q] = 0.0;
    __syncthreads();                     // the staging area may still be read by a previous call
    fill(0, stage);
    __syncthreads();
    int cur = 0;
    for (long long c0 = 0; c0 < n; c0 += SER_CH) {
        const double *buf = stage + (size_t)cur * SER_MAXJ * SER_CH;
        if (c0 + SER_CH < n) fill(c0 + SER_CH, stage + (size_t)(cur ^ 1) * SER_MAXJ * SER_CH);
        if (tid == 0) {
            const int len = (int)((n - c0 < SER_CH) ? (n - c0) : SER_CH);
            for (int e = 0; e < len; ++e) {
#pragma unroll
                for (int q = 0; q < SER_MAXJ; ++q)
                    if (q < nj) acc[q] = acc[q] + buf[q * SER_CH + e];
            }
        }
        __syncthreads();
        cur ^= 1;
    }
    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < SER_MAXJ; ++q)
            if (q < nj) out[q] = acc[q];
    }
}

struct SerialArgs {
    GridDesc g;
    SerialJob job[SER_MAXJ];
    int nj;
    double *out;
};

__global__ void __launch_bounds__(256) k_serial_reduce(const SerialArgs a) {
    __shared__ double stage[2 * SER_MAXJ * SER_CH];
    serial_reduce(a.g, a.job, a.nj, stage, a.out);
}

// ---------------------------------------------------------------- peer-memory exchange

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Posted sums travel as "LL" words: every 8-byte word carries 32 bits of
// payload and the low 32 bits of the phase number, so a word is its own flag
// (an aligned 8-byte store is indivisible) and the poster needs no fence
// between data and flag -- one NVLink trip instead of a round trip plus a trip.
// Three doubles = six words per (slot, rank).  A slot is reused every
// PEER_SLOTS phases; stale words carry another tag and never match.
__device__ __forceinline__ unsigned long long ll_word(unsigned int half, unsigned long long phase) {
    return ((unsigned long long)half << 32) | (unsigned long long)(unsigned int)phase;
}

// All threads of the CTA call this.  Lane r of warp 0 watches the six words
// rank r posts into this rank's comm block -- all ranks at once, six loads in
// flight per lane -- until they carry this phase's tag, reads one of them
// again as an acquire (the neighbours' ghost rows of r were complete before they posted),
// and the eight lanes add the sums in a fixed tree: the same on every rank and
// in every kernel, so every rank gets the same bits.  A peer that never posts
// (crashed) makes the wait give up after PeerComm::timeout_clocks GPU clocks
// (PDEODE_PEER_TIMEOUT_MS, default 10 s) and raise DevScal::error instead of
// hanging the device.  Returns false in that case; the caller must then leave
// without storing anything (the sums are zeros) and mark the solve finished so
// that the launches queued behind it return at once.
__device__ __forceinline__ bool peer_wait_sums(const PeerComm &pc, unsigned long long phase,
                                               DevScal *S, double &s0, double &s1,
                                               double *s2 = nullptr) {
    __shared__ double sh[4];
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const int slot = (int)(phase % PEER_SLOTS);
        const unsigned int tag = (unsigned int)phase;
        double a = 0.0, b = 0.0, c = 0.0;
        bool ok = true;
        if (lane < pc.nranks) {
            const unsigned long long *w = reinterpret_cast<const unsigned long long *>(pc.slots) +
                                          ((size_t)slot * PEER_MAX_RANKS + lane) * PEER_VALS;
            unsigned long long w0, w1, w2, w3, w4, w5;
            const long long t0 = clock64();
            for (;;) {
                w0 = ld_relaxed_sys_u64(w); w1 = ld_relaxed_sys_u64(w + 1);
                w2 = ld_relaxed_sys_u64(w + 2); w3 = ld_relaxed_sys_u64(w + 3);
                w4 = ld_relaxed_sys_u64(w + 4); w5 = ld_relaxed_sys_u64(w + 5);
                if ((unsigned int)w0 == tag && (unsigned int)w1 == tag && (unsigned int)w2 == tag &&
                    (unsigned int)w3 == tag && (unsigned int)w4 == tag && (unsigned int)w5 == tag)
                    break;
                if (clock64() - t0 > pc.timeout_clocks) { ok = false; break; }
            }
            // acquire without a MEMBAR.SYS (which costs microseconds in every CTA): one more
            // look at a word that is already there, as an acquire load
            (void)ld_acquire_sys(w);
            a = __hiloint2double((int)(w1 >> 32), (int)(w0 >> 32));
            b = __hiloint2double((int)(w3 >> 32), (int)(w2 >> 32));
            c = __hiloint2double((int)(w5 >> 32), (int)(w4 >> 32));
            if (!ok) { a = 0.0; b = 0.0; c = 0.0; }
        }
        __syncwarp();
        static_assert(PEER_MAX_RANKS == 8, "the tree below spans eight lanes");
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            a += __shfl_xor_sync(FULL, a, o);
            b += __shfl_xor_sync(FULL, b, o);
            c += __shfl_xor_sync(FULL, c, o);
        }
        ok = __all_sync(FULL, ok);
        if (lane == 0) {
            if (!ok) S->error = 1;
            sh[0] = a; sh[1] = b; sh[2] = c; sh[3] = ok ? 1.0 : 0.0;
        }
    }
    __syncthreads();
    s0 = sh[0]; s1 = sh[1];
    if (s2) *s2 = sh[2];
    const bool all_ok = sh[3] != 0.0;
    __syncthreads();
    return all_ok;
}

// One thread: store this rank's sums, as LL words, into every rank's comm
// block (its own included).  No fence here: whoever calls this has already
// taken a system-scope fence after observing that all CTAs of this rank are
// done (ticket / arrival counter), and the CTAs that stored ghost rows into a
// neighbour GPU fenced system-wide before they signalled -- those rows are
// complete at their destination before these words are even sent.
__device__ __forceinline__ void peer_post(const PeerComm &pc, unsigned long long phase, double a,
                                          double b, double c = 0.0) {
    const int slot = (int)(phase % PEER_SLOTS);
    const unsigned long long w0 = ll_word((unsigned int)__double2loint(a), phase);
    const unsigned long long w1 = ll_word((unsigned int)__double2hiint(a), phase);
    const unsigned long long w2 = ll_word((unsigned int)__double2loint(b), phase);
    const unsigned long long w3 = ll_word((unsigned int)__double2hiint(b), phase);
    const unsigned long long w4 = ll_word((unsigned int)__double2loint(c), phase);
    const unsigned long long w5 = ll_word((unsigned int)__double2hiint(c), phase);
    for (int r = 0; r < pc.nranks; ++r) {
        unsigned long long *w = reinterpret_cast<unsigned long long *>(pc.peer_slots[r]) +
                                ((size_t)slot * PEER_MAX_RANKS + pc.rank) * PEER_VALS;
        st_relaxed_sys(w, w0); st_relaxed_sys(w + 1, w1); st_relaxed_sys(w + 2, w2);
        st_relaxed_sys(w + 3, w3); st_relaxed_sys(w + 4, w4); st_relaxed_sys(w + 5, w5);
    }
}

// The same from a whole warp that holds the sums in every lane: lane r stores into rank r's
// comm block, so the ranks' six words leave side by side instead of one rank after the other.
__device__ __forceinline__ void peer_post_lanes(const PeerComm &pc, unsigned long long phase,
                                                double a, double b, double c) {
    const int r = threadIdx.x & 31;
    if (r >= pc.nranks) return;
    const int slot = (int)(phase % PEER_SLOTS);
    unsigned long long *w = reinterpret_cast<unsigned long long *>(pc.peer_slots[r]) +
                            ((size_t)slot * PEER_MAX_RANKS + pc.rank) * PEER_VALS;
    st_relaxed_sys(w, ll_word((unsigned int)__double2loint(a), phase));
    st_relaxed_sys(w + 1, ll_word((unsigned int)__double2hiint(a), phase));
    st_relaxed_sys(w + 2, ll_word((unsigned int)__double2loint(b), phase));
    st_relaxed_sys(w + 3, ll_word((unsigned int)__double2hiint(b), phase));
    st_relaxed_sys(w + 4, ll_word((unsigned int)__double2loint(c), phase));
    st_relaxed_sys(w + 5, ll_word((unsigned int)__double2hiint(c), phase));
}

// ---------------------------------------------------------------- D(M)

// One IEEE divide and two integer powers per cell: a long dependent chain.  Every thread keeps
// UNROLL double2 (eight cells at 4) in flight so that the chains overlap; measured 0.111 ms at
// 4096^2 with one double2 per thread and iteration (fp64 pipe 39 % busy, latency bound).
template <int UNROLL>
__global__ void __launch_bounds__(256) k_diffcoef(GridDesc g, NumParams np,
                                                  const double *__restrict__ M,
                                                  double *__restrict__ d, long long n2) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n2; i += UNROLL * stride) {
        double2 m[UNROLL], o[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) m[u] = ld2(M + 2 * (i + u * stride));
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) { o[u].x = dfunc(m[u].x, np); o[u].y = dfunc(m[u].y, np); }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) st2(d + 2 * (i + u * stride), o[u]);
    }
    for (; i < n2; i += stride) {
        const double2 m = ld2(M + 2 * i);
        double2 o;
        o.x = dfunc(m.x, np);
        o.y = dfunc(m.y, np);
        st2(d + 2 * i, o);
    }
}

// ---------------------------------------------------------------- 5-point marching kernel

enum { MODE_INIT = 0, MODE_CG = 1 };

// Value of the vector the operator is applied to, at one cell.
//   INIT: x0                                  (CG:46)
//   CG:   r            first iteration        (CG:64)
//         r + beta p   afterwards             (CG:71)
template <int MODE>
struct VecLoad {
    const double *a, *b;
    double beta;
    bool first;
    __device__ __forceinline__ double2 two(long long off) const {
        if (MODE == MODE_INIT) return ld2(a + off);
        const double2 r = ld2(a + off);
        if (first) return r;
        const double2 p = ld2(b + off);
        double2 v;
        v.x = r.x + beta * p.x;
        v.y = r.y + beta * p.y;
        return v;
    }
    __device__ __forceinline__ double one(long long off) const {
        if (MODE == MODE_INIT) return a[off];
        const double r = a[off];
        if (first) return r;
        return r + beta * b[off];
    }
};

// Array pointers of one tile march: the members of StencilArgs that change
// from solve to solve.
struct TilePtrs {
    const double *d, *Cc, *Mprev, *x0;
    double *ac, *r, *x, *p, *q;
};

// One tile march: rows [i_begin, i_end) of the column pair (j0, j0+1).  Thread
// t of a warp owns the pair next to thread t-1's and keeps the rows above / at /
// below in registers.  West / east neighbours come from the adjacent lanes;
// the two edge lanes of each warp fetch theirs from memory.  All 32 lanes of a
// warp must call this together (shuffles).
template <int MODE>
__device__ __forceinline__ void stencil_tile(const StencilArgs &a, const TilePtrs &T,
                                             const VecLoad<MODE> &V, const int j0,
                                             const int i_begin, const int i_end,
                                             double &acc0, double &acc1) {
    const GridDesc &g = a.g;
    const NumParams &np = a.np;
    const int lane = threadIdx.x & 31;
    const bool inrow = j0 < g.P;
    const double h = np.h;
    const double2 zero2 = make_double2(0.0, 0.0);

    // column roles (P:560, P:568)
    const bool firstc0 = (j0 == 0);
    const bool lastc0 = (j0 == g.col - 1), lastc1 = (j0 + 1 == g.col - 1);
    const bool valid0 = j0 < g.col, valid1 = j0 + 1 < g.col;

    long long off = (long long)i_begin * g.P + j0;
    double2 vN = zero2, vC = zero2, dN = zero2, dC = zero2;
    if (inrow) {
        vN = V.two(off - g.P);
        vC = V.two(off);
        dN = ld2(T.d + off - g.P);
        dC = ld2(T.d + off);
    }
    if (MODE == MODE_CG && a.write_ghost_top && i_begin == 0 && inrow)
        st2(T.p + off - g.P, vN);

    for (int i = i_begin; i < i_end; ++i, off += g.P) {
        double2 vS = zero2, dS = zero2;
        if (inrow) {
            vS = V.two(off + g.P);
            dS = ld2(T.d + off + g.P);
        }
        double vW0 = __shfl_up_sync(FULL, vC.y, 1);
        double dW0 = __shfl_up_sync(FULL, dC.y, 1);
        double vE1 = __shfl_down_sync(FULL, vC.x, 1);
        double dE1 = __shfl_down_sync(FULL, dC.x, 1);
        if (lane == 0 && inrow) { vW0 = V.one(off - 1); dW0 = T.d[off - 1]; }
        if (lane == 31 && inrow) { vE1 = V.one(off + 2); dE1 = T.d[off + 2]; }

        if (inrow) {
            const int gi = g.row0 + i;
            const bool first = (gi == 0);                               // P:552
            const bool lastrow = (gi == g.grow - 1);
            const bool quirkrow = (gi == g.grow - 2);                   // P:576 '>='
            // face coefficients xCof*0.5*(diff(k)+diff(i))  (P:553)
            const double aN0 = h * (dN.x + dC.x), aN1 = h * (dN.y + dC.y);
            const double aS0 = h * (dS.x + dC.x), aS1 = h * (dS.y + dC.y);
            const double aW0 = h * (dW0 + dC.x);
            const double aM = h * (dC.y + dC.x);   // east of cell 0 == west of cell 1
            const double aE1 = h * (dE1 + dC.y);

            double2 ac2, q2;
            double2 c2 = zero2, mp2 = zero2;
            if (MODE == MODE_INIT) {
                c2 = ld2(T.Cc + off);
                mp2 = ld2(T.Mprev + off);
            } else {
                ac2 = ld2(T.ac + off);
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const bool firstc = c ? false : firstc0;
                const bool lastc = c ? lastc1 : lastc0;
                const bool lastq = lastrow || (quirkrow && lastc);
                const double aN = c ? aN1 : aN0, aS = c ? aS1 : aS0;
                const double aW = c ? aM : aW0, aE = c ? aE1 : aM;
                const double xN = c ? vN.y : vN.x, xS = c ? vS.y : vS.x;
                const double xC = c ? vC.y : vC.x;
                const double xW = c ? vC.x : vW0, xE = c ? vE1 : vC.y;
                // DIA entries after the four if/else blocks of P:552-582
                const double A1 = first ? 0.0 : (lastq ? -(aN + aN) : -aN);
                const double A5 = lastq ? 0.0 : (first ? -(aS + aS) : -aS);
                const double A2 = firstc ? 0.0 : (lastc ? -(aW + aW) : -aW);
                const double A4 = lastc ? 0.0 : (firstc ? -(aE + aE) : -aE);
                double A3;
                if (MODE == MODE_INIT) {
                    // centre: blocks 1..4 in order, then - f + (1/tDel)  (P:554-585)
                    const double t1 = first ? aS : aN;
                    const double t2 = firstc ? aE : aW;
                    const double t3 = lastc ? aW : aE;
                    const double t4 = lastq ? aN : aS;
                    const double f = ffunc(c ? mp2.y : mp2.x, c ? c2.y : c2.x, np);
                    A3 = (((t1 + t2) + t3) + t4) - f + np.inv_dt;
                    if (c) ac2.y = A3; else ac2.x = A3;
                } else {
                    A3 = c ? ac2.y : ac2.x;
                }
                // amuxd: diagonals 1..5 accumulated in order (P:1093-1100)
                double y = A1 * xN;
                y = y + A2 * xW;
                y = y + A3 * xC;
                y = y + A4 * xE;
                y = y + A5 * xS;
                const bool valid = c ? valid1 : valid0;
                if (MODE == MODE_INIT) {
                    const double rhs = (c ? mp2.y : mp2.x) / np.tDel;       // P:586
                    const double rr = valid ? (rhs - y) : 0.0;              // CG:49
                    if (c) q2.y = rr; else q2.x = rr;
                    acc0 += rr * rr;
                    acc1 += valid ? xC * xC : 0.0;
                    if (!valid) { if (c) ac2.y = 0.0; else ac2.x = 0.0; }
                } else {
                    const double qq = valid ? y : 0.0;
                    if (c) q2.y = qq; else q2.x = qq;
                    acc0 += xC * qq;                                        // CG:76
                    acc1 += valid ? xC * xC : 0.0;                          // CG:87
                }
            }
            if (MODE == MODE_INIT) {
                st2(T.ac + off, ac2);
                st2(T.r + off, q2);
                if (T.x != T.x0) st2(T.x + off, vC);
                if (a.fuse_scalars == SUMS_PEER) {
                    // my boundary rows of r are the neighbours' ghost rows
                    if (i == 0 && a.pc.r_up)
                        st2(a.pc.r_up + (long long)a.pc.rows_up * g.P + j0, q2);
                    if (i == g.rows - 1 && a.pc.r_dn) st2(a.pc.r_dn - g.P + j0, q2);
                }
            } else {
                st2(T.p + off, vC);
                st2(T.q + off, q2);
            }
        }
        vN = vC; vC = vS; dN = dC; dC = dS;
    }
    if (MODE == MODE_CG && a.write_ghost_bot && i_end == g.rows && inrow)
        st2(T.p + off, vC);
}

// One CTA marches TR rows down a strip of 2*BS columns.
template <int MODE, int BS>
__global__ void __launch_bounds__(BS) k_stencil(const StencilArgs a) {
    __shared__ double sm[64];
    DevScal *S = a.S;
    VecLoad<MODE> V;
    if (MODE == MODE_CG) {
        if (S->done) return;
        V.a = a.r; V.b = a.p_in;
        V.first = (S->nit == 0);
        V.beta = V.first ? 0.0 : S->rho / S->rho_old;       // CG:68
    } else {
        V.a = a.x0; V.b = nullptr; V.first = true; V.beta = 0.0;
    }
    const GridDesc g = a.g;
    const NumParams np = a.np;
    const int j0 = (blockIdx.x * BS + threadIdx.x) * 2;
    const int i_begin = blockIdx.y * a.TR;
    const int i_end = min(i_begin + a.TR, g.rows);
    TilePtrs T;
    T.d = a.d; T.Cc = a.Cc; T.Mprev = a.Mprev; T.x0 = a.x0;
    T.ac = a.ac; T.r = a.r; T.x = a.x; T.p = a.p; T.q = a.q;
    double acc0 = 0.0, acc1 = 0.0;
    stencil_tile<MODE>(a, T, V, j0, i_begin, i_end, acc0, acc1);

    const int nb = gridDim.x * gridDim.y;
    const int bid = blockIdx.y * gridDim.x + blockIdx.x;
    if (grid_sum2(acc0, acc1, a.partials, nb, bid, &S->ticket[TICKET_STENCIL], sm,
                  a.fuse_scalars == SUMS_PEER)) {
        if (a.fuse_scalars == SUMS_FUSED) {
            if (MODE == MODE_INIT) scal_after_init(S, acc0, acc1);
            else scal_after_spmv(S, acc0, acc1, g.n_glob, np.tol2, a.trace);
        } else if (a.fuse_scalars == SUMS_PEER && MODE == MODE_INIT) {
            // rho and sum(sol*sol) become known to every rank once all have posted; the
            // first K_A adds the slots.  Local part of CG:53 happens here.
            scal_after_init(S, 0.0, 0.0);
            peer_post(a.pc, a.post_phase, acc0, acc1);
        } else {
            S->part[0] = acc0; S->part[1] = acc1;
        }
    }
}

// ---------------------------------------------------------------- CG SpMV, bulk-async staged

// mbarrier / bulk-copy (TMA engine, 1-D form) primitives
__device__ __forceinline__ unsigned smem_u32(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes,
                                         unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Hint: bring [src, src + bytes) into L2 (no destination, no completion to wait for).
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// Same arithmetic as k_stencil<MODE_CG>, different data movement: one producer
// lane streams the row segments of r, p, d and ac of a 2*BS-column strip into a
// DEPTH-deep shared-memory ring with cp.async.bulk (the TMA engine), each row
// arriving on its own mbarrier; BS consumer threads pick a landed row out of
// shared memory, keep three rows in registers and march down.  Loads are in
// flight DEPTH rows ahead of the arithmetic without costing registers, which
// is what the register-staged version lacked (long-scoreboard bound at 35 %
// occupancy).  Row segments carry a 2-double halo on both sides so the west /
// east neighbours of the strip's edge columns come from the same copy.
template <int BS, int DEPTH>
struct Ring {
    static constexpr int TC = 2 * BS;          // columns per strip
    static constexpr int SEG = TC + 4;         // doubles per staged row segment (halo 2 + 2)
    static constexpr size_t BYTES = (size_t)DEPTH * (3 * SEG + 2 * TC) * sizeof(double) + 2 * DEPTH * 8;
    double *s_r, *s_p, *s_d, *s_ac, *s_x;
    unsigned long long *full, *empty;
    __device__ __forceinline__ void carve(unsigned char *raw) {
        s_r = reinterpret_cast<double *>(raw);                     // [DEPTH][SEG]
        s_p = s_r + DEPTH * SEG;                                   // [DEPTH][SEG]
        s_d = s_p + DEPTH * SEG;                                   // [DEPTH][SEG]
        s_ac = s_d + DEPTH * SEG;                                  // [DEPTH][TC]
        s_x = s_ac + DEPTH * TC;                                   // [DEPTH][TC] (88-byte iteration)
        full = reinterpret_cast<unsigned long long *>(s_x + DEPTH * TC);
        empty = full + DEPTH;
    }
    __device__ __forceinline__ void init_barriers() {
        for (int s = 0; s < DEPTH; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], BS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
};

// Producer lane: rows i_begin-1 .. i_end of one strip into the ring.  K counts
// the rows this CTA has ever pushed through the ring (slot = K % DEPTH, use
// number = K / DEPTH gives the mbarrier parity), so tiles and CG iterations
// can follow one another without re-initialising the barriers.
template <int BS, int DEPTH>
__device__ __forceinline__ void bulk_produce(const Ring<BS, DEPTH> &R, const GridDesc &g,
                                             const double *r, const double *p_in,
                                             const double *d, const double *ac, bool firstit,
                                             int jt, int i_begin, int i_end, long long K,
                                             const double *xsrc = nullptr) {
    constexpr int TC = Ring<BS, DEPTH>::TC, SEG = Ring<BS, DEPTH>::SEG;
    const int nstage = (i_end - i_begin) + 2;
    // segment [jt-2, jt-2+seglen): never past 2 doubles beyond the row (guard)
    const unsigned segbytes = (unsigned)min(SEG, g.P + 4 - jt) * 8u;
    const unsigned acbytes = (unsigned)min(TC, g.P - jt) * 8u;
    for (int k = 0; k < nstage; ++k, ++K) {
        const int s = (int)(K % DEPTH);
        if (K >= DEPTH) mbar_wait(&R.empty[s], (unsigned)((K / DEPTH) - 1) & 1u);
        const int i = i_begin - 1 + k;
        const long long off = (long long)i * g.P + (jt - 2);
        const bool has_ac = (k >= 1 && k <= nstage - 2);
        const bool has_x = has_ac && xsrc != nullptr;
        const unsigned tx = segbytes * (firstit ? 2u : 3u) + (has_ac ? acbytes : 0u) +
                            (has_x ? acbytes : 0u);
        mbar_expect_tx(&R.full[s], tx);
        bulk_g2s(R.s_r + s * SEG, r + off, segbytes, &R.full[s]);
        if (!firstit) bulk_g2s(R.s_p + s * SEG, p_in + off, segbytes, &R.full[s]);
        bulk_g2s(R.s_d + s * SEG, d + off, segbytes, &R.full[s]);
        if (has_ac) bulk_g2s(R.s_ac + s * TC, ac + off + 2, acbytes, &R.full[s]);
        if (has_x) bulk_g2s(R.s_x + s * TC, xsrc + off + 2, acbytes, &R.full[s]);
    }
}

// Consumers: p = r + beta p ; q = A p on rows [i_begin, i_end) of one strip;
// adds p.q and p.p of those rows to acc0 / acc1.
template <int BS, int DEPTH, bool DEFER = false>
__device__ __forceinline__ void bulk_consume(const Ring<BS, DEPTH> &R, const GridDesc &g,
                                             double h, double *p_out, double *q_out,
                                             bool firstit, double beta, int jt, int i_begin,
                                             int i_end, long long K, bool ghost_top,
                                             bool ghost_bot, double &acc0, double &acc1,
                                             double *xdef = nullptr, double alfa_prev = 0.0,
                                             double *acc2 = nullptr) {
    constexpr int TC = Ring<BS, DEPTH>::TC, SEG = Ring<BS, DEPTH>::SEG;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int j0 = jt + 2 * tid;
    const bool inrow = j0 < g.P;
    const bool firstc0 = (j0 == 0);
    const bool lastc0 = (j0 == g.col - 1), lastc1 = (j0 + 1 == g.col - 1);
    const bool valid0 = j0 < g.col, valid1 = j0 + 1 < g.col;
    const int c0 = 2 * tid + 2;                  // my pair inside a staged segment
    const int ce = (lane == 0) ? c0 - 1 : c0 + 2;  // edge lanes: west / east halo cell

    double2 vN, vC, vS, dN, dC, dS, acC, acS;
    double veC = 0.0, deC = 0.0, veS = 0.0, deS = 0.0;   // edge-lane neighbours
    vN = vC = vS = dN = dC = dS = acC = acS = make_double2(0.0, 0.0);

    // fetch one staged row into (v, d, ac, ve, de)
    // deferred update (the 88-byte iteration): x += alfa_prev * p_prev rides along here,
    // on the cells this thread owns, using the p_prev it reads anyway
    const bool defer = DEFER && (xdef != nullptr) && !firstit;
    double2 pC = make_double2(0.0, 0.0), pS = pC, pdum = pC;
    double2 xC = make_double2(0.0, 0.0), xS = xC, xdum = xC;
    auto take = [&](long long kk, double2 &v, double2 &d, double2 &ac, double &ve, double &de,
                    double2 &pv, double2 &xv) {
        const int s = (int)(kk % DEPTH);
        mbar_wait(&R.full[s], (unsigned)(kk / DEPTH) & 1u);
        const double *sr = R.s_r + s * SEG, *sp = R.s_p + s * SEG, *sd = R.s_d + s * SEG;
        if (inrow) {
            const double2 r2 = *reinterpret_cast<const double2 *>(sr + c0);
            d = *reinterpret_cast<const double2 *>(sd + c0);
            ac = *reinterpret_cast<const double2 *>(R.s_ac + s * TC + 2 * tid);
            if (defer) xv = *reinterpret_cast<const double2 *>(R.s_x + s * TC + 2 * tid);
            if (firstit) {
                v = r2;                                            // CG:64
            } else {
                const double2 p2 = *reinterpret_cast<const double2 *>(sp + c0);
                v.x = r2.x + beta * p2.x;                          // CG:71
                v.y = r2.y + beta * p2.y;
                pv = p2;
            }
            if (lane == 0 || lane == 31) {
                ve = firstit ? sr[ce] : sr[ce] + beta * sp[ce];
                de = sd[ce];
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&R.empty[s]);
    };

    double2 dummy;
    double dum1, dum2;
    take(K, vN, dN, dummy, dum1, dum2, pdum, xdum);
    take(K + 1, vC, dC, acC, veC, deC, pC, xC);
    long long off = (long long)i_begin * g.P + j0;
    if (ghost_top && i_begin == 0 && inrow) st2(p_out + off - g.P, vN);

    for (int i = i_begin; i < i_end; ++i, off += g.P) {
        take(K + (i - i_begin) + 2, vS, dS, acS, veS, deS, pS, xS);
        double2 x2 = xC;
        double vW0 = __shfl_up_sync(FULL, vC.y, 1);
        double dW0 = __shfl_up_sync(FULL, dC.y, 1);
        double vE1 = __shfl_down_sync(FULL, vC.x, 1);
        double dE1 = __shfl_down_sync(FULL, dC.x, 1);
        if (lane == 0) { vW0 = veC; dW0 = deC; }
        if (lane == 31) { vE1 = veC; dE1 = deC; }
        if (inrow) {
            const int gi = g.row0 + i;
            const bool first = (gi == 0);
            const bool lastrow = (gi == g.grow - 1);
            const bool quirkrow = (gi == g.grow - 2);
            const double aN0 = h * (dN.x + dC.x), aN1 = h * (dN.y + dC.y);
            const double aS0 = h * (dS.x + dC.x), aS1 = h * (dS.y + dC.y);
            const double aW0 = h * (dW0 + dC.x);
            const double aM = h * (dC.y + dC.x);
            const double aE1 = h * (dE1 + dC.y);
            double2 q2;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const bool firstc = c ? false : firstc0;
                const bool lastc = c ? lastc1 : lastc0;
                const bool lastq = lastrow || (quirkrow && lastc);
                const double aN = c ? aN1 : aN0, aS = c ? aS1 : aS0;
                const double aW = c ? aM : aW0, aE = c ? aE1 : aM;
                const double xN = c ? vN.y : vN.x, xS = c ? vS.y : vS.x;
                const double xC = c ? vC.y : vC.x;
                const double xW = c ? vC.x : vW0, xE = c ? vE1 : vC.y;
                const double A1 = first ? 0.0 : (lastq ? -(aN + aN) : -aN);
                const double A5 = lastq ? 0.0 : (first ? -(aS + aS) : -aS);
                const double A2 = firstc ? 0.0 : (lastc ? -(aW + aW) : -aW);
                const double A4 = lastc ? 0.0 : (firstc ? -(aE + aE) : -aE);
                const double A3 = c ? acC.y : acC.x;
                double y = A1 * xN;
                y = y + A2 * xW;
                y = y + A3 * xC;
                y = y + A4 * xE;
                y = y + A5 * xS;
                const bool valid = c ? valid1 : valid0;
                const double qq = valid ? y : 0.0;
                if (c) q2.y = qq; else q2.x = qq;
                acc0 += xC * qq;
                acc1 += valid ? xC * xC : 0.0;
            }
            st2(p_out + off, vC);
            st2(q_out + off, q2);
            if (defer) {
                x2.x = x2.x + alfa_prev * pC.x;                    // CG:80 of the previous iteration
                x2.y = x2.y + alfa_prev * pC.y;
                st2(xdef + off, x2);
                *acc2 += x2.x * x2.x;                              // CG:57 of this iteration
                *acc2 += x2.y * x2.y;
            }
        }
        vN = vC; vC = vS; dN = dC; dC = dS; acC = acS; veC = veS; deC = deS; pC = pS; xC = xS;
    }
    if (ghost_bot && i_end == g.rows && inrow) st2(p_out + off, vC);
}

// DEFER = the 88-byte iteration (x += alfa p of the previous iteration rides along); kept out
// of the plain instantiation so that it does not pay the registers.
template <int BS, int DEPTH, bool DEFER>
__global__ void __launch_bounds__(BS + 32, (BS == 256 ? 2 : (BS == 128 ? 3 : 6)))
k_spmv_bulk(const StencilArgs a) {
    constexpr int TC = Ring<BS, DEPTH>::TC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double sm[64];
    Ring<BS, DEPTH> R;
    R.carve(smem_raw);

    DevScal *S = a.S;
    if (S->done) return;
    const bool firstit = (S->nit == 0);
    double beta, rr_glob = 0.0, xx_glob = 0.0;
    if (a.fuse_scalars == SUMS_PEER) {
        // dotProd(r,r) and sum(sol*sol) of the previous phase arrive from all ranks;
        // the neighbours' flags also cover the ghost rows of r they wrote here
        if (!peer_wait_sums(a.pc, a.wait_phase, S, rr_glob, xx_glob)) {
            if (threadIdx.x == 0) { S->done = 1; S->finished = 1; }   // error: nothing is stored
            return;
        }
        beta = firstit ? 0.0 : rr_glob / S->rho;                   // S->rho still holds rho_old
    } else {
        beta = firstit ? 0.0 : S->rho / S->rho_old;                // CG:68
    }
    const GridDesc g = a.g;
    const int tid = threadIdx.x;
    const int jt = blockIdx.x * TC;                 // first column of the strip
    const int i_begin = blockIdx.y * a.TR;
    const int i_end = min(i_begin + a.TR, g.rows);

    if (tid == 0) R.init_barriers();
    __syncthreads();

    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    const double alfa_prev = S->alfa;               // of the previous iteration (deferred x update)
    if (tid >= BS) {
        if (tid == BS)
            bulk_produce<BS, DEPTH>(R, g, a.r, a.p_in, a.d, a.ac, firstit, jt, i_begin, i_end, 0,
                                    (DEFER && !firstit) ? a.x : nullptr);
    } else {
        bulk_consume<BS, DEPTH, DEFER>(R, g, a.np.h, a.p, a.q, firstit, beta, jt, i_begin, i_end, 0,
                                       a.write_ghost_top != 0, a.write_ghost_bot != 0, acc0, acc1,
                                       DEFER ? a.x : nullptr, alfa_prev, &acc2);
    }

    const int nb = gridDim.x * gridDim.y;
    const int bid = blockIdx.y * gridDim.x + blockIdx.x;
    if (DEFER) {
        // 88-byte iteration: sum(sol*sol) of the iterate this kernel has just completed
        // comes out of the same pass as p.q and p.p
        if (grid_sum3(acc0, acc1, acc2, a.partials, nb, bid, &S->ticket[TICKET_STENCIL], sm,
                      a.fuse_scalars == SUMS_PEER)) {
            if (a.fuse_scalars == SUMS_FUSED) {
                if (!firstit) S->xx = acc2;                          // CG:57
                scal_after_spmv(S, acc0, acc1, g.n_glob, a.np.tol2, a.trace);
            } else if (a.fuse_scalars == SUMS_PEER) {
                // S->xx is set by the update kernel from the all-rank sum posted here
                S->rho_old = S->rho; S->rho = rr_glob;
                if (firstit) S->xx = xx_glob;                        // sum(x0*x0), from INIT
                peer_post(a.pc, a.post_phase, acc0, acc1, acc2);
            } else {
                S->part[0] = acc0; S->part[1] = acc1; S->part[2] = acc2;
            }
        }
        return;
    }
    if (grid_sum2(acc0, acc1, a.partials, nb, bid, &S->ticket[TICKET_STENCIL], sm,
                  a.fuse_scalars == SUMS_PEER)) {
        if (a.fuse_scalars == SUMS_FUSED) {
            scal_after_spmv(S, acc0, acc1, g.n_glob, a.np.tol2, a.trace);
        } else if (a.fuse_scalars == SUMS_PEER) {
            // close the previous phase (CG:56,57,59) now that every CTA has used rho_old,
            // then hand this rank's p.q and p.p to all ranks
            S->rho_old = S->rho; S->rho = rr_glob; S->xx = xx_glob;
            peer_post(a.pc, a.post_phase, acc0, acc1);
        } else {
            S->part[0] = acc0; S->part[1] = acc1;
        }
    }
}

// ---------------------------------------------------------------- INIT, bulk-async staged

// k_stencil<MODE_INIT>'s arithmetic behind the same shared-memory ring as
// k_spmv_bulk: the ring's four lanes carry x0 (with halo), Mprev, d (with halo)
// and C instead of r, p, d, ac.  The register-staged INIT reached 0.58 of the
// measured HBM peak; this one streams like K_A.
template <int BS, int DEPTH>
__global__ void __launch_bounds__(BS + 32) k_init_bulk(const StencilArgs a) {
    constexpr int TC = Ring<BS, DEPTH>::TC, SEG = Ring<BS, DEPTH>::SEG;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double sm[64];
    Ring<BS, DEPTH> R;
    R.carve(smem_raw);
    DevScal *S = a.S;
    const GridDesc g = a.g;
    const NumParams np = a.np;
    const int tid = threadIdx.x;
    const int jt = blockIdx.x * TC;
    const int i_begin = blockIdx.y * a.TR;
    const int i_end = min(i_begin + a.TR, g.rows);
    if (tid == 0) R.init_barriers();
    __syncthreads();

    double acc0 = 0.0, acc1 = 0.0;
    if (tid >= BS) {
        // lanes: s_r <- x0, s_p <- Mprev, s_d <- d, s_ac <- C (rows i_begin..i_end-1 only)
        if (tid == BS)
            bulk_produce<BS, DEPTH>(R, g, a.x0, a.Mprev, a.d, a.Cc, false, jt, i_begin, i_end, 0);
    } else {
        const int lane = tid & 31;
        const int j0 = jt + 2 * tid;
        const bool inrow = j0 < g.P;
        const double h = np.h;
        const bool firstc0 = (j0 == 0);
        const bool lastc0 = (j0 == g.col - 1), lastc1 = (j0 + 1 == g.col - 1);
        const bool valid0 = j0 < g.col, valid1 = j0 + 1 < g.col;
        const int c0 = 2 * tid + 2;
        const int ce = (lane == 0) ? c0 - 1 : c0 + 2;
        double2 vN, vC, vS, dN, dC, dS, cC, cS, mC, mS;
        double veC = 0.0, deC = 0.0, veS = 0.0, deS = 0.0;
        vN = vC = vS = dN = dC = dS = cC = cS = mC = mS = make_double2(0.0, 0.0);
        auto take = [&](long long kk, double2 &v, double2 &d, double2 &cc, double2 &mp, double &ve,
                        double &de) {
            const int s = (int)(kk % DEPTH);
            mbar_wait(&R.full[s], (unsigned)(kk / DEPTH) & 1u);
            const double *sx = R.s_r + s * SEG, *sd = R.s_d + s * SEG;
            if (inrow) {
                v = *reinterpret_cast<const double2 *>(sx + c0);
                d = *reinterpret_cast<const double2 *>(sd + c0);
                cc = *reinterpret_cast<const double2 *>(R.s_ac + s * TC + 2 * tid);
                mp = *reinterpret_cast<const double2 *>(R.s_p + s * SEG + c0);
                if (lane == 0 || lane == 31) { ve = sx[ce]; de = sd[ce]; }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&R.empty[s]);
        };
        double2 dm1, dm2;
        double du1, du2;
        take(0, vN, dN, dm1, dm2, du1, du2);
        take(1, vC, dC, cC, mC, veC, deC);
        long long off = (long long)i_begin * g.P + j0;
        for (int i = i_begin; i < i_end; ++i, off += g.P) {
            take((long long)(i - i_begin) + 2, vS, dS, cS, mS, veS, deS);
            double vW0 = __shfl_up_sync(FULL, vC.y, 1);
            double dW0 = __shfl_up_sync(FULL, dC.y, 1);
            double vE1 = __shfl_down_sync(FULL, vC.x, 1);
            double dE1 = __shfl_down_sync(FULL, dC.x, 1);
            if (lane == 0) { vW0 = veC; dW0 = deC; }
            if (lane == 31) { vE1 = veC; dE1 = deC; }
            if (inrow) {
                const int gi = g.row0 + i;
                const bool first = (gi == 0);
                const bool lastrow = (gi == g.grow - 1);
                const bool quirkrow = (gi == g.grow - 2);
                const double aN0 = h * (dN.x + dC.x), aN1 = h * (dN.y + dC.y);
                const double aS0 = h * (dS.x + dC.x), aS1 = h * (dS.y + dC.y);
                const double aW0 = h * (dW0 + dC.x);
                const double aM = h * (dC.y + dC.x);
                const double aE1 = h * (dE1 + dC.y);
                double2 ac2, r2;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const bool firstc = c ? false : firstc0;
                    const bool lastc = c ? lastc1 : lastc0;
                    const bool lastq = lastrow || (quirkrow && lastc);
                    const double aN = c ? aN1 : aN0, aS = c ? aS1 : aS0;
                    const double aW = c ? aM : aW0, aE = c ? aE1 : aM;
                    const double xN = c ? vN.y : vN.x, xS = c ? vS.y : vS.x;
                    const double xC = c ? vC.y : vC.x;
                    const double xW = c ? vC.x : vW0, xE = c ? vE1 : vC.y;
                    const double A1 = first ? 0.0 : (lastq ? -(aN + aN) : -aN);
                    const double A5 = lastq ? 0.0 : (first ? -(aS + aS) : -aS);
                    const double A2 = firstc ? 0.0 : (lastc ? -(aW + aW) : -aW);
                    const double A4 = lastc ? 0.0 : (firstc ? -(aE + aE) : -aE);
                    // centre: blocks 1..4 in order, then - f + (1/tDel)  (P:554-585)
                    const double t1 = first ? aS : aN;
                    const double t2 = firstc ? aE : aW;
                    const double t3 = lastc ? aW : aE;
                    const double t4 = lastq ? aN : aS;
                    const double mpc = c ? mC.y : mC.x;
                    const double f = ffunc(mpc, c ? cC.y : cC.x, np);
                    const double A3 = (((t1 + t2) + t3) + t4) - f + np.inv_dt;
                    double y = A1 * xN;                                    // P:1093-1100
                    y = y + A2 * xW;
                    y = y + A3 * xC;
                    y = y + A4 * xE;
                    y = y + A5 * xS;
                    const bool valid = c ? valid1 : valid0;
                    const double rhs = mpc / np.tDel;                      // P:586
                    const double rr = valid ? (rhs - y) : 0.0;             // CG:49
                    if (c) { r2.y = rr; ac2.y = valid ? A3 : 0.0; } else { r2.x = rr; ac2.x = valid ? A3 : 0.0; }
                    acc0 += rr * rr;
                    acc1 += valid ? xC * xC : 0.0;
                }
                st2(a.ac + off, ac2);
                st2(a.r + off, r2);
                if (a.x != a.x0) st2(a.x + off, vC);
                if (a.fuse_scalars == SUMS_PEER) {
                    if (i == 0 && a.pc.r_up) st2(a.pc.r_up + (long long)a.pc.rows_up * g.P + j0, r2);
                    if (i == g.rows - 1 && a.pc.r_dn) st2(a.pc.r_dn - g.P + j0, r2);
                }
            }
            vN = vC; vC = vS; dN = dC; dC = dS; cC = cS; mC = mS; veC = veS; deC = deS;
        }
    }
    const int nb = gridDim.x * gridDim.y;
    const int bid = blockIdx.y * gridDim.x + blockIdx.x;
    if (grid_sum2(acc0, acc1, a.partials, nb, bid, &S->ticket[TICKET_STENCIL], sm,
                  a.fuse_scalars == SUMS_PEER)) {
        if (a.fuse_scalars == SUMS_FUSED) {
            scal_after_init(S, acc0, acc1);
        } else if (a.fuse_scalars == SUMS_PEER) {
            scal_after_init(S, 0.0, 0.0);
            peer_post(a.pc, a.post_phase, acc0, acc1);
        } else {
            S->part[0] = acc0; S->part[1] = acc1;
        }
    }
}

// ---------------------------------------------------------------- CG vector update

// sol += alfa p ; r -= alfa q (CG:79-82) and the two sums the next iteration
// opens with: dotProd(r,r) (CG:59) and sum(sol*sol) (CG:57).  Pad cells hold
// zeros in all four vectors, so the arrays are streamed flat.
__global__ void __launch_bounds__(256) k_update(const UpdateArgs a) {
    __shared__ double sm[64];
    DevScal *S = a.S;
    if (S->finished) return;
    if (a.tail_ok && S->done) {
        // the last iteration (CG:88-90 said stop): r -= alfa q is dead, the two sums are dead,
        // and sol += alfa p is applied by k_solvec, which reads sol anyway
        if (blockIdx.x == 0 && threadIdx.x == 0) { S->tail = 1; S->finished = 1; }
        return;
    }
    const bool peer = (a.fuse_scalars == SUMS_PEER);
    const bool r_only = a.r_only != 0;          // 88-byte iteration: x belongs to the SpMV kernel
    double alfa, pq_glob = 0.0, pp_glob = 0.0, xx_post = 0.0;
    if (peer) {
        if (!peer_wait_sums(a.pc, a.wait_phase, S, pq_glob, pp_glob, &xx_post)) {   // dotProd(p,q), sum(p*p)
            if (threadIdx.x == 0) { S->done = 1; S->finished = 1; }   // error: nothing is stored
            return;
        }
        alfa = S->rho / pq_glob;                                   // CG:77
    } else {
        alfa = S->alfa;
    }
    // peer mode: rows 0 and rows-1 of r are also the neighbours' ghost rows
    const long long last_row0 = (long long)(a.g.rows - 1) * a.g.P;
    double *const up = peer && a.pc.r_up ? a.pc.r_up + (long long)a.pc.rows_up * a.g.P : nullptr;
    double *const dn = peer && a.pc.r_dn ? a.pc.r_dn - a.g.P - last_row0 : nullptr;
    auto put_r = [&](long long k, double2 v) {
        st2(a.r + k, v);
        if (up && k < a.g.P) st2(up + k, v);
        if (dn && k >= last_row0) st2(dn + k, v);
    };
    double rr = 0.0, xx = 0.0;
    const bool rev = a.reverse != 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < a.n2; i += 2 * stride) {
        // reverse: sweep from the high addresses down, i.e. start where the SpMV kernel
        // just finished, so that the tails of p, q, r it left in L2 are hit
        const long long k0 = 2 * (rev ? a.n2 - 1 - i : i);
        const long long k1 = 2 * (rev ? a.n2 - 1 - (i + stride) : i + stride);
        double2 r0 = ld2(a.r + k0), q0 = ld2(a.q + k0), r1 = ld2(a.r + k1), q1 = ld2(a.q + k1);
        double2 x0 = r0, p0 = r0, x1 = r1, p1 = r1;
        if (!r_only) { x0 = ld2(a.x + k0); p0 = ld2(a.p + k0); x1 = ld2(a.x + k1); p1 = ld2(a.p + k1); }
        r0.x = r0.x - alfa * q0.x; r0.y = r0.y - alfa * q0.y;
        r1.x = r1.x - alfa * q1.x; r1.y = r1.y - alfa * q1.y;
        put_r(k0, r0); put_r(k1, r1);
        if (!r_only) {
            x0.x = x0.x + alfa * p0.x; x0.y = x0.y + alfa * p0.y;
            x1.x = x1.x + alfa * p1.x; x1.y = x1.y + alfa * p1.y;
            st2(a.x + k0, x0); st2(a.x + k1, x1);
        }
        rr += r0.x * r0.x; rr += r0.y * r0.y; rr += r1.x * r1.x; rr += r1.y * r1.y;
        xx += x0.x * x0.x; xx += x0.y * x0.y; xx += x1.x * x1.x; xx += x1.y * x1.y;
    }
    for (; i < a.n2; i += stride) {
        const long long k0 = 2 * (rev ? a.n2 - 1 - i : i);
        double2 r0 = ld2(a.r + k0), q0 = ld2(a.q + k0);
        double2 x0 = r0, p0 = r0;
        if (!r_only) { x0 = ld2(a.x + k0); p0 = ld2(a.p + k0); }
        r0.x = r0.x - alfa * q0.x; r0.y = r0.y - alfa * q0.y;
        put_r(k0, r0);
        if (!r_only) {
            x0.x = x0.x + alfa * p0.x; x0.y = x0.y + alfa * p0.y;
            st2(a.x + k0, x0);
        }
        rr += r0.x * r0.x; rr += r0.y * r0.y;
        xx += x0.x * x0.x; xx += x0.y * x0.y;
    }
    if (grid_sum2(rr, xx, a.partials, gridDim.x, blockIdx.x, &S->ticket[TICKET_UPDATE], sm, peer)) {
        if (a.fuse_scalars == SUMS_FUSED) {
            scal_after_update(S, rr, xx);
        } else if (peer) {
            // the stop test of this iteration (CG:55,77,87-90) with the all-rank sums, then
            // this rank's dotProd(r,r) and sum(sol*sol) go to all ranks
            if (r_only && S->nit > 0) S->xx = xx_post;   // sum(sol*sol) came with p.q and p.p
            scal_after_spmv(S, pq_glob, pp_glob, a.g.n_glob, a.np.tol2, a.trace);
            S->finished = S->done;
            peer_post(a.pc, a.post_phase, rr, xx);
        } else {
            S->part[0] = rr; S->part[1] = xx;
        }
    }
}

// The update kernel of the 88-byte iteration: r -= alfa q and dotProd(r,r) only
// (CG:81, CG:59); sol += alfa p (CG:80) is applied by the next SpMV kernel, which
// reads p anyway, and by k_xtail after the last iteration.
__global__ void __launch_bounds__(256) k_update_r(const UpdateArgs a) {
    __shared__ double sm[64];
    DevScal *S = a.S;
    if (S->finished) return;
    if (a.tail_ok && S->done) {                 // see k_update
        if (blockIdx.x == 0 && threadIdx.x == 0) { S->tail = 1; S->finished = 1; }
        return;
    }
    const double alfa = S->alfa;
    const bool rev = a.reverse != 0;
    double rr = 0.0, zz = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < a.n2; i += 2 * stride) {
        const long long k0 = 2 * (rev ? a.n2 - 1 - i : i);
        const long long k1 = 2 * (rev ? a.n2 - 1 - (i + stride) : i + stride);
        double2 r0 = ld2(a.r + k0), q0 = ld2(a.q + k0);
        double2 r1 = ld2(a.r + k1), q1 = ld2(a.q + k1);
        r0.x = r0.x - alfa * q0.x; r0.y = r0.y - alfa * q0.y;
        r1.x = r1.x - alfa * q1.x; r1.y = r1.y - alfa * q1.y;
        st2(a.r + k0, r0); st2(a.r + k1, r1);
        rr += r0.x * r0.x; rr += r0.y * r0.y; rr += r1.x * r1.x; rr += r1.y * r1.y;
    }
    for (; i < a.n2; i += stride) {
        const long long k0 = 2 * (rev ? a.n2 - 1 - i : i);
        double2 r0 = ld2(a.r + k0), q0 = ld2(a.q + k0);
        r0.x = r0.x - alfa * q0.x; r0.y = r0.y - alfa * q0.y;
        st2(a.r + k0, r0);
        rr += r0.x * r0.x; rr += r0.y * r0.y;
    }
    if (grid_sum2(rr, zz, a.partials, gridDim.x, blockIdx.x, &S->ticket[TICKET_UPDATE], sm)) {
        if (a.fuse_scalars == SUMS_FUSED)
            scal_after_update(S, rr, S->xx);  // sum(sol*sol) is the SpMV kernel's now
        else
            S->part[0] = rr;
    }
}

// sol += alfa p of the last iteration (CG:80), once per solve.
__global__ void __launch_bounds__(256) k_xtail(double *__restrict__ x, const double *__restrict__ p,
                                               const DevScal *S, long long n2) {
    const double alfa = S->alfa;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        double2 x2 = ld2(x + 2 * i);
        const double2 p2 = ld2(p + 2 * i);
        x2.x = x2.x + alfa * p2.x; x2.y = x2.y + alfa * p2.y;
        st2(x + 2 * i, x2);
    }
}

// ---------------------------------------------------------------- persistent CG solve

// Sum of a triple over the block with one barrier pair; result in thread 0.  sm: 96 doubles.
__device__ __forceinline__ void block_sum3(double &a, double &b, double &c, double *sm) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if (lane == 0) { sm[w] = a; sm[32 + w] = b; sm[64 + w] = c; }
    __syncthreads();
    if (w == 0) {
        a = (lane < nw) ? sm[lane] : 0.0;
        b = (lane < nw) ? sm[32 + lane] : 0.0;
        c = (lane < nw) ? sm[64 + lane] : 0.0;
        a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    }
}

// Grid-wide (and, in peer mode, all-rank) sum of a triple for the persistent
// solve.  Every CTA stores its partial sums in its own slot, fences, and adds
// one to an arrival counter that only grows (sum number `gen`, counted over the
// life of the context, is complete at gen * G) -- see step_allsum for the
// single-GPU timing of this against a counter + generation word.
//   one rank:  thread 0 of every CTA watches the counter (acquire), then all
//              CTAs add the G slots in slot order: same bits everywhere.
//   peer mode: nobody watches the counter.  The CTA whose arrival completes it
//              adds the slots and posts the rank's sums into every rank's comm
//              block over NVLink; every CTA of every rank waits for all ranks'
//              flags there and adds the posted sums in rank order.  The local
//              release word, the re-add in every CTA and the separate
//              post-after-barrier step of a two-level scheme are gone.
// Slots alternate between two sets (gen & 1): a CTA reaches gen+2 only after
// gen+1 is complete, i.e. after whoever read the slots of gen has arrived
// there.  slots: [2][G][4] doubles.
__device__ __forceinline__ bool persist_allsum(double &a, double &b, double &c, double *slots,
                                               unsigned long long *counter,
                                               const unsigned long long gen, const int G,
                                               const bool peer, const bool wrote_peer,
                                               const bool rows_went_out,
                                               const PeerComm &pc, const unsigned long long ph,
                                               DevScal *S, double *sm, double *s_tot) {
    __shared__ int s_last, s_bad;
    const int tid = threadIdx.x;
    if (tid == 0) s_bad = 0;
    double *const my_slots = slots + (size_t)(gen & 1ULL) * G * 4;
    const unsigned long long target = gen * (unsigned long long)G;
    block_sum3(a, b, c, sm);
    if (tid == 0) {
        double *o = my_slots + 4 * blockIdx.x;
        st2(o, make_double2(a, b));
        st2(o + 2, make_double2(c, 0.0));
        // release: my CTA's field stores (ordered before this by the barrier inside
        // block_sum3) and the slot; system-wide if rows went into a neighbour GPU
        if (wrote_peer) __threadfence_system();
        else asm volatile("fence.acq_rel.gpu;" ::: "memory");
        if (peer) {
            const unsigned long long old = atomicAdd(counter, 1ULL);
            s_last = (old + 1ULL == target);
            // acquire the others' slots.  System scope where this phase stored ghost rows
            // into the neighbour GPUs (the post that follows releases them to all ranks);
            // after the SpMV phase nothing but the sums themselves crosses NVLink
            if (s_last) {
                if (rows_went_out) asm volatile("fence.acq_rel.sys;" ::: "memory");
                else asm volatile("fence.acq_rel.gpu;" ::: "memory");
            }
        } else {
            atomicAdd(counter, 1ULL);
            const long long t0 = clock64();
            unsigned spins = 0;
            while (ld_acquire_gpu(counter) < target) {
                if ((++spins & 255u) == 0u && clock64() - t0 > 4000000000LL) { S->error = 1; s_bad = 1; break; }
            }
            s_last = 1;                  // one rank: every CTA adds the slots
        }
    }
    __syncthreads();
    if (s_last) {
        double x = 0.0, y = 0.0, z = 0.0;
        for (int i = tid; i < G; i += blockDim.x) {
            const double *o = my_slots + 4 * i;
            const double2 u = __ldcg(reinterpret_cast<const double2 *>(o));
            const double2 w = __ldcg(reinterpret_cast<const double2 *>(o + 2));
            x += u.x; y += u.y; z += w.x;
        }
        __syncthreads();                 // block_sum3's scratch is free again
        block_sum3(x, y, z, sm);                 // totals in every lane of warp 0
        if (peer) {
            if (tid < 32) peer_post_lanes(pc, ph, x, y, z);
        } else if (tid == 0) {
            s_tot[0] = x; s_tot[1] = y; s_tot[2] = z;
        }
    }
    if (peer) return peer_wait_sums(pc, ph, S, a, b, &c);     // barriers inside
    __syncthreads();
    a = s_tot[0]; b = s_tot[1]; c = s_tot[2];
    const bool ok = (s_bad == 0);
    __syncthreads();
    return ok;
}

// PDEODE_SUM_ORDER=serial (one rank) inside the persistent solve: behind the grid-wide sum --
// which is then only the barrier -- CTA 0 re-does the two sums in the reference's cell order,
// and a second barrier hands the result to every CTA.
__device__ __forceinline__ bool persist_serial(const SerialJob *jb, double &v0, double &v1,
                                            const PersistArgs &A, const GridDesc &g, double *stage,
                                            unsigned &nser, unsigned long long &gen,
                                            const unsigned long long ph, double *sm, double *s_tot) {
    double *slot = A.ser_out + (size_t)(nser++ & 1u) * SER_MAXJ;
    if (blockIdx.x == 0) serial_reduce(g, jb, 2, stage, slot);
    double z0 = 0.0, z1 = 0.0, z2 = 0.0;
    ++gen;
    if (!persist_allsum(z0, z1, z2, A.slots, A.counter, gen, (int)gridDim.x, false, false, false,
                        A.st.pc, ph, A.st.S, sm, s_tot))
        return false;
    v0 = __ldcg(slot); v1 = __ldcg(slot + 1);
    return true;
}

// The whole CG solve (CG:54-94) in one cooperative launch: every CTA owns one
// tile (a 2*BS-column strip x a row range) for the entire solve and alternates
//   A: p = r + beta p ; q = A p ; p.q ; p.p      (bulk-async staged, as k_spmv_bulk)
//   B: x += alfa p ; r -= alfa q ; r.r ; x.x     (its own cells only)
// with one grid-wide sum after each.  All CTAs hold the same scalars and take
// the same stop decision (CG:87-90), so no kernel boundary, host round trip
// or scalar broadcast is needed per iteration; in peer mode the CTA that
// completes a grid-wide sum also posts it to the other GPUs (persist_allsum)
// and boundary rows of r go straight into the neighbours' ghost rows.  Used where launch latency would dominate (small
// grids, thin slabs of a multi-GPU run); results are bit-identical to the
// kernel-per-phase path only up to the order of the partial sums.
// SERIAL = PDEODE_SUM_ORDER=serial, a separate instantiation so that the production kernel
// pays neither registers nor shared memory for it.
template <int BS, int DEPTH, bool DEFER, bool SERIAL = false>
__global__ void __launch_bounds__(BS + 32, (BS == 256 ? 2 : (BS == 128 ? 3 : 5)))
k_cg_persist(const PersistArgs A) {
    constexpr int TC = Ring<BS, DEPTH>::TC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double sm[96];
    __shared__ double s_tot[3];
    __shared__ double s_stage[SERIAL ? 2 * SER_MAXJ * SER_CH : 1];
    Ring<BS, DEPTH> R;
    R.carve(smem_raw);

    const StencilArgs &a = A.st;
    constexpr bool defer = DEFER;          // 88-byte iteration: x += alfa p rides in phase A
    DevScal *S = a.S;
    const GridDesc g = a.g;
    const int G = gridDim.x;
    const int tid = threadIdx.x;
    const bool peer = (a.fuse_scalars == SUMS_PEER);
    const int strip = blockIdx.x % A.nstrips, chunk = blockIdx.x / A.nstrips;
    const int jt = strip * TC;
    const int i_begin = (int)((long long)chunk * g.rows / A.nchunks);
    const int i_end = (int)((long long)(chunk + 1) * g.rows / A.nchunks);
    const int j0 = jt + 2 * tid;
    const bool worker = (tid < BS) && (j0 < g.P);
    const unsigned pf_bytes = (unsigned)max(0, min(TC, g.P - jt)) * 8u;   // a row of my strip
    // the first A.prefetch rows of my tile (what all CTAs pull together has to fit L2 next to
    // the rows the running phase keeps hot)
    const int pf_rows = pf_bytes ? min(A.prefetch, i_end - i_begin) : 0;

    if (tid == 0) R.init_barriers();
    __syncthreads();

    unsigned long long ph = a.wait_phase;
    double rho, xx, rho_old = 0.0;
    if (peer) {
        if (!peer_wait_sums(a.pc, ph, S, rho, xx)) {   // posted by the INIT kernel
            if (blockIdx.x == 0 && tid == 0) { S->nit = 0; S->stop = -1; S->done = 1; S->finished = 1; }
            return;
        }
    } else {
        rho = S->rho; xx = S->xx;
    }
    const long long maxit = S->maxit;
    const double n_glob = g.n_glob, tol2 = a.np.tol2;
    // peer mode: rows 0 and rows-1 of r are also the neighbours' ghost rows
    double *const up = peer && a.pc.r_up ? a.pc.r_up + (long long)a.pc.rows_up * g.P : nullptr;
    double *const dn = peer && a.pc.r_dn ? a.pc.r_dn - g.P : nullptr;

    long long K = 0, it = 0;
    int ip = A.ip, stop = 0;
    unsigned long long gen = A.gen0;
    double alfa = 0.0, err2 = 0.0, normx = 0.0, pq = 0.0, pp = 0.0;
    unsigned nser = 0;                   // PDEODE_SUM_ORDER=serial: see persist_serial
    for (;;) {
        ++it;                                                        // CG:55
        const bool firstit = (it == 1);
        const double beta = firstit ? 0.0 : rho / rho_old;           // CG:68
        const double *p_in = A.pbuf[ip];
        double *p_out = A.pbuf[ip ^ 1];
        // ---- phase A
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
        if (tid >= BS) {
            if (tid == BS) {
                asm volatile("fence.proxy.async;" ::: "memory");     // r, p, x were written by plain stores
                bulk_produce<BS, DEPTH>(R, g, a.r, p_in, a.d, a.ac, firstit, jt, i_begin, i_end, K,
                                        (defer && !firstit) ? A.x : nullptr);
            }
        } else {
            // alfa still holds the previous iteration's value here
            bulk_consume<BS, DEPTH, DEFER>(R, g, a.np.h, p_out, a.q, firstit, beta, jt, i_begin, i_end, K,
                                    a.write_ghost_top != 0, a.write_ghost_bot != 0, acc0, acc1,
                                    defer ? A.x : nullptr, alfa, &acc2);
        }
        K += (i_end - i_begin) + 2;
        // While the sum is under way (one rank: ~3 us; all ranks over NVLink: 6-7 us) HBM would
        // idle.  The next phase's loads need none of the scalars, so the rows of my tile that
        // phase B reads and that are not L2-hot from this phase (x, r; p and q were just written)
        // are pulled into L2 now (the first pf_rows of them), one row segment per thread, in the
        // order they will be used.
        if (pf_rows)
            for (int t = tid; t < pf_rows; t += blockDim.x) {
                const long long o = (long long)(i_begin + t) * g.P + jt;
                if (!defer) bulk_prefetch_l2(A.x + o, pf_bytes);
                bulk_prefetch_l2(a.r + o, pf_bytes);
            }
        ++ph; ++gen;
        if (!persist_allsum(acc0, acc1, acc2, A.slots, A.counter, gen, G, peer, false, false, a.pc, ph, S, sm, s_tot)) {
            stop = -1;                       // an exchange timed out (DevScal::error is set): leave
            break;
        }
        if (SERIAL) {                        // dotProd(p,q) (CG:76), sum(p*p) (CG:87)
            const SerialJob jb[2] = {{p_out, a.q, SER_DOT, 0}, {p_out, p_out, SER_DOT, 0}};
            if (!persist_serial(jb, acc0, acc1, A, g, s_stage, nser, gen, ph, sm, s_tot)) { stop = -1; break; }
        }
        pq = acc0; pp = acc1;
        if (defer && !firstit) xx = acc2;                            // CG:57
        alfa = rho / pq;                                             // CG:77
        err2 = fabs(alfa) * sqrt(pp) / n_glob;                       // CG:87
        normx = sqrt(xx) / n_glob;                                   // CG:57
        stop = 0;
        if (it > maxit) stop = -1;                                   // CG:88
        if (err2 < tol2 * normx) stop = stop - 100;                  // CG:90
        if (blockIdx.x == 0 && tid == 0 && a.trace && it <= (long long)S->trace_cap) {
            double *t = a.trace + 5 * (it - 1);
            t[0] = rho; t[1] = pq; t[2] = alfa; t[3] = err2; t[4] = normx;
        }
        // ---- phase B: my own cells (the ones whose p and q I just wrote)
        double rr = 0.0, xn = 0.0;
        // the slab's last row goes first: its stores into the lower neighbour GPU then have the
        // whole phase to land before the system-wide fence that precedes my arrival at the sum
        // (the first row, for the upper neighbour, is first anyway)
        const bool last_first = dn && i_end == g.rows && i_end - i_begin > 1;
        const int i_hi = last_first ? i_end - 1 : i_end;
        if (worker && defer) {
            long long off = (long long)i_begin * g.P + j0;
            if (last_first) {
                const long long o = (long long)(i_end - 1) * g.P + j0;
                double2 r2 = ld2(a.r + o);
                const double2 q2 = ld2(a.q + o);
                r2.x = r2.x - alfa * q2.x; r2.y = r2.y - alfa * q2.y;  // CG:81
                st2(a.r + o, r2);
                st2(dn + j0, r2);
                rr += r2.x * r2.x; rr += r2.y * r2.y;
            }
            for (int i = i_begin; i < i_hi; ++i, off += g.P) {
                double2 r2 = ld2(a.r + off);
                const double2 q2 = ld2(a.q + off);
                r2.x = r2.x - alfa * q2.x; r2.y = r2.y - alfa * q2.y;  // CG:81
                st2(a.r + off, r2);
                if (up && i == 0) st2(up + j0, r2);
                if (dn && i == g.rows - 1) st2(dn + j0, r2);
                rr += r2.x * r2.x; rr += r2.y * r2.y;
            }
            asm volatile("fence.proxy.async;" ::: "memory");
        } else if (worker) {
            long long off = (long long)i_begin * g.P + j0;
            auto upd = [&](int i, long long o, double2 x2, double2 p2, double2 r2, double2 q2) {
                x2.x = x2.x + alfa * p2.x; x2.y = x2.y + alfa * p2.y;  // CG:80
                r2.x = r2.x - alfa * q2.x; r2.y = r2.y - alfa * q2.y;  // CG:81
                st2(A.x + o, x2);
                st2(a.r + o, r2);
                if (up && i == 0) st2(up + j0, r2);
                if (dn && i == g.rows - 1) st2(dn + j0, r2);
                rr += r2.x * r2.x; rr += r2.y * r2.y;
                xn += x2.x * x2.x; xn += x2.y * x2.y;
            };
            if (last_first) {
                const long long o = (long long)(i_end - 1) * g.P + j0;
                upd(i_end - 1, o, ld2(A.x + o), ld2(p_out + o), ld2(a.r + o), ld2(a.q + o));
            }
            int i = i_begin;
            for (; i + 1 < i_hi; i += 2, off += 2 * g.P) {           // two rows in flight
                const long long o1 = off + g.P;
                const double2 xa = ld2(A.x + off), pa = ld2(p_out + off);
                const double2 ra = ld2(a.r + off), qa = ld2(a.q + off);
                const double2 xb = ld2(A.x + o1), pb = ld2(p_out + o1);
                const double2 rb = ld2(a.r + o1), qb = ld2(a.q + o1);
                upd(i, off, xa, pa, ra, qa);
                upd(i + 1, o1, xb, pb, rb, qb);
            }
            if (i < i_hi) upd(i, off, ld2(A.x + off), ld2(p_out + off), ld2(a.r + off), ld2(a.q + off));
            asm volatile("fence.proxy.async;" ::: "memory");         // next phase A reads r via bulk copies
        }
        // the same for phase A of the next iteration: d, ac and the direction just written
        // (r is L2-hot: phase B stored it a moment ago)
        if (pf_rows)
            for (int t = tid; t < pf_rows; t += blockDim.x) {
                const long long o = (long long)(i_begin + t) * g.P + jt;
                bulk_prefetch_l2(a.d + o, pf_bytes);
                bulk_prefetch_l2(a.ac + o, pf_bytes);
                bulk_prefetch_l2(p_out + o, pf_bytes);
                if (defer) bulk_prefetch_l2(A.x + o, pf_bytes);
            }
        // only CTAs that stored into a neighbour GPU need the system-wide fence
        const bool wrote_peer = (up && i_begin == 0) || (dn && i_end == g.rows);
        double zz = 0.0;
        ++ph; ++gen;
        if (!persist_allsum(rr, xn, zz, A.slots, A.counter, gen, G, peer, wrote_peer, peer, a.pc, ph, S, sm, s_tot)) {
            stop = -1;
            break;
        }
        if (SERIAL) {                        // dotProd(r,r) (CG:59), sum(sol*sol) (CG:57)
            const SerialJob jb[2] = {{a.r, a.r, SER_DOT, 0}, {A.x, A.x, SER_DOT, 0}};
            if (!persist_serial(jb, rr, xn, A, g, s_stage, nser, gen, ph, sm, s_tot)) { stop = -1; break; }
        }
        rho_old = rho; rho = rr;                                     // CG:56,59
        if (!defer) xx = xn;                                         // CG:57
        ip ^= 1;
        if (stop != 0) break;
    }
    if (defer && worker) {
        // the last iteration's sol += alfa p (CG:80), on my own cells
        const double *p_last = A.pbuf[ip];
        long long off = (long long)i_begin * g.P + j0;
        for (int i = i_begin; i < i_end; ++i, off += g.P) {
            double2 x2 = ld2(A.x + off);
            const double2 p2 = ld2(p_last + off);
            x2.x = x2.x + alfa * p2.x; x2.y = x2.y + alfa * p2.y;
            st2(A.x + off, x2);
        }
    }
    if (blockIdx.x == 0 && tid == 0) {
        S->nit = it; S->stop = stop; S->done = 1; S->finished = 1;
        S->rho = rho; S->rho_old = rho_old; S->xx = xx;
        S->alfa = alfa; S->err2 = err2; S->normx = normx; S->pq = pq; S->pp = pp;
        if (a.trace) S->trace_len = (int)min(it, (long long)S->trace_cap);
    }
}

// ---------------------------------------------------------------- C update + Picard residuals

// P:504-506
__device__ __forceinline__ double solve_c_cell(double Mp, double Mn, double Cp,
                                               const NumParams &np) {
    const double g = Cp * Mp / (np.kappa + Cp);
    const double b = np.kappa - Cp + np.aux * (Mn + g);
    const double c = -np.kappa * Cp + np.aux * np.kappa * Cp * Mp / (np.kappa + Cp);
    return 0.5 * (-b + sqrt(b * b - 4.0 * c));
}

// C update (P:504-506), the two Picard residuals (P:948-963) and the field sums of the new
// iterate (calcMass, P:783-787) in one pass.  With DevScal::tail pending it also applies the
// last CG iteration's sol += alfa p (CG:80) on the way -- the update kernel of that iteration
// did nothing -- so that the final iteration of a solve and the C update cost 64 B per cell
// instead of 96.
template <int UNROLL>
__global__ void __launch_bounds__(256) k_solvec(const SolveCArgs a) {
    __shared__ double sm[64];
    DevScal *S = a.S;
    if (a.gated && (!S->finished || S->solvec_seq == a.seq)) return;
    const GridDesc g = a.g;
    const NumParams np = a.np;
    const bool tail = S->tail != 0;
    const double alfa = S->alfa;
    const double *p_last = a.pbuf[(a.ip0 ^ (int)(S->nit & 1)) & 1];
    double v[4] = {0.0, 0.0, 0.0, 0.0};       // dC, dM, sum Mnew, sum Cnew
    const long long stride = (long long)gridDim.x * blockDim.x;
    // two divisions and a square root per cell: UNROLL double2 per thread in flight
    auto cell_pair = [&](const long long k, const double2 mp, const double2 cp, const double2 cc,
                         const double2 mc, double2 mn, const double2 p2) {
        const int j = (int)(k % g.P);
        const bool v0 = j < g.col, v1 = j + 1 < g.col;
        if (tail) {
            mn.x = mn.x + alfa * p2.x; mn.y = mn.y + alfa * p2.y;     // CG:80 of the last iteration
            st2(a.Mnew + k, mn);
        }
        double2 cn;
        cn.x = v0 ? solve_c_cell(mp.x, mn.x, cp.x, np) : 0.0;
        cn.y = v1 ? solve_c_cell(mp.y, mn.y, cp.y, np) : 0.0;
        if (np.gSelect == 1) st2(a.Cnew + k, cn);        // P:501
        else cn = ld2(a.Cnew + k);                        // Cnew left as it was
        v[0] += v0 ? fabs(cc.x - cn.x) : 0.0;             // P:958
        v[0] += v1 ? fabs(cc.y - cn.y) : 0.0;
        v[1] += v0 ? fabs(mc.x - mn.x) : 0.0;
        v[1] += v1 ? fabs(mc.y - mn.y) : 0.0;
        v[2] += v0 ? mn.x : 0.0; v[2] += v1 ? mn.y : 0.0; // P:785
        v[3] += v0 ? cn.x : 0.0; v[3] += v1 ? cn.y : 0.0;
    };
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < a.n2; i += UNROLL * stride) {
        double2 mp[UNROLL], cp[UNROLL], cc[UNROLL], mc[UNROLL], mn[UNROLL], p2[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long long k = 2 * (i + u * stride);
            mp[u] = ld2(a.Mprev + k); cp[u] = ld2(a.Cprev + k); cc[u] = ld2(a.Ccur + k);
            mc[u] = ld2(a.Mcur + k); mn[u] = ld2(a.Mnew + k);
            p2[u] = tail ? ld2(p_last + k) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            cell_pair(2 * (i + u * stride), mp[u], cp[u], cc[u], mc[u], mn[u], p2[u]);
    }
    for (; i < a.n2; i += stride) {
        const long long k = 2 * i;
        cell_pair(k, ld2(a.Mprev + k), ld2(a.Cprev + k), ld2(a.Ccur + k), ld2(a.Mcur + k),
                  ld2(a.Mnew + k), tail ? ld2(p_last + k) : make_double2(0.0, 0.0));
    }
    if (grid_sum<4>(v, a.partials, gridDim.x, blockIdx.x, &S->ticket[TICKET_SOLVEC], sm)) {
        if (a.fuse_scalars) { scal_after_solvec(S, v[0], v[1], g.n_glob); scal_mass_sums(S, v[2], v[3]); }
        else { S->part[0] = v[0]; S->part[1] = v[1]; S->part[2] = v[2]; S->part[3] = v[3]; }
        S->tail = 0;
        S->solvec_seq = a.seq;
    }
}

// ---------------------------------------------------------------- whole time step, one launch

// Grid-wide sum of NV values that is also the grid barrier, for a cooperative
// launch (all CTAs resident).  Every CTA publishes its partial sums in its own
// slot, fences, and adds one to an arrival counter that only ever grows: sum
// number `gen` (counted over the life of the context) is complete once the
// counter reads gen * G.  No reset, no release word: one fence and one atomic
// on the way in, one acquire load on the way out.  Then every CTA adds the G
// slots in slot order -- the same numbers in the same order, so all CTAs hold
// the same bits.  Slots alternate between two sets (gen & 1): a CTA can only
// reach sum gen+2 after every CTA has arrived at gen+1, i.e. has finished
// reading the slots of gen.  A CTA that waits longer than ~2 s of GPU clocks
// raises *err; everybody then leaves (returns false) instead of hanging the
// device.  Measured on B200, 257 CTAs (profiles/README.md): ~2.8 us per sum,
// against 4.6 us with an arrival counter + generation word + three seq_cst
// fences, and 6+ us with per-CTA flags that every CTA polls.
//   slots: [2][G][4] doubles, counter: one u64, sm: 4*32 doubles
template <int NV>
__device__ __forceinline__ bool step_allsum(double (&v)[NV], double *slots,
                                            unsigned long long *counter, unsigned int *err,
                                            const unsigned long long gen, const int G, double *sm,
                                            double *s_tot) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    double *const my_slots = slots + (size_t)(gen & 1ULL) * G * 4;
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) sm[32 * k + w] = v[k];
    }
    __syncthreads();                     // also: all my CTA's field stores precede the fence below
    int bad = 0;
    if (w == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) v[k] = warp_sum(lane < nw ? sm[32 * k + lane] : 0.0);
        if (lane == 0) {
            double *o = my_slots + 4 * blockIdx.x;
            st2(o, make_double2(v[0], v[1]));
            if (NV > 2) st2(o + 2, make_double2(v[2], v[NV > 3 ? 3 : 2]));
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            atomicAdd(counter, 1ULL);
            const unsigned long long target = gen * (unsigned long long)G;
            const long long t0 = clock64();
            unsigned spins = 0;
            while (ld_acquire_gpu(counter) < target) {
                if ((++spins & 255u) != 0u) continue;
                if (*(volatile unsigned int *)err == 0u && clock64() - t0 <= 4000000000LL) continue;
                *(volatile unsigned int *)err = 1u;
                bad = 1;
                break;
            }
        }
    }
    bad = __syncthreads_or(bad);         // thread 0's acquire precedes everybody's later loads
    double x[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) x[k] = 0.0;
    for (int i = tid; i < G; i += blockDim.x) {
        const double *o = my_slots + 4 * i;
        const double2 a = __ldcg(reinterpret_cast<const double2 *>(o));
        x[0] += a.x; x[1] += a.y;
        if (NV > 2) {
            const double2 b = __ldcg(reinterpret_cast<const double2 *>(o + 2));
            x[2] += b.x;
            if (NV > 3) x[3] += b.y;
        }
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) x[k] = warp_sum(x[k]);
    if (lane == 0) {                     // sm is free: the barrier above is behind its last reader
#pragma unroll
        for (int k = 0; k < NV; ++k) sm[32 * k + w] = x[k];
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) x[k] = warp_sum(lane < nw ? sm[32 * k + lane] : 0.0);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < NV; ++k) s_tot[k] = x[k];
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = s_tot[k];
    return bad == 0;
}

// PDEODE_SUM_ORDER=serial for the whole-step kernel: behind a grid-wide sum (then only the
// barrier) CTA 0 re-does the NV sums in the reference's cell order; a second barrier hands
// the result to every CTA.
template <int NV>
__device__ __forceinline__ bool step_serial(double (&v)[NV], const SerialJob (&jb)[NV],
                                            const GridDesc &g, double *stage, double *ser_out,
                                            unsigned &nser, double *slots,
                                            unsigned long long *counter, unsigned int *err,
                                            unsigned long long &gen, const int G, double *sm,
                                            double *s_tot) {
    double *slot = ser_out + (size_t)(nser++ & 1u) * SER_MAXJ;
    if (blockIdx.x == 0) serial_reduce(g, jb, NV, stage, slot);
    double z[2] = {0.0, 0.0};
    if (!step_allsum<2>(z, slots, counter, err, ++gen, G, sm, s_tot)) return false;
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = __ldcg(slot + k);
    return true;
}

// The loop body of solveOrder2 (P:700-742) in one cooperative launch, for grids
// so small that a time step is a chain of launch latencies and host round trips
// (the shipped 257 x 257 run: ~6 CG iterations per step in ~3 Picard iterates).
// Every CTA owns one tile -- a strip of columns (the whole row up to P = 1024),
// a few rows -- for the whole step and walks through
//     d = D(Mprev)                                              P:544-546
//     per Picard iterate (P:706):
//         centre diagonal, r = b - A x0, x = x0                 P:551-586, CG:46-50
//         per CG iteration: p, q = A p | p.q, p.p               CG:61-76,87
//                           x += alfa p, r -= alfa q | r.r, x.x CG:57,59,79-86
//         C update, Picard residuals                            P:504-506, P:948-963
// with one grid-wide sum (= barrier, step_allsum) where the reference has a
// reduction.  Two barriers are saved: a CTA computes D for its halo rows
// itself, and the centre diagonal / first residual of the NEXT Picard iterate
// are built right behind the C update (they need nothing from other CTAs that
// is not already there), so that their sums travel with the Picard residuals;
// if the residuals then say "converged" (P:706) that work is simply dropped --
// it only wrote scratch arrays (ac, r, the spare M buffer).
// All CTAs add the partial sums in the same order, hold the same scalars and
// take the same decisions (CG:88-90, P:706, P:732); nothing goes through the
// host until the step is over.  Cell arithmetic is the kernels' own
// (stencil_tile, solve_c_cell, dfunc): only the summation order differs.
template <bool SERIAL>
__global__ void __launch_bounds__(512) k_step_persist(const StepArgs A) {
    __shared__ double sm[128];
    __shared__ double s_tot[4];
    __shared__ double s_stage[SERIAL ? 2 * SER_MAXJ * SER_CH : 1];   // PDEODE_SUM_ORDER=serial
    unsigned nser = 0;
    const StencilArgs &a = A.st;
    const GridDesc g = a.g;
    const NumParams np = a.np;
    DevScal *S = a.S;
    const int G = gridDim.x;
    const int tid = threadIdx.x;
    const int strip = blockIdx.x % A.nstrips, chunk = blockIdx.x / A.nstrips;
    // A.cw threads (whole warps) span the strip; the CTA's blockDim.x / cw row groups split
    // the chunk's rows between them, so that a thread marches as few rows as possible
    const int RG = (int)blockDim.x / A.cw, rg = tid / A.cw, tcol = tid - rg * A.cw;
    const int j0 = (strip * A.cw + tcol) * 2;
    const bool inrow = j0 < g.P;
    const bool v0 = j0 < g.col, v1 = j0 + 1 < g.col;
    const int c_begin = (int)((long long)chunk * g.rows / A.nchunks);
    const int c_rows = (int)((long long)(chunk + 1) * g.rows / A.nchunks) - c_begin;
    const int i_begin = c_begin + rg * c_rows / RG;
    const int i_end = c_begin + (rg + 1) * c_rows / RG;
    const long long off0 = (long long)i_begin * g.P + j0;
    const long long maxit = S->maxit;
    const double n_glob = g.n_glob, tol2 = np.tol2;
    unsigned long long gen = A.gen0;
    unsigned int *const err = A.bar + 2;
    int nst = 0;
    auto stamp = [&]() {
        if (A.stamps && blockIdx.x == 0 && tid == 0 && nst < 48) {
            unsigned long long ns;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
            A.out->stamp_ns[nst] = ns;
            A.out->stamp_clk[nst] = (unsigned long long)clock64();
            ++nst;
        }
    };
    stamp();

    int iM = A.iM, iC = A.iC, iMnew = A.iMnew0, ip = A.ip;
    bool ok = true, warm = A.warm != 0;
    double diffC = 1.0, diffM = 1.0;
    int count = 0, cg_cap = 0, picard_cap = 0;
    long long nsum = 0, nmax = 0;
    double rho = 0.0, rho_old = 0.0, xx = 0.0, alfa = 0.0, err2 = 0.0, normx = 0.0, pq = 0.0, pp = 0.0;
    long long it = 0;
    int stop = 0;
    // what the caller accumulates over time steps (P:722-723, P:739-740)
    int steps_done = 0, count_max = 0;
    long long count_sum = 0, nit_sum = 0, nit_max = 0;

    // A.nsteps time steps back to back: nothing between two steps needs the host
    for (int sidx = 0; sidx < A.nsteps; ++sidx) {
    const int iMprev = iM, iCprev = iC;              // Mprev = M ; Cprev = C (P:703-704)

    // d = D(Mprev) on my rows and the rows above and below (the neighbours store
    // the same bits there); d left and right of a strip comes from other CTAs
    if (inrow) {
        const double *Mp = A.Mb[iMprev];
        double *d = const_cast<double *>(a.d);
        const int lo = max(i_begin - 1, 0), hi = min(i_end + 1, g.rows);
        long long off = (long long)lo * g.P + j0;
        for (int i = lo; i < hi; ++i, off += g.P) {
            const double2 m = ld2(Mp + off);
            double2 o;
            o.x = dfunc(m.x, np);
            o.y = dfunc(m.y, np);
            st2(d + off, o);
        }
    }
    if (A.nstrips > 1) {
        double z[2] = {0.0, 0.0};
        ok = step_allsum<2>(z, A.slots, A.counter, err, ++gen, G, sm, s_tot);
    } else {
        __syncthreads();
    }
    stamp();

    diffC = 1.0; diffM = 1.0;                        // P:700-701
    count = 0;                                       // P:702
    nsum = 0; nmax = 0;
    const bool run = ok && (diffC + diffM > A.eSoln);   // P:706 on entry (eSoln >= 2: no iterate at all)

    // target of a solve: neither Mprev nor the current iterate; the very first
    // solve after set_state starts from (and into) the zeroed buffer (D8)
    int inew = 0;
    while (inew == iMprev || inew == iM) inew++;
    if (!warm) inew = A.iMnew0;
    TilePtrs T;
    T.d = a.d; T.ac = a.ac; T.r = a.r; T.q = a.q; T.p = nullptr;
    T.Mprev = A.Mb[iMprev];

    // ---- first Picard iterate: centre diagonal, r = b - A x0 (P:712, CG:46-50)
    if (run) {
        T.Cc = A.Cb[iC];
        T.x0 = warm ? A.Mb[iM] : A.Mb[inew];
        T.x = A.Mb[inew];
        VecLoad<MODE_INIT> V;
        V.a = T.x0; V.b = nullptr; V.first = true; V.beta = 0.0;
        double acc[2] = {0.0, 0.0};
        stencil_tile<MODE_INIT>(a, T, V, j0, i_begin, i_end, acc[0], acc[1]);
        ok = step_allsum<2>(acc, A.slots, A.counter, err, ++gen, G, sm, s_tot);
        if (SERIAL && ok) {
            const SerialJob jb[2] = {{a.r, a.r, SER_DOT, 0}, {T.x, T.x, SER_DOT, 0}};
            ok = step_serial<2>(acc, jb, g, s_stage, A.ser_out, nser, A.slots, A.counter, err, gen, G, sm, s_tot);
        }
        rho = acc[0]; xx = acc[1];                   // CG:53
        stamp();
    }
    while (run && ok) {                              // P:706, left by the breaks below
        double *const x = A.Mb[inew];
        // ---- solveLSDIAG (CG:54-94)
        it = 0; stop = 0; rho_old = 0.0;             // CG:44,53
        for (;;) {
            ++it;                                                        // CG:55
            const bool firstit = (it == 1);
            VecLoad<MODE_CG> V;
            V.a = a.r; V.b = A.pbuf[ip];
            V.first = firstit;
            V.beta = firstit ? 0.0 : rho / rho_old;                      // CG:68
            T.p = A.pbuf[ip ^ 1];
            double acc[2] = {0.0, 0.0};
            stencil_tile<MODE_CG>(a, T, V, j0, i_begin, i_end, acc[0], acc[1]);
            ok = step_allsum<2>(acc, A.slots, A.counter, err, ++gen, G, sm, s_tot);
            if (SERIAL && ok) {                    // dotProd(p,q) (CG:76), sum(p*p) (CG:87)
                const SerialJob jb[2] = {{T.p, a.q, SER_DOT, 0}, {T.p, T.p, SER_DOT, 0}};
                ok = step_serial<2>(acc, jb, g, s_stage, A.ser_out, nser, A.slots, A.counter, err, gen, G, sm, s_tot);
            }
            if (!ok) break;
            stamp();
            pq = acc[0]; pp = acc[1];
            alfa = rho / pq;                                             // CG:77
            err2 = fabs(alfa) * sqrt(pp) / n_glob;                       // CG:87
            normx = sqrt(xx) / n_glob;                                   // CG:57
            stop = 0;
            if (it > maxit) stop = -1;                                   // CG:88
            if (err2 < tol2 * normx) stop = stop - 100;                  // CG:90
            double rx[2] = {0.0, 0.0};
            if (inrow) {
                const double *p_out = T.p;
                long long off = off0;
                for (int i = i_begin; i < i_end; ++i, off += g.P) {
                    double2 x2 = ld2(x + off), r2 = ld2(a.r + off);
                    const double2 p2 = ld2(p_out + off), q2 = ld2(a.q + off);
                    x2.x = x2.x + alfa * p2.x; x2.y = x2.y + alfa * p2.y;  // CG:80
                    r2.x = r2.x - alfa * q2.x; r2.y = r2.y - alfa * q2.y;  // CG:81
                    st2(x + off, x2);
                    st2(a.r + off, r2);
                    rx[0] += r2.x * r2.x; rx[0] += r2.y * r2.y;
                    rx[1] += x2.x * x2.x; rx[1] += x2.y * x2.y;
                }
            }
            ok = step_allsum<2>(rx, A.slots, A.counter, err, ++gen, G, sm, s_tot);
            if (SERIAL && ok) {                    // dotProd(r,r) (CG:59), sum(sol*sol) (CG:57)
                const SerialJob jb[2] = {{a.r, a.r, SER_DOT, 0}, {x, x, SER_DOT, 0}};
                ok = step_serial<2>(rx, jb, g, s_stage, A.ser_out, nser, A.slots, A.counter, err, gen, G, sm, s_tot);
            }
            if (!ok) break;
            stamp();
            rho_old = rho; rho = rx[0]; xx = rx[1];                      // CG:56,59,57
            ip ^= 1;
            if (stop != 0) break;
        }
        if (!ok) break;
        iMnew = inew;
        if (stop == -1 || stop == -101) cg_cap = 1;

        // ---- solveC (P:717) and the sums of calcDiff (P:725-726) on my cells
        int icnew = 0;
        while (icnew == iCprev || icnew == iC) icnew++;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};        // dC, dM, then r.r and x.x of the next iterate
        if (inrow) {
            const double *Mp = A.Mb[iMprev], *Mn = A.Mb[inew], *Cp = A.Cb[iCprev];
            const double *Cc = A.Cb[iC], *Mc = A.Mb[iM];
            double *Cn = A.Cb[icnew];
            long long off = off0;
            for (int i = i_begin; i < i_end; ++i, off += g.P) {
                const double2 mp = ld2(Mp + off), mn = ld2(Mn + off), cp = ld2(Cp + off);
                const double2 cc = ld2(Cc + off), mc = ld2(Mc + off);
                double2 cn;
                cn.x = v0 ? solve_c_cell(mp.x, mn.x, cp.x, np) : 0.0;
                cn.y = v1 ? solve_c_cell(mp.y, mn.y, cp.y, np) : 0.0;
                if (np.gSelect == 1) st2(Cn + off, cn);      // P:501
                else cn = ld2(Cn + off);                      // Cnew left as it was
                acc[0] += v0 ? fabs(cc.x - cn.x) : 0.0;       // P:958
                acc[0] += v1 ? fabs(cc.y - cn.y) : 0.0;
                acc[1] += v0 ? fabs(mc.x - mn.x) : 0.0;
                acc[1] += v1 ? fabs(mc.y - mn.y) : 0.0;
            }
        }
        const double *const ser_Cc = A.Cb[iC], *const ser_Mc = A.Mb[iM];   // the previous iterate
        iC = icnew;                                  // C = Cnew (P:728)
        iM = inew;                                   // M = Mnew (P:729)

        // ---- the next iterate's centre diagonal and r = b - A x0.  It reads my own new C,
        // and M = Mnew whose halo rows were complete two barriers ago: it can be built ahead
        // of the verdict, so that its sums travel with the Picard residuals.
        inew = 0;
        while (inew == iMprev || inew == iM) inew++;
        auto build_next = [&](double &s0, double &s1) {
            T.Cc = A.Cb[iC];
            T.x0 = A.Mb[iM];
            T.x = A.Mb[inew];
            VecLoad<MODE_INIT> V;
            V.a = T.x0; V.b = nullptr; V.first = true; V.beta = 0.0;
            stencil_tile<MODE_INIT>(a, T, V, j0, i_begin, i_end, s0, s1);
        };
        // serial-order verification: calcDiff x2 (P:725-726) in cell order NOW -- build_next
        // below overwrites the buffer of the previous iterate (it becomes the next solve's target)
        double dser[2] = {0.0, 0.0};
        if (SERIAL) {
            double z[2] = {0.0, 0.0};
            ok = step_allsum<2>(z, A.slots, A.counter, err, ++gen, G, sm, s_tot);   // every CTA's Cnew is there
            if (ok) {
                const SerialJob jb[2] = {{ser_Cc, A.Cb[iC], SER_ABSDIFF, 0}, {ser_Mc, A.Mb[iM], SER_ABSDIFF, 0}};
                ok = step_serial<2>(dser, jb, g, s_stage, A.ser_out, nser, A.slots, A.counter, err, gen, G, sm, s_tot);
            }
            if (!ok) break;
        }
        build_next(acc[2], acc[3]);
        ok = step_allsum<4>(acc, A.slots, A.counter, err, ++gen, G, sm, s_tot);
        if (SERIAL && ok) {                          // CG:59,57 of the next solve
            const SerialJob jb[2] = {{a.r, a.r, SER_DOT, 0}, {T.x, T.x, SER_DOT, 0}};
            double nx[2] = {0.0, 0.0};
            ok = step_serial<2>(nx, jb, g, s_stage, A.ser_out, nser, A.slots, A.counter, err, gen, G, sm, s_tot);
            acc[0] = dser[0]; acc[1] = dser[1]; acc[2] = nx[0]; acc[3] = nx[1];
        }
        if (!ok) break;
        stamp();
        diffC = acc[0] / n_glob;                     // P:961
        diffM = acc[1] / n_glob;

        if (it > nmax) nmax = it;                    // P:722
        nsum += it;                                  // P:723
        if (blockIdx.x == 0 && tid == 0 && count < STEP_TRACE) {
            A.out->nit[count] = it;
            A.out->diffs[2 * count] = diffC;
            A.out->diffs[2 * count + 1] = diffM;
        }
        count++;                                     // P:730
        if (count > A.picard_cap) { picard_cap = 1; break; }   // P:732
        if (!(diffC + diffM > A.eSoln)) break;       // P:706
        rho = acc[2]; xx = acc[3];                   // CG:53 of the next solve
    }
    if (!ok || picard_cap) break;
    warm = warm || count > 0;
    ++steps_done;
    count_sum += count;                              // P:740
    if (count > count_max) count_max = count;        // P:739
    nit_sum += nsum;
    if (nmax > nit_max) nit_max = nmax;
    }   // time steps
    // calcMass of the state the launch leaves (P:783-787): the diagnostics that follow a stretch
    // of steps (P:665-669) then need no pass and no launch of their own
    double msum[2] = {0.0, 0.0};
    bool have_sums = false;
    if (!SERIAL && ok) {
        if (inrow) {
            const double *Mf = A.Mb[iM], *Cf = A.Cb[iC];
            long long off = off0;
            for (int i = i_begin; i < i_end; ++i, off += g.P) {
                const double2 m = ld2(Mf + off), c2 = ld2(Cf + off);
                msum[0] += v0 ? m.x : 0.0; msum[0] += v1 ? m.y : 0.0;     // P:785
                msum[1] += v0 ? c2.x : 0.0; msum[1] += v1 ? c2.y : 0.0;
            }
        }
        have_sums = step_allsum<2>(msum, A.slots, A.counter, err, ++gen, G, sm, s_tot);
    }
    if (blockIdx.x == 0 && tid == 0) {
        StepOut *o = A.out;
        o->sumM = msum[0]; o->sumC = msum[1]; o->have_sums = have_sums ? 1 : 0;
        o->steps_done = steps_done; o->count_sum = count_sum; o->count_max = count_max;
        o->nit_sum_all = nit_sum; o->nit_max_all = nit_max;
        o->nit_sum = nsum; o->nit_max = nmax; o->count = count;
        o->cg_cap = cg_cap; o->picard_cap = picard_cap;
        o->error = (*(volatile unsigned int *)err != 0u) ? 1 : 0;
        o->iM = iM; o->iC = iC; o->iMnew = iMnew; o->ip = ip;
        o->gen = gen;
        o->nstamps = nst;
        S->nit = it; S->stop = stop; S->done = 1; S->finished = 1;
        S->rho = rho; S->rho_old = rho_old; S->xx = xx;
        S->alfa = alfa; S->err2 = err2; S->normx = normx; S->pq = pq; S->pp = pp;
        S->diffC = diffC; S->diffM = diffM;
        S->trace_len = 0;
        __threadfence_system();
        *(volatile unsigned long long *)&o->seq = A.seq;
    }
}

// ---------------------------------------------------------------- diagnostics

// calcMass (P:783-787) for M and C, and calcPeakInterface (P:989-1001): the
// last cell in row-major order holding the maximum of M (ties: '>=' keeps the
// later cell, P:993; the running maximum starts at 0, P:988) and the last row
// with some M > 0.1 (P:997).
__global__ void __launch_bounds__(256) k_mass(const MassArgs a) {
    __shared__ double sm[64];
    __shared__ double s_v[8];
    __shared__ long long s_i[8];
    __shared__ int s_r[8];
    __shared__ int s_last;
    const GridDesc g = a.g;
    double sM = 0.0, sC = 0.0;
    double bv = -1.0; long long bi = -1; int ir = -1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n2; i += stride) {
        const long long k = 2 * i;
        const int li = (int)(k / g.P);
        const int j = (int)(k - (long long)li * g.P);
        const double2 m = ld2(a.M + k), c = ld2(a.C + k);
        const long long gidx = (long long)(g.row0 + li) * g.col + j;
#pragma unroll
        for (int cdx = 0; cdx < 2; ++cdx) {
            if (j + cdx < g.col) {
                const double mv = cdx ? m.y : m.x;
                sM += mv;
                sC += cdx ? c.y : c.x;
                if (mv >= 0.0 && (mv > bv || (mv == bv && gidx + cdx > bi))) { bv = mv; bi = gidx + cdx; }
                if (mv > 0.1) ir = max(ir, g.row0 + li);
            }
        }
    }
    // block reduce
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(FULL, bv, o);
        const long long oi = __shfl_xor_sync(FULL, bi, o);
        const int orr = __shfl_xor_sync(FULL, ir, o);
        if (ov > bv || (ov == bv && oi > bi)) { bv = ov; bi = oi; }
        ir = max(ir, orr);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { s_v[w] = bv; s_i[w] = bi; s_r[w] = ir; }
    block_sum2(sM, sC, sm);   // contains __syncthreads
    const int nb = gridDim.x;
    double *pS = a.partials;
    double *pV = a.partials + 2 * nb;
    long long *pI = reinterpret_cast<long long *>(a.partials + 3 * nb);
    long long *pR = reinterpret_cast<long long *>(a.partials + 4 * nb);
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
            if (s_v[k] > bv || (s_v[k] == bv && s_i[k] > bi)) { bv = s_v[k]; bi = s_i[k]; }
            ir = max(ir, s_r[k]);
        }
        pS[blockIdx.x] = sM; pS[nb + blockIdx.x] = sC;
        pV[blockIdx.x] = bv; pI[blockIdx.x] = bi; pR[blockIdx.x] = ir;
        __threadfence();
        const unsigned int t = atomicAdd(&a.S->ticket[TICKET_MASS], 1u);
        s_last = (t == (unsigned int)(nb - 1));
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) {
        // serial, fixed order: nb is a few hundred at most
        double tM = 0.0, tC = 0.0, v = -1.0; long long vi = -1; long long r = -1;
        for (int k = 0; k < nb; ++k) {
            tM += __ldcg(pS + k); tC += __ldcg(pS + nb + k);
            const double ov = __ldcg(pV + k); const long long oi = __ldcg(pI + k);
            if (ov > v || (ov == v && oi > vi)) { v = ov; vi = oi; }
            const long long orr = __ldcg(pR + k);
            if (orr > r) r = orr;
        }
        a.S->sumM = tM; a.S->sumC = tC;
        a.S->peak_val = v; a.S->peak_idx = vi; a.S->intfac_row = (int)r;
        a.S->ticket[TICKET_MASS] = 0u;
    }
}

// ---------------------------------------------------------------- frame decimation

// What printToFile writes (P:834-842): every filter-th row and column.  One
// thread per written point; out = [M frame | C frame], nro x nco each.
__global__ void __launch_bounds__(256) k_decimate(GridDesc g, const double *__restrict__ M,
                                                  const double *__restrict__ C, int filter,
                                                  int first_row, int nro, int nco,
                                                  double *__restrict__ out) {
    const long long n = (long long)nro * nco;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (long long)gridDim.x * blockDim.x) {
        const int oi = (int)(t / nco), oj = (int)(t - (long long)oi * nco);
        const long long src = (long long)(first_row - g.row0 + oi * filter) * g.P + (long long)oj * filter;
        out[t] = M[src];
        out[n + t] = C[src];
    }
}

// What printToFile2D writes (P:898-925): for every filter-th row the mean over
// the columns, added left to right as the reference does.  One thread per row
// (the layout that uses this has col = 4).  out = [mean M | mean C].
__global__ void __launch_bounds__(128) k_rowmeans(GridDesc g, const double *__restrict__ M,
                                                  const double *__restrict__ C, int filter,
                                                  int first_row, int nro, double *__restrict__ out) {
    const int oi = blockIdx.x * blockDim.x + threadIdx.x;
    if (oi >= nro) return;
    const double *m = M + (long long)(first_row - g.row0 + oi * filter) * g.P;
    const double *c = C + (long long)(first_row - g.row0 + oi * filter) * g.P;
    double am = 0.0, ac = 0.0;
    for (int j = 0; j < g.col; ++j) { am = am + m[j]; ac = ac + c[j]; }   // P:908-909
    out[oi] = am / g.col;                                                 // P:923
    out[nro + oi] = ac / g.col;                                           // P:924
}

// ---------------------------------------------------------------- initial conditions on the device

// setInitialConditions (P:262-439) cell by cell.  The cases that loop over
// inoculation points (4, 5, 9, 10) cost O(n * points) on the host -- 1.6e10
// point evaluations at 16384^2 with the shipped 60 points; here
// every cell walks the point list itself, in the reference's order, so the
// values are the host's bit for bit.  px / py hold the points: the random
// draws of P:312-313,332-334,409-414 (made by the caller's seeded generator)
// or the spacing formula of P:393.  Case 7 draws one random number per cell
// in sequence (P:365-372); the cells that draw are a prefix 1..K of the cell order (the test
// (double)i * xDel / col <= depth is monotone in i), so cell i takes draw number i of the
// caller's generator -- splitmix64, whose draw k depends on seed + k * constant only.  The seed
// arrives as px[0] (low 32 bits) and py[0] (high 32 bits).
__global__ void __launch_bounds__(256) k_init_cond(GridDesc g, int ic, double depth, double height,
                                                   double xDel, int npts,
                                                   const double *__restrict__ px,
                                                   const double *__restrict__ py,
                                                   double *__restrict__ M, double *__restrict__ C) {
    const long long n_loc = (long long)g.rows * g.P;
    const long long n = (long long)g.grow * g.col;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_loc;
         t += (long long)gridDim.x * blockDim.x) {
        const int li = (int)(t / g.P), j = (int)(t - (long long)li * g.P);
        if (j >= g.col) { M[t] = 0.0; C[t] = 0.0; continue; }
        const long long i = (long long)(g.row0 + li) * g.col + j + 1;     // the reference's 1-based cell
        double m = 0.0;                                                   // P:276
        if (ic == 1) {                                                    // P:279-288
            const double a = -height / ((depth * depth) * (depth * depth));
            const long long x = (i - 1) / g.col;
            const double tt = (double)x * xDel;
            const double f = a * ((tt * tt) * (tt * tt)) + height;
            m = (f <= 0) ? 0.0 : f;
        } else if (ic == 2) {                                             // P:291-292
            m = height;
        } else if (ic == 3) {                                             // P:295-305
            const double a = -height / (depth * depth);
            const long long x = i % g.col, y = i / g.grow;
            const double u = x * xDel - 0.5, v = y * xDel - 0.5;
            const double f = a * (u * u + v * v) + height;
            m = (f <= 0) ? 0.0 : f;
        } else if (ic == 4 || ic == 9) {                                  // P:308-325, P:390-402
            const double a = -height / (depth * depth);
            const long long x = i % g.col, y = i / g.grow;
            for (int k = 0; k < npts; ++k) {
                if (m <= 0) {
                    const double u = x * xDel - px[k], v = y * xDel - py[k];
                    m = m + a * (u * u + v * v) + height;
                    if (m <= 0) m = 0;
                }
            }
        } else if (ic == 5 || ic == 10) {                                 // P:328-341, P:405-421
            const double a = -height / (depth * depth);
            const long long x = i % g.col, y = i / g.grow;
            for (int k = 0; k < npts; ++k) {
                const double u = x * xDel - px[k], v = y * xDel - py[k];
                const double innoc = a * (u * u + v * v) + height;
                if (innoc >= 0) m = m + innoc;
            }
        } else if (ic == 6) {                                             // P:356-362
            const long long x = (i - 1) / g.col;
            if (x * xDel <= depth) m = height;
        } else if (ic == 7) {                                             // P:365-372
            if ((double)i * xDel / g.col <= depth) {
                const unsigned long long seed =
                    ((unsigned long long)py[0] << 32) | (unsigned long long)px[0];
                unsigned long long z = seed + (unsigned long long)i * 0x9E3779B97F4A7C15ull;
                z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
                z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
                z ^= z >> 31;
                m = ((double)(z >> 11) * (1.0 / 9007199254740992.0)) * 0.2;
            }
        } else if (ic == 8) {                                             // P:375-387
            m = (double)i / (double)n / 4.0;
            if (m >= 0.1) m = 0;
        }
        M[t] = m;
        C[t] = 1.0;                                                       // P:276
    }
}

// P:343-349: number and sum of the cells with M > 0 (case 5's normalisation)
__global__ void __launch_bounds__(256) k_ic_count_sum(GridDesc g, const double *__restrict__ M,
                                                      DevScal *S, double *partials) {
    __shared__ double sm[64];
    const long long n_loc = (long long)g.rows * g.P;
    double cnt = 0.0, tot = 0.0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_loc;
         t += (long long)gridDim.x * blockDim.x) {
        const double m = M[t];
        if (m > 0) { cnt += 1.0; tot += m; }
    }
    if (grid_sum2(cnt, tot, partials, gridDim.x, blockIdx.x, &S->ticket[TICKET_MASS], sm)) {
        S->part[0] = cnt; S->part[1] = tot;
    }
}

// P:350-352: M(j) = M(j) * height/(total/real(i)), factor from S->part (all ranks' sums)
__global__ void __launch_bounds__(256) k_ic_scale(GridDesc g, double *__restrict__ M, const DevScal *S,
                                                  double height) {
    const long long n_loc = (long long)g.rows * g.P;
    const double denom = S->part[1] / S->part[0];
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_loc;
         t += (long long)gridDim.x * blockDim.x)
        M[t] = M[t] * height / denom;
}

}  // namespace

cudaError_t launch_init_cond(const GridDesc &g, int ic, double depth, double height, double xDel,
                             int npts, const double *px, const double *py, double *M, double *C,
                             cudaStream_t st) {
    k_init_cond<<<stream_blocks(), 256, 0, st>>>(g, ic, depth, height, xDel, npts, px, py, M, C);
    return cudaGetLastError();
}
cudaError_t launch_ic_count_sum(const GridDesc &g, const double *M, DevScal *S, double *partials,
                                cudaStream_t st) {
    k_ic_count_sum<<<296, 256, 0, st>>>(g, M, S, partials);
    return cudaGetLastError();
}
cudaError_t launch_ic_scale(const GridDesc &g, double *M, const DevScal *S, double height,
                            cudaStream_t st) {
    k_ic_scale<<<stream_blocks(), 256, 0, st>>>(g, M, S, height);
    return cudaGetLastError();
}

cudaError_t launch_decimate(const GridDesc &g, const double *M, const double *C, int filter,
                            int first_row, int nro, int nco, double *out, cudaStream_t st) {
    const long long n = (long long)nro * nco;
    const int nb = (int)((n + 255) / 256 < 2048 ? (n + 255) / 256 : 2048);
    k_decimate<<<nb > 0 ? nb : 1, 256, 0, st>>>(g, M, C, filter, first_row, nro, nco, out);
    return cudaGetLastError();
}

cudaError_t launch_rowmeans(const GridDesc &g, const double *M, const double *C, int filter,
                            int first_row, int nro, double *out, cudaStream_t st) {
    k_rowmeans<<<(nro + 127) / 128 > 0 ? (nro + 127) / 128 : 1, 128, 0, st>>>(g, M, C, filter,
                                                                               first_row, nro, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- launchers

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int PERSIST_DEPTH = 4;

template <int BS>
static int persist_occupancy() {
    constexpr size_t smem = Ring<BS, PERSIST_DEPTH>::BYTES;
    // the 88-byte instantiation needs more registers; size the grid for whichever fits fewer
    if (cudaFuncSetAttribute(k_cg_persist<BS, PERSIST_DEPTH, true>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if (cudaFuncSetAttribute(k_cg_persist<BS, PERSIST_DEPTH, false>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if (cudaFuncSetAttribute(k_cg_persist<BS, PERSIST_DEPTH, false, true>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int n = 0;
    const bool want_defer = getenv("PDEODE_PERSIST_DEFER") && atoi(getenv("PDEODE_PERSIST_DEFER")) != 0;
    const bool ser = getenv("PDEODE_SUM_ORDER") && getenv("PDEODE_SUM_ORDER")[0] == 's';
    cudaError_t e = ser
        ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_cg_persist<BS, PERSIST_DEPTH, false, true>, BS + 32, smem)
        : want_defer
        ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_cg_persist<BS, PERSIST_DEPTH, true>, BS + 32, smem)
        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_cg_persist<BS, PERSIST_DEPTH, false>, BS + 32, smem);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

static int persist_ctas_per_sm(int bs) {
    if (bs == 64) return persist_occupancy<64>();
    if (bs == 128) return persist_occupancy<128>();
    return persist_occupancy<256>();
}

template <int BS>
static cudaError_t launch_cg_persist_t(const PersistArgs &a, cudaStream_t st) {
    void *args[] = {const_cast<PersistArgs *>(&a)};
    const void *fn = a.serial ? (const void *)k_cg_persist<BS, PERSIST_DEPTH, false, true>
                 : a.st.defer_x ? (const void *)k_cg_persist<BS, PERSIST_DEPTH, true>
                                : (const void *)k_cg_persist<BS, PERSIST_DEPTH, false>;
    return cudaLaunchCooperativeKernel(fn,
                                       dim3(a.nstrips * a.nchunks), dim3(BS + 32), args,
                                       Ring<BS, PERSIST_DEPTH>::BYTES, st);
}

cudaError_t launch_cg_persist(const PersistArgs &a, int BS, cudaStream_t st) {
    if (BS == 64) return launch_cg_persist_t<64>(a, st);
    if (BS == 128) return launch_cg_persist_t<128>(a, st);
    return launch_cg_persist_t<256>(a, st);
}

cudaError_t launch_step_persist(const StepArgs &a, int BS, cudaStream_t st) {
    void *args[] = {const_cast<StepArgs *>(&a)};
    const void *fn = a.serial ? (const void *)k_step_persist<true> : (const void *)k_step_persist<false>;
    return cudaLaunchCooperativeKernel(fn, dim3(a.nstrips * a.nchunks), dim3(BS), args, 0, st);
}

LaunchCfg launch_config(const GridDesc &g) {
    if (const char *e = getenv("PDEODE_STREAM_CTAS")) {
        const int v = atoi(e);
        if (v >= 1 && v <= 32) g_stream_ctas_per_sm = v;
    }
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
        g_num_sms = sms;

    // Measured on B200 at 4096^2 (profiles/r01_sweep_*.txt): the bulk-async
    // variant peaks at BS 256, depth 4, 64 rows per CTA.  Narrow grids take a
    // narrower strip; short grids take fewer rows per CTA so that about two
    // CTAs per SM exist.
    LaunchCfg c;
    c.variant = 1;
    c.DEPTH = 4;
    c.BS = (g.P <= 128) ? 64 : (g.P <= 256 ? 128 : 256);
    const int nstrips = cdiv(g.P, 2 * c.BS);
    const int want_chunks = cdiv(2 * g_num_sms, nstrips);
    c.TR = g.rows / want_chunks;
    if (c.TR > 64) c.TR = 64;
    if (c.TR < 8) c.TR = 8;
    if (const char *e = getenv("PDEODE_BS")) {
        const int v = atoi(e);
        if (v == 64 || v == 128 || v == 256) c.BS = v;
    }
    if (const char *e = getenv("PDEODE_TR")) {
        const int v = atoi(e);
        if (v >= 1 && v <= 4096) c.TR = v;
    }
    if (const char *e = getenv("PDEODE_VARIANT")) c.variant = (e[0] == 'b') ? 1 : 0;   // bulk | reg
    if (const char *e = getenv("PDEODE_DEPTH")) {
        const int v = atoi(e);
        if (v >= 2 && v <= 8) c.DEPTH = v;
    }

    // Persistent solve: one tile per CTA, all CTAs resident (cooperative launch).
    // Narrower strips for narrower grids so that enough CTAs exist.
    c.p_BS = (g.P <= 1024) ? 64 : (g.P <= 2048 ? 128 : 256);
    if (const char *e = getenv("PDEODE_PERSIST_BS")) {
        const int v = atoi(e);
        if (v == 64 || v == 128 || v == 256) c.p_BS = v;
    }
    const int per_sm = persist_ctas_per_sm(c.p_BS);
    c.p_nstrips = cdiv(g.P, 2 * c.p_BS);
    int nch = per_sm * g_num_sms / c.p_nstrips;
    const int by_rows = g.rows / 4 > 0 ? g.rows / 4 : 1;
    if (nch > by_rows) nch = by_rows;
    // several contexts sharing one GPU (ensembles): leave room for the others' grids --
    // PDEODE_GPU_SHARE=n takes 1/n of the resident CTAs, PDEODE_PERSIST_MAX_CTAS caps them
    int share = 1;
    if (const char *e = getenv("PDEODE_GPU_SHARE")) share = atoi(e) > 1 ? atoi(e) : 1;
    if (share > 1) nch = nch / share > 0 ? nch / share : 1;
    if (const char *e = getenv("PDEODE_PERSIST_MAX_CTAS")) {
        const int cap = atoi(e) / c.p_nstrips;
        if (cap >= 1 && nch > cap) nch = cap;
    }
    c.p_nchunks = nch;
    // default: where a CG iteration is short enough for launch latency to matter
    // (measured: one GPU, faster up to 2048^2 and 4 % slower at 4096^2; as a slab of a
    // multi-GPU run, where every kernel boundary also waits for the peers, still 11 %
    // faster at 8.4M cells per GPU)
    const long long cells = (long long)g.rows * g.P;
    const long long limit = (g.rows != g.grow) ? (16LL << 20) : (4LL << 20);
    c.persist = (nch >= 1 && c.variant == 1 && cells <= limit) ? 1 : 0;
    if (const char *e = getenv("PDEODE_PERSIST")) c.persist = (atoi(e) != 0 && nch >= 1) ? 1 : 0;
    // measured at 4096^2 (bench.py): 6.29e10 with the reverse sweep and 16 CTAs per SM
    // against 5.95e10 forward with 8
    c.kb_reverse = 1;
    // 88-byte iteration (one GPU, kernel-per-phase engine): 6.53e10 against 6.2e10 at 4096^2
    c.defer_x = 1;
    c.persist_defer = 0;
    // sweep at 4096^2: the 88-byte kernel runs 0.183 ms at 32 rows per CTA, 0.194 ms at 64
    c.TR_x = c.TR > 32 ? 32 : c.TR;
    if (const char *e = getenv("PDEODE_PERSIST_DEFER")) c.persist_defer = atoi(e) != 0;
    // L2 prefetch of the next phase's rows while a grid-wide sum is under way, in rows per tile:
    // -1 = as a slab of a multi-GPU run only (the all-rank sum takes 6-7 us and the slab does
    // not fit L2), as many rows as fit a 40 MB budget over all CTAs
    c.persist_prefetch = -1;
    if (const char *e = getenv("PDEODE_PERSIST_PREFETCH")) c.persist_prefetch = atoi(e) > 0 ? atoi(e) : 0;
    if (const char *e = getenv("PDEODE_DEFER_X")) c.defer_x = atoi(e) != 0;
    // INIT (sweep at 4096^2): bulk BS 128 x 32 rows 0.200 ms, BS 256 x 64 rows 0.322 ms
    c.i_BS = c.BS > 128 ? 128 : c.BS;
    c.i_TR = c.TR > 32 ? 32 : c.TR;
    if (const char *e = getenv("PDEODE_INIT_BS")) {
        const int v = atoi(e);
        if (v == 64 || v == 128 || v == 256) c.i_BS = v;
    }
    if (const char *e = getenv("PDEODE_INIT_TR")) {
        const int v = atoi(e);
        if (v >= 1 && v <= 4096) c.i_TR = v;
    }
    if (const char *e = getenv("PDEODE_KB_REVERSE")) c.kb_reverse = atoi(e) != 0;

    // Whole time step in one launch: s_cw threads span a strip (the whole row where P <= 1024),
    // a CTA holds as many row groups of s_cw threads as fit in 512 threads, a thread marches
    // PDEODE_STEP_ROWS (default 1) rows or more, all CTAs resident.
    c.s_cw = (g.P / 2 + 31) / 32 * 32;
    if (c.s_cw > 512) c.s_cw = 512;
    int rgroups = 512 / c.s_cw;
    if (const char *e = getenv("PDEODE_STEP_GROUPS")) {
        const int v = atoi(e);
        if (v >= 1 && v * c.s_cw <= 512) rgroups = v;
    }
    c.s_BS = c.s_cw * rgroups;
    c.s_nstrips = cdiv(g.P, 2 * c.s_cw);
    {
        int n = 0;
        // sized for the serial-order instantiation too (it holds 16 KB of staging)
        const bool ser = getenv("PDEODE_SUM_ORDER") && getenv("PDEODE_SUM_ORDER")[0] == 's';
        if ((ser ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_step_persist<true>, c.s_BS, 0)
                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_step_persist<false>, c.s_BS, 0)) != cudaSuccess) {
            cudaGetLastError();
            n = 0;
        }
        int maxg = n * g_num_sms;
        if (share > 1) maxg = maxg / share > c.s_nstrips ? maxg / share : c.s_nstrips;
        if (const char *e = getenv("PDEODE_PERSIST_MAX_CTAS")) {
            const int cap = atoi(e);
            if (cap >= 1 && maxg > cap) maxg = cap;
        }
        int rows_min = 1;
        if (const char *e = getenv("PDEODE_STEP_ROWS")) {
            const int v = atoi(e);
            if (v >= 1 && v <= 4096) rows_min = v;
        }
        int nch = maxg / c.s_nstrips;
        const int by_rows = cdiv(g.rows, rows_min * rgroups);
        if (nch > by_rows) nch = by_rows;
        c.s_nchunks = nch;
        // where a step is a chain of launch latencies: the shipped 257^2 grid, ensemble members.
        // Measured per step (shipped parameters, tools/step_bench.py), whole step / persistent
        // solve / kernel per phase: 257^2 59 / 136 / 182 us, 512^2 74 / 144 / 180, 1024^2
        // 151 / 228 / 247; 1448^2 449 / 333, 2048^2 702 / 549: the streaming engines win there.
        c.whole_step = (nch >= 1 && g.rows == g.grow && cells <= (1280LL << 10)) ? 1 : 0;
        if (const char *e = getenv("PDEODE_WHOLE_STEP")) c.whole_step = (atoi(e) != 0 && nch >= 1) ? 1 : 0;
    }
    return c;
}

int stream_blocks() { return g_num_sms * g_stream_ctas_per_sm; }

static int flat_blocks(long long n2, int ctas_per_sm = 0) {
    const long long want = (n2 + 255) / 256;
    const long long cap = ctas_per_sm > 0 ? (long long)g_num_sms * ctas_per_sm : stream_blocks();
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

int stencil_max_blocks(const GridDesc &g, int TR) {
    return cdiv(g.P, 2 * 64) * cdiv(g.rows, TR);
}

cudaError_t launch_diffcoef(const GridDesc &g, const NumParams &np, const double *M, double *d,
                            cudaStream_t st) {
    const long long n2 = (long long)g.rows * g.P / 2;
    // sweep at 4096^2 (tools/sweep_pointwise.sh): 0.104 ms with one or two double2 per thread in
    // flight, 0.112 with four; 0.093 once the integer powers are unrolled (powi_fast); the default
    // re-timed on the final code: 0.082 ms (profiles/r02_kernel_times_4096.jsonl)
    static const int unroll = getenv("PDEODE_DIFFCOEF_UNROLL") ? atoi(getenv("PDEODE_DIFFCOEF_UNROLL")) : 2;
    if (unroll >= 4) k_diffcoef<4><<<flat_blocks((n2 + 3) / 4), 256, 0, st>>>(g, np, M, d, n2);
    else if (unroll >= 2) k_diffcoef<2><<<flat_blocks((n2 + 1) / 2), 256, 0, st>>>(g, np, M, d, n2);
    else k_diffcoef<1><<<flat_blocks(n2), 256, 0, st>>>(g, np, M, d, n2);
    return cudaGetLastError();
}

template <int MODE>
static cudaError_t launch_stencil(const StencilArgs &a, cudaStream_t st) {
    const int bs = a.BS;
    dim3 grid(cdiv(a.g.P, 2 * bs), cdiv(a.g.rows, a.TR));
    if (bs == 64) k_stencil<MODE, 64><<<grid, 64, 0, st>>>(a);
    else if (bs == 256) k_stencil<MODE, 256><<<grid, 256, 0, st>>>(a);
    else k_stencil<MODE, 128><<<grid, 128, 0, st>>>(a);
    return cudaGetLastError();
}

template <int BS, int DEPTH>
static cudaError_t launch_init_bulk(const StencilArgs &a, cudaStream_t st) {
    constexpr size_t smem = Ring<BS, DEPTH>::BYTES;
    // the attribute is per device: one flag per device, set by whichever thread gets there
    // first (setting it twice is harmless)
    static std::atomic<bool> attr_set[MAX_DEVICES];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= MAX_DEVICES || !attr_set[dev].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(k_init_bulk<BS, DEPTH>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < MAX_DEVICES) attr_set[dev].store(true, std::memory_order_release);
    }
    dim3 grid(cdiv(a.g.P, 2 * BS), cdiv(a.g.rows, a.TR));
    k_init_bulk<BS, DEPTH><<<grid, BS + 32, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_stencil_init(const StencilArgs &a, cudaStream_t st) {
    if (a.variant == 1) {
        // ring depth of the INIT kernel: two IEEE divides per cell make it latency bound, so a
        // shallower ring (more CTAs per SM) can beat a deeper one (PDEODE_INIT_DEPTH = 2 | 4)
        static const int depth = getenv("PDEODE_INIT_DEPTH") ? atoi(getenv("PDEODE_INIT_DEPTH")) : 4;
        if (depth <= 2) {
            if (a.BS == 64) return launch_init_bulk<64, 2>(a, st);
            if (a.BS == 128) return launch_init_bulk<128, 2>(a, st);
            return launch_init_bulk<256, 2>(a, st);
        }
        if (a.BS == 64) return launch_init_bulk<64, 4>(a, st);
        if (a.BS == 128) return launch_init_bulk<128, 4>(a, st);
        return launch_init_bulk<256, 4>(a, st);
    }
    return launch_stencil<MODE_INIT>(a, st);
}
template <int BS, int DEPTH>
static cudaError_t launch_spmv_bulk(const StencilArgs &a, cudaStream_t st) {
    constexpr int TC = 2 * BS;
    constexpr size_t smem = Ring<BS, DEPTH>::BYTES;
    static std::atomic<bool> attr_set[MAX_DEVICES];      // per device, see launch_init_bulk
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= MAX_DEVICES || !attr_set[dev].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(k_spmv_bulk<BS, DEPTH, false>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_spmv_bulk<BS, DEPTH, true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < MAX_DEVICES) attr_set[dev].store(true, std::memory_order_release);
    }
    dim3 grid(cdiv(a.g.P, TC), cdiv(a.g.rows, a.TR));
    if (a.defer_x) k_spmv_bulk<BS, DEPTH, true><<<grid, BS + 32, smem, st>>>(a);
    else k_spmv_bulk<BS, DEPTH, false><<<grid, BS + 32, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_stencil_cg(const StencilArgs &a, cudaStream_t st) {
    if (a.variant == 1) {
        const int bs = a.BS, dp = a.DEPTH;
        if (bs == 256) {
            if (dp <= 2) return launch_spmv_bulk<256, 2>(a, st);
            if (dp <= 4) return launch_spmv_bulk<256, 4>(a, st);
            return launch_spmv_bulk<256, 6>(a, st);
        }
        if (bs == 64) {
            if (dp <= 4) return launch_spmv_bulk<64, 4>(a, st);
            return launch_spmv_bulk<64, 8>(a, st);
        }
        if (dp <= 2) return launch_spmv_bulk<128, 2>(a, st);
        if (dp <= 4) return launch_spmv_bulk<128, 4>(a, st);
        if (dp <= 6) return launch_spmv_bulk<128, 6>(a, st);
        return launch_spmv_bulk<128, 8>(a, st);
    }
    return launch_stencil<MODE_CG>(a, st);
}

cudaError_t launch_update(const UpdateArgs &a, cudaStream_t st) {
    k_update<<<flat_blocks((a.n2 + 1) / 2), 256, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_update_r(const UpdateArgs &a, cudaStream_t st) {
    k_update_r<<<flat_blocks((a.n2 + 1) / 2), 256, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_xtail(double *x, const double *p, const DevScal *S, long long n2,
                         cudaStream_t st) {
    k_xtail<<<flat_blocks(n2), 256, 0, st>>>(x, p, S, n2);
    return cudaGetLastError();
}

cudaError_t launch_solvec(const SolveCArgs &a, cudaStream_t st) {
    // The C update is bound by instruction issue (two IEEE divisions and a square root per
    // cell: 65 % issue slots busy, fp64 pipe 39 %, profiles/r01_ncu_full_v3_production.md), not
    // by loads in flight: the same sweep has one double2 per thread and 8 CTAs per SM at
    // 0.180 ms, two at 0.196, 16 CTAs per SM at 0.186 / 0.207.
    static const int unroll = getenv("PDEODE_SOLVEC_UNROLL") ? atoi(getenv("PDEODE_SOLVEC_UNROLL")) : 1;
    static const int ctas = getenv("PDEODE_STREAM_CTAS") ? 0 : 8;     // 0: the global setting
    if (unroll >= 2) k_solvec<2><<<flat_blocks((a.n2 + 1) / 2, ctas), 256, 0, st>>>(a);
    else k_solvec<1><<<flat_blocks(a.n2, ctas), 256, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_mass(const MassArgs &a, cudaStream_t st) {
    int nb = flat_blocks(a.n2);
    if (nb > 296) nb = 296;
    k_mass<<<nb, 256, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_serial_reduce(const GridDesc &g, const SerialJob *jobs, int nj, double *out,
                                 cudaStream_t st) {
    SerialArgs a{};
    a.g = g; a.nj = nj < SER_MAXJ ? nj : SER_MAXJ; a.out = out;
    for (int q = 0; q < a.nj; ++q) a.job[q] = jobs[q];
    k_serial_reduce<<<1, 256, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_scal_after_init(DevScal *S, cudaStream_t st) {
    k_scal_after_init<<<1, 1, 0, st>>>(S);
    return cudaGetLastError();
}
cudaError_t launch_scal_after_spmv(DevScal *S, GridDesc g, double tol2, double *trace,
                                   cudaStream_t st) {
    k_scal_after_spmv<<<1, 1, 0, st>>>(S, g, tol2, trace);
    return cudaGetLastError();
}
cudaError_t launch_scal_after_update(DevScal *S, cudaStream_t st) {
    k_scal_after_update<<<1, 1, 0, st>>>(S);
    return cudaGetLastError();
}
cudaError_t launch_scal_after_solvec(DevScal *S, GridDesc g, cudaStream_t st) {
    k_scal_after_solvec<<<1, 1, 0, st>>>(S, g);
    return cudaGetLastError();
}

}  // namespace pdeode
