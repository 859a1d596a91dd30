/*
 * pdeode_oracle.c -- CPU restatement of the reference's implicit time step.
 * TEST INFRASTRUCTURE ONLY (see pdeode_oracle.h).  PARITY UNPINNED against
 * the reference binary (no Fortran compiler in this image); pinned by KATs.
 *
 * Written from the reference's behaviour, statement by statement, keeping
 * its 1-based index arithmetic visible (arrays are indexed [i-1]) so every
 * condition can be compared with the cited line.  Build: see Makefile
 * (-O2 -ffp-contract=off: the reference's x86-64 baseline build has no FMA
 * and gfortran does not reassociate without -ffast-math, so sums are serial
 * and every product/sum rounds separately).
 *
 * With -fopenmp the loops the reference parallelises in its *_paral.sh
 * builds (runAll_paral.sh:6) carry the same directives; the loops it leaves
 * serial (sum(), maxval, CG:41,57,86-87) stay serial.
 */
#include "pdeode_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* gfortran lowers real**integer_variable to __builtin_powi -> libgcc's
 * __powidf2 (libgcc2.c): binary exponentiation from the low bit up. */
double oracle_powi(double x, int m)
{
    unsigned int n = (m < 0) ? -(unsigned int)m : (unsigned int)m;
    double y = (n % 2) ? x : 1.0;
    while (n >>= 1) {
        x = x * x;
        if (n % 2)
            y *= x;
    }
    return (m < 0) ? 1.0 / y : y;
}

/* P:464-472.  d is left as-is (0 here) for an unknown dSelect. */
double oracle_dfunc(double M, double delta, int alpha, int beta, int dSelect)
{
    double d = 0.0;
    if (dSelect == 1) d = delta;                                   /* P:469 */
    if (dSelect == 2) d = delta * oracle_powi(M, alpha);           /* P:470 */
    if (dSelect == 3)                                              /* P:471 */
        d = delta * oracle_powi(M, alpha) / oracle_powi(1.0 - M, beta);
    return d;
}

/* P:445-458 */
double oracle_ffunc(double M, double C, double kappa, double nu, int fSelect)
{
    double f = 0.0;
    double eps = C * 0.00000001;                                   /* P:451 */
    if (fSelect == 1) f = C / (kappa + C) - nu;
    if (fSelect == 2) f = C / (kappa + C);
    if (fSelect == 3) f = C / (kappa * M + C + eps) - nu;
    if (fSelect == 4) f = C / (kappa * M + C + eps);
    if (fSelect == 5) f = 1.0;
    if (fSelect == 6) f = C / (kappa + C) * (1.0 - M / (C + eps)) - nu;
    return f;
}

#define A_(i, k) A[(size_t)((k) - 1) * (size_t)n + (size_t)((i) - 1)]

/* P:517-590 */
void oracle_gen_matrix(const double *M, const double *C, double *A, int *ioff,
                       double *rhs, int row, int col, const oracle_params *p)
{
    const long n = (long)row * (long)col;
    const double tDel = p->tDel, xDel = p->xDel;
    const double xCof = 1.0 / (xDel * xDel);                       /* P:537 */
    double *diff = (double *)malloc(sizeof(double) * (size_t)n);   /* P:530 */
    long i;

    ioff[0] = -col; ioff[1] = -1; ioff[2] = 0; ioff[3] = 1; ioff[4] = col; /* P:539 */
    memset(A, 0, sizeof(double) * 5u * (size_t)n);                 /* P:540 */

#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (i = 1; i <= n; i++)                                       /* P:544-546 */
        diff[i - 1] = oracle_dfunc(M[i - 1], p->delta, p->alpha, p->beta,
                                   p->dSelect);

#define D_(k) diff[(k) - 1]
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (i = 1; i <= n; i++) {                                     /* P:551-587 */
        double f;
        if (i <= col) {                                            /* P:552 */
            A_(i, 5) = A_(i, 5) - xCof * 0.5 * (D_(i + col) + D_(i));
            A_(i, 3) = A_(i, 3) + xCof * 0.5 * (D_(i + col) + D_(i));
        } else {
            A_(i, 1) = A_(i, 1) - xCof * 0.5 * (D_(i - col) + D_(i));
            A_(i, 3) = A_(i, 3) + xCof * 0.5 * (D_(i - col) + D_(i));
        }
        if (i % col == 1) {                                        /* P:560 */
            A_(i, 4) = A_(i, 4) - xCof * 0.5 * (D_(i + 1) + D_(i));
            A_(i, 3) = A_(i, 3) + xCof * 0.5 * (D_(i + 1) + D_(i));
        } else {
            A_(i, 2) = A_(i, 2) - xCof * 0.5 * (D_(i - 1) + D_(i));
            A_(i, 3) = A_(i, 3) + xCof * 0.5 * (D_(i - 1) + D_(i));
        }
        if (i % col == 0) {                                        /* P:568 */
            A_(i, 2) = A_(i, 2) - xCof * 0.5 * (D_(i - 1) + D_(i));
            A_(i, 3) = A_(i, 3) + xCof * 0.5 * (D_(i - 1) + D_(i));
        } else {
            A_(i, 4) = A_(i, 4) - xCof * 0.5 * (D_(i + 1) + D_(i));
            A_(i, 3) = A_(i, 3) + xCof * 0.5 * (D_(i + 1) + D_(i));
        }
        if (i >= n - col) {                     /* P:576: '>=' kept (quirk Q1) */
            A_(i, 1) = A_(i, 1) - xCof * 0.5 * (D_(i - col) + D_(i));
            A_(i, 3) = A_(i, 3) + xCof * 0.5 * (D_(i - col) + D_(i));
        } else {
            A_(i, 5) = A_(i, 5) - xCof * 0.5 * (D_(i + col) + D_(i));
            A_(i, 3) = A_(i, 3) + xCof * 0.5 * (D_(i + col) + D_(i));
        }
        f = oracle_ffunc(M[i - 1], C[i - 1], p->kappa, p->nu, p->fSelect); /* P:584 */
        A_(i, 3) = A_(i, 3) - f + (1.0 / tDel);                    /* P:585 */
        rhs[i - 1] = M[i - 1] / tDel;                              /* P:586 */
    }
#undef D_
    free(diff);
}

/* P:1044-1105 */
void oracle_amuxd(long n, const double *x, double *y, const double *diag,
                  int idiag, const int *ioff)
{
    long i;
    int j;
#ifdef _OPENMP
#pragma omp parallel private(j)
#endif
    {
#ifdef _OPENMP
#pragma omp for
#endif
        for (i = 1; i <= n; i++)                                   /* P:1087-1089 */
            y[i - 1] = 0.0;
        for (j = 1; j <= idiag; j++) {                             /* P:1093 */
            const long io = ioff[j - 1];
            const long i1 = (1 - io > 1) ? 1 - io : 1;             /* P:1095 */
            const long i2 = (n - io < n) ? n - io : n;             /* P:1096 */
            const double *dj = diag + (size_t)(j - 1) * (size_t)n;
            long ii;
#ifdef _OPENMP
#pragma omp for
#endif
            for (ii = i1; ii <= i2; ii++)                          /* P:1098-1100 */
                y[ii - 1] = y[ii - 1] + dj[ii - 1] * x[ii + io - 1];
        }
    }
}

/* CG:99-114 */
double oracle_dotprod(long n, const double *u, const double *v)
{
    double sol = 0.0;
    long i;
#ifdef _OPENMP
#pragma omp parallel for reduction(+ : sol)
#endif
    for (i = 0; i < n; i++)
        sol = sol + u[i] * v[i];
    return sol;
}

/* sum(v*v): a serial intrinsic in the reference (CG:57,86,87). */
static double sum_sq(long n, const double *v)
{
    double s = 0.0;
    long i;
    for (i = 0; i < n; i++)
        s = s + v[i] * v[i];
    return s;
}

/* CG:1-95 */
void oracle_solve_ls_diag(long n, int ndiag, const int *ioff, const double *A,
                          double *sol, const double *rhs, long long *nit,
                          double *err1, double *err2, int *stopcrit,
                          double *scratch, oracle_cg_trace *trace,
                          long trace_cap)
{
    double *r = scratch, *p = scratch + n, *q = scratch + 2 * n;
    double rho, beta, rhoold, alfa, dot;
    double tol2, normx;
    long long maxit, it;
    long i;
    int stop;

    maxit = *nit; tol2 = *err2;                                    /* CG:40 */
    /* CG:41 normA/normb are computed and never used: omitted. */
    rho = 0.0;                                                     /* CG:44 */

    oracle_amuxd(n, sol, r, A, ndiag, ioff);                       /* CG:46 */
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (i = 0; i < n; i++)                                        /* CG:48-50 */
        r[i] = rhs[i] - r[i];

    it = 0; stop = 0;                                              /* CG:53 */
    while (stop == 0) {                                            /* CG:54 */
        it = it + 1;
        rhoold = rho;
        normx = sqrt(sum_sq(n, sol)) / (double)n;                  /* CG:57 */
        rho = oracle_dotprod(n, r, r);                             /* CG:59 */
        if (it == 1) {                                             /* CG:61-66 */
#ifdef _OPENMP
#pragma omp parallel for
#endif
            for (i = 0; i < n; i++)
                p[i] = r[i];
        } else {                                                   /* CG:68-72 */
            beta = rho / rhoold;
#ifdef _OPENMP
#pragma omp parallel for
#endif
            for (i = 0; i < n; i++)
                p[i] = r[i] + beta * p[i];
        }
        oracle_amuxd(n, p, q, A, ndiag, ioff);                     /* CG:75 */
        dot = oracle_dotprod(n, p, q);                             /* CG:76 */
        alfa = rho / dot;                                          /* CG:77 */
#ifdef _OPENMP
#pragma omp parallel for
#endif
        for (i = 0; i < n; i++) {                                  /* CG:79-83 */
            sol[i] = sol[i] + alfa * p[i];
            r[i] = r[i] - alfa * q[i];
        }
        *err1 = sqrt(sum_sq(n, r)) / (double)n;                    /* CG:86 */
        *err2 = fabs(alfa) * sqrt(sum_sq(n, p)) / (double)n;       /* CG:87 */
        if (it > maxit) stop = -1;                                 /* CG:88 */
        if (*err2 < tol2 * normx) stop = stop - 100;               /* CG:90 */
        if (trace && it <= trace_cap) {
            oracle_cg_trace *t = &trace[it - 1];
            t->rho = rho; t->dot = dot; t->alfa = alfa;
            t->err2 = *err2; t->normx = normx;
        }
    }
    *nit = it;
    *stopcrit = stop;
}

/* P:482-511 */
void oracle_solve_c(const double *M, const double *Mnew, const double *Csol,
                    double *Cnew, long n, double kappa, double gama,
                    double tDel, int gSelect)
{
    const double aux = tDel * 0.5 * gama;                          /* P:499 */
    long i;
    if (gSelect == 1) {                                            /* P:501 */
#ifdef _OPENMP
#pragma omp parallel for
#endif
        for (i = 0; i < n; i++) {                                  /* P:503-507 */
            double b = kappa - Csol[i]
                     + aux * (Mnew[i] + Csol[i] * M[i] / (kappa + Csol[i]));
            double c = -kappa * Csol[i]
                     + aux * kappa * Csol[i] * M[i] / (kappa + Csol[i]);
            Cnew[i] = 0.5 * (-b + sqrt(b * b - 4.0 * c));
        }
    }
    /* any other gSelect leaves Cnew unassigned in the reference (S8) */
}

/* P:948-963 */
double oracle_calc_diff(const double *C, const double *Cnew, int row, int col)
{
    const long n = (long)row * (long)col;
    double diff = 0.0;
    long i;
#ifdef _OPENMP
#pragma omp parallel for reduction(+ : diff)
#endif
    for (i = 0; i < n; i++)
        diff = diff + fabs(C[i] - Cnew[i]);
    return diff / (double)n;
}

/* P:772-791 */
double oracle_calc_mass(const double *X, long n)
{
    double total = 0.0;
    long i;
#ifdef _OPENMP
#pragma omp parallel for reduction(+ : total)
#endif
    for (i = 0; i < n; i++)
        total = total + X[i];
    return total / (double)n;
}

/* P:978-1004 */
void oracle_calc_peak_interface(const double *M, int row, int col,
                                double *peak, double *height, double *intfac)
{
    double hei = 0.0;
    int i, j;
    for (i = 1; i <= row; i++) {
        double y = (double)(i - 1) / (double)(row - 1);            /* P:990 */
        for (j = 1; j <= col; j++) {
            size_t pidx = (size_t)j + (size_t)(i - 1) * (size_t)col; /* P:992 */
            if (M[pidx - 1] >= hei) {                              /* P:993 */
                *peak = y;
                hei = M[pidx - 1];
            }
            if (M[pidx - 1] > 0.1)                                 /* P:997 */
                *intfac = y;
        }
    }
    *height = hei;
}

/* ------------------------------------------------------------------ */

struct oracle_state {
    int row, col;
    long n;
    long long maxit;
    oracle_params p;
    double *M, *C;                 /* P:73 */
    double *Mnew, *Cnew;           /* P:623-624 */
    double *A, *rhs;               /* P:625-626 */
    double *Cprev, *Mprev;         /* P:627 */
    double *scratch;               /* CG:32 r,p,q (z is never used) */
    int ioff[5];                   /* P:628 */
};

oracle_state *oracle_create(int row, int col, const oracle_params *p)
{
    oracle_state *s = (oracle_state *)calloc(1, sizeof(*s));
    size_t n = (size_t)row * (size_t)col;
    if (!s) return NULL;
    s->row = row; s->col = col; s->n = (long)n; s->p = *p;
    s->maxit = 100LL * (long long)n;                               /* P:707 */
    s->M = (double *)calloc(n, sizeof(double));
    s->C = (double *)calloc(n, sizeof(double));
    s->Mnew = (double *)calloc(n, sizeof(double));
    s->Cnew = (double *)calloc(n, sizeof(double));
    s->A = (double *)calloc(5 * n, sizeof(double));
    s->rhs = (double *)calloc(n, sizeof(double));
    s->Cprev = (double *)calloc(n, sizeof(double));
    s->Mprev = (double *)calloc(n, sizeof(double));
    s->scratch = (double *)calloc(3 * n, sizeof(double));
    if (!s->M || !s->C || !s->Mnew || !s->Cnew || !s->A || !s->rhs ||
        !s->Cprev || !s->Mprev || !s->scratch) {
        oracle_destroy(s);
        return NULL;
    }
    return s;
}

void oracle_destroy(oracle_state *s)
{
    if (!s) return;
    free(s->M); free(s->C); free(s->Mnew); free(s->Cnew); free(s->A);
    free(s->rhs); free(s->Cprev); free(s->Mprev); free(s->scratch);
    free(s);
}

void oracle_set_state(oracle_state *s, const double *M, const double *C)
{
    memcpy(s->M, M, sizeof(double) * (size_t)s->n);
    memcpy(s->C, C, sizeof(double) * (size_t)s->n);
    memset(s->Mnew, 0, sizeof(double) * (size_t)s->n);             /* D8 */
}

void oracle_get_state(const oracle_state *s, double *M, double *C)
{
    memcpy(M, s->M, sizeof(double) * (size_t)s->n);
    memcpy(C, s->C, sizeof(double) * (size_t)s->n);
}

/* Continue a run from a step boundary: there Mnew == M (P:729), which is the warm start of the
 * next step's first solve (P:714); oracle_set_state alone would start it from zeros. */
void oracle_set_mnew(oracle_state *s, const double *Mnew)
{
    memcpy(s->Mnew, Mnew, sizeof(double) * (size_t)s->n);
}

void oracle_get_mnew(const oracle_state *s, double *Mnew)
{
    memcpy(Mnew, s->Mnew, sizeof(double) * (size_t)s->n);
}

void oracle_set_maxit(oracle_state *s, long long maxit)
{
    s->maxit = (maxit >= 0) ? maxit : 100LL * (long long)s->n;
}

/* P:700-742 */
int oracle_step(oracle_state *s, int *countIters, long long *nitSum,
                long long *nitMax, long long *nit_per_solve, double *diffs,
                int cap)
{
    const size_t bytes = sizeof(double) * (size_t)s->n;
    double diffC = 1.0, diffM = 1.0;                               /* P:700-701 */
    int count = 0;                                                 /* P:702 */
    long long sum = 0, mx = 0;
    memcpy(s->Cprev, s->C, bytes);                                 /* P:703 */
    memcpy(s->Mprev, s->M, bytes);                                 /* P:704 */

    while (diffC + diffM > s->p.eSoln) {                           /* P:706 */
        long long nit = s->maxit;                                  /* P:707 */
        double e1 = s->p.e1, e2 = s->p.e2;                         /* P:708-709 */
        int stop;
        oracle_gen_matrix(s->Mprev, s->C, s->A, s->ioff, s->rhs, s->row,
                          s->col, &s->p);                          /* P:712 */
        oracle_solve_ls_diag(s->n, 5, s->ioff, s->A, s->Mnew, s->rhs, &nit,
                             &e1, &e2, &stop, s->scratch, NULL, 0); /* P:714 */
        oracle_solve_c(s->Mprev, s->Mnew, s->Cprev, s->Cnew, s->n,
                       s->p.kappa, s->p.gama, s->p.tDel, s->p.gSelect); /* P:717 */
        if (nit > mx) mx = nit;                                    /* P:722 */
        sum += nit;                                                /* P:723 */
        diffC = oracle_calc_diff(s->C, s->Cnew, s->row, s->col);   /* P:725 */
        diffM = oracle_calc_diff(s->M, s->Mnew, s->row, s->col);   /* P:726 */
        memcpy(s->C, s->Cnew, bytes);                              /* P:728 */
        memcpy(s->M, s->Mnew, bytes);                              /* P:729 */
        if (nit_per_solve && count < cap) nit_per_solve[count] = nit;
        if (diffs && count < cap) {
            diffs[2 * count] = diffC;
            diffs[2 * count + 1] = diffM;
        }
        count = count + 1;                                         /* P:730 */
        if (count > 10000) {                                       /* P:732 */
            *countIters = count; *nitSum = sum; *nitMax = mx;
            return 1;
        }
    }
    *countIters = count; *nitSum = sum; *nitMax = mx;
    return 0;
}

long long oracle_first_solve_trace(oracle_state *s, oracle_cg_trace *trace,
                                   long trace_cap)
{
    long long nit = s->maxit;
    double e1 = s->p.e1, e2 = s->p.e2;
    int stop;
    /* first Picard iterate: Mprev = M, C = current C */
    oracle_gen_matrix(s->M, s->C, s->A, s->ioff, s->rhs, s->row, s->col,
                      &s->p);
    oracle_solve_ls_diag(s->n, 5, s->ioff, s->A, s->Mnew, s->rhs, &nit, &e1,
                         &e2, &stop, s->scratch, trace, trace_cap);
    return nit;
}
