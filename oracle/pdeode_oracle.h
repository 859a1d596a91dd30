/*
 * pdeode_oracle.h -- CPU oracle for the implicit time step of
 * ericmjalbert/PDE-ODEsolver (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (pde-odesolver_b200/) never links, imports or executes it.
 *
 * PARITY PIN: no Fortran compiler exists in this image, so the gfortran binary
 * of runAll.sh:6 cannot be built.  Instead oracle/f90c/f90toc.py translates the
 * reference's three .f90 files, where they lie, statement by statement to C
 * (language rules only; it knows nothing about the algorithm), oracle/Makefile
 * compiles that with gcc -O3 into oracle/_ref/ (program + library), and
 * tests/test_reference_pin.py holds this restatement to it BIT FOR BIT:
 * every subroutine on random inputs, solveLSDIAG's iteration counts, whole
 * solveOrder2 runs, and the whole program's output files.  Vectors produced by
 * that build are committed as tests/golden/ref_*.npz (tests/test_golden_ref.py)
 * so that the pin travels to boxes without /root/reference.  What remains
 * unpinned is gfortran's own code generation (e.g. a different libgcc powi or
 * an FMA-contracting target), which no test here can see.
 *
 * "P:" = /root/reference/PDE-ODEsolver.f90, "CG:" = /root/reference/CGdiag.f90.
 * All reals are 8 bytes (reference build: gfortran -O3 -fdefault-real-8,
 * runAll.sh:6); no FMA contraction (x86-64 baseline has none), serial
 * summation order.
 */
#ifndef PDEODE_ORACLE_H
#define PDEODE_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Same layout as pdeode_params in include/pdeode_b200.h. */
typedef struct {
    double delta, nu, kappa, gama; /* P:61 */
    double tDel, xDel;             /* P:69; xDel = 1/row (P:110) */
    double e1, e2, eSoln;          /* P:69 */
    int alpha, beta;               /* P:62 (integer exponents) */
    int dSelect, fSelect, gSelect; /* P:64 */
    int reserved;
} oracle_params;

/* One record per CG iteration, for per-iteration parity checks. */
typedef struct {
    double rho, dot, alfa, err2, normx;
} oracle_cg_trace;

/* P:464-472 */
double oracle_dfunc(double M, double delta, int alpha, int beta, int dSelect);
/* P:445-458 */
double oracle_ffunc(double M, double C, double kappa, double nu, int fSelect);
/* libgcc __powidf2, what gfortran emits for real**integer-variable */
double oracle_powi(double x, int m);

/* P:517-590.  A is n x 5 column-major (diagonal-major), ioff[5], rhs[n]. */
void oracle_gen_matrix(const double *M, const double *C, double *A, int *ioff,
                       double *rhs, int row, int col, const oracle_params *p);

/* P:1044-1105 */
void oracle_amuxd(long n, const double *x, double *y, const double *diag,
                  int idiag, const int *ioff);

/* CG:99-114 */
double oracle_dotprod(long n, const double *u, const double *v);

/* CG:1-95.  nit: in = maxit, out = iterations taken.  trace may be NULL;
 * at most trace_cap records are written.  scratch = 3n doubles (r,p,q). */
void oracle_solve_ls_diag(long n, int ndiag, const int *ioff, const double *A,
                          double *sol, const double *rhs, long long *nit,
                          double *err1, double *err2, int *stopcrit,
                          double *scratch, oracle_cg_trace *trace,
                          long trace_cap);

/* P:482-511 */
void oracle_solve_c(const double *M, const double *Mnew, const double *Csol,
                    double *Cnew, long n, double kappa, double gama,
                    double tDel, int gSelect);

/* P:948-963 */
double oracle_calc_diff(const double *C, const double *Cnew, int row, int col);
/* P:772-791 */
double oracle_calc_mass(const double *X, long n);
/* P:978-1004.  Outputs are left untouched when the reference would not
 * assign them (intfac with no M > 0.1: SURVEY D8). */
void oracle_calc_peak_interface(const double *M, int row, int col,
                                double *peak, double *height, double *intfac);

/* ---- stepper state: the locals of solveOrder2, P:623-628 ---- */
typedef struct oracle_state oracle_state;

oracle_state *oracle_create(int row, int col, const oracle_params *p);
void oracle_destroy(oracle_state *s);
/* P:133-138: copy M, C in; Mnew (CG warm start) := 0. */
void oracle_set_state(oracle_state *s, const double *M, const double *C);
void oracle_get_state(const oracle_state *s, double *M, double *C);
/* Mnew as left by the last solve (the next warm start). */
void oracle_get_mnew(const oracle_state *s, double *Mnew);
void oracle_set_mnew(oracle_state *s, const double *Mnew);
/* Override the CG iteration cap (default 100*n, P:707).  <0 restores it. */
void oracle_set_maxit(oracle_state *s, long long maxit);

/* One pass of P:700-742.  Returns 0, or 1 when countIters exceeded 10000
 * (P:732-736).  nit_per_solve (may be NULL) receives the CG count of each
 * Picard iterate, at most cap entries; diffs (may be NULL) receives
 * diffC,diffM pairs. */
int oracle_step(oracle_state *s, int *countIters, long long *nitSum,
                long long *nitMax, long long *nit_per_solve, double *diffs,
                int cap);

/* First Picard iterate's operator + CG only, with a per-iteration trace;
 * used by the per-CG-iteration parity test.  Leaves the state untouched
 * except Mnew.  Returns iterations taken. */
long long oracle_first_solve_trace(oracle_state *s, oracle_cg_trace *trace,
                                   long trace_cap);

int oracle_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
