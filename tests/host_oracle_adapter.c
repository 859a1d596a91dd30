/*
 * host_oracle_adapter.c -- TEST INFRASTRUCTURE.  Implements the stepper table
 * the host program expects (pdeode_host::Stepper: the C ABI's entry points)
 * on top of the CPU oracle, so that the host logic (parameter file, cadences,
 * console, output files) can be exercised end to end on a machine without a
 * GPU.  Never linked into the product.
 */
#include <stdlib.h>

#include "../include/pdeode_b200.h"
#include "../oracle/pdeode_oracle.h"

struct pdeode_ctx {
    oracle_state *o;
    int row, col;
    long n;
    double *M, *C;
};

static int a_create(pdeode_ctx **out, int row, int col, const pdeode_params *p, int device)
{
    (void)device;
    pdeode_ctx *c = (pdeode_ctx *)calloc(1, sizeof(*c));
    if (!c) return PDEODE_ERR_NOMEM;
    c->o = oracle_create(row, col, (const oracle_params *)p);   /* identical layout */
    c->row = row; c->col = col; c->n = (long)row * col;
    c->M = (double *)malloc(sizeof(double) * c->n);
    c->C = (double *)malloc(sizeof(double) * c->n);
    if (!c->o || !c->M || !c->C) return PDEODE_ERR_NOMEM;
    *out = c;
    return PDEODE_OK;
}
static int a_set_state(pdeode_ctx *c, const double *M, const double *C)
{
    /* ORACLE_PERTURB=eps: start from M*(1 +- eps), a pseudo-random sign per cell -- the size of
     * rounding differences between two builders of the same initial condition.  Used to see how
     * far two runs of the reference's own algorithm drift apart from such a start. */
    const char *e = getenv("ORACLE_PERTURB");
    if (e && atof(e) != 0.0) {
        const double eps = atof(e);
        const long n = (long)c->row * c->col;
        double *Mp = (double *)malloc(sizeof(double) * (size_t)n);
        if (!Mp) return PDEODE_ERR_NOMEM;
        unsigned int h = 12345u;
        for (long i = 0; i < n; ++i) {
            h = h * 1664525u + 1013904223u;
            Mp[i] = M[i] * (1.0 + ((h >> 16) & 1u ? eps : -eps));
        }
        oracle_set_state(c->o, Mp, C);
        free(Mp);
        return PDEODE_OK;
    }
    oracle_set_state(c->o, M, C);
    return PDEODE_OK;
}
static int a_get_state(pdeode_ctx *c, double *M, double *C)
{
    oracle_get_state(c->o, M, C);
    return PDEODE_OK;
}
static int a_step(pdeode_ctx *c, int *cnt, long long *sum, long long *mx)
{
    return oracle_step(c->o, cnt, sum, mx, NULL, NULL, 0) ? PDEODE_ERR_PICARD_CAP : PDEODE_OK;
}
static int a_mass(pdeode_ctx *c, double *mM, double *mC)
{
    oracle_get_state(c->o, c->M, c->C);
    *mM = oracle_calc_mass(c->M, c->n);
    *mC = oracle_calc_mass(c->C, c->n);
    return PDEODE_OK;
}
static int a_peak(pdeode_ctx *c, double *peak, double *height, double *intfac)
{
    oracle_get_state(c->o, c->M, c->C);
    oracle_calc_peak_interface(c->M, c->row, c->col, peak, height, intfac);
    return PDEODE_OK;
}
static int a_destroy(pdeode_ctx *c)
{
    if (!c) return PDEODE_OK;
    oracle_destroy(c->o);
    free(c->M); free(c->C); free(c);
    return PDEODE_OK;
}
static const char *a_strerror(int code) { return code ? "oracle adapter error" : "ok"; }

typedef struct {
    int (*create)(pdeode_ctx **, int, int, const pdeode_params *, int);
    int (*set_state)(pdeode_ctx *, const double *, const double *);
    int (*get_state)(pdeode_ctx *, double *, double *);
    int (*step)(pdeode_ctx *, int *, long long *, long long *);
    int (*mass)(pdeode_ctx *, double *, double *);
    int (*peak)(pdeode_ctx *, double *, double *, double *);
    int (*destroy)(pdeode_ctx *);
    const char *(*strerror_)(int);
    /* optional device-side frame decimation: absent here, so the host takes its
     * whole-field writer path -- the GPU run takes the other, and the two are compared */
    void *frame_size, *get_frame, *get_frame_rowmeans;
    void *set_initial_condition;    /* absent: the host builds the fields itself */
    void *step_many;                /* absent: one step call per time step */
} stepper_table;

static const stepper_table TABLE = {a_create, a_set_state, a_get_state, a_step,
                                    a_mass,   a_peak,      a_destroy,   a_strerror,
                                    0,        0,           0,           0,
                                    0};

const void *oracle_stepper_table(void) { return &TABLE; }

#ifdef ADAPTER_EXPORT_ABI
/* The same functions under the C ABI's own names, so that a host program LINKED against
 * libpdeode_b200.so's interface (the reference's Fortran host patched by
 * pde-odesolver_b200/fortran/patch_reference.py, tests/test_fortran_host.py) can be run on a
 * machine without a GPU.  A separate test-only library; never loaded next to the real one. */
int pdeode_create(pdeode_ctx **c, int row, int col, const pdeode_params *p, int dev) { return a_create(c, row, col, p, dev); }
int pdeode_set_state(pdeode_ctx *c, const double *M, const double *C) { return a_set_state(c, M, C); }
int pdeode_get_state(pdeode_ctx *c, double *M, double *C) { return a_get_state(c, M, C); }
int pdeode_step(pdeode_ctx *c, int *cnt, long long *sum, long long *mx) { return a_step(c, cnt, sum, mx); }
int pdeode_mass(pdeode_ctx *c, double *mM, double *mC) { return a_mass(c, mM, mC); }
int pdeode_peak(pdeode_ctx *c, double *peak, double *height, double *intfac) { return a_peak(c, peak, height, intfac); }
int pdeode_destroy(pdeode_ctx *c) { return a_destroy(c); }
#endif
