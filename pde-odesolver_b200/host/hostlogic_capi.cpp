// hostlogic_capi.cpp -- C entry points of libpdeode_hostlogic.so: the host
// logic (parameter file, initial conditions, writers, time loop) without any
// stepper linked in, so that it can be exercised on a machine with no GPU.
// The stepper arrives as a table of function pointers (pdeode_host::Stepper).
#include <cstdio>
#include <cstring>

#include "pdeode_host.h"

using namespace pdeode_host;

extern "C" {

// Flat mirror of RunParams for ctypes.
struct pdeode_host_params {
    int pSize, true2D, numInnocu, alpha, beta, nOuts, fSelect, dSelect, gSelect, MinitialCond;
    int row, col;
    long long n, nit;
    double depth, height, nu, kappa, gama, delta, tEnd, tDel, e1, e2, eSoln, xDel;
};

static void flatten(const RunParams &p, pdeode_host_params *o) {
    o->pSize = p.pSize; o->true2D = p.true2D; o->numInnocu = p.numInnocu; o->alpha = p.alpha;
    o->beta = p.beta; o->nOuts = p.nOuts; o->fSelect = p.fSelect; o->dSelect = p.dSelect;
    o->gSelect = p.gSelect; o->MinitialCond = p.MinitialCond; o->row = p.row; o->col = p.col;
    o->n = p.n; o->nit = p.nit; o->depth = p.depth; o->height = p.height; o->nu = p.nu;
    o->kappa = p.kappa; o->gama = p.gama; o->delta = p.delta; o->tEnd = p.tEnd; o->tDel = p.tDel;
    o->e1 = p.e1; o->e2 = p.e2; o->eSoln = p.eSoln; o->xDel = p.xDel;
}

static RunParams unflatten(const pdeode_host_params *o) {
    RunParams p;
    p.pSize = o->pSize; p.true2D = o->true2D; p.numInnocu = o->numInnocu; p.alpha = o->alpha;
    p.beta = o->beta; p.nOuts = o->nOuts; p.fSelect = o->fSelect; p.dSelect = o->dSelect;
    p.gSelect = o->gSelect; p.MinitialCond = o->MinitialCond; p.row = o->row; p.col = o->col;
    p.n = o->n; p.nit = o->nit; p.depth = o->depth; p.height = o->height; p.nu = o->nu;
    p.kappa = o->kappa; p.gama = o->gama; p.delta = o->delta; p.tEnd = o->tEnd; p.tDel = o->tDel;
    p.e1 = o->e1; p.e2 = o->e2; p.eSoln = o->eSoln; p.xDel = o->xDel;
    return p;
}

int pdeode_host_read_parameters(const char *path, pdeode_host_params *out, char *err, int errlen) {
    RunParams p;
    std::string e;
    if (!read_parameter_file(path, p, e)) {
        if (err && errlen > 0) { strncpy(err, e.c_str(), errlen - 1); err[errlen - 1] = 0; }
        return 1;
    }
    flatten(p, out);
    return 0;
}

int pdeode_host_initial_conditions(const pdeode_host_params *pp, unsigned long long seed,
                                   double *M, double *C) {
    RunParams p = unflatten(pp);
    std::vector<double> m, c;
    set_initial_conditions(c, m, p, seed);
    memcpy(M, m.data(), sizeof(double) * m.size());
    memcpy(C, c.data(), sizeof(double) * c.size());
    return 0;
}

// The inoculation points set_initial_conditions uses for this case and seed
// (what the host hands to pdeode_set_initial_condition).  Returns the count.
int pdeode_host_inoculation_points(const pdeode_host_params *pp, unsigned long long seed,
                                   double *px, double *py, int cap) {
    RunParams p = unflatten(pp);
    std::vector<double> x, y;
    inoculation_points(p, seed, x, y);
    const int n = (int)x.size() < cap ? (int)x.size() : cap;
    for (int k = 0; k < n; ++k) { px[k] = x[(size_t)k]; py[k] = y[(size_t)k]; }
    return (int)x.size();
}

int pdeode_host_format_real(double x, char *out, int len) {
    const std::string s = fmt_real(x);
    strncpy(out, s.c_str(), len - 1);
    out[len - 1] = 0;
    return (int)s.size();
}

int pdeode_host_write_frame(const char *path, const pdeode_host_params *pp, const double *M,
                            const double *C, int append) {
    RunParams p = unflatten(pp);
    FILE *f = fopen(path, append ? "a" : "w");
    if (!f) return 1;
    if (p.true2D == 1) print_to_file_2d(f, p, M, C);
    else print_to_file(f, p, M, C);
    fclose(f);
    return 0;
}

// The writer the GPU program uses: frame data already decimated (what
// pdeode_get_frame returns).  Decimates M, C here the way the device does.
int pdeode_host_write_frame_decimated(const char *path, const pdeode_host_params *pp,
                                      const double *M, const double *C, int append) {
    RunParams p = unflatten(pp);
    if (p.true2D == 1) return 2;
    const int filter = (p.row > 257) ? (p.row - 1) / 256 : 1;                    // P:829-832
    const int nro = (p.row - 1) / filter + 1, nco = (p.col - 1) / filter + 1;
    std::vector<double> Md((size_t)nro * nco), Cd((size_t)nro * nco);
    for (int oi = 0; oi < nro; ++oi)
        for (int oj = 0; oj < nco; ++oj) {
            const size_t q = (size_t)(oi * filter) * p.col + (size_t)(oj * filter);
            Md[(size_t)oi * nco + oj] = M[q];
            Cd[(size_t)oi * nco + oj] = C[q];
        }
    FILE *f = fopen(path, append ? "a" : "w");
    if (!f) return 1;
    print_frame(f, p, nro, nco, Md.data(), Cd.data());
    fclose(f);
    return 0;
}

int pdeode_host_steps_until_next_look(long long counter, int filter, double tDel, double tEnd) {
    return steps_until_next_look(counter, filter, tDel, tEnd);
}

// Whole program (P:20-175) with the caller's stepper; param_file name is fed
// to the prompt as stdin would be.  console_path receives the console text.
int pdeode_host_run(const char *param_file, const char *outdir, const char *console_path,
                    const Stepper *st, int device) {
    FILE *in = tmpfile();
    if (!in) return 3;
    fprintf(in, "%s\n", param_file);
    rewind(in);
    FILE *con = fopen(console_path, "w");
    if (!con) { fclose(in); return 3; }
    const int rc = run_program(in, con, *st, device, outdir ? outdir : "");
    fclose(con);
    fclose(in);
    return rc;
}

}  // extern "C"
