// pdeode_host.h -- host program of the B200 build: everything of
// `program cThermoPDEODE` (PDE-ODEsolver.f90:20-175) that is NOT the time
// step's inner machinery: parameter file, initial conditions, the time loop's
// cadences, console lines and output files.  The time step itself is reached
// only through the C ABI of include/pdeode_b200.h, via a table of function
// pointers so that tests can drive the same host logic with another stepper.
//
// The reference's host is Fortran; this image has no Fortran compiler, so the
// host side is restated in C++ (names and argument meaning follow the
// reference's subroutines).  fortran/ holds the bind(C) route for a Fortran host.
//
// "P:" = PDE-ODEsolver.f90 of the reference.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/pdeode_b200.h"

namespace pdeode_host {

// The 21 values of parameter.txt:4-29, in file order (P:184-250).
struct RunParams {
    int pSize = 0, true2D = 0;                       // P:193-194
    double depth = 0, height = 0;                    // P:225-226
    int numInnocu = 0, alpha = 0, beta = 0;          // P:227-229
    double nu = 0, kappa = 0, gama = 0, delta = 0;   // P:230-233
    int nOuts = 0;                                   // P:234
    double tEnd = 0, tDel = 0;                       // P:235-236
    double e1 = 0, e2 = 0, eSoln = 0;                // P:238-240
    int fSelect = 0, dSelect = 0, gSelect = 0, MinitialCond = 0;  // P:243-246
    // derived by the program, not read from the file
    int row = 0, col = 0;                            // P:91-99
    long long n = 0;                                 // P:100
    double xDel = 0;                                 // P:110 (the file's xDel line is skipped, P:237)
    long long nit = 0;                               // P:111 (the file's nit line is skipped, P:241)
};

// getProbSize + paramSet (P:184-250): strictly positional, names ignored.
// Returns false and fills err when the file cannot be read that way.
bool read_parameter_file(const std::string &path, RunParams &p, std::string &err);

// setInitialConditions (P:262-439).  seed drives the clock-seeded cases
// (MinitialCond 4, 5, 7, 10; init_random_seed.f90:11-14); the uninitialised
// accumulator `total` of P:272,347 is taken as 0.
void set_initial_conditions(std::vector<double> &C, std::vector<double> &M, const RunParams &p,
                            uint64_t seed);

// gfortran list-directed rendering of one REAL(8) / default INTEGER item
// (leading separator blank included), used by every `write(u,*)` of the
// reference that lands in a file or on the console.
std::string fmt_real(double x);
std::string fmt_int(long long v);

// printToFile (P:810-849) / printToFile2D (P:871-933): append one frame.
void print_to_file(FILE *f, const RunParams &p, const double *M, const double *C);
void print_to_file_2d(FILE *f, const RunParams &p, const double *M, const double *C);
// The same frames from data decimated on the device (pdeode_get_frame*).
void print_frame(FILE *f, const RunParams &p, int nro, int nco, const double *Md, const double *Cd);
// Steps the time loop takes from `counter` on before it looks at the state again (P:663,665,683).
int steps_until_next_look(long long counter, int filter, double tDel, double tEnd);
void print_frame_2d(FILE *f, const RunParams &p, int nro, const double *meanM, const double *meanC);
// reportStats (P:1019-1040)
bool report_stats(const std::string &path, double avgIters, double maxIters, double avgNit,
                  double maxNit, double seconds);

// The time step behind function pointers: exactly the C ABI's entry points.
struct Stepper {
    int (*create)(pdeode_ctx **, int, int, const pdeode_params *, int);
    int (*set_state)(pdeode_ctx *, const double *, const double *);
    int (*get_state)(pdeode_ctx *, double *, double *);
    int (*step)(pdeode_ctx *, int *, long long *, long long *);
    int (*mass)(pdeode_ctx *, double *, double *);
    int (*peak)(pdeode_ctx *, double *, double *, double *);
    int (*destroy)(pdeode_ctx *);
    const char *(*strerror_)(int);
    // optional (may be null): frame data decimated on the device; without them the
    // host downloads the whole fields at frame times
    int (*frame_size)(const pdeode_ctx *, int, int *, int *);
    int (*get_frame)(pdeode_ctx *, int, double *, double *);
    int (*get_frame_rowmeans)(pdeode_ctx *, int, double *, double *);
    // optional: build the initial fields on the device instead of uploading them
    int (*set_initial_condition)(pdeode_ctx *, int, double, double, int, const double *,
                                 const double *);
    // optional: several time steps with one host wait (pdeode_step_many); without it the
    // loop calls step once per time step
    int (*step_many)(pdeode_ctx *, int, int *, long long *, int *, long long *, long long *);
};

// What pdeode_set_initial_condition needs: the case and its inoculation points.
struct DeviceIc {
    int ic = 0;
    double depth = 0, height = 0;
    std::vector<double> px, py;
};
bool inoculation_points(const RunParams &p, uint64_t seed, std::vector<double> &px,
                        std::vector<double> &py);

struct RunStats {
    double avgIters = 0, maxIters = 0, avgNit = 0, maxNit = 0;   // P:76-77
    double seconds = 0;                                          // P:157
    long long steps = 0, frames = 0;
    int status = 0;      // 0, or PDEODE_ERR_* (PDEODE_ERR_PICARD_CAP = the `stop` of P:735)
};

// solveOrder2 (P:600-763) with the step body replaced by stepper calls.
// Files are written into outdir ("" = cwd); console goes to `con`.
RunStats solve_order2(const RunParams &p, std::vector<double> &M, std::vector<double> &C,
                      const Stepper &st, int device, const std::string &outdir, FILE *con,
                      const DeviceIc *dic = nullptr);

// program cThermoPDEODE (P:20-175): prompts for the parameter file name on
// `in`, echoes the setup, runs, reports.  Returns the process exit code.
int run_program(FILE *in, FILE *con, const Stepper &st, int device, const std::string &outdir);

}  // namespace pdeode_host
