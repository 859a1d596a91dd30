// main.cpp -- the executable the run scripts call as ./a.out
// (runAll.sh:7, twoDimension_runall.sh:7: `echo FILE.txt | ./a.out`).
// Same stdin / console / output-file contract as `program cThermoPDEODE`
// (PDE-ODEsolver.f90:20-175); the time step runs on the B200 through
// libpdeode_b200.so.  Environment (never the parameter file, whose format is
// fixed): PDEODE_DEVICE = CUDA device ordinal (default 0), PDEODE_SEED = seed
// of the random initial conditions (default: clock, like init_random_seed.f90),
// PDEODE_NGPUS = N > 1: the grid is split into N row slabs over devices
// PDEODE_DEVICE .. PDEODE_DEVICE+N-1 of this one process (pdeode_create_group).
#include <cstdlib>
#include <vector>

#include "pdeode_host.h"

namespace {
int g_ngpus = 1;
// Stepper::create with the group form behind it when PDEODE_NGPUS asks for several GPUs
int create_maybe_group(pdeode_ctx **ctx, int row, int col, const pdeode_params *p, int device) {
    if (g_ngpus <= 1) return pdeode_create(ctx, row, col, p, device);
    std::vector<int> dev((size_t)g_ngpus);
    for (int k = 0; k < g_ngpus; ++k) dev[(size_t)k] = device + k;
    return pdeode_create_group(ctx, row, col, p, g_ngpus, dev.data());
}
}  // namespace

int main() {
    pdeode_host::Stepper st{};
    if (const char *n = getenv("PDEODE_NGPUS")) g_ngpus = atoi(n) > 1 ? atoi(n) : 1;
    st.create = create_maybe_group;
    st.set_state = pdeode_set_state;
    st.get_state = pdeode_get_state;
    st.step = pdeode_step; st.step_many = pdeode_step_many;
    st.mass = pdeode_mass;
    st.peak = pdeode_peak;
    st.destroy = pdeode_destroy;
    st.strerror_ = pdeode_strerror;
    st.frame_size = pdeode_frame_size;
    st.get_frame = pdeode_get_frame;
    st.get_frame_rowmeans = pdeode_get_frame_rowmeans;
    st.set_initial_condition = pdeode_set_initial_condition;
    int device = 0;
    if (const char *d = getenv("PDEODE_DEVICE")) device = atoi(d);
    return pdeode_host::run_program(stdin, stdout, st, device, "");
}
