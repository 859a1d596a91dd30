// ensemble_main.cpp -- parameter-sweep driver (BASELINE config 5).
//
// The reference has no sweep driver: runAll_paral.sh is the same single run
// built with OpenMP; its .gitignore ("Extra Parameter Files
// [1-9]*.txt", :34-35) shows how the author swept: numbered parameter files,
// one output directory each.  This program does that in one process:
//
//     pdeode_ensemble 1.txt 2.txt ... N.txt
//
// Every file is an independent run of `program cThermoPDEODE` (same console
// text, same output files) written into <name>Output/ like the run scripts'
// DIRNAME (runAll.sh:3-4); the console goes to <name>Output/console.txt.
// Members are replicas: no communication between them.  They are dealt
// round-robin over the visible GPUs (PDEODE_DEVICES = "0,1,..." to restrict)
// and PDEODE_MEMBERS_PER_GPU (default 16) of them run concurrently per GPU,
// each on its own context and stream, so that small grids fill the device.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <vector>

#include "pdeode_host.h"

namespace {

// One context per worker thread, kept from member to member: creating and destroying a context
// is some 25 allocations plus their cudaFree (each a device-wide synchronisation while the
// other members' kernels are running) behind the driver's process-wide lock -- with 128 worker
// threads on 8 GPUs that was 50 ms per member, 12 of the 13 s a 256-member sweep took.  A
// member of the same grid re-uses the thread's context with its own parameters
// (pdeode_reconfigure); another grid size gets a new one.
struct Cached {
    pdeode_ctx *ctx = nullptr;
    int row = 0, col = 0, device = -1;
    ~Cached() { if (ctx) pdeode_destroy(ctx); }
};
thread_local Cached t_cache;

int pooled_create(pdeode_ctx **out, int row, int col, const pdeode_params *p, int device) {
    Cached &c = t_cache;
    if (c.ctx && c.row == row && c.col == col && c.device == device) {
        const int rc = pdeode_reconfigure(c.ctx, p);
        *out = c.ctx;
        return rc;
    }
    if (c.ctx) { pdeode_destroy(c.ctx); c.ctx = nullptr; }
    const int rc = pdeode_create(out, row, col, p, device);
    if (rc == PDEODE_OK) { c.ctx = *out; c.row = row; c.col = col; c.device = device; }
    return rc;                      // on failure the caller destroys *out itself (not cached)
}

int pooled_destroy(pdeode_ctx *ctx) {
    if (ctx && ctx == t_cache.ctx) return PDEODE_OK;      // stays with the thread
    return pdeode_destroy(ctx);
}

bool copy_file(const std::string &from, const std::string &to) {
    FILE *a = fopen(from.c_str(), "rb");
    if (!a) return false;
    FILE *b = fopen(to.c_str(), "wb");
    if (!b) { fclose(a); return false; }
    char buf[1 << 14];
    size_t k;
    while ((k = fread(buf, 1, sizeof buf, a)) > 0) fwrite(buf, 1, k, b);
    fclose(a);
    return fclose(b) == 0;
}

}  // namespace

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s PARAM1.txt [PARAM2.txt ...]\n", argv[0]);
        return 2;
    }
    pdeode_host::Stepper st{};
    const bool pooled = !(getenv("PDEODE_ENSEMBLE_POOL") && getenv("PDEODE_ENSEMBLE_POOL")[0] == '0');
    st.create = pooled ? pooled_create : pdeode_create; st.set_state = pdeode_set_state; st.get_state = pdeode_get_state;
    st.step = pdeode_step; st.step_many = pdeode_step_many; st.mass = pdeode_mass; st.peak = pdeode_peak;
    st.destroy = pooled ? pooled_destroy : pdeode_destroy; st.strerror_ = pdeode_strerror;
    st.frame_size = pdeode_frame_size; st.get_frame = pdeode_get_frame;
    st.get_frame_rowmeans = pdeode_get_frame_rowmeans;
    st.set_initial_condition = pdeode_set_initial_condition;

    std::vector<int> devices;
    if (const char *d = getenv("PDEODE_DEVICES")) {
        std::string s(d);
        size_t pos = 0;
        while (pos < s.size()) {
            devices.push_back(atoi(s.c_str() + pos));
            pos = s.find(',', pos);
            if (pos == std::string::npos) break;
            ++pos;
        }
    } else {
        int n = 0;
        if (pdeode_device_count(&n) != PDEODE_OK || n <= 0) {
            fprintf(stderr, "pdeode_ensemble: no CUDA device (there is no CPU fallback)\n");
            return 1;
        }
        for (int i = 0; i < n; ++i) devices.push_back(i);
    }
    int per_gpu = 16;    // measured best for 512^2 members on a B200 (profiles/README.md)
    if (const char *m = getenv("PDEODE_MEMBERS_PER_GPU")) per_gpu = atoi(m) > 0 ? atoi(m) : 1;

    // members share a GPU: each member's cooperative launches (the whole-step kernel, the
    // persistent solve) are capped to 64 CTAs, so that at least two members' grids are
    // resident side by side -- one fills the other's barrier waits -- and the rest take turns.
    // Measured with 32 members of 512^2, 16 in flight (profiles/README.md): caps of 37 to 74
    // CTAs all take 2.2-2.4 s, 24 CTAs 4.5 s, 92 CTAs (one member resident at a time) 6.5 s.
    if (per_gpu > 1 && !getenv("PDEODE_PERSIST_MAX_CTAS") && !getenv("PDEODE_GPU_SHARE"))
        setenv("PDEODE_PERSIST_MAX_CTAS", "64", 1);
    // more member threads than cores is normal here: waiting threads give their core away
    if (!getenv("PDEODE_STEP_YIELD")) setenv("PDEODE_STEP_YIELD", "1", 1);

    const int nmember = argc - 1;
    std::vector<int> rc((size_t)nmember, -1);
    std::atomic<int> next{0};
    auto worker = [&](int device) {
        for (;;) {
            const int k = next.fetch_add(1);
            if (k >= nmember) return;
            const std::string file = argv[1 + k];
            std::string stem = file.substr(0, file.find('.'));          // `cut -d '.' -f 1`, runAll.sh:3
            const std::string dir = stem + "Output";
            mkdir(dir.c_str(), 0777);
            FILE *in = tmpfile();
            FILE *con = fopen((dir + "/console.txt").c_str(), "w");
            if (!in || !con) { rc[(size_t)k] = 3; if (in) fclose(in); if (con) fclose(con); continue; }
            fprintf(in, "%s\n", file.c_str());
            rewind(in);
            rc[(size_t)k] = pdeode_host::run_program(in, con, st, device, dir);
            fclose(in);
            fclose(con);
            // the run scripts also copy the parameter file next to the results (runAll.sh:19);
            // not through system(): forking a process with 128 threads and eight GPUs mapped
            // costs milliseconds apiece
            const size_t slash = file.find_last_of('/');
            copy_file(file, dir + "/" + (slash == std::string::npos ? file : file.substr(slash + 1)));
        }
    };
    std::vector<std::thread> pool;
    for (size_t d = 0; d < devices.size(); ++d)
        for (int s = 0; s < per_gpu; ++s) pool.emplace_back(worker, devices[d]);
    for (auto &t : pool) t.join();
    int bad = 0;
    for (int k = 0; k < nmember; ++k) {
        printf("%s -> %s\n", argv[1 + k], rc[(size_t)k] == 0 ? "ok" : "FAILED");
        bad += rc[(size_t)k] != 0;
    }
    return bad ? 1 : 0;
}
