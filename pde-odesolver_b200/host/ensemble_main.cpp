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

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s PARAM1.txt [PARAM2.txt ...]\n", argv[0]);
        return 2;
    }
    pdeode_host::Stepper st{};
    st.create = pdeode_create; st.set_state = pdeode_set_state; st.get_state = pdeode_get_state;
    st.step = pdeode_step; st.step_many = pdeode_step_many; st.mass = pdeode_mass; st.peak = pdeode_peak;
    st.destroy = pdeode_destroy; st.strerror_ = pdeode_strerror;
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
            // the run scripts also copy the parameter file next to the results (runAll.sh:19)
            std::string cmd = "cp '" + file + "' '" + dir + "/' 2>/dev/null";
            if (system(cmd.c_str()) != 0) { /* best effort */ }
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
