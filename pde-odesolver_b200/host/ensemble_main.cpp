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
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <cerrno>
#include <dirent.h>
#include <new>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>
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

    // The GPUs to use.  Counted without touching CUDA where possible (the process-per-GPU mode
    // below must fork before the first CUDA call): PDEODE_DEVICES = "0,1,..." (ordinals as CUDA
    // numbers them for this process), else every entry of CUDA_VISIBLE_DEVICES, else every
    // GPU the driver lists under /proc/driver/nvidia/gpus.
    std::vector<int> devices;
    auto split_ints = [](const std::string &s2, std::vector<int> &out) {
        size_t pos = 0;
        while (pos < s2.size()) {
            out.push_back(atoi(s2.c_str() + pos));
            pos = s2.find(',', pos);
            if (pos == std::string::npos) break;
            ++pos;
        }
    };
    std::vector<std::string> visible;           // entries of CUDA_VISIBLE_DEVICES, if it is set
    if (const char *v = getenv("CUDA_VISIBLE_DEVICES")) {
        std::string s2(v);
        size_t pos = 0;
        while (pos <= s2.size()) {
            const size_t c = s2.find(',', pos);
            visible.push_back(s2.substr(pos, c == std::string::npos ? std::string::npos : c - pos));
            if (c == std::string::npos) break;
            pos = c + 1;
        }
        if (visible.size() == 1 && visible[0].empty()) visible.clear();
    }
    bool counted_without_cuda = true;
    if (const char *d = getenv("PDEODE_DEVICES")) {
        split_ints(d, devices);
    } else if (!visible.empty()) {
        for (size_t i = 0; i < visible.size(); ++i) devices.push_back((int)i);
    } else {
        int n = 0;
        if (DIR *dp = opendir("/proc/driver/nvidia/gpus")) {
            while (struct dirent *e = readdir(dp)) n += e->d_name[0] != '.';
            closedir(dp);
        }
        if (n <= 0) {
            counted_without_cuda = false;
            if (pdeode_device_count(&n) != PDEODE_OK || n <= 0) {
                fprintf(stderr, "pdeode_ensemble: no CUDA device (there is no CPU fallback)\n");
                return 1;
            }
        }
        for (int i = 0; i < n; ++i) devices.push_back(i);
    }
    if (devices.empty()) { fprintf(stderr, "pdeode_ensemble: empty device list\n"); return 1; }
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
    // Results and the member counter live in memory shared with the per-GPU child processes.
    struct Shared { std::atomic<int> next; };
    const size_t shbytes = sizeof(Shared) + (size_t)nmember * (sizeof(int) + 2 * sizeof(double)) + 64;
    void *shm = mmap(nullptr, shbytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (shm == MAP_FAILED) { perror("pdeode_ensemble: mmap"); return 1; }
    Shared *sh = new (shm) Shared();
    sh->next.store(0);
    double *began = reinterpret_cast<double *>(static_cast<char *>(shm) + 64);
    double *first_done = began + nmember;
    int *rc = reinterpret_cast<int *>(first_done + nmember);
    for (int k = 0; k < nmember; ++k) { rc[k] = -1; began[k] = 0.0; first_done[k] = 0.0; }
    // PDEODE_ENSEMBLE_TIMING=1: where the wall time of a sweep goes (stderr)
    const bool timing = getenv("PDEODE_ENSEMBLE_TIMING") && getenv("PDEODE_ENSEMBLE_TIMING")[0] == '1';
    const auto t_start = std::chrono::steady_clock::now();      // CLOCK_MONOTONIC: the same in the children
    auto since = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count(); };
    auto worker = [&](int device) {
        for (;;) {
            const int k = sh->next.fetch_add(1);
            if (k >= nmember) return;
            began[k] = since();
            const std::string file = argv[1 + k];
            std::string stem = file.substr(0, file.find('.'));          // `cut -d '.' -f 1`, runAll.sh:3
            const std::string dir = stem + "Output";
            mkdir(dir.c_str(), 0777);
            FILE *in = tmpfile();
            FILE *con = fopen((dir + "/console.txt").c_str(), "w");
            if (!in || !con) { rc[k] = 3; if (in) fclose(in); if (con) fclose(con); continue; }
            fprintf(in, "%s\n", file.c_str());
            rewind(in);
            rc[k] = pdeode_host::run_program(in, con, st, device, dir);
            fclose(in);
            fclose(con);
            first_done[k] = since();
            // the run scripts also copy the parameter file next to the results (runAll.sh:19);
            // not through system(): forking a process with 128 threads and eight GPUs mapped
            // costs milliseconds apiece
            const size_t slash = file.find_last_of('/');
            copy_file(file, dir + "/" + (slash == std::string::npos ? file : file.substr(slash + 1)));
        }
    };
    // One process per GPU (default with several GPUs; PDEODE_ENSEMBLE_FORK=0 keeps everything in
    // this process): a process that sees eight GPUs spends seconds in the driver bringing all
    // of them up one after the other before its first kernel runs, and as long again tearing
    // them down (measured on an 8 x B200 box: first member done after 6.3 s of an 8.2 s sweep);
    // children that each see one GPU do that side by side.  Members are still dealt from one
    // counter, so a GPU that finishes early takes the next file.
    const char *fk = getenv("PDEODE_ENSEMBLE_FORK");
    const bool fork_mode = devices.size() > 1 && counted_without_cuda && !(fk && fk[0] == '0');
    double t_spawned = 0.0;
    if (fork_mode) {
        fflush(nullptr);
        std::vector<pid_t> kids;
        for (size_t d = 0; d < devices.size(); ++d) {
            const pid_t pid = fork();
            if (pid < 0) { perror("pdeode_ensemble: fork"); break; }
            if (pid == 0) {
                const int dev = devices[d];
                const std::string only = !visible.empty()
                    ? ((size_t)dev < visible.size() ? visible[(size_t)dev] : std::string("-1"))
                    : std::to_string(dev);
                setenv("CUDA_VISIBLE_DEVICES", only.c_str(), 1);
                std::vector<std::thread> pool;
                for (int s2 = 0; s2 < per_gpu; ++s2) pool.emplace_back(worker, 0);
                for (auto &t : pool) t.join();
                fflush(nullptr);
                _exit(0);
            }
            kids.push_back(pid);
        }
        t_spawned = since();
        for (pid_t pid : kids) {
            int status = 0;
            while (waitpid(pid, &status, 0) < 0 && errno == EINTR) {}
        }
        // files no child got to (a child that died early): run them here
        if (sh->next.load() < nmember) worker(devices[0]);
    } else {
        std::vector<std::thread> pool;
        for (size_t d = 0; d < devices.size(); ++d)
            for (int s2 = 0; s2 < per_gpu; ++s2) pool.emplace_back(worker, devices[d]);
        t_spawned = since();
        for (auto &t : pool) t.join();
    }
    if (timing) {
        double first = 1e30, last = 0, longest = 0, b0 = 1e30, b1 = 0;
        for (int k = 0; k < nmember; ++k) {
            first = std::min(first, first_done[k]); last = std::max(last, first_done[k]);
            longest = std::max(longest, first_done[k] - began[k]);
            b0 = std::min(b0, began[k]); b1 = std::max(b1, began[k]);
        }
        fprintf(stderr, "pdeode_ensemble timing (%s): workers started %.3f s, members began %.3f .. %.3f s, first "
                        "member done %.3f s, last member done %.3f s, longest member %.3f s, workers finished "
                        "(contexts destroyed) %.3f s\n", fork_mode ? "one process per GPU" : "one process",
                t_spawned, b0, b1, first, last, longest, since());
    }
    int bad = 0;
    for (int k = 0; k < nmember; ++k) {
        printf("%s -> %s\n", argv[1 + k], rc[k] == 0 ? "ok" : "FAILED");
        bad += rc[k] != 0;
    }
    // Every member's files are closed and every context destroyed by now.  Leave without the
    // runtime's exit handlers: unloading the modules and unmapping the address space of eight
    // primary contexts one by one took 8 of the 11 s a 256-member sweep spent in this program
    // (measured), and the kernel reclaims all of it when the process ends anyway.
    fflush(nullptr);
    if (!(getenv("PDEODE_ENSEMBLE_SLOW_EXIT") && getenv("PDEODE_ENSEMBLE_SLOW_EXIT")[0] == '1'))
        _exit(bad ? 1 : 0);
    return bad ? 1 : 0;
}
