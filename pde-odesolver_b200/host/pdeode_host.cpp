// pdeode_host.cpp -- see pdeode_host.h.  Host-side restatement of the parts of
// `program cThermoPDEODE` that stay on the host.  "P:" = PDE-ODEsolver.f90.
#include "pdeode_host.h"

#include <charconv>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace pdeode_host {

// ---------------------------------------------------------------- formatting

// gfortran list-directed REAL(8): one separator blank, then G25.17E3 --
// F20.(17-k) followed by five blanks when 0.1 <= |x| < 1e17 (k = digits before
// the point), else 1PE25.16E3.
// Writes the 26 characters of one list-directed REAL(8) item into out (no NUL).
// One conversion per number: the 17 significant digits of "%.16E" are re-pointed
// for the F form instead of being converted a second time.
static void fmt_real_to(char *out, double x) {
    memset(out, ' ', 26);
    if (std::isnan(x)) { memcpy(out + 23, "NaN", 3); return; }
    if (std::isinf(x)) {
        if (x > 0) memcpy(out + 18, "Infinity", 8); else memcpy(out + 17, "-Infinity", 9);
        return;
    }
    if (x == 0.0) {                               // F20.16 + 5 blanks
        const bool neg = std::signbit(x);
        memcpy(out + 3, "0.0000000000000000", 18);
        if (neg) out[2] = '-';
        return;
    }
    // [-]d.dddddddddddddddde[+-]XX, correctly rounded to 17 digits; the same digits as
    // printf("%.16E") at a third of the cost
    char e[48];
    char *end = std::to_chars(e, e + sizeof e - 1, x, std::chars_format::scientific, 16).ptr;
    *end = '\0';
    const bool neg = (e[0] == '-');
    const char *m = e + (neg ? 1 : 0);            // d.dddd...
    const char *ep = strchr(m, 'e');
    const int ex = atoi(ep + 1);
    const int k = ex + 1;                         // digits before the point in F form
    char dig[17];
    dig[0] = m[0];
    memcpy(dig + 1, m + 2, 16);
    if (k >= 0 && k <= 17) {
        // F20.(17-k) right-justified in 20 columns after the separator blank, then 5 blanks
        char body[24];
        int n = 0;
        if (neg) body[n++] = '-';
        if (k == 0) { body[n++] = '0'; body[n++] = '.'; memcpy(body + n, dig, 17); n += 17; }
        else { memcpy(body + n, dig, k); n += k; body[n++] = '.'; memcpy(body + n, dig + k, 17 - k); n += 17 - k; }
        memcpy(out + 1 + (20 - n), body, n);
        return;
    }
    // 1PE25.16E3
    char body[28];
    int n = 0;
    if (neg) body[n++] = '-';
    memcpy(body + n, m, 18); n += 18;             // d.dddddddddddddddd
    n += snprintf(body + n, sizeof body - n, "E%c%03d", ex < 0 ? '-' : '+', ex < 0 ? -ex : ex);
    memcpy(out + 1 + (25 - n), body, n);
}

std::string fmt_real(double x) {
    char buf[26];
    fmt_real_to(buf, x);
    return std::string(buf, 26);
}

std::string fmt_int(long long v) {
    char buf[32];
    snprintf(buf, sizeof buf, " %11lld", v);
    return buf;
}

static void put(FILE *f, const std::string &s) { fputs(s.c_str(), f); fputc('\n', f); }

// ---------------------------------------------------------------- parameter file

namespace {

// One `read(19,*) dum, dum2, value`: three list-directed items; the first two
// land in character(1) variables (P:189,219), the third is the value.
bool third_item(const std::string &line, std::string &tok) {
    std::string s = line;
    for (char &c : s) if (c == ',') c = ' ';
    std::istringstream is(s);
    std::string a, b;
    return static_cast<bool>(is >> a >> b >> tok);
}

bool item_real(const std::string &line, double &v) {
    std::string t;
    if (!third_item(line, t)) return false;
    for (char &c : t) if (c == 'd' || c == 'D') c = 'e';     // Fortran double exponent
    char *end = nullptr;
    v = strtod(t.c_str(), &end);
    return end != t.c_str();
}

bool item_int(const std::string &line, int &v) {
    std::string t;
    if (!third_item(line, t)) return false;
    char *end = nullptr;
    const long x = strtol(t.c_str(), &end, 10);
    if (end == t.c_str() || (*end != '\0' && *end != ' ')) return false;  // "4.0" is an error in Fortran too
    v = (int)x;
    return true;
}

}  // namespace

bool read_parameter_file(const std::string &path, RunParams &p, std::string &err) {
    std::ifstream f(path);
    if (!f) { err = "cannot open " + path; return false; }
    std::vector<std::string> L;
    std::string line;
    while (std::getline(f, line)) L.push_back(line);
    if (L.size() < 29) { err = path + ": fewer than 29 lines"; return false; }
    auto bad = [&](int ln) {
        err = path + ": cannot read a value on line " + std::to_string(ln);
        return false;
    };
    // getProbSize: skip 3, then pSize, true2D (P:192-194)
    if (!item_int(L[3], p.pSize)) return bad(4);
    if (!item_int(L[4], p.true2D)) return bad(5);
    // paramSet: skip 5 (P:223), 12 values (P:225-236), skip xDel (P:237),
    // 3 tolerances (P:238-240), skip nit (P:241), skip 3 (P:242), 4 selectors (P:243-246)
    if (!item_real(L[5], p.depth)) return bad(6);
    if (!item_real(L[6], p.height)) return bad(7);
    if (!item_int(L[7], p.numInnocu)) return bad(8);
    if (!item_int(L[8], p.alpha)) return bad(9);
    if (!item_int(L[9], p.beta)) return bad(10);
    if (!item_real(L[10], p.nu)) return bad(11);
    if (!item_real(L[11], p.kappa)) return bad(12);
    if (!item_real(L[12], p.gama)) return bad(13);
    if (!item_real(L[13], p.delta)) return bad(14);
    if (!item_int(L[14], p.nOuts)) return bad(15);
    if (!item_real(L[15], p.tEnd)) return bad(16);
    if (!item_real(L[16], p.tDel)) return bad(17);
    if (!item_real(L[18], p.e1)) return bad(19);
    if (!item_real(L[19], p.e2)) return bad(20);
    if (!item_real(L[20], p.eSoln)) return bad(21);
    if (!item_int(L[25], p.fSelect)) return bad(26);
    if (!item_int(L[26], p.dSelect)) return bad(27);
    if (!item_int(L[27], p.gSelect)) return bad(28);
    if (!item_int(L[28], p.MinitialCond)) return bad(29);
    // P:91-100, 110-111
    p.row = p.pSize;
    p.col = (p.true2D == 1) ? 4 : p.pSize;      // any other true2D leaves col unset in the reference
    p.n = (long long)p.row * p.col;
    p.xDel = 1.0 / (double)p.row;
    p.nit = p.n * 100;
    return true;
}

// ---------------------------------------------------------------- initial conditions

namespace {

// stand-in for random_number(): splitmix64, uniform in [0,1)
struct Rng {
    uint64_t s;
    double next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        return (double)(z >> 11) * (1.0 / 9007199254740992.0);
    }
};

}  // namespace

// The inoculation points of MinitialCond 4, 5, 9, 10 in the order
// set_initial_conditions below draws / computes them; handed to
// pdeode_set_initial_condition so that the device builds the same field.
bool inoculation_points(const RunParams &p, uint64_t seed, std::vector<double> &px,
                        std::vector<double> &py) {
    px.clear(); py.clear();
    Rng rng{seed};
    const int ic = p.MinitialCond, np = p.numInnocu;
    if (ic == 4) {
        for (int k = 1; k <= np; ++k) { px.push_back(rng.next()); py.push_back(rng.next()); }
    } else if (ic == 5) {
        for (int k = 1; k <= np; ++k) { px.push_back(rng.next()); py.push_back(rng.next() * 0.1); }
    } else if (ic == 9) {
        for (int k = 1; k <= np; ++k) {
            px.push_back(p.depth + (k - 1) * 2 * p.depth +
                         ((k - 1) * (1 - np * p.depth * 2)) / (double)(np - 1));
            py.push_back(0.0);
        }
    } else if (ic == 10) {
        for (int k = 1; k <= 2 * np; ++k) {
            px.push_back(rng.next());
            double yr = rng.next() * 0.1;
            if (k > np) yr = yr + 0.9;
            py.push_back(yr);
        }
    } else if (ic == 7) {
        // one draw per cell, in cell order (P:365-372): the generator is a counter (splitmix64:
        // draw k is a function of seed + k * constant), so the device needs only the seed --
        // handed over as two 32-bit halves, each exact in a double
        px.push_back((double)(seed & 0xFFFFFFFFull));
        py.push_back((double)(seed >> 32));
    } else {
        return false;
    }
    return true;
}

void set_initial_conditions(std::vector<double> &C, std::vector<double> &M, const RunParams &p,
                            uint64_t seed) {
    const long long n = p.n;
    const int row = p.row, col = p.col;
    const double depth = p.depth, height = p.height, xDel = p.xDel;
    C.assign((size_t)n, 1.0);                                 // P:276
    M.assign((size_t)n, 0.0);
    Rng rng{seed};
    auto sq = [](double v) { return v * v; };
    const int ic = p.MinitialCond;
    if (ic == 1) {                                            // P:279-288
        const double a = -height / ((depth * depth) * (depth * depth));
        for (long long i = 1; i <= n; ++i) {
            const long long x = (i - 1) / col;
            const double t = (double)x * xDel;
            const double f = a * ((t * t) * (t * t)) + height;
            M[i - 1] = (f <= 0) ? 0.0 : f;
        }
    } else if (ic == 2) {                                     // P:291-292
        M.assign((size_t)n, height);
    } else if (ic == 3) {                                     // P:295-305
        const double a = -height / (depth * depth);
        for (long long i = 1; i <= n; ++i) {
            const long long x = i % col, y = i / row;
            const double f = a * (sq(x * xDel - 0.5) + sq(y * xDel - 0.5)) + height;
            M[i - 1] = (f <= 0) ? 0.0 : f;
        }
    } else if (ic == 4) {                                     // P:308-325
        const double a = -height / (depth * depth);
        for (int k = 1; k <= p.numInnocu; ++k) {
            const double xr = rng.next(), yr = rng.next();
            for (long long j = 1; j <= n; ++j) {
                if (M[j - 1] <= 0) {
                    const long long x = j % col, y = j / row;
                    M[j - 1] = M[j - 1] + a * (sq(x * xDel - xr) + sq(y * xDel - yr)) + height;
                    if (M[j - 1] <= 0) M[j - 1] = 0;
                }
            }
        }
    } else if (ic == 5) {                                     // P:328-352
        const double a = -height / (depth * depth);
        for (int k = 1; k <= p.numInnocu; ++k) {
            const double xr = rng.next();
            const double yr = rng.next() * 0.1;
            for (long long j = 1; j <= n; ++j) {
                const long long x = j % col, y = j / row;
                const double innoc = a * (sq(x * xDel - xr) + sq(y * xDel - yr)) + height;
                if (innoc >= 0) M[j - 1] = M[j - 1] + innoc;
            }
        }
        long long cnt = 0;
        double total = 0.0;                                   // uninitialised in the reference (D5)
        for (long long j = 1; j <= n; ++j)
            if (M[j - 1] > 0) { cnt++; total = total + M[j - 1]; }
        for (long long j = 1; j <= n; ++j) M[j - 1] = M[j - 1] * height / (total / (double)cnt);
    } else if (ic == 6) {                                     // P:356-362
        for (long long i = 1; i <= n; ++i) {
            const long long x = (i - 1) / col;
            if (x * xDel <= depth) M[i - 1] = height;
        }
    } else if (ic == 7) {                                     // P:365-372
        for (long long i = 1; i <= n; ++i)
            if ((double)i * xDel / col <= depth) M[i - 1] = rng.next() * 0.2;
    } else if (ic == 8) {                                     // P:375-387
        for (long long i = 1; i <= n; ++i) {
            M[i - 1] = (double)i / (double)n / 4.0;
            if (M[i - 1] >= 0.1) M[i - 1] = 0;
        }
    } else if (ic == 9) {                                     // P:390-402
        const double a = -height / (depth * depth);
        const int np = p.numInnocu;
        for (int k = 1; k <= np; ++k) {
            const double xr = depth + (k - 1) * 2 * depth +
                              ((k - 1) * (1 - np * depth * 2)) / (double)(np - 1);
            for (long long j = 1; j <= n; ++j) {
                if (M[j - 1] <= 0) {
                    const long long x = j % col, y = j / row;
                    M[j - 1] = M[j - 1] + a * (sq(x * xDel - xr) + sq(y * xDel)) + height;
                    if (M[j - 1] <= 0) M[j - 1] = 0;
                }
            }
        }
    } else if (ic == 10) {                                    // P:405-421
        const double a = -height / (depth * depth);
        for (int k = 1; k <= 2 * p.numInnocu; ++k) {
            const double xr = rng.next();
            double yr = rng.next() * 0.1;
            if (k > p.numInnocu) yr = yr + 0.9;
            for (long long j = 1; j <= n; ++j) {
                const long long x = j % col, y = j / row;
                const double innoc = a * (sq(x * xDel - xr) + sq(y * xDel - yr)) + height;
                if (innoc >= 0) M[j - 1] = M[j - 1] + innoc;
            }
        }
    }
}

// ---------------------------------------------------------------- writers

static int frame_filter(int row) { return (row > 257) ? (row - 1) / 256 : 1; }   // P:829-832

void print_to_file(FILE *f, const RunParams &p, const double *M, const double *C) {
    const int row = p.row, col = p.col, filter = frame_filter(row);
    for (int i = 1; i <= row; ++i) {
        if ((i - 1) % filter != 0) continue;
        for (int j = 1; j <= col; ++j) {
            if ((j - 1) % filter != 0) continue;
            const size_t q = (size_t)j + (size_t)(i - 1) * col;                  // P:838
            std::string s = fmt_real((double)(j - 1) / (double)(col - 1)) +
                            fmt_real((double)(i - 1) / (double)(row - 1)) + fmt_real(M[q - 1]) +
                            fmt_real(C[q - 1]);
            put(f, s);                                                           // P:839-840
        }
        put(f, "  ");                                                            // P:843 write(11,*) ' '
    }
    put(f, "");                                                                  // P:846 frame terminator
}

void print_to_file_2d(FILE *f, const RunParams &p, const double *M, const double *C) {
    const int row = p.row, col = p.col, filter = frame_filter(row);
    for (int i = 1; i <= row; ++i) {
        if ((i - 1) % filter != 0) continue;
        double averageM = 0, averageC = 0;
        for (int j = 1; j <= col; ++j) {
            const size_t q = (size_t)j + (size_t)(i - 1) * col;
            averageM = averageM + M[q - 1];
            averageC = averageC + C[q - 1];
        }
        averageM = averageM / col;
        averageC = averageC / col;
        const double y = (double)(i - 1) / (double)(row - 1);
        fprintf(f, "%20.12f%20.12f%20.12f\n", y, averageM, averageC);            // P:926
    }
    put(f, "");                                                                  // P:930
}

// The same two frames from data already decimated on the device
// (pdeode_get_frame / pdeode_get_frame_rowmeans): nro x nco points, point
// (oi, oj) being cell (oi*filter, oj*filter).
// A frame is fixed-width text -- 105 characters per point (P:839-840), a
// 3-character separator line per row (P:843), one empty line at the end
// (P:846) -- so its rows can be formatted in place by several threads.
void print_frame(FILE *f, const RunParams &p, int nro, int nco, const double *Md, const double *Cd) {
    const int filter = frame_filter(p.row);
    const size_t row_bytes = (size_t)nco * 105 + 3;
    std::vector<char> text((size_t)nro * row_bytes + 1);
    std::vector<char> xs((size_t)nco * 26);          // the x column repeats in every row
    for (int oj = 0; oj < nco; ++oj)
        fmt_real_to(xs.data() + (size_t)oj * 26, (double)(oj * filter) / (double)(p.col - 1));
    auto rows = [&](int lo, int hi) {
        for (int oi = lo; oi < hi; ++oi) {
            char ys[26];
            fmt_real_to(ys, (double)(oi * filter) / (double)(p.row - 1));
            char *line = text.data() + (size_t)oi * row_bytes;
            for (int oj = 0; oj < nco; ++oj, line += 105) {
                memcpy(line, xs.data() + (size_t)oj * 26, 26);
                memcpy(line + 26, ys, 26);
                fmt_real_to(line + 52, Md[(size_t)oi * nco + oj]);
                fmt_real_to(line + 78, Cd[(size_t)oi * nco + oj]);
                line[104] = '\n';
            }
            memcpy(line, "  \n", 3);
        }
    };
    unsigned nt = std::thread::hardware_concurrency();
    if (nt > 8) nt = 8;
    if ((size_t)nro * nco < 16384 || nt < 2) {
        rows(0, nro);
    } else {
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nt; ++t)
            pool.emplace_back(rows, (int)((long long)nro * t / nt), (int)((long long)nro * (t + 1) / nt));
        for (auto &t : pool) t.join();
    }
    text.back() = '\n';
    fwrite(text.data(), 1, text.size(), f);
}

void print_frame_2d(FILE *f, const RunParams &p, int nro, const double *meanM, const double *meanC) {
    const int filter = frame_filter(p.row);
    for (int oi = 0; oi < nro; ++oi)
        fprintf(f, "%20.12f%20.12f%20.12f\n", (double)(oi * filter) / (double)(p.row - 1), meanM[oi],
                meanC[oi]);
    put(f, "");
}

bool report_stats(const std::string &path, double avgIters, double maxIters, double avgNit,
                  double maxNit, double seconds) {
    FILE *f = fopen(path.c_str(), "w");                                          // P:1025-1027
    if (!f) return false;
    put(f, " Statsitcs:");
    put(f, " -------------------------------------------------------------");
    put(f, " Time to compute = " + fmt_real(seconds));
    put(f, " ");
    put(f, " Avg Iters for iterating betn. soln. =" + fmt_real(avgIters));
    put(f, " Max Iters for iterating betn. soln. =" + fmt_real(maxIters));
    put(f, " Avg Iters for linear solver =" + fmt_real(avgNit));
    put(f, " Max Iters for linear solver =" + fmt_real(maxNit));
    fclose(f);
    return true;
}

// ---------------------------------------------------------------- the time loop

// How many time steps the loop P:663-742 takes from `counter` on before it looks
// at the state again: diagnostics when mod(counter, filter/100+1) == 0 (P:665), a
// frame when mod(counter, filter) == 0 (P:683), the end when counter*tDel > tEnd
// (P:663).  Always >= 1: the step at `counter` itself is due.
int steps_until_next_look(long long counter, int filter, double tDel, double tEnd) {
    const long long cad = (long long)(filter / 100 + 1);
    int batch = 1;
    while (batch < 4096 && (counter + batch) % cad != 0 && (counter + batch) % filter != 0 &&
           (double)(counter + batch) * tDel <= tEnd)
        ++batch;
    return batch;
}

static std::string join(const std::string &dir, const char *name) {
    return dir.empty() ? std::string(name) : dir + "/" + name;
}

namespace {

// Frames leave the time loop through a short queue: one background thread
// formats and writes them, in order, while the device keeps stepping.  (The
// shipped 257 x 257 run writes 631 MB of frame text; formatted inside the time
// loop that costs twice what the 30 001 time steps do.)
class FrameSink {
  public:
    explicit FrameSink(const RunParams &p) : p_(p) {}
    ~FrameSink() { finish(); }
    void push(FILE *f, int nro, int nco, const std::vector<double> &M, const std::vector<double> &C) {
        std::unique_lock<std::mutex> lk(mu_);
        if (!started_) {
            done_ = false;
            th_ = std::thread([this] { run(); });
            started_ = true;
        }
        space_.wait(lk, [&] { return q_.size() < 4; });
        q_.push_back(Job{f, nro, nco, M, C});
        work_.notify_one();
    }
    // returns once every queued frame is in its file
    void finish() {
        if (!started_) return;
        {
            std::lock_guard<std::mutex> lk(mu_);
            done_ = true;
        }
        work_.notify_one();
        th_.join();
        started_ = false;
    }

  private:
    struct Job {
        FILE *f;
        int nro, nco;
        std::vector<double> M, C;
    };
    void run() {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu_);
                work_.wait(lk, [&] { return done_ || !q_.empty(); });
                if (q_.empty()) return;
                j = std::move(q_.front());
                q_.pop_front();
            }
            space_.notify_one();
            if (p_.true2D == 1) print_frame_2d(j.f, p_, j.nro, j.M.data(), j.C.data());
            else print_frame(j.f, p_, j.nro, j.nco, j.M.data(), j.C.data());
        }
    }
    const RunParams p_;
    std::mutex mu_;
    std::condition_variable work_, space_;
    std::deque<Job> q_;
    std::thread th_;
    bool started_ = false, done_ = false;
};

}  // namespace

RunStats solve_order2(const RunParams &p, std::vector<double> &M, std::vector<double> &C,
                      const Stepper &st, int device, const std::string &outdir, FILE *con,
                      const DeviceIc *dic) {
    RunStats rs;
    const double tDel = p.tDel, tEnd = p.tEnd;
    const int filter = (int)(tEnd / (p.nOuts * tDel));                           // P:638
    if (filter <= 0) {   // mod(counter, 0) at P:683 is undefined in the reference (S8)
        fprintf(con, " [!] nOuts*tDel > tEnd: output cadence is zero. Exit!\n");
        rs.status = PDEODE_ERR_ARG;
        return rs;
    }
    pdeode_params pp{};
    pp.delta = p.delta; pp.nu = p.nu; pp.kappa = p.kappa; pp.gama = p.gama;
    pp.tDel = p.tDel; pp.xDel = p.xDel; pp.e1 = p.e1; pp.e2 = p.e2; pp.eSoln = p.eSoln;
    pp.alpha = p.alpha; pp.beta = p.beta;
    pp.dSelect = p.dSelect; pp.fSelect = p.fSelect; pp.gSelect = p.gSelect;
    pdeode_ctx *ctx = nullptr;
    int rc = st.create(&ctx, p.row, p.col, &pp, device);
    auto fail = [&](const char *what) {
        fprintf(con, " [!] %s: %s\n", what, st.strerror_(rc));
        if (ctx) st.destroy(ctx);
        rs.status = rc & PDEODE_ERR_MASK;
        return rs;
    };
    if (rc) return fail("pdeode_create");
    const bool device_ic = dic && st.set_initial_condition;
    if (device_ic) {
        // the fields are built where they will live: nothing to upload
        rc = st.set_initial_condition(ctx, dic->ic, dic->depth, dic->height, (int)dic->px.size(),
                                      dic->px.data(), dic->py.data());
        if (rc) return fail("pdeode_set_initial_condition");
    } else if ((rc = st.set_state(ctx, M.data(), C.data()))) {
        return fail("pdeode_set_state");
    }

    long long counter = 0;                                                       // P:640
    double avgIters = 0, avgNit = 0;                                             // P:642-643
    double maxIters = 0, maxNit = 0;       // read before set in the reference (D8): taken as 0
    double prevMassC = 1;                                                        // P:645
    double totalMassM = 0, totalMassC = 0;
    double peak = 0, height = 0, intfac = 0;                                     // P:633
    FILE *ftotal = fopen(join(outdir, "total.dat").c_str(), "w");                // P:647-649
    FILE *fco = fopen(join(outdir, "COprod.dat").c_str(), "w");                  // P:651-653
    FILE *fpeak = (p.true2D == 1) ? fopen(join(outdir, "peakInfo.dat").c_str(), "w") : nullptr;
    FILE *fframe = nullptr;                // opened by the first writer call (P:824-827, P:888-891)
    FrameSink sink(p);
    const char *so = getenv("PDEODE_SYNC_OUTPUT");
    const bool async_frames = !(so && so[0] == '1');
    auto close_all = [&]() {
        sink.finish();
        if (ftotal) fclose(ftotal);
        if (fco) fclose(fco);
        if (fpeak) fclose(fpeak);
        if (fframe) fclose(fframe);
    };
    fprintf(con, "    time    avgIter  maxIter      avgNit  maxNit        avgM        avgC\n");  // P:661

    const auto t0 = std::chrono::steady_clock::now();
    // PDEODE_HOST_TIMING=1: where the wall time of the loop goes (stderr, at the end)
    const bool timing = getenv("PDEODE_HOST_TIMING") != nullptr;
    double t_step = 0, t_diag = 0, t_frame = 0;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto since = [](std::chrono::steady_clock::time_point a) {
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count();
    };
    bool have_host_copy = !device_ic;      // M, C on the host equal the device state
    std::vector<double> fbufM, fbufC;      // decimated frame data
    while ((double)counter * tDel <= tEnd) {                                     // P:663
        const auto td = now();
        if (counter % (long long)(filter / 100 + 1) == 0) {                      // P:665
            if ((rc = st.mass(ctx, &totalMassM, &totalMassC))) { close_all(); return fail("pdeode_mass"); }
            if (ftotal) put(ftotal, fmt_real(counter * tDel) + fmt_real(totalMassM) + fmt_real(totalMassC));
            if (fco) put(fco, fmt_real(counter * tDel) + fmt_real(prevMassC - totalMassC) +
                                   fmt_real(1 - totalMassC));                    // P:672
            prevMassC = totalMassC;
            if (p.true2D == 1) {                                                 // P:676-679
                if ((rc = st.peak(ctx, &peak, &height, &intfac))) { close_all(); return fail("pdeode_peak"); }
                if (fpeak) put(fpeak, fmt_real(tDel * counter) + fmt_real(peak) + fmt_real(height) +
                                          fmt_real(intfac));
            }
        }
        t_diag += since(td);
        const auto tf = now();
        if (counter % filter == 0) {                                             // P:683
            if (!fframe)
                fframe = fopen(join(outdir, p.true2D == 1 ? "2D_output.dat" : "output.dat").c_str(), "w");
            if (st.frame_size && st.get_frame && st.get_frame_rowmeans) {
                // only the points a frame contains cross the bus
                int nro = 0, nco = 0;
                const int filter = frame_filter(p.row);
                if ((rc = st.frame_size(ctx, filter, &nro, &nco))) { close_all(); return fail("pdeode_frame_size"); }
                fbufM.resize((size_t)nro * (p.true2D == 1 ? 1 : nco));
                fbufC.resize(fbufM.size());
                if (p.true2D == 1) rc = st.get_frame_rowmeans(ctx, filter, fbufM.data(), fbufC.data());
                else rc = st.get_frame(ctx, filter, fbufM.data(), fbufC.data());
                if (rc) { close_all(); return fail("pdeode_get_frame"); }
                if (fframe && async_frames) {
                    sink.push(fframe, nro, nco, fbufM, fbufC);
                } else if (fframe) {
                    if (p.true2D == 1) print_frame_2d(fframe, p, nro, fbufM.data(), fbufC.data());
                    else print_frame(fframe, p, nro, nco, fbufM.data(), fbufC.data());
                }
            } else {
                if (!have_host_copy) {
                    if ((rc = st.get_state(ctx, M.data(), C.data()))) { close_all(); return fail("pdeode_get_state"); }
                    have_host_copy = true;
                }
                if (fframe) {
                    if (p.true2D == 1) print_to_file_2d(fframe, p, M.data(), C.data());
                    else print_to_file(fframe, p, M.data(), C.data());
                }
            }
            rs.frames++;
            if (counter == 0)                                                    // P:690-692
                fprintf(con, "%8.2f%12.2f%8d%12.2f%8d%12.6f%12.6f\n", 0.0, 0.0, (int)maxIters, 0.0,
                        (int)maxNit, totalMassM, totalMassC);
            else                                                                 // P:694-696
                fprintf(con, "%8.2f%12.2f%8d%12.2f%8d%12.6f%12.6f\n", tDel * counter,
                        avgIters / counter, (int)maxIters, avgNit / avgIters, (int)maxNit,
                        totalMassM, totalMassC);
            fflush(con);
        }
        t_frame += since(tf);
        // the stretch up to the next point where the loop looks at the state (P:665, P:683)
        // or ends (P:663) goes to the device in one call
        const int batch = st.step_many ? steps_until_next_look(counter, filter, tDel, tEnd) : 1;
        int countIters = 0, stepsDone = 1;
        long long countSum = 0, nitSum = 0, nitMax = 0;
        const auto ts = now();
        if (batch > 1) {
            rc = st.step_many(ctx, batch, &stepsDone, &countSum, &countIters, &nitSum, &nitMax);
        } else {
            rc = st.step(ctx, &countIters, &nitSum, &nitMax);                    // P:700-738
            countSum = countIters;
        }
        t_step += since(ts);
        have_host_copy = false;
        if ((rc & PDEODE_ERR_MASK) == PDEODE_ERR_PICARD_CAP) {                   // P:732-736
            fprintf(con, " [!] Over 10000 iterations in one timestep\n");
            fprintf(con, " [!] Solutions not converging. Exit!\n");
            close_all();
            st.destroy(ctx);
            rs.status = PDEODE_ERR_PICARD_CAP;
            return rs;
        }
        if (rc & PDEODE_ERR_MASK) { close_all(); return fail("pdeode_step"); }
        if ((double)nitMax > maxNit) maxNit = (double)nitMax;                    // P:722
        avgNit = avgNit + (double)nitSum;                                        // P:723
        if (countIters > maxIters) maxIters = countIters;                        // P:739
        avgIters = avgIters + (double)countSum;                                  // P:740
        counter = counter + stepsDone;                                           // P:742
    }
    avgNit = avgNit / avgIters;                                                  // P:744
    avgIters = avgIters / counter;                                               // P:745
    // the final state is never written by the reference (P:747-755); the host
    // arrays are still refreshed, as solveOrder2's intent(out) M, C would be
    rc = st.get_state(ctx, M.data(), C.data());
    if (const char *fs = getenv("PDEODE_FINAL_STATE")) {
        // the "Print final solution" the reference has commented out (P:747-755), at
        // full resolution so that MinitialCond = 11 can continue from it
        FILE *ff = fopen(join(outdir, fs).c_str(), "w");
        if (ff) {
            for (int i = 1; i <= p.row; ++i) {
                for (int j = 1; j <= p.col; ++j) {
                    const size_t q = (size_t)j + (size_t)(i - 1) * p.col;
                    put(ff, fmt_real((double)(j - 1) / (double)(p.col - 1)) +
                                fmt_real((double)(i - 1) / (double)(p.row - 1)) + fmt_real(M[q - 1]) +
                                fmt_real(C[q - 1]));
                }
                put(ff, "  ");
            }
            fclose(ff);
        }
    }
    const auto tq = now();
    sink.finish();          // "Time to compute" covers the frames still in the queue
    const auto t1 = std::chrono::steady_clock::now();
    if (timing)
        fprintf(stderr, "host timing: steps %.3f s, diagnostics %.3f s, frames %.3f s, frame queue drain %.3f s\n",
                t_step, t_diag, t_frame, since(tq));
    close_all();
    st.destroy(ctx);
    rs.avgIters = avgIters; rs.maxIters = maxIters; rs.avgNit = avgNit; rs.maxNit = maxNit;
    rs.seconds = std::chrono::duration<double>(t1 - t0).count();
    rs.steps = counter;
    rs.status = rc & PDEODE_ERR_MASK;
    return rs;
}

// ---------------------------------------------------------------- program cThermoPDEODE

int run_program(FILE *in, FILE *con, const Stepper &st, int device, const std::string &outdir) {
    auto say = [&](const std::string &s) { put(con, " " + s); };
    say("Enter parameter file name: ex. 'parameter.txt' ");                      // P:81
    fputs("    ", con);                                                          // P:82
    char name[256] = {0};
    if (fscanf(in, "%255s", name) != 1) { say("[!] no parameter file name given"); return 2; }
    std::string filename(name);
    if (filename.size() >= 2 && (filename.front() == '\'' || filename.front() == '"'))
        filename = filename.substr(1, filename.size() - 2);       // list-directed read strips quotes

    say("Setting Parameters that shouldn't change");                             // P:85
    fprintf(con, "     ndiag = %5d\n", 5);                                       // P:87
    say("Getting problem size from file");                                       // P:89
    say(filename);                                                               // P:190
    RunParams p;
    std::string err;
    if (!read_parameter_file(filename, p, err)) { say("[!] " + err); return 2; }
    if (p.true2D == 1) say("    Problem is 2D");                                 // P:91-97
    else if (p.true2D == 0) say("    Problem is 3D");
    else { say("[!] true2D must be 0 or 1"); return 2; }
    say("    pSize = " + fmt_int(p.pSize));
    say("    row   = " + fmt_int(p.row));
    say("    col   = " + fmt_int(p.col));
    say("    n     = " + fmt_int(p.n));
    say("Opening Parameter file");                                               // P:105
    say("Parameters set:");                                                      // P:112-131
    say("    depth       = " + fmt_real(p.depth));
    say("    height      = " + fmt_real(p.height));
    say("    alpha       = " + fmt_int(p.alpha));
    say("    beta        = " + fmt_int(p.beta));
    say("    gama        = " + fmt_real(p.gama));
    say("    nu          = " + fmt_real(p.nu));
    say("    delta       = " + fmt_real(p.delta));
    say("    nOuts       = " + fmt_int(p.nOuts));
    say("    tEnd        = " + fmt_real(p.tEnd));
    say("    tDel        = " + fmt_real(p.tDel));
    say("    e1          = " + fmt_real(p.e1));
    say("    e2          = " + fmt_real(p.e2));
    say("    eSoln       = " + fmt_real(p.eSoln));
    say("    nit         = " + fmt_int(p.nit));
    say("    xDel        = " + fmt_real(p.xDel));
    say("    fSelect     = " + fmt_int(p.fSelect));
    say("    dSelect     = " + fmt_int(p.dSelect));
    say("    gSelect     = " + fmt_int(p.gSelect));
    say("    MinitialCond= " + fmt_int(p.MinitialCond));
    say("Allocating the size of arrays");                                        // P:132
    fprintf(con, "    C and M are now dimension(%12lld) arrays\n", p.n);         // P:134
    say("Setting Initial Conditions");                                           // P:136
    uint64_t seed;
    if (const char *s = getenv("PDEODE_SEED")) seed = strtoull(s, nullptr, 10);
    else seed = (uint64_t)std::chrono::system_clock::now().time_since_epoch().count();  // clock, as init_random_seed.f90:11
    std::vector<double> C, M;
    // Build the fields on the device when the stepper offers it (all cases but the
    // restart, 11); PDEODE_HOST_IC=1 forces the host loops.
    DeviceIc dic;
    const int icn = p.MinitialCond;
    const bool use_dic = st.set_initial_condition && st.frame_size && icn >= 1 && icn <= 10 &&
                         !(getenv("PDEODE_HOST_IC") && getenv("PDEODE_HOST_IC")[0] == '1');
    if (use_dic) {
        dic.ic = icn; dic.depth = p.depth; dic.height = p.height;
        inoculation_points(p, seed, dic.px, dic.py);
        M.assign((size_t)p.n, 0.0);
        C.assign((size_t)p.n, 1.0);
    } else {
        set_initial_conditions(C, M, p, seed);
    }
    if (p.MinitialCond == 11) {
        // The restart the reference's author sketched and left commented out as
        // "CANNOT GET WORKING" (P:423-435): continue from a previous run's state.
        // initialCond.dat holds one "x y M C" line per cell in printToFile's order
        // (blank separator lines ignored), e.g. the file PDEODE_FINAL_STATE writes.
        std::ifstream f(join(outdir, "initialCond.dat"));
        std::string line;
        long long k = 0;
        while (f && k < p.n && std::getline(f, line)) {
            std::istringstream is(line);
            double x, y, m, c;
            if (is >> x >> y >> m >> c) { M[(size_t)k] = m; C[(size_t)k] = c; ++k; }
        }
        if (k != p.n) {
            say("[!] initialCond.dat: expected " + std::to_string(p.n) + " cells, read " + std::to_string(k));
            return 2;
        }
        say("    M, C continue from initialCond.dat");
    }
    say("    C = 1");                                                            // P:139
    if (p.MinitialCond == 1) say("    M has non-homogenous Initial Conditions");
    else if (p.MinitialCond == 2) {
        say("    M has homogenous initial conditions");
        say("        M = " + fmt_real(p.height));
    }
    say("Starting Solver");                                                      // P:148
    RunStats rs = solve_order2(p, M, C, st, device, outdir, con, use_dic ? &dic : nullptr);
    if (rs.status == PDEODE_ERR_PICARD_CAP) return 0;       // Fortran `stop`: exit code 0
    if (rs.status) return 1;
    say("-------------------------------------------------------------");        // P:161-169
    say("Statsitcs:");
    say("-------------------------------------------------------------");
    say("Time to compute = " + fmt_real(rs.seconds));
    say("");
    say("Avg Iters for iterating betn. soln. =" + fmt_real(rs.avgIters));
    say("Max Iters for iterating betn. soln. =" + fmt_real(rs.maxIters));
    say("Avg Iters for linear solver =" + fmt_real(rs.avgNit));
    say("Max Iters for linear solver =" + fmt_real(rs.maxNit));
    say("Writing statistics to file");
    report_stats(join(outdir, "statReport.dat"), rs.avgIters, rs.maxIters, rs.avgNit, rs.maxNit,
                 rs.seconds);                                                    // P:171
    say("Execution completed");                                                  // P:173
    return 0;
}

}  // namespace pdeode_host
