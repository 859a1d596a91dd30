/* f90rt.c -- runtime of the Fortran-to-C translation (see f90toc.py, f90rt.h).
 * TEST INFRASTRUCTURE.  Implements just the I/O, clock and random-number
 * behaviour the reference's three source files rely on, following gfortran:
 *   - list-directed output: one leading blank per record, INTEGER(4) in 12
 *     columns, REAL(8) as G25.17E3 behind one blank; an empty output list is an
 *     empty record ("\n"), which is what seperateOutput.py:14 keys on;
 *   - OPEN of a file that is connected to another unit fails (IOSTAT /= 0),
 *     OPEN of a unit already connected to the same file changes nothing, so the
 *     frame writers append after their first call (PDE-ODEsolver.f90:824-827);
 *   - list-directed input: blank / comma separated items, a CHARACTER(1) item
 *     takes the first character of its token, '/' ends the read.
 * random_number is NOT gfortran's xoshiro256**: with PDEODE_SEED set it is the
 * splitmix64 stream the B200 host program draws from the same seed, so that
 * both programs build the same random inoculation points;
 * otherwise it is seeded from random_seed(put=). */
#include "f90rt.h"

#include <ctype.h>
#include <stdint.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

int f90_ipow(int x, int n) {
    int y = 1;
    if (n < 0) return (x == 1) ? 1 : ((x == -1) ? ((n & 1) ? -1 : 1) : 0);
    while (n) { if (n & 1) y *= x; x *= x; n >>= 1; }
    return y;
}

void *f90_calloc(long n, long size) {
    void *p = calloc((size_t)(n > 0 ? n : 1), (size_t)size);
    if (!p) { fprintf(stderr, "f90rt: out of memory\n"); exit(2); }
    return p;
}

/* ---------------------------------------------------------------- units */

#define MAXUNITS 64
typedef struct {
    int unit;
    FILE *f;
    char name[512];
    int used;
} Unit;
static Unit units[MAXUNITS];

static Unit *find_unit(int unit) {
    for (int i = 0; i < MAXUNITS; ++i)
        if (units[i].used && units[i].unit == unit) return &units[i];
    return NULL;
}
static Unit *find_file(const char *name) {
    for (int i = 0; i < MAXUNITS; ++i)
        if (units[i].used && strcmp(units[i].name, name) == 0) return &units[i];
    return NULL;
}
static void trimcpy(char *dst, size_t cap, const char *s, int len) {
    while (len > 0 && (s[len - 1] == ' ' || s[len - 1] == '\0')) --len;
    if ((size_t)len >= cap) len = (int)cap - 1;
    memcpy(dst, s, (size_t)len);
    dst[len] = '\0';
}
static int streq_ci(const char *s, int len, const char *lit) {
    char b[32];
    trimcpy(b, sizeof b, s ? s : "", s ? len : 0);
    for (char *p = b; *p; ++p) *p = (char)tolower((unsigned char)*p);
    return strcmp(b, lit) == 0;
}

void f90_flush_all(void) {
    for (int i = 0; i < MAXUNITS; ++i)
        if (units[i].used && units[i].f) fflush(units[i].f);
    fflush(stdout);
}

/* tests that call the time loop several times in one process: forget all connections
 * (a fresh process starts with none) */
void f90_close_all(void) {
    for (int i = 0; i < MAXUNITS; ++i)
        if (units[i].used) { if (units[i].f) fclose(units[i].f); units[i].used = 0; }
}

void f90_stop(void) {
    f90_flush_all();
    exit(0);
}

void f90_open(int unit, const char *file, int filelen, const char *status, int statuslen,
              const char *action, int actionlen, const char *position, int positionlen,
              int *iostat) {
    char name[512];
    trimcpy(name, sizeof name, file ? file : "", file ? filelen : 0);
    const int old = streq_ci(status, statuslen, "old");
    const int rd = streq_ci(action, actionlen, "read");
    const int app = streq_ci(position, positionlen, "append");
    int err = 0;
    Unit *u = find_unit(unit);
    if (u && strcmp(u->name, name) == 0) {       /* same unit, same file: nothing changes */
        if (iostat) *iostat = 0;
        return;
    }
    if (find_file(name)) {
        err = 5004;                              /* already connected to another unit */
    } else {
        struct stat sb;
        const int exists = (stat(name, &sb) == 0);
        if (old && !exists) err = 2;
        if (!err) {
            if (u) { if (u->f) fclose(u->f); u->used = 0; }
            FILE *f = fopen(name, rd ? "r" : (app ? "a+" : (old ? "r+" : "w+")));
            if (!f) err = 13;
            else {
                for (int i = 0; i < MAXUNITS; ++i)
                    if (!units[i].used) {
                        units[i].used = 1; units[i].unit = unit; units[i].f = f;
                        snprintf(units[i].name, sizeof units[i].name, "%s", name);
                        break;
                    }
            }
        }
    }
    if (iostat) *iostat = err;
    else if (err) { fprintf(stderr, "f90rt: cannot open '%s' on unit %d (error %d)\n", name, unit, err); exit(2); }
}

void f90_close(int unit, const char *status, int statuslen) {
    Unit *u = find_unit(unit);
    if (!u) return;
    if (u->f) fclose(u->f);
    if (streq_ci(status, statuslen, "delete")) unlink(u->name);
    u->used = 0;
}

/* ---------------------------------------------------------------- WRITE */

static struct {
    FILE *f;
    const char *fmt;      /* NULL: list-directed */
    int noadv;
    int nitems;
    const char *fp;       /* cursor in fmt */
} W;

static FILE *unit_file_w(int unit) {
    if (unit < 0 || unit == 6) return stdout;
    Unit *u = find_unit(unit);
    if (!u) {                                    /* implicit fort.N */
        char nm[32];
        snprintf(nm, sizeof nm, "fort.%d", unit);
        f90_open(unit, nm, (int)strlen(nm), NULL, 0, NULL, 0, NULL, 0, NULL);
        u = find_unit(unit);
    }
    return u->f;
}

void f90_w_begin(int unit, const char *fmt, int no_advance) {
    W.f = unit_file_w(unit);
    W.fmt = fmt; W.noadv = no_advance; W.nitems = 0;
    W.fp = fmt ? strchr(fmt, '(') : NULL;
    if (W.fp) ++W.fp;
}

/* next edit descriptor: letter, width, digits */
static int next_desc(char *letter, int *w, int *d) {
    if (!W.fp) return 0;
    while (*W.fp == ' ' || *W.fp == ',') ++W.fp;
    if (*W.fp == ')' || *W.fp == '\0') {         /* format reversion: start over */
        W.fp = strchr(W.fmt, '(') + 1;
        fputc('\n', W.f);
        while (*W.fp == ' ' || *W.fp == ',') ++W.fp;
    }
    *letter = (char)toupper((unsigned char)*W.fp++);
    *w = 0; *d = 0;
    while (isdigit((unsigned char)*W.fp)) *w = *w * 10 + (*W.fp++ - '0');
    if (*W.fp == '.') { ++W.fp; while (isdigit((unsigned char)*W.fp)) *d = *d * 10 + (*W.fp++ - '0'); }
    return 1;
}

static void put_field(const char *s, int w) {
    const int n = (int)strlen(s);
    if (w <= 0) { fputs(s, W.f); return; }
    if (n > w) { for (int i = 0; i < w; ++i) fputc('*', W.f); return; }
    for (int i = 0; i < w - n; ++i) fputc(' ', W.f);
    fputs(s, W.f);
}

static void fmt_real_ld(char *out, double x) {   /* out: 96 bytes */
    /* gfortran list-directed REAL(8): blank + G25.17E3 */
    if (isnan(x)) { snprintf(out, 96, "%26s", "NaN"); return; }
    if (isinf(x)) { snprintf(out, 96, "%26s", x > 0 ? "Infinity" : "-Infinity"); return; }
    if (x == 0.0) { snprintf(out, 96, "%s0.0000000000000000     ", signbit(x) ? "  -" : "   "); return; }
    char e[64];
    snprintf(e, sizeof e, "%.16E", x);
    const int ex = atoi(strchr(e, 'E') + 1);
    const int k = ex + 1;
    if (k >= 0 && k <= 17) {
        char body[64];
        snprintf(body, sizeof body, "%.*f", 17 - k, x);
        /* %.17f of a number in [0.1, 1) prints "0.1...": 19 chars, as gfortran's F20.17 does */
        snprintf(out, 96, " %20s     ", body);
    } else {
        char m[32];
        snprintf(m, sizeof m, "%.16E", x);
        *strchr(m, 'E') = '\0';
        char body[64];
        snprintf(body, sizeof body, "%sE%c%03d", m, ex < 0 ? '-' : '+', ex < 0 ? -ex : ex);
        snprintf(out, 96, " %25s", body);
    }
}

void f90_w_int(int v) {
    char b[64];
    if (!W.fmt) {                                /* blank + I11; the blank doubles as the record's */
        W.nitems++;
        snprintf(b, sizeof b, " %11d", v);
        fputs(b, W.f);
        return;
    }
    char L; int w, d;
    next_desc(&L, &w, &d);
    snprintf(b, sizeof b, "%d", v);
    put_field(b, w);
    W.nitems++;
}

void f90_w_real(double v) {
    char b[96];
    if (!W.fmt) {                                /* blank + G25.17E3 */
        W.nitems++;
        fmt_real_ld(b, v);
        fputs(b, W.f);
        return;
    }
    char L; int w, d;
    next_desc(&L, &w, &d);
    snprintf(b, sizeof b, "%.*f", d, v);
    /* gfortran drops the optional leading zero when the field is too narrow */
    if ((int)strlen(b) > w && w > 0) {
        if (strncmp(b, "0.", 2) == 0) memmove(b, b + 1, strlen(b));
        else if (strncmp(b, "-0.", 3) == 0) memmove(b + 1, b + 2, strlen(b + 1));
    }
    put_field(b, w);
    W.nitems++;
}

void f90_w_str(const char *s, int len) {
    if (!W.fmt) {
        if (W.nitems++ == 0) fputc(' ', W.f);
        fwrite(s, 1, (size_t)len, W.f);
        return;
    }
    char L; int w, d;
    next_desc(&L, &w, &d);
    if (w > 0 && w < len) len = w;
    for (int i = 0; i < w - len; ++i) fputc(' ', W.f);
    fwrite(s, 1, (size_t)len, W.f);
    W.nitems++;
}

void f90_w_end(void) {
    if (!W.noadv) fputc('\n', W.f);
    if (W.f == stdout) fflush(stdout);
}

/* ---------------------------------------------------------------- READ */

static struct {
    FILE *f;
    char line[4096];
    char *p;
    int have;        /* a record is loaded */
    int ended;       /* '/' seen or EOF */
} R;

static void load_record(void) {
    if (!fgets(R.line, sizeof R.line, R.f)) { R.line[0] = '\0'; R.ended = 1; }
    R.p = R.line;
    R.have = 1;
}

void f90_r_begin(int unit) {
    if (unit < 0 || unit == 5) R.f = stdin;
    else {
        Unit *u = find_unit(unit);
        if (!u) { fprintf(stderr, "f90rt: read from unconnected unit %d\n", unit); exit(2); }
        R.f = u->f;
    }
    R.ended = 0;
    load_record();                               /* every READ starts a new record */
}

/* next list-directed token; returns 0 when the read has ended */
static int next_token(char *tok, size_t cap) {
    for (;;) {
        if (R.ended) return 0;
        while (*R.p == ' ' || *R.p == '\t' || *R.p == ',' || *R.p == '\r') ++R.p;
        if (*R.p == '\n' || *R.p == '\0') { load_record(); continue; }
        break;
    }
    if (*R.p == '/') { R.ended = 1; return 0; }
    size_t n = 0;
    if (*R.p == '\'' || *R.p == '"') {
        const char q = *R.p++;
        while (*R.p && *R.p != '\n' && *R.p != q) { if (n + 1 < cap) tok[n++] = *R.p; ++R.p; }
        if (*R.p == q) ++R.p;
    } else {
        while (*R.p && !strchr(" \t,\n\r/", *R.p)) { if (n + 1 < cap) tok[n++] = *R.p; ++R.p; }
    }
    tok[n] = '\0';
    return 1;
}

void f90_r_int(int *v) {
    char t[256];
    if (!next_token(t, sizeof t)) return;
    char *end;
    const long x = strtol(t, &end, 10);
    if (end == t || *end) { fprintf(stderr, "f90rt: bad integer in list-directed input: '%s'\n", t); exit(2); }
    *v = (int)x;
}

void f90_r_real(double *v) {
    char t[256];
    if (!next_token(t, sizeof t)) return;
    for (char *p = t; *p; ++p) if (*p == 'd' || *p == 'D') *p = 'e';
    char *end;
    const double x = strtod(t, &end);
    if (end == t || *end) { fprintf(stderr, "f90rt: bad real in list-directed input: '%s'\n", t); exit(2); }
    *v = x;
}

void f90_r_str(char *s, int len) {
    char t[1024];
    if (!next_token(t, sizeof t)) return;
    const int n = (int)strlen(t);
    for (int i = 0; i < len; ++i) s[i] = (i < n) ? t[i] : ' ';
}

void f90_r_end(void) {}

/* ---------------------------------------------------------------- clock, random numbers */

void f90_system_clock(int *count, int *count_rate, int *count_max) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    if (count) *count = (int)(((long long)ts.tv_sec * 1000 + ts.tv_nsec / 1000000) % 2147483647LL);
    if (count_rate) *count_rate = 1000;
    if (count_max) *count_max = 2147483647;
}

static uint64_t rng_state = 0x853c49e6748fea9bULL;
static int rng_env = -1;

static void rng_check_env(void) {
    if (rng_env >= 0) return;
    const char *s = getenv("PDEODE_SEED");
    rng_env = s ? 1 : 0;
    if (s) rng_state = strtoull(s, NULL, 10);
}

void f90_random_seed_size(int *n) { *n = 8; }

void f90_random_seed_put(const int *seed, long n) {
    rng_check_env();
    if (rng_env) return;                         /* the environment's seed wins */
    uint64_t h = 0x9E3779B97F4A7C15ULL;
    for (long i = 0; i < n; ++i) h = (h ^ (uint64_t)(uint32_t)seed[i]) * 0x100000001B3ULL;
    rng_state = h;
}

/* tests: restart the stream from a given seed (same as PDEODE_SEED=seed in a fresh process) */
void f90_rng_reseed(unsigned long long seed) { rng_env = 1; rng_state = seed; }

void f90_random_number(double *x) {
    rng_check_env();
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    *x = (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

#ifdef F90_MAIN
int main(void) { return f90_program(); }
#endif
