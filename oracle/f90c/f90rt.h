/* f90rt.h -- runtime of the Fortran-to-C translation (oracle/f90c/f90toc.py).
 * TEST INFRASTRUCTURE: only what the reference's three source files use. */
#ifndef F90RT_H
#define F90RT_H

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define F90_MAX(a, b) ((a) > (b) ? (a) : (b))
#define F90_MIN(a, b) ((a) < (b) ? (a) : (b))

int f90_ipow(int x, int n);
void *f90_calloc(long n, long size);
void f90_stop(void);
void f90_flush_all(void);

/* OPEN / CLOSE.  Names are blank-padded Fortran strings (pointer, length). */
void f90_open(int unit, const char *file, int filelen, const char *status, int statuslen,
              const char *action, int actionlen, const char *position, int positionlen,
              int *iostat);
void f90_close(int unit, const char *status, int statuslen);

/* WRITE: unit -1 = "*"; fmt NULL = list-directed */
void f90_w_begin(int unit, const char *fmt, int no_advance);
void f90_w_int(int v);
void f90_w_real(double v);
void f90_w_str(const char *s, int len);
void f90_w_end(void);

/* READ, list-directed only */
void f90_r_begin(int unit);
void f90_r_int(int *v);
void f90_r_real(double *v);
void f90_r_str(char *s, int len);
void f90_r_end(void);

void f90_system_clock(int *count, int *count_rate, int *count_max);
void f90_random_seed_size(int *n);
void f90_random_seed_put(const int *seed, long n);
void f90_random_number(double *x);

void f90_close_all(void);
void f90_rng_reseed(unsigned long long seed);

int f90_program(void);

#endif
