/*
 * ref_shim.h -- hooks compiled into the *patched* reference solver object (oracle/Makefile).
 * TEST INFRASTRUCTURE (oracle/).  The reference hard-codes EPS / MAX_ITER / OUT_ITER with
 * unconditional #defines (solver.c:3-9) and reports the residual only through printf
 * (solver.c:124, 135-139); the Makefile pipes solver.c through sed (no copy is written to disk)
 * so that those three constants read from here and printf goes through orc_ref_printf, which
 * records the full-precision doubles handed to it before formatting, and malloc() of the work vectors goes
 * through orc_ref_zmalloc (zero-filled).  Arithmetic is untouched.
 */
#ifndef ORACLE_REF_SHIM_H
#define ORACLE_REF_SHIM_H
#ifdef __cplusplus
extern "C" {
#endif
double orc_ref_eps(void);        /* default 1.0e-15  (solver.c:3) ; env REF_EPS      */
int    orc_ref_max_iter(void);   /* default 1000     (solver.c:4) ; env REF_MAX_ITER */
int    orc_ref_out_iter(void);   /* default 100      (solver.c:9) ; env REF_OUT_ITER */
int    orc_ref_printf(const char *fmt, ...);
#include <stddef.h>
void  *orc_ref_zmalloc(size_t bytes);   /* calloc: the solver's malloc'ed work vectors start at zero, which is what
                                          fresh mmap pages give the reference for vectors >= 128 KiB */

void   orc_ref_config(double eps, int max_iter, int out_iter, int quiet);
void   orc_ref_hist_reset(void);
int    orc_ref_hist_count(void);
int    orc_ref_hist_iter(int i);
double orc_ref_hist_res(int i);          /* sqrt(dot_r/dot_zero) exactly as passed to printf */
double orc_ref_final_res(void);
int    orc_ref_total_iter(void);
double orc_ref_total_time(void);
double orc_ref_avg_time(void);
#ifdef __cplusplus
}
#endif
#endif
