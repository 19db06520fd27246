/* ref_shim.c -- see ref_shim.h.  TEST INFRASTRUCTURE (oracle/). */
#include "ref_shim.h"
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static double g_eps = -1.0;
static int g_max_iter = -1, g_out_iter = -1, g_quiet = -1;

static int    *g_hk = NULL;
static double *g_hr = NULL;
static int g_hn = 0, g_hcap = 0;
static double g_final = 0.0, g_ttime = 0.0, g_atime = 0.0;
static int g_titer = -1;

double orc_ref_eps(void)
{
    if (g_eps < 0.0) { const char *e = getenv("REF_EPS"); g_eps = e ? atof(e) : 1.0e-15; }
    return g_eps;
}
int orc_ref_max_iter(void)
{
    if (g_max_iter < 0) { const char *e = getenv("REF_MAX_ITER"); g_max_iter = e ? atoi(e) : 1000; }
    return g_max_iter;
}
int orc_ref_out_iter(void)
{
    if (g_out_iter < 0) { const char *e = getenv("REF_OUT_ITER"); g_out_iter = e ? atoi(e) : 100; }
    return g_out_iter;
}
static int quiet(void)
{
    if (g_quiet < 0) { const char *e = getenv("REF_QUIET"); g_quiet = e ? atoi(e) : 0; }
    return g_quiet;
}
void orc_ref_config(double eps, int max_iter, int out_iter, int q)
{
    g_eps = eps; g_max_iter = max_iter; g_out_iter = out_iter; g_quiet = q;
}
void orc_ref_hist_reset(void) { g_hn = 0; g_titer = -1; g_final = g_ttime = g_atime = 0.0; }
int orc_ref_hist_count(void) { return g_hn; }
int orc_ref_hist_iter(int i) { return g_hk[i]; }
double orc_ref_hist_res(int i) { return g_hr[i]; }
double orc_ref_final_res(void) { return g_final; }
int orc_ref_total_iter(void) { return g_titer; }
double orc_ref_total_time(void) { return g_ttime; }
double orc_ref_avg_time(void) { return g_atime; }

void *orc_ref_zmalloc(size_t bytes) { return calloc(bytes ? bytes : 1, 1); }

int orc_ref_printf(const char *fmt, ...)
{
    va_list ap, aq;
    va_start(ap, fmt);
    va_copy(aq, ap);
    if (strncmp(fmt, "Iteration:", 10) == 0) {            /* solver.c:124 */
        int k = va_arg(aq, int);
        double res = va_arg(aq, double);
        if (g_hn == g_hcap) {
            g_hcap = g_hcap ? 2 * g_hcap : 1024;
            g_hk = (int *)realloc(g_hk, (size_t)g_hcap * sizeof(int));
            g_hr = (double *)realloc(g_hr, (size_t)g_hcap * sizeof(double));
        }
        g_hk[g_hn] = k; g_hr[g_hn] = res; ++g_hn;
    } else if (strncmp(fmt, "Total iter", 10) == 0) {      /* solver.c:135 */
        g_titer = va_arg(aq, int);
    } else if (strncmp(fmt, "Final r", 7) == 0) {          /* solver.c:136 */
        g_final = va_arg(aq, double);
    } else if (strncmp(fmt, "Total time", 10) == 0) {      /* solver.c:138 */
        g_ttime = va_arg(aq, double);
    } else if (strncmp(fmt, "Avg time/iter", 13) == 0) {   /* solver.c:139 */
        g_atime = va_arg(aq, double);
    }
    va_end(aq);
    int rc = 0;
    if (!quiet()) rc = vprintf(fmt, ap);
    va_end(ap);
    return rc;
}
