/*
 * bicg_oracle.c -- CPU restatement of the reference BiCGStab hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker.  The product (mpi-bicgstab_b200/) never links or calls this file.
 *
 * Parity pin: tests/test_oracle_golden.py runs this restatement side by side with the
 * reference's own sources compiled in place (oracle/_ref/libref_strict.so, recipe in
 * oracle/Makefile) and requires bit-identical histories under strict IEEE evaluation
 * (-ffp-contract=off); tests/golden/ holds histories produced by that reference build
 * (generator: tests/golden/make_golden.py) so the pin also travels to boxes without
 * /root/reference.
 *
 * What is restated (all citations are /root/reference/src/...):
 *   orc_rowsum / orc_spmv_part   <- mult()                     matrix.c:498-516
 *   orc_spmv                     <- MPI_csr_spmv_ovlap()       matrix.c:428-441
 *   orc_axpy/orc_dot/...         <- my_daxpy/my_ddot/my_dscal/my_dcopy   vector.c:3-27
 *   orc_bicgstab                 <- bicgstab()                 solver.c:35-146
 *   orc_ca_bicgstab              <- ca_bicgstab()              solver.c:160-278
 *   orc_pipe_bicgstab(_rr)       <- pipe_bicgstab()/pipe_bicgstab_rr()   solver.c:292-417, 433-576
 *   orc_partition                <- row partition rule         matrix.c:295-308
 *
 * Shape of the restatement: the reference runs P MPI ranks, each holding a row block split into a
 * diagonal part (local columns) and an off-diagonal part (global columns).  Here ONE process holds
 * the global CSR and *emulates* the P ranks:
 *   - a row's products are summed left to right over the entries whose column lies inside the
 *     owning rank's range first (the "diag" pass, matrix.c:437), then over the remaining entries
 *     (the "offd" pass, matrix.c:440), and the two partial sums are added to a zeroed y in that
 *     order (matrix.c:434-436, 514) -- this is exactly the association the reference produces;
 *   - every dot product is P sequential partial sums (vector.c:9-15) combined in rank order
 *     (MPI_SUM over MPI_IN_PLACE, solver.c:79 etc.; the MPI standard leaves the order to the
 *     implementation, rank order is what oracle/_ref's mini-MPI does too).
 * With P = 1 this degenerates to the single-rank reference bit for bit.
 *
 * Semantics the reference leaves undefined and this file pins: omega starts at 0
 * and the work vectors p, s, z, v start at 0 in the CA / pipelined variants (the reference reads
 * them uninitialised, solver.c:179->217, 310->352; zero is what a fresh mmap'd malloc gives).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int n;              /* global rows == cols */
    const double *val;  /* nnz values */
    const unsigned *col;/* nnz global column indices */
    const unsigned *ptr;/* n+1 row starts, ptr[0] == 0 */
    int P;              /* emulated rank count */
    int *first;         /* P+1 first global row of each rank (matrix.c:295-308) */
} orc_sys;

/* ---- partition rule: matrix.c:295-308 ---------------------------------------------------- */
void orc_partition(int n, int P, int *counts, int *displs)
{
    int base = n / P, extra = n % P;
    for (int p = 0; p < P; ++p) {
        int lo = p * base + (p < extra ? p : extra);
        counts[p] = base + (p < extra ? 1 : 0);
        displs[p] = lo;
    }
}

/* ---- BLAS-1: vector.c:3-27 (sequential, scalar accumulator) ------------------------------ */
static void orc_axpy(int n, double a, const double *x, double *y)
{
    for (int i = 0; i < n; ++i) y[i] += a * x[i];
}
static void orc_scal(int n, double a, double *x)
{
    for (int i = 0; i < n; ++i) x[i] *= a;
}
static void orc_copy(int n, const double *x, double *y)
{
    for (int i = 0; i < n; ++i) y[i] = x[i];
}
/* dot over the whole (global) vector, emulating P per-rank partial sums added in rank order */
static double orc_dot(const orc_sys *S, const double *x, const double *y)
{
    double total = 0.0;
    for (int p = 0; p < S->P; ++p) {
        double part = 0.0;
        for (int i = S->first[p]; i < S->first[p + 1]; ++i) part += x[i] * y[i];
        if (p == 0) total = part; else total += part;
    }
    return total;
}

/* exported single-rank helpers so the tests can check kernels one by one */
double orc_ddot(int n, const double *x, const double *y)
{
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += x[i] * y[i];
    return s;
}
void orc_daxpy(int n, double a, const double *x, double *y) { orc_axpy(n, a, x, y); }
void orc_dscal(int n, double a, double *x) { orc_scal(n, a, x); }

/* ---- SpMV: matrix.c:428-441 + 498-516 ------------------------------------------------------ */
static void orc_spmv_sys(const orc_sys *S, const double *x, double *y)
{
    for (int p = 0; p < S->P; ++p) {
        unsigned lo = (unsigned)S->first[p], hi = (unsigned)S->first[p + 1];
        for (unsigned i = lo; i < hi; ++i) {
            double acc_own = 0.0, acc_far = 0.0;
            /* diag pass: columns owned by rank p, in stored order (matrix.c:437) */
            for (unsigned j = S->ptr[i]; j < S->ptr[i + 1]; ++j) {
                unsigned c = S->col[j];
                if (c >= lo && c < hi) acc_own += S->val[j] * x[c];
            }
            /* offd pass: every other column (matrix.c:440) */
            for (unsigned j = S->ptr[i]; j < S->ptr[i + 1]; ++j) {
                unsigned c = S->col[j];
                if (c < lo || c >= hi) acc_far += S->val[j] * x[c];
            }
            double yi = 0.0;       /* matrix.c:434-436 */
            yi += acc_own;         /* matrix.c:514, first call  */
            yi += acc_far;         /* matrix.c:514, second call */
            y[i] = yi;
        }
    }
}

static void orc_sys_init(orc_sys *S, int n, const double *val, const unsigned *col,
                         const unsigned *ptr, int P)
{
    S->n = n; S->val = val; S->col = col; S->ptr = ptr; S->P = P;
    S->first = (int *)malloc((size_t)(P + 1) * sizeof(int));
    int *cnt = (int *)malloc((size_t)P * sizeof(int));
    orc_partition(n, P, cnt, S->first);
    S->first[P] = n;
    free(cnt);
}
static void orc_sys_free(orc_sys *S) { free(S->first); }

/* y = A x for the P-rank emulation (exported) */
void orc_spmv(int n, const double *val, const unsigned *col, const unsigned *ptr, int P,
              const double *x, double *y)
{
    orc_sys S; orc_sys_init(&S, n, val, col, ptr, P);
    orc_spmv_sys(&S, x, y);
    orc_sys_free(&S);
}

/* ---- common solver scaffolding ------------------------------------------------------------- */
typedef struct {
    double tol;        /* EPS, solver.c:3  */
    int max_iter;      /* MAX_ITER, solver.c:4 */
    double *hist;      /* optional: hist[k] = dot_r/dot_zero after iteration k (hist[0] = 1) */
    int hist_cap;
} orc_opts;

static double *vnew(int n) { return (double *)calloc((size_t)n, sizeof(double)); }

static void hist_put(const orc_opts *o, int k, double dot_r, double dot_zero)
{
    if (o->hist && k < o->hist_cap) o->hist[k] = dot_r / dot_zero;
}

/* r <- b - A x0, r_hat <- r, returns (r,r): the shared opening of all four solvers
 * (solver.c:74-79, 200-203, 333-336, 476-479) */
static double orc_open(const orc_sys *S, const double *x, double *r, double *Ax, double *r_hat)
{
    orc_spmv_sys(S, x, Ax);
    orc_axpy(S->n, -1.0, Ax, r);
    orc_copy(S->n, r, r_hat);
    return orc_dot(S, r, r);
}

/* ---- bicgstab: solver.c:35-146 ------------------------------------------------------------- */
int orc_bicgstab(int n, const double *val, const unsigned *col, const unsigned *ptr, int P,
                 double *x, double *r, double tol, int max_iter, double *hist, int hist_cap)
{
    orc_sys S; orc_sys_init(&S, n, val, col, ptr, P);
    orc_opts o = { tol, max_iter, hist, hist_cap };
    double *Ax = vnew(n), *rh = vnew(n), *s = vnew(n), *y = vnew(n), *p = vnew(n);

    double rTr = orc_open(&S, x, r, Ax, rh);
    orc_copy(n, r, p);                                   /* solver.c:77 */
    double dot_r = rTr, dot_zero = rTr;                  /* solver.c:82-83 */
    hist_put(&o, 0, dot_r, dot_zero);
    int k = 0;
    while (dot_r > tol * tol * dot_zero && k < max_iter) {     /* solver.c:86 */
        orc_spmv_sys(&S, p, s);                                 /* :88  s = A p */
        double rTs = orc_dot(&S, rh, s);                        /* :89  */
        double alpha = rTr / rTs;                               /* :93  */
        orc_axpy(n, -alpha, s, r);                              /* :94  q (kept in r) */
        orc_spmv_sys(&S, r, y);                                 /* :96  y = A q */
        double rTy = orc_dot(&S, r, y);                         /* :97  */
        double yTy = orc_dot(&S, y, y);                         /* :99  */
        double omega = rTy / yTy;                               /* :104 */
        orc_axpy(n, alpha, p, x);                               /* :105 */
        orc_axpy(n, omega, r, x);                               /* :106 */
        orc_axpy(n, -omega, y, r);                              /* :107 */
        dot_r = orc_dot(&S, r, r);                              /* :108 */
        double rTr_old = rTr;                                   /* :110 */
        rTr = orc_dot(&S, rh, r);                               /* :111 */
        double beta = (alpha / omega) * (rTr / rTr_old);        /* :116 */
        orc_scal(n, beta, p);                                   /* :117 */
        orc_axpy(n, 1.0, r, p);                                 /* :118 */
        orc_axpy(n, -beta * omega, s, p);                       /* :119 */
        ++k;
        hist_put(&o, k, dot_r, dot_zero);
    }
    free(Ax); free(rh); free(s); free(y); free(p);
    orc_sys_free(&S);
    return k;
}

/* ---- ca_bicgstab: solver.c:160-278 --------------------------------------------------------- */
int orc_ca_bicgstab(int n, const double *val, const unsigned *col, const unsigned *ptr, int P,
                    double *x, double *r, double tol, int max_iter, double *hist, int hist_cap)
{
    orc_sys S; orc_sys_init(&S, n, val, col, ptr, P);
    orc_opts o = { tol, max_iter, hist, hist_cap };
    double *Ax = vnew(n), *rh = vnew(n), *s = vnew(n), *z = vnew(n), *w = vnew(n), *p = vnew(n);

    double rTr = orc_open(&S, x, r, Ax, rh);
    orc_spmv_sys(&S, r, w);                              /* :205 w = A r */
    double rTw = orc_dot(&S, r, w);                      /* :206 */
    double alpha = rTr / rTw, beta = 0.0, omega = 0.0;   /* :210-211 (omega pinned to 0) */
    double dot_r = rTr, dot_zero = rTr;
    hist_put(&o, 0, dot_r, dot_zero);
    int k = 0;
    while (dot_r > tol * tol * dot_zero && k < max_iter) {
        orc_axpy(n, -omega, s, p); orc_scal(n, beta, p); orc_axpy(n, 1.0, r, p);   /* :217-219 */
        orc_axpy(n, -omega, z, s); orc_scal(n, beta, s); orc_axpy(n, 1.0, w, s);   /* :220-222 */
        orc_spmv_sys(&S, s, z);                                  /* :224 z = A s */
        orc_axpy(n, -alpha, s, r);                               /* :225 q */
        orc_axpy(n, -alpha, z, w);                               /* :226 y */
        double qy = orc_dot(&S, r, w);                           /* :227 */
        double yy = orc_dot(&S, w, w);                           /* :228 */
        omega = qy / yy;                                         /* :232 */
        orc_axpy(n, alpha, p, x);                                /* :233 */
        orc_axpy(n, omega, r, x);                                /* :234 */
        orc_axpy(n, -omega, w, r);                               /* :235 */
        dot_r = orc_dot(&S, r, r);                               /* :236 */
        orc_spmv_sys(&S, r, w);                                  /* :238 w = A r */
        double rTr_old = rTr;
        rTr = orc_dot(&S, rh, r);                                /* :240 */
        rTw = orc_dot(&S, rh, w);                                /* :241 */
        double rTs = orc_dot(&S, rh, s);                         /* :242 */
        double rTz = orc_dot(&S, rh, z);                         /* :243 */
        beta = (alpha / omega) * (rTr / rTr_old);                /* :248 */
        alpha = rTr / (rTw + beta * (rTs - omega * rTz));        /* :249 */
        ++k;
        hist_put(&o, k, dot_r, dot_zero);
    }
    free(Ax); free(rh); free(s); free(z); free(w); free(p);
    orc_sys_free(&S);
    return k;
}

/* ---- pipe_bicgstab / pipe_bicgstab_rr: solver.c:292-417, 433-576 -----------------------------
 * krr <= 0 selects the plain pipelined variant (no replacement branch exists there). */
int orc_pipe_bicgstab_rr(int n, const double *val, const unsigned *col, const unsigned *ptr, int P,
                         double *x, double *r, int krr, int nrr,
                         double tol, int max_iter, double *hist, int hist_cap)
{
    orc_sys S; orc_sys_init(&S, n, val, col, ptr, P);
    orc_opts o = { tol, max_iter, hist, hist_cap };
    double *b = vnew(n), *Ax = vnew(n), *rh = vnew(n), *s = vnew(n), *z = vnew(n), *w = vnew(n),
           *p = vnew(n), *v = vnew(n), *t = vnew(n);
    const int with_rr = krr > 0;

    if (with_rr) orc_copy(n, r, b);                      /* :475 */
    double rTr = orc_open(&S, x, r, Ax, rh);
    orc_spmv_sys(&S, r, w);                              /* :338 / :481 */
    double rTw = orc_dot(&S, r, w);
    orc_spmv_sys(&S, w, t);                              /* :341 / :484 */
    double alpha = rTr / rTw, beta = 0.0, omega = 0.0;
    double dot_r = rTr, dot_zero = rTr;
    hist_put(&o, 0, dot_r, dot_zero);
    int k = 0;
    while (dot_r > tol * tol * dot_zero && k < max_iter) {
        int replace = with_rr && (k % krr == 0) && k > 0 && k <= krr * nrr;      /* :498, :522 */
        orc_axpy(n, -omega, s, p); orc_scal(n, beta, p); orc_axpy(n, 1.0, r, p);  /* :352-354 */
        if (replace) {
            orc_spmv_sys(&S, p, s);                              /* :499 */
            orc_spmv_sys(&S, s, z);                              /* :500 */
        } else {
            orc_axpy(n, -omega, z, s); orc_scal(n, beta, s); orc_axpy(n, 1.0, w, s);  /* :355-357 */
            orc_axpy(n, -omega, v, z); orc_scal(n, beta, z); orc_axpy(n, 1.0, t, z);  /* :358-360 */
        }
        orc_axpy(n, -alpha, s, r);                               /* :361 q */
        orc_axpy(n, -alpha, z, w);                               /* :362 y */
        double qy = orc_dot(&S, r, w);                           /* :363 */
        double yy = orc_dot(&S, w, w);                           /* :364 */
        orc_spmv_sys(&S, z, v);                                  /* :365 v = A z */
        omega = qy / yy;                                         /* :369 */
        orc_axpy(n, alpha, p, x);                                /* :370 */
        orc_axpy(n, omega, r, x);                                /* :371 */
        if (replace) {
            orc_spmv_sys(&S, x, Ax);                             /* :523 */
            orc_copy(n, b, r);                                   /* :524 */
            orc_axpy(n, -1.0, Ax, r);                            /* :525 */
            orc_spmv_sys(&S, r, w);                              /* :526 */
        } else {
            orc_axpy(n, -omega, w, r);                           /* :372 */
            orc_axpy(n, -alpha, v, t);                           /* :374 */
            orc_axpy(n, -omega, t, w);                           /* :375 */
        }
        dot_r = orc_dot(&S, r, r);                               /* :373 / :533 */
        double rTr_old = rTr;
        rTr = orc_dot(&S, rh, r);                                /* :377 */
        rTw = orc_dot(&S, rh, w);                                /* :378 */
        double rTs = orc_dot(&S, rh, s);                         /* :379 */
        double rTz = orc_dot(&S, rh, z);                         /* :380 */
        orc_spmv_sys(&S, w, t);                                  /* :381 t = A w */
        beta = (alpha / omega) * (rTr / rTr_old);                /* :387 */
        alpha = rTr / (rTw + beta * (rTs - omega * rTz));        /* :388 */
        ++k;
        hist_put(&o, k, dot_r, dot_zero);
    }
    free(b); free(Ax); free(rh); free(s); free(z); free(w); free(p); free(v); free(t);
    orc_sys_free(&S);
    return k;
}

int orc_pipe_bicgstab(int n, const double *val, const unsigned *col, const unsigned *ptr, int P,
                      double *x, double *r, double tol, int max_iter, double *hist, int hist_cap)
{
    return orc_pipe_bicgstab_rr(n, val, col, ptr, P, x, r, 0, 0, tol, max_iter, hist, hist_cap);
}

/* ---- long-double SpMV, used by the kernel-level (K-level) tolerance tests ------------------- */
void orc_spmv_ld(int n, const double *val, const unsigned *col, const unsigned *ptr,
                 const double *x, double *y)
{
    for (int i = 0; i < n; ++i) {
        long double acc = 0.0L;
        for (unsigned j = ptr[i]; j < ptr[i + 1]; ++j) acc += (long double)val[j] * (long double)x[col[j]];
        y[i] = (double)acc;
    }
}

/* ---- shifted_lopbicg_switching: shifted_switching_solver.c:260-602 ------------------------------
 * Seed-switching shifted BiCGStab(1): one seed system (A + sigma[seed] I) x = b is iterated with BiCGStab; every
 * other shift j is advanced from the seed's Krylov data with collinear-residual recurrences (eta, pi, zeta), and when the
 * seed converges before the others the slowest remaining shift becomes the new seed (its scalars are re-derived from
 * the archived alpha / beta / omega / pi history).  Restated line by line; the P-rank emulation of the dots and of the
 * SpMV is the same as above.  x_set: sigma_len blocks of n (in: initial guess, reference drivers pass zeros), r: b in,
 * seed residual out.  Returns the reference's return value (k, one more than the iterations performed);
 * hist[k] = dot_r / dot_zero after iteration k, seed_out = final seed, stop_iter[j] = iteration at which shift j stopped
 * (0 if it never did).  EPS 1e-12, MAX_ITER 1000 in the reference (:5-6). */
int orc_shifted_lopbicg_switching(int n, const double *val, const unsigned *col, const unsigned *ptr, int P,
                                  double *x_set, double *r, const double *sigma, int sigma_len, int seed,
                                  double tol, int max_iter_opt, double *hist, int hist_cap, int *seed_out, int *stop_iter)
{
    orc_sys S; orc_sys_init(&S, n, val, col, ptr, P);
    int i, j;
    int k = 1, max_iter = max_iter_opt + 1, stop_count = 0;             /* :291-294 */
    double max_zeta_pi, abs_zeta_pi;
    int max_sigma = seed;
    double *r_old = vnew(n), *r_hat = vnew(n), *s = vnew(n), *y = vnew(n), *q_copy = vnew(n);
    double *p_set = (double *)calloc((size_t)n * (size_t)sigma_len, sizeof(double));           /* :304 */
    double *alpha_set = vnew(sigma_len), *beta_set = vnew(sigma_len), *omega_set = vnew(sigma_len),
           *eta_set = vnew(sigma_len), *zeta_set = vnew(sigma_len);
    double *alpha_arch = vnew(max_iter), *beta_arch = vnew(max_iter), *omega_arch = vnew(max_iter);
    double *pi_arch = (double *)calloc((size_t)max_iter * (size_t)sigma_len, sizeof(double));
    char *stop_flag = (char *)calloc((size_t)sigma_len, 1);
    double dot_r, dot_zero, rTr, rTs, qTq, qTy, rTr_old;
#define PSET(jj) (p_set + (size_t)(jj) * (size_t)n)
#define XSET(jj) (x_set + (size_t)(jj) * (size_t)n)
#define PI(jj, kk) pi_arch[(size_t)(jj) * (size_t)max_iter + (size_t)(kk)]

    rTr = orc_dot(&S, r, r);                                            /* :342 */
    orc_copy(n, r, r_hat);                                              /* :344 */
    for (i = 0; i < sigma_len; i++) {                                   /* :345-353 */
        orc_copy(n, r, PSET(i));
        alpha_set[i] = 1.0; beta_set[i] = 0.0; eta_set[i] = 0.0;
        PI(i, 0) = 1.0; PI(i, 1) = 1.0;
        zeta_set[i] = 1.0;
    }
    orc_copy(n, r, PSET(seed));                                         /* :354 */
    dot_r = rTr; dot_zero = rTr; max_zeta_pi = 1.0;                     /* :357-359 */
    alpha_arch[0] = 1.0; beta_arch[0] = 0.0;                            /* :361-362 */
    if (hist && hist_cap > 0) hist[0] = 1.0;
    if (stop_iter) for (j = 0; j < sigma_len; j++) stop_iter[j] = 0;

    while (stop_count < sigma_len && k < max_iter) {                    /* :372 */
        orc_copy(n, r, r_old);                                          /* :374 */
        orc_spmv_sys(&S, PSET(seed), s);                                /* :377-384  s <- A p[seed]            */
        orc_axpy(n, sigma[seed], PSET(seed), s);                        /* :386      s <- s + sigma[seed] p    */
        rTs = orc_dot(&S, r_hat, s);                                    /* :387 */
        alpha_arch[k] = rTr / rTs;                                      /* :390 */
        orc_axpy(n, -alpha_arch[k], s, r);                              /* :391  q <- r - alpha s */
        orc_copy(n, r, q_copy);                                         /* :392 */
        orc_spmv_sys(&S, r, y);                                         /* :395-402  y <- A q */
        orc_axpy(n, sigma[seed], r, y);                                 /* :404 */
        qTq = orc_dot(&S, r, r);                                        /* :405 */
        qTy = orc_dot(&S, r, y);                                        /* :406 */
        omega_arch[k] = qTq / qTy;                                      /* :410 */
        orc_axpy(n, alpha_arch[k], PSET(seed), XSET(seed));             /* :411 */
        orc_axpy(n, omega_arch[k], r, XSET(seed));                      /* :412 */
        orc_axpy(n, -omega_arch[k], y, r);                              /* :413 */
        dot_r = orc_dot(&S, r, r);                                      /* :414 */
        rTr_old = rTr;                                                  /* :415 */
        rTr = orc_dot(&S, r_hat, r);                                    /* :416 */
        beta_arch[k] = (alpha_arch[k] / omega_arch[k]) * (rTr / rTr_old);   /* :420 */
        orc_scal(n, beta_arch[k], PSET(seed));                          /* :421 */
        orc_axpy(n, 1.0, r, PSET(seed));                                /* :422 */
        orc_axpy(n, -beta_arch[k] * omega_arch[k], s, PSET(seed));      /* :423 */

        for (j = 0; j < sigma_len; j++) {                               /* :429-446 */
            if (j == seed) continue;
            if (stop_flag[j]) continue;
            eta_set[j] = (beta_arch[k - 1] / alpha_arch[k - 1]) * alpha_arch[k] * eta_set[j]
                       - (sigma[seed] - sigma[j]) * alpha_arch[k] * PI(j, k - 1);
            PI(j, k) = eta_set[j] + PI(j, k - 1);
            alpha_set[j] = (PI(j, k - 1) / PI(j, k)) * alpha_arch[k];
            omega_set[j] = omega_arch[k] / (1.0 - omega_arch[k] * (sigma[seed] - sigma[j]));
            orc_axpy(n, omega_set[j] / (PI(j, k) * zeta_set[j]), q_copy, XSET(j));
            orc_axpy(n, alpha_set[j], PSET(j), XSET(j));
            orc_axpy(n, omega_set[j] / (alpha_set[j] * zeta_set[j] * PI(j, k)), q_copy, PSET(j));
            orc_axpy(n, -omega_set[j] / (alpha_set[j] * zeta_set[j] * PI(j, k - 1)), r_old, PSET(j));
            zeta_set[j] = (1.0 - omega_arch[k] * (sigma[seed] - sigma[j])) * zeta_set[j];
            beta_set[j] = (PI(j, k - 1) / PI(j, k)) * (PI(j, k - 1) / PI(j, k)) * beta_arch[k];
            orc_scal(n, beta_set[j], PSET(j));
            orc_axpy(n, 1.0 / (PI(j, k) * zeta_set[j]), r, PSET(j));
        }

        max_zeta_pi = 1.0;                                              /* :451-476 */
        for (j = 0; j < sigma_len; j++) {
            if (stop_flag[j]) continue;
            if (j == seed) abs_zeta_pi = 1.0;
            else abs_zeta_pi = fabs(1.0 / (zeta_set[j] * PI(j, k)));
            if (abs_zeta_pi * abs_zeta_pi * dot_r <= tol * tol * dot_zero) {
                stop_flag[j] = 1; stop_count++;
                if (stop_iter) stop_iter[j] = k;
            } else if (abs_zeta_pi > max_zeta_pi) {
                max_zeta_pi = abs_zeta_pi; max_sigma = j;
            }
        }

        if (stop_flag[seed] && stop_count < sigma_len) {                /* :490-527 seed switching */
            for (i = 1; i <= k; i++) {
                alpha_arch[i] = (PI(max_sigma, i - 1) / PI(max_sigma, i)) * alpha_arch[i];
                beta_arch[i] = (PI(max_sigma, i - 1) / PI(max_sigma, i)) * (PI(max_sigma, i - 1) / PI(max_sigma, i)) * beta_arch[i];
                omega_arch[i] = omega_arch[i] / (1.0 - omega_arch[i] * (sigma[seed] - sigma[max_sigma]));
            }
            orc_scal(n, 1.0 / (zeta_set[max_sigma] * PI(max_sigma, k)), r);
            for (j = 0; j < sigma_len; j++) { eta_set[j] = 0.0; zeta_set[j] = 1.0; }
            for (i = 1; i <= k; i++) {
                for (j = 0; j < sigma_len; j++) {
                    if (stop_flag[j]) continue;
                    if (j == max_sigma) continue;
                    eta_set[j] = (beta_arch[i - 1] / alpha_arch[i - 1]) * alpha_arch[i] * eta_set[j]
                               - (sigma[max_sigma] - sigma[j]) * alpha_arch[i] * PI(j, i - 1);
                    PI(j, i) = eta_set[j] + PI(j, i - 1);
                    zeta_set[j] = (1.0 - omega_arch[i] * (sigma[max_sigma] - sigma[j])) * zeta_set[j];
                }
            }
            seed = max_sigma;
        }
        if (hist && k < hist_cap) hist[k] = dot_r / dot_zero;
        k++;                                                            /* :537 */
    }
    if (seed_out) *seed_out = seed;
    free(r_old); free(r_hat); free(s); free(y); free(q_copy); free(p_set);
    free(alpha_set); free(beta_set); free(omega_set); free(eta_set); free(zeta_set);
    free(alpha_arch); free(beta_arch); free(omega_arch); free(pi_arch); free(stop_flag);
    orc_sys_free(&S);
#undef PSET
#undef XSET
#undef PI
    return k;
}
