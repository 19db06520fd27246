// shifted.cu -- shifted_lopbicg_switching (shifted_switching_solver.c:260-602; prototype shifted_switching_solver.h:12):
// seed-switching shifted BiCGStab for (A + sigma_j I) x_j = b, j = 0 .. sigma_len - 1, behind the reference's own entry point.
//
// One seed system is iterated with BiCGStab on the arena vectors with the same fused SpMV (+ sigma_seed x in the
// epilogue) and reduction tails as the un-shifted solvers; every other shift is advanced from the seed's Krylov data:
//   * the per-shift scalar recurrences (eta, pi, zeta, alpha_j, omega_j, beta_j; :431-445), the convergence tests
//     (:451-476) and the seed switch (:490-527, history of alpha / beta / omega / pi re-derived for the new seed) run in
//     ONE small kernel per iteration (sh_scalar_iter), entirely on the device;
//   * the six daxpy / dscal passes per shift and iteration of the reference (:435-445: up to 512 shifts x 6 passes over
//     length-n vectors) are ONE multi-vector kernel (sh_vec_shift): q, r_old and r are loaded once per row, then x_j and
//     p_j of every active shift are read and written exactly once -- 32 B per row and shift, the HBM floor of the method.
// Element-wise operation order = the reference's call order with gcc's FMA contraction (y += a x -> fma(a, x, y)).
// The host only enqueues batches of iterations and polls a done flag (as solve.cu does).
#include "engine.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace bicg {

namespace {

constexpr int SH_COEF = 6;          // per active shift: c1, alpha_j, c2, c3, beta_j, c4
constexpr int SH_EVENTS = 16;       // seed switches remembered for the reference's printf lines

struct ShiftDev {
    int L, max_iter;                // sigma_len, MAX_ITER + 1                                    (:291-293)
    double tol;                     // EPS                                                       (:292)
    int seed, k, stop_count, done, live, switched, max_sigma, n_active, n_events;
    double rTr, rTs, qTq, qTy, dot_r, dot_zero, rTr_old, r_scale;
    double sigma_seed;              // read by the SpMV epilogue
    double *sigma, *alpha_set, *beta_set, *omega_set, *eta_set, *zeta_set;      // [L]
    unsigned char *stop_flag;                                                  // [L]
    int *stop_iter;                                                             // [L]
    double *alpha_arch, *beta_arch, *omega_arch;                                // [max_iter]
    double *pi_arch;                                                            // [L][max_iter]
    double *coef;                                                               // [L][SH_COEF], compacted with `active`
    int *active;                                                                // [L]
    double *hist;                                                               // [max_iter + 1]
    int *ev_k, *ev_seed, *ev_remain;                                            // [SH_EVENTS]
    double *ev_vals;                                                            // [SH_EVENTS][L][3]  eta, pi, zeta (:518)
};

#define PI(j, kk) sd->pi_arch[(size_t)(j) * (size_t)sd->max_iter + (size_t)(kk)]

// ---- scalar kernels ------------------------------------------------------------------------------------------------
__global__ void sh_scalar_init(ShiftDev *sd, Scalars *sc)                           // :342-362
{
    const int t = threadIdx.x;
    if (t == 0) {
        sd->rTr = sc->pend[0]; sd->dot_r = sd->rTr; sd->dot_zero = sd->rTr;
        sd->k = 1; sd->stop_count = 0; sd->done = 0; sd->live = 0; sd->switched = 0; sd->n_events = 0; sd->max_sigma = sd->seed;
        sd->alpha_arch[0] = 1.0; sd->beta_arch[0] = 0.0;
        sd->sigma_seed = sd->sigma[sd->seed];
        sd->hist[0] = 1.0;
        if (!(sd->k < sd->max_iter) || sd->L <= 0) sd->done = 1;
    }
    for (int j = t; j < sd->L; j += blockDim.x) {
        sd->alpha_set[j] = 1.0; sd->beta_set[j] = 0.0; sd->eta_set[j] = 0.0; sd->zeta_set[j] = 1.0;
        PI(j, 0) = 1.0; PI(j, 1) = 1.0;
        sd->stop_flag[j] = 0; sd->stop_iter[j] = 0;
    }
}
__global__ void sh_scalar_alpha(ShiftDev *sd, Scalars *sc)                          // :387-390
{
    sd->live = !sd->done;
    if (sd->done) return;
    sd->switched = 0;
    sd->rTs = sc->pend[0];
    const double al = sd->rTr / sd->rTs;
    sd->alpha_arch[sd->k] = al;
    sc->alpha = al;
}
__global__ void sh_scalar_omega(ShiftDev *sd, Scalars *sc)                          // :405-410
{
    if (sd->done) return;
    sd->qTq = sc->pend[0]; sd->qTy = sc->pend[1];
    const double om = sd->qTq / sd->qTy;
    sd->omega_arch[sd->k] = om;
    sc->omega = om;
}
// beta, the shifts' recurrences, convergence tests, seed switch, loop test: one block
__global__ void __launch_bounds__(512) sh_scalar_iter(ShiftDev *sd, Scalars *sc)
{
    if (sd->done) return;
    const int t = threadIdx.x, T = blockDim.x;
    const int k = sd->k, L = sd->L;
    __shared__ int s_n;
    if (t == 0) {
        sd->dot_r = sc->pend[0];                                                    // :414
        sd->rTr_old = sd->rTr;                                                      // :415
        sd->rTr = sc->pend[1];                                                      // :416
        const double be = (sd->alpha_arch[k] / sd->omega_arch[k]) * (sd->rTr / sd->rTr_old);      // :420
        sd->beta_arch[k] = be;
        sc->beta = be;
        s_n = 0;
    }
    __syncthreads();
    const int seed = sd->seed;
    const double al_k = sd->alpha_arch[k], om_k = sd->omega_arch[k], be_k = sd->beta_arch[k];
    const double al_o = sd->alpha_arch[k - 1], be_o = sd->beta_arch[k - 1], sg_s = sd->sigma[seed];
    for (int j = t; j < L; j += T) {                                                // :429-446 (scalars; vectors: sh_vec_shift)
        if (j == seed || sd->stop_flag[j]) continue;
        const double pi_o = PI(j, k - 1), zeta_o = sd->zeta_set[j];
        const double eta = (be_o / al_o) * al_k * sd->eta_set[j] - (sg_s - sd->sigma[j]) * al_k * pi_o;
        const double pi_n = eta + pi_o;
        const double al_j = (pi_o / pi_n) * al_k;
        const double om_j = om_k / (1.0 - om_k * (sg_s - sd->sigma[j]));
        const double c1 = om_j / (pi_n * zeta_o);
        const double c2 = om_j / (al_j * zeta_o * pi_n);
        const double c3 = -om_j / (al_j * zeta_o * pi_o);
        const double zeta_n = (1.0 - om_k * (sg_s - sd->sigma[j])) * zeta_o;
        const double be_j = (pi_o / pi_n) * (pi_o / pi_n) * be_k;
        const double c4 = 1.0 / (pi_n * zeta_n);
        sd->eta_set[j] = eta; PI(j, k) = pi_n; sd->alpha_set[j] = al_j; sd->omega_set[j] = om_j;
        sd->zeta_set[j] = zeta_n; sd->beta_set[j] = be_j;
        const int slot = atomicAdd(&s_n, 1);
        sd->active[slot] = j;
        double *c = sd->coef + (size_t)slot * SH_COEF;
        c[0] = c1; c[1] = al_j; c[2] = c2; c[3] = c3; c[4] = be_j; c[5] = c4;
    }
    __syncthreads();
    if (t == 0) {
        sd->n_active = s_n;
        double max_zeta_pi = 1.0;                                                   // :451-476
        for (int j = 0; j < L; ++j) {
            if (sd->stop_flag[j]) continue;
            const double azp = (j == seed) ? 1.0 : fabs(1.0 / (sd->zeta_set[j] * PI(j, k)));
            if (azp * azp * sd->dot_r <= sd->tol * sd->tol * sd->dot_zero) {
                sd->stop_flag[j] = 1; sd->stop_count += 1; sd->stop_iter[j] = k;
            } else if (azp > max_zeta_pi) {
                max_zeta_pi = azp; sd->max_sigma = j;
            }
        }
    }
    __syncthreads();
    const bool sw = sd->stop_flag[seed] && sd->stop_count < L;                      // :490
    if (sw) {
        const int ms = sd->max_sigma;
        const double dsg = sg_s - sd->sigma[ms];
        for (int i = 1 + t; i <= k; i += T) {                                       // :494-498
            const double ratio = PI(ms, i - 1) / PI(ms, i);
            sd->alpha_arch[i] = ratio * sd->alpha_arch[i];
            sd->beta_arch[i] = ratio * ratio * sd->beta_arch[i];
            sd->omega_arch[i] = sd->omega_arch[i] / (1.0 - sd->omega_arch[i] * dsg);
        }
        if (t == 0) sd->r_scale = 1.0 / (sd->zeta_set[ms] * PI(ms, k));            // :499
        __syncthreads();
        for (int j = t; j < L; j += T) { sd->eta_set[j] = 0.0; sd->zeta_set[j] = 1.0; }      // :501-505
        __syncthreads();
        const double sg_m = sd->sigma[ms];
        for (int j = t; j < L; j += T) {                                            // :509-517
            if (sd->stop_flag[j] || j == ms) continue;
            double eta = 0.0, zeta = 1.0;
            for (int i = 1; i <= k; ++i) {
                eta = (sd->beta_arch[i - 1] / sd->alpha_arch[i - 1]) * sd->alpha_arch[i] * eta - (sg_m - sd->sigma[j]) * sd->alpha_arch[i] * PI(j, i - 1);
                PI(j, i) = eta + PI(j, i - 1);
                zeta = (1.0 - sd->omega_arch[i] * (sg_m - sd->sigma[j])) * zeta;
            }
            sd->eta_set[j] = eta; sd->zeta_set[j] = zeta;
        }
        __syncthreads();
        if (sd->n_events < SH_EVENTS) {                                             // what the reference prints (:519-526)
            const int e = sd->n_events;
            for (int j = t; j < L; j += T) {
                double *v = sd->ev_vals + ((size_t)e * L + j) * 3;
                const bool shown = !(sd->stop_flag[j] || j == ms);
                v[0] = shown ? sd->eta_set[j] : NAN; v[1] = PI(j, k); v[2] = sd->zeta_set[j];
            }
        }
        __syncthreads();
        if (t == 0) {
            if (sd->n_events < SH_EVENTS) {
                sd->ev_k[sd->n_events] = k; sd->ev_seed[sd->n_events] = ms; sd->ev_remain[sd->n_events] = L - sd->stop_count;
                sd->n_events += 1;
            }
            sd->seed = ms; sd->sigma_seed = sd->sigma[ms]; sd->switched = 1;        // :524
        }
    }
    __syncthreads();
    if (t == 0) {
        sd->hist[k] = sd->dot_r / sd->dot_zero;
        sd->k = k + 1;                                                              // :537
        if (!(sd->stop_count < L && sd->k < sd->max_iter)) { sd->done = 1; sc->done = 1; }      // :372 (sc->done stops the shared kernels)
    }
}

// ---- vector kernels -----------------------------------------------------------------------------------------------
struct ShVec {
    KernelCommon kc;
    const ShiftDev *sd;
    double *r, *rh, *p, *s, *y, *qc, *rold;      // arena vectors (own parts)
    double *x_set, *p_set;
    long long stride;                            // doubles between consecutive shifts in x_set / p_set
    int n, L;
};

// r# = r, p[seed] (arena) = r, p[j] = r for every shift, (r,r)                                        :342-354
__global__ void __launch_bounds__(256) sh_vec_init(const __grid_constant__ ShVec a)
{
    __shared__ double scratch[32 * MAX_DOTS];
    double dot[1] = {0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
        const double r = a.r[i];
        a.rh[i] = r; a.p[i] = r;
        for (int j = 0; j < a.L; ++j) a.p_set[(size_t)j * a.stride + i] = r;
        dot[0] = fma(r, r, dot[0]);
    }
    block_sum<1>(dot, scratch);
    kernel_tail<1>(a.kc, dot, scratch);
}
// r_old = r; q = r - alpha s (in r); q_copy = q                                                       :374, 391-392
__global__ void __launch_bounds__(256) sh_vec_q(const __grid_constant__ ShVec a)
{
    if (a.sd->done) return;
    const double al = a.kc.sc->alpha;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
        const double r = a.r[i];
        const double q = fma(-al, a.s[i], r);
        a.rold[i] = r; a.r[i] = q; a.qc[i] = q;
    }
}
// x[seed] += alpha p + omega q; r = q - omega y; (r,r), (r#,r)                                         :411-416
__global__ void __launch_bounds__(256) sh_vec_xr(const __grid_constant__ ShVec a)
{
    if (a.sd->done) return;
    __shared__ double scratch[32 * MAX_DOTS];
    const double al = a.kc.sc->alpha, om = a.kc.sc->omega;
    double *x = a.x_set + (size_t)a.sd->seed * a.stride;
    double dot[2] = {0.0, 0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
        const double q = a.r[i];
        double xv = fma(al, a.p[i], x[i]);
        xv = fma(om, q, xv);
        const double r = fma(-om, a.y[i], q);
        x[i] = xv; a.r[i] = r;
        dot[0] = fma(r, r, dot[0]);
        dot[1] = fma(a.rh[i], r, dot[1]);
    }
    block_sum<2>(dot, scratch);
    kernel_tail<2>(a.kc, dot, scratch);
}
// all active shifts at once                                                                            :435-445
//   x_j += c1 q + alpha_j p_j ;  p_j += c2 q + c3 r_old ;  p_j = beta_j p_j + c4 r
__global__ void __launch_bounds__(256) sh_vec_shift(const __grid_constant__ ShVec a)
{
    const ShiftDev *sd = a.sd;
    if (!sd->live) return;
    extern __shared__ double s_coef[];                       // [n_active][SH_COEF] then the shift indices
    const int na = sd->n_active;
    int *s_idx = reinterpret_cast<int *>(s_coef + (size_t)na * SH_COEF);
    for (int t = threadIdx.x; t < na * SH_COEF; t += blockDim.x) s_coef[t] = sd->coef[t];
    for (int t = threadIdx.x; t < na; t += blockDim.x) s_idx[t] = sd->active[t];
    __syncthreads();
    for (int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x); i < a.n; i += 2 * gridDim.x * blockDim.x) {
        const bool two = i + 1 < a.n;                        // stride is even and every block starts 16-byte aligned
        double q0 = a.qc[i], q1 = two ? a.qc[i + 1] : 0.0;
        double o0 = a.rold[i], o1 = two ? a.rold[i + 1] : 0.0;
        double r0 = a.r[i], r1 = two ? a.r[i + 1] : 0.0;
#pragma unroll 2
        for (int t = 0; t < na; ++t) {
            const double *c = s_coef + (size_t)t * SH_COEF;
            double *xj = a.x_set + (size_t)s_idx[t] * a.stride + i, *pj = a.p_set + (size_t)s_idx[t] * a.stride + i;
            double x0, x1, p0, p1;
            if (two) {
                const double2 xv = *reinterpret_cast<const double2 *>(xj), pv = *reinterpret_cast<const double2 *>(pj);
                x0 = xv.x; x1 = xv.y; p0 = pv.x; p1 = pv.y;
            } else { x0 = xj[0]; p0 = pj[0]; x1 = p1 = 0.0; }
            x0 = fma(c[0], q0, x0); x0 = fma(c[1], p0, x0);
            x1 = fma(c[0], q1, x1); x1 = fma(c[1], p1, x1);
            p0 = fma(c[2], q0, p0); p0 = fma(c[3], o0, p0); p0 = c[4] * p0; p0 = fma(c[5], r0, p0);
            p1 = fma(c[2], q1, p1); p1 = fma(c[3], o1, p1); p1 = c[4] * p1; p1 = fma(c[5], r1, p1);
            if (two) {
                *reinterpret_cast<double2 *>(xj) = make_double2(x0, x1);
                *reinterpret_cast<double2 *>(pj) = make_double2(p0, p1);
            } else { xj[0] = x0; pj[0] = p0; }
        }
    }
}
// The next SpMV input: normally p[seed] = r + beta p[seed] - beta omega s (:421-423, same operation order as solver.c:117-119);
// on a seed switch the old seed's p is dead, r <- r / (zeta pi) (:499) and the new seed's p (just advanced by sh_vec_shift)
// takes its place in the arena.
__global__ void __launch_bounds__(256) sh_vec_p(const __grid_constant__ ShVec a)
{
    const ShiftDev *sd = a.sd;
    if (!sd->live) return;
    if (sd->switched) {
        const double sc = sd->r_scale;
        const double *pn = a.p_set + (size_t)sd->seed * a.stride;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
            a.r[i] = sc * a.r[i];
            a.p[i] = pn[i];
        }
    } else {
        const double be = a.kc.sc->beta, nbo = -a.kc.sc->beta * a.kc.sc->omega;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
            double t = be * a.p[i];
            t = fma(1.0, a.r[i], t);
            a.p[i] = fma(nbo, a.s[i], t);
        }
    }
}

inline TailDesc tail_none() { return TailDesc{TAIL_NONE, FIN_NONE, 0, 0, 0, 0, 0}; }
inline TailDesc tail_store(int ndot) { return TailDesc{TAIL_ALLREDUCE, FIN_STORE_PEND, ndot, 0, 0, 0, 0}; }

__global__ void sh_reset_scalars(Scalars *s)
{
    s->alpha = s->beta = s->omega = 0.0;
    for (int k = 0; k < MAX_DOTS; ++k) s->pend[k] = 0.0;
    s->k = 0; s->max_iter = 0; s->done = 0; s->converged = 0; s->error = 0; s->ticket = 0u;
}

struct ShRun {
    bicg_matrix *m;
    Context &c;
    ShiftDev *d_sd;
    ShVec base{};
    int launches = 0;
    explicit ShRun(bicg_matrix *mm) : m(mm), c(ctx()), d_sd(nullptr) {}

    PushDesc make_push(int id) const
    {
        PushDesc pd{};
        if (m->world == 1) return pd;
        pd.npeers = m->npush; pd.fence_writers = c.cfg.fence_writers;
        pd.src = m->vec(id);
        for (int s = 0; s < m->npush; ++s) {
            const int d = m->push_peer[s];
            pd.dst[s] = (double *)((char *)m->peer_base[d] + m->peer_vec_off[d]) + (long long)id * m->peer_vstride[d] + m->peer_ghost_off[d];
            pd.runs[s] = m->d_push_runs[s]; pd.nruns[s] = m->push_nruns[s];
        }
        return pd;
    }
    VecArgs vec_args(TailDesc tail) const
    {
        VecArgs a{};
        a.kc.sc = m->d_sc; a.kc.partials = m->d_partials; a.kc.hist = m->d_hist; a.kc.comm = m->comm; a.kc.tail = tail;
        a.v.x = m->vec(V_X); a.v.r = m->vec(V_R); a.v.rh = m->vec(V_RH); a.v.p = m->vec(V_P); a.v.s = m->vec(V_S);
        a.v.y = m->vec(V_Y); a.v.z = m->vec(V_Z); a.v.w = m->vec(V_W); a.v.v = m->vec(V_V); a.v.t = m->vec(V_T);
        a.v.b = m->vec(V_B); a.v.ax = m->vec(V_AX);
        a.n = m->n_loc; a.chunk = m->vchunk;
        return a;
    }
    void push(int id)                             // halo of arena vector `id` for the next SpMV (kernel-per-phase protocol)
    {
        if (m->world == 1) return;
        VecArgs a = vec_args(tail_none());
        a.kc.tail.signal_halo = 1;
        a.push = make_push(id);
        int rc = launch_vec(PH_PUSH, m->vgrid, a, c.stream);
        if (rc) fatal("bicgstab_b200: push kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
        ++launches; ++c.launches;
    }
    void spmv(int x_id, int y_id, int ndot, const double *a0, const double *b0, const double *a1 = nullptr, const double *b1 = nullptr)
    {
        SpmvArgs a = make_spmv_args(m, m->plan, x_id, y_id);
        a.kc.tail = tail_store(ndot);
        a.shift_sigma = &d_sd->sigma_seed;
        epi_add_dot(a.epi, a0, b0);
        if (ndot > 1) epi_add_dot(a.epi, a1, b1);
        launch_spmv_plan(m, m->plan, a, 0);
        ++launches;
    }
    ShVec vargs(TailDesc tail) const
    {
        ShVec v = base;
        v.kc.sc = m->d_sc; v.kc.partials = m->d_partials; v.kc.hist = m->d_hist; v.kc.comm = m->comm; v.kc.tail = tail;
        return v;
    }
    void iteration()
    {
        const int G = m->vgrid;
        const double *Y = nullptr;
        spmv(V_P, V_S, 1, m->vec(V_RH), Y);                                           // s = (A + sigma I) p, (r#,s)   :377-387
        sh_scalar_alpha<<<1, 1, 0, c.stream>>>(d_sd, m->d_sc);
        sh_vec_q<<<G, 256, 0, c.stream>>>(vargs(tail_none()));                        // q, r_old, q_copy            :374, 391-392
        push(V_R);
        spmv(V_R, V_Y, 2, m->vec(V_R), m->vec(V_R), m->vec(V_R), Y);                  // y = (A + sigma I) q, (q,q), (q,y)  :395-406
        sh_scalar_omega<<<1, 1, 0, c.stream>>>(d_sd, m->d_sc);
        sh_vec_xr<<<G, 256, 0, c.stream>>>(vargs(tail_store(2)));                     // x[seed], r, (r,r), (r#,r)   :411-416
        sh_scalar_iter<<<1, 512, 0, c.stream>>>(d_sd, m->d_sc);                       // beta ... loop test          :420, 429-537
        const size_t smem = (size_t)base.L * (SH_COEF * sizeof(double) + sizeof(int));
        sh_vec_shift<<<std::max(1, std::min(c.sm_count * 8, (m->n_loc + 511) / 512)), 256, smem, c.stream>>>(vargs(tail_none()));
        sh_vec_p<<<G, 256, 0, c.stream>>>(vargs(tail_none()));                        // p[seed] (or the switch)     :421-423 / :499
        push(V_P);
        launches += 7; c.launches += 7;
    }
};

} // namespace

int shifted_solve(bicg_matrix *m, double *x_set, double *r, const double *sigma, int L, int seed, double tol, int max_iter_opt)
{
    Context &c = ctx();
    c.ensure();
    if (L <= 0 || seed < 0 || seed >= L) return -1;
    const int n = m->n_loc;
    const int max_iter = max_iter_opt + 1;                                            // :293
    const long long stride = ((long long)n + 15) / 16 * 16;

    // ---- device state -------------------------------------------------------------------------------------------------
    ShiftDev h{};
    h.L = L; h.max_iter = max_iter; h.tol = tol; h.seed = seed;
    auto dalloc = [&](size_t bytes) { return c.dev_alloc(std::max<size_t>(bytes, 16)); };
    h.sigma = (double *)dalloc(L * sizeof(double));
    h.alpha_set = (double *)dalloc(L * sizeof(double)); h.beta_set = (double *)dalloc(L * sizeof(double));
    h.omega_set = (double *)dalloc(L * sizeof(double)); h.eta_set = (double *)dalloc(L * sizeof(double));
    h.zeta_set = (double *)dalloc(L * sizeof(double));
    h.stop_flag = (unsigned char *)dalloc(L); h.stop_iter = (int *)dalloc(L * sizeof(int));
    h.alpha_arch = (double *)dalloc(max_iter * sizeof(double)); h.beta_arch = (double *)dalloc(max_iter * sizeof(double));
    h.omega_arch = (double *)dalloc(max_iter * sizeof(double));
    h.pi_arch = (double *)dalloc((size_t)L * max_iter * sizeof(double));
    h.coef = (double *)dalloc((size_t)L * SH_COEF * sizeof(double)); h.active = (int *)dalloc(L * sizeof(int));
    h.hist = (double *)dalloc(((size_t)max_iter + 1) * sizeof(double));
    h.ev_k = (int *)dalloc(SH_EVENTS * sizeof(int)); h.ev_seed = (int *)dalloc(SH_EVENTS * sizeof(int));
    h.ev_remain = (int *)dalloc(SH_EVENTS * sizeof(int));
    h.ev_vals = (double *)dalloc((size_t)SH_EVENTS * L * 3 * sizeof(double));
    ShiftDev *d_sd = (ShiftDev *)dalloc(sizeof(ShiftDev));
    double *d_x = (double *)dalloc((size_t)L * stride * sizeof(double));
    double *d_p = (double *)dalloc((size_t)L * stride * sizeof(double));
    BICG_CUDA(cudaMemcpyAsync(d_sd, &h, sizeof(ShiftDev), cudaMemcpyHostToDevice, c.stream));
    BICG_CUDA(cudaMemcpyAsync(h.sigma, sigma, L * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    BICG_CUDA(cudaMemsetAsync(h.pi_arch, 0, (size_t)L * max_iter * sizeof(double), c.stream));
    BICG_CUDA(cudaMemsetAsync(h.hist, 0, ((size_t)max_iter + 1) * sizeof(double), c.stream));
    BICG_CUDA(cudaMemcpy2DAsync(d_x, stride * sizeof(double), x_set, (size_t)n * sizeof(double), (size_t)n * sizeof(double), L,
                                cudaMemcpyHostToDevice, c.stream));
    BICG_CUDA(cudaMemcpyAsync(m->vec(V_R), r, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    sh_reset_scalars<<<1, 1, 0, c.stream>>>(m->d_sc);

    ShRun run(m);
    run.d_sd = d_sd;
    run.base.sd = d_sd;
    run.base.r = m->vec(V_R); run.base.rh = m->vec(V_RH); run.base.p = m->vec(V_P); run.base.s = m->vec(V_S);
    run.base.y = m->vec(V_Y); run.base.qc = m->vec(V_W); run.base.rold = m->vec(V_V);
    run.base.x_set = d_x; run.base.p_set = d_p; run.base.stride = stride; run.base.n = n; run.base.L = L;
    const size_t smem = (size_t)L * (SH_COEF * sizeof(double) + sizeof(int));
    if (smem > 48 * 1024) BICG_CUDA(cudaFuncSetAttribute(sh_vec_shift, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    cudaEvent_t e0, e1;
    BICG_CUDA(cudaEventCreate(&e0)); BICG_CUDA(cudaEventCreate(&e1));
    const int launches0 = c.launches;
    BICG_CUDA(cudaEventRecord(e0, c.stream));                                         // the reference's timed region :364
    sh_vec_init<<<m->vgrid, 256, 0, c.stream>>>(run.vargs(tail_store(1)));            // :342-354
    sh_scalar_init<<<1, 256, 0, c.stream>>>(d_sd, m->d_sc);
    run.push(V_P);
    c.launches += 2;

    const int U = 8, DEPTH = 2, RING = 64;
    std::vector<cudaEvent_t> ring((size_t)RING, nullptr);
    const int batches = (max_iter + U - 1) / U;
    for (int b = 0; b < batches; ++b) {
        if (b >= DEPTH) {
            const int o = (b - DEPTH) % RING;
            BICG_CUDA(cudaEventSynchronize(ring[(size_t)o]));
            if (c.h_flags[o * 4 + 0]) break;                                          // done was raised in batch b - DEPTH
        }
        for (int u = 0; u < U; ++u) run.iteration();
        const int o = b % RING;
        BICG_CUDA(cudaMemcpyAsync(&c.h_flags[o * 4], &d_sd->done, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
        if (!ring[(size_t)o]) BICG_CUDA(cudaEventCreateWithFlags(&ring[(size_t)o], cudaEventDisableTiming));
        BICG_CUDA(cudaEventRecord(ring[(size_t)o], c.stream));
    }
    BICG_CUDA(cudaEventRecord(e1, c.stream));

    // ---- results ------------------------------------------------------------------------------------------------------
    BICG_CUDA(cudaMemcpy2DAsync(x_set, (size_t)n * sizeof(double), d_x, stride * sizeof(double), (size_t)n * sizeof(double), L,
                                cudaMemcpyDeviceToHost, c.stream));
    BICG_CUDA(cudaMemcpyAsync(r, m->vec(V_R), (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    ShiftDev out{};
    BICG_CUDA(cudaMemcpyAsync(&out, d_sd, sizeof(ShiftDev), cudaMemcpyDeviceToHost, c.stream));
    Scalars hs;
    BICG_CUDA(cudaMemcpyAsync(&hs, m->d_sc, sizeof(Scalars), cudaMemcpyDeviceToHost, c.stream));
    BICG_CUDA(cudaStreamSynchronize(c.stream));
    for (cudaEvent_t e : ring) if (e) cudaEventDestroy(e);
    if (hs.error) fatal("bicgstab_b200: rank %d timed out waiting for a peer GPU in the shifted solver", m->rank);
    float ms = 0.f;
    BICG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);

    const int k = out.k;
    c.last_hist.assign((size_t)std::max(k, 1), 0.0);
    BICG_CUDA(cudaMemcpy(c.last_hist.data(), out.hist, (size_t)std::max(k, 1) * sizeof(double), cudaMemcpyDeviceToHost));
    c.last_shift_stop.assign((size_t)L, 0);
    BICG_CUDA(cudaMemcpy(c.last_shift_stop.data(), out.stop_iter, (size_t)L * sizeof(int), cudaMemcpyDeviceToHost));
    c.last_shift_seed = out.seed;
    bicg_stats st{};
    st.iters = k - 1; st.converged = out.stop_count >= L; st.final_res = sqrt(out.dot_r / out.dot_zero); st.loop_ms = ms;
    st.kernel_launches = c.launches - launches0;
    st.h2d_bytes = (uint64_t)L * n * 8 + (uint64_t)n * 8; st.d2h_bytes = st.h2d_bytes;
    c.last_stats = st;

    if (c.rank == 0 && !c.cfg.quiet) {
        // what the reference prints: the seed switches (:518-526), then the MEASURE_TIME lines (:557-561)
        std::vector<int> ek(SH_EVENTS), es(SH_EVENTS), er(SH_EVENTS);
        std::vector<double> ev((size_t)SH_EVENTS * L * 3);
        BICG_CUDA(cudaMemcpy(ek.data(), out.ev_k, SH_EVENTS * sizeof(int), cudaMemcpyDeviceToHost));
        BICG_CUDA(cudaMemcpy(es.data(), out.ev_seed, SH_EVENTS * sizeof(int), cudaMemcpyDeviceToHost));
        BICG_CUDA(cudaMemcpy(er.data(), out.ev_remain, SH_EVENTS * sizeof(int), cudaMemcpyDeviceToHost));
        BICG_CUDA(cudaMemcpy(ev.data(), out.ev_vals, ev.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int e = 0; e < out.n_events && e < SH_EVENTS; ++e) {
            for (int j = 0; j < L; ++j) {
                const double *v = &ev[((size_t)e * L + j) * 3];
                if (!std::isnan(v[0])) printf("sigma[%d] eta: %f, pi: %f, zeta: %f\n", j, v[0], v[1], v[2]);
            }
            printf("k: %d, seed: %d, remain: %d\n", ek[(size_t)e], es[(size_t)e], er[(size_t)e]);
        }
        const double t = ms * 1e-3;
        printf("Total iter   : %d\n", k - 1);
        printf("Total time   : %e [sec.] \n", t);
        printf("Avg time/iter: %e [sec.] \n", t / k);
        fflush(stdout);
    }

    for (void *p : {(void *)h.sigma, (void *)h.alpha_set, (void *)h.beta_set, (void *)h.omega_set, (void *)h.eta_set, (void *)h.zeta_set,
                    (void *)h.stop_flag, (void *)h.stop_iter, (void *)h.alpha_arch, (void *)h.beta_arch, (void *)h.omega_arch,
                    (void *)h.pi_arch, (void *)h.coef, (void *)h.active, (void *)h.hist, (void *)h.ev_k, (void *)h.ev_seed,
                    (void *)h.ev_remain, (void *)h.ev_vals, (void *)d_sd, (void *)d_x, (void *)d_p})
        c.dev_free(p);
    return k;                                                                         // :600
}

} // namespace bicg
