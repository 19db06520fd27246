// mega.cu -- one persistent, cooperative kernel per solve: the whole iteration loop of solver.c on the device.
//
// Why: with the matrix split over 8 GPUs an iteration is ~13 us of memory traffic, and five kernel boundaries
// (launch latency + last-CTA election + reduction tail each) cost several times that.  Here one CTA per SM stays
// resident for the entire solve.  Every CTA owns a fixed, contiguous range of rows:
//   * SpMV phases run the warp-specialised TMA pipeline of spmv.cu over the CTA's own tiles; the producer warp
//     never stops -- while the consumers run a vector phase or sit in a barrier it is already streaming the first
//     stages of the NEXT SpMV (the matrix never changes), so the DRAM pipe stays primed across phases;
//   * vector phases touch only the CTA's own rows (same element-wise bodies as vec.cu), so nothing but the
//     gathered x crosses SMs;
//   * phases are separated by a grid barrier (atomic arrive + generation flag) whose last arriver -- the master
//     warp -- combines the per-CTA dot partials in a fixed order, performs the cross-GPU reduction over the peer
//     mailboxes, evaluates the scalar recurrence (dev.cuh: tail_warp, identical to the multi-kernel path), signals
//     and awaits the neighbours' halo epochs, and only then opens the barrier.  The loop test of solver.c:86 is
//     evaluated there too, so the kernel leaves the loop at exactly the reference's iteration.
// Coherence: x is gathered with plain (L1-cached) loads; every thread-0 that observes the barrier opening issues a
// gpu-scope fence (L1 invalidate) before its CTA continues, so rows rewritten by other SMs are re-fetched from L2.
//
// Used for bicgstab / ca_bicgstab / pipe_bicgstab when the SpMV plan is thread-per-row (LANES = 1); everything else
// (and pipe_bicgstab_rr) stays on the multi-kernel path of solve.cu.  BICG_MEGA=0 disables it.
#include "mega.cuh"
#include "vec_body.cuh"

namespace bicg {

namespace {

constexpr int PROW_PAD = 8;
struct StageHdr { int row0, row1; unsigned a0; int rowa; };

__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cbar(int nthreads)          // consumer-only CTA barrier (the producer warp free-runs)
{
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}
__device__ __forceinline__ double ld_coherent(const double *p)    // plain ld.global: L1-cached, never the .nc path
{
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// CTA-wide sum over the consumer threads only; result valid in every lane of warp 0
template <int N, int CT>
__device__ __forceinline__ void cblock_sum(double (&v)[N], double *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = CT / 32;
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = warp_sum(v[k]);
    cbar(CT);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) scratch[warp * N + k] = v[k];
    }
    cbar(CT);
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            double t = (lane < NW) ? scratch[lane * N + k] : 0.0;
            v[k] = warp_sum(t);
        }
    }
}

enum Epi : int { EPI_NONE = 0, EPI_RH_Y, EPI_QY_YY, EPI_CA4 };

template <int CT>
struct Mega {
    static constexpr int RPT = CT, PROW = RPT + PROW_PAD, NCW = CT / 32, UNR = 16;

    const MegaArgs &a;
    unsigned char *dyn;
    unsigned long long *full_bar, *empty_bar;
    StageHdr *hdr;
    double *scratch;
    volatile int *s_flags;        // [0] master, [1] stop, [2] consumed visits
    int tid, t0, my_tiles, row_lo, row_hi;
    unsigned vis;                 // SpMV tile visits consumed so far (mirrors the producer's counter)
    unsigned my_gen;              // grid-barrier generation (same value in every CTA)
    size_t stage_bytes;
    bool failed;

    __device__ Mega(const MegaArgs &args) : a(args) {}

    // ---------------------------------------------------------------- grid barrier + master work ----------
    template <int NDOT>
    __device__ void barrier(double (&dot)[NDOT > 0 ? NDOT : 1], TailDesc td, bool pushed = false)
    {
        if (NDOT > 0) cblock_sum<(NDOT > 0 ? NDOT : 1), CT>(dot, scratch);
        cbar(CT);                                     // every consumer's stores happen-before thread 0's fence
        const unsigned gen = my_gen;                  // generation this barrier closes; every CTA counts them locally
        if (tid == 0) {
#pragma unroll
            for (int k = 0; k < NDOT; ++k) __stcg(&a.partials[(size_t)blockIdx.x * MAX_DOTS + k], dot[k]);
            // release: the CTA barrier above + this fence order every store of the CTA (at system scope when this
            // CTA's rows went to peers) before the arrival
            if (pushed) __threadfence_system(); else __threadfence();
            const unsigned prev = atomicAdd(&a.bar->count, 1u);
            s_flags[0] = (prev == gridDim.x - 1);
        }
        cbar(CT);
        if (s_flags[0]) {
            // last arriver = master: everyone else of the grid is parked, so its extra work delays nobody twice
            if (tid < 32) {
                __threadfence();
                double tot[NDOT > 0 ? NDOT : 1];
#pragma unroll
                for (int k = 0; k < (NDOT > 0 ? NDOT : 1); ++k) tot[k] = 0.0;
                if (NDOT > 0) {
                    for (unsigned b = tid; b < gridDim.x; b += 32) {
#pragma unroll
                        for (int k = 0; k < NDOT; ++k) tot[k] += __ldcg(&a.partials[(size_t)b * MAX_DOTS + k]);
                    }
#pragma unroll
                    for (int k = 0; k < NDOT; ++k) tot[k] = warp_sum(tot[k]);
                }
                KernelCommon kc;
                kc.sc = a.sc; kc.partials = a.partials; kc.hist = a.hist; kc.comm = a.comm; kc.tail = td;
                tail_warp<NDOT>(kc, tot, true);
                if (tid == 0) {
                    a.bar->count = 0u;
                    st_release_gpu(&a.bar->gen, gen + 1u);      // release: orders the scalars and the counter reset
                }
            }
        } else if (tid == 0) {
            const unsigned long long t_start = globaltimer_ns();
            unsigned spins = 0;
            while (ld_acquire_gpu(&a.bar->gen) == gen) {        // acquire: later loads of this SM see the released data
                __nanosleep(32);
                if ((++spins & 1023u) == 0 && globaltimer_ns() - t_start > 2 * PEER_TIMEOUT_NS) { a.sc->error = 1; break; }
            }
        }
        my_gen = gen + 1u;
        cbar(CT);
    }

    // ---------------------------------------------------------------- SpMV over this CTA's tiles ----------
    template <int EPI>
    __device__ void spmv(const double *x, double *y, double (&dot)[4])
    {
        const int stages = a.stages, cap = a.cap;
        for (int lt = 0; lt < my_tiles; ++lt, ++vis) {
            const int s = (int)(vis % (unsigned)stages);
            mbar_wait(smem_u32(&full_bar[s]), (vis / (unsigned)stages) & 1u);
            const unsigned char *st = dyn + (size_t)s * stage_bytes;
            const double   *sval = reinterpret_cast<const double *>(st);
            const unsigned *scol = reinterpret_cast<const unsigned *>(sval + cap);
            const unsigned *sptr = scol + cap;
            const StageHdr h = hdr[s];
            const int row = h.row0 + tid;
            const bool valid = row < h.row1;
            int j = 0, e = 0;
            double e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;     // epilogue operands: in flight during the gathers
            if (valid) {
                j = (int)(sptr[row - h.rowa] - h.a0);
                e = (int)(sptr[row - h.rowa + 1] - h.a0);
                if (EPI == EPI_RH_Y) e0 = a.v.rh[row];
                if (EPI == EPI_QY_YY) e0 = a.v.r[row];
                if (EPI == EPI_CA4) { e0 = a.v.rh[row]; e1 = a.v.r[row]; e2 = a.v.s[row]; e3 = a.v.z[row]; }
            }
            double acc = 0.0;
            while (j < e) {
                unsigned c[UNR];
                double v[UNR], xv[UNR];
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int idx = min(j + u, e - 1);
                    c[u] = scol[idx];
                    v[u] = sval[idx];
                }
#pragma unroll
                for (int u = 0; u < UNR; ++u) xv[u] = ld_coherent(x + c[u]);
#pragma unroll
                for (int u = 0; u < UNR; ++u)
                    if (j + u < e) acc = fma(v[u], xv[u], acc);
                j += UNR;
            }
            if (valid) {
                y[row] = acc;
                if (EPI == EPI_RH_Y) dot[0] = fma(e0, acc, dot[0]);
                if (EPI == EPI_QY_YY) { dot[0] = fma(e0, acc, dot[0]); dot[1] = fma(acc, acc, dot[1]); }
                if (EPI == EPI_CA4) {
                    dot[0] = fma(e0, e1, dot[0]); dot[1] = fma(e0, acc, dot[1]);
                    dot[2] = fma(e0, e2, dot[2]); dot[3] = fma(e0, e3, dot[3]);
                }
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(smem_u32(&empty_bar[s]));
        }
    }

    // ---------------------------------------------------------------- vector phase over the CTA's own rows --
    template <int PH>
    __device__ void vec(double *dot)
    {
        Coef c;
        c.al = __ldcg(&a.sc->alpha); c.be = __ldcg(&a.sc->beta); c.om = __ldcg(&a.sc->omega);
        c.nbo = -c.be * c.om;
        int i = row_lo + tid;
        for (; i + 3 * CT < row_hi; i += 4 * CT) body<PH, Strided<4, CT>>(a.v, i, c, dot);
        for (; i + CT < row_hi; i += 2 * CT) body<PH, Strided<2, CT>>(a.v, i, c, dot);
        for (; i < row_hi; i += CT) body<PH, Strided<1, CT>>(a.v, i, c, dot);
    }
    __device__ bool push(const PushDesc &pd)          // true: this CTA stored to a peer
    {
        if (pd.npeers == 0 || !push_touches(pd, row_lo, row_hi)) return false;
        cbar(CT);                                     // the rows being pushed are final
        push_chunk(pd, row_lo, row_hi, tid, CT);
        return true;
    }
    int trace_it = 0;
    __device__ void mark(int slot)
    {
        if (a.trace && blockIdx.x == 0 && tid == 0 && trace_it < MEGA_TRACE_ITERS)
            a.trace[trace_it * MEGA_TRACE_SLOTS + slot] = globaltimer_ns();
    }
    __device__ bool stop_now() { return __ldcg(&a.sc->done) != 0 || __ldcg(&a.sc->error) != 0; }

    static __device__ TailDesc td_red(int fin, int ndot, int npend = 0) { return TailDesc{TAIL_ALLREDUCE, fin, ndot, npend, 0, 0, 0}; }
    static __device__ TailDesc td_post(int ndot) { return TailDesc{TAIL_POST, FIN_NONE, ndot, 0, 0, 0, 0}; }
    static __device__ TailDesc td_complete(int fin, int nred) { return TailDesc{TAIL_COMPLETE, fin, 0, 0, 0, nred, 0}; }
    static __device__ TailDesc td_pend(int ndot) { return TailDesc{TAIL_PEND, FIN_NONE, ndot, 0, 0, 0, 0}; }
    __device__ TailDesc with_halo(TailDesc t) const { t.signal_halo = a.comm.world > 1 ? 1 : 0; return t; }
    static __device__ TailDesc td_none() { return TailDesc{TAIL_NONE, FIN_NONE, 0, 0, 0, 0, 0}; }

    // ---------------------------------------------------------------- solver.c:86-127 -----------------------
    __device__ void run_bicgstab()
    {
        double d4[4], d2[2], d0[1];
        if (stop_now()) return;                                             // solver.c:86 before the first pass
        while (true) {
            mark(0);
            d4[0] = d4[1] = d4[2] = d4[3] = 0.0;
            spmv<EPI_RH_Y>(a.v.p, a.v.s, d4);                               // s = A p, (r#,s)           :88-91
            mark(1);
            { double t1[1] = {d4[0]}; barrier<1>(t1, td_red(FIN_BICG_ALPHA, 1)); }
            mark(2);
            if (stop_now()) break;
            vec<PH_BICG_Q>(d0);                                             // q = r - alpha s            :94
            const bool pr = push(a.push_r);
            mark(3);
            barrier<0>(d0, with_halo(td_none()), pr);
            mark(4);
            d4[0] = d4[1] = 0.0;
            spmv<EPI_QY_YY>(a.v.r, a.v.y, d4);                              // y = A q, (q,y), (y,y)      :96-102
            mark(5);
            d2[0] = d4[0]; d2[1] = d4[1];
            barrier<2>(d2, td_red(FIN_BICG_OMEGA, 2));
            mark(6);
            d2[0] = d2[1] = 0.0;
            vec<PH_BICG_XR>(d2);                                            // x, r, (r,r), (r#,r)        :105-114
            mark(7);
            barrier<2>(d2, td_red(FIN_BICG_BETA, 2));                       // beta, k++, loop test       :116-120
            mark(8);
            if (stop_now()) break;
            vec<PH_BICG_P>(d0);                                             // p                          :117-119
            const bool pp = push(a.push_p);
            mark(9);
            barrier<0>(d0, with_halo(td_none()), pp);
            mark(10);
            ++trace_it;
        }
    }
    // EXPERIMENTAL (BICG_MEGA_FUSEQ=1, off by default; parity-green on one GPU, not yet timed, not yet run multi-GPU): y = A q with
    // q[col] = r[col] - alpha s[col] gathered on the fly (solver.c:94 folded into :96), q[row] written to the spare
    // vector ax, dots (q,y), (y,y).  Removes the q vector phase and one grid barrier per iteration.
    __device__ void spmv_fq(double alpha, double (&dot)[4])
    {
        constexpr int UQ = 8;
        const int stages = a.stages, cap = a.cap;
        const double *rv = a.v.r, *sv = a.v.s;
        for (int lt = 0; lt < my_tiles; ++lt, ++vis) {
            const int s = (int)(vis % (unsigned)stages);
            mbar_wait(smem_u32(&full_bar[s]), (vis / (unsigned)stages) & 1u);
            const unsigned char *st = dyn + (size_t)s * stage_bytes;
            const double   *sval = reinterpret_cast<const double *>(st);
            const unsigned *scol = reinterpret_cast<const unsigned *>(sval + cap);
            const unsigned *sptr = scol + cap;
            const StageHdr h = hdr[s];
            const int row = h.row0 + tid;
            const bool valid = row < h.row1;
            int j = 0, e = 0;
            double r_row = 0.0, s_row = 0.0;
            if (valid) {
                j = (int)(sptr[row - h.rowa] - h.a0);
                e = (int)(sptr[row - h.rowa + 1] - h.a0);
                r_row = rv[row]; s_row = sv[row];
            }
            double acc = 0.0;
            while (j < e) {
                unsigned c[UQ];
                double v[UQ], xr[UQ], xs[UQ];
#pragma unroll
                for (int u = 0; u < UQ; ++u) {
                    const int idx = min(j + u, e - 1);
                    c[u] = scol[idx];
                    v[u] = sval[idx];
                }
#pragma unroll
                for (int u = 0; u < UQ; ++u) { xr[u] = ld_coherent(rv + c[u]); xs[u] = ld_coherent(sv + c[u]); }
#pragma unroll
                for (int u = 0; u < UQ; ++u)
                    if (j + u < e) acc = fma(v[u], fma(-alpha, xs[u], xr[u]), acc);
                j += UQ;
            }
            if (valid) {
                const double q_row = fma(-alpha, s_row, r_row);
                a.v.y[row] = acc;
                a.v.ax[row] = q_row;
                dot[0] = fma(q_row, acc, dot[0]);
                dot[1] = fma(acc, acc, dot[1]);
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(smem_u32(&empty_bar[s]));
        }
    }
    __device__ void run_bicgstab_fq()
    {
        double d4[4], d2[2], d0[1];
        if (stop_now()) return;
        while (true) {
            d4[0] = d4[1] = d4[2] = d4[3] = 0.0;
            spmv<EPI_RH_Y>(a.v.p, a.v.s, d4);                               // s = A p, (r#,s)
            const bool ps = push(a.push_s);                                 // neighbours gather s in the next SpMV
            { double t1[1] = {d4[0]}; barrier<1>(t1, with_halo(td_red(FIN_BICG_ALPHA, 1)), ps); }
            if (stop_now()) break;
            d4[0] = d4[1] = 0.0;
            spmv_fq(__ldcg(&a.sc->alpha), d4);                              // y = A (r - alpha s), q -> ax
            d2[0] = d4[0]; d2[1] = d4[1];
            barrier<2>(d2, td_red(FIN_BICG_OMEGA, 2));
            d2[0] = d2[1] = 0.0;
            vec<PH_BICG_XR_Q>(d2);                                          // x, r = q - omega y, (r,r), (r#,r)
            const bool pr = push(a.push_r);
            barrier<2>(d2, with_halo(td_red(FIN_BICG_BETA, 2)), pr);
            if (stop_now()) break;
            vec<PH_BICG_P>(d0);                                             // p
            const bool pp = push(a.push_p);
            barrier<0>(d0, with_halo(td_none()), pp);
        }
    }
    // ---------------------------------------------------------------- solver.c:216-259 ----------------------
    __device__ void run_ca()
    {
        double d4[4], d2[2], d1[1], d0[1];
        if (stop_now()) return;
        while (true) {
            vec<PH_CA_PS>(d0);                                              // p, s                       :217-222
            const bool ps = push(a.push_s);
            barrier<0>(d0, with_halo(td_none()), ps);
            d4[0] = 0.0;
            spmv<EPI_NONE>(a.v.s, a.v.z, d4);                               // z = A s                    :224
            cbar(CT);                                                       // own rows of z written by other warps
            d2[0] = d2[1] = 0.0;
            vec<PH_QY>(d2);                                                 // q, y, (q,y), (y,y)         :225-230
            barrier<2>(d2, td_red(FIN_OMEGA2, 2));
            d1[0] = 0.0;
            vec<PH_CA_XR>(d1);                                              // x, r, local (r,r)          :233-236
            const bool pr = push(a.push_r);
            barrier<1>(d1, with_halo(td_pend(1)), pr);
            d4[0] = d4[1] = d4[2] = d4[3] = 0.0;
            spmv<EPI_CA4>(a.v.r, a.v.w, d4);                                // w = A r, 4 dots            :238-247
            barrier<4>(d4, td_red(FIN_CAPIPE_END, 4, 1));                   // beta, alpha, k++, test     :248-253
            if (stop_now()) break;
        }
    }
    // ---------------------------------------------------------------- solver.c:351-398 ----------------------
    __device__ void run_pipe()
    {
        double d5[5], d4[4], d2[2], d0[1];
        if (stop_now()) return;
        while (true) {
            d2[0] = d2[1] = 0.0;
            vec<PH_PIPE_1>(d2);                                             // p,s,z,q,y + (q,y),(y,y)    :352-364
            const bool pz = push(a.push_z);
            barrier<2>(d2, with_halo(td_post(2)), pz);                          // MPI_Iallreduce x2
            d4[0] = 0.0;
            spmv<EPI_NONE>(a.v.z, a.v.v, d4);                               // v = A z                    :365
            barrier<0>(d0, td_complete(FIN_OMEGA2, 2));                     // MPI_Wait x2 -> omega       :366-369
            d5[0] = d5[1] = d5[2] = d5[3] = d5[4] = 0.0;
            vec<PH_PIPE_3>(d5);                                             // x, r, w + 5 dots           :370-380
            const bool pw = push(a.push_w);
            barrier<5>(d5, with_halo(td_post(5)), pw);
            spmv<EPI_NONE>(a.v.w, a.v.t, d4);                               // t = A w                    :381
            barrier<0>(d0, td_complete(FIN_CAPIPE_END, 5));                 // MPI_Wait x5 -> beta, alpha :382-388
            if (stop_now()) break;
        }
    }
};

template <int CT, bool FQ>
__global__ void __launch_bounds__(CT + 32, 1) bicg_mega_kernel(const __grid_constant__ MegaArgs a)
{
    using M = Mega<CT>;
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ __align__(8) unsigned long long full_bar[4], empty_bar[4];
    __shared__ StageHdr hdr[4];
    __shared__ double scratch[32 * MAX_DOTS];
    __shared__ int s_flags[4];

    const int tid = threadIdx.x;
    const int stages = a.stages, cap = a.cap;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1u);
            mbar_init(smem_u32(&empty_bar[s]), (unsigned)M::NCW);
        }
        s_flags[0] = s_flags[1] = s_flags[2] = 0;
        mbar_fence_init();
    }
    __syncthreads();

    const int t0 = a.cta_tile[blockIdx.x], t1 = a.cta_tile[blockIdx.x + 1];
    const int my_tiles = t1 - t0;
    const size_t stage_bytes = (size_t)cap * 12 + (size_t)M::PROW * 4;

    if (tid >= CT) {
        // ============================ producer warp: streams this CTA's tiles round and round ==============
        if (tid == CT && my_tiles > 0) {
            volatile int *flags = s_flags;
            unsigned v = 0;
            bool stop = false;
            for (;; ++v) {
                const int s = (int)(v % (unsigned)stages);
                if (v >= (unsigned)stages) {
                    const unsigned par = (v / (unsigned)stages - 1u) & 1u;
                    while (!mbar_try_wait(smem_u32(&empty_bar[s]), par)) {
                        if (flags[1]) { stop = true; break; }
                    }
                }
                if (stop || flags[1]) break;
                const int t = t0 + (int)(v % (unsigned)my_tiles);
                const int row0 = a.tile_row[t], row1 = a.tile_row[t + 1];
                const unsigned p0 = a.tile_nz[t], p1 = a.tile_nz[t + 1];
                const unsigned a0 = p0 & ~3u, cnt = ((p1 + 3u) & ~3u) - a0;
                const int rowa = row0 & ~3, cntp = ((row1 + 1 + 3) & ~3) - rowa;
                unsigned char *st = dyn_smem + (size_t)s * stage_bytes;
                double   *sval = reinterpret_cast<double *>(st);
                unsigned *scol = reinterpret_cast<unsigned *>(sval + cap);
                unsigned *sptr = scol + cap;
                hdr[s] = StageHdr{row0, row1, a0, rowa};
                const unsigned bar = smem_u32(&full_bar[s]);
                mbar_arrive_expect_tx(bar, cnt * 12u + (unsigned)cntp * 4u);
                if (cnt) {
                    tma_load_1d(smem_u32(sval), a.val + a0, cnt * 8u, bar);
                    tma_load_1d(smem_u32(scol), a.col + a0, cnt * 4u, bar);
                }
                tma_load_1d(smem_u32(sptr), a.ptr + rowa, (unsigned)cntp * 4u, bar);
            }
            // drain: bulk copies already issued must land before the CTA may retire its shared memory
            const unsigned consumed = (unsigned)flags[2];
            for (unsigned w = consumed; w < v; ++w)
                mbar_wait(smem_u32(&full_bar[w % (unsigned)stages]), (w / (unsigned)stages) & 1u);
        }
    } else {
        // ============================ consumer warps: the solver ============================================
        M m(a);
        m.dyn = dyn_smem; m.full_bar = full_bar; m.empty_bar = empty_bar; m.hdr = hdr; m.scratch = scratch;
        m.s_flags = s_flags; m.tid = tid; m.t0 = t0; m.my_tiles = my_tiles; m.vis = 0u; m.stage_bytes = stage_bytes;
        m.row_lo = a.tile_row[t0]; m.row_hi = a.tile_row[t1]; m.failed = false;
        m.my_gen = ld_acquire_gpu(&a.bar->gen);       // left by the previous solve; nobody can have advanced it yet
        if constexpr (FQ) m.run_bicgstab_fq();
        else if (a.method == 0) m.run_bicgstab();
        else if (a.method == 1) m.run_ca();
        else m.run_pipe();
        cbar(CT);
        if (tid == 0) { s_flags[2] = (int)m.vis; __threadfence_block(); s_flags[1] = 1; }
    }
}

template <int CT, bool FQ>
cudaError_t launch(const MegaArgs &a, int grid, size_t smem, cudaStream_t st)
{
    void *params[1] = {(void *)&a};
    return cudaLaunchCooperativeKernel((const void *)bicg_mega_kernel<CT, FQ>, dim3(grid), dim3(CT + 32), params, smem, st);
}
template <int CT, bool FQ>
cudaError_t set_attr()
{
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, bicg_mega_kernel<CT, FQ>);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(bicg_mega_kernel<CT, FQ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                227 * 1024 - (int)fa.sharedSizeBytes);
}

} // namespace

size_t mega_smem_bytes(int cap, int stages, int threads) { return (size_t)stages * ((size_t)cap * 12u + (size_t)(threads + PROW_PAD) * 4u); }

int mega_setup_attributes()
{
    cudaError_t e;
    if ((e = set_attr<256, false>()) != cudaSuccess) return (int)e;
    if ((e = set_attr<512, false>()) != cudaSuccess) return (int)e;
    if ((e = set_attr<256, true>()) != cudaSuccess) return (int)e;
    if ((e = set_attr<512, true>()) != cudaSuccess) return (int)e;
    return 0;
}

int launch_mega(int threads, bool fuse_q, int grid, size_t smem, const MegaArgs &a, cudaStream_t st)
{
    switch (threads) {
    case 256: return (int)(fuse_q ? launch<256, true>(a, grid, smem, st) : launch<256, false>(a, grid, smem, st));
    case 512: return (int)(fuse_q ? launch<512, true>(a, grid, smem, st) : launch<512, false>(a, grid, smem, st));
    default:  return (int)cudaErrorInvalidValue;
    }
}

} // namespace bicg
