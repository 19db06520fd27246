// dev.cuh -- device-side building blocks shared by the SpMV and the fused-vector kernels (sm_100a).
//
//   * PTX wrappers: mbarrier + 1-D TMA bulk copies (cp.async.bulk -> SASS UBLKCP), system-scope
//     release/acquire accesses for the NVLink peer mailboxes;
//   * deterministic block / grid reductions (fixed order -> bitwise reproducible dot products);
//   * the kernel "tail": what the last CTA of a kernel does once the grid's partial dots are combined --
//     the cross-GPU all-reduce over peer memory (post / wait split-phase, summed in rank order so every
//     rank gets bitwise identical scalars), the scalar recurrences of solver.c (alpha, beta, omega, the
//     loop test), and the halo-ready signal to the peers.  This is the device-side replacement of the
//     reference's MPI_Iallreduce / MPI_Wait pairs and host-side scalar code (solver.c:89-126, 227-258,
//     363-397).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bicg {

constexpr int MAX_RANKS  = 8;     // one NVSwitch box
constexpr int MAX_DOTS   = 8;     // dot products one kernel can reduce
constexpr int MAIL_VALS  = 8;

// ------------------------------------------------------------------------------------------------
// device-resident solver state (one per matrix arena)
// ------------------------------------------------------------------------------------------------
struct Scalars {
    // reference scalars, same names as solver.c:55-56, 182-183
    double rTr, rTr_old, rTs, rTy, yTy, rTw, wTw, rTz, dot_r, dot_zero, alpha, beta, omega;
    double tol2;                 // tol * tol                                (solver.c:86)
    double pend[MAX_DOTS];       // locally reduced values waiting for a later cross-GPU reduction
    int    k, max_iter;          // iteration counter / MAX_ITER             (solver.c:4, 120)
    int    done;                 // loop test failed -> every later kernel of the batch returns at once
    int    converged;
    int    error;                // 1: peer wait timed out
    unsigned int ticket;         // last-CTA election
    unsigned int red_epoch;      // sequence number of cross-GPU reductions posted by this rank
    unsigned int red_done;       // ... completed by this rank
    unsigned int halo_epoch;     // sequence number of halo pushes issued by this rank
    unsigned int pad_;
};

// one mailbox = what rank `src` contributes to one reduction: MAIL_VALS doubles, each as TWO self-validating 8-byte
// words {flag32 | data32} (NCCL's LL protocol): an aligned 8-byte store is single-copy atomic, so a word whose flag
// equals the expected epoch carries valid data -- no fence between "flag" and data, and nothing depends on a 16-byte
// vector store arriving un-torn over NVLink.  128 B so that no two mailboxes share a line.
struct alignas(128) Mailbox { unsigned long long w[2 * MAIL_VALS]; };
struct alignas(128) HaloFlag { unsigned long long epoch; unsigned long long pad_[15]; };

// peer-memory view of the job, passed by value to every kernel
struct CommDev {
    int rank, world;
    Mailbox  *mail[MAX_RANKS];   // mail[p] = base of rank p's mailbox array [2 parities][MAX_RANKS sources]
    HaloFlag *hflag[MAX_RANKS];  // hflag[p] = base of rank p's halo flags [MAX_RANKS sources]
    unsigned send_mask;          // peers this rank pushes halo data to
    unsigned recv_mask;          // peers this rank receives halo data from
    unsigned long long timeout_ns;   // bound of every device-side wait for a peer / another CTA (BICG_PEER_TIMEOUT_S)
};

// finalize ids: which scalar recurrence the tail evaluates once the reduced values are known
enum Fin : int {
    FIN_NONE = 0,
    FIN_BICG_INIT,     // tot0=(r,r)                                   solver.c:78-83
    FIN_BICG_ALPHA,    // tot0=(r#,s)            -> alpha              solver.c:89-93
    FIN_BICG_OMEGA,    // tot0=(q,y) tot1=(y,y)  -> omega              solver.c:97-104
    FIN_BICG_BETA,     // tot0=(r,r) tot1=(r#,r) -> beta, k++, test    solver.c:108-120
    FIN_STORE_RTR,     // tot0=(r,r) -> rTr                            solver.c:203
    FIN_CAPIPE_INIT,   // tot0=(r,w) -> alpha, beta=0, omega=0, test   solver.c:206-213
    FIN_OMEGA2,        // tot0=(q,y) tot1=(y,y) -> omega               solver.c:227-232 / 363-369
    FIN_CAPIPE_END,    // tot0..3=(r#,r),(r#,w),(r#,s),(r#,z) tot4=(r,r) -> beta, alpha, k++, test   solver.c:240-253
    FIN_STORE_PEND,    // reduced values -> Scalars::pend[] (the shifted solver's own scalar kernels take it from there)
};

// what the tail does with the locally reduced dots
enum TailOp : int {
    TAIL_NONE = 0,
    TAIL_PEND,         // keep them in Scalars::pend[pend_off ...] for a later kernel (no communication)
    TAIL_ALLREDUCE,    // [pend values +] local dots -> post + wait -> finalize          (blocking sync point)
    TAIL_POST,         // [pend values +] local dots -> post only                        (MPI_Iallreduce)
    TAIL_COMPLETE,     // wait for the reduction posted earlier -> finalize              (MPI_Wait)
};

struct TailDesc {
    int op;            // TailOp
    int fin;           // Fin
    int ndot;          // local dots produced by this kernel
    int npend;         // values taken from Scalars::pend and appended after the local dots (ALLREDUCE/POST)
    int pend_off;      // TAIL_PEND: where to store
    int nred;          // TAIL_COMPLETE: number of values of the pending reduction
    int signal_halo;   // 1: this kernel pushed halo data; tell the receivers
};

struct KernelCommon {
    Scalars *sc;
    double  *partials;     // [gridDim.x][MAX_DOTS]
    double  *hist;         // hist[k] = dot_r / dot_zero
    CommDev  comm;
    TailDesc tail;
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    while (!mbar_try_wait(bar, parity)) { }
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier.  16-byte aligned
// addresses and a multiple-of-16 size are required.
__device__ __forceinline__ void tma_load_1d(unsigned dst_smem, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_flag_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// ---- LL words: {flag32 | data32}; a double travels as two of them ------------------------------------------
__device__ __forceinline__ unsigned long long ll_pack(unsigned data, unsigned flag)
{
    return ((unsigned long long)flag << 32) | (unsigned long long)data;
}
__device__ __forceinline__ void ll_encode(double v, unsigned flag, unsigned long long &w0, unsigned long long &w1)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    w0 = ll_pack((unsigned)b, flag);
    w1 = ll_pack((unsigned)(b >> 32), flag);
}
__device__ __forceinline__ bool ll_valid(unsigned long long w0, unsigned long long w1, unsigned flag)
{
    return (unsigned)(w0 >> 32) == flag && (unsigned)(w1 >> 32) == flag;
}
__device__ __forceinline__ double ll_decode(unsigned long long w0, unsigned long long w1)
{
    return __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
}
// both words of one value with one 16-byte access; each word validates itself, so tearing is harmless
__device__ __forceinline__ void st_ll_sys(unsigned long long *p, unsigned long long w0, unsigned long long w1)
{
    asm volatile("st.relaxed.sys.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ void ld_ll_sys(const unsigned long long *p, unsigned long long &w0, unsigned long long &w1)
{
    asm volatile("ld.relaxed.sys.global.v2.b64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_ll_gpu(unsigned long long *p, unsigned long long w)
{
    asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long ld_ll_gpu(const unsigned long long *p)
{
    unsigned long long w;
    asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return w;
}
__device__ __forceinline__ void ld_ll_gpu2(const unsigned long long *p, unsigned long long &w0, unsigned long long &w1)
{
    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
// plain ld.global: L1-cached but never the non-coherent (.nc) path -- for vectors that other SMs / peer GPUs rewrite
// while the kernel is running (visibility comes from the acquire fence that follows the flag wait)
// L2-only gather (ld.global.cg): never allocates in L1, so it cannot return a line that went stale in this SM's L1
__device__ __forceinline__ double ld_l2(const double *p)
{
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_coherent(const double *p)
{
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double *p)
{
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr unsigned long long PEER_TIMEOUT_NS = 20000000000ull;  // default bound (20 s): a lost peer must not hang the GPU, but
                                                                // ordinary host-side skew between ranks must not kill the job

// ------------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum N per-thread values over the CTA.  Result is valid in every lane of warp 0.  `scratch` holds
// 32 * N doubles.  Fixed combination order -> deterministic.
template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], double *scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();                     // scratch may still be in use by a previous call
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) scratch[warp * N + k] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            double t = (lane < nwarps) ? scratch[lane * N + k] : 0.0;
            v[k] = warp_sum(t);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// cross-GPU reduction over peer mailboxes (called by warp 0 of the last CTA, all 32 lanes)
// ------------------------------------------------------------------------------------------------
// post: lane p stores this rank's `nv` values (at least one, so a pure barrier works too) into rank p's
// mailbox[parity][me] as LL words carrying the epoch.
__device__ __forceinline__ void xg_post(const CommDev &c, unsigned epoch, const double *vals, int nv)
{
    const int lane = threadIdx.x & 31;
    if (lane < c.world) {
        Mailbox *mb = c.mail[lane] + (epoch & 1u) * MAX_RANKS + c.rank;
        const int n = nv > 0 ? nv : 1;
        for (int k = 0; k < n; ++k) {
            unsigned long long w0, w1;
            ll_encode(nv > 0 ? vals[k] : 0.0, epoch, w0, w1);
            st_ll_sys(&mb->w[2 * k], w0, w1);
        }
    }
    __syncwarp();
}
// wait: lane p polls the words rank p sent to this rank until they carry `epoch`, keeps rank p's values; lane k
// then adds value k over ranks 0..world-1 in rank order (the same order on every rank -> bitwise identical
// results everywhere).  Returns false on timeout.
__device__ __forceinline__ bool xg_wait_sum(const CommDev &c, unsigned epoch, double *vals, int nv)
{
    __shared__ double s_contrib[MAX_RANKS][MAIL_VALS];
    const int lane = threadIdx.x & 31;
    const Mailbox *mine = c.mail[c.rank] + (epoch & 1u) * MAX_RANKS;
    bool ok = true;
    if (lane < c.world) {
        const unsigned long long t0 = globaltimer_ns();
        const int n = nv > 0 ? nv : 1;
        for (int k = 0; k < n && ok; ++k) {
            unsigned long long w0, w1;
            for (;;) {
                ld_ll_sys(&mine[lane].w[2 * k], w0, w1);
                if (ll_valid(w0, w1, epoch)) break;
                if (globaltimer_ns() - t0 > c.timeout_ns) { ok = false; break; }
            }
            s_contrib[lane][k] = ll_decode(w0, w1);
        }
    }
    ok = __all_sync(0xffffffffu, ok);
    __syncwarp();
    if (ok && lane < nv) {
        double acc = s_contrib[0][lane];
        for (int p = 1; p < c.world; ++p) acc += s_contrib[p][lane];
        vals[lane] = acc;
    }
    __syncwarp();
    return ok;
}

// ------------------------------------------------------------------------------------------------
// scalar recurrences -- one thread, same operation order as the reference
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void loop_test(Scalars *s)
{
    // solver.c:86  while (dot_r > tol * tol * dot_zero && k < max_iter)
    const bool go = (s->dot_r > s->tol2 * s->dot_zero) && (s->k < s->max_iter);
    if (!go) {
        s->done = 1;
        s->converged = !(s->dot_r > s->tol2 * s->dot_zero);
    }
}

// hist == nullptr: the caller keeps a private copy of the scalars and somebody else records the history
__device__ __forceinline__ void finalize(int fin, Scalars *s, double *hist, const double *t)
{
    double hdummy;
#define BICG_HIST(i) (*(hist ? &hist[i] : &hdummy))
    switch (fin) {
    case FIN_BICG_INIT:
        s->rTr = t[0]; s->dot_r = t[0]; s->dot_zero = t[0]; s->k = 0;         // solver.c:78-83
        BICG_HIST(0) = s->dot_r / s->dot_zero;
        loop_test(s);
        break;
    case FIN_BICG_ALPHA:
        s->rTs = t[0];
        s->alpha = s->rTr / s->rTs;                                           // solver.c:93
        break;
    case FIN_BICG_OMEGA:
        s->rTy = t[0]; s->yTy = t[1];
        s->omega = s->rTy / s->yTy;                                           // solver.c:104
        break;
    case FIN_BICG_BETA:
        s->dot_r = t[0];
        s->rTr_old = s->rTr;                                                  // solver.c:110
        s->rTr = t[1];
        s->beta = (s->alpha / s->omega) * (s->rTr / s->rTr_old);              // solver.c:116
        s->k += 1;
        BICG_HIST(s->k) = s->dot_r / s->dot_zero;
        loop_test(s);
        break;
    case FIN_STORE_RTR:
        s->rTr = t[0];
        break;
    case FIN_CAPIPE_INIT:
        s->rTw = t[0];
        s->alpha = s->rTr / s->rTw;                                           // solver.c:210
        s->beta = 0.0; s->omega = 0.0;                                        // :211 (omega pinned)
        s->dot_r = s->rTr; s->dot_zero = s->rTr; s->k = 0;                    // :212-213
        BICG_HIST(0) = s->dot_r / s->dot_zero;
        loop_test(s);
        break;
    case FIN_OMEGA2:
        s->rTw = t[0]; s->wTw = t[1];
        s->omega = s->rTw / s->wTw;                                           // solver.c:232 / 369
        break;
    case FIN_CAPIPE_END:
        s->rTr_old = s->rTr;                                                  // solver.c:239
        s->rTr = t[0]; s->rTw = t[1]; s->rTs = t[2]; s->rTz = t[3];
        s->dot_r = t[4];
        s->beta = (s->alpha / s->omega) * (s->rTr / s->rTr_old);              // :248
        s->alpha = s->rTr / (s->rTw + s->beta * (s->rTs - s->omega * s->rTz)); // :249
        s->k += 1;
        BICG_HIST(s->k) = s->dot_r / s->dot_zero;
        loop_test(s);
        break;
    case FIN_STORE_PEND:
        for (int k = 0; k < MAX_DOTS; ++k) s->pend[k] = t[k];
        break;
    default: break;
    }
#undef BICG_HIST
}

// Wait until every peer in recv_mask has published halo epoch >= `expect` (called by warp 0 of a CTA).
__device__ __forceinline__ bool halo_wait_epoch(const CommDev &c, unsigned expect)
{
    const int lane = threadIdx.x & 31;
    bool ok = true;
    if (lane < c.world && ((c.recv_mask >> lane) & 1u)) {
        const unsigned long long *f = &c.hflag[c.rank][lane].epoch;
        const unsigned long long t0 = globaltimer_ns();
        while (ld_acquire_sys(f) < (unsigned long long)expect) {
            if (globaltimer_ns() - t0 > c.timeout_ns) { ok = false; break; }
        }
    }
    return __all_sync(0xffffffffu, ok);
}

__device__ __forceinline__ bool halo_wait(const CommDev &c, unsigned expect) { return halo_wait_epoch(c, expect); }

// What the elected warp does once the grid's dots are combined: `tot` holds the totals in every lane.
// Runs the TailDesc (cross-GPU reduction over the peer mailboxes, scalar recurrence, halo-ready signal) and, when
// wait_halo is set (persistent kernel), also waits for the peers' halo signal of the same epoch.
template <int NDOT>
__device__ __forceinline__ void tail_warp(const KernelCommon &kc, double (&tot)[NDOT > 0 ? NDOT : 1], bool wait_halo)
{
    Scalars *sc = kc.sc;
    const int tid = threadIdx.x & 31;
    __shared__ double s_vals[MAIL_VALS];
    const TailDesc &td = kc.tail;
    const int lane = tid;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NDOT; ++k) s_vals[k] = tot[k];
        if (td.op == TAIL_ALLREDUCE || td.op == TAIL_POST)
            for (int k = 0; k < td.npend; ++k) s_vals[td.ndot + k] = sc->pend[k];
    }
    __syncwarp();

    bool ok = true;
    switch (td.op) {
    case TAIL_PEND:
        if (lane == 0) for (int k = 0; k < td.ndot; ++k) sc->pend[td.pend_off + k] = s_vals[k];
        break;
    case TAIL_ALLREDUCE: {
        const int nv = td.ndot + td.npend;
        unsigned ep = sc->red_epoch + 1u;
        xg_post(kc.comm, ep, s_vals, nv);
        ok = xg_wait_sum(kc.comm, ep, s_vals, nv);
        __syncwarp();
        if (lane == 0) { sc->red_epoch = ep; sc->red_done = ep; }
        break;
    }
    case TAIL_POST: {
        const int nv = td.ndot + td.npend;
        unsigned ep = sc->red_epoch + 1u;
        xg_post(kc.comm, ep, s_vals, nv);
        if (lane == 0) sc->red_epoch = ep;
        break;
    }
    case TAIL_COMPLETE: {
        unsigned ep = sc->red_done + 1u;
        ok = xg_wait_sum(kc.comm, ep, s_vals, td.nred);
        __syncwarp();
        if (lane == 0) sc->red_done = ep;
        break;
    }
    default: break;
    }

    if (lane == 0) {
        if (!ok) { sc->error = 1; sc->done = 1; }
        else if (td.fin != FIN_NONE) finalize(td.fin, sc, kc.hist, s_vals);
    }
    if (td.signal_halo) {
        // all CTAs fenced their peer stores before taking a ticket; publish the new epoch to receivers
        const unsigned he = sc->halo_epoch + 1u;
        // every CTA that pushed fenced at system scope BEFORE its ticket, and this warp observed all tickets; the
        // flag itself is a system-scope RELEASE store so the pattern is a proper release chain at system scope
        // (one warp per kernel pays for it)
        if (lane < kc.comm.world && ((kc.comm.send_mask >> lane) & 1u))
            st_release_sys(&kc.comm.hflag[lane][kc.comm.rank].epoch, (unsigned long long)he);
        __syncwarp();
        if (wait_halo && !halo_wait_epoch(kc.comm, he) && lane == 0) { sc->error = 1; sc->done = 1; }
        if (lane == 0) sc->halo_epoch = he;
    }
}


// ------------------------------------------------------------------------------------------------
// kernel tail: elect the last CTA, combine the grid's partial dots in a fixed order, run the TailDesc
// ------------------------------------------------------------------------------------------------
// `local` holds this CTA's dots (valid in warp 0 after block_sum).  `scratch`: >= 32*MAX_DOTS doubles.
// Every thread of the CTA must call this.  All global writes of the CTA that the tail's signals cover
// (halo pushes) are ordered by the CTA barrier + thread 0's system-scope fence at the top of this function.
template <int NDOT>
__device__ __forceinline__ void kernel_tail(const KernelCommon &kc, double (&local)[NDOT > 0 ? NDOT : 1],
                                            double *scratch)
{
    __shared__ int s_last;
    Scalars *sc = kc.sc;
    const int tid = threadIdx.x;
    __syncthreads();                     // every thread's global / peer stores of this CTA happen-before the fence
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NDOT; ++k) __stcg(&kc.partials[(size_t)blockIdx.x * MAX_DOTS + k], local[k]);
        if (kc.tail.signal_halo) __threadfence_system();   // peer (NVLink) stores of the halo push
        else __threadfence();
        const unsigned t = atomicAdd(&sc->ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    // grid-level combination: thread t sums CTAs t, t+T, ... ascending, then a fixed CTA tree
    double tot[NDOT > 0 ? NDOT : 1];
#pragma unroll
    for (int k = 0; k < (NDOT > 0 ? NDOT : 1); ++k) tot[k] = 0.0;
    if (NDOT > 0) {
        for (unsigned b = tid; b < gridDim.x; b += blockDim.x) {
#pragma unroll
            for (int k = 0; k < NDOT; ++k) tot[k] += __ldcg(&kc.partials[(size_t)b * MAX_DOTS + k]);
        }
        block_sum<(NDOT > 0 ? NDOT : 1)>(tot, scratch);
    }
    if (tid >= 32) return;                       // warp 0 finishes the job
    tail_warp<NDOT>(kc, tot, false);
    if (tid == 0) { __threadfence(); sc->ticket = 0u; }
}

} // namespace bicg
