// vec.cu -- the BLAS-1 chains of solver.c as fused single-pass kernels (sm_100a).
//
// The reference walks each length-n vector once per my_daxpy / my_dscal / my_ddot call (vector.c:3-27):
// 7 / 11 / 16 separate passes per iteration for bicgstab / ca_bicgstab / pipe_bicgstab.  Here each group of
// calls between two synchronisation points is ONE kernel: every vector is read once (double2 loads) and
// written once, the dot products that follow are accumulated in the same pass, the grid's partial sums are
// combined in a fixed order by the last CTA, which then also performs the cross-GPU reduction over peer
// memory, evaluates the scalar recurrence on the device (dev.cuh: kernel_tail) and, when the updated vector
// is the next SpMV's input, pushes the boundary runs the neighbours need straight into their ghost regions
// over NVLink (the reference: MPI_Iallgatherv of the whole vector, matrix.c:432).
//
// Per element the operation order and the FMA contraction match what gcc emits for the reference
// (y += a*x -> fma(a, x, y); x *= a -> a*x), so elementwise results are bitwise those of the CPU code.
//
//   phase          reference lines replaced
//   PH_BICG_INIT   solver.c:75-78          PH_BICG_Q    :94              PH_BICG_XR  :105-111
//   PH_BICG_P      :117-119                PH_INIT_R    :201-203         PH_CA_PS    :217-222
//   PH_QY          :225-228 / 361-364      PH_CA_XR     :233-236         PH_PIPE_1   :352-364
//   PH_PIPE_3      :370-380                PH_RR_P      :494-496         PH_RR_X     :518-519
//   PH_RR_R        :524-525                PH_RR_DOTS   :533-539
#include "vec.cuh"

namespace bicg {

namespace {

template <int W> struct Pk { double v[W]; };

template <int W> __device__ __forceinline__ Pk<W> ld(const double *p, int i);
template <> __device__ __forceinline__ Pk<1> ld<1>(const double *p, int i) { Pk<1> r; r.v[0] = p[i]; return r; }
template <> __device__ __forceinline__ Pk<2> ld<2>(const double *p, int i)
{
    const double2 t = *reinterpret_cast<const double2 *>(p + i);
    Pk<2> r; r.v[0] = t.x; r.v[1] = t.y; return r;
}
template <> __device__ __forceinline__ Pk<4> ld<4>(const double *p, int i)
{
    const double2 t = *reinterpret_cast<const double2 *>(p + i);
    const double2 u = *reinterpret_cast<const double2 *>(p + i + 2);
    Pk<4> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = u.x; r.v[3] = u.y; return r;
}
template <int W> __device__ __forceinline__ void st(double *p, int i, const Pk<W> &a);
template <> __device__ __forceinline__ void st<1>(double *p, int i, const Pk<1> &a) { p[i] = a.v[0]; }
template <> __device__ __forceinline__ void st<2>(double *p, int i, const Pk<2> &a)
{
    *reinterpret_cast<double2 *>(p + i) = make_double2(a.v[0], a.v[1]);
}
template <> __device__ __forceinline__ void st<4>(double *p, int i, const Pk<4> &a)
{
    *reinterpret_cast<double2 *>(p + i) = make_double2(a.v[0], a.v[1]);
    *reinterpret_cast<double2 *>(p + i + 2) = make_double2(a.v[2], a.v[3]);
}

__host__ __device__ constexpr int phase_ndot(int ph)
{
    return ph == PH_BICG_INIT ? 1 : ph == PH_BICG_XR ? 2 : ph == PH_INIT_R ? 1 : ph == PH_QY ? 2
         : ph == PH_CA_XR ? 1 : ph == PH_PIPE_1 ? 2 : ph == PH_PIPE_3 ? 5 : ph == PH_RR_DOTS ? 5 : 0;
}

struct Coef { double al, be, om, nbo; };

template <int PH, int W>
__device__ __forceinline__ void body(const VecPtrs &v, int i, const Coef &c, double *dot)
{
    if constexpr (PH == PH_BICG_INIT || PH == PH_INIT_R) {
        Pk<W> ax = ld<W>(v.ax, i), r = ld<W>(v.r, i);
#pragma unroll
        for (int k = 0; k < W; ++k) {
            r.v[k] = fma(-1.0, ax.v[k], r.v[k]);
            dot[0] = fma(r.v[k], r.v[k], dot[0]);
        }
        st<W>(v.r, i, r); st<W>(v.rh, i, r);
        if constexpr (PH == PH_BICG_INIT) st<W>(v.p, i, r);
    } else if constexpr (PH == PH_BICG_Q) {
        Pk<W> s = ld<W>(v.s, i), r = ld<W>(v.r, i);
#pragma unroll
        for (int k = 0; k < W; ++k) r.v[k] = fma(-c.al, s.v[k], r.v[k]);
        st<W>(v.r, i, r);
    } else if constexpr (PH == PH_BICG_XR) {
        Pk<W> x = ld<W>(v.x, i), p = ld<W>(v.p, i), r = ld<W>(v.r, i), y = ld<W>(v.y, i), rh = ld<W>(v.rh, i);
#pragma unroll
        for (int k = 0; k < W; ++k) {
            x.v[k] = fma(c.al, p.v[k], x.v[k]);
            x.v[k] = fma(c.om, r.v[k], x.v[k]);
            r.v[k] = fma(-c.om, y.v[k], r.v[k]);
            dot[0] = fma(r.v[k], r.v[k], dot[0]);
            dot[1] = fma(rh.v[k], r.v[k], dot[1]);
        }
        st<W>(v.x, i, x); st<W>(v.r, i, r);
    } else if constexpr (PH == PH_BICG_P) {
        Pk<W> p = ld<W>(v.p, i), r = ld<W>(v.r, i), s = ld<W>(v.s, i);
#pragma unroll
        for (int k = 0; k < W; ++k) {
            double t = c.be * p.v[k];
            t = fma(1.0, r.v[k], t);
            p.v[k] = fma(c.nbo, s.v[k], t);
        }
        st<W>(v.p, i, p);
    } else if constexpr (PH == PH_CA_PS) {
        Pk<W> p = ld<W>(v.p, i), s = ld<W>(v.s, i), z = ld<W>(v.z, i), r = ld<W>(v.r, i), w = ld<W>(v.w, i);
#pragma unroll
        for (int k = 0; k < W; ++k) {
            double t = fma(-c.om, s.v[k], p.v[k]);
            t = c.be * t;
            p.v[k] = fma(1.0, r.v[k], t);
            double u = fma(-c.om, z.v[k], s.v[k]);
            u = c.be * u;
            s.v[k] = fma(1.0, w.v[k], u);
        }
        st<W>(v.p, i, p); st<W>(v.s, i, s);
    } else if constexpr (PH == PH_QY) {
        Pk<W> r = ld<W>(v.r, i), s = ld<W>(v.s, i), w = ld<W>(v.w, i), z = ld<W>(v.z, i);
#pragma unroll
        for (int k = 0; k < W; ++k) {
            r.v[k] = fma(-c.al, s.v[k], r.v[k]);
            w.v[k] = fma(-c.al, z.v[k], w.v[k]);
            dot[0] = fma(r.v[k], w.v[k], dot[0]);
            dot[1] = fma(w.v[k], w.v[k], dot[1]);
        }
        st<W>(v.r, i, r); st<W>(v.w, i, w);
    } else if constexpr (PH == PH_CA_XR) {
        Pk<W> x = ld<W>(v.x, i), p = ld<W>(v.p, i), r = ld<W>(v.r, i), w = ld<W>(v.w, i);
#pragma unroll
        for (int k = 0; k < W; ++k) {
            x.v[k] = fma(c.al, p.v[k], x.v[k]);
            x.v[k] = fma(c.om, r.v[k], x.v[k]);
            r.v[k] = fma(-c.om, w.v[k], r.v[k]);
            dot[0] = fma(r.v[k], r.v[k], dot[0]);
        }
        st<W>(v.x, i, x); st<W>(v.r, i, r);
    } else if constexpr (PH == PH_PIPE_1) {
        Pk<W> p = ld<W>(v.p, i), s = ld<W>(v.s, i), z = ld<W>(v.z, i), vv = ld<W>(v.v, i), t = ld<W>(v.t, i),
              r = ld<W>(v.r, i), w = ld<W>(v.w, i);
#pragma unroll
        for (int k = 0; k < W; ++k) {
            double a = fma(-c.om, s.v[k], p.v[k]);  a = c.be * a;  p.v[k] = fma(1.0, r.v[k], a);
            double b = fma(-c.om, z.v[k], s.v[k]);  b = c.be * b;  s.v[k] = fma(1.0, w.v[k], b);
            double d = fma(-c.om, vv.v[k], z.v[k]); d = c.be * d;  z.v[k] = fma(1.0, t.v[k], d);
            r.v[k] = fma(-c.al, s.v[k], r.v[k]);
            w.v[k] = fma(-c.al, z.v[k], w.v[k]);
            dot[0] = fma(r.v[k], w.v[k], dot[0]);
            dot[1] = fma(w.v[k], w.v[k], dot[1]);
        }
        st<W>(v.p, i, p); st<W>(v.s, i, s); st<W>(v.z, i, z); st<W>(v.r, i, r); st<W>(v.w, i, w);
    } else if constexpr (PH == PH_PIPE_3) {
        Pk<W> x = ld<W>(v.x, i), p = ld<W>(v.p, i), r = ld<W>(v.r, i), w = ld<W>(v.w, i), t = ld<W>(v.t, i),
              vv = ld<W>(v.v, i), rh = ld<W>(v.rh, i), s = ld<W>(v.s, i), z = ld<W>(v.z, i);
#pragma unroll
        for (int k = 0; k < W; ++k) {
            x.v[k] = fma(c.al, p.v[k], x.v[k]);
            x.v[k] = fma(c.om, r.v[k], x.v[k]);
            r.v[k] = fma(-c.om, w.v[k], r.v[k]);
            const double tt = fma(-c.al, vv.v[k], t.v[k]);     // t - alpha v; t itself is overwritten by t = A w next
            w.v[k] = fma(-c.om, tt, w.v[k]);
            dot[0] = fma(rh.v[k], r.v[k], dot[0]);     // order expected by FIN_CAPIPE_END
            dot[1] = fma(rh.v[k], w.v[k], dot[1]);
            dot[2] = fma(rh.v[k], s.v[k], dot[2]);
            dot[3] = fma(rh.v[k], z.v[k], dot[3]);
            dot[4] = fma(r.v[k], r.v[k], dot[4]);
        }
        st<W>(v.x, i, x); st<W>(v.r, i, r); st<W>(v.w, i, w);
    } else if constexpr (PH == PH_RR_P) {
        Pk<W> p = ld<W>(v.p, i), s = ld<W>(v.s, i), r = ld<W>(v.r, i);
#pragma unroll
        for (int k = 0; k < W; ++k) {
            double a = fma(-c.om, s.v[k], p.v[k]); a = c.be * a; p.v[k] = fma(1.0, r.v[k], a);
        }
        st<W>(v.p, i, p);
    } else if constexpr (PH == PH_RR_X) {
        Pk<W> x = ld<W>(v.x, i), p = ld<W>(v.p, i), r = ld<W>(v.r, i);
#pragma unroll
        for (int k = 0; k < W; ++k) {
            x.v[k] = fma(c.al, p.v[k], x.v[k]);
            x.v[k] = fma(c.om, r.v[k], x.v[k]);
        }
        st<W>(v.x, i, x);
    } else if constexpr (PH == PH_RR_R) {
        Pk<W> b = ld<W>(v.b, i), ax = ld<W>(v.ax, i);
#pragma unroll
        for (int k = 0; k < W; ++k) b.v[k] = fma(-1.0, ax.v[k], b.v[k]);
        st<W>(v.r, i, b);
    } else if constexpr (PH == PH_RR_DOTS) {
        Pk<W> r = ld<W>(v.r, i), rh = ld<W>(v.rh, i), w = ld<W>(v.w, i), s = ld<W>(v.s, i), z = ld<W>(v.z, i);
#pragma unroll
        for (int k = 0; k < W; ++k) {
            dot[0] = fma(rh.v[k], r.v[k], dot[0]);     // order expected by FIN_CAPIPE_END
            dot[1] = fma(rh.v[k], w.v[k], dot[1]);
            dot[2] = fma(rh.v[k], s.v[k], dot[2]);
            dot[3] = fma(rh.v[k], z.v[k], dot[3]);
            dot[4] = fma(r.v[k], r.v[k], dot[4]);
        }
    }
    (void)v; (void)i; (void)c; (void)dot;
}

// copy the parts of this CTA's chunk [lo, hi) that peers need into their ghost regions (peer stores)
__device__ __forceinline__ void push_chunk(const PushDesc &pd, int lo, int hi)
{
    bool stored = false;
    for (int pi = 0; pi < pd.npeers; ++pi) {
        const PushRun *runs = pd.runs[pi];
        const int nr = pd.nruns[pi];
        int a = 0, b = nr;                       // first run that ends after lo
        while (a < b) {
            const int m = (a + b) >> 1;
            if (runs[m].src + runs[m].len <= lo) a = m + 1; else b = m;
        }
        for (int ri = a; ri < nr; ++ri) {
            const PushRun r = runs[ri];
            if (r.src >= hi) break;
            const int s = max(r.src, lo), e = min(r.src + r.len, hi);
            double *d = pd.dst[pi] + ((long long)r.dst_off - (long long)r.src);
            for (int i = s + (int)threadIdx.x; i < e; i += (int)blockDim.x) { d[i] = __ldcg(pd.src + i); stored = true; }
        }
    }
    // a thread that wrote to a peer orders its own NVLink stores before anything that follows (the halo flag
    // is released by the tail after a CTA barrier, a grid-wide ticket and another system fence)
    if (stored && pd.fence_writers) __threadfence_system();
}

template <int PH>
__global__ void __launch_bounds__(256) vec_kernel(const __grid_constant__ VecArgs a)
{
    const Scalars *sc = a.kc.sc;
    if (sc->done) return;
    __shared__ double scratch[32 * MAX_DOTS];
    constexpr int ND = phase_ndot(PH);
    constexpr int NDA = ND > 0 ? ND : 1;

    Coef c;
    c.al = sc->alpha; c.be = sc->beta; c.om = sc->omega; c.nbo = -c.be * c.om;   // solver.c:119  -beta*omega

    const int lo = (int)blockIdx.x * a.chunk;
    const int hi = min(a.n, lo + a.chunk);
    double dot[NDA];
#pragma unroll
    for (int k = 0; k < NDA; ++k) dot[k] = 0.0;

    if constexpr (PH != PH_PUSH) {
        // 4 doubles (two 16-byte loads per vector) per thread and step: twice the bytes in flight of a
        // double2 loop; lo is a multiple of 4 and every vector is 128-byte aligned
        int i = lo + 4 * (int)threadIdx.x;
        for (; i + 3 < hi; i += 4 * (int)blockDim.x) body<PH, 4>(a.v, i, c, dot);
        for (; i < hi; ++i) body<PH, 1>(a.v, i, c, dot);  // < 4 trailing elements of the last chunk (one thread)
    }

    const bool pushing = a.push.npeers > 0;
    if (pushing) {
        __syncthreads();                 // this CTA's elements are final
        push_chunk(a.push, lo, hi);
        // the peer stores are ordered before the halo flag by ONE system-scope fence per CTA: kernel_tail's
        // thread 0 fences after the __syncthreads that follows (cumulativity covers the whole CTA's stores);
        // a fence.sys in every thread costs microseconds per kernel
    }
    if (ND == 0 && a.kc.tail.op == TAIL_NONE && !a.kc.tail.signal_halo) return;
    if (ND > 0) block_sum<NDA>(dot, scratch);
    kernel_tail<ND>(a.kc, dot, scratch);
}

template <int PH>
cudaError_t launch(int grid, const VecArgs &a, cudaStream_t st)
{
    vec_kernel<PH><<<grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}

} // namespace

int launch_vec(int phase, int grid, const VecArgs &a, cudaStream_t st)
{
    switch (phase) {
    case PH_BICG_INIT: return (int)launch<PH_BICG_INIT>(grid, a, st);
    case PH_BICG_Q:    return (int)launch<PH_BICG_Q>(grid, a, st);
    case PH_BICG_XR:   return (int)launch<PH_BICG_XR>(grid, a, st);
    case PH_BICG_P:    return (int)launch<PH_BICG_P>(grid, a, st);
    case PH_INIT_R:    return (int)launch<PH_INIT_R>(grid, a, st);
    case PH_CA_PS:     return (int)launch<PH_CA_PS>(grid, a, st);
    case PH_QY:        return (int)launch<PH_QY>(grid, a, st);
    case PH_CA_XR:     return (int)launch<PH_CA_XR>(grid, a, st);
    case PH_PIPE_1:    return (int)launch<PH_PIPE_1>(grid, a, st);
    case PH_PIPE_3:    return (int)launch<PH_PIPE_3>(grid, a, st);
    case PH_RR_P:      return (int)launch<PH_RR_P>(grid, a, st);
    case PH_RR_X:      return (int)launch<PH_RR_X>(grid, a, st);
    case PH_RR_R:      return (int)launch<PH_RR_R>(grid, a, st);
    case PH_RR_DOTS:   return (int)launch<PH_RR_DOTS>(grid, a, st);
    case PH_PUSH:      return (int)launch<PH_PUSH>(grid, a, st);
    default:           return (int)cudaErrorInvalidValue;
    }
}

} // namespace bicg
