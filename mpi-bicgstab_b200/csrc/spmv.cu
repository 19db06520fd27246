// spmv.cu -- fp64 CSR SpMV for sm_100a with the dot products of the solver fused into its epilogue and the
// cross-GPU reduction / scalar recurrence in its tail.  Replaces mult() + MPI_csr_spmv_ovlap()
// (matrix.c:498-516, 428-441) and the my_ddot + MPI_Iallreduce pairs that follow them (solver.c:88-91,
// 96-102, 238-247, 365-367, 381-385).
//
// Two kernels:
//
//  spmv_tma_kernel<LANES, THREADS>   (kind 0, the default)
//      Persistent CTAs walk a precomputed tile plan (<= THREADS/LANES rows and <= cap entries per tile).
//      One elected thread streams each tile's val[] / col[] slices from HBM into a multi-stage shared-memory
//      ring with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx; UBLKCP in SASS), so the DRAM
//      stream is fully coalesced, asynchronous and independent of the row structure.  LANES threads then
//      consume one row from shared memory.  With LANES = 1 a warp's 32 gathers x[col] at step j hit the
//      same stencil offset of 32 consecutive rows -> 2 cache lines instead of ~15 for banded matrices, and
//      the row is summed left to right exactly like the reference's scalar loop.  LANES > 1 is for long /
//      irregular rows (shuffle reduction inside the LANES group).
//
//  spmv_rowsplit_kernel<LANES>       (kind 1)
//      Classic sub-warp-per-row kernel reading val/col straight from global memory; fallback for matrices
//      with rows longer than a stage, and the comparison point for the TMA kernel.
//
// Both write y exactly once (no zero-fill + accumulate passes as in matrix.c:434-440).
#include "spmv.cuh"

namespace bicg {

namespace {

template <int LANES>
__device__ __forceinline__ double lanes_sum(double v)
{
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void row_epilogue(const SpmvArgs &a, int row, double yi, double (&dot)[4])
{
    a.y[row] = yi;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < a.epi.ndot) {
            const double av = a.epi.a[k] ? a.epi.a[k][row] : yi;
            const double bv = a.epi.b[k] ? a.epi.b[k][row] : yi;
            dot[k] = fma(av, bv, dot[k]);
        }
    }
}

__device__ __forceinline__ bool needs_tail(const KernelCommon &kc)
{
    return kc.tail.op != TAIL_NONE || kc.tail.signal_halo;
}

template <int LANES, int THREADS>
__global__ void __launch_bounds__(THREADS) spmv_tma_kernel(const __grid_constant__ SpmvArgs a)
{
    if (a.kc.sc->done) return;

    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ __align__(8) unsigned long long bars[4];
    __shared__ double scratch[32 * 4];

    const int tid = threadIdx.x;
    const int lane = tid % LANES;
    const int row_in_tile = tid / LANES;
    const int stages = a.stages, cap = a.cap;
    double   *sval = reinterpret_cast<double *>(dyn_smem);
    unsigned *scol = reinterpret_cast<unsigned *>(dyn_smem + (size_t)stages * cap * sizeof(double));
    const int my_tiles = (a.ntiles > (int)blockIdx.x) ? (a.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(smem_u32(&bars[s]), 1u);
        mbar_fence_init();
    }
    __syncthreads();

    // producer: thread 0 arms the stage's mbarrier with the byte count and fires two bulk copies
    auto produce = [&](int i) {
        const int t = (int)blockIdx.x + i * (int)gridDim.x, s = i % stages;
        const unsigned p0 = a.tile_nz[t], p1 = a.tile_nz[t + 1];
        const unsigned a0 = p0 & ~3u, cnt = ((p1 + 3u) & ~3u) - a0;       // 16-byte aligned window
        const unsigned bar = smem_u32(&bars[s]);
        mbar_arrive_expect_tx(bar, cnt * 12u);
        if (cnt) {
            tma_load_1d(smem_u32(sval + (size_t)s * cap), a.val + a0, cnt * 8u, bar);
            tma_load_1d(smem_u32(scol + (size_t)s * cap), a.col + a0, cnt * 4u, bar);
        }
    };
    if (tid == 0)
        for (int i = 0; i < stages - 1 && i < my_tiles; ++i) produce(i);

    // The matrix stream is already in flight; now make sure the peers' halo values of x have landed
    // (the reference's MPI_Wait on the allgather, matrix.c:439).
    if (a.wait_halo) {
        if (tid < 32) {
            const bool ok = halo_wait(a.kc.comm, a.kc.sc->halo_epoch);
            if (!ok && tid == 0) a.kc.sc->error = 1;
        }
        __syncthreads();
    }

    double dot[4] = {0.0, 0.0, 0.0, 0.0};
    const double *__restrict__ x = a.x;

    for (int i = 0; i < my_tiles; ++i) {
        if (tid == 0 && i + stages - 1 < my_tiles) produce(i + stages - 1);

        const int t = (int)blockIdx.x + i * (int)gridDim.x, s = i % stages;
        const int row0 = a.tile_row[t], row1 = a.tile_row[t + 1];
        const unsigned a0 = a.tile_nz[t] & ~3u;
        const int row = row0 + row_in_tile;
        const bool valid = row < row1;
        unsigned pb = 0, pe = 0;
        if (valid) { pb = a.ptr[row]; pe = a.ptr[row + 1]; }

        mbar_wait(smem_u32(&bars[s]), (unsigned)(i / stages) & 1u);

        const double   *sv = sval + (size_t)s * cap;
        const unsigned *sc = scol + (size_t)s * cap;
        double acc = 0.0;
        int j = (int)(pb - a0) + lane;
        const int e = (int)(pe - a0);
        for (; j + 3 * LANES < e; j += 4 * LANES) {
            const unsigned c0 = sc[j], c1 = sc[j + LANES], c2 = sc[j + 2 * LANES], c3 = sc[j + 3 * LANES];
            const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
            const double v0 = sv[j], v1 = sv[j + LANES], v2 = sv[j + 2 * LANES], v3 = sv[j + 3 * LANES];
            acc = fma(v0, x0, acc);
            acc = fma(v1, x1, acc);
            acc = fma(v2, x2, acc);
            acc = fma(v3, x3, acc);
        }
        for (; j < e; j += LANES) acc = fma(sv[j], __ldg(x + sc[j]), acc);

        acc = lanes_sum<LANES>(acc);
        if (valid && lane == 0) row_epilogue(a, row, acc, dot);

        __syncthreads();     // every thread is done with stage s -> the producer may refill it
    }

    if (!needs_tail(a.kc)) return;
    block_sum<4>(dot, scratch);
    kernel_tail<4>(a.kc, dot, scratch);
}

template <int LANES>
__global__ void __launch_bounds__(256) spmv_rowsplit_kernel(const __grid_constant__ SpmvArgs a)
{
    if (a.kc.sc->done) return;
    __shared__ double scratch[32 * 4];
    const int tid = threadIdx.x;
    if (a.wait_halo) {
        if (tid < 32) {
            const bool ok = halo_wait(a.kc.comm, a.kc.sc->halo_epoch);
            if (!ok && tid == 0) a.kc.sc->error = 1;
        }
        __syncthreads();
    }
    constexpr int RPB = 256 / LANES;
    const int lane = tid % LANES;
    const double *__restrict__ x = a.x;
    const double *__restrict__ val = a.val;
    const unsigned *__restrict__ col = a.col;
    double dot[4] = {0.0, 0.0, 0.0, 0.0};
    for (long long base = (long long)blockIdx.x * RPB; base < a.rows; base += (long long)gridDim.x * RPB) {
        const int row = (int)base + tid / LANES;
        const bool valid = row < a.rows;
        unsigned pb = 0, pe = 0;
        if (valid) { pb = a.ptr[row]; pe = a.ptr[row + 1]; }
        double acc = 0.0;
        unsigned j = pb + lane;
        for (; j + 3 * LANES < pe; j += 4 * LANES) {
            const unsigned c0 = col[j], c1 = col[j + LANES], c2 = col[j + 2 * LANES], c3 = col[j + 3 * LANES];
            const double v0 = val[j], v1 = val[j + LANES], v2 = val[j + 2 * LANES], v3 = val[j + 3 * LANES];
            acc = fma(v0, __ldg(x + c0), acc);
            acc = fma(v1, __ldg(x + c1), acc);
            acc = fma(v2, __ldg(x + c2), acc);
            acc = fma(v3, __ldg(x + c3), acc);
        }
        for (; j < pe; j += LANES) acc = fma(val[j], __ldg(x + col[j]), acc);
        acc = lanes_sum<LANES>(acc);
        if (valid && lane == 0) row_epilogue(a, row, acc, dot);
    }
    if (!needs_tail(a.kc)) return;
    block_sum<4>(dot, scratch);
    kernel_tail<4>(a.kc, dot, scratch);
}

template <int LANES, int THREADS>
cudaError_t launch_tma(int grid, size_t smem, const SpmvArgs &a, cudaStream_t st)
{
    spmv_tma_kernel<LANES, THREADS><<<grid, THREADS, smem, st>>>(a);
    return cudaGetLastError();
}
template <int LANES>
cudaError_t launch_tma_t(int threads, int grid, size_t smem, const SpmvArgs &a, cudaStream_t st)
{
    switch (threads) {
    case 128: return launch_tma<LANES, 128>(grid, smem, a, st);
    case 256: return launch_tma<LANES, 256>(grid, smem, a, st);
    case 512: return launch_tma<LANES, 512>(grid, smem, a, st);
    default:  return cudaErrorInvalidValue;
    }
}
template <int LANES>
cudaError_t launch_rowsplit(int grid, const SpmvArgs &a, cudaStream_t st)
{
    spmv_rowsplit_kernel<LANES><<<grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}

template <int LANES, int THREADS>
cudaError_t set_attr()
{
    // opt-in limit is 227 KB per CTA *including* the kernel's static shared memory
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, spmv_tma_kernel<LANES, THREADS>);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(spmv_tma_kernel<LANES, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                227 * 1024 - (int)fa.sharedSizeBytes);
}
template <int LANES>
cudaError_t set_attr_l()
{
    cudaError_t e;
    if ((e = set_attr<LANES, 128>()) != cudaSuccess) return e;
    if ((e = set_attr<LANES, 256>()) != cudaSuccess) return e;
    return set_attr<LANES, 512>();
}

} // namespace

size_t spmv_tma_smem_bytes(int cap, int stages) { return (size_t)stages * (size_t)cap * 12u; }

int spmv_setup_attributes()
{
    cudaError_t e;
    if ((e = set_attr_l<1>()) != cudaSuccess) return (int)e;
    if ((e = set_attr_l<2>()) != cudaSuccess) return (int)e;
    if ((e = set_attr_l<4>()) != cudaSuccess) return (int)e;
    if ((e = set_attr_l<8>()) != cudaSuccess) return (int)e;
    if ((e = set_attr_l<16>()) != cudaSuccess) return (int)e;
    if ((e = set_attr_l<32>()) != cudaSuccess) return (int)e;
    return 0;
}

int launch_spmv(int kind, int lanes, int threads, int grid, size_t smem, const SpmvArgs &a, cudaStream_t st)
{
    if (kind == 0) {
        switch (lanes) {
        case 1:  return (int)launch_tma_t<1>(threads, grid, smem, a, st);
        case 2:  return (int)launch_tma_t<2>(threads, grid, smem, a, st);
        case 4:  return (int)launch_tma_t<4>(threads, grid, smem, a, st);
        case 8:  return (int)launch_tma_t<8>(threads, grid, smem, a, st);
        case 16: return (int)launch_tma_t<16>(threads, grid, smem, a, st);
        case 32: return (int)launch_tma_t<32>(threads, grid, smem, a, st);
        default: return (int)cudaErrorInvalidValue;
        }
    }
    switch (lanes) {
    case 1:  return (int)launch_rowsplit<1>(grid, a, st);
    case 2:  return (int)launch_rowsplit<2>(grid, a, st);
    case 4:  return (int)launch_rowsplit<4>(grid, a, st);
    case 8:  return (int)launch_rowsplit<8>(grid, a, st);
    case 16: return (int)launch_rowsplit<16>(grid, a, st);
    case 32: return (int)launch_rowsplit<32>(grid, a, st);
    default: return (int)cudaErrorInvalidValue;
    }
}

} // namespace bicg
