// spmv.cuh -- argument block + launch interface of the CSR SpMV kernels (spmv.cu)
#pragma once
#include "dev.cuh"

namespace bicg {

// dots fused into the SpMV epilogue: dot[k] += U[row] * V[row] where U, V are either one of up to four
// epilogue vectors (index into vec[]) or the y just computed (index -1)
struct EpiArgs {
    int nvec;
    const double *vec[4];
    int ndot;
    int ia[4], ib[4];
};

struct SpmvArgs {
    KernelCommon kc;
    // device CSR of this rank's rows over the extended local column space [own columns | ghost columns]
    // (diag block entries first, then offd entries, matrix.c:437-440 order).  val/col are padded by >= 8
    // entries so 16-byte aligned over-reads of a tile stay in bounds.
    const double   *val;
    const unsigned *col;
    const unsigned *ptr;
    int rows;
    // tile plan of the TMA kernel
    const int      *tile_row;   // ntiles + 1
    const unsigned *tile_nz;    // ntiles + 1 : ptr[tile_row[t]]
    int ntiles;
    int cap;                    // stage capacity in entries (multiple of 32)
    int stages;                 // 2..4
    const double *x;            // input vector, extended layout
    double       *y;            // output, own rows
    EpiArgs epi;
    int wait_halo;              // 1: x's ghost part is filled by peers; wait for their halo flags first
    const double *shift_sigma;  // not null: y = A x + (*shift_sigma) x  (shifted systems, shifted_switching_solver.c:386, 404)
};

// kind 0: warp-specialised TMA tile kernel, kind 1: row-split kernel.  threads (consumer threads) only matters for kind 0.
// Returns cudaError_t as int.
int launch_spmv(int kind, int lanes, int threads, int grid, size_t smem_bytes, const SpmvArgs &a, cudaStream_t st);
// one-time opt-in to > 48 KB dynamic shared memory for every instantiation
int spmv_setup_attributes();
size_t spmv_tma_smem_bytes(int cap, int stages, int threads, int lanes);
// add a dot term (a . b) to an epilogue; null pointer = the y just computed
void epi_add_dot(EpiArgs &e, const double *a, const double *b);

} // namespace bicg
