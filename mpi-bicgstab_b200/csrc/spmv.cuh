// spmv.cuh -- argument block + launch interface of the CSR SpMV kernels (spmv.cu)
#pragma once
#include "dev.cuh"

namespace bicg {

// dots fused into the SpMV epilogue: dot[k] += A_k[row] * B_k[row]; a null pointer means "the y just computed"
struct EpiArgs {
    int ndot;
    const double *a[4];
    const double *b[4];
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
};

// kind 0: TMA-staged tile kernel, kind 1: row-split kernel.  threads only matters for kind 0.
// Returns cudaError_t as int.
int launch_spmv(int kind, int lanes, int threads, int grid, size_t smem_bytes, const SpmvArgs &a, cudaStream_t st);
// one-time opt-in to > 48 KB dynamic shared memory for every instantiation
int spmv_setup_attributes();
size_t spmv_tma_smem_bytes(int cap, int stages);

} // namespace bicg
