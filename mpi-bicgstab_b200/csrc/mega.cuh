// mega.cuh -- argument block + launch interface of the persistent solver kernel (mega.cu)
#pragma once
#include "dev.cuh"
#include "vec.cuh"

namespace bicg {

// grid barrier state (HBM, zero-initialised): arrival counter and generation flag on separate lines, so the
// pollers of `gen` do not slow down the arrivals
struct alignas(256) GridBar { unsigned count; unsigned pad0_[31]; unsigned gen; unsigned pad1_[31]; };

constexpr int MEGA_TRACE_ITERS = 256, MEGA_TRACE_SLOTS = 16;

struct MegaArgs {
    Scalars *sc;
    double  *partials;          // [grid][MAX_DOTS]
    double  *hist;
    CommDev  comm;
    GridBar *bar;
    const double   *val;
    const unsigned *col;
    const unsigned *ptr;
    const int      *tile_row;   // ntiles + 1
    const unsigned *tile_nz;    // ntiles + 1
    const int      *cta_tile;   // grid + 1 : first tile of every CTA (contiguous ownership)
    int cap, stages;
    VecPtrs v;
    PushDesc push_p, push_r, push_s, push_z, push_w;
    int method;                 // 0 bicgstab, 1 ca_bicgstab, 2 pipe_bicgstab
    unsigned long long *trace;  // optional [MEGA_TRACE_ITERS][MEGA_TRACE_SLOTS] globaltimer checkpoints of CTA 0 (BICG_MEGA_TRACE)
};

// fuse_q: experimental 4-barrier BiCGStab (q gathered on the fly in the second SpMV), bicgstab only
int    launch_mega(int threads, bool fuse_q, int grid, size_t smem, const MegaArgs &a, cudaStream_t st);
int    mega_setup_attributes();
size_t mega_smem_bytes(int cap, int stages, int threads);

} // namespace bicg
