// plan.cpp -- host-side planning that needs no CUDA: row partition, SpMV tile plan, halo plan.
// Exposed through the C ABI so the CPU-only test-suite can check it (include/bicgstab_b200.h, Part 2).
#include "bicgstab_b200.h"
#include "plan.hpp"

#include <algorithm>
#include <cstring>
#include <vector>

// matrix.c:295-308 -- the first (n % world) ranks own one extra row; blocks are contiguous.
extern "C" void bicg_plan_partition(int n, int world, int *counts, int *displs)
{
    const int base = n / world, extra = n % world;
    for (int p = 0; p < world; ++p) {
        counts[p] = base + (p < extra ? 1 : 0);
        displs[p] = p * base + std::min(p, extra);
    }
}

namespace bicg {

// Greedy tiling: a tile closes when adding the next row would exceed rows_per_tile rows or cap_nnz
// entries.  Rows longer than cap_nnz cannot be staged and are reported to the caller.
int plan_tiles(const unsigned *ptr, int rows, int rows_per_tile, int cap_nnz, std::vector<int> &tile_row)
{
    tile_row.clear();
    tile_row.push_back(0);
    int r0 = 0;
    while (r0 < rows) {
        int r1 = r0;
        const unsigned base = ptr[r0];
        while (r1 < rows && r1 - r0 < rows_per_tile && ptr[r1 + 1] - base <= (unsigned)cap_nnz) ++r1;
        if (r1 == r0) return -2;          // single row longer than a stage
        tile_row.push_back(r1);
        r0 = r1;
    }
    return (int)tile_row.size() - 1;
}

// Runs of global columns this rank must receive.  Columns inside the rank's own range never occur in
// an offd block (matrix.c:387-388) and are ignored if they do.
void plan_halo_runs(const CSR_Matrix *offd, const INFO_Matrix *info, int world, int gap, int self,
                    std::vector<HaloRun> &runs)
{
    runs.clear();
    if (!offd || offd->nz == 0 || world == 1) return;
    const unsigned n = info->cols;
    std::vector<unsigned char> need(n, 0);
    for (unsigned j = 0; j < offd->nz; ++j) need[offd->col[j]] = 1;
    for (int p = 0; p < world; ++p) {
        if (p == self) continue;
        const unsigned lo = (unsigned)info->displs[p], hi = lo + (unsigned)info->recvcounts[p];
        unsigned c = lo;
        while (c < hi) {
            if (!need[c]) { ++c; continue; }
            unsigned first = c, last = c;         // grow the run while the next needed column is within `gap`
            unsigned probe = c + 1;
            while (probe < hi && probe - last <= (unsigned)gap + 1) {
                if (need[probe]) last = probe;
                ++probe;
            }
            runs.push_back({(int)first, (int)(last - first + 1), p});
            c = last + 1;
        }
    }
}

} // namespace bicg

extern "C" int bicg_plan_tiles(const unsigned int *ptr, int rows, int rows_per_tile, int cap_nnz,
                               int *tile_row, int tile_row_cap)
{
    std::vector<int> t;
    int nt = bicg::plan_tiles(ptr, rows, rows_per_tile, cap_nnz, t);
    if (nt < 0) return nt;
    if ((int)t.size() > tile_row_cap) return -1;
    std::memcpy(tile_row, t.data(), t.size() * sizeof(int));
    return nt;
}

extern "C" int bicg_plan_halo_runs(const CSR_Matrix *offd, const INFO_Matrix *info, int self, int world,
                                   int gap, int *runs_out, int runs_cap)
{
    std::vector<bicg::HaloRun> runs;
    bicg::plan_halo_runs(offd, info, world, gap, self, runs);
    if ((int)runs.size() > runs_cap) return -(int)runs.size();
    for (size_t i = 0; i < runs.size(); ++i) {
        runs_out[3 * i + 0] = runs[i].first;
        runs_out[3 * i + 1] = runs[i].len;
        runs_out[3 * i + 2] = runs[i].owner;
    }
    return (int)runs.size();
}
