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

// nnz-balanced variant (the reference's archived DYNAMIC_ROWS option, archive/matrix.c:407-420): ranks take consecutive
// rows until their entry count reaches nnz / world; the last rank takes the rest.  Contiguous blocks, so everything
// downstream (diag / offd split, halo plan, recvcounts / displs) is unchanged.
extern "C" void bicg_plan_partition_nnz(const unsigned int *row_nnz, int n, int world, int *counts, int *displs)
{
    unsigned long long total = 0;
    for (int i = 0; i < n; ++i) total += row_nnz[i];
    const unsigned long long target = total / (unsigned long long)world;
    int start = 0;
    for (int p = 0; p < world; ++p) {
        int end = n;
        if (p < world - 1) {
            unsigned long long cum = 0;
            for (int i = start; i < n; ++i) {
                cum += row_nnz[i];
                if (cum >= target) { end = i + 1; break; }
            }
        }
        counts[p] = end - start;
        displs[p] = start;
        start = end;
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

// Merge the reference's two blocks of one rank into a single CSR over the extended local column space:
// columns [0, n_loc) are the rank's own, columns ghost_off + g are ghost slot g.  Per row the diag entries come
// first, then the offd entries -- the order in which the reference accumulates them (matrix.c:437, 440).
// recv gets quadruples (first_col, len, owner, ghost_idx) describing which global columns fill the ghost slots.
void merge_blocks(const CSR_Matrix *diag, const CSR_Matrix *offd, const INFO_Matrix *info, int self, int world, int gap,
                  int ghost_off, std::vector<unsigned> &mptr, std::vector<unsigned> &mcol, std::vector<double> &mval,
                  std::vector<int> &recv, int &n_ghost)
{
    std::vector<HaloRun> runs;
    const bool far = offd && offd->nz > 0 && world > 1;
    if (far) plan_halo_runs(offd, info, world, gap, self, runs);
    recv.clear();
    int ghost = 0;
    for (const HaloRun &r : runs) { recv.insert(recv.end(), {r.first, r.len, r.owner, ghost}); ghost += r.len; }
    n_ghost = ghost;
    const int n_loc = (int)diag->rows;
    const size_t nnz = (size_t)diag->nz + (far ? (size_t)offd->nz : 0);
    mptr.resize((size_t)n_loc + 1); mcol.resize(nnz); mval.resize(nnz);
    std::vector<int> run_first(runs.size());
    for (size_t i = 0; i < runs.size(); ++i) run_first[i] = runs[i].first;
    size_t k = 0;
    mptr[0] = 0;
    for (int i = 0; i < n_loc; ++i) {
        for (unsigned j = diag->ptr[i]; j < diag->ptr[i + 1]; ++j) { mval[k] = diag->val[j]; mcol[k] = diag->col[j]; ++k; }
        if (far)
            for (unsigned j = offd->ptr[i]; j < offd->ptr[i + 1]; ++j) {
                const int gc = (int)offd->col[j];
                const size_t ri = (size_t)(std::upper_bound(run_first.begin(), run_first.end(), gc) - run_first.begin()) - 1;
                mval[k] = offd->val[j];
                mcol[k] = (unsigned)(ghost_off + recv[4 * ri + 3] + (gc - runs[ri].first));
                ++k;
            }
        mptr[(size_t)i + 1] = (unsigned)k;
    }
}

// The layout half of merge_blocks without touching the entries: receive list (quadruples first_col, len, owner,
// ghost_idx), number of ghost slots and the merged row pointer mptr[i] = diag->ptr[i] + offd->ptr[i].  The entries
// themselves are merged on the GPU (matrix.cu: merge_rows_kernel) from the two blocks uploaded as they are.
void plan_merged_layout(const CSR_Matrix *diag, const CSR_Matrix *offd, const INFO_Matrix *info, int self, int world, int gap,
                        std::vector<unsigned> &mptr, std::vector<int> &recv, int &n_ghost)
{
    std::vector<HaloRun> runs;
    const bool far = offd && offd->nz > 0 && world > 1;
    if (far) plan_halo_runs(offd, info, world, gap, self, runs);
    recv.clear();
    int ghost = 0;
    for (const HaloRun &r : runs) { recv.insert(recv.end(), {r.first, r.len, r.owner, ghost}); ghost += r.len; }
    n_ghost = ghost;
    const int n_loc = (int)diag->rows;
    mptr.resize((size_t)n_loc + 1);
    for (int i = 0; i <= n_loc; ++i) mptr[(size_t)i] = diag->ptr[i] + (far ? offd->ptr[i] : 0u);
}

// Tile plan of the persistent solver kernel (mega.cu).  CTA g of `ctas` owns a contiguous row range; the ranges are
// balanced by the cost of a row per iteration -- 24 B per entry (two SpMVs) + row_weight (216 B of vector traffic is the
// pure byte count; measured on B200 the thread-per-row SpMV and the vector phases behave like ~1000 B per row: CTA time
// follows the row count much more than the entry count, profiles/r02_*), plus `extra_weight` for every peer the row is pushed to (NVLink stores are slow per SM, and the CTAs at the partition
// boundary also sit on the critical path of the halo exchange) -- and start at multiples of 16 rows, so every CTA's
// slice of every vector is 128-byte aligned.  Each range is cut into ceil(len / rows_per_tile) tiles of (almost)
// equal height.  tile_row gets the first row of every tile plus a final `rows`; cta_tile[g] is the index of CTA g's
// first tile (cta_tile[ctas] = #tiles).  Returns the largest number of entries in any tile.
unsigned plan_cta_tiles(const unsigned *ptr, int rows, int ctas, int rows_per_tile, const unsigned char *row_extra,
                        int extra_weight, std::vector<int> &tile_row, std::vector<int> &cta_tile, int cap_limit,
                        std::vector<unsigned> *tile_nz, std::vector<int> *tile_flag, int row_weight)
{
    if (row_weight <= 0) row_weight = 216;
    tile_row.clear();
    if (tile_nz) tile_nz->clear();
    if (tile_flag) tile_flag->clear();
    cta_tile.assign((size_t)ctas + 1, 0);
    std::vector<int> first((size_t)ctas + 1, rows);
    first[0] = 0;
    auto weight = [&](int i) -> long long {
        return 24ll * (long long)(ptr[i + 1] - ptr[i]) + (long long)row_weight + (row_extra ? (long long)extra_weight * row_extra[i] : 0ll);
    };
    long long total = 24ll * (long long)(ptr[rows] - ptr[0]) + (long long)row_weight * rows;
    if (row_extra && extra_weight)
        for (int i = 0; i < rows; ++i) total += (long long)extra_weight * row_extra[i];
    {
        long long prefix = 0;         // weight of rows [0, r)
        int r = 0;
        for (int g = 1; g < ctas; ++g) {
            const long long target = (long long)((__int128)total * g / ctas);
            while (r < rows && prefix + weight(r) / 2 < target) { prefix += weight(r); ++r; }
            int cut = (int)(((long long)r + 8) / 16 * 16);           // nearest multiple of 16
            cut = std::min(rows, std::max(cut, first[(size_t)g - 1]));
            first[(size_t)g] = cut;
        }
    }
    // pass 1: equal-height tiles (what regular matrices get)
    unsigned max_tile_nnz = 0;
    for (int g = 0; g < ctas; ++g) {
        const int lo = first[(size_t)g], hi = first[(size_t)g + 1];
        cta_tile[(size_t)g] = (int)tile_row.size();
        const int len = hi - lo;
        const int k = (len + rows_per_tile - 1) / rows_per_tile;
        for (int t = 0; t < k; ++t) {
            const int r0 = (int)(lo + (long long)len * t / k), r1 = (int)(lo + (long long)len * (t + 1) / k);
            tile_row.push_back(r0);
            max_tile_nnz = std::max(max_tile_nnz, ptr[r1] - ptr[r0]);
        }
    }
    cta_tile[(size_t)ctas] = (int)tile_row.size();
    tile_row.push_back(rows);
    if (cap_limit <= 0 || max_tile_nnz <= (unsigned)cap_limit || !tile_nz || !tile_flag) {
        if (tile_nz) { tile_nz->resize(tile_row.size()); for (size_t i = 0; i < tile_row.size(); ++i) (*tile_nz)[i] = ptr[tile_row[i]]; }
        if (tile_flag) tile_flag->assign(tile_row.size(), 0);
        return max_tile_nnz;
    }
    // pass 2 (irregular matrices): some tile does not fit a shared-memory stage.  Greedy tiles of <= rows_per_tile rows
    // and <= cap_limit entries; a row longer than cap_limit becomes a run of CHUNK tiles (flag 1: more chunks of this
    // row follow, flag 2: last chunk) that the whole CTA multiplies cooperatively, carrying the row's partial sum from
    // chunk to chunk (mega.cu).  tile_nz then holds the first ENTRY of every tile (rows alone no longer determine it).
    tile_row.clear(); tile_nz->clear(); tile_flag->clear();
    max_tile_nnz = 0;
    for (int g = 0; g < ctas; ++g) {
        const int lo = first[(size_t)g], hi = first[(size_t)g + 1];
        cta_tile[(size_t)g] = (int)tile_row.size();
        int r0 = lo;
        while (r0 < hi) {
            const unsigned base = ptr[r0];
            const unsigned len0 = ptr[r0 + 1] - base;
            if (len0 > (unsigned)cap_limit) {
                for (unsigned off = 0; off < len0; off += (unsigned)cap_limit) {
                    tile_row.push_back(r0); tile_nz->push_back(base + off);
                    tile_flag->push_back(off + (unsigned)cap_limit < len0 ? 1 : 2);
                    max_tile_nnz = std::max(max_tile_nnz, std::min<unsigned>((unsigned)cap_limit, len0 - off));
                }
                ++r0;
                continue;
            }
            int r1 = r0;
            while (r1 < hi && r1 - r0 < rows_per_tile && ptr[r1 + 1] - base <= (unsigned)cap_limit) ++r1;
            tile_row.push_back(r0); tile_nz->push_back(base); tile_flag->push_back(0);
            max_tile_nnz = std::max(max_tile_nnz, ptr[r1] - base);
            r0 = r1;
        }
    }
    cta_tile[(size_t)ctas] = (int)tile_row.size();
    tile_row.push_back(rows); tile_nz->push_back(ptr[rows]); tile_flag->push_back(0);
    return max_tile_nnz;
}

// From every rank's receive list (quadruples first_col, len, owner, ghost_idx; `stride` ints per rank, cnts[p]
// valid quadruples) derive what rank `self` must push to rank `dest`: local source run -> ghost offset on dest.
void plan_push_runs(const int *all_recv, const int *cnts, int stride, int self, int dest, int my_first,
                    std::vector<PushRunHost> &out)
{
    out.clear();
    const int *rr = all_recv + (size_t)dest * (size_t)stride;
    for (int i = 0; i < cnts[dest]; ++i)
        if (rr[4 * i + 2] == self) out.push_back(PushRunHost{rr[4 * i] - my_first, rr[4 * i + 1], rr[4 * i + 3]});
}

} // namespace bicg

extern "C" int bicg_plan_push_runs(const int *all_recv, const int *cnts, int stride, int self, int dest, int my_first,
                                   int *out, int out_cap)
{
    std::vector<bicg::PushRunHost> v;
    bicg::plan_push_runs(all_recv, cnts, stride, self, dest, my_first, v);
    if ((int)v.size() > out_cap) return -(int)v.size();
    for (size_t i = 0; i < v.size(); ++i) { out[3 * i] = v[i].src; out[3 * i + 1] = v[i].len; out[3 * i + 2] = v[i].dst_off; }
    return (int)v.size();
}

extern "C" long long bicg_plan_merge(const CSR_Matrix *diag, const CSR_Matrix *offd, const INFO_Matrix *info, int self,
                                     int world, int gap, int ghost_off, unsigned *ptr_out, unsigned *col_out,
                                     double *val_out, int *recv_out, int recv_cap, int *n_ghost_out)
{
    std::vector<unsigned> mptr, mcol;
    std::vector<double> mval;
    std::vector<int> recv;
    int n_ghost = 0;
    bicg::merge_blocks(diag, offd, info, self, world, gap, ghost_off, mptr, mcol, mval, recv, n_ghost);
    if ((int)recv.size() > recv_cap) return -(long long)recv.size();
    std::memcpy(ptr_out, mptr.data(), mptr.size() * sizeof(unsigned));
    std::memcpy(col_out, mcol.data(), mcol.size() * sizeof(unsigned));
    std::memcpy(val_out, mval.data(), mval.size() * sizeof(double));
    std::memcpy(recv_out, recv.data(), recv.size() * sizeof(int));
    *n_ghost_out = n_ghost;
    return (long long)(recv.size() / 4);
}

extern "C" int bicg_plan_cta_tiles(const unsigned int *ptr, int rows, int ctas, int rows_per_tile, const unsigned char *row_extra,
                                   int extra_weight, int *tile_row, int tile_row_cap, int *cta_tile, unsigned int *max_tile_nnz)
{
    std::vector<int> tr, ct;
    const unsigned mx = bicg::plan_cta_tiles(ptr, rows, ctas, rows_per_tile, row_extra, extra_weight, tr, ct, 0, nullptr, nullptr, 216);
    if ((int)tr.size() > tile_row_cap) return -(int)tr.size();
    std::memcpy(tile_row, tr.data(), tr.size() * sizeof(int));
    std::memcpy(cta_tile, ct.data(), ct.size() * sizeof(int));
    if (max_tile_nnz) *max_tile_nnz = mx;
    return (int)tr.size() - 1;
}

// cap-limited variant: tiles of <= cap_limit entries; rows longer than that become chunk tiles (tile_flag 1 / 2).  tile_nz[t] is
// the first entry of tile t.  All three arrays need room for tile_cap ints; returns ntiles or -needed.
extern "C" int bicg_plan_cta_tiles_capped(const unsigned int *ptr, int rows, int ctas, int rows_per_tile, int cap_limit, int *tile_row,
                                          unsigned int *tile_nz, int *tile_flag, int tile_cap, int *cta_tile, unsigned int *max_tile_nnz)
{
    std::vector<int> tr, ct, fl;
    std::vector<unsigned> nz;
    const unsigned mx = bicg::plan_cta_tiles(ptr, rows, ctas, rows_per_tile, nullptr, 0, tr, ct, cap_limit, &nz, &fl, 216);
    if ((int)tr.size() > tile_cap) return -(int)tr.size();
    std::memcpy(tile_row, tr.data(), tr.size() * sizeof(int));
    std::memcpy(tile_nz, nz.data(), nz.size() * sizeof(unsigned));
    std::memcpy(tile_flag, fl.data(), fl.size() * sizeof(int));
    std::memcpy(cta_tile, ct.data(), ct.size() * sizeof(int));
    if (max_tile_nnz) *max_tile_nnz = mx;
    return (int)tr.size() - 1;
}

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
