// gen.cpp -- deterministic synthetic CSR inputs for the BASELINE.json configs.
//
// Every entry's value is a pure function of (seed, global row, slot) through splitmix64, so any rank can
// generate exactly its own row block without touching the others' -- the 16 M-row / 512 M-nnz config is
// produced 2 M rows at a time, one block per GPU.  Output is the reference's block layout:
// diag block with local column indices, offd block with global column indices, both in ascending
// column order inside a row (what a column-major SuiteSparse file gives after the reference's stable
// row sort, matrix.c:135-183, 380-392), and the partition of matrix.c:295-308 in info.
#include "bicgstab_b200.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
inline double u01(uint64_t seed, uint64_t row, uint64_t slot)   // [0,1)
{
    uint64_t h = splitmix64(seed ^ splitmix64(row * 64ull + slot));
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

struct Entry { long long col; double val; };

// One global row of the selected matrix family, columns ascending.
struct RowGen {
    int kind; long long g; double p0; uint64_t seed; long long n;

    void row(long long i, std::vector<Entry> &out) const
    {
        out.clear();
        switch (kind) {
        case 0: stencil15(i, out); break;
        case 1: laplace5(i, out); break;
        case 2: random_row(i, out); break;
        default: convdiff(i, out); break;
        }
    }

    // 15-point stencil: centre, 6 faces, 8 corners of a g^3 grid; diag = p0, off = -(0.5 + u)
    void stencil15(long long i, std::vector<Entry> &out) const
    {
        long long z = i / (g * g), y = (i / g) % g, x = i % g;
        int slot = 0;
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    int nzero = (dx != 0) + (dy != 0) + (dz != 0);
                    if (nzero == 2) continue;             // edges are not part of the stencil
                    ++slot;
                    long long xx = x + dx, yy = y + dy, zz = z + dz;
                    if (xx < 0 || xx >= g || yy < 0 || yy >= g || zz < 0 || zz >= g) continue;
                    long long c = (zz * g + yy) * g + xx;
                    double v = (nzero == 0) ? p0 : -(0.5 + u01(seed, (uint64_t)i, (uint64_t)slot));
                    out.push_back({c, v});
                }
    }

    void laplace5(long long i, std::vector<Entry> &out) const
    {
        long long y = i / g, x = i % g;
        if (y > 0) out.push_back({i - g, -1.0});
        if (x > 0) out.push_back({i - 1, -1.0});
        out.push_back({i, 4.0});
        if (x < g - 1) out.push_back({i + 1, -1.0});
        if (y < g - 1) out.push_back({i + g, -1.0});
    }

    // k entries per row: the diagonal (k + 1) and k - 1 distinct uniformly random columns, values -(0,1]
    void random_row(long long i, std::vector<Entry> &out) const
    {
        int k = (int)p0;
        std::vector<long long> cols;
        cols.reserve((size_t)k);
        cols.push_back(i);
        uint64_t ctr = 0;
        while ((int)cols.size() < k && (long long)cols.size() < n) {
            uint64_t h = splitmix64(seed ^ splitmix64((uint64_t)i * 0x100000001B3ull + ctr++));
            long long c = (long long)(h % (uint64_t)n);
            if (std::find(cols.begin(), cols.end(), c) == cols.end()) cols.push_back(c);
        }
        std::sort(cols.begin(), cols.end());
        for (size_t s = 0; s < cols.size(); ++s) {
            long long c = cols[s];
            double v = (c == i) ? (double)(k + 1) : -(1.0 - u01(seed, (uint64_t)i, (uint64_t)(s + 1)));
            out.push_back({c, v});
        }
    }

    // first-order upwind convection-diffusion on a g x g grid: nonsymmetric, not diagonally dominant by much
    void convdiff(long long i, std::vector<Entry> &out) const
    {
        long long y = i / g, x = i % g;
        double cx = p0, cy = 0.5 * p0;             // convection strengths in x and y
        if (y > 0) out.push_back({i - g, -1.0 - cy});
        if (x > 0) out.push_back({i - 1, -1.0 - cx});
        out.push_back({i, 4.0 + cx + cy + 1e-3 * u01(seed, (uint64_t)i, 0)});
        if (x < g - 1) out.push_back({i + 1, -1.0});
        if (y < g - 1) out.push_back({i + g, -1.0});
    }
};

} // namespace

extern "C" int bicg_gen_block(int kind, long long g, double p0, uint64_t seed, int rank, int world,
                              CSR_Matrix *diag, CSR_Matrix *offd, INFO_Matrix *info)
{
    if (world < 1 || rank < 0 || rank >= world || g < 1) return -1;
    long long n;
    switch (kind) {
    case 0: n = g * g * g; break;
    case 1: case 3: n = g * g; break;
    case 2: n = g; break;
    default: return -1;
    }
    if (n > 0x7fffffffLL) return -2;
    RowGen gen{kind, g, p0, seed, n};

    std::vector<Entry> row;
    const char *part = getenv("BICG_PARTITION");
    if (part && !strcmp(part, "nnz") && world > 1 && kind != 2) {        // kind 2: every row has the same length
        std::vector<unsigned> row_nnz((size_t)n);
        for (long long i = 0; i < n; ++i) { gen.row(i, row); row_nnz[(size_t)i] = (unsigned)row.size(); }
        bicg_plan_partition_nnz(row_nnz.data(), (int)n, world, info->recvcounts, info->displs);     // archive/matrix.c:407-420
    } else {
        bicg_plan_partition((int)n, world, info->recvcounts, info->displs);                        // matrix.c:295-308
    }
    const long long lo = info->displs[rank], nloc = info->recvcounts[rank], hi = lo + nloc;

    // pass 1: count
    unsigned long long nd = 0, no = 0, ntot_est = 0;
    for (long long i = lo; i < hi; ++i) {
        gen.row(i, row);
        for (const Entry &e : row) (e.col >= lo && e.col < hi) ? ++nd : ++no;
    }
    if (nd > 0xfffffff0ull || no > 0xfffffff0ull) return -2;
    (void)ntot_est;

    csr_init_matrix(diag); csr_init_matrix(offd);
    diag->rows = (unsigned)nloc; diag->cols = (unsigned)nloc; diag->nz = (unsigned)nd;   // matrix.c:343-345
    offd->rows = (unsigned)nloc; offd->cols = (unsigned)n;    offd->nz = (unsigned)no;   // matrix.c:350-352
    diag->val = (double *)malloc((nd + 1) * sizeof(double));
    diag->col = (unsigned *)malloc((nd + 1) * sizeof(unsigned));
    diag->ptr = (unsigned *)malloc(((size_t)nloc + 1) * sizeof(unsigned));
    offd->val = (double *)malloc((no + 1) * sizeof(double));
    offd->col = (unsigned *)malloc((no + 1) * sizeof(unsigned));
    offd->ptr = (unsigned *)malloc(((size_t)nloc + 1) * sizeof(unsigned));
    if (!diag->val || !diag->col || !diag->ptr || !offd->val || !offd->col || !offd->ptr) return -3;

    // pass 2: fill
    size_t kd = 0, ko = 0;
    diag->ptr[0] = offd->ptr[0] = 0;
    for (long long i = lo; i < hi; ++i) {
        gen.row(i, row);
        for (const Entry &e : row) {
            if (e.col >= lo && e.col < hi) { diag->val[kd] = e.val; diag->col[kd] = (unsigned)(e.col - lo); ++kd; }
            else                           { offd->val[ko] = e.val; offd->col[ko] = (unsigned)e.col;        ++ko; }
        }
        diag->ptr[i - lo + 1] = (unsigned)kd;
        offd->ptr[i - lo + 1] = (unsigned)ko;
    }

    // global info: nz of the whole matrix needs every row; it is cheap for the structured kinds and is
    // known in closed form for kind 2
    unsigned long long gnz = 0;
    if (kind == 2) {
        gnz = (unsigned long long)n * (unsigned long long)std::min<long long>((long long)p0, n);
    } else if (world == 1) {
        gnz = nd + no;
    } else {
        for (long long i = 0; i < n; ++i) { gen.row(i, row); gnz += row.size(); }
    }
    info->rows = info->cols = (unsigned)n;
    info->nz = (unsigned)std::min<unsigned long long>(gnz, 0xffffffffull);
    memcpy(info->code, "MCRG", 4);            // matrix / coordinate / real / general (mmio.h:31-44)
    return 0;
}
