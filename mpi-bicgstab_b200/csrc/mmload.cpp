// mmload.cpp -- Matrix-Market coordinate file -> one rank's diag / offd CSR blocks.
//
// Replaces the reference's loader chain MPI_csr_load_matrix_block -> MPI_coo_load_matrix_block -> coo2csr
// (matrix.c:268-419, 206-232) behind the same entry point.  Same result, different method: the file is
// mapped once and parsed in a single pass with from_chars (the reference runs fscanf over the whole file
// twice per rank, matrix.c:315-393), entries of this rank's rows are bucketed by a stable counting sort
// on the row (the reference: stable merge sort on the row, matrix.c:135-183), so the in-row order is the
// file order in both, and the diag / offd split and column conventions are those of matrix.c:380-392.
//
// Behaviour kept from the reference: 1-based -> 0-based indices; partition matrix.c:295-308; `symmetric`
// files are NOT mirrored (the block loader ignores the flag, matrix.c:93 vs 268-399) unless
// BICG_MM_EXPAND_SYMMETRIC=1.  Behaviour fixed: `pattern` entries get 1.0 and `integer` entries their
// value (the reference's block loader leaves val unset for both, matrix.c:316-331, 384, 389; its serial
// loader does what we do here, matrix.c:74-91).
#include "bicgstab_b200.h"

#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>

extern "C" void csr_init_matrix(CSR_Matrix *m)          // matrix.c:188-193
{
    m->val = nullptr; m->col = nullptr; m->ptr = nullptr;
    m->nz = m->rows = m->cols = 0;
}

extern "C" void csr_free_matrix(CSR_Matrix *m)          // matrix.c:195-204
{
    bicg_matrix_invalidate(m);                          // forget any cached device copy keyed by these arrays
    free(m->val); free(m->col); free(m->ptr);
    csr_init_matrix(m);
}

// matrix.c:536-551: add sigma to every diagonal entry of the local diagonal block, in place, on the caller's HOST arrays
// (A + sigma I for the caller who solves the shifted systems one by one; main_shifted.c:132 keeps the call commented out).
// A row without a stored diagonal entry is an error, as in the reference.  The cached device copy keyed by these arrays is
// dropped explicitly (the content fingerprint would notice the change too).
extern "C" void csr_shift_diagonal(CSR_Matrix *A_diag, double sigma)
{
    for (unsigned i = 0; i < A_diag->rows; ++i) {
        bool found = false;
        for (unsigned j = A_diag->ptr[i]; j < A_diag->ptr[i + 1] && !found; ++j)
            if (A_diag->col[j] == i) { A_diag->val[j] += sigma; found = true; }
        if (!found) {
            fprintf(stderr, "Error: Diagonal element not found in row %u.\n", i);
            exit(EXIT_FAILURE);
        }
    }
    bicg_matrix_invalidate(A_diag);
}

namespace {

[[noreturn]] void fail(const char *msg)
{
    fprintf(stderr, "%s\n", msg);                       // matrix.c:281-288 convention
    exit(EXIT_FAILURE);
}

struct Cursor {
    const char *p, *end;
    void skip_ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) ++p; }
    void skip_line() { while (p < end && *p != '\n') ++p; if (p < end) ++p; }
    bool at_end() const { return p >= end; }
    template <class T> bool number(T &out)
    {
        skip_ws();
        if (p < end && *p == '+') ++p;
        auto res = std::from_chars(p, end, out);
        if (res.ec != std::errc()) return false;
        p = res.ptr;
        return true;
    }
};

std::string lower_token(Cursor &c)
{
    c.skip_ws();
    std::string t;
    while (c.p < c.end && *c.p != ' ' && *c.p != '\t' && *c.p != '\n' && *c.p != '\r') {
        char ch = *c.p++;
        t.push_back((char)((ch >= 'A' && ch <= 'Z') ? ch - 'A' + 'a' : ch));
    }
    return t;
}

} // namespace

extern "C" void MPI_csr_load_matrix_block(char *filename, CSR_Matrix *D, CSR_Matrix *O, INFO_Matrix *info)
{
    const int world = bicg_comm_world(), me = bicg_comm_rank();

    int fd = open(filename, O_RDONLY);
    if (fd < 0) { fprintf(stderr, "ERROR: can't open file \"%s\"\n", filename); exit(EXIT_FAILURE); }
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size == 0) fail("ERROR: Could not process Matrix Market banner.");
    const char *base = (const char *)mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (base == MAP_FAILED) fail("ERROR: Could not process Matrix Market banner.");
    Cursor c{base, base + sb.st_size};

    // banner: %%MatrixMarket matrix coordinate <field> <symmetry>     (mmio.c:96-186)
    if (lower_token(c) != "%%matrixmarket") fail("ERROR: Could not process Matrix Market banner.");
    std::string obj = lower_token(c), fmt = lower_token(c), field = lower_token(c), symm = lower_token(c);
    if (obj != "matrix" || fmt != "coordinate") fail("ERROR: Could not process Matrix Market banner.");
    char code[4] = {'M', 'C', ' ', ' '};
    if (field == "real") code[2] = 'R'; else if (field == "integer") code[2] = 'I';
    else if (field == "pattern") code[2] = 'P'; else if (field == "complex") code[2] = 'C';
    else fail("ERROR: Could not process Matrix Market banner.");
    if (symm == "general") code[3] = 'G'; else if (symm == "symmetric") code[3] = 'S';
    else if (symm == "skew-symmetric") code[3] = 'K'; else if (symm == "hermitian") code[3] = 'H';
    else fail("ERROR: Could not process Matrix Market banner.");
    if (code[2] == 'C') fail("ERROR: reading matrix data.");      // complex entries cannot be scanned as %lg
    c.skip_line();

    // comments and blank lines, then the size line                      (mmio.c:189-223)
    for (;;) {
        c.skip_ws();
        if (c.at_end()) fail("ERROR: Could not read matrix size.");
        if (*c.p == '%') { c.skip_line(); continue; }
        break;
    }
    long long m = 0, n = 0, nz = 0;
    if (!c.number(m) || !c.number(n) || !c.number(nz)) fail("ERROR: Could not read matrix size.");

    info->nz = (unsigned)nz; info->rows = (unsigned)m; info->cols = (unsigned)n;
    memcpy(info->code, code, 4);
    const bool pattern = code[2] == 'P', integer = code[2] == 'I';
    const char *ex = getenv("BICG_MM_EXPAND_SYMMETRIC");
    const bool expand = ex && atoi(ex) && code[3] == 'S';

    // Partition: the reference's equal-rows rule (matrix.c:295-308), or -- BICG_PARTITION=nnz -- its archived nnz-balanced
    // rule (archive/matrix.c:407-420), which needs the row lengths of the whole file first: one extra counting pass
    // over the mapped text (still far cheaper than the reference's two fscanf passes).
    const char *part = getenv("BICG_PARTITION");
    if (part && !strcmp(part, "nnz") && world > 1) {
        std::vector<unsigned> row_nnz((size_t)m, 0u);
        Cursor c2 = c;
        for (long long e = 0; e < nz; ++e) {
            long long r, cc;
            if (!c2.number(r) || !c2.number(cc)) fail("ERROR: reading matrix data.");
            if (!pattern) { double v; long long iv; if (integer ? !c2.number(iv) : !c2.number(v)) fail("ERROR: reading matrix data."); }
            if (r < 1 || r > m || cc < 1 || cc > n) fail("ERROR: reading matrix data.");
            ++row_nnz[(size_t)(r - 1)];
            if (expand && r != cc) ++row_nnz[(size_t)(cc - 1)];
        }
        bicg_plan_partition_nnz(row_nnz.data(), (int)m, world, info->recvcounts, info->displs);
    } else {
        bicg_plan_partition((int)m, world, info->recvcounts, info->displs);
    }
    const long long lo = info->displs[me], nloc = info->recvcounts[me], hi = lo + nloc;

    struct Ent { unsigned row; unsigned col; double val; };
    std::vector<Ent> dent, oent;
    auto keep = [&](long long r, long long cc, double v) {
        if (r < lo || r >= hi) return;
        if (cc >= lo && cc < hi) dent.push_back({(unsigned)(r - lo), (unsigned)(cc - lo), v});   // matrix.c:381-385
        else                      oent.push_back({(unsigned)(r - lo), (unsigned)cc, v});          // matrix.c:386-390
    };
    for (long long e = 0; e < nz; ++e) {
        long long r, cc; double v = 1.0;
        if (!c.number(r) || !c.number(cc)) fail("ERROR: reading matrix data.");
        if (!pattern) {
            if (integer) { long long iv; if (!c.number(iv)) fail("ERROR: reading matrix data."); v = (double)iv; }
            else if (!c.number(v)) fail("ERROR: reading matrix data.");
        }
        --r; --cc;                                                   // matrix.c:333-334
        keep(r, cc, v);
        if (expand && r != cc) keep(cc, r, v);
    }
    munmap((void *)base, (size_t)sb.st_size);
    close(fd);

    auto to_csr = [&](std::vector<Ent> &ent, CSR_Matrix *out, unsigned cols) {
        csr_init_matrix(out);
        out->rows = (unsigned)nloc; out->cols = cols; out->nz = (unsigned)ent.size();
        out->val = (double *)malloc((ent.size() + 1) * sizeof(double));
        out->col = (unsigned *)malloc((ent.size() + 1) * sizeof(unsigned));
        out->ptr = (unsigned *)calloc((size_t)nloc + 1, sizeof(unsigned));
        for (const Ent &e : ent) ++out->ptr[e.row + 1];
        for (long long i = 0; i < nloc; ++i) out->ptr[i + 1] += out->ptr[i];
        std::vector<unsigned> fill(out->ptr, out->ptr + nloc);
        for (const Ent &e : ent) {                                   // stable: file order inside a row
            unsigned k = fill[e.row]++;
            out->val[k] = e.val; out->col[k] = e.col;
        }
        std::vector<Ent>().swap(ent);
    };
    to_csr(dent, D, (unsigned)nloc);        // matrix.c:343-345
    to_csr(oent, O, (unsigned)n);           // matrix.c:350-352
}
