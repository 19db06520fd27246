// gen_main.cpp -- stand-alone writer of the benchmark's synthetic matrices for the CPU reference arm.
// TEST INFRASTRUCTURE (oracle/): `bench.py --impl reference` must time the reference's own code without loading the
// product library, so the matrix it solves is produced by this small executable (same generator source as the
// library's bicg_gen_block: mpi-bicgstab_b200/csrc/gen.cpp, compiled here as plain host C++) and handed to
// oracle/_ref/ref_driver_* as a binary CSR file (format of oracle.py: write_csr_bin).
//   gen_csr <kind 0..3> <g> <p0> <out.bin> [seed]
#include "bicgstab_b200.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

// the generator only needs the reference's trivial initialiser (matrix.c:188-194) from the rest of the library
extern "C" void csr_init_matrix(CSR_Matrix *m) { m->val = nullptr; m->col = nullptr; m->ptr = nullptr; m->nz = m->rows = m->cols = 0; }

// ... and the partition rule (matrix.c:295-308), restated here so that plan.cpp need not be linked
extern "C" void bicg_plan_partition(int n, int world, int *counts, int *displs)
{
    const int base = n / world, extra = n % world;
    for (int p = 0; p < world; ++p) { counts[p] = base + (p < extra ? 1 : 0); displs[p] = p * base + (p < extra ? p : extra); }
}

extern "C" void bicg_plan_partition_nnz(const unsigned int *, int n, int world, int *counts, int *displs)
{
    bicg_plan_partition(n, world, counts, displs);          // never reached: this tool always generates world = 1
}

int main(int argc, char **argv)
{
    if (argc < 5) { fprintf(stderr, "usage: %s kind g p0 out.bin [seed]\n", argv[0]); return 2; }
    const int kind = atoi(argv[1]);
    const long long g = atoll(argv[2]);
    const double p0 = atof(argv[3]);
    const uint64_t seed = argc > 5 ? strtoull(argv[5], nullptr, 10) : 12345ull;
    CSR_Matrix diag{}, offd{};
    INFO_Matrix info{};
    int rc[1], ds[1];
    info.recvcounts = rc; info.displs = ds;
    if (bicg_gen_block(kind, g, p0, seed, 0, 1, &diag, &offd, &info) != 0) { fprintf(stderr, "gen_csr: generator failed\n"); return 1; }
    FILE *f = fopen(argv[4], "wb");
    if (!f) { perror(argv[4]); return 1; }
    const long long hdr[2] = {(long long)diag.rows, (long long)diag.nz};
    fwrite(hdr, sizeof(long long), 2, f);
    fwrite(diag.ptr, sizeof(unsigned), (size_t)diag.rows + 1, f);
    fwrite(diag.col, sizeof(unsigned), diag.nz, f);
    if (((size_t)diag.rows + 1 + diag.nz) % 2) { const unsigned z = 0; fwrite(&z, sizeof(unsigned), 1, f); }
    fwrite(diag.val, sizeof(double), diag.nz, f);
    fclose(f);
    printf("%lld %lld\n", hdr[0], hdr[1]);
    return 0;
}
