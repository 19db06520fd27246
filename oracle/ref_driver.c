/*
 * ref_driver.c -- in-memory driver around the reference's own solver entry points.
 * TEST INFRASTRUCTURE (oracle/).  Linked against objects compiled from /root/reference/src
 * (oracle/Makefile) it plays the role of the reference's main.c (main.c:79-141) for inputs that
 * are too large for the reference's fscanf loader: it maps a binary CSR file, gives every rank its
 * row block split into the diagonal / off-diagonal parts exactly as the block loader would
 * (partition matrix.c:295-308, split + local/global column convention matrix.c:380-392, in-row
 * order preserved as by the stable row sort matrix.c:135-183), builds b and x0 as main.c:109-117
 * does, and calls bicgstab / ca_bicgstab / pipe_bicgstab / pipe_bicgstab_rr (solver.h:10-13).
 *
 * usage: ref_driver <csr.bin> <method> <rhs: a1|ones> <out_prefix|-> [krr nrr]
 *   csr.bin : int64 n, int64 nnz, uint32 ptr[n+1], uint32 col[nnz], double val[nnz]
 *   outputs : <out_prefix>.x.<rank>, <out_prefix>.r.<rank> (raw doubles), and from rank 0
 *             <out_prefix>.hist = int32 count, then count x (int32 k, double residual) packed,
 *             plus one JSON line on stdout.
 */
#define _GNU_SOURCE
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "solver.h"     /* the reference's own header: CSR_Matrix, INFO_Matrix, prototypes */
#include "ref_shim.h"

static void die(const char *msg) { fprintf(stderr, "ref_driver: %s\n", msg); exit(EXIT_FAILURE); }

static void dump(const char *prefix, const char *what, int rank, const double *v, int n)
{
    char path[1024];
    snprintf(path, sizeof path, "%s.%s.%d", prefix, what, rank);
    FILE *f = fopen(path, "wb");
    if (!f) die("cannot write output");
    fwrite(v, sizeof(double), (size_t)n, f);
    fclose(f);
}

int main(int argc, char **argv)
{
    MPI_Init(&argc, &argv);
    int np, me;
    MPI_Comm_size(MPI_COMM_WORLD, &np);
    MPI_Comm_rank(MPI_COMM_WORLD, &me);
    if (argc < 5) die("usage: ref_driver <csr.bin> <method> <a1|ones> <out_prefix|-> [krr nrr]");
    const char *method = argv[2], *rhs = argv[3], *prefix = argv[4];

    int fd = open(argv[1], O_RDONLY);
    if (fd < 0) die("cannot open csr file");
    struct stat sb; fstat(fd, &sb);
    const char *base = (const char *)mmap(NULL, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (base == MAP_FAILED) die("mmap failed");
    int64_t n = ((const int64_t *)base)[0], nnz = ((const int64_t *)base)[1];
    const uint32_t *gptr = (const uint32_t *)(base + 16);
    const uint32_t *gcol = gptr + (n + 1);
    const double *gval = (const double *)(base + 16 + 4 * (n + 1) + 4 * nnz + ((4 * (n + 1 + nnz)) % 8 ? 4 : 0));

    INFO_Matrix info;
    info.rows = info.cols = (unsigned)n; info.nz = (unsigned)nnz;
    memcpy(info.code, "MCRG", 4);
    info.recvcounts = (int *)malloc((size_t)np * sizeof(int));
    info.displs = (int *)malloc((size_t)np * sizeof(int));
    int per = (int)(n / np), extra = (int)(n % np);
    for (int p = 0; p < np; ++p) {
        info.displs[p] = p * per + (p < extra ? p : extra);
        info.recvcounts[p] = per + (p < extra ? 1 : 0);
    }
    unsigned lo = (unsigned)info.displs[me], nloc = (unsigned)info.recvcounts[me], hi = lo + nloc;

    CSR_Matrix D, O;
    csr_init_matrix(&D); csr_init_matrix(&O);
    size_t nd = 0, no = 0;
    for (unsigned i = lo; i < hi; ++i)
        for (uint32_t j = gptr[i]; j < gptr[i + 1]; ++j)
            if (gcol[j] >= lo && gcol[j] < hi) ++nd; else ++no;
    D.rows = nloc; D.cols = nloc; D.nz = (unsigned)nd;
    O.rows = nloc; O.cols = (unsigned)n; O.nz = (unsigned)no;
    D.val = (double *)malloc((nd + 1) * sizeof(double)); D.col = (unsigned *)malloc((nd + 1) * sizeof(unsigned));
    O.val = (double *)malloc((no + 1) * sizeof(double)); O.col = (unsigned *)malloc((no + 1) * sizeof(unsigned));
    D.ptr = (unsigned *)malloc(((size_t)nloc + 1) * sizeof(unsigned));
    O.ptr = (unsigned *)malloc(((size_t)nloc + 1) * sizeof(unsigned));
    nd = no = 0; D.ptr[0] = O.ptr[0] = 0;
    for (unsigned i = lo; i < hi; ++i) {
        for (uint32_t j = gptr[i]; j < gptr[i + 1]; ++j) {
            unsigned c = gcol[j];
            if (c >= lo && c < hi) { D.val[nd] = gval[j]; D.col[nd] = c - lo; ++nd; }
            else                   { O.val[no] = gval[j]; O.col[no] = c;      ++no; }
        }
        D.ptr[i - lo + 1] = (unsigned)nd; O.ptr[i - lo + 1] = (unsigned)no;
    }

    double *x_loc = (double *)malloc((size_t)nloc * sizeof(double));
    double *r_loc = (double *)malloc((size_t)nloc * sizeof(double));
    double *xfull = (double *)malloc((size_t)n * sizeof(double));
    if (strcmp(rhs, "a1") == 0) {                         /* main.c:109-113 */
        for (unsigned i = 0; i < nloc; ++i) x_loc[i] = 1.0;
        MPI_csr_spmv_ovlap(&D, &O, &info, x_loc, xfull, r_loc);
    } else {
        for (unsigned i = 0; i < nloc; ++i) r_loc[i] = 1.0;
    }
    for (unsigned i = 0; i < nloc; ++i) x_loc[i] = 0.0;   /* main.c:115-117 */

    orc_ref_hist_reset();
    int iters;
    if      (strcmp(method, "bicgstab") == 0)      iters = bicgstab(&D, &O, &info, x_loc, r_loc);
    else if (strcmp(method, "ca_bicgstab") == 0)   iters = ca_bicgstab(&D, &O, &info, x_loc, r_loc);
    else if (strcmp(method, "pipe_bicgstab") == 0) iters = pipe_bicgstab(&D, &O, &info, x_loc, r_loc);
    else if (strcmp(method, "pipe_bicgstab_rr") == 0) {
        if (argc < 7) die("pipe_bicgstab_rr needs <krr> <nrr>");
        iters = pipe_bicgstab_rr(&D, &O, &info, x_loc, r_loc, atoi(argv[5]), atoi(argv[6]));
    } else { die("unknown method"); return 1; }

    if (strcmp(prefix, "-") != 0) {
        dump(prefix, "x", me, x_loc, (int)nloc);
        dump(prefix, "r", me, r_loc, (int)nloc);
        if (me == 0) {
            char path[1024];
            snprintf(path, sizeof path, "%s.hist", prefix);
            FILE *f = fopen(path, "wb");
            int32_t cnt = orc_ref_hist_count();
            fwrite(&cnt, 4, 1, f);
            for (int i = 0; i < cnt; ++i) {
                int32_t k = orc_ref_hist_iter(i); double res = orc_ref_hist_res(i);
                fwrite(&k, 4, 1, f); fwrite(&res, 8, 1, f);
            }
            fclose(f);
        }
    }
    if (me == 0) {
        fprintf(stdout, "{\"ranks\": %d, \"method\": \"%s\", \"n\": %lld, \"nnz\": %lld, \"iters\": %d, "
                        "\"final_res\": %.17e, \"total_time_s\": %.9e, \"avg_time_per_iter_s\": %.9e}\n",
                np, method, (long long)n, (long long)nnz, iters, orc_ref_final_res(),
                orc_ref_total_time(), orc_ref_avg_time());
    }
    csr_free_matrix(&D); csr_free_matrix(&O);
    free(x_loc); free(r_loc); free(xfull); free(info.recvcounts); free(info.displs);
    MPI_Finalize();
    return 0;
}
