// hostvec.cpp -- vector.h's four entry points (vector.c:3-27) for the reference's DRIVERS: main_shifted.c:114-135,
// main_repeat.c:121 and main_seed_diff.c:118-121 build right-hand sides and keep copies of them on HOST arrays with
// my_daxpy / my_dcopy, so these symbols have to exist for those programs to link unchanged.  Host loops on host memory,
// like the reference's; the solvers never come here -- their vector work is fused into the device kernels
// (vec_body.cuh, mega.cu), and nothing in this file is a fallback for them.
// Arithmetic: one multiply and one add per element, no contraction (what the reference's strict build and the oracle do).
#include "bicgstab_b200.h"

#define BICG_NO_CONTRACT __attribute__((optimize("fp-contract=off")))

extern "C" {

BICG_NO_CONTRACT void my_daxpy(int n, double alpha, const double *x, double *y)        // vector.c:3-7
{
    for (int i = 0; i < n; ++i) y[i] += alpha * x[i];
}

BICG_NO_CONTRACT double my_ddot(int n, const double *x, const double *y)               // vector.c:9-15
{
    double sum = 0.0;
    for (int i = 0; i < n; ++i) sum += x[i] * y[i];
    return sum;
}

void my_dscal(int n, double alpha, double *x)                                          // vector.c:17-21
{
    for (int i = 0; i < n; ++i) x[i] *= alpha;
}

void my_dcopy(int n, const double *x, double *y)                                       // vector.c:23-27
{
    for (int i = 0; i < n; ++i) y[i] = x[i];
}

} // extern "C"
