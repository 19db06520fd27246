/*
 * mini_mpi.c -- fork + shared-memory implementation of the MPI subset in mpi_stub/mpi.h.
 * TEST INFRASTRUCTURE (oracle/): lets the reference sources under /root/reference/src run with
 * P > 1 ranks on the host cores of a box that has no MPI (SURVEY.md section 8(c) caveat 3).
 *
 * MINI_MPI_NP=P   ranks to create in MPI_Init (default 1)
 * MINI_MPI_SHM_MB size of the shared scratch mapping in MiB (default 1024, lazily committed)
 * MINI_MPI_PIN=1  pin rank r to core r (sched_setaffinity)
 */
#define _GNU_SOURCE
#include "mpi.h"
#include <sched.h>
#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

typedef struct {
    atomic_int arrived;
    atomic_int sense;
    char pad[56];
} shm_head;

static int g_np = 1, g_rank = 0;
static shm_head *g_head = NULL;
static char *g_data = NULL;
static size_t g_data_bytes = 0;
static int g_local_sense = 0;
static pid_t *g_children = NULL;

static size_t dt_size(MPI_Datatype dt)
{
    switch (dt) { case MPI_DOUBLE: return 8; case MPI_CHAR: return 1; case MPI_INT: return 4; }
    fprintf(stderr, "mini_mpi: unsupported datatype %d\n", dt);
    abort();
}

static void shm_barrier(void)
{
    if (g_np == 1) return;
    g_local_sense = !g_local_sense;
    if (atomic_fetch_add(&g_head->arrived, 1) == g_np - 1) {
        atomic_store(&g_head->arrived, 0);
        atomic_store(&g_head->sense, g_local_sense);
    } else {
        unsigned spins = 0;
        while (atomic_load(&g_head->sense) != g_local_sense)
            if ((++spins & 0x3ff) == 0) sched_yield();
    }
}

static void need(size_t bytes)
{
    if (bytes > g_data_bytes) {
        fprintf(stderr, "mini_mpi: collective needs %zu bytes of shared scratch, have %zu "
                        "(raise MINI_MPI_SHM_MB)\n", bytes, g_data_bytes);
        abort();
    }
}

int MPI_Init(int *argc, char ***argv)
{
    (void)argc; (void)argv;
    const char *e = getenv("MINI_MPI_NP");
    g_np = e ? atoi(e) : 1;
    if (g_np < 1) g_np = 1;
    if (g_np == 1) return MPI_SUCCESS;

    const char *m = getenv("MINI_MPI_SHM_MB");
    size_t mb = m ? (size_t)atol(m) : 1024;
    g_data_bytes = mb << 20;
    void *map = mmap(NULL, sizeof(shm_head) + g_data_bytes, PROT_READ | PROT_WRITE,
                     MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (map == MAP_FAILED) { perror("mini_mpi: mmap"); exit(EXIT_FAILURE); }
    g_head = (shm_head *)map;
    g_data = (char *)map + sizeof(shm_head);
    atomic_init(&g_head->arrived, 0);
    atomic_init(&g_head->sense, 0);

    g_children = (pid_t *)calloc((size_t)g_np, sizeof(pid_t));
    fflush(NULL);
    for (int r = 1; r < g_np; ++r) {
        pid_t pid = fork();
        if (pid < 0) { perror("mini_mpi: fork"); exit(EXIT_FAILURE); }
        if (pid == 0) { g_rank = r; free(g_children); g_children = NULL; break; }
        g_children[r] = pid;
    }
    const char *pin = getenv("MINI_MPI_PIN");
    if (pin && atoi(pin)) {
        cpu_set_t set; CPU_ZERO(&set); CPU_SET(g_rank, &set);
        sched_setaffinity(0, sizeof(set), &set);
    }
    shm_barrier();
    return MPI_SUCCESS;
}

int MPI_Finalize(void)
{
    fflush(NULL);
    shm_barrier();
    if (g_np > 1 && g_rank == 0 && g_children) {
        for (int r = 1; r < g_np; ++r) { int st; waitpid(g_children[r], &st, 0); }
        free(g_children); g_children = NULL;
    }
    return MPI_SUCCESS;
}

int MPI_Comm_size(MPI_Comm c, int *size) { (void)c; *size = g_np; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm c, int *rank) { (void)c; *rank = g_rank; return MPI_SUCCESS; }

int MPI_Get_processor_name(char *name, int *len)
{
    if (gethostname(name, MPI_MAX_PROCESSOR_NAME - 1) != 0) strcpy(name, "localhost");
    name[MPI_MAX_PROCESSOR_NAME - 1] = '\0';
    *len = (int)strlen(name);
    return MPI_SUCCESS;
}

double MPI_Wtime(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int MPI_Barrier(MPI_Comm c) { (void)c; shm_barrier(); return MPI_SUCCESS; }

int MPI_Gather(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, int rcount, MPI_Datatype rt,
               int root, MPI_Comm comm)
{
    (void)comm; (void)rcount; (void)rt;
    size_t bytes = (size_t)scount * dt_size(st);
    if (g_np == 1) { memcpy(rbuf, sbuf, bytes); return MPI_SUCCESS; }
    need(bytes * (size_t)g_np);
    memcpy(g_data + bytes * (size_t)g_rank, sbuf, bytes);
    shm_barrier();
    if (g_rank == root) memcpy(rbuf, g_data, bytes * (size_t)g_np);
    shm_barrier();
    return MPI_SUCCESS;
}

int MPI_Iallgatherv(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, const int *rcounts,
                    const int *displs, MPI_Datatype rt, MPI_Comm comm, MPI_Request *req)
{
    (void)comm; (void)rt;
    size_t es = dt_size(st);
    if (req) *req = 0;
    if (g_np == 1) {
        memcpy((char *)rbuf + es * (size_t)displs[0], sbuf, es * (size_t)scount);
        return MPI_SUCCESS;
    }
    size_t total = 0;
    for (int p = 0; p < g_np; ++p) {
        size_t end = (size_t)displs[p] + (size_t)rcounts[p];
        if (end > total) total = end;
    }
    need(total * es);
    memcpy(g_data + es * (size_t)displs[g_rank], sbuf, es * (size_t)scount);
    shm_barrier();
    for (int p = 0; p < g_np; ++p)
        memcpy((char *)rbuf + es * (size_t)displs[p], g_data + es * (size_t)displs[p], es * (size_t)rcounts[p]);
    shm_barrier();
    return MPI_SUCCESS;
}

int MPI_Allreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm)
{
    (void)comm;
    if (dt != MPI_DOUBLE || op != MPI_SUM) { fprintf(stderr, "mini_mpi: only SUM of doubles\n"); abort(); }
    const double *src = (sbuf == MPI_IN_PLACE) ? (const double *)rbuf : (const double *)sbuf;
    if (g_np == 1) {
        if (sbuf != MPI_IN_PLACE) memcpy(rbuf, sbuf, 8u * (size_t)count);
        return MPI_SUCCESS;
    }
    need(8u * (size_t)count * (size_t)g_np);
    double *slots = (double *)g_data;
    memcpy(slots + (size_t)g_rank * (size_t)count, src, 8u * (size_t)count);
    shm_barrier();
    double *out = (double *)rbuf;
    for (int i = 0; i < count; ++i) {
        double acc = slots[i];
        for (int p = 1; p < g_np; ++p) acc += slots[(size_t)p * (size_t)count + (size_t)i];
        out[i] = acc;   /* rank order, identical on every rank */
    }
    shm_barrier();
    return MPI_SUCCESS;
}

int MPI_Iallreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm,
                   MPI_Request *req)
{
    if (req) *req = 0;
    return MPI_Allreduce(sbuf, rbuf, count, dt, op, comm);
}

int MPI_Wait(MPI_Request *req, MPI_Status *st) { (void)req; (void)st; return MPI_SUCCESS; }

static int dead_call(const char *what)
{
    fprintf(stderr, "mini_mpi: %s is not implemented (only reachable from dead reference code)\n", what);
    abort();
    return 1;
}
int MPI_Isend(const void *b, int c, MPI_Datatype d, int dest, int tag, MPI_Comm comm, MPI_Request *r)
{ (void)b; (void)c; (void)d; (void)dest; (void)tag; (void)comm; (void)r; return dead_call("MPI_Isend"); }
int MPI_Irecv(void *b, int c, MPI_Datatype d, int src, int tag, MPI_Comm comm, MPI_Request *r)
{ (void)b; (void)c; (void)d; (void)src; (void)tag; (void)comm; (void)r; return dead_call("MPI_Irecv"); }
int MPI_Waitsome(int n, MPI_Request *r, int *oc, int *idx, MPI_Status *s)
{ (void)n; (void)r; (void)oc; (void)idx; (void)s; return dead_call("MPI_Waitsome"); }
