// shm_boot.cpp -- host-side rendezvous for multi-process jobs started without MPI or torch: a POSIX
// shared-memory segment with a sense-reversing barrier and an allgather, used (a) as the bicg_allgather_fn of
// bicg_comm_init and (b) to implement the handful of MPI calls main.c makes (include/compat/mpi.h).
// It only carries bootstrap data (IPC handles, halo plans, processor names); solver traffic never touches it.
#include "bicgstab_b200.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <sched.h>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

namespace {

struct ShmHead {
    std::atomic<int> magic;
    std::atomic<int> arrived;
    std::atomic<int> sense;
    char pad[52];
};
constexpr int SHM_READY = 0x42494347;     // "BICG"

struct Boot {
    int rank = 0, world = 1;
    ShmHead *head = nullptr;
    char *data = nullptr;
    size_t data_bytes = 0;
    int local_sense = 0;
    std::string name;
} g;

int env_int(const char *a, const char *b, int dflt)
{
    const char *v = getenv(a);
    if (!v && b) v = getenv(b);
    return v ? atoi(v) : dflt;
}

void barrier()
{
    if (g.world == 1) return;
    g.local_sense = !g.local_sense;
    if (g.head->arrived.fetch_add(1) == g.world - 1) {
        g.head->arrived.store(0);
        g.head->sense.store(g.local_sense);
    } else {
        unsigned spins = 0;
        while (g.head->sense.load() != g.local_sense)
            if ((++spins & 0xff) == 0) sched_yield();
    }
}

int shm_allgather(void *, const void *send, void *recv, size_t bytes)
{
    if (g.world == 1) { memcpy(recv, send, bytes); return 0; }
    if (bytes * (size_t)g.world > g.data_bytes) {
        fprintf(stderr, "bicgstab_b200: bootstrap allgather of %zu bytes/rank exceeds the shm segment (%zu); "
                        "raise BICG_SHM_MB\n", bytes, g.data_bytes);
        return -1;
    }
    memcpy(g.data + bytes * (size_t)g.rank, send, bytes);
    barrier();
    memcpy(recv, g.data, bytes * (size_t)g.world);
    barrier();
    return 0;
}

} // namespace

extern "C" {

// Join the job described by the environment.  Exposed so non-MPI C programs can bootstrap too.
int bicg_shm_bootstrap(void)
{
    g.rank = env_int("BICG_RANK", "RANK", 0);
    g.world = env_int("BICG_WORLD", "WORLD_SIZE", 1);
    if (g.world <= 1) { g.world = 1; g.rank = 0; return bicg_comm_init(0, 1, nullptr, nullptr); }
    const char *job = getenv("BICG_JOB_ID");
    const char *port = getenv("MASTER_PORT");
    g.name = std::string("/bicg_b200_") + (job ? job : (port ? port : "default"));
    const size_t mb = (size_t)env_int("BICG_SHM_MB", nullptr, 256);
    g.data_bytes = mb << 20;
    const size_t total = sizeof(ShmHead) + g.data_bytes;
    int fd = -1;
    if (g.rank == 0) {
        shm_unlink(g.name.c_str());                       // leftovers of a crashed job
        fd = shm_open(g.name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)total) != 0) { perror("bicgstab_b200: shm_open"); return -1; }
    } else {
        for (int tries = 0; tries < 60000; ++tries) {     // up to ~60 s for rank 0 to appear
            fd = shm_open(g.name.c_str(), O_RDWR, 0600);
            struct stat sb;
            if (fd >= 0 && fstat(fd, &sb) == 0 && (size_t)sb.st_size >= total) break;
            if (fd >= 0) { close(fd); fd = -1; }
            usleep(1000);
        }
        if (fd < 0) { fprintf(stderr, "bicgstab_b200: rank %d could not join %s\n", g.rank, g.name.c_str()); return -1; }
    }
    void *map = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (map == MAP_FAILED) { perror("bicgstab_b200: mmap"); return -1; }
    g.head = (ShmHead *)map;
    g.data = (char *)map + sizeof(ShmHead);
    if (g.rank == 0) g.head->magic.store(SHM_READY);
    else while (g.head->magic.load() != SHM_READY) usleep(100);
    barrier();
    return bicg_comm_init(g.rank, g.world, shm_allgather, nullptr);
}

void bicg_shm_shutdown(void)
{
    bicg_comm_finalize();
    if (g.world > 1 && g.head) {
        barrier();
        if (g.rank == 0) { g.head->magic.store(0); shm_unlink(g.name.c_str()); }
    }
}

// ---- include/compat/mpi.h -------------------------------------------------------------------------------
int bicg_shim_MPI_Init(int *, char ***) { return bicg_shm_bootstrap(); }
int bicg_shim_MPI_Finalize(void) { fflush(nullptr); bicg_shm_shutdown(); return 0; }
int bicg_shim_MPI_Comm_size(int, int *size) { *size = bicg_comm_world(); return 0; }
int bicg_shim_MPI_Comm_rank(int, int *rank) { *rank = bicg_comm_rank(); return 0; }
int bicg_shim_MPI_Get_processor_name(char *name, int *len)
{
    if (gethostname(name, 127) != 0) strcpy(name, "localhost");
    name[127] = '\0';
    *len = (int)strlen(name);
    return 0;
}
double bicg_shim_MPI_Wtime(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
int bicg_shim_MPI_Barrier(int) { barrier(); return 0; }
int bicg_shim_MPI_Gather(const void *sbuf, int scount, int st, void *rbuf, int, int, int root, int)
{
    const size_t es = (st == 1) ? 8 : (st == 3) ? 4 : 1;      // MPI_DOUBLE / MPI_INT / MPI_CHAR of compat/mpi.h
    const size_t bytes = es * (size_t)scount;
    const int world = bicg_comm_world();
    if (world == 1) { memcpy(rbuf, sbuf, bytes); return 0; }
    char *tmp = (char *)malloc(bytes * (size_t)world);
    int rc = shm_allgather(nullptr, sbuf, tmp, bytes);
    if (rc == 0 && bicg_comm_rank() == root) memcpy(rbuf, tmp, bytes * (size_t)world);
    free(tmp);
    return rc;
}

} // extern "C"
