/*
 * fake_nccl.c — TEST INFRASTRUCTURE.  The five NCCL entry points the library resolves at run time
 * (nhd_b200/csrc/nhd_api.cu: ncclGetUniqueId, ncclCommInitRank, ncclCommDestroy, ncclAllGather,
 * ncclGetErrorString) for ranks that are ordinary processes on one machine without GPUs: the
 * communicator is a POSIX shared-memory segment named by the unique id, the all-gather copies every
 * rank's buffer into its slot, meets at a barrier, copies all slots out, meets again.  Built as libnccl.so.2
 * (that soname is what the library looks for) and loaded by the test processes before the emulated
 * library; lets the world_size > 1 path of nhd_solve_batch (sharded filter + one all-gather + identical
 * sweeps) run in the CPU suite.  Only what that path uses: bytes.
 */
#define _GNU_SOURCE
#include <fcntl.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { FAKE_MAX_RANKS = 8, FAKE_SLOT_BYTES = 8 << 20 };

typedef struct {
    volatile int arrived, generation;
    volatile int attached;
    char pad[52];
    unsigned char slots[];                 /* n ranks x FAKE_SLOT_BYTES */
} shared_t;

typedef struct ncclComm { int rank, n; shared_t* sh; size_t bytes; char name[64]; } comm_t;

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static int barrier(comm_t* c)
{
    shared_t* s = c->sh;
    const int gen = __atomic_load_n(&s->generation, __ATOMIC_ACQUIRE);
    if (__atomic_add_fetch(&s->arrived, 1, __ATOMIC_ACQ_REL) == c->n) {
        __atomic_store_n(&s->arrived, 0, __ATOMIC_RELEASE);
        __atomic_add_fetch(&s->generation, 1, __ATOMIC_ACQ_REL);
        return 0;
    }
    const double t0 = now_s();
    while (__atomic_load_n(&s->generation, __ATOMIC_ACQUIRE) == gen) {
        sched_yield();
        if (now_s() - t0 > 120.0) { fprintf(stderr, "fake_nccl: rank %d waited 120 s at a barrier\n", c->rank); return 1; }
    }
    return 0;
}

ncclResult_t ncclGetUniqueId(ncclUniqueId* id)
{
    memset(id, 0, sizeof(*id));
    snprintf(id->internal, sizeof(id->internal), "/nhd_fake_nccl_%d_%ld", (int)getpid(), (long)(now_s() * 1e6));
    return 0;
}

ncclResult_t ncclCommInitRank(comm_t** out, int n, ncclUniqueId id, int rank)
{
    if (n < 1 || n > FAKE_MAX_RANKS || rank < 0 || rank >= n || id.internal[0] != '/') return 4;
    comm_t* c = (comm_t*)calloc(1, sizeof(comm_t));
    id.internal[sizeof(c->name) - 1] = 0;
    c->rank = rank; c->n = n;
    snprintf(c->name, sizeof(c->name), "%s", id.internal);
    c->bytes = sizeof(shared_t) + (size_t)n * FAKE_SLOT_BYTES;
    int fd = shm_open(c->name, O_CREAT | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)c->bytes) != 0) { free(c); return 2; }
    c->sh = (shared_t*)mmap(NULL, c->bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (c->sh == MAP_FAILED) { free(c); return 2; }
    __atomic_add_fetch(&c->sh->attached, 1, __ATOMIC_ACQ_REL);
    const double t0 = now_s();
    while (__atomic_load_n(&c->sh->attached, __ATOMIC_ACQUIRE) < n) {       /* like NCCL: returns when everybody is in */
        sched_yield();
        if (now_s() - t0 > 120.0) { fprintf(stderr, "fake_nccl: rank %d alone after 120 s\n", rank); return 6; }
    }
    *out = c;
    return 0;
}

ncclResult_t ncclCommDestroy(comm_t* c)
{
    if (!c) return 0;
    if (c->rank == 0) shm_unlink(c->name);
    munmap(c->sh, c->bytes);
    free(c);
    return 0;
}

ncclResult_t ncclAllGather(const void* send, void* recv, size_t count, int dtype, comm_t* c, void* stream)
{
    (void)stream;
    if (dtype != 0) return 4;                                  /* ncclInt8 */
    if (count > FAKE_SLOT_BYTES) return 4;
    memcpy(c->sh->slots + (size_t)c->rank * FAKE_SLOT_BYTES, send, count);
    if (barrier(c)) return 1;
    for (int r = 0; r < c->n; r++)
        memmove((unsigned char*)recv + (size_t)r * count, c->sh->slots + (size_t)r * FAKE_SLOT_BYTES, count);
    return barrier(c);                                         /* nobody overwrites a slot somebody still reads */
}

const char* ncclGetErrorString(ncclResult_t r) { return r == 0 ? "no error" : "fake_nccl error"; }
