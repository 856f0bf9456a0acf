/*
 * hostsim_cuda.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A minimal stand-in for the parts of the CUDA runtime API that wmb_context.cu uses, so
 * that the library's host logic and the kernels' phase functions (wmb_kernels.cuh) can be
 * executed on a machine without a GPU by the `-m "not gpu"` tests.  "Device memory" is
 * plain malloc memory, streams execute immediately, kernels are loops
 * (hostsim_launch.inl).  This is never compiled into libwmbus_b200.so and never loaded by
 * the rtl-wmbus_b200 package: the product has no CPU path.
 */
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int cudaError_t;
#define cudaSuccess 0
#define cudaErrorNotReady 600
typedef struct hs_stream { int dummy; } *cudaStream_t;
typedef struct hs_event { int dummy; } *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
#define cudaStreamNonBlocking 1
#define cudaDevAttrMultiProcessorCount 16

static inline const char *cudaGetErrorString(cudaError_t) { return "hostsim"; }
static inline cudaError_t cudaGetLastError(void) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n)
{
    /* the simulation pretends to have one device unless told otherwise (used to test the
     * "no device -> WMB_E_NODEVICE" path) */
    const char *e = getenv("WMB_HOSTSIM_NO_DEVICE");
    *n = (e && *e == '1') ? 0 : 1;
    return cudaSuccess;
}
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
/* cudaMalloc does not clear what it returns (and a process that creates context after context gets its own stale data
 * back), so the simulation hands out memory filled with a pattern -- 0xA5 by default, WMB_HOSTSIM_POISON=<byte> for
 * another -- and a kernel that reads what nothing has written yet changes a result here too.  The buffers are sized for
 * 1 GiB batches and mostly untouched by a test, so the pattern must not cost a write per page: the allocation is a
 * private mapping of one pattern-filled 32 MiB memory file, over and over (read faults share its pages, writes copy). */
#include <sys/mman.h>
#include <unistd.h>
#define HS_PAT_BYTES ((size_t)32 << 20)
struct hs_alloc_rec { void *p; size_t n; };
static inline struct hs_alloc_rec *hs_alloc_table(void) { static struct hs_alloc_rec t[4096]; return t; }
static inline int hs_pattern_fd(void)
{
    static int fd = -1;
    if (fd >= 0) return fd;
    const char *e = getenv("WMB_HOSTSIM_POISON");
    const int fill = e ? (int)(strtoul(e, NULL, 0) & 0xFF) : 0xA5;
    fd = memfd_create("wmb_hostsim_pattern", 0);
    if (fd < 0 || ftruncate(fd, (off_t)HS_PAT_BYTES) != 0) abort();
    void *m = mmap(NULL, HS_PAT_BYTES, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (m == MAP_FAILED) abort();
    memset(m, fill, HS_PAT_BYTES);
    munmap(m, HS_PAT_BYTES);
    return fd;
}
static inline cudaError_t cudaMalloc(void **p, size_t n)
{
    const size_t page = 4096, len = ((n ? n : 1) + page - 1) / page * page;
    const int fd = hs_pattern_fd();
    uint8_t *base = (uint8_t *)mmap(NULL, len, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (base == MAP_FAILED) return 2;
    for (size_t off = 0; off < len; off += HS_PAT_BYTES) {
        const size_t k = len - off < HS_PAT_BYTES ? len - off : HS_PAT_BYTES;
        if (mmap(base + off, k, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_FIXED | MAP_NORESERVE, fd, 0) == MAP_FAILED) { munmap(base, len); return 2; }
    }
    struct hs_alloc_rec *t = hs_alloc_table();
    for (int i = 0; i < 4096; i++) if (!t[i].p) { t[i].p = base; t[i].n = len; *p = base; return cudaSuccess; }
    munmap(base, len);
    return 2;
}
static inline cudaError_t cudaFree(void *p)
{
    if (!p) return cudaSuccess;
    struct hs_alloc_rec *t = hs_alloc_table();
    for (int i = 0; i < 4096; i++) if (t[i].p == p) { munmap(p, t[i].n); t[i].p = NULL; return cudaSuccess; }
    abort();                                                  /* not ours */
}
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { return cudaFree(p); }
static inline cudaError_t cudaMemset(void *p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)calloc(1, sizeof(struct hs_stream)); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
#define cudaEventDisableTiming 2
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t)calloc(1, sizeof(struct hs_event)); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
