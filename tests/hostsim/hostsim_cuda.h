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
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? cudaSuccess : 2; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
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
