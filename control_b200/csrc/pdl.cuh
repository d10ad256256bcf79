// Programmatic dependent launch on the device side, and loads that keep their place before the wait.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifdef __CUDACC__
// Programmatic dependent launch (sm_90+): the time sweeps are chains of ~10^4 small dependent
// kernels per preconditioner application, so the gap between two launches matters as much as
// the kernels.  A kernel launched through pdl_launch may be scheduled while its predecessor
// drains; it calls pdl_sync() before its first global memory access, which (a) lets ITS
// successor be scheduled early and (b) waits until the predecessor grid has completed and its
// writes are visible.
__device__ __forceinline__ void pdl_sync()
{
#if __CUDA_ARCH__ >= 900
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

// The two halves of pdl_sync() for kernels that start with loads of CONSTANT data (matrix entries, inverse
// diagonals: written at setup, never inside a sweep): pdl_trigger() first, the constant loads next, pdl_wait()
// before the first access to anything a kernel of the chain writes.  The constant loads then overlap the tail of
// the predecessor -- in a chain of ~5 us kernels made of three or four dependent memory round trips each, that is
// one or two round trips per kernel taken off the critical path.  pdl_wait() may be called more than once.
// Note that before the wait not even the predecessor's predecessor is known to be complete (a grid triggers its
// dependents before its own wait), so nothing but constant data may be touched there.
__device__ __forceinline__ void pdl_trigger()
{
#if __CUDA_ARCH__ >= 900
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

__device__ __forceinline__ void pdl_wait()
{
#if __CUDA_ARCH__ >= 900
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

// Loads of constant data that must be ISSUED before pdl_wait().  ptxas sinks a read-only (ld.global.nc / __ldg)
// load below the wait even when it is a volatile asm statement -- checked in the SASS: ACQBULK first, LDG.CONSTANT
// after it -- but keeps a plain ld.global (and ld.global.cs) where it was written.  STREAM: evict-first (a matrix stream that does not fit L2 next to the vectors).
template <bool STREAM = false>
__device__ __forceinline__ int pre_ld(const int *p)
{
    int v;
    if (STREAM) asm volatile("ld.global.cs.s32 %0, [%1];" : "=r"(v) : "l"(p));
    else asm volatile("ld.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
template <bool STREAM = false>
__device__ __forceinline__ unsigned pre_ld(const uint16_t *p)
{
    unsigned short v;
    if (STREAM) asm volatile("ld.global.cs.u16 %0, [%1];" : "=h"(v) : "l"(p));
    else asm volatile("ld.global.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return v;
}
template <bool STREAM = false>
__device__ __forceinline__ unsigned pre_ld(const uint8_t *p)
{
    unsigned v;
    if (STREAM) asm volatile("ld.global.cs.u8 %0, [%1];" : "=r"(v) : "l"(p));
    else asm volatile("ld.global.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
template <bool STREAM = false>
__device__ __forceinline__ double pre_ld(const double *p)
{
    double v;
    if (STREAM) asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int2 pre_ld(const int2 *p)
{
    int2 v;
    asm volatile("ld.global.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ int4 pre_ld(const int4 *p)
{
    int4 v;
    asm volatile("ld.global.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 pre_ld(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
#endif
