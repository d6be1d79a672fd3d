// cuda_runtime.h - TEST INFRASTRUCTURE: a host emulation of the CUDA execution model, just large enough to run the
// byte / index kernels of libinfercnv_b200 (icnv_regions.cu, icnv_ingest.cu, icnv_reduce.cu) and the host entry points
// of icnv_api.cu on a machine without a GPU, from the SAME source text.  It shadows <cuda_runtime.h> on the include
// path of tests/host/build_emu.py only; nothing in the package refers to it and the library it produces is never
// loaded by infercnv_b200/ (the product has no CPU path - see tests/test_capi_symbols.py).
//
// Model: a launch runs its blocks one after another on the calling thread; the threads of a block are fibers
// (ucontext) scheduled round-robin, each running until it finishes or reaches a synchronisation point:
//   __syncthreads()            - released when every live thread of the block waits there
//   __shfl_*_sync / __reduce_* - released when every live lane of the warp waits at a warp collective
// A block in which some live threads wait at the block barrier and others (of a not yet complete warp) at a warp
// collective for ever is reported as a divergence error and aborts: the emulation doubles as a barrier-divergence check.
// Atomics are plain operations (one OS thread).  "Device memory" is host memory; streams and events are no-ops.
#pragma once

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <functional>
#include <vector>

// ---- language extensions ------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x = 0, y = 0, z = 0; };
struct dim3 {
    unsigned x = 1, y = 1, z = 1;
    dim3() {}
    dim3(unsigned x_, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
struct double2 { double x, y; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

inline uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

using std::max;
using std::min;

// ---- runtime API --------------------------------------------------------------------------------------------------
typedef int cudaError_t;
typedef struct emuStream *cudaStream_t;
typedef struct emuEvent *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNoDevice = 100 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
struct cudaDeviceProp {
    int major, minor, multiProcessorCount;
    size_t sharedMemPerBlockOptin;
};

static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    p->major = 10; p->minor = 0; p->multiProcessorCount = 4; p->sharedMemPerBlockOptin = 227 * 1024;
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { return cudaStreamCreateWithFlags(s, 0); }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (cudaEvent_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void **p, size_t n) {
    // poison fresh "device" memory so a kernel that reads what it never wrote is noticed
    *p = malloc(n);
    if (!*p) return cudaErrorMemoryAllocation;
    memset(*p, 0xA5, n);
    return cudaSuccess;
}
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) {
    if (n) memmove(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) {
    if (n) memset(d, v, n);
    return cudaSuccess;
}
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }

// ---- execution model ----------------------------------------------------------------------------------------------
namespace emu {

enum State { RUN = 0, WAIT_BLOCK, WAIT_WARP, DONE };
struct Fiber {
    ucontext_t ctx;
    State state = DONE;
    unsigned gen = 0;          // number of warp collectives this thread has entered
};
constexpr size_t STACK_BYTES = 64 * 1024;
constexpr int MAX_THREADS = 1024;

inline std::vector<Fiber> fibers;
inline std::vector<char> stacks;
inline ucontext_t sched_ctx;
inline Fiber *cur = nullptr;
inline const std::function<void()> *body = nullptr;
inline uint64_t warp_slot[2][MAX_THREADS / 32][32];
inline unsigned long long launches = 0, blocks_run = 0;

inline void yield_to_scheduler() { swapcontext(&cur->ctx, &sched_ctx); }
inline void trampoline() {
    (*body)();
    cur->state = DONE;
    swapcontext(&cur->ctx, &sched_ctx);
}

inline void run_block(unsigned nt, const std::function<void()> &fn) {
    if (nt == 0 || nt > MAX_THREADS || (nt & 31)) {
        fprintf(stderr, "emu: block size %u not supported (multiple of 32, <= 1024)\n", nt);
        abort();
    }
    body = &fn;
    if (fibers.size() < nt) fibers.resize(nt);
    if (stacks.size() < (size_t)nt * STACK_BYTES) stacks.resize((size_t)nt * STACK_BYTES);
    for (unsigned t = 0; t < nt; ++t) {
        Fiber &f = fibers[t];
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = stacks.data() + (size_t)t * STACK_BYTES;
        f.ctx.uc_stack.ss_size = STACK_BYTES;
        f.ctx.uc_link = &sched_ctx;
        makecontext(&f.ctx, trampoline, 0);
        f.state = RUN;
        f.gen = 0;
    }
    unsigned alive = nt;
    while (alive > 0) {
        bool ran = false;
        for (unsigned t = 0; t < nt; ++t) {
            Fiber &f = fibers[t];
            if (f.state != RUN) continue;
            cur = &f;
            threadIdx.x = t;
            swapcontext(&sched_ctx, &f.ctx);
            ran = true;
            if (f.state == DONE) --alive;
        }
        if (alive == 0) break;
        bool released = false;
        for (unsigned w = 0; w < nt / 32; ++w) {      // warp collectives first
            unsigned live = 0, waiting = 0;
            for (unsigned l = 0; l < 32; ++l) {
                const State s = fibers[w * 32 + l].state;
                live += s != DONE;
                waiting += s == WAIT_WARP;
            }
            if (live && waiting == live) {
                for (unsigned l = 0; l < 32; ++l)
                    if (fibers[w * 32 + l].state == WAIT_WARP) fibers[w * 32 + l].state = RUN;
                released = true;
            }
        }
        if (!released) {
            unsigned waiting = 0;
            for (unsigned t = 0; t < nt; ++t) waiting += fibers[t].state == WAIT_BLOCK;
            if (waiting == alive) {
                for (unsigned t = 0; t < nt; ++t)
                    if (fibers[t].state == WAIT_BLOCK) fibers[t].state = RUN;
                released = true;
            }
        }
        if (!released && !ran) {
            fprintf(stderr, "emu: divergent synchronisation in block (%u,%u): the live threads wait at different barriers\n",
                    blockIdx.x, blockIdx.y);
            abort();
        }
    }
    ++blocks_run;
}

struct Cfg {
    dim3 grid, block;
    Cfg(dim3 g, dim3 b, size_t = 0, cudaStream_t = nullptr) : grid(g), block(b) {}
};

inline void launch(const Cfg &cfg, const std::function<void()> &fn) {
    if (cfg.block.y != 1 || cfg.block.z != 1 || cfg.grid.z != 1) {
        fprintf(stderr, "emu: only (x, y) grids of 1-D blocks are supported\n");
        abort();
    }
    gridDim = cfg.grid;
    blockDim = cfg.block;
    ++launches;
    for (unsigned by = 0; by < cfg.grid.y; ++by)
        for (unsigned bx = 0; bx < cfg.grid.x; ++bx) {
            blockIdx.x = bx;
            blockIdx.y = by;
            run_block(cfg.block.x, fn);
        }
}

template <class T>
inline T warp_exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    Fiber *f = cur;
    const unsigned tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = f->gen++ & 1u;
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    warp_slot[g][w][lane] = bits;
    f->state = WAIT_WARP;
    yield_to_scheduler();
    if (src_lane < 0 || src_lane > 31 || fibers[w * 32 + (unsigned)src_lane].state == DONE) return v;
    T out;
    memcpy(&out, &warp_slot[g][w][src_lane], sizeof(T));
    return out;
}

}  // namespace emu

#define EMU_LAUNCH(kernel, cfg, args) emu::launch(emu::Cfg cfg, [&]() { kernel args; })

static inline void __syncthreads() {
    emu::cur->state = emu::WAIT_BLOCK;
    emu::yield_to_scheduler();
}
static inline void __syncwarp(unsigned = 0xffffffffu) { (void)emu::warp_exchange<int>(0, 0); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emu::warp_exchange(v, src & 31); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu::warp_exchange(v, (int)((threadIdx.x & 31) ^ (unsigned)m)); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    const int lane = (int)(threadIdx.x & 31);
    return emu::warp_exchange(v, lane >= (int)d ? lane - (int)d : lane);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d) {
    const int lane = (int)(threadIdx.x & 31);
    return emu::warp_exchange(v, lane + (int)d <= 31 ? lane + (int)d : lane);
}

// ---- intrinsics ---------------------------------------------------------------------------------------------------
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __double2hiint(double v) { uint64_t u; memcpy(&u, &v, 8); return (int)(u >> 32); }
static inline int __double2loint(double v) { uint64_t u; memcpy(&u, &v, 8); return (int)(u & 0xffffffffu); }
static inline long long __double_as_longlong(double v) { long long u; memcpy(&u, &v, 8); return u; }
static inline double __longlong_as_double(long long u) { double v; memcpy(&v, &u, 8); return v; }

static inline double atomicAdd(double *p, double v) { const double o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
static inline int atomicAdd(int *p, int v) { const int o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
static inline int atomicOr(int *p, int v) { const int o = *p; *p = o | v; return o; }
static inline int atomicExch(int *p, int v) { const int o = *p; *p = v; return o; }
