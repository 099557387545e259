// tools/emu_include/cuda_runtime.h -- TEST INFRASTRUCTURE: a host stand-in for the CUDA runtime and the SIMT execution model, just
// large enough to compile csrc/fm.cu with g++ (-DSBMF_SIMT_EMU -I tools/emu_include) and EXECUTE its kernels on the CPU:
//   * device memory = host memory (cudaMalloc -> malloc, cudaMemcpy -> memcpy), streams are synchronous;
//   * a kernel launch runs the kernel body once per (block, thread): the threads of a block are real host threads from a pool,
//     blocks run one after the other; __syncthreads() is a block barrier, __shfl_xor_sync() exchanges through a per-block buffer
//     between two warp barriers, __shared__ is a function-local static (one block at a time), atomics are __atomic builtins.
// What it checks: indexing, work lists, launch geometry, reduction trees, barrier placement, the host orchestration -- everything
// except the hardware itself.  Never part of the product; used by tests/test_fm_simt_emulation.py only.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct float2 { float x, y; };
struct double2 { double x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
template <class T> inline T min(T a, T b) { return b < a ? b : a; }
template <class T> inline T max(T a, T b) { return a < b ? b : a; }
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

// ---- runtime API ----------------------------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1 };
typedef struct emu_stream* cudaStream_t;
struct cudaDeviceProp { int major, minor, multiProcessorCount; };
// fresh "device" memory is filled with a byte pattern (SIMT_EMU_FILL, default 0xCD = garbage floats / huge indices): a kernel that
// relies on cudaMalloc returning zeros -- which fresh GPU memory often happens to be -- shows up here
inline cudaError_t cudaMalloc(void** p, size_t n)
{
    static const int fill = getenv("SIMT_EMU_FILL") ? (int)strtol(getenv("SIMT_EMU_FILL"), NULL, 0) : 0xCD;
    const size_t bytes = ((n ? n : 1) + 255) / 256 * 256;
    *p = aligned_alloc(256, bytes);   // cudaMalloc's alignment: the 256-bit accesses of the kernels are checked against it (UBSan)
    if (*p) memset(*p, fill, bytes);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t) { return cudaMemcpy(d, s, n, k); }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { return cudaMemset(d, v, n); }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)malloc(1); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated error"; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->major = 10; p->minor = 0; p->multiProcessorCount = 2; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { return cudaFree(p); }
inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
#define __constant__
#define cudaMemcpyToSymbol(sym, src, n) (memcpy((void*)&(sym), (src), (n)), cudaSuccess)
// events: no timing on the host; graphs, memory pools and IPC are refused, which the library treats as "not available"
enum { cudaErrorNotSupported = 801, cudaEventDisableTiming = 2, cudaStreamCaptureModeThreadLocal = 1, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold = 4 };
typedef struct emu_event* cudaEvent_t;
typedef struct emu_graph* cudaGraph_t;
typedef struct emu_graph_exec* cudaGraphExec_t;
typedef struct emu_pool* cudaMemPool_t;
struct cudaIpcMemHandle_t { char reserved[64]; };
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)malloc(1); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, int) { return cudaErrorNotSupported; }
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = nullptr; return cudaErrorNotSupported; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t*, cudaGraph_t, unsigned long long) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t*, int) { return cudaErrorNotSupported; }
inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void*) { return cudaErrorNotSupported; }
inline cudaError_t cudaMemPoolTrimTo(cudaMemPool_t, size_t) { return cudaErrorNotSupported; }
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorNotSupported; }
inline cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaErrorNotSupported; }

// ---- SIMT execution -------------------------------------------------------------------------------------------------------------
namespace simt {
struct Idx { unsigned x, y, z; };
struct Barrier {   // reusable barrier; waiters yield (far more host threads than cores, and episodes are short).  A thread that leaves the
                   // kernel early drops out of the barriers of its CTA and warp, like an exited thread on the GPU.
                   // Lock-free: state = live threads << 32 | threads waiting.  Whoever makes waiting reach live (the last arrival,
                   // or a drop that leaves only waiters) is alone at that moment -- every other live thread spins on gen -- so it
                   // may reset the count with a plain store before it opens the next generation.
    std::atomic<uint64_t> state{0};
    std::atomic<unsigned> gen{0};
    void arm(int n) { state.store((uint64_t)(unsigned)n << 32, std::memory_order_relaxed); }
    void wait()
    {
        const unsigned g = gen.load(std::memory_order_acquire);   // cannot advance before this thread has arrived
        const uint64_t old = state.fetch_add(1, std::memory_order_acq_rel);
        const uint64_t n = old >> 32, w = (old & 0xffffffffu) + 1;
        if (w >= n) {
            state.store(n << 32, std::memory_order_relaxed);
            gen.store(g + 1, std::memory_order_release);
            return;
        }
        int spins = 0;
        while (gen.load(std::memory_order_acquire) == g)
            if (++spins > 16) std::this_thread::yield();   // (sleeping -- futex or nanosleep back-off -- was measured 5-8x slower: episodes are short)
    }
    void drop()
    {
        const uint64_t old = state.fetch_sub((uint64_t)1 << 32, std::memory_order_acq_rel);
        const uint64_t n = (old >> 32) - 1, w = old & 0xffffffffu;
        if (n > 0 && w >= n) {
            state.store(n << 32, std::memory_order_relaxed);
            gen.fetch_add(1, std::memory_order_release);
        }
    }
};
struct Block {
    Barrier all, done;
    Barrier warp[32];
    uint64_t xch[2][1024];   // shuffle exchange, double-buffered: one barrier per shuffle (the next one writes the other half)
    std::vector<unsigned char> dyn_smem;
};
inline Block& block()
{
    static Block b;
    return b;
}
struct Pool {
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv_start, cv_done;
    uint64_t gen = 0;
    int active = 0, remaining = 0;
    bool stop = false;
    std::function<void(int)> job;
    void worker(int i)
    {
        uint64_t last = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m);
                cv_start.wait(lk, [&] { return stop || gen != last; });
                if (stop) return;
                last = gen;
                if (i >= active) continue;
            }
            job(i);
            std::unique_lock<std::mutex> lk(m);
            if (--remaining == 0) cv_done.notify_one();
        }
    }
    void run(int n, const std::function<void(int)>& f)
    {
        while ((int)th.size() < n) {
            const int i = (int)th.size();
            th.emplace_back([this, i] { worker(i); });
        }
        {
            std::unique_lock<std::mutex> lk(m);
            job = f;
            active = remaining = n;
            ++gen;
        }
        cv_start.notify_all();
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return remaining == 0; });
    }
    ~Pool()
    {
        {
            std::unique_lock<std::mutex> lk(m);
            stop = true;
        }
        cv_start.notify_all();
        for (auto& t : th) t.join();
    }
};
inline Pool& pool()
{
    static Pool p;
    return p;
}
inline thread_local Idx t_thread{0, 0, 0}, t_block{0, 0, 0};
inline thread_local dim3 t_bdim, t_gdim;
inline thread_local unsigned t_xpar = 0;   // which half of Block::xch this thread's next shuffle uses (same for all threads of a warp)

template <class F>
inline void launch(dim3 grid, dim3 blk, F&& body, size_t dyn_smem_bytes = 0)
{
    const int nt = (int)blk.x;
    if (nt <= 0 || nt > 1024 || blk.y != 1 || blk.z != 1 || grid.x == 0 || grid.y == 0) abort();
    Block& b = block();
    b.dyn_smem.assign(dyn_smem_bytes + 16, 0);
    auto arm = [&] {
        b.all.arm(nt);
        for (int w = 0; w < 32; ++w) b.warp[w].arm(std::max(0, std::min(32, nt - 32 * w)));
    };
    arm();
    b.done.arm(nt);
    pool().run(nt, [&](int t) {
        t_bdim = blk;
        t_gdim = grid;
        t_thread = Idx{(unsigned)t, 0, 0};
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                t_block = Idx{bx, by, 0};
                t_xpar = 0;
                body();
                b.all.drop();            // exited: later __syncthreads / shuffles of this CTA no longer wait for this thread
                b.warp[t >> 5].drop();
                b.done.wait();           // one CTA at a time (function-local __shared__ statics, exchange buffer)
                if (t == 0) arm();
                b.done.wait();
            }
    });
}
}  // namespace simt

#define threadIdx simt::t_thread
#define blockIdx simt::t_block
#define blockDim simt::t_bdim
#define gridDim simt::t_gdim

inline void __syncthreads() { simt::block().all.wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { simt::block().warp[threadIdx.x >> 5].wait(); }
template <class T>
inline T simt_exchange(T v, unsigned src_lane)
{
    static_assert(sizeof(T) <= 8, "shuffle payload");
    simt::Block& b = simt::block();
    const unsigned t = threadIdx.x, w = t >> 5;
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    // A thread can only reach the shuffle after the next (and overwrite this half) once every thread of the warp has passed the
    // next shuffle's barrier, i.e. has finished reading here.
    uint64_t* x = b.xch[simt::t_xpar];
    simt::t_xpar ^= 1u;
    x[t] = bits;
    b.warp[w].wait();
    const uint64_t got = x[(t & ~31u) | (src_lane & 31u)];
    T r;
    memcpy(&r, &got, sizeof(T));
    return r;
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask, int width = 32)
{
    const unsigned lane = threadIdx.x & 31u;
    unsigned src = lane ^ (unsigned)lane_mask;
    if ((src & ~(unsigned)(width - 1)) != (lane & ~(unsigned)(width - 1))) src = lane;   // outside the segment: own value
    return simt_exchange(v, src);
}
template <class T>
inline T __shfl_sync(unsigned, T v, int src_lane, int width = 32)
{
    const unsigned lane = threadIdx.x & 31u;
    return simt_exchange(v, (lane & ~(unsigned)(width - 1)) | ((unsigned)src_lane & (unsigned)(width - 1)));
}
template <class T>
inline T __ldg(const T* p) { return *p; }
template <class T>
inline T __ldcg(const T* p) { return *(const volatile T*)p; }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline uint32_t atomicMin(uint32_t* a, uint32_t v)
{
    uint32_t old = __atomic_load_n(a, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(a, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
inline uint32_t atomicMax(uint32_t* a, uint32_t v)
{
    uint32_t old = __atomic_load_n(a, __ATOMIC_RELAXED);
    while (v > old && !__atomic_compare_exchange_n(a, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
inline uint32_t atomicExch(uint32_t* a, uint32_t v) { return __atomic_exchange_n(a, v, __ATOMIC_RELAXED); }
inline uint32_t atomicAdd(uint32_t* a, uint32_t v) { return __atomic_fetch_add(a, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long* a, unsigned long long v) { return __atomic_fetch_add(a, v, __ATOMIC_RELAXED); }
