// cuda_sim.hpp — TEST INFRASTRUCTURE ONLY.  A minimal "CUDA on the CPU" layer that lets g++ compile the product's .cu/.cuh
// sources unchanged (after tests/hostsim/build.py rewrites the <<<...>>> launches) and run every kernel FUNCTIONALLY on the host:
// each CUDA thread of a block is a fiber (ucontext) of one OS thread, blocks run one after the other, __syncthreads and the warp
// intrinsics are rendez-vous points between the fibers.  It exists because the C-ABI plumbing (table packing, struct layouts,
// argument order, kernel control flow) can then be exercised by the `-m gpu` tests in a container that has no GPU:
//     python tests/hostsim/build.py && CCK_SO=tests/hostsim/_build/libcck_b200.so CCK_HOSTSIM=1 python -m pytest tests -m gpu
// It is NOT a product path (the product library refuses to load without a GPU), says nothing about performance, and does not model
// memory spaces, asynchrony, cp.async visibility or divergence - a kernel that passes here can still be wrong on the device.
#pragma once
#include <ucontext.h>
#include <sys/mman.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <vector>

#define CCK_HOSTSIM 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define CCK_SIM_NOINLINE __attribute__((noinline))   // build.py rewrites __noinline__ (libstdc++ uses that spelling itself)
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct __attribute__((aligned(16))) double2 { double x, y; };
struct __attribute__((aligned(16))) double4 { double x, y, z, w; };
inline double4 make_double4(double x, double y, double z, double w) { return double4{x, y, z, w}; }
struct __attribute__((aligned(8))) float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(8))) int2 { int x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(8))) uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline int2 make_int2(int x, int y) { return int2{x, y}; }
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace sim {

struct Fiber {
    ucontext_t ctx;
    void* stack = nullptr;
    bool done = false, released = false;
    uint3 tid3{0, 0, 0};
    int tid = 0;
};
struct Block {
    int nthreads = 0, n_alive = 0, b_arrived = 0;
    unsigned b_gen = 0;
    uint3 bid{0, 0, 0};
    dim3 bdim, gdim;
    std::vector<Fiber> fibers;
    unsigned alive[32] = {}, arrived[32] = {};
    uint64_t slot[32][32] = {};
    unsigned char* dyn = nullptr;
    ucontext_t main;
    std::function<void()> body;
    long progress = 0;
};
inline Block* g_blk = nullptr;
inline Fiber* g_cur = nullptr;
inline std::recursive_mutex g_launch_mutex;
constexpr size_t kStack = 256 << 10;
inline std::vector<void*>& stack_pool() { static std::vector<void*> p; return p; }

inline void yield() { swapcontext(&g_cur->ctx, &g_blk->main); }

inline void trampoline() {
    Block& B = *g_blk;
    Fiber& f = *g_cur;
    B.body();
    f.done = true;
    B.alive[f.tid >> 5] &= ~(1u << (f.tid & 31));
    B.n_alive--;
    B.progress++;
    // returns to uc_link = B.main
}

inline void warp_sync(unsigned mask) {
    Block& B = *g_blk;
    Fiber& f = *g_cur;
    const int w = f.tid >> 5;
    const unsigned bit = 1u << (f.tid & 31);
    B.arrived[w] |= bit;
    f.released = false;
    B.progress++;
    for (;;) {
        if (f.released) return;
        const unsigned need = mask & B.alive[w];
        if ((B.arrived[w] & need) == need) {
            for (int l = 0; l < 32; ++l)
                if (need >> l & 1) B.fibers[(w << 5) + l].released = true;
            B.arrived[w] &= ~need;
            B.progress++;
            return;
        }
        yield();
    }
}
inline void syncthreads() {
    Block& B = *g_blk;
    const unsigned gen = B.b_gen;
    B.b_arrived++;
    B.progress++;
    for (;;) {
        if (B.b_gen != gen) return;
        if (B.b_arrived >= B.n_alive) { B.b_arrived = 0; B.b_gen++; B.progress++; return; }
        yield();
    }
}

inline unsigned char* dyn_smem() { return g_blk->dyn; }

struct Cfg { dim3 grid, block; size_t smem; };
template <class S = void*>
inline Cfg cfg(dim3 g, dim3 b, size_t smem = 0, S = S()) { return Cfg{g, b, smem}; }

inline void run_block(Block& B) {
    const int n = B.nthreads;
    B.fibers.assign(n, Fiber());
    for (int w = 0; w < 32; ++w) { B.alive[w] = 0; B.arrived[w] = 0; }
    B.n_alive = n; B.b_arrived = 0; B.b_gen = 0;
    auto& pool = stack_pool();
    for (int i = 0; i < n; ++i) {
        Fiber& f = B.fibers[i];
        if (pool.empty()) {
            void* s = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (s == MAP_FAILED) { std::fprintf(stderr, "hostsim: mmap of a fiber stack failed\n"); std::abort(); }
            f.stack = s;
        } else { f.stack = pool.back(); pool.pop_back(); }
        f.tid = i;
        f.tid3 = uint3{(unsigned)(i % B.bdim.x), (unsigned)((i / B.bdim.x) % B.bdim.y), (unsigned)(i / (B.bdim.x * B.bdim.y))};
        B.alive[i >> 5] |= 1u << (i & 31);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = &B.main;
        makecontext(&f.ctx, (void (*)())trampoline, 0);
    }
    // CCK_HOSTSIM_SCHED=reverse | random[:seed]: the order in which runnable threads are resumed between rendez-vous points.
    // A kernel whose results depend on it has a race (a missing __syncthreads / __syncwarp, an unordered shared-memory update).
    static const int sched = [] { const char* e = std::getenv("CCK_HOSTSIM_SCHED"); return !e ? 0 : (e[0] == 'r' && e[1] == 'e') ? 1 : (e[0] == 'r' && e[1] == 'a') ? 2 : 0; }();
    static uint64_t rng = [] { const char* e = std::getenv("CCK_HOSTSIM_SCHED"); const char* c = e ? std::strchr(e, ':') : nullptr; return c ? std::strtoull(c + 1, nullptr, 10) * 2 + 1 : 0x9E3779B97F4A7C15ull; }();
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = sched == 1 ? n - 1 - i : i;
    for (;;) {
        bool any = false;
        const long p0 = B.progress;
        if (sched == 2)
            for (int i = n - 1; i > 0; --i) { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; std::swap(order[i], order[(int)(rng % (uint64_t)(i + 1))]); }
        for (int oi = 0; oi < n; ++oi) {
            const int i = order[oi];
            if (B.fibers[i].done) continue;
            any = true;
            g_cur = &B.fibers[i];
            swapcontext(&B.main, &B.fibers[i].ctx);
        }
        if (!any) break;
        if (B.progress == p0) {
            std::fprintf(stderr, "hostsim: deadlock in block (%u,%u,%u): no thread made progress (mismatched barrier or spin on a value nobody writes)\n",
                         B.bid.x, B.bid.y, B.bid.z);
            std::abort();
        }
    }
    for (int i = 0; i < n; ++i) pool.push_back(B.fibers[i].stack);
    g_cur = nullptr;
}

inline long& launch_counter() { static long n = 0; return n; }

template <class... P, class... A>
inline void launch(const Cfg& c, void (*kernel)(P...), A&&... args) {
    std::lock_guard<std::recursive_mutex> lk(g_launch_mutex);
    launch_counter()++;
    Block B;
    B.bdim = c.block; B.gdim = c.grid;
    B.nthreads = (int)(c.block.x * c.block.y * c.block.z);
    if (B.nthreads <= 0 || B.nthreads > 1024) { std::fprintf(stderr, "hostsim: bad block size %d\n", B.nthreads); std::abort(); }
    std::vector<unsigned char> dyn(c.smem + 256);
    B.dyn = (unsigned char*)(((uintptr_t)dyn.data() + 255) & ~(uintptr_t)255);
    auto call = [&]() { kernel(static_cast<P>(args)...); };
    B.body = call;
    Block* prev = g_blk;
    g_blk = &B;
    for (unsigned z = 0; z < c.grid.z; ++z)
        for (unsigned y = 0; y < c.grid.y; ++y)
            for (unsigned x = 0; x < c.grid.x; ++x) {
                B.bid = uint3{x, y, z};
                run_block(B);
            }
    g_blk = prev;
}

template <class T>
inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, "shuffle type"); std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T>
inline T from_bits(uint64_t b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }

template <class T>
inline T shfl_from(unsigned mask, T v, int src) {
    Block& B = *g_blk;
    const int w = g_cur->tid >> 5, lane = g_cur->tid & 31;
    B.slot[w][lane] = to_bits(v);
    warp_sync(mask);
    const T r = (src >= 0 && src < 32) ? from_bits<T>(B.slot[w][src]) : v;
    warp_sync(mask);
    return r;
}
inline unsigned ballot(unsigned mask, bool pred) {
    Block& B = *g_blk;
    const int w = g_cur->tid >> 5, lane = g_cur->tid & 31;
    B.slot[w][lane] = pred ? 1 : 0;
    warp_sync(mask);
    unsigned r = 0;
    const unsigned part = mask & (B.alive[w] | (1u << lane));
    for (int l = 0; l < 32; ++l)
        if ((part >> l & 1) && B.slot[w][l]) r |= 1u << l;
    warp_sync(mask);
    return r;
}

}  // namespace sim

#define threadIdx (sim::g_cur->tid3)
#define blockIdx (sim::g_blk->bid)
#define blockDim (sim::g_blk->bdim)
#define gridDim (sim::g_blk->gdim)

// ---- synchronisation and warp intrinsics -------------------------------------------------------------------------------------
inline void __syncthreads() { sim::syncthreads(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { sim::warp_sync(mask); }
template <class T> inline T __shfl_sync(unsigned m, T v, int src) { return sim::shfl_from(m, v, src & 31); }
template <class T> inline T __shfl_xor_sync(unsigned m, T v, int x) { return sim::shfl_from(m, v, (sim::g_cur->tid & 31) ^ x); }
template <class T> inline T __shfl_down_sync(unsigned m, T v, int d) { return sim::shfl_from(m, v, (sim::g_cur->tid & 31) + d); }
template <class T> inline T __shfl_up_sync(unsigned m, T v, int d) { return sim::shfl_from(m, v, (sim::g_cur->tid & 31) - d); }
inline unsigned __ballot_sync(unsigned m, int pred) { return sim::ballot(m, pred != 0); }
inline int __all_sync(unsigned m, int pred) {
    const unsigned b = sim::ballot(m, pred != 0);
    const unsigned part = m & (sim::g_blk->alive[sim::g_cur->tid >> 5]);
    return (b & part) == part;
}
inline int __any_sync(unsigned m, int pred) { return sim::ballot(m, pred != 0) != 0; }
inline int __reduce_min_sync(unsigned m, int v) {
    int r = v;
    for (int x = 16; x >= 1; x >>= 1) { const int o = __shfl_xor_sync(m, r, x); r = o < r ? o : r; }
    return r;
}
inline unsigned __reduce_min_sync(unsigned m, unsigned v) {
    unsigned r = v;
    for (int x = 16; x >= 1; x >>= 1) { const unsigned o = __shfl_xor_sync(m, r, x); r = o < r ? o : r; }
    return r;
}

// ---- atomics (one OS thread runs all fibers of a launch; launches are serialised) ---------------------------------------------
template <class T> inline T atomicAdd(T* p, T v) { const T o = *p; *p = o + v; return o; }
inline unsigned long long atomicAdd(unsigned long long* p, long long v) { const unsigned long long o = *p; *p = o + (unsigned long long)v; return o; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long v) { const unsigned long long o = *p; *p = o + v; return o; }
inline int atomicAdd(int* p, unsigned v) { const int o = *p; *p = o + (int)v; return o; }
inline unsigned atomicAdd(unsigned* p, int v) { const unsigned o = *p; *p = o + (unsigned)v; return o; }
template <class T> inline T atomicMin(T* p, T v) { const T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicMax(T* p, T v) { const T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicExch(T* p, T v) { const T o = *p; *p = v; return o; }
template <class T> inline T atomicCAS(T* p, T cmp, T v) { const T o = *p; if (o == cmp) *p = v; return o; }
template <class T> inline T atomicOr(T* p, T v) { const T o = *p; *p = o | v; return o; }
template <class T> inline T atomicAnd(T* p, T v) { const T o = *p; *p = o & v; return o; }
inline void __threadfence() {}
inline void __threadfence_block() {}

// ---- scalar intrinsics ----------------------------------------------------------------------------------------------------------
template <class T> inline T __ldg(const T* p) { return *p; }
inline float __double2float_rn(double d) { return (float)d; }
inline float __double2float_ru(double d) { float f = (float)d; if ((double)f < d) f = std::nextafterf(f, INFINITY); return f; }
inline float __double2float_rd(double d) { float f = (float)d; if ((double)f > d) f = std::nextafterf(f, -INFINITY); return f; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
// directed rounding through the exact double result (float +,-,* of two floats are exact in double unless the exponents are > 29 apart)
inline float __fsub_rd(float a, float b) { return __double2float_rd((double)a - (double)b); }
inline float __fsub_ru(float a, float b) { return __double2float_ru((double)a - (double)b); }
inline float __fadd_rd(float a, float b) { return __double2float_rd((double)a + (double)b); }
inline float __fadd_ru(float a, float b) { return __double2float_ru((double)a + (double)b); }
inline float __fmul_rd(float a, float b) { return __double2float_rd((double)a * (double)b); }
inline float __fmul_ru(float a, float b) { return __double2float_ru((double)a * (double)b); }
inline int __float2int_rz(float f) { return std::isnan(f) ? 0 : (f >= 2147483648.f ? INT32_MAX : (f <= -2147483648.f ? INT32_MIN : (int)f)); }
inline int __float2int_rd(float f) { return __float2int_rz(std::floor(f)); }
inline int __double2int_rz(double f) { return std::isnan(f) ? 0 : (f >= 2147483648.0 ? INT32_MAX : (f <= -2147483648.0 ? INT32_MIN : (int)f)); }
inline int __double2hiint(double d) { long long l; std::memcpy(&l, &d, 8); return (int)(l >> 32); }
inline int __double2loint(double d) { long long l; std::memcpy(&l, &d, 8); return (int)(l & 0xffffffffll); }
inline double __hiloint2double(int hi, int lo) { const unsigned long long l = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo; double d; std::memcpy(&d, &l, 8); return d; }
inline long long __double_as_longlong(double d) { long long l; std::memcpy(&l, &d, 8); return l; }
inline double __longlong_as_double(long long l) { double d; std::memcpy(&d, &l, 8); return d; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline float __fdividef(float a, float b) { return a / b; }
inline float __saturatef(float a) { return a < 0.f ? 0.f : (a > 1.f ? 1.f : a); }
// erfinv: CUDA's is accurate to a few ulp; here a rational start refined by Newton steps on std::erf (test infrastructure: every
// decision that feeds on it is band-asserted in the tests)
inline double erfinv(double y) {
    if (std::isnan(y) || y < -1.0 || y > 1.0) return NAN;
    if (y == 1.0) return INFINITY;
    if (y == -1.0) return -INFINITY;
    if (y == 0.0) return y;
    const double w0 = -std::log((1.0 - y) * (1.0 + y));
    double x;
    if (w0 < 5.0) {
        const double w = w0 - 2.5;
        double p = 2.81022636e-08;
        p = 3.43273939e-07 + p * w; p = -3.5233877e-06 + p * w; p = -4.39150654e-06 + p * w; p = 0.00021858087 + p * w;
        p = -0.00125372503 + p * w; p = -0.00417768164 + p * w; p = 0.246640727 + p * w; p = 1.50140941 + p * w;
        x = p * y;
    } else {
        const double w = std::sqrt(w0) - 3.0;
        double p = -0.000200214257;
        p = 0.000100950558 + p * w; p = 0.00134934322 + p * w; p = -0.00367342844 + p * w; p = 0.00573950773 + p * w;
        p = -0.0076224613 + p * w; p = 0.00943887047 + p * w; p = 1.00167406 + p * w; p = 2.83297682 + p * w;
        x = p * y;
    }
    const double k = 0.88622692545275801365;   // sqrt(pi) / 2
    for (int it = 0; it < 3; ++it) {
        const double ax = std::fabs(x);
        // erf(x) - y, evaluated through erfc in the tails to keep relative accuracy
        const double err = ax > 0.5 ? (y > 0 ? (1.0 - y) - std::erfc(x) : std::erfc(-x) - (1.0 + y)) : std::erf(x) - y;
        const double d = err * k * std::exp(x * x);
        x -= d / (1.0 + x * d);
    }
    return x;
}
inline double erfcinv(double y) { return erfinv(1.0 - y); }
inline double normcdf(double x) { return 0.5 * std::erfc(-x * 0.70710678118654752440); }
inline double normcdfinv(double p) { return 1.41421356237309504880 * erfinv(2.0 * p - 1.0); }
using std::isnan;
using std::isinf;
using std::isfinite;

// ---- the slice of the CUDA runtime API the library uses: device memory is host memory, streams are synchronous ---------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
typedef void* cudaStream_t;
struct SimEvent { std::chrono::steady_clock::time_point t; };
typedef SimEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp {
    char name[256];
    int major, minor, multiProcessorCount;
    size_t sharedMemPerBlockOptin, totalGlobalMem;
};
inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : (e == cudaErrorMemoryAllocation ? "out of memory" : "invalid value"); }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) {      // CCK_HOSTSIM_DEVICES: pretend to have several (multi-rank dry runs)
    const char* e = std::getenv("CCK_HOSTSIM_DEVICES");
    *n = e ? std::max(1, std::atoi(e)) : 1;
    return cudaSuccess;
}
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    std::memset(p, 0, sizeof(*p));
    std::snprintf(p->name, sizeof(p->name), "hostsim (CPU fibers, test infrastructure)");
    p->major = 10; p->minor = 0;
    p->multiProcessorCount = 2;            // small grids: blocks run one after the other
    p->sharedMemPerBlockOptin = 232448;    // 227 KB, as on sm_100
    p->totalGlobalMem = (size_t)8 << 30;
    return cudaSuccess;
}
template <class T>
inline cudaError_t cudaMalloc(T** p, size_t n) {
    void* q = nullptr;
    if (posix_memalign(&q, 256, (n + 255) & ~(size_t)255 ? (n + 255) & ~(size_t)255 : 256) != 0) { *p = nullptr; return cudaErrorMemoryAllocation; }
    std::memset(q, 0xCD, n);               // poison: reads of never-written device memory show up
    *p = (T*)q;
    return cudaSuccess;
}
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
template <class T>
inline cudaError_t cudaHostAlloc(T** p, size_t n, unsigned) { return cudaMalloc(p, n); }
template <class T>
inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) {
    std::lock_guard<std::recursive_mutex> lk(sim::g_launch_mutex);
    if (n) std::memmove(d, s, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t = nullptr) {
    std::lock_guard<std::recursive_mutex> lk(sim::g_launch_mutex);
    for (size_t r = 0; r < height; ++r) std::memmove((char*)d + r * dpitch, (const char*)s + r * spitch, width);
    return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { if (n) std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = std::malloc(8); return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = std::malloc(8); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new SimEvent(); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::max(1e-3f, std::chrono::duration<float, std::milli>(b->t - a->t).count());
    return cudaSuccess;
}
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
