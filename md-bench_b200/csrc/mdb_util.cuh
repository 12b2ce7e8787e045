// mdb_util.cuh -- error handling, device buffers, launch accounting, exact-rounding helpers.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>
#include <nvtx3/nvToolsExt.h> // header-only (CUDA 12): a no-op unless a tool (nsys, ncu --nvtx) is attached

namespace mdb {

// NVTX ranges in place of the reference's LIKWID regions ("force", "reneighbour": verletlist/main.c:80-92,137-143,
// force_lj.c / force_eam.c) plus one range per C-ABI entry point, named after it.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline std::string fmt(const char* f, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, f);
    vsnprintf(buf, sizeof buf, f, ap);
    va_end(ap);
    return buf;
}

#define MDB_CUDA(expr)                                                                           \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            throw ::mdb::Error(::mdb::fmt("[CUDA Error]: %s: %s (%s:%d)", #expr,                 \
                cudaGetErrorString(e__), __FILE__, __LINE__));                                   \
    } while (0)

// Every kernel launch of the library goes through this macro: it counts launches (bench.py's
// gpu_launches) and checks the launch status.
#define MDB_LAUNCH(counter, kernel, grid, block, smem, stream, ...)                              \
    do {                                                                                         \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                              \
        ++(counter);                                                                             \
        MDB_CUDA(cudaPeekAtLastError());                                                         \
    } while (0)

inline size_t round_up(size_t n, size_t m) { return (n + m - 1) / m * m; }
inline unsigned grid_for(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

// Growable device buffer. Growth is geometric (x1.25) so that 67M-atom systems do not reallocate
// in the reference's 20000-element steps (verletlist/atom.c:17).
template <class T> struct DBuf {
    T* p       = nullptr;
    size_t cap = 0;
    // Allocations exported through CUDA IPC (the peer-store halo maps x, y, z of the neighbor GPUs) must not be freed while
    // another process still has them mapped: with `retire` set, an outgrown allocation is parked there instead of freed and
    // its owner (DomainGroup) frees it once every importer has closed its mapping.
    std::vector<void*>* retire = nullptr;
    void ensure(size_t n, bool keep, cudaStream_t s)
    {
        if (n <= cap) return;
        size_t ncap = round_up(n > cap + cap / 4 ? n : cap + cap / 4, 1024);
        T* q        = nullptr;
        MDB_CUDA(cudaMalloc(&q, ncap * sizeof(T)));
        if (p && keep && cap) MDB_CUDA(cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, s));
        if (p) {
            MDB_CUDA(cudaStreamSynchronize(s));
            if (retire) retire->push_back(p);
            else MDB_CUDA(cudaFree(p));
        }
        p   = q;
        cap = ncap;
    }
    void release()
    {
        if (p) {
            if (retire) retire->push_back(p);
            else cudaFree(p);
        }
        p   = nullptr;
        cap = 0;
    }
};

// Addressing of the neighbor list: element (i,k) lives at base(i) + k*sk.  Atoms are grouped in tiles
// of G = 2^gshift consecutive atoms; inside a tile the entries are stored k-major ([k][atom]), tiles
// follow each other.  G = 1: the reference's row-major rows; G = 32: one warp's k-th entries are one
// coalesced 128-byte line and a warp streams through one contiguous tile; G >= Nlocal (gshift 31):
// fully transposed.
struct NbLayout {
    size_t tile_stride; // ints per tile = G * rowlen
    size_t sk;          // = G
    int gshift;
    __host__ __device__ __forceinline__ size_t base(int i) const
    {
        return (size_t)((unsigned)i >> gshift) * tile_stride + ((unsigned)i & ((1u << gshift) - 1u));
    }
};

// Round-to-nearest single operations that the compiler may not contract or re-associate.  Used
// wherever bits decide list membership: bin index, ghost coordinates, list distance.
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double fma_rn(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// ---- packed FP32 (sm_100 add/mul/fma.f32x2 -> FADD2 / FMUL2 / FFMA2): two lanes' worth of arithmetic per issue slot.
// IEEE round-to-nearest per half, i.e. bit-identical to the scalar __f*_rn operations.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

} // namespace mdb
