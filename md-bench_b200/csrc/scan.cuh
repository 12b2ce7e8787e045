// scan.cuh -- exclusive prefix sum over int32 on the device (bin offsets, ghost offsets).
// Three-phase tile scan (tile scan -> recursive scan of tile totals -> add offsets); deterministic.
#pragma once
#include "mdb_util.cuh"

namespace mdb {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS   = 8;
constexpr int SCAN_TILE    = SCAN_THREADS * SCAN_ITEMS;

static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile(
    const int* __restrict__ in, int* __restrict__ out, int* __restrict__ tile_sums, size_t n)
{
    __shared__ int warp_sums[SCAN_THREADS / 32];
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        sum += v[k];
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
        int t = s;
#pragma unroll
        for (int d = 1; d < SCAN_THREADS / 32; d <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, t, d);
            if (lane >= d) t += u;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = t - s; // exclusive warp offsets
        if (lane == SCAN_THREADS / 32 - 1) tile_sums[blockIdx.x] = t;
    }
    __syncthreads();
    int run = inc - sum + warp_sums[w];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

static __global__ void k_scan_add(int* __restrict__ out, const int* __restrict__ tile_off, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += tile_off[i / SCAN_TILE];
}

static __global__ void k_scan_total(const int* __restrict__ tile_sums, int* __restrict__ dst)
{
    *dst = tile_sums[0];
}

struct Scanner {
    DBuf<int> lvl[4]; // tile sums per recursion level (+ their scans in place)
    long long* launches = nullptr;

    // out[0..n) = exclusive scan of in[0..n); *total_dst (device pointer, may be null) = sum.
    void preload() // see Sim::preload_kernels
    {
        cudaFuncAttributes a;
        cudaFuncGetAttributes(&a, (const void*)k_scan_tile);
        cudaFuncGetAttributes(&a, (const void*)k_scan_add);
        cudaFuncGetAttributes(&a, (const void*)k_scan_total);
    }
    void exclusive(const int* in, int* out, size_t n, int* total_dst, cudaStream_t s, int depth = 0)
    {
        if (depth >= 4) throw Error("scan: input too large");
        const size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
        lvl[depth].ensure(tiles + 1, false, s);
        int* sums = lvl[depth].p;
        if (n == 0) {
            if (total_dst) MDB_CUDA(cudaMemsetAsync(total_dst, 0, sizeof(int), s));
            return;
        }
        MDB_LAUNCH(*launches, k_scan_tile, (unsigned)tiles, SCAN_THREADS, 0, s, in, out, sums, n);
        if (tiles == 1) {
            if (total_dst) MDB_LAUNCH(*launches, k_scan_total, 1, 1, 0, s, sums, total_dst);
            return;
        }
        exclusive(sums, sums, tiles, total_dst, s, depth + 1); // in-place is safe: tile reads precede writes
        MDB_LAUNCH(*launches, k_scan_add, grid_for(n, 256), 256, 0, s, out, sums, n);
    }
    void release()
    {
        for (auto& b : lvl) b.release();
    }
};

} // namespace mdb
