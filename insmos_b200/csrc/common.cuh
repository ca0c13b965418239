// Shared device helpers for libinsmos_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>
#include <stdio.h>
#include "../../include/insmos_b200.h"

#define INSMOS_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define INSMOS_ROW_BITS 25
#define INSMOS_ROW_MASK ((1u << INSMOS_ROW_BITS) - 1u)

void insmos_set_last_error(const char* what, cudaError_t e);
extern unsigned long long g_insmos_launches;      // kernels launched by this library (insmos_launch_count)

// follows EVERY kernel launch of the library: counts it and turns a launch error into INSMOS_ERR_CUDA
#define INSMOS_CHECK_LAUNCH(what)                                   \
    do {                                                            \
        __atomic_fetch_add(&g_insmos_launches, 1ull, __ATOMIC_RELAXED); \
        cudaError_t _e = cudaGetLastError();                        \
        if (_e != cudaSuccess) {                                    \
            insmos_set_last_error(what, _e);                        \
            return INSMOS_ERR_CUDA;                                 \
        }                                                           \
    } while (0)

#define INSMOS_CHECK_CUDA(expr)                                     \
    do {                                                            \
        cudaError_t _e = (expr);                                    \
        if (_e != cudaSuccess) {                                    \
            insmos_set_last_error(#expr, _e);                       \
            return INSMOS_ERR_CUDA;                                 \
        }                                                           \
    } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Opt-in to > 48 KB of dynamic shared memory.  cudaFuncSetAttribute is PER DEVICE, so the "already configured" cache of
// every launcher is keyed by the current device (a process may drive several GPUs).
struct insmos_smem_cfg_t { size_t v[16] = {0}; };
template <class Kern>
static inline cudaError_t insmos_ensure_smem(Kern kern, size_t smem, insmos_smem_cfg_t& cfg) {
    if (smem <= 48 * 1024) return cudaSuccess;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    size_t& have = cfg.v[dev & 15];
    if (smem > have) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) have = smem;
    }
    return e;
}

// ---- 64-bit key packing: c0,c1,c2 16 bit (biased 32768), c3 8 bit (biased 128), batch 8 bit ----
__device__ __forceinline__ bool coord_in_range(int b, int c0, int c1, int c2, int c3) {
    return (unsigned)b < 255u && (unsigned)(c0 + 32768) < 65536u && (unsigned)(c1 + 32768) < 65536u &&
           (unsigned)(c2 + 32768) < 65536u && (unsigned)(c3 + 128) < 256u;
}
__device__ __forceinline__ uint64_t pack_key(int b, int c0, int c1, int c2, int c3) {
    return ((uint64_t)(uint32_t)(c0 + 32768)) | ((uint64_t)(uint32_t)(c1 + 32768) << 16) |
           ((uint64_t)(uint32_t)(c2 + 32768) << 32) | ((uint64_t)(uint32_t)(c3 + 128) << 48) |
           ((uint64_t)(uint32_t)b << 56);
}
// 64-bit key -> 32-bit slot hash, all 32-bit integer ops (the rule-book build is instruction bound on this:
// profiles/r01_conv_v2_sass_notes.md).  Two odd multipliers on the halves + xorshift-multiply finalisers.
__device__ __forceinline__ uint64_t hash64(uint64_t k) {
    uint32_t lo = (uint32_t)k, hi = (uint32_t)(k >> 32);
    uint32_t h = lo * 0x9E3779B1u ^ hi * 0x85EBCA77u;
    h ^= h >> 15; h *= 0x2C1B3C6Du;
    h ^= h >> 12; h *= 0x297A2D39u;
    h ^= h >> 15;
    return (uint64_t)h;
}

// Home slot of a voxel key.  A SPATIALLY BLOCKED variant (the 4 x 4 x 2 voxels of a block share one hashed 32-slot region)
// was measured on B200 and lost: the map builds went 1.33 -> 1.48 ms per forward -- the longer probe chains of the
// clustered regions cost more than the better line reuse gains (kept behind INSMOS_BLOCKED_HASH for the record).
__device__ __forceinline__ uint64_t home_slot(uint64_t key, uint64_t mask) {
#ifdef INSMOS_BLOCKED_HASH
    const uint64_t local = (key & 3ull) | (((key >> 16) & 3ull) << 2) | (((key >> 32) & 1ull) << 4);
    const uint64_t block = key & ~(3ull | (3ull << 16) | (1ull << 32));
    return ((hash64(block) << 5) | local) & mask;
#else
    return hash64(key) & mask;
#endif
}

// insert (or find) key; returns slot index. Table load factor <= 0.5 guarantees termination.
__device__ __forceinline__ int64_t table_insert(insmos_slot_t* table, uint64_t mask, uint64_t key) {
    uint64_t slot = home_slot(key, mask);
    while (true) {
        unsigned long long* kp = reinterpret_cast<unsigned long long*>(&table[slot].key);
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(kp);
        if (cur == key) return (int64_t)slot;
        if (cur == INSMOS_EMPTY_KEY) {
            unsigned long long prev = atomicCAS(kp, INSMOS_EMPTY_KEY, (unsigned long long)key);
            if (prev == INSMOS_EMPTY_KEY || prev == key) return (int64_t)slot;
        }
        slot = (slot + 1) & mask;
    }
}

// read-only lookup; returns row id or -1. One 16-byte load per probe step.
__device__ __forceinline__ int table_find_row(const insmos_slot_t* __restrict__ table, uint64_t mask, uint64_t key) {
    uint64_t slot = home_slot(key, mask);
    while (true) {
        const int4 v = __ldg(reinterpret_cast<const int4*>(&table[slot]));
        const uint64_t k = ((uint64_t)(uint32_t)v.y << 32) | (uint32_t)v.x;
        if (k == key) return v.w;
        if (k == INSMOS_EMPTY_KEY) return -1;
        slot = (slot + 1) & mask;
    }
}

__device__ __forceinline__ int floor_div(int c, int q) {
    return (c >= 0) ? (c / q) : -((-c + q - 1) / q);
}

// ---- block-wide exclusive scan of one 64-bit value per thread (blockDim.x multiple of 32, <= 1024) ----
// Every thread receives its exclusive prefix; *total (per-thread variable) receives the block total.
__device__ __forceinline__ unsigned long long block_exclusive_scan_u64(unsigned long long v, unsigned long long* total) {
    __shared__ unsigned long long warp_sums[32];
    __shared__ unsigned long long s_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = (lane < nwarps) ? warp_sums[lane] : 0ull;
        unsigned long long winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        warp_sums[lane] = winc - w;                 // exclusive prefix of warp sums
        if (lane == 31) s_total = winc;             // lanes >= nwarps contribute 0
    }
    __syncthreads();
    const unsigned long long r = warp_sums[warp] + inc - v;
    if (total) *total = s_total;
    __syncthreads();                                // shared scratch reusable by the next call
    return r;
}

#define INSMOS_SCAN_BLOCK 1024

