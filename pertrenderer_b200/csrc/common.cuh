// Shared device helpers of the perturbed shading kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/pertshade.h"

namespace pert {

constexpr int NT = 128;  // threads per CTA of the stand-alone operator kernels
constexpr int NW = NT / 32;
// The fused shader kernels run ONE WARP per CTA: tiles differ a lot in cost (sparse fragments), and a
// CTA only gives its shared memory back when its slowest warp is done; with one-warp CTAs the hardware
// CTA scheduler balances the load per tile.
constexpr int FNT = 32;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_prod(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v *= __shfl_xor_sync(FULL, v, o);
    return v;
}
// (value, index) max with the FIRST index winning ties
__device__ __forceinline__ void warp_argmax(float& v, int& i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(FULL, v, o);
        const int oi = __shfl_xor_sync(FULL, i, o);
        if (ov > v || (ov == v && oi < i)) {
            v = ov;
            i = oi;
        }
    }
}

// ---- reductions over an aligned group of G lanes (G a power of two, warp-uniform; a compile-time
// constant in the production instantiations, so these loops unroll) -------------------------------
__device__ __forceinline__ float group_sum(float v, int G) {
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ int group_sum_i(int v, int G) {
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ int group_min_i(int v, int G) {
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ float group_prod(float v, int G) {
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) v *= __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ void group_argmax(float& v, int& i, int G) {
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(FULL, v, o);
        const int oi = __shfl_xor_sync(FULL, i, o);
        if (ov > v || (ov == v && oi < i)) {
            v = ov;
            i = oi;
        }
    }
}

__host__ __device__ __forceinline__ int pow2_ceil(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}
__host__ __device__ __forceinline__ int pow2_floor(int v) {
    int p = 1;
    while (p * 2 <= v) p <<= 1;
    return p;
}

// warp-aggregated append of `flag`-ed items to a shared work list (CTA-shared counter)
__device__ __forceinline__ void list_append(bool flag, uint16_t item, uint16_t* list, int* count) {
    const unsigned b = __ballot_sync(FULL, flag);
    if (b == 0) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == __ffs(b) - 1) base = atomicAdd(count, __popc(b));
    base = __shfl_sync(FULL, base, __ffs(b) - 1);
    if (flag) list[base + __popc(b & ((1u << lane) - 1u))] = item;
}

struct Carver {
    unsigned char* base;
    unsigned off;
    __device__ explicit Carver(unsigned char* b) : base(b), off(0) {}
    template <typename T>
    __device__ T* take(int n) {
        T* r = reinterpret_cast<T*>(base + off);
        off += ((unsigned)n * (unsigned)sizeof(T) + 15u) & ~15u;
        return r;
    }
};
static inline size_t carve(size_t n, size_t sz) { return ((n * sz) + 15) & ~(size_t)15; }

__device__ __forceinline__ void store_winner(void* winners, int win_bytes, int64_t idx, int v) {
    if (win_bytes == 1) reinterpret_cast<uint8_t*>(winners)[idx] = (uint8_t)v;
    else reinterpret_cast<uint16_t*>(winners)[idx] = (uint16_t)v;
}
__device__ __forceinline__ int load_winner(const void* winners, int win_bytes, int64_t idx) {
    return win_bytes == 1 ? (int)reinterpret_cast<const uint8_t*>(winners)[idx]
                          : (int)reinterpret_cast<const uint16_t*>(winners)[idx];
}

// Geometry of one launch of the fused kernels.  One WARP owns a tile of `tp` consecutive pixels
// (tp*K fragment entries are one contiguous run of every (P,K) tensor); G = 32/tp lanes work on each
// pixel of the tile in the per-pixel phases.
struct Launch {
    int tp;         // pixels per warp tile: 4, 8, 16 or 32
    int G;          // lanes per pixel = 32 / tp
    int gshift;     // log2(G)
    int64_t P;      // pixels in this call
    int64_t HW;     // pixels per batch element
    int64_t ntiles; // ceil(P / tp)
    int win_bytes;  // 1 or 2
    int sa_loc;     // local aggregation samples (s_agg_end - s_agg_begin)
    int sc;         // backward: samples per shared-memory chunk (multiple of 32)
    int nchunks;    // backward: ceil(sa_loc / sc)
    int lpe_r, lpe_r_shift;  // coverage sampling: lanes per entry = min(32, pow2_ceil(local quads))
    int lpe_a, lpe_a_shift;  // argmax sampling: lanes per pixel = min(32, pow2_ceil(local quads))
    int lpp, lpp_shift;      // backward: lanes per (pixel, logit) pair = min(8, pow2_floor(local quads))
    int warp_smem;  // bytes of shared memory per warp
    int vec_ok;     // tile rows are 16-byte aligned in every (P,K) tensor
    float invK;     // 1/K for the entry -> pixel division
};

// 16-byte asynchronous global -> shared copy (LDGSTS) and its completion wait
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// entry index within a tile -> pixel of the tile.  Exact for e < 2^16, K < 2^10: (e + .5)/K is at
// least .5/K away from an integer, far more than the fp32 rounding of the product.
__device__ __forceinline__ int entry_pixel(int e, float invK) { return __float2int_rz(((float)e + 0.5f) * invK); }

}  // namespace pert
