// Fused perturbed shading kernels for sm_100a (B200).  C ABI in include/pertshade.h.
//
// Path replaced (reference = quentinll/pertrenderer): RandomSimpleShader.forward ->
// smooth_rgb_blend (randomras/random_rasterizer.py:34-56) -> GaussianRast.rasterize
// (randomras/smoothrast.py:144-147, randomHeaviside :12-59) -> GaussianAgg.aggregate
// (randomras/smoothagg.py:196-205, randomArgmax :10-73) and the autograd backward of that chain.
//
// Design: one CTA owns a tile of `tp` consecutive pixels (tp*K fragment entries
// are one contiguous run of every (N,H,W,K) input).  The tile is staged in shared memory, the
// Monte-Carlo work of the tile is compacted into work lists (only entries / pixels / samples whose
// noise can change an output bit are drawn) and spread over all lanes as (entry, sample-quad)
// items; noise comes from Philox counters (philox.cuh), so nothing sample-sized touches HBM except
// one winner index per pixel and sample.  Integer histograms and fixed-order shuffle / tree
// reductions make every output deterministic.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/pertshade.h"
#include "philox.cuh"

namespace pert {

constexpr int NT = 128;  // threads per CTA
constexpr int NW = NT / 32;
constexpr unsigned FULL = 0xffffffffu;

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
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
__device__ __forceinline__ int pow2_ceil(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// warp-aggregated append of `flag`-ed items to a shared work list
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
    unsigned char* p;
    __device__ explicit Carver(unsigned char* base) : p(base) {}
    template <typename T>
    __device__ T* take(int n) {
        T* r = reinterpret_cast<T*>(p);
        p += (((size_t)n * sizeof(T)) + 15) & ~(size_t)15;
        return r;
    }
};
static inline size_t carve(size_t n, size_t sz) { return ((n * sz) + 15) & ~(size_t)15; }

struct Launch {
    int tp;          // pixels per tile
    int64_t P;       // pixels in this call
    int64_t HW;      // pixels per batch element
    int win_bytes;   // 1 or 2
    int sa_loc;      // local aggregation samples (s_agg_end - s_agg_begin)
    int sc;          // backward: samples per shared-memory chunk (multiple of 4)
};

// ------------------------------------------------------------------------------------------------
// stage 1 on a tile: coverage samples of the listed entries
//   randomras/smoothrast.py:32-36: h = 1[x + sigma*U >= 0] (rounded multiply, rounded add), mean over s
//   saved for backward: cnt = sum_s h, rs = sum_s (h - h0) U   (smoothrast.py:46)
// ------------------------------------------------------------------------------------------------
template <class NoiseT>
__device__ __forceinline__ void rast_sample_list(const NoiseT& noise, const uint16_t* list, int nlist, const float* xs,
                                                 uint16_t* cnt, float* rs, int K, int64_t pix0, float sigma,
                                                 int s_begin, int s_end) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qb = s_begin >> 2, qe = (s_end + 3) >> 2;
    const int lpe = min(32, pow2_ceil(qe - qb));  // lanes per entry
    const int gpw = 32 / lpe;                     // entries per warp pass
    const int lig = lane & (lpe - 1);
    for (int base = warp * gpw; base < nlist; base += NW * gpw) {
        const int e = base + lane / lpe;
        const bool active = e < nlist;
        const int i = active ? list[e] : 0;
        const int pix = i / K, k = i - pix * K;
        const float x = xs[i];
        const bool h0 = x >= 0.0f;
        int c = 0;
        float r = 0.0f;
        if (active) {
            for (int q = qb + lig; q < qe; q += lpe) {
                float n[4];
                noise.get4(q, k, pix0 + pix, n);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int s = q * 4 + t;
                    const bool h = __fadd_rn(x, __fmul_rn(sigma, n[t])) >= 0.0f;
                    if (s >= s_begin && s < s_end) {
                        c += h ? 1 : 0;
                        if (h != h0) r += h ? n[t] : -n[t];
                    }
                }
            }
        }
        for (int o = lpe >> 1; o > 0; o >>= 1) {
            c += __shfl_xor_sync(FULL, c, o);
            r += __shfl_xor_sync(FULL, r, o);
        }
        if (active && lig == 0) {
            cnt[i] = (uint16_t)c;
            rs[i] = r;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// per-pixel preparation shared by forward and backward (one warp per pixel)
//   random_rasterizer.py:47-48: P = p_hat*mask, alpha = prod(1-P)
//   smoothagg.py:198-202: zi, zmax, zeta_k = (gamma/alpha) log P_k + zi_k - zmax, zeta_K = eps - zmax
// ------------------------------------------------------------------------------------------------
struct PixelInfo {
    float zmax, zimax, prod_nz, zeta_max;
    int argzi, nzero, a0;
};

__device__ __forceinline__ PixelInfo prep_pixel(int pix, int K, const uint16_t* cnt, const unsigned char* msk,
                                                const float* zr, float* zeta /* [K+1] of this pixel */, float zn,
                                                float zf, int S_rast, float gal, float eps) {
    const int lane = threadIdx.x & 31;
    const int e0 = pix * K;
    float zimax = -CUDART_INF_F;
    int argzi = 0x7fffffff;
    float prod = 1.0f;
    int nzero = 0;
    const float denom = zf - zn;
    for (int k = lane; k < K; k += 32) {
        const float m = msk[e0 + k] ? 1.0f : 0.0f;
        const float pk = ((float)cnt[e0 + k] / (float)S_rast) * m;
        const float om = 1.0f - pk;
        if (om == 0.0f) nzero++; else prod *= om;
        const float zi = __fmul_rn(__fdiv_rn(zf - zr[e0 + k], denom), m);
        zeta[k] = zi;
        if (zi > zimax) {
            zimax = zi;
            argzi = k;
        }
    }
    warp_argmax(zimax, argzi);
    prod = warp_prod(prod);
    nzero = warp_sum_i(nzero);
    const float zmax = fmaxf(zimax, eps);
    float best = __fadd_rn(eps, -zmax);  // background logit
    int a0 = K;
    for (int k = lane; k < K; k += 32) {
        const float m = msk[e0 + k] ? 1.0f : 0.0f;
        const float pk = ((float)cnt[e0 + k] / (float)S_rast) * m;
        const float z = __fadd_rn(__fadd_rn(__fmul_rn(gal, logf(pk)), zeta[k]), -zmax);
        zeta[k] = z;
        if (z > best || (z == best && k < a0)) {
            best = z;
            a0 = k;
        }
    }
    if (lane == 0) zeta[K] = __fadd_rn(eps, -zmax);
    warp_argmax(best, a0);
    PixelInfo pi;
    pi.zmax = zmax;
    pi.zimax = zimax;
    pi.prod_nz = prod;
    pi.nzero = nzero;
    pi.argzi = argzi;
    pi.a0 = a0;
    pi.zeta_max = best;
    return pi;
}

__device__ __forceinline__ void store_winner(void* winners, int win_bytes, int64_t idx, int v) {
    if (win_bytes == 1) reinterpret_cast<uint8_t*>(winners)[idx] = (uint8_t)v;
    else reinterpret_cast<uint16_t*>(winners)[idx] = (uint16_t)v;
}
__device__ __forceinline__ int load_winner(const void* winners, int win_bytes, int64_t idx) {
    return win_bytes == 1 ? (int)reinterpret_cast<const uint8_t*>(winners)[idx]
                          : (int)reinterpret_cast<const uint16_t*>(winners)[idx];
}

// ------------------------------------------------------------------------------------------------
// forward kernel
// ------------------------------------------------------------------------------------------------
struct FwdArgs {
    pert_problem pb;
    Launch L;
    float* image;
    uint16_t* counts;
    float* rsum;
    void* winners;
    int32_t* hist;
};

static size_t fwd_smem_bytes(int tp, int K) {
    const size_t E = (size_t)tp * K, E1 = (size_t)tp * (K + 1);
    return carve(E, 4) * 3 + carve(E1, 4) * 2 + carve(E, 2) * 2 + carve(E1, 2) + carve(E, 1) + carve(tp, 4) * 2 +
           carve(tp, 2) + 64;
}

template <class NoiseR, class NoiseA>
__global__ void __launch_bounds__(NT) shade_fwd_kernel(const FwdArgs a, const NoiseR noise_r, const NoiseA noise_a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const pert_problem& pb = a.pb;
    const int K = pb.K, K1 = K + 1, tp = a.L.tp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t pix0 = (int64_t)blockIdx.x * tp;
    const int npx = (int)min((int64_t)tp, a.L.P - pix0);
    const int E = npx * K;
    const uint32_t flags = pb.flags;
    const bool do_rast = flags & PERT_PH_RAST, do_agg = flags & PERT_PH_AGG, do_blend = flags & PERT_PH_BLEND;

    Carver cv(smem_raw);
    float* xs = cv.take<float>(tp * K);      // x = -dists
    float* zr = cv.take<float>(tp * K);      // raw zbuf
    float* rs = cv.take<float>(tp * K);      // sum_s (h-h0) U
    float* zeta = cv.take<float>(tp * K1);   // logits
    int* hist = cv.take<int>(tp * K1);       // winner histogram
    uint16_t* cnt = cv.take<uint16_t>(tp * K);
    uint16_t* list = cv.take<uint16_t>(tp * K);   // coverage work list (entries)
    uint16_t* live = cv.take<uint16_t>(tp * K1);  // per-pixel list of logits that can win
    unsigned char* msk = cv.take<unsigned char>(tp * K);
    int* nlive = cv.take<int>(tp);
    float* px_alpha = cv.take<float>(tp);
    uint16_t* plist = cv.take<uint16_t>(tp);  // pixels that need aggregation samples
    int* counters = cv.take<int>(4);          // [0] entries listed, [1] pixels listed

    if (tid < 4) counters[tid] = 0;
    __syncthreads();

    // ---- phase 0: stage the tile; decide which entries need coverage samples --------------------
    const int sr_loc = pb.s_rast_end - pb.s_rast_begin;
    const float thr = NoiseR::kBounded ? pb.sigma * kNoiseAbsMax * 1.0001f : CUDART_INF_F;
    const bool no_skip = flags & PERT_F_NO_SKIP;
    {
        const int64_t g0 = pix0 * K;
        const int Eround = (E + 31) & ~31;
        for (int i = tid; i < Eround; i += NT) {
            bool need = false;
            if (i < E) {
                const float x = -pb.dists[g0 + i];
                const bool m = pb.pix_to_face[g0 + i] >= 0;
                xs[i] = x;
                zr[i] = pb.zbuf[g0 + i];
                msk[i] = m ? 1 : 0;
                if (do_rast) {
                    // masked entries never reach an output (P = p_hat*mask, grad * mask); |x| beyond the
                    // largest possible sigma*|U| cannot flip: both are exact, not approximations
                    need = no_skip || (m && fabsf(x) <= thr);
                    if (!need) {
                        cnt[i] = (x >= 0.0f) ? (uint16_t)sr_loc : (uint16_t)0;
                        rs[i] = 0.0f;
                    }
                } else {
                    cnt[i] = a.counts[g0 + i];
                }
            }
            if (do_rast) list_append(need, (uint16_t)i, list, &counters[0]);
        }
    }
    __syncthreads();

    // ---- phase 1: coverage samples ---------------------------------------------------------------
    if (do_rast) {
        rast_sample_list(noise_r, list, counters[0], xs, cnt, rs, K, pix0, pb.sigma, pb.s_rast_begin, pb.s_rast_end);
        __syncthreads();
        const int64_t g0 = pix0 * K;
        for (int i = tid; i < E; i += NT) {
            a.counts[g0 + i] = cnt[i];
            a.rsum[g0 + i] = rs[i];
        }
    }
    if (!do_agg && !do_blend) return;

    // ---- phase 2: per-pixel logits, alpha, live lists --------------------------------------------
    const float gal = pb.gamma / pb.alpha;  // fp32 scalar division, smoothagg.py:201
    const int sa_loc = a.L.sa_loc;
    const float cut = NoiseA::kBounded ? 2.0f * pb.gamma * kNoiseAbsMax * 1.0001f : CUDART_INF_F;
    for (int pix = warp; pix < npx; pix += NW) {
        const int64_t gp = pix0 + pix;
        const int b = pb.depth_len > 1 ? (int)(gp / a.L.HW) : 0;
        const float zn = pb.znear[b], zf = pb.zfar[b];
        float* zt = zeta + pix * K1;
        PixelInfo pi = prep_pixel(pix, K, cnt, msk, zr, zt, zn, zf, pb.S_rast, gal, pb.eps);
        __syncwarp();
        if (lane == 0) px_alpha[pix] = 1.0f - (pi.nzero ? 0.0f : pi.prod_nz);
        if (do_agg) {
            // a logit can win some sample only if zeta_j + gamma*Umax >= zeta_max - gamma*Umax
            const float floor_v = no_skip ? -CUDART_INF_F : pi.zeta_max - cut;
            int n = 0;
            for (int j0 = 0; j0 < K1; j0 += 32) {
                const int j = j0 + lane;
                const float z = j < K1 ? zt[j] : -CUDART_INF_F;
                const bool lv = j < K1 && (no_skip || (z > -CUDART_INF_F && z >= floor_v));
                const unsigned bal = __ballot_sync(FULL, lv);
                if (lv) live[pix * K1 + n + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)j;
                n += __popc(bal);
                if (j < K1) hist[pix * K1 + j] = 0;
            }
            __syncwarp();
            if (n == 1) {
                // a single candidate: every sample picks it, no noise needed
                const int j = live[pix * K1];
                if (lane == 0) hist[pix * K1 + j] = sa_loc;
                for (int s = lane; s < sa_loc; s += 32) store_winner(a.winners, a.L.win_bytes, gp * sa_loc + s, j);
            }
            if (lane == 0) {
                nlive[pix] = n;
                if (n > 1) plist[atomicAdd(&counters[1], 1)] = (uint16_t)pix;
            }
        }
    }
    __syncthreads();

    // ---- phase 3: perturbed argmax samples (smoothagg.py:33-36) ----------------------------------
    if (do_agg) {
        const int np = counters[1];
        const int qb = pb.s_agg_begin >> 2, qe = (pb.s_agg_end + 3) >> 2;
        const int lpe = min(32, pow2_ceil(qe - qb));
        const int gpw = 32 / lpe;
        const int lig = lane & (lpe - 1);
        const float gamma = pb.gamma;
        for (int base = warp * gpw; base < np; base += NW * gpw) {
            const int pe = base + lane / lpe;
            if (pe >= np) continue;
            const int pix = plist[pe];
            const int n = nlive[pix];
            const uint16_t* lv = live + pix * K1;
            const float* zt = zeta + pix * K1;
            const int64_t gp = pix0 + pix;
            for (int q = qb + lig; q < qe; q += lpe) {
                float best[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
                int bi[4] = {0, 0, 0, 0};
                for (int l = 0; l < n; ++l) {
                    const int j = lv[l];
                    const float z = zt[j];
                    float nz[4];
                    noise_a.get4(q, j, gp, nz);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const float v = __fadd_rn(z, __fmul_rn(gamma, nz[t]));
                        if (v > best[t]) {
                            best[t] = v;
                            bi[t] = j;
                        }
                    }
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int s = q * 4 + t;
                    if (s >= pb.s_agg_begin && s < pb.s_agg_end) {
                        atomicAdd(&hist[pix * K1 + bi[t]], 1);
                        store_winner(a.winners, a.L.win_bytes, gp * sa_loc + (s - pb.s_agg_begin), bi[t]);
                    }
                }
            }
        }
        __syncthreads();
        if (a.hist) {
            const int64_t g1 = pix0 * K1;
            for (int i = tid; i < npx * K1; i += NT) a.hist[g1 + i] = hist[i];
        }
    } else {
        const int64_t g1 = pix0 * K1;
        for (int i = tid; i < npx * K1; i += NT) hist[i] = a.hist[g1 + i];
        __syncthreads();
    }
    if (!do_blend) return;

    // ---- phase 4: blend (random_rasterizer.py:50-54) ---------------------------------------------
    for (int pix = warp; pix < npx; pix += NW) {
        const int64_t gp = pix0 + pix;
        float r = 0.f, g = 0.f, bl = 0.f;
        for (int j = lane; j < K1; j += 32) {
            const int hcount = hist[pix * K1 + j];
            if (hcount > 0) {
                const float w = (float)hcount / (float)pb.S_agg;
                if (j < K) {
                    const float* c = pb.colors + (gp * K + j) * 3;
                    r += w * c[0];
                    g += w * c[1];
                    bl += w * c[2];
                } else {
                    r += w * pb.background[0];
                    g += w * pb.background[1];
                    bl += w * pb.background[2];
                }
            }
        }
        r = warp_sum(r);
        g = warp_sum(g);
        bl = warp_sum(bl);
        if (lane == 0) reinterpret_cast<float4*>(a.image)[gp] = make_float4(r, g, bl, px_alpha[pix]);
    }
}

// ------------------------------------------------------------------------------------------------
// backward kernel
// ------------------------------------------------------------------------------------------------
struct BwdArgs {
    pert_problem pb;
    Launch L;
    const float* grad_image;
    const uint16_t* counts;
    const float* rsum;
    const void* winners;
    float* grad_dists;
    float* grad_zbuf;
    float* grad_colors;
    float* partials;
    float* acc;
    float* pixstat;
    const int32_t* hist;
};

static size_t bwd_smem_bytes(int tp, int K, int sc) {
    const size_t E = (size_t)tp * K, E1 = (size_t)tp * (K + 1);
    return carve(E, 4) * 3 + carve(E1, 4) * 4 + carve(E1, 4) + carve(E, 2) + carve(E, 1) + carve((size_t)tp * sc, 4) +
           carve(tp, 4) * 12 + carve(tp, 16) + carve(NW * 4, 4) + 64;
}

template <class NoiseA>
__global__ void __launch_bounds__(NT) shade_bwd_kernel(const BwdArgs a, const NoiseA noise_a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const pert_problem& pb = a.pb;
    const int K = pb.K, K1 = K + 1, tp = a.L.tp, sc = a.L.sc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t pix0 = (int64_t)blockIdx.x * tp;
    const int npx = (int)min((int64_t)tp, a.L.P - pix0);
    const int E = npx * K, E1 = npx * K1;
    const uint32_t flags = pb.flags;
    const bool do_sample = flags & PERT_PH_BWD_SAMPLE, do_finish = flags & PERT_PH_BWD_FINISH;
    const bool skip_dead = flags & PERT_F_SKIP_DEAD_NOISE;
    const bool no_skip = flags & PERT_F_NO_SKIP;
    const int sa_loc = a.L.sa_loc;
    const int wb = a.L.win_bytes;

    Carver cv(smem_raw);
    float* xs = cv.take<float>(tp * K);
    float* zr = cv.take<float>(tp * K);
    float* rs = cv.take<float>(tp * K);
    float* zeta = cv.take<float>(tp * K1);
    float* gsel = cv.take<float>(tp * K1);  // g_j = <G_rgb, colour_j>
    float* accs = cv.take<float>(tp * K1);  // sum_s c_s V_sj
    float* t2s = cv.take<float>(tp * K1);   // sum_s c_s V_sj^2
    int* hist = cv.take<int>(tp * K1);
    uint16_t* cnt = cv.take<uint16_t>(tp * K);
    unsigned char* msk = cv.take<unsigned char>(tp * K);
    float* cs = cv.take<float>(tp * sc);  // c_s of the current sample chunk
    float* px_zmax = cv.take<float>(tp);
    float* px_prod = cv.take<float>(tp);
    float* px_csum = cv.take<float>(tp);
    float* px_zn = cv.take<float>(tp);
    float* px_zf = cv.take<float>(tp);
    int* px_nzero = cv.take<int>(tp);
    int* px_argzi = cv.take<int>(tp);
    int* px_pass = cv.take<int>(tp);
    int* px_a0 = cv.take<int>(tp);
    int* px_active = cv.take<int>(tp);
    int* px_pad0 = cv.take<int>(tp);
    int* px_pad1 = cv.take<int>(tp);
    float4* px_G = cv.take<float4>(tp);
    float* red = cv.take<float>(NW * 4);
    (void)px_pad0;
    (void)px_pad1;

    // ---- phase 0: stage the tile ------------------------------------------------------------------
    {
        const int64_t g0 = pix0 * K;
        for (int i = tid; i < E; i += NT) {
            xs[i] = -pb.dists[g0 + i];
            zr[i] = pb.zbuf[g0 + i];
            msk[i] = pb.pix_to_face[g0 + i] >= 0 ? 1 : 0;
            cnt[i] = a.counts[g0 + i];
            rs[i] = a.rsum[g0 + i];
        }
        for (int i = tid; i < npx; i += NT) px_G[i] = reinterpret_cast<const float4*>(a.grad_image)[pix0 + i];
        for (int i = tid; i < E1; i += NT) {
            hist[i] = 0;
            accs[i] = 0.f;
            t2s[i] = 0.f;
            gsel[i] = 0.f;
        }
    }
    __syncthreads();

    // ---- phase 1: per-pixel logits, histogram of saved winners, g_j, grad_colors --------------------
    const float gal = pb.gamma / pb.alpha;
    for (int pix = warp; pix < npx; pix += NW) {
        const int64_t gp = pix0 + pix;
        const int b = pb.depth_len > 1 ? (int)(gp / a.L.HW) : 0;
        const float zn = pb.znear[b], zf = pb.zfar[b];
        float* zt = zeta + pix * K1;
        PixelInfo pi = prep_pixel(pix, K, cnt, msk, zr, zt, zn, zf, pb.S_rast, gal, pb.eps);
        __syncwarp();
        const float4 G = px_G[pix];
        int* hp = hist + pix * K1;
        // histogram: the unperturbed winner a0 is by far the most frequent, count it with ballots
        int n_a0 = 0;
        for (int s0 = 0; s0 < sa_loc; s0 += 32) {
            const int s = s0 + lane;
            const int w = s < sa_loc ? load_winner(a.winners, wb, gp * sa_loc + s) : -1;
            n_a0 += __popc(__ballot_sync(FULL, w == pi.a0));
            if (w >= 0 && w != pi.a0) atomicAdd(&hp[w], 1);
        }
        if (lane == 0) hp[pi.a0] = n_a0;
        __syncwarp();
        // g_j for logits that were selected at least once (others never enter c_s), and for a0
        float* gp_sel = gsel + pix * K1;
        for (int j = lane; j < K1; j += 32) {
            if (hp[j] > 0 || j == pi.a0) {
                float gj;
                if (j < K) {
                    const float* c = pb.colors + (gp * K + j) * 3;
                    gj = G.x * c[0] + G.y * c[1] + G.z * c[2];
                } else {
                    gj = G.x * pb.background[0] + G.y * pb.background[1] + G.z * pb.background[2];
                }
                gp_sel[j] = gj;
            }
        }
        // grad_colors = w_k * G_rgb (dense tensor: 3K contiguous floats per pixel)
        if (a.grad_colors && do_finish) {
            const float invS = 1.0f / (float)pb.S_agg;
            const int32_t* hg = a.hist ? a.hist + gp * K1 : hp;  // all-shard histogram when sample-sharded
            float* gcs = a.grad_colors + gp * K * 3;
            if (((K * 3) & 1) == 0) {  // every pixel row is 8-byte aligned: 64-bit stores
                float2* gc = reinterpret_cast<float2*>(gcs);
                const int n2 = (K * 3) >> 1;
                for (int e = lane; e < n2; e += 32) {
                    const int f0 = 2 * e, f1 = 2 * e + 1;
                    const int k0 = f0 / 3, c0 = f0 - 3 * k0, k1 = f1 / 3, c1 = f1 - 3 * k1;
                    const float w0 = (float)hg[k0] * invS, w1 = (float)hg[k1] * invS;
                    const float g0v = c0 == 0 ? G.x : (c0 == 1 ? G.y : G.z);
                    const float g1v = c1 == 0 ? G.x : (c1 == 1 ? G.y : G.z);
                    gc[e] = make_float2(w0 * g0v, w1 * g1v);
                }
            } else {
                for (int f = lane; f < K * 3; f += 32) {
                    const int k = f / 3, c = f - 3 * k;
                    gcs[f] = (float)hg[k] * invS * (c == 0 ? G.x : (c == 1 ? G.y : G.z));
                }
            }
        }
        if (lane == 0) {
            px_zmax[pix] = pi.zmax;
            px_prod[pix] = pi.prod_nz;
            px_nzero[pix] = pi.nzero;
            px_argzi[pix] = pi.argzi;
            px_pass[pix] = pi.zimax >= pb.eps ? 1 : 0;
            px_a0[pix] = pi.a0;
            px_csum[pix] = 0.f;
            px_zn[pix] = zn;
            px_zf[pix] = zf;
            // a pixel whose samples all picked a0 has c_s = 0 for every s: no score noise needed
            px_active[pix] = (n_a0 != sa_loc) ? 1 : 0;
        }
    }
    __syncthreads();

    // ---- phase 2: score-function sums (smoothagg.py:51-56) ------------------------------------------
    //   c_s = g[a_s] - g[a_0];  acc_j = sum_s c_s V_sj;  t2_j = sum_s c_s V_sj^2
    if (do_sample) {
        const int qb = pb.s_agg_begin >> 2;
        for (int c0 = 0; c0 < sa_loc; c0 += sc) {  // sample chunks (c0 multiple of 4)
            const int cn = min(sc, sa_loc - c0);
            const int cn4 = (cn + 3) & ~3;
            for (int pix = warp; pix < npx; pix += NW) {
                const int64_t gp = pix0 + pix;
                const float* gp_sel = gsel + pix * K1;
                const float g0v = gp_sel[px_a0[pix]];
                float part = 0.f;
                for (int s = lane; s < cn4; s += 32) {
                    float c = 0.f;
                    if (s < cn) {
                        const int w = load_winner(a.winners, wb, gp * sa_loc + c0 + s);
                        c = gp_sel[w] - g0v;
                    }
                    cs[pix * sc + s] = c;
                    part += c;
                }
                part = warp_sum(part);
                if (lane == 0) px_csum[pix] += part;
            }
            __syncthreads();
            for (int it = tid; it < E1; it += NT) {
                const int pix = it / K1, j = it - pix * K1;
                if (!px_active[pix] && !no_skip) continue;
                if (skip_dead && !(zeta[it] > -CUDART_INF_F)) continue;
                const float4* c4p = reinterpret_cast<const float4*>(cs + pix * sc);
                float acc = accs[it], t2 = t2s[it];
                const int64_t gp = pix0 + pix;
                for (int ql = 0; ql < (cn4 >> 2); ++ql) {
                    const float4 c4 = c4p[ql];
                    if (!no_skip && c4.x == 0.f && c4.y == 0.f && c4.z == 0.f && c4.w == 0.f) continue;
                    float nz[4];
                    noise_a.get4(qb + (c0 >> 2) + ql, j, gp, nz);
                    float cv0 = c4.x * nz[0], cv1 = c4.y * nz[1], cv2 = c4.z * nz[2], cv3 = c4.w * nz[3];
                    acc += (cv0 + cv1) + (cv2 + cv3);
                    t2 += (cv0 * nz[0] + cv1 * nz[1]) + (cv2 * nz[2] + cv3 * nz[3]);
                }
                accs[it] = acc;
                t2s[it] = t2;
            }
            __syncthreads();
        }
        if (!do_finish) {
            // sample-sharded job: publish the partial sums, the caller all-reduces them
            const int64_t g1 = pix0 * K1;
            for (int i = tid; i < E1; i += NT) a.acc[g1 + i] = accs[i];
            for (int pix = warp; pix < npx; pix += NW) {
                float t = 0.f;
                for (int j = lane; j < K1; j += 32) {
                    const bool dead = !(zeta[pix * K1 + j] > -CUDART_INF_F);
                    t += (skip_dead && dead) ? px_csum[pix] : t2s[pix * K1 + j];
                }
                t = warp_sum(t);
                if (lane == 0) {
                    a.pixstat[(pix0 + pix) * 2 + 0] = t;
                    a.pixstat[(pix0 + pix) * 2 + 1] = px_csum[pix];
                }
            }
            return;
        }
    }

    // ---- phase 3: chain rule per pixel --------------------------------------
    float p_sigma = 0.f, p_gamma = 0.f, p_q = 0.f;
    {
        const float invSg = 1.0f / ((float)pb.S_agg * pb.gamma);
        for (int pix = warp; pix < npx; pix += NW) {
            const int64_t gp = pix0 + pix;
            float t2sum, csum;
            if (!do_sample) {
                const int64_t g1 = gp * K1;
                for (int j = lane; j < K1; j += 32) accs[pix * K1 + j] = a.acc[g1 + j];
                t2sum = a.pixstat[gp * 2];
                csum = a.pixstat[gp * 2 + 1];
                __syncwarp();
            } else {
                float t = 0.f;
                for (int j = lane; j < K1; j += 32) {
                    const bool dead = !(zeta[pix * K1 + j] > -CUDART_INF_F);
                    t += (skip_dead && dead) ? px_csum[pix] : t2s[pix * K1 + j];
                }
                t2sum = warp_sum(t);
                csum = px_csum[pix];
            }
            // grad_zeta_j = acc_j / (S gamma);  gzmax = -sum_j grad_zeta_j
            float sg = 0.f;
            for (int j = lane; j < K1; j += 32) sg += accs[pix * K1 + j] * invSg;
            const float gzmax = -warp_sum(sg);
            if (lane == 0) p_gamma += (t2sum - csum) * invSg;
            const float G_a = px_G[pix].w;
            const float denom = px_zf[pix] - px_zn[pix];
            const int nzero = px_nzero[pix];
            const float prod_nz = px_prod[pix];
            const int argzi = px_argzi[pix];
            const bool pass = px_pass[pix];
            const float inv_sr = 1.0f / ((float)pb.S_rast * pb.sigma);
            for (int k = lane; k < K; k += 32) {
                const int i = pix * K + k;
                const float m = msk[i] ? 1.0f : 0.0f;
                const float gz = accs[pix * K1 + k] * invSg;
                const float gzi = gz + ((k == argzi && pass) ? gzmax : 0.f);
                a.grad_zbuf[gp * K + k] = -(gzi * m) / denom;
                const float pk = ((float)cnt[i] / (float)pb.S_rast) * m;
                const float lp = logf(pk);
                if (pk > 0.f) p_q += lp * gz;  // prod_corrected: inf -> 0 on the scalar side
                float gP = 0.f;
                if (pk > 0.f) gP = (gal * gz) / pk;  // log_corrected: 1/0 -> 0
                const float om = 1.0f - pk;
                float excl;
                if (nzero == 0) excl = prod_nz / om;
                else if (nzero == 1) excl = (om == 0.f) ? prod_nz : 0.f;
                else excl = 0.f;
                gP += G_a * excl;
                const float gx = (gP * m) * (rs[i] * inv_sr);
                a.grad_dists[gp * K + k] = -gx;
                p_sigma += gx;
            }
        }
    }
    p_sigma = warp_sum(p_sigma);
    p_gamma = warp_sum(p_gamma);
    p_q = warp_sum(p_q);
    if (lane == 0) {
        red[warp * 4 + 0] = p_sigma;
        red[warp * 4 + 1] = p_gamma;
        red[warp * 4 + 2] = p_q;
    }
    __syncthreads();
    if (tid < 3) {
        float s = 0.f;
        for (int w = 0; w < NW; ++w) s += red[w * 4 + tid];
        a.partials[(int64_t)blockIdx.x * 4 + tid] = s;
    }
}

// deterministic final reduction of per-tile scalar partials (one CTA)
//   out[0] = d/dsigma = sum gx                         (smoothrast.py:57-58)
//   out[1] = d/dgamma = score term + q/alpha           (smoothagg.py:72 and :329-332 through gamma/alpha)
//   out[2] = d/dalpha = -q gamma / alpha^2
__global__ void __launch_bounds__(256) finalize_scalars_kernel(const float* partials, int64_t ntiles, float gamma,
                                                               float alpha, float* out) {
    __shared__ double red[3][256];
    double s0 = 0, s1 = 0, s2 = 0;
    for (int64_t t = threadIdx.x; t < ntiles; t += 256) {
        s0 += partials[t * 4 + 0];
        s1 += partials[t * 4 + 1];
        s2 += partials[t * 4 + 2];
    }
    red[0][threadIdx.x] = s0;
    red[1][threadIdx.x] = s1;
    red[2][threadIdx.x] = s2;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            red[0][threadIdx.x] += red[0][threadIdx.x + o];
            red[1][threadIdx.x] += red[1][threadIdx.x + o];
            red[2][threadIdx.x] += red[2][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double q = red[2][0];
        out[0] = (float)red[0][0];
        out[1] = (float)(red[1][0] + q / (double)alpha);
        out[2] = (float)(-q * (double)gamma / ((double)alpha * (double)alpha));
    }
}

// mode 1: out[0] = sum of column 0 only (stand-alone ops)
__global__ void __launch_bounds__(256) finalize_single_kernel(const float* partials, int64_t n, float* out) {
    __shared__ double red[256];
    double s = 0;
    for (int64_t t = threadIdx.x; t < n; t += 256) s += partials[t];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (float)red[0];
}

// ------------------------------------------------------------------------------------------------
// stand-alone perturbed Heaviside (randomras/smoothrast.py:12-59)
// ------------------------------------------------------------------------------------------------
constexpr int RAST_TILE = 1024;  // entries per CTA

template <class NoiseT>
__global__ void __launch_bounds__(NT) rast_fwd_kernel(const float* x, int64_t n, int K, int S, int s_begin, int s_end,
                                                      float sigma, uint32_t flags, const NoiseT noise, float* prob,
                                                      float* rsum) {
    __shared__ float xs[RAST_TILE];
    __shared__ float rs[RAST_TILE];
    __shared__ uint16_t cnt[RAST_TILE];
    __shared__ uint16_t list[RAST_TILE];
    __shared__ int nlist;
    const int tid = threadIdx.x;
    const int64_t g0 = (int64_t)blockIdx.x * RAST_TILE;
    const int E = (int)min((int64_t)RAST_TILE, n - g0);
    if (tid == 0) nlist = 0;
    __syncthreads();
    const float thr = NoiseT::kBounded ? sigma * kNoiseAbsMax * 1.0001f : CUDART_INF_F;
    const bool no_skip = flags & PERT_F_NO_SKIP;
    const int s_loc = s_end - s_begin;
    for (int i = tid; i < RAST_TILE; i += NT) {
        bool need = false;
        if (i < E) {
            const float v = x[g0 + i];
            xs[i] = v;
            need = no_skip || fabsf(v) <= thr;
            if (!need) {
                cnt[i] = v >= 0.f ? (uint16_t)s_loc : (uint16_t)0;
                rs[i] = 0.f;
            }
        }
        list_append(need, (uint16_t)i, list, &nlist);
    }
    __syncthreads();
    // entries of this tile are (pixel, k) pairs of the flat (P,K) tensor: recover them for the counters
    {
        const int lane = tid & 31, warp = tid >> 5;
        const int qb = s_begin >> 2, qe = (s_end + 3) >> 2;
        const int lpe = min(32, pow2_ceil(qe - qb));
        const int gpw = 32 / lpe;
        const int lig = lane & (lpe - 1);
        const int nl = nlist;
        for (int base = warp * gpw; base < nl; base += NW * gpw) {
            const int e = base + lane / lpe;
            const bool active = e < nl;
            const int i = active ? list[e] : 0;
            const int64_t gi = g0 + i;
            const int64_t pixel = gi / K;
            const int k = (int)(gi - pixel * K);
            const float v = xs[i];
            const bool h0 = v >= 0.f;
            int c = 0;
            float r = 0.f;
            if (active) {
                for (int q = qb + lig; q < qe; q += lpe) {
                    float nz[4];
                    noise.get4(q, k, pixel, nz);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int s = q * 4 + t;
                        const bool h = __fadd_rn(v, __fmul_rn(sigma, nz[t])) >= 0.f;
                        if (s >= s_begin && s < s_end) {
                            c += h ? 1 : 0;
                            if (h != h0) r += h ? nz[t] : -nz[t];
                        }
                    }
                }
            }
            for (int o = lpe >> 1; o > 0; o >>= 1) {
                c += __shfl_xor_sync(FULL, c, o);
                r += __shfl_xor_sync(FULL, r, o);
            }
            if (active && lig == 0) {
                cnt[i] = (uint16_t)c;
                rs[i] = r;
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < E; i += NT) {
        prob[g0 + i] = (float)cnt[i] / (float)S;
        rsum[g0 + i] = rs[i];
    }
}

// grad_x = grad_l * rsum / (S sigma); partial sums of grad_x for sigma.grad (smoothrast.py:53-58)
__global__ void __launch_bounds__(256) rast_bwd_kernel(const float* grad_l, const float* rsum, int64_t n, float inv,
                                                       float* grad_x, float* partials) {
    __shared__ float red[8];
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    float gx = 0.f;
    if (i < n) {
        gx = grad_l[i] * (rsum[i] * inv);
        grad_x[i] = gx;
    }
    gx = warp_sum(gx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = gx;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w];
        partials[blockIdx.x] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// stand-alone perturbed argmax (randomras/smoothagg.py:10-73): one warp per pixel
// ------------------------------------------------------------------------------------------------
template <class NoiseT>
__global__ void __launch_bounds__(NT) argmax_fwd_kernel(const float* z, int64_t P, int K1, int S, int s_begin, int s_end,
                                                        float gamma, uint32_t flags, int win_bytes, const NoiseT noise,
                                                        float* weights, void* winners) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t pixel = (int64_t)blockIdx.x * NW + warp;
    if (pixel >= P) return;
    Carver cv(smem_raw);
    float* zt = cv.take<float>(NW * K1) + warp * K1;
    int* hist = cv.take<int>(NW * K1) + warp * K1;
    uint16_t* live = cv.take<uint16_t>(NW * K1) + warp * K1;
    const int s_loc = s_end - s_begin;
    float zmax = -CUDART_INF_F;
    for (int j = lane; j < K1; j += 32) {
        const float v = z[pixel * K1 + j];
        zt[j] = v;
        hist[j] = 0;
        zmax = fmaxf(zmax, v);
    }
    zmax = warp_max(zmax);
    __syncwarp();
    const bool no_skip = flags & PERT_F_NO_SKIP;
    const float floor_v = (NoiseT::kBounded && !no_skip) ? zmax - 2.0f * gamma * kNoiseAbsMax * 1.0001f : -CUDART_INF_F;
    int n = 0;
    for (int j0 = 0; j0 < K1; j0 += 32) {
        const int j = j0 + lane;
        const bool lv = j < K1 && (no_skip || (zt[j] > -CUDART_INF_F && zt[j] >= floor_v));
        const unsigned bal = __ballot_sync(FULL, lv);
        if (lv) live[n + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)j;
        n += __popc(bal);
    }
    __syncwarp();
    const int qb = s_begin >> 2, qe = (s_end + 3) >> 2;
    for (int q = qb + lane; q < qe; q += 32) {
        float best[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
        int bi[4] = {0, 0, 0, 0};
        if (n > 0) {
            bi[0] = bi[1] = bi[2] = bi[3] = live[0];
        }
        for (int l = 0; l < n; ++l) {
            const int j = live[l];
            const float zj = zt[j];
            float nz[4];
            noise.get4(q, j, pixel, nz);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float v = __fadd_rn(zj, __fmul_rn(gamma, nz[t]));
                if (v > best[t]) {
                    best[t] = v;
                    bi[t] = j;
                }
            }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int s = q * 4 + t;
            if (s >= s_begin && s < s_end) {
                atomicAdd(&hist[bi[t]], 1);
                store_winner(winners, win_bytes, pixel * s_loc + (s - s_begin), bi[t]);
            }
        }
    }
    __syncwarp();
    for (int j = lane; j < K1; j += 32) weights[pixel * K1 + j] = (float)hist[j] / (float)S;
}

template <class NoiseT>
__global__ void __launch_bounds__(NT) argmax_bwd_kernel(const float* grad_l, const float* z, const void* winners,
                                                        int64_t P, int K1, int S, int s_begin, int s_end, float gamma,
                                                        uint32_t flags, int win_bytes, const NoiseT noise, float* grad_z,
                                                        float* partials) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t pixel = (int64_t)blockIdx.x * NW + warp;
    __shared__ float red[NW];
    Carver cv(smem_raw);
    float* gl = cv.take<float>(NW * K1) + warp * K1;
    float* cs = cv.take<float>(NW * 128) + warp * 128;  // chunk of 128 samples
    float p_gamma = 0.f;
    if (pixel < P) {
        const int s_loc = s_end - s_begin;
        float best = -CUDART_INF_F;
        int a0 = 0x7fffffff;
        for (int j = lane; j < K1; j += 32) {
            gl[j] = grad_l[pixel * K1 + j];
            const float v = z[pixel * K1 + j];
            if (v > best) {
                best = v;
                a0 = j;
            }
        }
        warp_argmax(best, a0);
        __syncwarp();
        const float g0 = gl[a0];
        const bool skip_dead = flags & PERT_F_SKIP_DEAD_NOISE;
        const bool no_skip = flags & PERT_F_NO_SKIP;
        const int nj = (K1 + 31) / 32;
        float csum = 0.f, t2 = 0.f;
        const float invSg = 1.0f / ((float)S * gamma);
        // accumulators for up to 8 logits per lane (K1 <= 256); larger K1 loops in passes
        for (int jpass = 0; jpass < nj; jpass += 8) {
            float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int c0 = 0; c0 < s_loc; c0 += 128) {
                const int cn = min(128, s_loc - c0), cn4 = (cn + 3) & ~3;
                __syncwarp();
                for (int s = lane; s < cn4; s += 32) {
                    float c = 0.f;
                    if (s < cn) c = gl[load_winner(winners, win_bytes, pixel * s_loc + c0 + s)] - g0;
                    cs[s] = c;
                    if (jpass == 0) csum += c;
                }
                __syncwarp();
                for (int ql = 0; ql < (cn4 >> 2); ++ql) {
                    const float4 c4 = reinterpret_cast<const float4*>(cs)[ql];
                    if (!no_skip && c4.x == 0.f && c4.y == 0.f && c4.z == 0.f && c4.w == 0.f) continue;
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int j = (jpass + u) * 32 + lane;
                        if (j < K1) {
                            if (skip_dead && !(z[pixel * K1 + j] > -CUDART_INF_F)) {
                                t2 += (c4.x + c4.y) + (c4.z + c4.w);
                                continue;
                            }
                            float nz[4];
                            noise.get4((s_begin >> 2) + (c0 >> 2) + ql, j, pixel, nz);
                            const float v0 = c4.x * nz[0], v1 = c4.y * nz[1], v2 = c4.z * nz[2], v3 = c4.w * nz[3];
                            acc[u] += (v0 + v1) + (v2 + v3);
                            t2 += (v0 * nz[0] + v1 * nz[1]) + (v2 * nz[2] + v3 * nz[3]);
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = (jpass + u) * 32 + lane;
                if (j < K1) grad_z[pixel * K1 + j] = acc[u] * invSg;
            }
        }
        csum = warp_sum(csum);
        t2 = warp_sum(t2);
        p_gamma = (t2 - csum) * invSg;
    }
    if (lane == 0) red[warp] = p_gamma;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < NW; ++w) s += red[w];
        partials[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256) noise_fill_kernel(uint64_t seed, int stage, int64_t P, int slots, int s_begin,
                                                         int s_end, int64_t pixel_offset, float* out) {
    // one thread per (quad, pixel, slot)
    const int64_t per_q = P * slots;
    const int qb = s_begin >> 2, qe = (s_end + 3) >> 2;
    const int64_t total = (int64_t)(qe - qb) * per_q;
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= total) return;
    const int q = qb + (int)(idx / per_q);
    const int64_t rem = idx - (int64_t)(q - qb) * per_q;
    const int64_t pixel = rem / slots;
    const int slot = (int)(rem - pixel * slots);
    PhiloxNoise noise(seed, stage, pixel_offset);
    float n[4];
    noise.get4(q, slot, pixel, n);
    for (int t = 0; t < 4; ++t) {
        const int s = q * 4 + t;
        if (s >= s_begin && s < s_end) out[((int64_t)(s - s_begin) * P + pixel) * slots + slot] = n[t];
    }
}

}  // namespace pert

// ================================================================================================
// C ABI
// ================================================================================================
using namespace pert;

static thread_local const char* g_last_cuda = "none";

static int cuda_fail(cudaError_t e) {
    g_last_cuda = cudaGetErrorName(e);
    return PERT_E_CUDA;
}

extern "C" int pert_version(void) { return PERT_ABI_VERSION; }

extern "C" const char* pert_last_cuda_error(void) { return g_last_cuda; }

extern "C" const char* pert_strerror(int code) {
    switch (code) {
        case PERT_OK: return "ok";
        case PERT_E_NULL: return "required pointer is NULL";
        case PERT_E_SHAPE: return "bad shape";
        case PERT_E_UNSUPPORTED: return "K or S not supported by the kernels";
        case PERT_E_ALIGN: return "pointer not aligned";
        case PERT_E_SAMPLES: return "bad sample shard (begin must be a multiple of 4, begin < end <= S)";
        case PERT_E_CUDA: return "CUDA error";
        case PERT_E_SCALAR: return "sigma, gamma and alpha must be finite and > 0";
        default: return "unknown error";
    }
}

static int pick_tp(int K) {
    // ~1024 fragment entries per tile, a multiple of 4 pixels so every tile starts 16-byte aligned
    int tp = 1024 / (K > 0 ? K : 1);
    tp &= ~3;
    if (tp < 4) tp = 4;
    if (tp > 32) tp = 32;
    return tp;
}

static int check_problem(const pert_problem* pb) {
    if (!pb) return PERT_E_NULL;
    if (pb->N <= 0 || pb->H <= 0 || pb->W <= 0 || pb->K <= 0) return PERT_E_SHAPE;
    if (pb->K > 1023) return PERT_E_UNSUPPORTED;
    if (pb->S_rast <= 0 || pb->S_agg <= 0 || pb->S_rast > 65535 || pb->S_agg > (1 << 24)) return PERT_E_UNSUPPORTED;
    if (pb->depth_len != 1 && pb->depth_len != pb->N) return PERT_E_SHAPE;
    if (!(pb->sigma > 0.f) || !(pb->gamma > 0.f) || !(pb->alpha > 0.f) || isinf(pb->sigma) || isinf(pb->gamma) ||
        isinf(pb->alpha))
        return PERT_E_SCALAR;
    if ((pb->s_rast_begin & 3) || pb->s_rast_begin < 0 || pb->s_rast_begin >= pb->s_rast_end || pb->s_rast_end > pb->S_rast)
        return PERT_E_SAMPLES;
    if ((pb->s_agg_begin & 3) || pb->s_agg_begin < 0 || pb->s_agg_begin >= pb->s_agg_end || pb->s_agg_end > pb->S_agg)
        return PERT_E_SAMPLES;
    if (!pb->pix_to_face || !pb->zbuf || !pb->dists || !pb->znear || !pb->zfar) return PERT_E_NULL;
    if (((uintptr_t)pb->pix_to_face & 7) || ((uintptr_t)pb->zbuf & 3) || ((uintptr_t)pb->dists & 3)) return PERT_E_ALIGN;
    return PERT_OK;
}

extern "C" int64_t pert_num_tiles(const pert_problem* pb) {
    if (!pb || pb->K <= 0) return 0;
    const int tp = pick_tp(pb->K);
    const int64_t P = pb->N * pb->H * pb->W;
    return (P + tp - 1) / tp;
}

extern "C" int pert_winner_bytes(int32_t K) { return (K + 1 <= 256) ? 1 : 2; }

static Launch make_launch(const pert_problem* pb) {
    Launch L;
    L.tp = pick_tp(pb->K);
    L.P = pb->N * pb->H * pb->W;
    L.HW = pb->H * pb->W;
    L.win_bytes = pert_winner_bytes(pb->K);
    L.sa_loc = pb->s_agg_end - pb->s_agg_begin;
    int sc = (L.sa_loc + 3) & ~3;
    if (sc > 256) sc = 256;
    L.sc = sc;
    return L;
}

template <class KernelT>
static int set_smem(KernelT k, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return cuda_fail(e);
    }
    return PERT_OK;
}

extern "C" int pert_shade_fwd(const pert_problem* pb_in, float* image, uint16_t* counts, float* rsum, void* winners,
                              int32_t* hist, void* stream) {
    int rc = check_problem(pb_in);
    if (rc) return rc;
    FwdArgs a;
    a.pb = *pb_in;
    if (!(a.pb.flags & (PERT_PH_RAST | PERT_PH_AGG | PERT_PH_BLEND))) a.pb.flags |= PERT_PH_RAST | PERT_PH_AGG | PERT_PH_BLEND;
    const uint32_t f = a.pb.flags;
    if (!counts) return PERT_E_NULL;
    if ((f & PERT_PH_RAST) && !rsum) return PERT_E_NULL;
    if ((f & PERT_PH_AGG) && !winners) return PERT_E_NULL;
    if ((f & PERT_PH_BLEND) && (!image || !a.pb.colors)) return PERT_E_NULL;
    if (((f & PERT_PH_AGG) != 0) != ((f & PERT_PH_BLEND) != 0) && !hist) return PERT_E_NULL;
    if (((uintptr_t)image & 15) || ((uintptr_t)counts & 1) || ((uintptr_t)rsum & 3) || ((uintptr_t)hist & 3) ||
        ((uintptr_t)a.pb.colors & 3))
        return PERT_E_ALIGN;
    a.L = make_launch(&a.pb);
    if (a.L.win_bytes == 2 && ((uintptr_t)winners & 1)) return PERT_E_ALIGN;
    a.image = image;
    a.counts = counts;
    a.rsum = rsum;
    a.winners = winners;
    a.hist = hist;
    const size_t smem = fwd_smem_bytes(a.L.tp, a.pb.K);
    if (smem > 200 * 1024) return PERT_E_UNSUPPORTED;
    const int64_t tiles = (a.L.P + a.L.tp - 1) / a.L.tp;
    if (tiles > 0x7fffffff) return PERT_E_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const bool er = a.pb.noise_rast != nullptr, ea = a.pb.noise_agg != nullptr;
    PhiloxNoise pr(a.pb.seed_rast, 0, a.pb.pixel_offset), pa(a.pb.seed_agg, 1, a.pb.pixel_offset);
    ExplicitNoise xr{a.pb.noise_rast, a.L.P, a.pb.K, a.pb.S_rast}, xa{a.pb.noise_agg, a.L.P, a.pb.K + 1, a.pb.S_agg};
    if (!er && !ea) {
        if ((rc = set_smem(shade_fwd_kernel<PhiloxNoise, PhiloxNoise>, smem))) return rc;
        shade_fwd_kernel<PhiloxNoise, PhiloxNoise><<<(unsigned)tiles, NT, smem, st>>>(a, pr, pa);
    } else if (er && ea) {
        if ((rc = set_smem(shade_fwd_kernel<ExplicitNoise, ExplicitNoise>, smem))) return rc;
        shade_fwd_kernel<ExplicitNoise, ExplicitNoise><<<(unsigned)tiles, NT, smem, st>>>(a, xr, xa);
    } else if (er) {
        if ((rc = set_smem(shade_fwd_kernel<ExplicitNoise, PhiloxNoise>, smem))) return rc;
        shade_fwd_kernel<ExplicitNoise, PhiloxNoise><<<(unsigned)tiles, NT, smem, st>>>(a, xr, pa);
    } else {
        if ((rc = set_smem(shade_fwd_kernel<PhiloxNoise, ExplicitNoise>, smem))) return rc;
        shade_fwd_kernel<PhiloxNoise, ExplicitNoise><<<(unsigned)tiles, NT, smem, st>>>(a, pr, xa);
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? PERT_OK : cuda_fail(e);
}

extern "C" int pert_shade_bwd(const pert_problem* pb_in, const float* grad_image, const uint16_t* counts,
                              const float* rsum, const void* winners, float* grad_dists, float* grad_zbuf,
                              float* grad_colors, float* scalar_partials, float* grad_scalars, float* acc,
                              float* pixstat, const int32_t* hist, void* stream) {
    int rc = check_problem(pb_in);
    if (rc) return rc;
    BwdArgs a;
    a.pb = *pb_in;
    if (!(a.pb.flags & (PERT_PH_BWD_SAMPLE | PERT_PH_BWD_FINISH))) a.pb.flags |= PERT_PH_BWD_SAMPLE | PERT_PH_BWD_FINISH;
    const uint32_t f = a.pb.flags;
    if (!grad_image || !counts || !rsum || !winners || !a.pb.colors) return PERT_E_NULL;
    if ((f & PERT_PH_BWD_FINISH) && (!grad_dists || !grad_zbuf || !scalar_partials || !grad_scalars)) return PERT_E_NULL;
    if (((f & PERT_PH_BWD_SAMPLE) != 0) != ((f & PERT_PH_BWD_FINISH) != 0) && (!acc || !pixstat)) return PERT_E_NULL;
    if (((uintptr_t)grad_image & 15) || ((uintptr_t)grad_colors & 7) || ((uintptr_t)grad_dists & 3) ||
        ((uintptr_t)grad_zbuf & 3) || ((uintptr_t)counts & 1) || ((uintptr_t)rsum & 3))
        return PERT_E_ALIGN;
    a.L = make_launch(&a.pb);
    a.grad_image = grad_image;
    a.counts = counts;
    a.rsum = rsum;
    a.winners = winners;
    a.grad_dists = grad_dists;
    a.grad_zbuf = grad_zbuf;
    a.grad_colors = grad_colors;
    a.partials = scalar_partials;
    a.acc = acc;
    a.pixstat = pixstat;
    a.hist = hist;
    const size_t smem = bwd_smem_bytes(a.L.tp, a.pb.K, a.L.sc);
    if (smem > 200 * 1024) return PERT_E_UNSUPPORTED;
    const int64_t tiles = (a.L.P + a.L.tp - 1) / a.L.tp;
    if (tiles > 0x7fffffff) return PERT_E_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (a.pb.noise_agg) {
        ExplicitNoise xa{a.pb.noise_agg, a.L.P, a.pb.K + 1, a.pb.S_agg};
        if ((rc = set_smem(shade_bwd_kernel<ExplicitNoise>, smem))) return rc;
        shade_bwd_kernel<ExplicitNoise><<<(unsigned)tiles, NT, smem, st>>>(a, xa);
    } else {
        PhiloxNoise pa(a.pb.seed_agg, 1, a.pb.pixel_offset);
        if ((rc = set_smem(shade_bwd_kernel<PhiloxNoise>, smem))) return rc;
        shade_bwd_kernel<PhiloxNoise><<<(unsigned)tiles, NT, smem, st>>>(a, pa);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e);
    if (f & PERT_PH_BWD_FINISH) {
        finalize_scalars_kernel<<<1, 256, 0, st>>>(scalar_partials, tiles, a.pb.gamma, a.pb.alpha, grad_scalars);
        e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e);
    }
    return PERT_OK;
}

extern "C" int pert_rast_fwd(const float* x, int64_t P, int32_t K, int32_t S, int32_t s_begin, int32_t s_end,
                             float sigma, uint64_t seed, int64_t pixel_offset, const float* noise, uint32_t flags,
                             float* prob, float* rsum, void* stream) {
    if (!x || !prob || !rsum) return PERT_E_NULL;
    if (P <= 0 || K <= 0) return PERT_E_SHAPE;
    if (S <= 0 || S > 65535) return PERT_E_UNSUPPORTED;
    if ((s_begin & 3) || s_begin < 0 || s_begin >= s_end || s_end > S) return PERT_E_SAMPLES;
    if (!(sigma > 0.f) || isinf(sigma)) return PERT_E_SCALAR;
    const int64_t n = P * K;
    const int64_t blocks = (n + RAST_TILE - 1) / RAST_TILE;
    if (blocks > 0x7fffffff) return PERT_E_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (noise) {
        ExplicitNoise xn{noise, P, K, S};
        rast_fwd_kernel<ExplicitNoise><<<(unsigned)blocks, NT, 0, st>>>(x, n, K, S, s_begin, s_end, sigma, flags, xn, prob, rsum);
    } else {
        PhiloxNoise pn(seed, 0, pixel_offset);
        rast_fwd_kernel<PhiloxNoise><<<(unsigned)blocks, NT, 0, st>>>(x, n, K, S, s_begin, s_end, sigma, flags, pn, prob, rsum);
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? PERT_OK : cuda_fail(e);
}

extern "C" int pert_rast_bwd(const float* grad_l, const float* rsum, int64_t n, int32_t S, float sigma, float* grad_x,
                             float* scalar_partials, float* grad_sigma, void* stream) {
    if (!grad_l || !rsum || !grad_x || !scalar_partials || !grad_sigma) return PERT_E_NULL;
    if (n <= 0 || S <= 0) return PERT_E_SHAPE;
    if (!(sigma > 0.f)) return PERT_E_SCALAR;
    const int64_t blocks = (n + 255) / 256;
    if (blocks > 0x7fffffff) return PERT_E_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    rast_bwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(grad_l, rsum, n, 1.0f / ((float)S * sigma), grad_x, scalar_partials);
    finalize_single_kernel<<<1, 256, 0, st>>>(scalar_partials, blocks, grad_sigma);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? PERT_OK : cuda_fail(e);
}

extern "C" int pert_argmax_fwd(const float* z, int64_t P, int32_t K1, int32_t S, int32_t s_begin, int32_t s_end,
                               float gamma, uint64_t seed, int64_t pixel_offset, const float* noise, uint32_t flags,
                               float* weights, void* winners, void* stream) {
    if (!z || !weights || !winners) return PERT_E_NULL;
    if (P <= 0 || K1 <= 0) return PERT_E_SHAPE;
    if (K1 > 1024 || S <= 0 || S > (1 << 24)) return PERT_E_UNSUPPORTED;
    if ((s_begin & 3) || s_begin < 0 || s_begin >= s_end || s_end > S) return PERT_E_SAMPLES;
    if (!(gamma > 0.f) || isinf(gamma)) return PERT_E_SCALAR;
    const int64_t blocks = (P + NW - 1) / NW;
    if (blocks > 0x7fffffff) return PERT_E_UNSUPPORTED;
    const size_t smem = carve((size_t)NW * K1, 4) * 2 + carve((size_t)NW * K1, 2);
    const int wb = pert_winner_bytes(K1 - 1);
    cudaStream_t st = (cudaStream_t)stream;
    if (noise) {
        ExplicitNoise xn{noise, P, K1, S};
        argmax_fwd_kernel<ExplicitNoise><<<(unsigned)blocks, NT, smem, st>>>(z, P, K1, S, s_begin, s_end, gamma, flags, wb, xn, weights, winners);
    } else {
        PhiloxNoise pn(seed, 1, pixel_offset);
        argmax_fwd_kernel<PhiloxNoise><<<(unsigned)blocks, NT, smem, st>>>(z, P, K1, S, s_begin, s_end, gamma, flags, wb, pn, weights, winners);
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? PERT_OK : cuda_fail(e);
}

extern "C" int pert_argmax_bwd(const float* grad_l, const float* z, const void* winners, int64_t P, int32_t K1,
                               int32_t S, int32_t s_begin, int32_t s_end, float gamma, uint64_t seed,
                               int64_t pixel_offset, const float* noise, uint32_t flags, float* grad_z,
                               float* scalar_partials, float* grad_gamma, void* stream) {
    if (!grad_l || !z || !winners || !grad_z || !scalar_partials || !grad_gamma) return PERT_E_NULL;
    if (P <= 0 || K1 <= 0) return PERT_E_SHAPE;
    if (K1 > 1024 || S <= 0) return PERT_E_UNSUPPORTED;
    if ((s_begin & 3) || s_begin < 0 || s_begin >= s_end || s_end > S) return PERT_E_SAMPLES;
    if (!(gamma > 0.f) || isinf(gamma)) return PERT_E_SCALAR;
    const int64_t blocks = (P + NW - 1) / NW;
    if (blocks > 0x7fffffff) return PERT_E_UNSUPPORTED;
    const size_t smem = carve((size_t)NW * K1, 4) + carve((size_t)NW * 128, 4);
    const int wb = pert_winner_bytes(K1 - 1);
    cudaStream_t st = (cudaStream_t)stream;
    if (noise) {
        ExplicitNoise xn{noise, P, K1, S};
        argmax_bwd_kernel<ExplicitNoise><<<(unsigned)blocks, NT, smem, st>>>(grad_l, z, winners, P, K1, S, s_begin, s_end, gamma, flags, wb, xn, grad_z, scalar_partials);
    } else {
        PhiloxNoise pn(seed, 1, pixel_offset);
        argmax_bwd_kernel<PhiloxNoise><<<(unsigned)blocks, NT, smem, st>>>(grad_l, z, winners, P, K1, S, s_begin, s_end, gamma, flags, wb, pn, grad_z, scalar_partials);
    }
    finalize_single_kernel<<<1, 256, 0, st>>>(scalar_partials, blocks, grad_gamma);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? PERT_OK : cuda_fail(e);
}

extern "C" int pert_noise_fill(uint64_t seed, int32_t stage, int64_t P, int32_t slots, int32_t s_begin, int32_t s_end,
                               int64_t pixel_offset, float* out, void* stream) {
    if (!out) return PERT_E_NULL;
    if (P <= 0 || slots <= 0) return PERT_E_SHAPE;
    if ((s_begin & 3) || s_begin < 0 || s_begin >= s_end) return PERT_E_SAMPLES;
    const int qn = ((s_end + 3) >> 2) - (s_begin >> 2);
    const int64_t total = (int64_t)qn * P * slots;
    const int64_t blocks = (total + 255) / 256;
    if (blocks > 0x7fffffff) return PERT_E_UNSUPPORTED;
    noise_fill_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(seed, stage, P, slots, s_begin, s_end, pixel_offset, out);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? PERT_OK : cuda_fail(e);
}
