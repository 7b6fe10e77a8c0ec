// Stand-alone perturbed operators for mixed operator pairs (e.g. GaussianRast + SoftAgg), where the
// fused shader kernels do not apply: pert_rast_*, pert_argmax_*, and pert_noise_fill (test aid).
//   randomHeaviside  randomras/smoothrast.py:12-59
//   randomArgmax     randomras/smoothagg.py:10-73
#include "kernels.h"
#include "philox.cuh"

namespace pert {

namespace {
struct DevCarver {  // carve dynamic shared memory into 16-byte aligned arrays
    unsigned char* p;
    __device__ explicit DevCarver(unsigned char* base) : p(base) {}
    template <typename T>
    __device__ T* take(int n) {
        T* r = reinterpret_cast<T*>(p);
        p += (((size_t)n * sizeof(T)) + 15) & ~(size_t)15;
        return r;
    }
};
}  // namespace

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

template <class NoiseT, int SCORE>
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
    const bool no_vr = flags & PERT_F_NO_VR;  // without the control variate every sample with h = 1 contributes
    const bool no_skip = (flags & PERT_F_NO_SKIP) || no_vr;
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
                            if (no_vr) {
                                if (h) r += noise_score<SCORE>(nz[t]);
                            } else if (h != h0) {
                                r += h ? noise_score<SCORE>(nz[t]) : -noise_score<SCORE>(nz[t]);
                            }
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
    DevCarver cv(smem_raw);
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

template <class NoiseT, int SCORE>
__global__ void __launch_bounds__(NT) argmax_bwd_kernel(const float* grad_l, const float* z, const void* winners,
                                                        int64_t P, int K1, int S, int s_begin, int s_end, float gamma,
                                                        uint32_t flags, int win_bytes, const NoiseT noise, float* grad_z,
                                                        float* partials) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t pixel = (int64_t)blockIdx.x * NW + warp;
    __shared__ float red[NW];
    DevCarver cv(smem_raw);
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
        const float g0 = (flags & PERT_F_NO_VR) ? 0.0f : gl[a0];  // no control variate: c_s = g[a_s]
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
                            const float v0 = c4.x * noise_score<SCORE>(nz[0]), v1 = c4.y * noise_score<SCORE>(nz[1]),
                                        v2 = c4.z * noise_score<SCORE>(nz[2]), v3 = c4.w * noise_score<SCORE>(nz[3]);
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

template <class NoiseT>
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
    NoiseT noise(seed, stage, pixel_offset);
    float n[4];
    noise.get4(q, slot, pixel, n);
    for (int t = 0; t < 4; ++t) {
        const int s = q * 4 + t;
        if (s >= s_begin && s < s_end) out[((int64_t)(s - s_begin) * P + pixel) * slots + slot] = n[t];
    }
}
int launch_rast_fwd(const float* x, int64_t P, int K, int S, int s_begin, int s_end, float sigma, uint64_t seed,
                    int64_t pixel_offset, const float* noise, uint32_t flags, float* prob, float* rsum, cudaStream_t st) {
    const int64_t n = P * K;
    const unsigned blocks = (unsigned)((n + RAST_TILE - 1) / RAST_TILE);
    const bool cauchy = flags & PERT_F_CAUCHY;
    if (noise) {
        ExplicitNoise xn{noise, P, K, S};
        if (cauchy) rast_fwd_kernel<ExplicitNoise, 1><<<blocks, NT, 0, st>>>(x, n, K, S, s_begin, s_end, sigma, flags, xn, prob, rsum);
        else rast_fwd_kernel<ExplicitNoise, 0><<<blocks, NT, 0, st>>>(x, n, K, S, s_begin, s_end, sigma, flags, xn, prob, rsum);
    } else if (cauchy) {
        PhiloxCauchy pn(seed, 0, pixel_offset);
        rast_fwd_kernel<PhiloxCauchy, 1><<<blocks, NT, 0, st>>>(x, n, K, S, s_begin, s_end, sigma, flags, pn, prob, rsum);
    } else {
        PhiloxNoise pn(seed, 0, pixel_offset);
        rast_fwd_kernel<PhiloxNoise, 0><<<blocks, NT, 0, st>>>(x, n, K, S, s_begin, s_end, sigma, flags, pn, prob, rsum);
    }
    return (int)cudaGetLastError();
}

int launch_rast_bwd(const float* grad_l, const float* rsum, int64_t n, int S, float sigma, float* grad_x,
                    float* partials, float* grad_sigma, cudaStream_t st) {
    const unsigned blocks = (unsigned)((n + 255) / 256);
    rast_bwd_kernel<<<blocks, 256, 0, st>>>(grad_l, rsum, n, 1.0f / ((float)S * sigma), grad_x, partials);
    finalize_single_kernel<<<1, 256, 0, st>>>(partials, blocks, grad_sigma);
    return (int)cudaGetLastError();
}

int launch_argmax_fwd(const float* z, int64_t P, int K1, int S, int s_begin, int s_end, float gamma, uint64_t seed,
                      int64_t pixel_offset, const float* noise, uint32_t flags, float* weights, void* winners,
                      cudaStream_t st) {
    const unsigned blocks = (unsigned)((P + NW - 1) / NW);
    const size_t smem = carve((size_t)NW * K1, 4) * 2 + carve((size_t)NW * K1, 2);
    const int wb = (K1 <= 256) ? 1 : 2;
    if (noise) {
        ExplicitNoise xn{noise, P, K1, S};
        argmax_fwd_kernel<ExplicitNoise><<<blocks, NT, smem, st>>>(z, P, K1, S, s_begin, s_end, gamma, flags, wb, xn, weights, winners);
    } else if (flags & PERT_F_CAUCHY) {
        PhiloxCauchy pn(seed, 1, pixel_offset);
        argmax_fwd_kernel<PhiloxCauchy><<<blocks, NT, smem, st>>>(z, P, K1, S, s_begin, s_end, gamma, flags, wb, pn, weights, winners);
    } else if (flags & PERT_F_UNIFORM) {
        PhiloxUniform pn(seed, 1, pixel_offset);
        argmax_fwd_kernel<PhiloxUniform><<<blocks, NT, smem, st>>>(z, P, K1, S, s_begin, s_end, gamma, flags, wb, pn, weights, winners);
    } else if (flags & PERT_F_GUMBEL) {
        PhiloxGumbel pn(seed, 1, pixel_offset);
        argmax_fwd_kernel<PhiloxGumbel><<<blocks, NT, smem, st>>>(z, P, K1, S, s_begin, s_end, gamma, flags, wb, pn, weights, winners);
    } else {
        PhiloxNoise pn(seed, 1, pixel_offset);
        argmax_fwd_kernel<PhiloxNoise><<<blocks, NT, smem, st>>>(z, P, K1, S, s_begin, s_end, gamma, flags, wb, pn, weights, winners);
    }
    return (int)cudaGetLastError();
}

int launch_argmax_bwd(const float* grad_l, const float* z, const void* winners, int64_t P, int K1, int S, int s_begin,
                      int s_end, float gamma, uint64_t seed, int64_t pixel_offset, const float* noise, uint32_t flags,
                      float* grad_z, float* partials, float* grad_gamma, cudaStream_t st) {
    const unsigned blocks = (unsigned)((P + NW - 1) / NW);
    const size_t smem = carve((size_t)NW * K1, 4) + carve((size_t)NW * 128, 4);
    const int wb = (K1 <= 256) ? 1 : 2;
    const bool cauchy = flags & PERT_F_CAUCHY;
    if (noise) {
        ExplicitNoise xn{noise, P, K1, S};
        if (cauchy) argmax_bwd_kernel<ExplicitNoise, 1><<<blocks, NT, smem, st>>>(grad_l, z, winners, P, K1, S, s_begin, s_end, gamma, flags, wb, xn, grad_z, partials);
        else argmax_bwd_kernel<ExplicitNoise, 0><<<blocks, NT, smem, st>>>(grad_l, z, winners, P, K1, S, s_begin, s_end, gamma, flags, wb, xn, grad_z, partials);
    } else if (cauchy) {
        PhiloxCauchy pn(seed, 1, pixel_offset);
        argmax_bwd_kernel<PhiloxCauchy, 1><<<blocks, NT, smem, st>>>(grad_l, z, winners, P, K1, S, s_begin, s_end, gamma, flags, wb, pn, grad_z, partials);
    } else {
        PhiloxNoise pn(seed, 1, pixel_offset);
        argmax_bwd_kernel<PhiloxNoise, 0><<<blocks, NT, smem, st>>>(grad_l, z, winners, P, K1, S, s_begin, s_end, gamma, flags, wb, pn, grad_z, partials);
    }
    finalize_single_kernel<<<1, 256, 0, st>>>(partials, blocks, grad_gamma);
    return (int)cudaGetLastError();
}

// splitmix64 step (Steele, Lea, Flood 2014) of the two device-side seeds
__global__ void seed_advance_kernel(uint64_t* s) {
    for (int i = 0; i < 2; ++i) {
        uint64_t z = (s[i] += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        s[i] = z ^ (z >> 31);
    }
}
int launch_seed_advance(uint64_t* seed_device, cudaStream_t st) {
    seed_advance_kernel<<<1, 1, 0, st>>>(seed_device);
    return (int)cudaGetLastError();
}

int launch_noise_fill(uint64_t seed, int stage, int64_t P, int slots, int s_begin, int s_end, int64_t pixel_offset,
                      float* out, cudaStream_t st) {
    const int qn = ((s_end + 3) >> 2) - (s_begin >> 2);
    const int64_t total = (int64_t)qn * P * slots;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    if (stage & 4) noise_fill_kernel<PhiloxUniform><<<blocks, 256, 0, st>>>(seed, stage & 1, P, slots, s_begin, s_end, pixel_offset, out);
    else if (stage & 8) noise_fill_kernel<PhiloxGumbel><<<blocks, 256, 0, st>>>(seed, stage & 1, P, slots, s_begin, s_end, pixel_offset, out);
    else if (stage & 2) noise_fill_kernel<PhiloxCauchy><<<blocks, 256, 0, st>>>(seed, stage & 1, P, slots, s_begin, s_end, pixel_offset, out);
    else if (stage & 16) noise_fill_kernel<PhiloxNoise7><<<blocks, 256, 0, st>>>(seed, stage & 1, P, slots, s_begin, s_end, pixel_offset, out);
    else noise_fill_kernel<PhiloxNoise><<<blocks, 256, 0, st>>>(seed, stage & 1, P, slots, s_begin, s_end, pixel_offset, out);
    return (int)cudaGetLastError();
}

}  // namespace pert
