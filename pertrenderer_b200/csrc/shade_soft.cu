// Fused SoftRas shader: SoftRast + SoftAgg (the DEFAULT operators of RandomSimpleShader), forward and backward.
//
// Reference path: smooth_rgb_blend (randomras/random_rasterizer.py:34-56) with
//   SoftRast.rasterize   randomras/smoothrast.py:132-134   P = sigmoid(-dists / sigma) * mask
//   SoftAgg.aggregate    randomras/smoothagg.py:172-182    w = softmax((1/gamma) * logits), same logits as GaussianAgg
//   log_corrected / prod_corrected backward rules          randomras/smoothagg.py:292-337
// Deterministic (no noise): a pure streaming kernel.  Same tile machinery as the perturbed shader: one warp per
// tile of tp pixels, compact list of valid entries, G = 32/tp lanes per pixel; backward recomputes the forward
// quantities instead of saving them (nothing is stored between the two passes).
#include "kernels.h"
#include "tile.cuh"

namespace pert {

void soft_smem_layout(int tp, int K, bool bwd, SmemLayout& L) {
    const size_t E = (size_t)tp * K;
    Carver cv(L);
    cv.take(E, 2);       // vlist
    cv.take(E, 4);       // ps: coverage probability
    cv.take(E, 4);       // zs: zbuf -> zi
    cv.take(E, 4);       // ys: logits -> exp -> weights
    cv.take(tp + 1, 4);  // vstart
    (void)bwd;
}

struct SoftPix {
    float zmax, zimax, prod_nz, ymax, ybg, zbg;
    int argzi, nzero, kpad;
};

// Shared by forward and backward: P, alpha pieces, zi, logits (randomras/smoothagg.py:174-180), y = (1/gamma) * logit.
// Out: ps[n] = P_k, zs[n] = zi_k, ys[n] = y_k.
__device__ __forceinline__ SoftPix soft_prep(int p, int lig, int G, bool pvalid, int K, const int* vstart,
                                             const uint16_t* vlist, const float* dists_t, const float* zbuf_t, float* ps,
                                             float* zs, float* ys, float zn, float zf, float sigma, float gal, float inv_g,
                                             float eps) {
    const int vs = pvalid ? vstart[p] : 0, ve = pvalid ? vstart[p + 1] : 0;
    const int nv = ve - vs;
    const int e_base = p * K;
    float zimax = -CUDART_INF_F, prod = 1.0f;
    int argzi = 0x7fffffff, nzero = 0, kpad = 0x7fffffff;
    const float denom = zf - zn;
#pragma unroll 1
    for (int n = vs + lig; n < ve; n += G) {
        const int e = vlist[n];
        const int k = e - e_base;
        const float u = __fdiv_rn(-__ldg(dists_t + e), sigma);       // smoothrast.py:133
        const float pk = __fdiv_rn(1.0f, 1.0f + expf(-u));            // torch.sigmoid
        ps[n] = pk;
        const float om = 1.0f - pk;
        if (om == 0.0f) nzero++; else prod *= om;
        const float zi = __fdiv_rn(zf - __ldg(zbuf_t + e), denom);
        zs[n] = zi;
        if (zi > zimax) {
            zimax = zi;
            argzi = k;
        }
        if (k != n - vs) kpad = min(kpad, n - vs);
    }
    __syncwarp();
    group_argmax(zimax, argzi, G);
    prod = group_prod(prod, G);
    nzero = group_sum_i(nzero, G);
    kpad = group_min_i(kpad, G);
    if (kpad == 0x7fffffff) kpad = nv;
    if (kpad < K && (0.0f > zimax || (0.0f == zimax && kpad < argzi))) {  // masked entries have zi = 0
        zimax = 0.0f;
        argzi = kpad;
    }
    const float zmax = fmaxf(zimax, eps);
    const float zbg = __fadd_rn(eps, -zmax);
    const float ybg = __fmul_rn(inv_g, zbg);
    float ymax = ybg;
#pragma unroll 1
    for (int n = vs + lig; n < ve; n += G) {
        const float pk = ps[n];
        float y = -CUDART_INF_F;
        if (pk > 0.0f) {
            const float zeta = __fadd_rn(__fadd_rn(__fmul_rn(gal, logf_exact(pk)), zs[n]), -zmax);
            y = __fmul_rn(inv_g, zeta);
        }
        ys[n] = y;
        ymax = fmaxf(ymax, y);
    }
    __syncwarp();
    for (int o = G >> 1; o > 0; o >>= 1) ymax = fmaxf(ymax, __shfl_xor_sync(FULL, ymax, o));
    SoftPix sp;
    sp.zmax = zmax;
    sp.zimax = zimax;
    sp.prod_nz = prod;
    sp.ymax = ymax;
    sp.ybg = ybg;
    sp.zbg = zbg;
    sp.argzi = argzi;
    sp.nzero = nzero;
    sp.kpad = kpad;
    return sp;
}

struct SoftArgs {
    pert_problem pb;
    Launch L;
    float* image;
    const float* grad_image;
    float* grad_dists;
    float* grad_zbuf;
    float* grad_colors;
    float* partials;
};

template <int GT, bool BWD>
__global__ void __launch_bounds__(FNT, 24) soft_shade_kernel(const SoftArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const pert_problem& pb = a.pb;
    const int lane = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int G = GT ? GT : a.L.G;
    const int gshift = GT ? (GT == 8 ? 3 : GT == 4 ? 2 : GT == 2 ? 1 : 0) : a.L.gshift;
    const int K = pb.K, tp = 32 >> gshift;
    const int64_t pix0 = tile * tp;
    const int npx = (int)min((int64_t)tp, a.L.P - pix0);
    const int E = npx * K;
    const int64_t g0 = pix0 * K;
    const int p = lane >> gshift, lig = lane & (G - 1);
    const bool pvalid = p < npx;
    const int64_t gp = pix0 + p;

    Taker cv(smem_raw, a.L.sm);
    uint16_t* vlist = cv.take<uint16_t>();
    float* ps = cv.take<float>();
    float* zs = cv.take<float>();
    float* ys = cv.take<float>();
    int* vstart = cv.take<int>();

    float4 Gi = make_float4(0.f, 0.f, 0.f, 0.f);
    float zn = 1.0f, zf = 100.0f;
    if (pvalid) {
        if (BWD) Gi = __ldg(reinterpret_cast<const float4*>(a.grad_image) + gp);
        const int b = pb.depth_len > 1 ? batch_of(pix0, p, a.L.HW) : 0;
        zn = __ldg(pb.znear + b);
        zf = __ldg(pb.zfar + b);
    }
    const int nv = scan_valid(pb.pix_to_face + g0, E, a.L.vec_ok, vlist, E);
    if (BWD) {
        zero_fill(a.grad_dists + g0, E, a.L.vec_ok);
        zero_fill(a.grad_zbuf + g0, E, a.L.vec_ok);
        if (a.grad_colors) zero_fill(a.grad_colors + g0 * 3, E * 3, a.L.vec_ok);
    }
    float p_sigma = 0.f, p_a1 = 0.f, p_q = 0.f;
    if (nv == 0) {
        if (!BWD && lane < npx)
            reinterpret_cast<float4*>(a.image)[pix0 + lane] =
                make_float4(pb.background[0], pb.background[1], pb.background[2], 0.0f);
    } else {
        __syncwarp();
        pixel_ranges(vlist, nv, K, tp, vstart);
        __syncwarp();
        const float gal = a.L.gal, inv_g = a.L.inv_gamma, sigma = pb.sigma;
        const float* const dists_t = pb.dists + g0;
        const float* const zbuf_t = pb.zbuf + g0;
        const float* const colors_t = pb.colors + g0 * 3;
        const SoftPix sp = soft_prep(p, lig, G, pvalid, K, vstart, vlist, dists_t, zbuf_t, ps, zs, ys, zn, zf, sigma, gal,
                                     inv_g, pb.eps);
        const int vs = pvalid ? vstart[p] : 0, ve = pvalid ? vstart[p + 1] : 0;
        // softmax over the K+1 logits (masked / zero-probability logits are -inf: weight 0)
        float se = 0.f;
#pragma unroll 1
        for (int n = vs + lig; n < ve; n += G) {
            const float ev = expf(ys[n] - sp.ymax);
            ys[n] = ev;
            se += ev;
        }
        se = group_sum(se, G);
        const float ebg = expf(sp.ybg - sp.ymax);
        se += ebg;
        const float wbg = __fdiv_rn(ebg, se);
        if (!BWD) {
            // ---- blend (random_rasterizer.py:50-54) ----------------------------------------------------
            float r = 0.f, g = 0.f, bl = 0.f;
#pragma unroll 1
            for (int n = vs + lig; n < ve; n += G) {
                const float w = __fdiv_rn(ys[n], se);
                const float* c = colors_t + (int)vlist[n] * 3;
                r += w * __ldg(c);
                g += w * __ldg(c + 1);
                bl += w * __ldg(c + 2);
            }
            r = group_sum(r, G) + wbg * pb.background[0];
            g = group_sum(g, G) + wbg * pb.background[1];
            bl = group_sum(bl, G) + wbg * pb.background[2];
            const float alpha_px = 1.0f - (sp.nzero ? 0.0f : sp.prod_nz);
            if (pvalid && lig == 0) reinterpret_cast<float4*>(a.image)[gp] = make_float4(r, g, bl, alpha_px);
        } else {
            // ---- backward (hand-derived chain rule, oracle/pert_oracle.py soft_shade_fwd_bwd) ----------
            const float gbg = Gi.x * pb.background[0] + Gi.y * pb.background[1] + Gi.z * pb.background[2];
            float s1 = 0.f;  // sum_j w_j g_j
#pragma unroll 1
            for (int n = vs + lig; n < ve; n += G) {
                const float w = __fdiv_rn(ys[n], se);
                const float* c = colors_t + (int)vlist[n] * 3;
                const float gj = Gi.x * __ldg(c) + Gi.y * __ldg(c + 1) + Gi.z * __ldg(c + 2);
                ys[n] = w;
                zs[n] = gj;  // zi is recomputed below (one division) to keep four arrays
                s1 += w * gj;
            }
            s1 = group_sum(s1, G) + wbg * gbg;
            // grad y_j = w_j (g_j - s1);  grad zeta_j = (1/gamma) grad y_j
            const float gy_bg = wbg * (gbg - s1);
            float sum_gzeta = 0.f, a1 = 0.f, q = 0.f;
            const float denom = zf - zn;
            const float inv_sigma = a.L.inv_sigma;
            const bool pass = sp.zimax >= pb.eps;
            float* const gd_t = a.grad_dists + g0;
            float* const gz_t = a.grad_zbuf + g0;
            float* const gc_t = a.grad_colors ? a.grad_colors + g0 * 3 : nullptr;
            // first the sums every entry needs
#pragma unroll 1
            for (int n = vs + lig; n < ve; n += G) {
                const float w = ys[n];
                const float gy = w * (zs[n] - s1);
                const float pk = ps[n];
                if (pk > 0.0f) {  // finite logit
                    const int e = vlist[n];
                    const float zi = __fdiv_rn(zf - __ldg(zbuf_t + e), denom);
                    const float lp = logf_exact(pk);
                    const float zeta = __fadd_rn(__fadd_rn(__fmul_rn(gal, lp), zi), -sp.zmax);
                    const float gzeta = inv_g * gy;
                    sum_gzeta += gzeta;
                    a1 += zeta * gy;
                    q += lp * gzeta;
                }
            }
            sum_gzeta = group_sum(sum_gzeta, G) + inv_g * gy_bg;
            const float a1g = group_sum(a1, G);
            if (lig == 0 && pvalid) p_a1 += a1g + sp.zbg * gy_bg;
            p_q += q;  // summed over lanes at the end
            const float gzmax = -sum_gzeta;
#pragma unroll 1
            for (int n = vs + lig; n < ve; n += G) {
                const int e = vlist[n];
                const int k = e - p * K;
                const float w = ys[n];
                const float pk = ps[n];
                const float gzeta = pk > 0.0f ? inv_g * (w * (zs[n] - s1)) : 0.0f;
                const float gzi = gzeta + ((k == sp.argzi && pass) ? gzmax : 0.f);
                gz_t[e] = -gzi / denom;
                float gP = pk > 0.0f ? (gal * gzeta) / pk : 0.0f;  // log_corrected: 1/0 -> 0
                const float om = 1.0f - pk;
                float excl;
                if (sp.nzero == 0) excl = sp.prod_nz / om;
                else if (sp.nzero == 1) excl = (om == 0.f) ? sp.prod_nz : 0.f;
                else excl = 0.f;
                gP += Gi.w * excl;
                const float gu = gP * pk * om;  // sigmoid backward
                const float d = __ldg(dists_t + e);
                gd_t[e] = -gu * inv_sigma;
                p_sigma += gu * d * inv_sigma * inv_sigma;
                if (gc_t && w > 0.0f) {
                    float* gc = gc_t + e * 3;
                    gc[0] = w * Gi.x;
                    gc[1] = w * Gi.y;
                    gc[2] = w * Gi.z;
                }
            }
            // masked argzi entry: its gradient is multiplied by the mask (zero) in the reference; nothing to write
        }
    }
    if (BWD) {
        p_sigma = warp_sum(p_sigma);
        p_a1 = warp_sum(p_a1);
        p_q = warp_sum(p_q);
        if (lane == 0) reinterpret_cast<float4*>(a.partials)[tile] = make_float4(p_sigma, p_a1, p_q, 0.f);
    }
}

// d/dsigma = sum;  d/dgamma = -A1/gamma^2 + q/alpha;  d/dalpha = -q gamma/alpha^2   (A1 = sum zeta_j grad y_j)
__global__ void __launch_bounds__(1024) soft_finalize_kernel(const float* partials, int64_t n, float gamma, float alpha, float* out) {
    __shared__ double red[3][32];
    double s0 = 0, s1 = 0, s2 = 0;
    const float4* const rows = reinterpret_cast<const float4*>(partials);
    int64_t t = threadIdx.x;
    for (; t + 7 * 1024 < n; t += 8 * 1024) {  // eight independent loads in flight per thread (see shade_bwd.cu)
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(rows + t + u * 1024);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            s0 += v[u].x;
            s1 += v[u].y;
            s2 += v[u].z;
        }
    }
    for (; t < n; t += 1024) {
        const float4 v = __ldg(rows + t);
        s0 += v.x;
        s1 += v.y;
        s2 += v.z;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(FULL, s0, o);
        s1 += __shfl_xor_sync(FULL, s1, o);
        s2 += __shfl_xor_sync(FULL, s2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = s0;
        red[1][threadIdx.x >> 5] = s1;
        red[2][threadIdx.x >> 5] = s2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double r0 = 0, a1 = 0, q = 0;
        for (int w = 0; w < 32; ++w) {
            r0 += red[0][w];
            a1 += red[1][w];
            q += red[2][w];
        }
        out[0] = (float)r0;
        out[1] = (float)(-a1 / ((double)gamma * (double)gamma) + q / (double)alpha);
        out[2] = (float)(-q * (double)gamma / ((double)alpha * (double)alpha));
    }
}

template <int GT, bool BWD>
static int launch_soft_t(const SoftArgs& a, cudaStream_t st) {
    const size_t smem = (size_t)a.L.warp_smem;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(soft_shade_kernel<GT, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    soft_shade_kernel<GT, BWD><<<(unsigned)a.L.ntiles, FNT, smem, st>>>(a);
    return (int)cudaGetLastError();
}

template <bool BWD>
static int launch_soft_g(const SoftArgs& a, cudaStream_t st) {
    switch (a.L.G) {
        case 4: return launch_soft_t<4, BWD>(a, st);
        case 8: return launch_soft_t<8, BWD>(a, st);
        default: return launch_soft_t<0, BWD>(a, st);
    }
}

int launch_soft_fwd(const pert_problem& pb, const Launch& L, float* image, cudaStream_t st) {
    SoftArgs a{pb, L, image, nullptr, nullptr, nullptr, nullptr, nullptr};
    return launch_soft_g<false>(a, st);
}

int launch_soft_bwd(const pert_problem& pb, const Launch& L, const float* grad_image, float* grad_dists, float* grad_zbuf,
                    float* grad_colors, float* partials, float* grad_scalars, cudaStream_t st) {
    SoftArgs a{pb, L, nullptr, grad_image, grad_dists, grad_zbuf, grad_colors, partials};
    if (int rc = launch_soft_g<true>(a, st)) return rc;
    soft_finalize_kernel<<<1, 1024, 0, st>>>(partials, L.ntiles, pb.gamma, pb.alpha, grad_scalars);
    return (int)cudaGetLastError();
}

}  // namespace pert
