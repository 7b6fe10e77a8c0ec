// Fused backward of the perturbed shader (pert_shade_bwd in include/pertshade.h).
//
// Reference path (autograd of smooth_rgb_blend, randomras/random_rasterizer.py:34-56):
//   randomArgmax.backward     randomras/smoothagg.py:45-73
//   cat / max / clamp, log_corrected / prod_corrected backward   randomras/smoothagg.py:198-202, 303-337
//   prod backward (alpha), mask backward                          randomras/random_rasterizer.py:47-48
//   randomHeaviside.backward  randomras/smoothrast.py:40-59
//
// One warp per tile of tp pixels.  Phases of a tile:
//   0  zero-fill the tile's rows of grad_dists / grad_zbuf / grad_colors (128-bit stores), scan
//      pix_to_face -> compact list of valid entries; zbuf and saved state only for those
//   1  per-pixel logits (same arithmetic as forward)
//   2  pixels where some sample left the unperturbed winner: histogram of the saved winners, g_j,
//      c_s = g[a_s] - g[a_0]
//   3  score sums acc_j = sum_s c_s V_sj, t2_j = sum_s c_s V_sj^2 over (pixel, logit) pairs, noise
//      regenerated from the Philox counters
//   4  chain rule per valid entry -> scattered stores over the zero-filled rows; scalar partials
#include "kernels.h"
#include "tile.cuh"

namespace pert {

size_t bwd_warp_smem(int tp, int K, int sc, int nchunks, int win_bytes) {
    const size_t E = (size_t)tp * K, E1 = (size_t)tp * (K + 1);
    return carve(E, 2) /*vlist*/ + carve(E, 4) /*zs*/ + carve(E, 2) /*cnt*/ + carve(E1, 4) * (nchunks > 1 ? 4 : 3)
           /*hj gsel accs [t2s]*/ + carve(E1, 2) /*pair_j*/ + carve(E1, 1) /*pair_p*/ + carve((size_t)tp * sc, 4) /*cs*/ +
           carve(tp + 1, 4) /*vstart*/ + carve(tp, 4) * 2 /*pa0, pg0*/ + carve(tp, 1) /*apx*/ +
           (nchunks == 1 ? carve((size_t)tp * sc, win_bytes) : 0) /*wst*/ + 16;
}

__device__ __forceinline__ void zero_fill(float* dst, int n, bool vec_ok) {
    const int lane = threadIdx.x & 31;
    if (vec_ok && (n & 3) == 0) {
        float4* d4 = reinterpret_cast<float4*>(dst);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int i = lane; i < (n >> 2); i += 32) d4[i] = z;
    } else {
#pragma unroll 1
        for (int i = lane; i < n; i += 32) dst[i] = 0.f;
    }
}

// GT = lanes per pixel as a compile-time constant (1, 2, 4, 8), or 0 to read it from the launch record.
// PHASED = false is the production instantiation: both phases in one launch, histogram rebuilt from the
// saved winners (the phase-split code of the sample-sharded job is compiled out to keep the hot code small).
// FACE = true: colours gathered through pix_to_face from pb.face_colors, grad scattered by atomics
template <class NoiseA, int GT, bool PHASED, bool FACE>
__global__ void __launch_bounds__(FNT) shade_bwd_kernel(const BwdArgs a, const NoiseA noise_a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const pert_problem& pb = a.pb;
    const int lane = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int G = GT ? GT : a.L.G;
    const int gshift = GT ? (GT == 8 ? 3 : GT == 4 ? 2 : GT == 2 ? 1 : 0) : a.L.gshift;
    const int K = pb.K, K1 = K + 1, tp = 32 >> gshift, sc = a.L.sc;
    const uint32_t flags = pb.flags;
    const bool do_sample = !PHASED || (flags & PERT_PH_BWD_SAMPLE), do_finish = !PHASED || (flags & PERT_PH_BWD_FINISH);
    const int32_t* const ghist = PHASED ? a.hist : nullptr;
    const bool no_skip = flags & PERT_F_NO_SKIP;
    // how the noise of logits that can never win (zero-mean, independent of everything) is handled
    const bool per_sample = !NoiseA::kBounded || no_skip || (flags & PERT_F_PER_SAMPLE_NOISE);
    const bool drop_dead = !per_sample && (flags & PERT_F_SKIP_DEAD_NOISE);
    float p_sigma = 0.f, p_gamma = 0.f, p_q = 0.f;

    const int64_t pix0 = tile * tp;
    const int npx = (int)min((int64_t)tp, a.L.P - pix0);
    const int E = npx * K;
    const int64_t g0 = pix0 * K;
    const int p = lane >> gshift, lig = lane & (G - 1);
    const bool pvalid = p < npx;
    const int64_t gp = pix0 + p;
    const unsigned lt = (1u << lane) - 1u;
    const int sa_loc = a.L.sa_loc;
    const int wb = a.L.win_bytes;

    Carver cv(smem_raw);
    uint16_t* vlist = cv.take<uint16_t>(tp * K);
    float* zs = cv.take<float>(tp * K);
    uint16_t* cnt = cv.take<uint16_t>(tp * K);
    int* hj = cv.take<int>(tp * K1);        // winner histogram, dense in j (active pixels)
    float* gsel = cv.take<float>(tp * K1);  // g_j = <G_rgb, colour_j> of the logits that can win
    float* accs = cv.take<float>(tp * K1);  // sum_s c_s V_sj
    // sum_s c_s V_sj^2: g_j is dead once the c_s of the (only) chunk are staged, so it reuses that array
    float* t2s = a.L.nchunks > 1 ? cv.take<float>(tp * K1) : gsel;
    uint16_t* pair_j = cv.take<uint16_t>(tp * K1);
    uint8_t* pair_p = cv.take<uint8_t>(tp * K1);
    float* cs = cv.take<float>(tp * sc);  // c_s of the current sample chunk, per active pixel
    int* vstart = cv.take<int>(tp + 1);
    int* pa0 = cv.take<int>(tp);
    float* pg0 = cv.take<float>(tp);
    uint8_t* apx = cv.take<uint8_t>(tp);
    // single-chunk jobs: the tile's saved winners (tp rows of sa_loc entries, contiguous) are copied to
    // shared memory asynchronously at the very start, off the critical path
    const bool early_w = a.L.nchunks == 1 && ((sa_loc * wb) & 15) == 0 && do_sample;
    unsigned char* wst = a.L.nchunks == 1 ? cv.take<unsigned char>(tp * sc * wb) : nullptr;
    if (early_w) {
        const unsigned char* src = reinterpret_cast<const unsigned char*>(a.winners) + pix0 * sa_loc * wb;
        const int nbytes = npx * sa_loc * wb;
#pragma unroll 1
        for (int o = lane * 16; o < nbytes; o += 32 * 16) cp_async16(wst + o, src + o);
    }

    // ---- phase 0 -----------------------------------------------------------------------------------
    // per-pixel inputs do not depend on the scan: issue their loads first
    float4 Gi = make_float4(0.f, 0.f, 0.f, 0.f);
    float zn = 1.0f, zf = 100.0f;
    int pstate = K;
    if (pvalid) {
        Gi = __ldg(reinterpret_cast<const float4*>(a.grad_image) + gp);
        const int b = pb.depth_len > 1 ? (int)(gp / a.L.HW) : 0;
        zn = __ldg(pb.znear + b);
        zf = __ldg(pb.zfar + b);
        pstate = a.pixstate[gp];
    }
    if (do_finish) {
        zero_fill(a.grad_dists + g0, E, a.L.vec_ok);
        zero_fill(a.grad_zbuf + g0, E, a.L.vec_ok);
        if (a.grad_colors && !FACE) zero_fill(a.grad_colors + g0 * 3, E * 3, a.L.vec_ok);
    } else {
        zero_fill(a.acc + pix0 * K1, npx * K1, false);
        if (lane < npx) {
            a.pixstat[(pix0 + lane) * 2] = 0.f;
            a.pixstat[(pix0 + lane) * 2 + 1] = 0.f;
        }
    }
    const int nv = scan_valid(pb.pix_to_face + g0, E, a.L.vec_ok, vlist);
    if (nv > 0) {
        __syncwarp();
        pixel_ranges(vlist, nv, K, tp, vstart);
        {
            const float* const zbuf_t = pb.zbuf + g0;
            const uint16_t* const counts_t = a.counts + g0;
            const float* const colors_t = pb.colors + g0 * 3;
            const float* const rsum_t = a.rsum + g0;
#pragma unroll 1
            for (int n = lane; n < nv; n += 32) {
                const int e = vlist[n];
                zs[n] = __ldg(zbuf_t + e);
                cnt[n] = counts_t[e];
                // needed a few round trips later (g_j of the logits that can win; chain rule): start them now
                if (!FACE) prefetch_l1(colors_t + e * 3);
                prefetch_l1(rsum_t + e);
            }
        }
        __syncwarp();

        // ---- phase 1 -------------------------------------------------------------------------------
        const float gal = pb.gamma / pb.alpha;
        const PixPrep pi = prep_pixels(p, lig, G, pvalid, K, vstart, vlist, cnt, zs, zn, zf, pb.S_rast, gal, pb.eps);
        const int vs = pvalid ? vstart[p] : 0, ve = pvalid ? vstart[p + 1] : 0;
        const int nvp = ve - vs;
        const int a0 = pi.a0;
        // a pixel whose samples all picked a0 has c_s = 0 for every s: every score sum is exactly 0
        // (and forward wrote no winners row for it)
        const bool act = pvalid && do_sample && (pstate & 0x8000);
        const float floor_v = pi.zeta_max - live_cut(pb.gamma, pi.zeta_max, NoiseA::kBounded && !no_skip);
        float t2sum = 0.f, csum = 0.f;
        const float* gacc = a.acc + gp * K1;  // FINISH-only: sums over all sample shards
        const float* const colors_p = pb.colors + gp * K * 3;
        const float* const fcol = FACE ? pb.face_colors : nullptr;  // per-face colours gathered through pix_to_face
        const int64_t* const p2f_p = pb.pix_to_face + gp * K;
        __syncwarp();

        // ---- phase 2 -------------------------------------------------------------------------------
        const unsigned actb = __ballot_sync(FULL, act && lig == 0);
        const int na = __popc(actb);
        const int my_ai = __popc(actb & ((1u << (p * G)) - 1u));
        if (na > 0) {
            // g_j of every logit that can win (the saved winners are among them), cleared sums, and the
            // list of (pixel, logit) pairs that need per-sample noise
            int np2 = 0;
            {
                if (act && lig == 0) {
                    apx[my_ai] = (uint8_t)p;
                    pa0[p] = a0;
                }
                if (act) {
#pragma unroll 1
                    for (int j = lig; j < K1; j += G) {
                        hj[p * K1 + j] = 0;
                        accs[p * K1 + j] = 0.f;
                    }
                }
                const int span = per_sample ? max(K1, nvp + 1) : nvp + 1;
                const int iters = warp_max_i(act ? (span + G - 1) / G : 0);
#pragma unroll 1
                for (int it = 0; it < iters; ++it) {
                    const int idx = it * G + lig;
                    bool want = false;
                    int j = 0;
                    if (act && idx <= nvp) {
                        j = idx < nvp ? (int)vlist[vs + idx] - p * K : K;
                        const float z = idx < nvp ? zs[vs + idx] : pi.zbg;
                        const bool live = z > -CUDART_INF_F && z >= floor_v;
                        if (live) {
                            float gj;
                            if (j < K) {
                                const float* c = fcol ? fcol + 3 * (int)__ldg(p2f_p + j) : colors_p + j * 3;
                                gj = Gi.x * __ldg(c) + Gi.y * __ldg(c + 1) + Gi.z * __ldg(c + 2);
                            } else {
                                gj = Gi.x * pb.background[0] + Gi.y * pb.background[1] + Gi.z * pb.background[2];
                            }
                            gsel[p * K1 + j] = gj;
                            if (j == a0) pg0[p] = gj;
                        }
                        want = live && !per_sample;
                    }
                    if (per_sample && act && idx < K1) {  // every logit, dense in j
                        j = idx;
                        want = true;
                    }
                    const unsigned wbal = __ballot_sync(FULL, want);
                    if (want) {
                        const int pos = np2 + __popc(wbal & lt);
                        pair_j[pos] = (uint16_t)j;
                        pair_p[pos] = (uint8_t)my_ai;
                    }
                    np2 += __popc(wbal);
                }
            }
            __syncwarp();

            // ---- phase 3: one pass over the saved winners per sample chunk: histogram + c_s, then the
            //      score sums of the pairs ------------------------------------------------------------------
            const int qb = pb.s_agg_begin >> 2;
            if (early_w) {
                cp_async_wait_all();
                __syncwarp();
            }
            // lanes per pair: as many as keep the warp full, at most a.L.lpp
            const int lpp_shift = np2 == 0 ? 0 : min(a.L.lpp_shift, np2 >= 32 ? 0 : 31 - __clz(32 / np2));
            const int LPP = 1 << lpp_shift;
            const int lq = lane & (LPP - 1);
            float C2 = 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < sa_loc; c0 += sc) {  // c0 multiple of 32
                const int cn = min(sc, sa_loc - c0);
                const int cn4 = (cn + 3) & ~3;
#pragma unroll 1
                for (int ai = 0; ai < na; ++ai) {
                    const int pp = apx[ai];
                    const float* gs = gsel + pp * K1;
                    int* hp = hj + pp * K1;
                    const float g0v = pg0[pp];
                    const int ppa0 = pa0[pp];
                    const int64_t wbase = (pix0 + pp) * sa_loc + c0;
#pragma unroll 1
                    for (int s = lane; s < cn4; s += 32) {
                        float c = 0.f;
                        if (s < cn) {
                            const int w = early_w ? load_winner(wst, wb, pp * sa_loc + s) : load_winner(a.winners, wb, wbase + s);
                            if (w != ppa0) {  // a0 (the most frequent by far) is counted as the remainder
                                atomicAdd(&hp[w], 1);
                                c = gs[w] - g0v;
                            }
                        }
                        cs[ai * sc + s] = c;
                    }
                }
                __syncwarp();
                if (c0 + sc >= sa_loc) {
                    // all winners seen: histogram remainder, sum_s c_s and sum_s c_s^2 from the histogram (read
                    // g_j now: the pair loop below may reuse its array).  Group reductions run on every lane:
                    // pixels of one warp differ in `act`.
                    int others = 0;
                    float c1 = 0.f, c2 = 0.f;
                    if (act) {
                        const float g0v = pg0[p];
#pragma unroll 1
                        for (int idx = lig; idx <= nvp; idx += G) {
                            const int j = idx < nvp ? (int)vlist[vs + idx] - p * K : K;
                            const int h = hj[p * K1 + j];
                            if (h > 0 && j != a0) {
                                others += h;
                                const float d = gsel[p * K1 + j] - g0v;
                                c1 += (float)h * d;
                                c2 += (float)h * d * d;
                            }
                        }
                    }
                    others = group_sum_i(others, G);
                    csum = group_sum(c1, G);
                    C2 = group_sum(c2, G);
                    if (act && lig == 0) hj[p * K1 + a0] = sa_loc - others;
                    __syncwarp();
                }
                const int nqc = cn4 >> 2;
#pragma unroll 1
                for (int it0 = 0; it0 < (np2 << lpp_shift); it0 += 32) {
                    const int it = it0 + lane;
                    const int pr = it >> lpp_shift;
                    const bool on = pr < np2;
                    const int j = on ? pair_j[pr] : 0;
                    const int ai = on ? pair_p[pr] : 0;
                    const int pp = apx[ai];
                    float acc = 0.f, t2 = 0.f;
                    if (on) {
                        const float4* c4p = reinterpret_cast<const float4*>(cs + ai * sc);
#pragma unroll 1
                        for (int ql = lq; ql < nqc; ql += LPP) {
                            const float4 c4 = c4p[ql];
                            if (!no_skip && c4.x == 0.f && c4.y == 0.f && c4.z == 0.f && c4.w == 0.f) continue;
                            float nz[4];
                            noise_a.get4(qb + (c0 >> 2) + ql, j, pix0 + pp, nz);
                            const float cv0 = c4.x * nz[0], cv1 = c4.y * nz[1], cv2 = c4.z * nz[2], cv3 = c4.w * nz[3];
                            acc += (cv0 + cv1) + (cv2 + cv3);
                            t2 += (cv0 * nz[0] + cv1 * nz[1]) + (cv2 * nz[2] + cv3 * nz[3]);
                        }
                    }
#pragma unroll 1
                    for (int o = LPP >> 1; o > 0; o >>= 1) {
                        acc += __shfl_xor_sync(FULL, acc, o);
                        t2 += __shfl_xor_sync(FULL, t2, o);
                    }
                    if (on && lq == 0) {
                        if (c0 == 0) {
                            accs[pp * K1 + j] = acc;
                            t2s[pp * K1 + j] = t2;
                        } else {
                            accs[pp * K1 + j] += acc;
                            t2s[pp * K1 + j] += t2;
                        }
                    }
                }
                __syncwarp();
            }

            // ---- logits that can never win: their noise is independent of every a_s, so given c the sum
            //      sum_s c_s V_sj is N(0, sum_s c_s^2) exactly: ONE draw per logit instead of S_agg -------------
            {
                float t = 0.f;
                if (per_sample) {
                    if (act) {
#pragma unroll 1
                        for (int j = lig; j < K1; j += G) t += t2s[p * K1 + j];
                    }
                    t2sum = group_sum(t, G);
                } else {
                    int nlive = 0;
                    const float sC2 = sqrtf(C2);
                    const uint32_t qx = 0xC0000000u + (uint32_t)qb;  // counters no sample quad uses
                    if (act) {
#pragma unroll 1
                        for (int idx = lig; idx <= nvp; idx += G) {
                            const int j = idx < nvp ? (int)vlist[vs + idx] - p * K : K;
                            const float z = idx < nvp ? zs[vs + idx] : pi.zbg;
                            if (z > -CUDART_INF_F && z >= floor_v) {
                                nlive++;
                                t += t2s[p * K1 + j];
                            } else if (!drop_dead) {
                                if constexpr (NoiseA::kBounded) {
                                    float nz[4];
                                    noise_a.get4(qx, j, gp, nz);
                                    accs[p * K1 + j] = sC2 * nz[0];
                                }
                            }
                        }
                    }
                    nlive = group_sum_i(nlive, G);
                    t2sum = group_sum(t, G);
                    if (act) {
                        const int n_nl = K1 - nlive;  // logits without per-sample noise
                        const int npad = K - nvp;     // of which masked (they only enter through gzmax)
                        float nz[4] = {0.f, 0.f, 0.f, 0.f};
                        if constexpr (NoiseA::kBounded) {
                            if (!drop_dead) noise_a.get4(qx, K1, gp, nz);
                        }
                        if (lig == 0 && pi.kpad < K) accs[p * K1 + pi.kpad] = sqrtf((float)npad * C2) * nz[0];
                        // sum_j sum_s c_s V_sj^2 over those logits: mean n*csum, variance 2 n sum_s c_s^2
                        t2sum += (float)n_nl * csum + sqrtf(2.0f * (float)n_nl * C2) * nz[1];
                    }
                }
            }
            __syncwarp();
        }

        if (!do_finish) {
            // sample-sharded job: publish the partial sums, the caller all-reduces them
            if (act) {
#pragma unroll 1
                for (int j = lig; j < K1; j += G) a.acc[gp * K1 + j] = accs[p * K1 + j];
                if (lig == 0) {
                    a.pixstat[gp * 2 + 0] = t2sum;
                    a.pixstat[gp * 2 + 1] = csum;
                }
            }
        } else {
            // ---- phase 4: chain rule per pixel ----------------------------------
            const float invSg = 1.0f / ((float)pb.S_agg * pb.gamma);
            const float inv_sr = 1.0f / ((float)pb.S_rast * pb.sigma);
            const float invS = 1.0f / (float)pb.S_agg;
            const bool from_global = PHASED && !do_sample;
            const bool has_acc = act || (from_global && pvalid);
            if (from_global && pvalid) {
                t2sum = a.pixstat[gp * 2];
                csum = a.pixstat[gp * 2 + 1];
            }
            // grad_zeta_j = acc_j / (S gamma);  gzmax = -sum_j grad_zeta_j over ALL K+1 logits
            float sg = 0.f;
            if (has_acc) {
                if (from_global) {
#pragma unroll 1
                    for (int j = lig; j < K1; j += G) sg += gacc[j] * invSg;
                } else if (per_sample) {
#pragma unroll 1
                    for (int j = lig; j < K1; j += G) sg += accs[p * K1 + j] * invSg;
                } else {
#pragma unroll 1
                    for (int idx = lig; idx <= nvp; idx += G) {
                        const int j = idx < nvp ? (int)vlist[vs + idx] - p * K : K;
                        sg += accs[p * K1 + j] * invSg;
                    }
                    if (lig == 0 && pi.kpad < K) sg += accs[p * K1 + pi.kpad] * invSg;
                }
            }
            const float gzmax = -group_sum(sg, G);
            if (has_acc && lig == 0) p_gamma += (t2sum - csum) * invSg;
            const float denom = zf - zn;
            const bool pass = pi.zimax >= pb.eps;
            const float fS = (float)pb.S_rast;
            const float* const rsum_t = a.rsum + g0;
            float* const gd_t = a.grad_dists + g0;
            float* const gz_t = a.grad_zbuf + g0;
            float* const gc_t = (a.grad_colors && !fcol) ? a.grad_colors + g0 * 3 : nullptr;
            float* const gfc = fcol ? a.grad_colors : nullptr;  // (num_faces,3), atomic scatter
#pragma unroll 1
            for (int n = vs + lig; n < ve; n += G) {
                const int e = vlist[n];
                const int k = e - p * K;
                const float rsn = rsum_t[e];
                float gz = 0.f;
                if (has_acc) gz = (from_global ? gacc[k] : accs[p * K1 + k]) * invSg;
                const float gzi = gz + ((k == pi.argzi && pass) ? gzmax : 0.f);
                gz_t[e] = -gzi / denom;
                const int c = cnt[n];
                const float pk = (float)c / fS;
                float gP = 0.f;
                if (c != 0) {
                    const float lp = (c == pb.S_rast) ? 0.0f : logf(pk);
                    p_q += lp * gz;        // prod_corrected: inf -> 0 on the scalar side
                    gP = (gal * gz) / pk;  // log_corrected: 1/0 -> 0
                }
                const float om = 1.0f - pk;
                float excl;
                if (pi.nzero == 0) excl = pi.prod_nz / om;
                else if (pi.nzero == 1) excl = (om == 0.f) ? pi.prod_nz : 0.f;
                else excl = 0.f;
                gP += Gi.w * excl;
                const float gx = gP * (rsn * inv_sr);
                gd_t[e] = -gx;
                p_sigma += gx;
                // grad_colors = w_k * G_rgb
                if (gc_t || gfc) {
                    int h;
                    if (ghist) h = ghist[gp * K1 + k];  // all-shard histogram when sample-sharded
                    else if (act) h = hj[p * K1 + k];
                    else h = (k == a0) ? sa_loc : 0;
                    if (h > 0) {
                        const float w = (float)h * invS;
                        if (gfc) {
                            float* gc = gfc + 3 * (int)__ldg(p2f_p + k);
                            atomicAdd(gc, w * Gi.x);
                            atomicAdd(gc + 1, w * Gi.y);
                            atomicAdd(gc + 2, w * Gi.z);
                        } else {
                            float* gc = gc_t + e * 3;
                            gc[0] = w * Gi.x;
                            gc[1] = w * Gi.y;
                            gc[2] = w * Gi.z;
                        }
                    }
                }
            }
        }
    }

    // ---- scalar partials of this warp (fixed order) --------------------------------------------------
    if (do_finish) {
        p_sigma = warp_sum(p_sigma);
        p_gamma = warp_sum(p_gamma);
        p_q = warp_sum(p_q);
        if (lane == 0) reinterpret_cast<float4*>(a.partials)[tile] = make_float4(p_sigma, p_gamma, p_q, 0.f);
    }
}

// deterministic final reduction of the per-tile scalar partials (one CTA)
//   out[0] = d/dsigma = sum gx                         (smoothrast.py:57-58)
//   out[1] = d/dgamma = score term + q/alpha           (smoothagg.py:72 and :329-332 through gamma/alpha)
//   out[2] = d/dalpha = -q gamma / alpha^2
__global__ void __launch_bounds__(1024) finalize_scalars_kernel(const float* partials, int64_t n, float gamma,
                                                                float alpha, float* out) {
    __shared__ double red[3][32];
    double s0 = 0, s1 = 0, s2 = 0;
    for (int64_t t = threadIdx.x; t < n; t += 1024) {
        const float4 v = reinterpret_cast<const float4*>(partials)[t];
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
        double r0 = 0, r1 = 0, q = 0;
        for (int w = 0; w < 32; ++w) {
            r0 += red[0][w];
            r1 += red[1][w];
            q += red[2][w];
        }
        out[0] = (float)r0;
        out[1] = (float)(r1 + q / (double)alpha);
        out[2] = (float)(-q * (double)gamma / ((double)alpha * (double)alpha));
    }
}

template <class NA, int GT, bool PHASED, bool FACE>
static int launch_bwd_f(const BwdArgs& a, const NA& na, cudaStream_t st) {
    const size_t smem = (size_t)a.L.warp_smem;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(shade_bwd_kernel<NA, GT, PHASED, FACE>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    shade_bwd_kernel<NA, GT, PHASED, FACE><<<(unsigned)a.L.ntiles, FNT, smem, st>>>(a, na);
    return (int)cudaGetLastError();
}
template <class NA, int GT, bool PHASED>
static int launch_bwd_t(const BwdArgs& a, const NA& na, cudaStream_t st) {
    return a.pb.face_colors ? launch_bwd_f<NA, GT, PHASED, true>(a, na, st) : launch_bwd_f<NA, GT, PHASED, false>(a, na, st);
}

int launch_shade_bwd(const BwdArgs& a, float* grad_scalars, cudaStream_t st) {
    int rc;
    const uint32_t both = PERT_PH_BWD_SAMPLE | PERT_PH_BWD_FINISH;
    const bool phased = (a.pb.flags & both) != both || a.hist != nullptr;
    if (a.pb.noise_agg) {
        ExplicitNoise xa{a.pb.noise_agg, a.L.P, a.pb.K + 1, a.pb.S_agg};
        rc = launch_bwd_t<ExplicitNoise, 0, true>(a, xa, st);
    } else {
        PhiloxNoise pa(a.pb.seed_agg, 1, a.pb.pixel_offset);
        if (phased) {
            rc = launch_bwd_t<PhiloxNoise, 0, true>(a, pa, st);
        } else {
            switch (a.L.G) {  // production path: lanes per pixel known at compile time
                case 1: rc = launch_bwd_t<PhiloxNoise, 1, false>(a, pa, st); break;
                case 2: rc = launch_bwd_t<PhiloxNoise, 2, false>(a, pa, st); break;
                case 4: rc = launch_bwd_t<PhiloxNoise, 4, false>(a, pa, st); break;
                default: rc = launch_bwd_t<PhiloxNoise, 8, false>(a, pa, st); break;
            }
        }
    }
    if (rc) return rc;
    if (a.pb.flags & PERT_PH_BWD_FINISH) {
        finalize_scalars_kernel<<<1, 1024, 0, st>>>(a.partials, a.L.ntiles, a.pb.gamma, a.pb.alpha, grad_scalars);
        return (int)cudaGetLastError();
    }
    return 0;
}

}  // namespace pert
