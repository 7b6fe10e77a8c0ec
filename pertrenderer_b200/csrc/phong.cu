// Phong lighting of the fragment entries (pert_phong_fwd / pert_phong_bwd in include/pertshade.h): the
// `colors` (P,K,3) RandomPhongShader hands to smooth_rgb_blend (randomras/random_rasterizer.py:103-113).
//
// Restates pytorch3d 0.4.0 (requirements.txt:7; its source is not under the reference tree):
//   shading.phong_shading:      pixel_coords / pixel_normals = interpolate_face_attributes(pix_to_face, bary, .)
//                               colors = (ambient + diffuse) * texels + specular
//   shading._apply_lighting:    ambient = materials.ambient * lights.ambient, diffuse = materials.diffuse *
//                               lights.diffuse(normals, points), specular = materials.specular * lights.specular(..)
//   lighting.diffuse/specular:  F.normalize(., eps=1e-6) of normal, light direction and view direction,
//                               relu(n.d), reflect = -d + 2 (n.d) n, relu(v.r) * [n.d > 0], pow(., shininess)
//
// A streaming pass over the (P,K) entries: pix_to_face is read coalesced (8 B per entry); everything else is
// touched for valid entries only (real fragments are sparse in K), and in backward only for entries whose
// colour gradient is non-zero (the blend gives weight to a few winners per pixel).  Face-table gradients go
// through a per-CTA shared-memory table when the mesh is small (every CTA would hammer the same few hundred
// L2 addresses otherwise) and straight to global atomics when it is large.
#include "kernels.h"

namespace pert {

namespace {

constexpr int PT = 256;  // threads per CTA
constexpr int PU = 4;    // entries per thread and chunk (loads of pix_to_face in flight)
constexpr int PCHUNK = PT * PU;
constexpr int TABLE_MAX_FACES = 512;  // 512 * 18 floats = 36 KB of shared memory

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 mk(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 ld3(const float* p) { return mk(__ldg(p), __ldg(p + 1), __ldg(p + 2)); }
__device__ __forceinline__ void st3(float* p, V3 v) {
    p[0] = v.x;
    p[1] = v.y;
    p[2] = v.z;
}

constexpr float kNormEps = 1e-6f;  // F.normalize(eps=1e-6) in pytorch3d's lighting.py

// x = raw / max(|raw|, eps)
__device__ __forceinline__ V3 normalize(V3 raw, float& len) {
    len = sqrtf(dot(raw, raw));
    return (1.0f / fmaxf(len, kNormEps)) * raw;
}
// gradient of normalize: (g - x (x.g)) / |raw| where the clamp is inactive, g / eps where it is
__device__ __forceinline__ V3 normalize_bwd(V3 x, float len, V3 g) {
    if (len > kNormEps) return (1.0f / len) * (g - dot(x, g) * x);
    return (1.0f / kNormEps) * g;
}

struct Row {  // one row of `lighting`
    V3 loc, amb, dif, spc, cam;
    float sh;
    bool directional;
};
__device__ __forceinline__ Row row_from(const float* r) {
    Row o;
    o.loc = mk(r[0], r[1], r[2]);
    o.amb = mk(r[3], r[4], r[5]);
    o.dif = mk(r[6], r[7], r[8]);
    o.spc = mk(r[9], r[10], r[11]);
    o.sh = r[12];
    o.cam = mk(r[13], r[14], r[15]);
    o.directional = r[16] != 0.0f;
    return o;
}

struct Lit {  // what backward needs of the forward evaluation
    V3 n, d, v, r;
    float nl, dl, vl, cosd, ang, vr, alpha, pw;
};

__device__ __forceinline__ Lit light_entry(const Row& L, V3 p, V3 n_raw) {
    Lit o;
    o.n = normalize(n_raw, o.nl);
    o.d = normalize(L.directional ? L.loc : L.loc - p, o.dl);
    o.cosd = dot(o.n, o.d);
    o.ang = fmaxf(o.cosd, 0.0f);
    o.v = normalize(L.cam - p, o.vl);
    o.r = (2.0f * o.cosd) * o.n - o.d;
    o.vr = dot(o.v, o.r);
    o.alpha = o.cosd > 0.0f ? fmaxf(o.vr, 0.0f) : 0.0f;
    o.pw = powf(o.alpha, L.sh);
    return o;
}

// The lighting rows of the batch elements a chunk of entries can touch: the chunk's first batch element and
// the next one sit in shared memory, anything further (tiny images) is read from global memory.
struct RowCache {
    float* s;  // 2 * PERT_PHONG_STRIDE floats
    int64_t b0;
    const float* g;
    __device__ __forceinline__ Row get(int64_t b) const {
        const int64_t o = b - b0;
        if (o == 0 || o == 1) return row_from(s + o * PERT_PHONG_STRIDE);
        float r[PERT_PHONG_STRIDE];
#pragma unroll
        for (int i = 0; i < 17; ++i) r[i] = __ldg(g + b * PERT_PHONG_STRIDE + i);
        return row_from(r);
    }
};

__device__ __forceinline__ void fill_rows(const pert_phong& ph, int64_t b0, float* srow) {
    if (threadIdx.x < 2 * PERT_PHONG_STRIDE) {
        const int64_t b = b0 + threadIdx.x / PERT_PHONG_STRIDE;
        const int64_t nb = ph.light_rows;
        srow[threadIdx.x] = __ldg(ph.lighting + (b < nb ? b : nb - 1) * PERT_PHONG_STRIDE + threadIdx.x % PERT_PHONG_STRIDE);
    }
}

__global__ void __launch_bounds__(PT) phong_fwd_kernel(const pert_phong ph, float* __restrict__ colors, int64_t E,
                                                       int64_t nchunks) {
    __shared__ float srow[2 * PERT_PHONG_STRIDE];
    const bool sparse = ph.flags & PERT_PHONG_SPARSE;
    const int64_t HWK = ph.HW * ph.K;
    int64_t cached_b0 = -1;
#pragma unroll 1
    for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const int64_t e_base = c * PCHUNK;
        const int64_t b0 = ph.light_rows > 1 ? e_base / HWK : 0;
        if (b0 != cached_b0) {  // block-uniform
            __syncthreads();
            fill_rows(ph, b0, srow);
            __syncthreads();
            cached_b0 = b0;
        }
        const RowCache rows{srow, b0, ph.lighting};
        long long f[PU];
#pragma unroll
        for (int u = 0; u < PU; ++u) {
            const int64_t e = e_base + u * PT + threadIdx.x;
            f[u] = e < E ? __ldg(ph.pix_to_face + e) : -2;
        }
#pragma unroll
        for (int u = 0; u < PU; ++u) {
            const int64_t e = e_base + u * PT + threadIdx.x;
            if (f[u] < 0) {
                // masked interpolation: zero point and normal -> no diffuse, alpha = 0; what is left is
                // ambient * texel (+ specular * 0^shininess, which is 1 for shininess 0)
                if (f[u] == -2 || sparse) continue;
                const Row L = rows.get(ph.light_rows > 1 ? e / HWK : 0);
                const V3 t = ph.face_colors ? mk(0.f, 0.f, 0.f) : ld3(ph.texels + e * 3);
                const float pw = L.sh == 0.0f ? 1.0f : 0.0f;
                st3(colors + e * 3, mk(L.amb.x * t.x + L.spc.x * pw, L.amb.y * t.y + L.spc.y * pw, L.amb.z * t.z + L.spc.z * pw));
                continue;
            }
            const V3 b = ld3(ph.bary + e * 3);
            const float* fv = ph.face_verts + f[u] * 9;
            const float* fn = ph.face_normals + f[u] * 9;
            const V3 p = b.x * ld3(fv) + b.y * ld3(fv + 3) + b.z * ld3(fv + 6);
            const V3 nr = b.x * ld3(fn) + b.y * ld3(fn + 3) + b.z * ld3(fn + 6);
            const V3 t = ph.face_colors ? ld3(ph.face_colors + f[u] * 3) : ld3(ph.texels + e * 3);
            const Row L = rows.get(ph.light_rows > 1 ? e / HWK : 0);
            const Lit o = light_entry(L, p, nr);
            st3(colors + e * 3, mk((L.amb.x + L.dif.x * o.ang) * t.x + L.spc.x * o.pw,
                                   (L.amb.y + L.dif.y * o.ang) * t.y + L.spc.y * o.pw,
                                   (L.amb.z + L.dif.z * o.ang) * t.z + L.spc.z * o.pw));
        }
    }
}

// TABLE: accumulate the (F,3,3) gradients of face_verts / face_normals (and the (F,3) gradient of
// face_colors) in shared memory and flush once per CTA.
template <bool TABLE>
__global__ void __launch_bounds__(PT) phong_bwd_kernel(const pert_phong ph, const float* __restrict__ grad_colors,
                                                       float* __restrict__ grad_texels, float* __restrict__ grad_bary,
                                                       float* __restrict__ grad_fv, float* __restrict__ grad_fn, int64_t E,
                                                       int64_t nchunks) {
    __shared__ float srow[2 * PERT_PHONG_STRIDE];
    extern __shared__ float table[];  // TABLE: F*9 (verts) | F*9 (normals) | F*3 (face colours)
    const int F = (int)ph.num_faces;
    float* const t_fv = table;
    float* const t_fn = table + F * 9;
    float* const t_fc = table + F * 18;
    const bool face_tex = ph.face_colors != nullptr;
    if (TABLE) {
        for (int i = threadIdx.x; i < F * 21; i += PT) table[i] = 0.0f;
        __syncthreads();
    }
    const bool sparse = ph.flags & PERT_PHONG_SPARSE;
    const int64_t HWK = ph.HW * ph.K;
    int64_t cached_b0 = -1;
#pragma unroll 1
    for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const int64_t e_base = c * PCHUNK;
        const int64_t b0 = ph.light_rows > 1 ? e_base / HWK : 0;
        if (b0 != cached_b0) {
            __syncthreads();
            fill_rows(ph, b0, srow);
            __syncthreads();
            cached_b0 = b0;
        }
        const RowCache rows{srow, b0, ph.lighting};
        long long f[PU];
#pragma unroll
        for (int u = 0; u < PU; ++u) {
            const int64_t e = e_base + u * PT + threadIdx.x;
            f[u] = e < E ? __ldg(ph.pix_to_face + e) : -2;
        }
#pragma unroll
        for (int u = 0; u < PU; ++u) {
            const int64_t e = e_base + u * PT + threadIdx.x;
            const V3 zero = mk(0.f, 0.f, 0.f);
            if (f[u] < 0) {
                if (f[u] == -2 || sparse) continue;
                if (grad_texels && !face_tex) {
                    const Row L = rows.get(ph.light_rows > 1 ? e / HWK : 0);
                    const V3 gc = ld3(grad_colors + e * 3);
                    st3(grad_texels + e * 3, mk(gc.x * L.amb.x, gc.y * L.amb.y, gc.z * L.amb.z));
                }
                if (grad_bary) st3(grad_bary + e * 3, zero);
                continue;
            }
            const V3 gc = ld3(grad_colors + e * 3);
            if (gc.x == 0.0f && gc.y == 0.0f && gc.z == 0.0f) {  // not a winner of any sample: every gradient is 0
                if (grad_texels && !face_tex) st3(grad_texels + e * 3, zero);
                if (grad_bary) st3(grad_bary + e * 3, zero);
                continue;
            }
            const V3 b = ld3(ph.bary + e * 3);
            const float* fv = ph.face_verts + f[u] * 9;
            const float* fn = ph.face_normals + f[u] * 9;
            const V3 v0 = ld3(fv), v1 = ld3(fv + 3), v2 = ld3(fv + 6);
            const V3 n0 = ld3(fn), n1 = ld3(fn + 3), n2 = ld3(fn + 6);
            const V3 p = b.x * v0 + b.y * v1 + b.z * v2;
            const V3 nr = b.x * n0 + b.y * n1 + b.z * n2;
            const V3 t = face_tex ? ld3(ph.face_colors + f[u] * 3) : ld3(ph.texels + e * 3);
            const Row L = rows.get(ph.light_rows > 1 ? e / HWK : 0);
            const Lit o = light_entry(L, p, nr);

            // colour = (amb + dif * ang) * t + spc * pw
            const V3 gt = mk(gc.x * (L.amb.x + L.dif.x * o.ang), gc.y * (L.amb.y + L.dif.y * o.ang),
                             gc.z * (L.amb.z + L.dif.z * o.ang));
            if (grad_texels) {
                if (face_tex) {
                    float* dst = TABLE ? t_fc + (int)f[u] * 3 : grad_texels + f[u] * 3;
                    atomicAdd(dst, gt.x);
                    atomicAdd(dst + 1, gt.y);
                    atomicAdd(dst + 2, gt.z);
                } else {
                    st3(grad_texels + e * 3, gt);
                }
            }
            const float g_ang = gc.x * L.dif.x * t.x + gc.y * L.dif.y * t.y + gc.z * L.dif.z * t.z;
            const float g_pw = gc.x * L.spc.x + gc.y * L.spc.y + gc.z * L.spc.z;
            // pw = alpha^sh; alpha = relu(v.r) * [n.d > 0]   (torch's pow backward: 0 at alpha = 0)
            const float g_vr = (o.alpha > 0.0f && L.sh != 0.0f) ? g_pw * L.sh * powf(o.alpha, L.sh - 1.0f) : 0.0f;
            const V3 g_v = g_vr * o.r, g_r = g_vr * o.v;
            // r = 2 cos n - d; ang = relu(cos); cos = n.d
            const float g_cos = 2.0f * dot(g_r, o.n) + (o.cosd > 0.0f ? g_ang : 0.0f);
            const V3 g_n = (2.0f * o.cosd) * g_r + g_cos * o.d;
            const V3 g_d = g_cos * o.n - g_r;
            const V3 g_nraw = normalize_bwd(o.n, o.nl, g_n);
            const V3 g_vraw = normalize_bwd(o.v, o.vl, g_v);
            V3 g_p = mk(-g_vraw.x, -g_vraw.y, -g_vraw.z);
            if (!L.directional) g_p = g_p - normalize_bwd(o.d, o.dl, g_d);
            if (grad_bary)
                st3(grad_bary + e * 3, mk(dot(g_p, v0) + dot(g_nraw, n0), dot(g_p, v1) + dot(g_nraw, n1),
                                          dot(g_p, v2) + dot(g_nraw, n2)));
            const float bw[3] = {b.x, b.y, b.z};
            if (grad_fv) {
                float* dst = TABLE ? t_fv + (int)f[u] * 9 : grad_fv + f[u] * 9;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    atomicAdd(dst + 3 * i, bw[i] * g_p.x);
                    atomicAdd(dst + 3 * i + 1, bw[i] * g_p.y);
                    atomicAdd(dst + 3 * i + 2, bw[i] * g_p.z);
                }
            }
            if (grad_fn) {
                float* dst = TABLE ? t_fn + (int)f[u] * 9 : grad_fn + f[u] * 9;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    atomicAdd(dst + 3 * i, bw[i] * g_nraw.x);
                    atomicAdd(dst + 3 * i + 1, bw[i] * g_nraw.y);
                    atomicAdd(dst + 3 * i + 2, bw[i] * g_nraw.z);
                }
            }
        }
    }
    if (TABLE) {
        __syncthreads();
        for (int i = threadIdx.x; i < F * 9; i += PT) {
            if (grad_fv && t_fv[i] != 0.0f) atomicAdd(grad_fv + i, t_fv[i]);
            if (grad_fn && t_fn[i] != 0.0f) atomicAdd(grad_fn + i, t_fn[i]);
        }
        if (face_tex && grad_texels)
            for (int i = threadIdx.x; i < F * 3; i += PT)
                if (t_fc[i] != 0.0f) atomicAdd(grad_texels + i, t_fc[i]);
    }
}

unsigned phong_grid(int64_t nchunks) {
    const int64_t cap = 148 * 8;  // 8 resident CTAs of 256 threads per SM
    return (unsigned)(nchunks < cap ? nchunks : cap);
}

}  // namespace

int launch_phong_fwd(const pert_phong& ph, float* colors, cudaStream_t st) {
    const int64_t E = ph.P * ph.K, nchunks = (E + PCHUNK - 1) / PCHUNK;
    phong_fwd_kernel<<<phong_grid(nchunks), PT, 0, st>>>(ph, colors, E, nchunks);
    return (int)cudaGetLastError();
}

int launch_phong_bwd(const pert_phong& ph, const float* grad_colors, float* grad_texels, float* grad_bary, float* grad_fv,
                     float* grad_fn, cudaStream_t st) {
    const int64_t E = ph.P * ph.K, nchunks = (E + PCHUNK - 1) / PCHUNK;
    const bool scatter = grad_fv || grad_fn || (ph.face_colors && grad_texels);
    if (scatter && ph.num_faces <= TABLE_MAX_FACES) {
        const size_t smem = (size_t)ph.num_faces * 21 * sizeof(float);
        cudaError_t e = cudaFuncSetAttribute(phong_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        // fewer, longer-lived CTAs: each one flushes its table once
        const int64_t cap = 148 * 4;
        phong_bwd_kernel<true><<<(unsigned)(nchunks < cap ? nchunks : cap), PT, smem, st>>>(ph, grad_colors, grad_texels,
                                                                                           grad_bary, grad_fv, grad_fn, E, nchunks);
    } else {
        phong_bwd_kernel<false><<<phong_grid(nchunks), PT, 0, st>>>(ph, grad_colors, grad_texels, grad_bary, grad_fv, grad_fn,
                                                                   E, nchunks);
    }
    return (int)cudaGetLastError();
}

}  // namespace pert
