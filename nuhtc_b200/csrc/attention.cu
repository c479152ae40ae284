// Cosine-attention global pooling of AttentionRoIExtractor (levels >= start_level) for sm_100a.
//
// Replaces /root/reference/nuhtc/models/roi_extractors_cus.py:220-238: for every RoI take the feature vector at its
// centre cell (b, y, x) of the level, weight EVERY position of that image's level map by
//     sim = relu(cos(roi_vec, feat[b,:,h,w]) - thres) + thres
// and return mean_{h,w}(feat[b,:,h,w] * sim): one C-vector per RoI, which the reference broadcasts over the P x P bins
// and adds to the pooled features.
//
// The reference gathers a [HW, C] copy of the level per unique cell.  Here a CTA owns 64 RoIs of one image and sweeps
// the level once in tiles of 64 positions: S = Rn * F^T (register-tiled 64x64x64, 8x4 per thread), the relu, then
// O += S * F (second 64x64x64), i.e. the map is read once per 64 RoIs and S is never written to memory.  fp32 SIMT: the result must
// match the reference's fp32 CPU path to 1e-5, so no tensor cores / TF32.
#include "common.cuh"

namespace {

constexpr int AR = 64;        // RoIs per CTA
constexpr int AT = 64;        // positions per tile
constexpr int ATH = 256;      // threads: 16 RoI groups (4 RoIs each) x 16 position / channel lanes
constexpr int SF = 65;        // row stride of the feature tile (odd: the 16 lanes of a row group hit 16 banks)
constexpr int SR = AR + 4;    // row stride of the transposed RoI-vector and similarity tiles (16-byte aligned rows)

// rois sorted by image are not required: the CTA's 64 RoIs are taken from a per-image index list.
// Thread tile 4 RoIs x 4 positions (pass 1) / 4 RoIs x 4 channels (pass 2): 16 FMAs for 1 LDS.128 + 4 LDS.32.  The grid is
// compact (one CTA per 64 RoIs of an image, found through the per-image counts): round 1 launched ceil(K/64) x B CTAs of
// which 1 in 16 had work, and the working ones landed on the SMs in clumps (ncu: SMSPs active 54 % of the kernel's time).
// The accumulation orders (channels ascending in pass 1, positions ascending in pass 2) are unchanged, so are the results.
__global__ void __launch_bounds__(ATH) attn_pool_kernel(const float *__restrict__ feat, int H, int W, int C,
                                                        const float *__restrict__ rois, const int32_t *__restrict__ order,
                                                        const int32_t *__restrict__ img_start, int B, float inv_stride2, float thres,
                                                        int accumulate, float *__restrict__ out) {
    // C == 64 in every NuHTC config; the kernel is written for C <= 64 (channels beyond C are zero padded)
    extern __shared__ __align__(16) float attn_smem[];
    float (*s_rT)[SR] = reinterpret_cast<float (*)[SR]>(attn_smem);                      // normalised RoI vectors, [channel][roi]
    float (*s_sT)[SR] = reinterpret_cast<float (*)[SR]>(attn_smem + 64 * SR);            // similarities of the tile, [position][roi]
    float (*s_f)[SF] = reinterpret_cast<float (*)[SF]>(attn_smem + 64 * SR + AT * SR);   // raw features of the tile, [position][channel]
    float *s_inv = attn_smem + 64 * SR + AT * SR + AT * SF;                              // 1 / max(||f_p||, eps)
    int b = 0, u = blockIdx.x;
    for (; b < B; ++b) {   // CTA u of the compact grid -> (image, first RoI)
        const int nb = (img_start[b + 1] - img_start[b] + AR - 1) / AR;
        if (u < nb) break;
        u -= nb;
    }
    if (b == B) return;
    const int n_img = img_start[b + 1] - img_start[b];
    const int r0 = u * AR;
    const int nr = min(AR, n_img - r0);
    const int tid = threadIdx.x;
    const int HW = H * W;
    const float *fb = feat + (size_t)b * HW * C;
    const float eps = 1e-8f;

    // ---- RoI vectors: feature at the centre cell, normalised
    for (int i = tid; i < AR * 64; i += ATH) {
        const int r = i >> 6, c = i & 63;
        float v = 0.f;
        if (r < nr && c < C) {
            const float *roi = rois + (size_t)order[img_start[b] + r0 + r] * 5;
            // torch.div(x1 + x2, 2 * stride, rounding_mode='floor'), clamped to the map (2*stride is a power of two)
            int cx = (int)floorf(__fmul_rn(__fadd_rn(roi[1], roi[3]), inv_stride2));
            int cy = (int)floorf(__fmul_rn(__fadd_rn(roi[2], roi[4]), inv_stride2));
            cx = min(max(cx, 0), W - 1);
            cy = min(max(cy, 0), H - 1);
            v = __ldg(fb + ((size_t)cy * W + cx) * C + c);
        }
        s_rT[c][r] = v;
    }
    __syncthreads();
    if (tid < AR) {
        float n2 = 0.f;
        for (int c = 0; c < 64; ++c) n2 += s_rT[c][tid] * s_rT[c][tid];
        const float inv = 1.0f / fmaxf(sqrtf(n2), eps);
        for (int c = 0; c < 64; ++c) s_rT[c][tid] *= inv;
    }
    const int tr = (tid >> 4) * 4, tl = tid & 15;   // RoIs tr..tr+3; positions / channels tl, tl+16, tl+32, tl+48
    float o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;

    for (int p0 = 0; p0 < HW; p0 += AT) {
        __syncthreads(); // previous tile fully consumed (also orders the s_rT normalisation before its first use)
        if (C == 64) {
            // all 8 loads of a thread are in flight together (a rolled loop of dependent load -> store pairs cost one L2
            // round trip per element: 19 k cycles per tile, most of the kernel's time before this was unrolled)
            float4 v[AT * 16 / ATH];
#pragma unroll
            for (int u = 0; u < AT * 16 / ATH; ++u) {
                const int i4 = tid + u * ATH, p = i4 >> 4, c4 = (i4 & 15) * 4;
                v[u] = p0 + p < HW ? __ldg(reinterpret_cast<const float4 *>(fb + (size_t)(p0 + p) * 64 + c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < AT * 16 / ATH; ++u) {
                const int i4 = tid + u * ATH, p = i4 >> 4, c4 = (i4 & 15) * 4;
                s_f[p][c4] = v[u].x;
                s_f[p][c4 + 1] = v[u].y;
                s_f[p][c4 + 2] = v[u].z;
                s_f[p][c4 + 3] = v[u].w;
            }
        } else {
            for (int i = tid; i < AT * 64; i += ATH) {
                const int p = i >> 6, c = i & 63;
                s_f[p][c] = (p0 + p < HW && c < C) ? __ldg(fb + (size_t)(p0 + p) * C + c) : 0.f;
            }
        }
        __syncthreads();
        if (tid < AT) {
            float n2 = 0.f;
            for (int c = 0; c < 64; ++c) n2 += s_f[tid][c] * s_f[tid][c];
            s_inv[tid] = 1.0f / fmaxf(sqrtf(n2), eps);
        }
        __syncthreads();
        // pass 1: S[r][p] = relu(<Rn[r], F[p]> * inv[p] - thres) + thres
        float s[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
        for (int c = 0; c < 64; ++c) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&s_rT[c][tr]);
            const float a[4] = {a0.x, a0.y, a0.z, a0.w};
            float f[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) f[j] = s_f[tl + 16 * j][c];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) s[i][j] = fmaf(a[i], f[j], s[i][j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int p = tl + 16 * j;
            const bool live = p0 + p < HW;
            const float inv = s_inv[p];
            float v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float cosv = s[i][j] * inv;
                v[i] = live ? fmaxf(cosv - thres, 0.f) + thres : 0.f;
            }
            *reinterpret_cast<float4 *>(&s_sT[p][tr]) = make_float4(v[0], v[1], v[2], v[3]);
        }
        __syncthreads();
        // pass 2: O[r][c] += sum_p S[r][p] * F[p][c]
#pragma unroll 8
        for (int p = 0; p < AT; ++p) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&s_sT[p][tr]);
            const float a[4] = {a0.x, a0.y, a0.z, a0.w};
            float f[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) f[j] = s_f[p][tl + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) o[i][j] = fmaf(a[i], f[j], o[i][j]);
        }
    }
    const float inv_hw = 1.0f / (float)HW;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = tr + i;
        if (r >= nr) continue;
        float *dst = out + (size_t)order[img_start[b] + r0 + r] * C;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = tl + 16 * j;
            if (c < C) {
                const float v = o[i][j] * inv_hw;
                dst[c] = accumulate ? dst[c] + v : v;
            }
        }
    }
}

// RoIs grouped by image: counting sort (K is small), stable so that results do not depend on scheduling
__global__ void attn_count_kernel(const float *__restrict__ rois, int K, int B, int32_t *cnt, int32_t *status) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int b = (int)rois[(size_t)k * 5];
    if (b < 0 || b >= B) {
        atomicExch(status, 2);
        return;
    }
    atomicAdd(cnt + b, 1);
}
__global__ void attn_scan_kernel(const int32_t *cnt, int B, int32_t *img_start) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int acc = 0;
        for (int b = 0; b < B; ++b) {
            img_start[b] = acc;
            acc += cnt[b];
        }
        img_start[B] = acc;
    }
}
__global__ void attn_fill_kernel(const float *__restrict__ rois, int K, int B, const int32_t *__restrict__ img_start,
                                 int32_t *__restrict__ cursor, int32_t *__restrict__ order) {
    // the order of the RoIs inside an image's list is irrelevant: every RoI's output depends on that RoI alone
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int b = (int)rois[(size_t)k * 5];
    if (b < 0 || b >= B) return;
    order[img_start[b] + atomicAdd(cursor + b, 1)] = k;
}

} // namespace

NUHTC_API size_t nuhtc_attention_pool_workspace_bytes(int K, int B) {
    return align_up(sizeof(int32_t) * (size_t)(K > 0 ? K : 1), 256) + 2 * align_up(sizeof(int32_t) * (size_t)(B + 1), 256) + 256;
}

NUHTC_API int nuhtc_attention_pool(const float *feat_nhwc, int B, int H, int W, int C, const float *rois, int K, float stride,
                                   float thres, int accumulate, float *out, int32_t *status, void *ws, size_t ws_bytes,
                                   void *stream) {
    NUHTC_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && C >= 1 && C <= 64 && K >= 0, "attention_pool: bad sizes (C must be <= 64)");
    NUHTC_CHECK_ARG(stride > 0.f && status != nullptr, "attention_pool: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    NUHTC_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    if (K == 0) return NUHTC_OK;
    NUHTC_CHECK_ARG(feat_nhwc && rois && out && ws, "attention_pool: null pointer");
    if (ws_bytes < nuhtc_attention_pool_workspace_bytes(K, B)) {
        nuhtc_set_error("attention_pool: workspace too small");
        return NUHTC_EWORKSPACE;
    }
    char *p = (char *)ws;
    int32_t *order = (int32_t *)p;
    p += align_up(sizeof(int32_t) * (size_t)K, 256);
    int32_t *cnt = (int32_t *)p;
    p += align_up(sizeof(int32_t) * (size_t)(B + 1), 256);
    int32_t *img_start = (int32_t *)p;
    NUHTC_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (B + 1), st));
    attn_count_kernel<<<(K + 255) / 256, 256, 0, st>>>(rois, K, B, cnt, status);
    attn_scan_kernel<<<1, 32, 0, st>>>(cnt, B, img_start);
    NUHTC_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (B + 1), st)); // reused as the fill cursors
    attn_fill_kernel<<<(K + 255) / 256, 256, 0, st>>>(rois, K, B, img_start, cnt, order);
    const unsigned grid = (unsigned)((K + AR - 1) / AR + B);   // >= sum over images of ceil(n_b / AR); the few spare CTAs exit at once
    constexpr size_t smem = sizeof(float) * (64 * SR + AT * SR + AT * SF + AT);
    static bool attr_done[kNuhtcMaxDevices] = {false};
    if (!attr_done[nuhtc_device()]) {
        NUHTC_CUDA(cudaFuncSetAttribute(attn_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done[nuhtc_device()] = true;
    }
    attn_pool_kernel<<<grid, ATH, smem, st>>>(feat_nhwc, H, W, C, rois, order, img_start, B, 1.0f / (2.0f * stride), thres, accumulate, out);
    NUHTC_LAUNCH_CHECK();
    return NUHTC_OK;
}
