// Mask paste for sm_100a: fused bilinear resample + threshold + store (uint8 / bit rows / fp32).
//
// Replaces `_do_paste_mask(masks, boxes, img_h, img_w, skip_empty=False)` followed by
// `(masks_chunk >= threshold).to(bool)` of FCNMaskHead.get_seg_masks
//   /root/reference/thirdparty/mmdetection/mmdet/models/roi_heads/mask_heads/fcn_mask_head.py:292-306,344-412
// without materialising the [N,H,W,2] sampling grid or the fp32 [N,H,W] intermediate: the
// normalised coordinate of every output pixel is recomputed from the box in the reference's own
// float op order, and ATen's grid_sampler_2d (bilinear, zeros padding, align_corners=False) is
// applied in registers.
//
// The dense output contract ([N,H,W] per mask) makes this a store stream: a nucleus covers ~1 % of its
// 256x256 frame and everything outside the box's reach is exactly zero (zeros padding).  The dense uint8 kind does
// both parts in one kernel (FUSE_FILL below: 0.103 ms for 8000 frames, 0.82 of the measured HBM peak, against 0.148 ms);
// the other kinds split the work in two launches on the same stream:
//   fill   : a grid-stride 128-bit streaming zero fill of the whole output (HBM-write bound, no per-mask
//            prologue on its critical path)
//   sparse : one CTA per mask evaluates only the 16-pixel segments (64-pixel words for bit rows) that the
//            box can reach, staged probability map in shared memory, and overwrites them; it also
//            reduces the per-mask area and tight bounding box that the mask NMS consumes.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kPasteThreads = 128;
constexpr int kMaxMaskElems = 64 * 64; // staged probability map (28x28 in every NuHTC config)

// normalised grid coordinate of pixel centre p+0.5 for a box side [a0,a1] (fcn_mask_head.py:388-400)
__device__ __forceinline__ float grid_coord(int p, float a0, float a1) {
    float g = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn((float)p + 0.5f, a0), __fsub_rn(a1, a0)), 2.0f), 1.0f);
    if (isinf(g)) g = 0.f;
    return g;
}
// ATen grid_sampler unnormalize, align_corners=False: ((g + 1) * size - 1) / 2
__device__ __forceinline__ float unnormalize(float g, int size) {
    return __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.0f), (float)size), 1.0f), 0.5f);
}

struct AxisTap {
    int i0;       // floor index; i0+1 is the other tap
    float w0, w1; // weights of i0 and i0+1
};
__device__ __forceinline__ AxisTap axis_tap(int p, float a0, float a1, int size) {
    const float s = unnormalize(grid_coord(p, a0, a1), size);
    const float f = floorf(s);
    AxisTap t;
    // NaN (degenerate box) or far-away coordinates: push the taps out of range => contributes 0
    t.i0 = (s >= -2.0f && s <= (float)size + 1.0f) ? (int)f : -4;
    t.w1 = __fsub_rn(s, f);
    t.w0 = __fsub_rn(__fadd_rn(f, 1.0f), s);
    return t;
}

__device__ __forceinline__ float sample(const float *m, int mh, int mw, const AxisTap ty, const AxisTap tx) {
    float v = 0.f;
    const bool y0 = ty.i0 >= 0 && ty.i0 < mh, y1 = ty.i0 + 1 >= 0 && ty.i0 + 1 < mh;
    const bool x0 = tx.i0 >= 0 && tx.i0 < mw, x1 = tx.i0 + 1 >= 0 && tx.i0 + 1 < mw;
    if (y0 && x0) v += m[ty.i0 * mw + tx.i0] * (tx.w0 * ty.w0);
    if (y0 && x1) v += m[ty.i0 * mw + tx.i0 + 1] * (tx.w1 * ty.w0);
    if (y1 && x0) v += m[(ty.i0 + 1) * mw + tx.i0] * (tx.w0 * ty.w1);
    if (y1 && x1) v += m[(ty.i0 + 1) * mw + tx.i0 + 1] * (tx.w1 * ty.w1);
    return v;
}

// conservative pixel range [lo,hi) outside of which every sample is exactly zero
__device__ __forceinline__ void active_range(float a0, float a1, int size, int extent, int &lo, int &hi) {
    const float side = a1 - a0;
    if (!(side > 0.f) || isinf(side)) { // degenerate / NaN box: evaluate everything literally
        lo = 0;
        hi = extent;
        return;
    }
    const float pad = side / (float)size + 2.0f; // half a mask pixel would do; be generous
    const float l = floorf(a0 - pad), h = ceilf(a1 + pad);
    lo = l < 0.f ? 0 : (l > (float)extent ? extent : (int)l);
    hi = h < 0.f ? 0 : (h > (float)extent ? extent : (int)h);
}

// ---- launch 1: zero fill (n16 16-byte chunks + a byte tail), grid-stride, 4 stores in flight per thread
__global__ void __launch_bounds__(256) fill_zero_kernel(uint4 *__restrict__ p, size_t n16, uint8_t *__restrict__ tail, int ntail) {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        st_stream_u4(p + i, z);
        st_stream_u4(p + i + stride, z);
        st_stream_u4(p + i + 2 * stride, z);
        st_stream_u4(p + i + 3 * stride, z);
    }
    for (; i < n16; i += stride) st_stream_u4(p + i, z);
    if (blockIdx.x == 0 && (int)threadIdx.x < ntail) tail[threadIdx.x] = 0;
}

// ---- launch 2: the reachable region of every mask.  FUSE_FILL (dense uint8 frames with 16-byte rows only): the CTA also
// zero-fills the part of its own frame that the box cannot reach, so the frame is written exactly once and the separate
// fill launch is skipped; the zero stores are issued first and drain while the reachable segments are evaluated.
template <int KIND, bool FUSE_FILL = false>
__global__ void __launch_bounds__(kPasteThreads) paste_sparse_kernel(const float *__restrict__ probs, const float *__restrict__ boxes,
                                                                     int mh, int mw, int H, int W, float thr, void *__restrict__ outv,
                                                                     int32_t *__restrict__ area, int32_t *__restrict__ bbox,
                                                                     unsigned long long *__restrict__ bits2 = nullptr) {
    __shared__ float s_m[kMaxMaskElems];
    __shared__ int s_red[5][kPasteThreads / 32];
    const int n = blockIdx.x, tid = threadIdx.x;
    const float bx0 = boxes[4 * n], by0 = boxes[4 * n + 1], bx1 = boxes[4 * n + 2], by1 = boxes[4 * n + 3];
    int ax0, ax1, ay0, ay1;
    active_range(bx0, bx1, mw, W, ax0, ax1);
    active_range(by0, by1, mh, H, ay0, ay1);
    if (KIND != NUHTC_PASTE_PROB && !(thr > 0.f)) { // 0 >= thr holds for the zero padding too: evaluate every pixel
        ax0 = ay0 = 0;
        ax1 = W;
        ay1 = H;
    }
    const bool any = ax1 > ax0 && ay1 > ay0;
    int cnt = 0, minx = 1 << 30, miny = 1 << 30, maxx = -1, maxy = -1;
    if (FUSE_FILL) { // 16-pixel segments outside the reachable rectangle (W % 16 == 0, checked by the launcher)
        const int spr = W / 16, nseg = H * spr;
        const int s0 = any ? ax0 / 16 : 0, s1 = any ? (ax1 + 15) / 16 : 0; // segment columns the evaluation below overwrites
        char *frame = (char *)outv + (size_t)n * H * W;
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int i = tid; i < nseg; i += kPasteThreads) {
            const int y = i / spr, sx = i - y * spr;
            if (!(any && y >= ay0 && y < ay1 && sx >= s0 && sx < s1)) st_stream_u4(frame + (size_t)i * 16, z);
        }
        if (bits2) { // second output of the same evaluation: the bit rows.  Zeroed whole first, then every 16-pixel segment of
                     // the reachable rectangle stores its own 16 bits (bit x % 64 of word x / 64 = the uint16 at byte x / 8)
            uint4 *bf = reinterpret_cast<uint4 *>(bits2 + (size_t)n * H * ((W + 63) / 64));
            const int n16 = H * ((W + 63) / 64) / 2;
            for (int i = tid; i < n16; i += kPasteThreads) bf[i] = z;
            __syncthreads();
        }
    }
    if (any) { // uniform over the CTA
        const float *src = probs + (size_t)n * mh * mw;
        for (int i = tid; i < mh * mw; i += kPasteThreads) s_m[i] = __ldg(src + i);
        __syncthreads();
        const int nrows = ay1 - ay0;
        if (KIND == NUHTC_PASTE_BITS) {
            const int wpr = (W + 63) / 64;
            unsigned long long *out = (unsigned long long *)outv + (size_t)n * H * wpr;
            const int w0 = ax0 >> 6, nw = ((ax1 + 63) >> 6) - w0;
            // 4 lanes share one 64-pixel word (16 pixels each) so that a nucleus keeps a whole CTA busy
            const int sub = tid & 3;
            const int ntask = nrows * nw;
            for (int base = 0; base < ntask; base += kPasteThreads / 4) { // trip count uniform over the CTA (shuffles below)
                const int i = base + (tid >> 2);
                const bool live = i < ntask;
                const int y = ay0 + (live ? i / nw : 0), xw = (w0 + (live ? i % nw : 0)) * 64;
                unsigned long long part = 0ull;
                if (live) {
                    const AxisTap ty = axis_tap(y, by0, by1, mh);
#pragma unroll 4
                    for (int u = 0; u < 16; ++u) {
                        const int x = xw + sub * 16 + u;
                        if (x >= ax0 && x < ax1 && x < W) {
                            const float v = sample(s_m, mh, mw, ty, axis_tap(x, bx0, bx1, mw));
                            if (v >= thr) {
                                part |= 1ull << (sub * 16 + u);
                                minx = min(minx, x);
                                maxx = max(maxx, x);
                            }
                        }
                    }
                }
                part |= __shfl_xor_sync(0xffffffffu, part, 1);
                part |= __shfl_xor_sync(0xffffffffu, part, 2);
                if (live && sub == 0) {
                    out[(size_t)y * wpr + (xw >> 6)] = part;
                    if (part) {
                        cnt += __popcll(part);
                        miny = min(miny, y);
                        maxy = max(maxy, y);
                    }
                }
            }
        } else {
            // 16 pixels (BIN: 16 bytes) or 4 pixels (PROB: 16 bytes) per thread per step
            constexpr int PX = KIND == NUHTC_PASTE_BIN ? 16 : 4;
            const size_t total = (size_t)H * W;
            char *outb = (char *)outv + (size_t)n * total * (KIND == NUHTC_PASTE_BIN ? 1 : 4);
            const bool vec_ok = (W % PX == 0) && (((uintptr_t)outb) % 16 == 0);
            if (vec_ok) {
                const int s0 = ax0 / PX, ns = (ax1 + PX - 1) / PX - s0;
                for (int i = tid; i < nrows * ns; i += kPasteThreads) {
                    const int y = ay0 + i / ns, xs = (s0 + i % ns) * PX;
                    const AxisTap ty = axis_tap(y, by0, by1, mh);
                    float v[PX];
#pragma unroll
                    for (int u = 0; u < PX; ++u) v[u] = sample(s_m, mh, mw, ty, axis_tap(xs + u, bx0, bx1, mw));
                    const size_t seg = ((size_t)y * W + xs) / PX;
                    if (KIND == NUHTC_PASTE_BIN) {
                        uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                        for (int u = 0; u < PX; ++u) {
                            if (v[u] >= thr) {
                                w[u >> 2] |= 1u << (8 * (u & 3));
                                ++cnt;
                                minx = min(minx, xs + u);
                                maxx = max(maxx, xs + u);
                                miny = min(miny, y);
                                maxy = max(maxy, y);
                            }
                        }
                        st_stream_u4(outb + seg * 16, make_uint4(w[0], w[1], w[2], w[3]));
                        if (FUSE_FILL && bits2) {
                            uint32_t b16 = 0u;
#pragma unroll
                            for (int u = 0; u < PX; ++u) b16 |= ((w[u >> 2] >> (8 * (u & 3))) & 1u) << u;
                            reinterpret_cast<uint16_t *>(bits2 + (size_t)n * H * ((W + 63) / 64))[((size_t)y * ((W + 63) / 64) * 64 + xs) / 16] =
                                (uint16_t)b16;
                        }
                    } else {
                        st_stream_f4((float *)outb + seg * 4, make_float4(v[0], v[1], v[2], v[3]));
                    }
                }
            } else {
                const int nc = ax1 - ax0;
                for (int i = tid; i < nrows * nc; i += kPasteThreads) {
                    const int y = ay0 + i / nc, x = ax0 + i % nc;
                    const float v = sample(s_m, mh, mw, axis_tap(y, by0, by1, mh), axis_tap(x, bx0, bx1, mw));
                    if (KIND == NUHTC_PASTE_BIN) {
                        const bool on = v >= thr;
                        ((uint8_t *)outb)[(size_t)y * W + x] = on ? 1 : 0;
                        if (on) {
                            ++cnt;
                            minx = min(minx, x);
                            maxx = max(maxx, x);
                            miny = min(miny, y);
                            maxy = max(maxy, y);
                        }
                    } else {
                        ((float *)outb)[(size_t)y * W + x] = v;
                    }
                }
            }
        }
    }

    if (KIND != NUHTC_PASTE_PROB && (area || bbox)) {
        const int lane = tid & 31, warp = tid >> 5;
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        minx = __reduce_min_sync(0xffffffffu, minx);
        miny = __reduce_min_sync(0xffffffffu, miny);
        maxx = __reduce_max_sync(0xffffffffu, maxx);
        maxy = __reduce_max_sync(0xffffffffu, maxy);
        if (lane == 0) {
            s_red[0][warp] = cnt;
            s_red[1][warp] = minx;
            s_red[2][warp] = miny;
            s_red[3][warp] = maxx;
            s_red[4][warp] = maxy;
        }
        __syncthreads();
        if (tid == 0) {
            int c = 0, x0 = 1 << 30, y0 = 1 << 30, x1 = -1, y1 = -1;
            for (int w = 0; w < kPasteThreads / 32; ++w) {
                c += s_red[0][w];
                x0 = min(x0, s_red[1][w]);
                y0 = min(y0, s_red[2][w]);
                x1 = max(x1, s_red[3][w]);
                y1 = max(y1, s_red[4][w]);
            }
            if (area) area[n] = c;
            if (bbox) {
                bbox[4 * n + 0] = c ? x0 : 0;
                bbox[4 * n + 1] = c ? y0 : 0;
                bbox[4 * n + 2] = c ? x1 + 1 : 0;
                bbox[4 * n + 3] = c ? y1 + 1 : 0;
            }
        }
    }
}

} // namespace

NUHTC_API int nuhtc_paste_masks_dense_bits(const float *probs, const float *boxes, int N, int mh, int mw, int img_h, int img_w,
                                           float thr, uint8_t *dense, uint64_t *bits, int32_t *area, int32_t *bbox, void *stream) {
    NUHTC_CHECK_ARG(N >= 0 && mh >= 1 && mw >= 1 && img_h >= 1 && img_w >= 1, "paste: bad sizes");
    NUHTC_CHECK_ARG(mh * mw <= kMaxMaskElems, "paste: mask %dx%d larger than the staged maximum", mh, mw);
    NUHTC_CHECK_ARG(img_w % 16 == 0 && (img_h * ((img_w + 63) / 64)) % 2 == 0, "paste_dense_bits: img_w must be a multiple of 16");
    if (N == 0) return NUHTC_OK;
    NUHTC_CHECK_ARG(probs && boxes && dense && bits, "paste: null pointer");
    NUHTC_CHECK_ARG(((uintptr_t)dense) % 16 == 0 && ((uintptr_t)bits) % 16 == 0, "paste: outputs must be 16-byte aligned");
    paste_sparse_kernel<NUHTC_PASTE_BIN, true><<<N, kPasteThreads, 0, (cudaStream_t)stream>>>(
        probs, boxes, mh, mw, img_h, img_w, thr, dense, area, bbox, (unsigned long long *)bits);
    NUHTC_LAUNCH_CHECK();
    return NUHTC_OK;
}

NUHTC_API int nuhtc_paste_masks(const float *probs, const float *boxes, int N, int mh, int mw, int img_h, int img_w, float thr,
                                int out_kind, void *out, int32_t *area, int32_t *bbox, void *stream) {
    NUHTC_CHECK_ARG(N >= 0 && mh >= 1 && mw >= 1 && img_h >= 1 && img_w >= 1, "paste: bad sizes");
    NUHTC_CHECK_ARG(mh * mw <= kMaxMaskElems, "paste: mask %dx%d larger than the staged maximum", mh, mw);
    NUHTC_CHECK_ARG(out_kind >= NUHTC_PASTE_PROB && out_kind <= NUHTC_PASTE_BITS, "paste: bad out_kind %d", out_kind);
    if (N == 0) return NUHTC_OK;
    NUHTC_CHECK_ARG(probs && boxes && out, "paste: null pointer");
    NUHTC_CHECK_ARG(((uintptr_t)out) % 16 == 0, "paste: output must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    size_t bytes;
    if (out_kind == NUHTC_PASTE_PROB) bytes = (size_t)N * img_h * img_w * 4;
    else if (out_kind == NUHTC_PASTE_BIN) bytes = (size_t)N * img_h * img_w;
    else bytes = (size_t)N * img_h * ((img_w + 63) / 64) * 8;
    const size_t n16 = bytes / 16;
    const int ntail = (int)(bytes % 16);
    size_t want = (n16 + 4 * 256 - 1) / (4 * 256);
    const size_t cap = (size_t)nuhtc_sm_count() * 8; // a multiple of the SM count, 8 resident CTAs each
    const unsigned fgrid = (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
    static const bool fuse_ok = !(getenv("NUHTC_PASTE_FUSE") && getenv("NUHTC_PASTE_FUSE")[0] == '0');
    if (out_kind == NUHTC_PASTE_BIN && img_w % 16 == 0 && fuse_ok) { // every frame starts 16-byte aligned: one pass
        paste_sparse_kernel<NUHTC_PASTE_BIN, true><<<N, kPasteThreads, 0, st>>>(probs, boxes, mh, mw, img_h, img_w, thr, out, area, bbox);
        NUHTC_LAUNCH_CHECK();
        return NUHTC_OK;
    }
    fill_zero_kernel<<<fgrid, 256, 0, st>>>((uint4 *)out, n16, (uint8_t *)out + n16 * 16, ntail);
    switch (out_kind) {
        case NUHTC_PASTE_PROB:
            paste_sparse_kernel<NUHTC_PASTE_PROB><<<N, kPasteThreads, 0, st>>>(probs, boxes, mh, mw, img_h, img_w, thr, out, area, bbox);
            break;
        case NUHTC_PASTE_BIN:
            paste_sparse_kernel<NUHTC_PASTE_BIN><<<N, kPasteThreads, 0, st>>>(probs, boxes, mh, mw, img_h, img_w, thr, out, area, bbox);
            break;
        default:
            paste_sparse_kernel<NUHTC_PASTE_BITS><<<N, kPasteThreads, 0, st>>>(probs, boxes, mh, mw, img_h, img_w, thr, out, area, bbox);
    }
    NUHTC_LAUNCH_CHECK();
    return NUHTC_OK;
}
