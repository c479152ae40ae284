// Connected components of a batch of binary masks for sm_100a: the host part of the watershed proposals.
// Replaces, per image, the scipy / skimage / Python-loop tail of HybridTaskCascadeRoIHead_Cus._watershed_proposal
// (nuhtc/models/htc_roi_head_cus.py:303-335, after a device->host copy of every mask):
//     ndi.binary_fill_holes -> ndi.distance_transform_edt -> ndi.label(distance > 0.25) -> skimage watershed(-distance,
//     markers, mask) -> torch.unique / relabel loop -> one-hot [n,H,W] -> areas -> area filter -> _inst_mask_to_bbox loop
// The Euclidean distance of a foreground pixel is >= 1, so `distance > 0.25` is the filled mask itself, every mask pixel
// is a marker and the watershed returns the markers: the instances ARE the 4-connected components of the hole-filled mask,
// numbered in raster order of their first pixel (scipy's label order, kept by torch.unique).  What is computed here:
//   1. union-find labelling of foreground AND background (4-connectivity, root = smallest pixel index);
//   2. holes = background components that do not touch the frame -> filled mask;
//   3. union-find labelling of the filled mask, per-component area and tight box by atomics on the root;
//   4. one CTA per image walks the pixels in raster order and emits the boxes (x0, y0, x1+1, y1+1, 1.0) of the
//      components with min_area < area < max_area, in label order.
// Label-equivalence union-find with atomicMin (Komura 2015 / Playne-Hawick 2018); all in global memory (a 512x512 frame
// has 1 MB of labels), the masks are tiny next to the RoI stage.
#include "common.cuh"

namespace {

__device__ __forceinline__ int uf_find(const int *L, int i) {
    int p = L[i];
    while (p != i) {
        i = p;
        p = L[i];
    }
    return i;
}
__device__ __forceinline__ void uf_union(int *L, int a, int b) {
    while (true) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a < b) {
            const int t = a;
            a = b;
            b = t;
        }
        const int old = atomicMin(&L[a], b);   // hang the larger root under the smaller one
        if (old == a) return;
        a = old;
    }
}

// Labels start as the first pixel of the horizontal run inside the pixel's 32-pixel segment (one ballot): a run of equal
// pixels is connected without a single atomic, and the union pass only has to stitch segment boundaries and the places where
// a run meets a NEW run of the row above.  p % 32 == lane because the block size is a multiple of 32.
__device__ __forceinline__ unsigned run_starts(bool start) { return __ballot_sync(0xffffffffu, start); }

__global__ void __launch_bounds__(256) ccl_init_kernel(const float *__restrict__ mask, int64_t n, int W, uint8_t *__restrict__ val,
                                                      int *__restrict__ label) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool in = p < n;
    const int v = in ? (mask[p] != 0.f) : 0;
    const int x = in ? (int)(p % W) : 0;
    const bool start = !in || lane == 0 || x == 0 || ((mask[p - 1] != 0.f) != (v != 0));
    const unsigned m = run_starts(start);
    if (!in) return;
    const int s = 31 - __clz(m & (0xffffffffu >> (31 - lane)));
    val[p] = (uint8_t)v;
    label[p] = (int)p - (lane - s);
}
__global__ void __launch_bounds__(256) ccl_runs_kernel(const uint8_t *__restrict__ val, int64_t n, int W, int *__restrict__ label) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool in = p < n;
    const int v = in ? val[p] : 0;
    const int x = in ? (int)(p % W) : 0;
    const bool start = !in || lane == 0 || x == 0 || val[p - 1] != v;
    const unsigned m = run_starts(start);
    if (!in) return;
    const int s = 31 - __clz(m & (0xffffffffu >> (31 - lane)));
    label[p] = (int)p - (lane - s);
}
// stitches: the left neighbour across a segment boundary, and the upper neighbour unless the pixel's left neighbour already
// made that connection (left and upper-left of the same value: p-1 ~ p-1-W by its own stitch, and both rows are runs)
__global__ void __launch_bounds__(256) ccl_union_kernel(const uint8_t *__restrict__ val, int *__restrict__ label, int64_t n, int H,
                                                       int W, int fg_only) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int v = val[p];
    if (fg_only && !v) return;
    const int lane = threadIdx.x & 31;
    const int x = (int)(p % W), y = (int)((p / W) % H);
    const bool left = x > 0 && val[p - 1] == v;
    if (left && lane == 0) uf_union(label, (int)p, (int)p - 1);
    if (y > 0 && val[p - W] == v) {
        const bool covered = left && val[p - W - 1] == v;
        if (!covered) uf_union(label, (int)p, (int)p - W);
    }
}
__global__ void ccl_border_kernel(const uint8_t *__restrict__ val, int *__restrict__ label, int64_t n, int H, int W,
                                  uint8_t *__restrict__ open_bg) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int r = uf_find(label, (int)p);
    label[p] = r;
    const int x = (int)(p % W), y = (int)((p / W) % H);
    if (!val[p] && (x == 0 || y == 0 || x == W - 1 || y == H - 1)) open_bg[r] = 1;
}
// filled mask + reset of the per-root statistics for the second pass
__global__ void ccl_fill_kernel(uint8_t *__restrict__ val, const int *__restrict__ label, const uint8_t *__restrict__ open_bg, int64_t n,
                                int *__restrict__ area, int *__restrict__ x0, int *__restrict__ y0, int *__restrict__ x1,
                                int *__restrict__ y1) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int r = label[p];   // the border kernel left every pixel pointing at its final root
    const bool fg = val[p] || !open_bg[r];
    val[p] = fg;
    area[p] = 0;
    x0[p] = 0x7fffffff;
    y0[p] = 0x7fffffff;
    x1[p] = -1;
    y1[p] = -1;
}
// per-component area and box: one set of atomics per run segment (its first lane), not per pixel
__global__ void __launch_bounds__(256) ccl_stats_kernel(const uint8_t *__restrict__ val, int *__restrict__ label, int64_t n, int H, int W,
                                                       int *__restrict__ area, int *__restrict__ x0, int *__restrict__ y0,
                                                       int *__restrict__ x1, int *__restrict__ y1) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool in = p < n;
    const int v = in ? val[p] : 0;
    const int x = in ? (int)(p % W) : 0;
    const bool start = !in || lane == 0 || x == 0 || val[p - 1] != v;
    const unsigned m = run_starts(start);
    if (!in || !v) return;
    const int r = uf_find(label, (int)p);
    label[p] = r;
    if (!start) return;
    const unsigned above = lane == 31 ? 0u : (m >> (lane + 1));
    const int len = above ? __ffs(above) : 32 - lane;          // pixels of this run inside the segment (they are all in range)
    const int y = (int)((p / W) % H);
    atomicAdd(area + r, len);
    atomicMin(x0 + r, x);
    atomicMax(x1 + r, x + len - 1);
    atomicMin(y0 + r, y);
    atomicMax(y1 + r, y);
}
// one CTA per image: component roots in raster order -> boxes.  Pass 1 counts the qualifying roots of every 32-pixel
// segment, one block scan turns the counts into ranks, pass 2 writes.  kEmitSeg segments (32 K pixels... x32) per round.
constexpr int kEmitSeg = 8192;
__global__ void __launch_bounds__(1024) ccl_emit_kernel(const uint8_t *__restrict__ val, const int *__restrict__ label, int HW,
                                                        const int *__restrict__ area, const int *__restrict__ x0,
                                                        const int *__restrict__ y0, const int *__restrict__ x1,
                                                        const int *__restrict__ y1, int min_area, int max_area, int max_boxes,
                                                        float *__restrict__ boxes, int32_t *__restrict__ counts) {
    __shared__ int s_cnt[kEmitSeg];
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t off = (int64_t)b * HW;
    if (tid == 0) s_base = 0;
    __syncthreads();
    auto qualifies = [&](int p) {
        if (p >= HW) return false;
        const int64_t g = off + p;
        if (!val[g] || label[g] != (int)g) return false;
        const int ar = area[g];
        return ar > min_area && ar < max_area;
    };
    for (int seg0 = 0; seg0 * 32 < HW; seg0 += kEmitSeg) {
        const int nseg = min(kEmitSeg, (HW - seg0 * 32 + 31) / 32);
        for (int sg = warp; sg < nseg; sg += 32) {
            const unsigned m = __ballot_sync(0xffffffffu, qualifies((seg0 + sg) * 32 + lane));
            if (lane == 0) s_cnt[sg] = __popc(m);
        }
        __syncthreads();
        // exclusive scan of s_cnt[0..nseg): every thread owns a contiguous run of kEmitSeg / 1024 entries
        constexpr int PER = kEmitSeg / 1024;
        int loc[PER], sum = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int e = tid * PER + i;
            loc[i] = e < nseg ? s_cnt[e] : 0;
            sum += loc[i];
        }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += v;
            }
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        int run = s_base + s_warp[warp] + incl - sum;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int e = tid * PER + i;
            if (e < nseg) s_cnt[e] = run;
            run += loc[i];
        }
        __syncthreads();
        for (int sg = warp; sg < nseg; sg += 32) {
            const int p = (seg0 + sg) * 32 + lane;
            const bool take = qualifies(p);
            const unsigned m = __ballot_sync(0xffffffffu, take);
            const int rank = s_cnt[sg] + __popc(m & ((1u << lane) - 1u));
            if (take && rank < max_boxes) {
                const int64_t g = off + p;
                float *o = boxes + ((size_t)b * max_boxes + rank) * 5;
                o[0] = (float)x0[g];
                o[1] = (float)y0[g];
                o[2] = (float)(x1[g] + 1);
                o[3] = (float)(y1[g] + 1);
                o[4] = 1.0f;
            }
        }
        __syncthreads();
        if (tid == 1023) s_base = run;      // the last thread's running total = all roots so far
        __syncthreads();
    }
    if (tid == 0) counts[b] = s_base;
}

struct CclWs {
    size_t val, open_bg, label, area, x0, y0, x1, y1, total;
};
CclWs ccl_layout(int64_t n) {
    CclWs w;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        const size_t at = o;
        o = (o + bytes + 255) / 256 * 256;
        return at;
    };
    w.val = take(n);
    w.open_bg = take(n);
    w.label = take(4 * n);
    w.area = take(4 * n);
    w.x0 = take(4 * n);
    w.y0 = take(4 * n);
    w.x1 = take(4 * n);
    w.y1 = take(4 * n);
    w.total = o;
    return w;
}

} // namespace

NUHTC_API size_t nuhtc_mask_components_workspace_bytes(int B, int H, int W) {
    if (B < 1 || H < 1 || W < 1) return 256;
    return ccl_layout((int64_t)B * H * W).total;
}

NUHTC_API int nuhtc_mask_components(const float *mask, int B, int H, int W, int min_area, int max_area, int max_boxes,
                                    float *boxes, int32_t *counts, uint8_t *filled, void *ws_, size_t ws_bytes, void *stream) {
    NUHTC_CHECK_ARG(B >= 0 && H >= 1 && W >= 1 && max_boxes >= 1, "mask_components: bad sizes");
    if (B == 0) return NUHTC_OK;
    const int64_t n = (int64_t)B * H * W;
    NUHTC_CHECK_ARG(n < (1ll << 31), "mask_components: more than 2^31 pixels");
    NUHTC_CHECK_ARG(mask && boxes && counts && ws_, "mask_components: null pointer");
    const CclWs w = ccl_layout(n);
    if (ws_bytes < w.total) {
        nuhtc_set_error("mask_components: workspace %zu < %zu bytes", ws_bytes, w.total);
        return NUHTC_EWORKSPACE;
    }
    char *ws = (char *)ws_;
    uint8_t *val = (uint8_t *)(ws + w.val), *open_bg = (uint8_t *)(ws + w.open_bg);
    int *label = (int *)(ws + w.label), *area = (int *)(ws + w.area);
    int *x0 = (int *)(ws + w.x0), *y0 = (int *)(ws + w.y0), *x1 = (int *)(ws + w.x1), *y1 = (int *)(ws + w.y1);
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned nb = (unsigned)((n + 255) / 256);
    NUHTC_CUDA(cudaMemsetAsync(open_bg, 0, n, st));
    ccl_init_kernel<<<nb, 256, 0, st>>>(mask, n, W, val, label);
    ccl_union_kernel<<<nb, 256, 0, st>>>(val, label, n, H, W, 0);
    ccl_border_kernel<<<nb, 256, 0, st>>>(val, label, n, H, W, open_bg);
    ccl_fill_kernel<<<nb, 256, 0, st>>>(val, label, open_bg, n, area, x0, y0, x1, y1);
    ccl_runs_kernel<<<nb, 256, 0, st>>>(val, n, W, label);
    ccl_union_kernel<<<nb, 256, 0, st>>>(val, label, n, H, W, 1);
    ccl_stats_kernel<<<nb, 256, 0, st>>>(val, label, n, H, W, area, x0, y0, x1, y1);
    ccl_emit_kernel<<<B, 1024, 0, st>>>(val, label, H * W, area, x0, y0, x1, y1, min_area, max_area, max_boxes, boxes, counts);
    if (filled) NUHTC_CUDA(cudaMemcpyAsync(filled, val, n, cudaMemcpyDeviceToDevice, st));
    NUHTC_LAUNCH_CHECK();
    return NUHTC_OK;
}
