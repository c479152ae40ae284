// Detection glue of the HTC test path for sm_100a: the element-wise steps between the heavy ops, which in the reference
// are dozens of small torch kernels per batch.  Four kernels:
//   delta2bbox            mmdet/core/bbox/coder/delta_xywh_bbox_coder.py:163-260 (class agnostic), same fp32 op order
//   multiclass candidates nuhtc/models/bbox_head.py:12-60: (roi, class) pairs, score threshold, class-agnostic box expand
//   detection slots       bbox_nms.py:86-102 `dets[:max_num]` as max_per_img fixed slots per tile (no host round trip)
//   tile filter           tools/infer_wsi.py:510-521 margin / min_area filter -> tile id or -1 for the mask NMS
// Every arithmetic step is a separately rounded fp32 operation (__fmul_rn / __fadd_rn), exactly what the chain of torch
// kernels computes; expf is the same libdevice routine ATen's exp kernel calls.
#include "common.cuh"
#include "decode.cuh"

namespace {

__global__ void delta2bbox_kernel(const float *__restrict__ rois, int with_batch, const float *__restrict__ deltas, int64_t K,
                                  F4 means, F4 stds, float max_ratio, int clamp, float max_w, float max_h, float inv_scale_div,
                                  int divide, float *__restrict__ out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int rs = with_batch ? 5 : 4;
    const float *r = rois + k * rs + (with_batch ? 1 : 0);
    const float4 dl = *reinterpret_cast<const float4 *>(deltas + k * 4);
    const float x1 = r[0], y1 = r[1], x2 = r[2], y2 = r[3];
    float o4[4];
    decode_box(x1, y1, x2, y2, dl, means, stds, max_ratio, clamp, max_w, max_h, o4);
    float o0 = o4[0], o1 = o4[1], o2 = o4[2], o3 = o4[3];
    if (divide) {
        o0 = __fdiv_rn(o0, inv_scale_div);
        o1 = __fdiv_rn(o1, inv_scale_div);
        o2 = __fdiv_rn(o2, inv_scale_div);
        o3 = __fdiv_rn(o3, inv_scale_div);
    }
    float *o = out + k * rs;
    if (with_batch) {
        o[0] = rois[k * 5];
        ++o;
    }
    o[0] = o0;
    o[1] = o1;
    o[2] = o2;
    o[3] = o3;
}

__global__ void candidates_kernel(const float *__restrict__ boxes, int box_stride, const float *__restrict__ scores, int score_stride,
                                  const float *__restrict__ rois_tile, int tile_stride, int64_t K, int C, float score_thr,
                                  float *__restrict__ cand_boxes, float *__restrict__ cand_scores, int64_t *__restrict__ cand_labels,
                                  int32_t *__restrict__ cand_tile, int32_t *__restrict__ groups) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * C) return;
    const int64_t k = i / C;
    const int c = (int)(i - k * C);
    const float s = scores[k * score_stride + c];
    const int32_t t = (int32_t)rois_tile[k * tile_stride];
    const float *bp = boxes + k * box_stride;
    *reinterpret_cast<float4 *>(cand_boxes + i * 4) = make_float4(bp[0], bp[1], bp[2], bp[3]);
    cand_scores[i] = s;
    cand_labels[i] = c;
    cand_tile[i] = t;
    groups[i] = s > score_thr ? t : -1;
}

__global__ void slots_kernel(const int64_t *__restrict__ keep, const int64_t *__restrict__ gstart, const int64_t *__restrict__ gcount,
                             int B, int M, const float *__restrict__ cand_boxes, const float *__restrict__ cand_scores,
                             const int64_t *__restrict__ cand_labels, const int32_t *__restrict__ cand_tile, float scale,
                             float *__restrict__ det_boxes, float *__restrict__ det_scores, int64_t *__restrict__ det_labels,
                             int32_t *__restrict__ det_tile, uint8_t *__restrict__ det_valid, int64_t *__restrict__ det_cand,
                             float *__restrict__ mask_rois) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * M) return;
    const int b = i / M, r = i - b * M;
    const bool valid = (int64_t)r < gcount[b];
    float4 bx = make_float4(-4096.f, -4096.f, -4095.f, -4095.f); // far outside every frame: RoIAlign and paste see nothing
    float sc = 0.f;
    int64_t lab = 0, cand = 0;
    int32_t tile = -1;
    if (valid) {
        cand = keep[gstart[b] + r];
        bx = *reinterpret_cast<const float4 *>(cand_boxes + cand * 4);
        sc = cand_scores[cand];
        lab = cand_labels[cand];
        tile = cand_tile[cand];
    } else {
        lab = cand_labels[0];
    }
    *reinterpret_cast<float4 *>(det_boxes + (size_t)i * 4) = bx;
    det_scores[i] = sc;
    det_labels[i] = lab;
    det_tile[i] = tile;
    det_valid[i] = valid ? 1 : 0;
    det_cand[i] = cand;
    float *mr = mask_rois + (size_t)i * 5;
    mr[0] = (float)max(tile, 0);
    mr[1] = __fmul_rn(bx.x, scale);
    mr[2] = __fmul_rn(bx.y, scale);
    mr[3] = __fmul_rn(bx.z, scale);
    mr[4] = __fmul_rn(bx.w, scale);
}

__global__ void tile_filter_kernel(const float *__restrict__ det_boxes, const int32_t *__restrict__ area,
                                   const int32_t *__restrict__ det_tile, int64_t D, float margin, float wmax, float hmax,
                                   int min_area, int32_t *__restrict__ tile_ids) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D) return;
    const float4 b = *reinterpret_cast<const float4 *>(det_boxes + i * 4);
    const bool ok = b.x >= margin && b.y >= margin && b.z <= wmax && b.w <= hmax && area[i] >= min_area;
    tile_ids[i] = ok ? det_tile[i] : -1;
}

__global__ void keep_flags_kernel(const int32_t *__restrict__ keep, const int32_t *__restrict__ tile_start,
                                  const int32_t *__restrict__ tile_count, int T, int max_tile, uint8_t *__restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T * max_tile) return;
    const int t = i / max_tile, r = i - t * max_tile;
    if (r < tile_count[t]) flags[keep[tile_start[t] + r]] = 1;
}

} // namespace

NUHTC_API int nuhtc_keep_flags(const int32_t *keep, const int32_t *tile_start, const int32_t *tile_count, int num_tiles,
                               int max_tile_size, int64_t n, uint8_t *flags, void *stream) {
    NUHTC_CHECK_ARG(num_tiles >= 1 && max_tile_size >= 1 && n >= 0, "keep_flags: bad sizes");
    if (n == 0) return NUHTC_OK;
    NUHTC_CHECK_ARG(keep && tile_start && tile_count && flags, "keep_flags: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    NUHTC_CUDA(cudaMemsetAsync(flags, 0, (size_t)n, st));
    const int tot = num_tiles * max_tile_size;
    keep_flags_kernel<<<(tot + 255) / 256, 256, 0, st>>>(keep, tile_start, tile_count, num_tiles, max_tile_size, flags);
    NUHTC_LAUNCH_CHECK();
    return NUHTC_OK;
}

NUHTC_API int nuhtc_delta2bbox(const float *rois, int with_batch, const float *deltas, int64_t K, const float *means,
                               const float *stds, int max_h, int max_w, double wh_ratio_clip, float divide_by, float *out,
                               void *stream) {
    NUHTC_CHECK_ARG(K >= 0 && means && stds && wh_ratio_clip > 0.0, "delta2bbox: bad arguments");
    if (K == 0) return NUHTC_OK;
    NUHTC_CHECK_ARG(rois && deltas && out, "delta2bbox: null pointer");
    NUHTC_CHECK_ARG(((uintptr_t)deltas & 15) == 0, "delta2bbox: deltas must be 16-byte aligned");
    F4 m, s;
    for (int i = 0; i < 4; ++i) {
        m.v[i] = means[i];
        s.v[i] = stds[i];
    }
    const float max_ratio = (float)fabs(log(wh_ratio_clip));
    const int clamp = max_h > 0 && max_w > 0;
    delta2bbox_kernel<<<(unsigned)((K + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        rois, with_batch, deltas, K, m, s, max_ratio, clamp, (float)max_w, (float)max_h, divide_by, divide_by != 0.f && divide_by != 1.f, out);
    NUHTC_LAUNCH_CHECK();
    return NUHTC_OK;
}

NUHTC_API int nuhtc_multiclass_candidates(const float *boxes, int box_stride, const float *scores, int score_stride, const float *roi_tile,
                                          int tile_stride, int64_t K, int num_classes, float score_thr, float *cand_boxes,
                                          float *cand_scores, int64_t *cand_labels, int32_t *cand_tile, int32_t *groups,
                                          void *stream) {
    NUHTC_CHECK_ARG(K >= 0 && num_classes >= 1 && score_stride >= num_classes && tile_stride >= 1 && box_stride >= 4,
                    "multiclass_candidates: bad sizes");
    if (K == 0) return NUHTC_OK;
    NUHTC_CHECK_ARG(boxes && scores && roi_tile && cand_boxes && cand_scores && cand_labels && cand_tile && groups,
                    "multiclass_candidates: null pointer");
    const int64_t n = K * num_classes;
    candidates_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        boxes, box_stride, scores, score_stride, roi_tile, tile_stride, K, num_classes, score_thr, cand_boxes, cand_scores, cand_labels, cand_tile, groups);
    NUHTC_LAUNCH_CHECK();
    return NUHTC_OK;
}

NUHTC_API int nuhtc_detection_slots(const int64_t *keep, const int64_t *group_start, const int64_t *group_count, int num_tiles,
                                    int max_per_img, const float *cand_boxes, const float *cand_scores, const int64_t *cand_labels,
                                    const int32_t *cand_tile, float scale_factor, float *det_boxes, float *det_scores,
                                    int64_t *det_labels, int32_t *det_tile, uint8_t *det_valid, int64_t *det_cand, float *mask_rois,
                                    void *stream) {
    NUHTC_CHECK_ARG(num_tiles >= 1 && max_per_img >= 1, "detection_slots: bad sizes");
    NUHTC_CHECK_ARG(keep && group_start && group_count && cand_boxes && cand_scores && cand_labels && cand_tile && det_boxes &&
                        det_scores && det_labels && det_tile && det_valid && det_cand && mask_rois,
                    "detection_slots: null pointer");
    const int n = num_tiles * max_per_img;
    slots_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(keep, group_start, group_count, num_tiles, max_per_img, cand_boxes,
                                                                    cand_scores, cand_labels, cand_tile, scale_factor, det_boxes,
                                                                    det_scores, det_labels, det_tile, det_valid, det_cand, mask_rois);
    NUHTC_LAUNCH_CHECK();
    return NUHTC_OK;
}

NUHTC_API int nuhtc_tile_filter(const float *det_boxes, const int32_t *area, const int32_t *det_tile, int64_t D, int margin,
                                int img_h, int img_w, int min_area, int32_t *tile_ids, void *stream) {
    NUHTC_CHECK_ARG(D >= 0, "tile_filter: bad sizes");
    if (D == 0) return NUHTC_OK;
    NUHTC_CHECK_ARG(det_boxes && area && det_tile && tile_ids, "tile_filter: null pointer");
    tile_filter_kernel<<<(unsigned)((D + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        det_boxes, area, det_tile, D, (float)margin, (float)(img_w - margin), (float)(img_h - margin), min_area, tile_ids);
    NUHTC_LAUNCH_CHECK();
    return NUHTC_OK;
}
