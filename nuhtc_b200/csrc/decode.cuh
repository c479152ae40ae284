// delta2bbox for one box (mmdet/core/bbox/coder/delta_xywh_bbox_coder.py:163-260, class agnostic).  Every arithmetic step is
// a separately rounded fp32 operation, exactly what the chain of torch kernels computes; expf is the libdevice routine
// ATen's exp kernel calls.  Shared by delta2bbox_kernel (det_glue.cu) and the RPN top-k kernel (rpn_topk.cu).
#pragma once
#include <cuda_runtime.h>

struct F4 {
    float v[4];
};

__device__ __forceinline__ void decode_box(float x1, float y1, float x2, float y2, float4 dl, const F4 &means, const F4 &stds,
                                           float max_ratio, int clamp, float max_w, float max_h, float (&o)[4]) {
    const float dx = __fadd_rn(__fmul_rn(dl.x, stds.v[0]), means.v[0]);
    const float dy = __fadd_rn(__fmul_rn(dl.y, stds.v[1]), means.v[1]);
    float dw = __fadd_rn(__fmul_rn(dl.z, stds.v[2]), means.v[2]);
    float dh = __fadd_rn(__fmul_rn(dl.w, stds.v[3]), means.v[3]);
    const float px = __fmul_rn(__fadd_rn(x1, x2), 0.5f), py = __fmul_rn(__fadd_rn(y1, y2), 0.5f);
    const float pw = __fsub_rn(x2, x1), ph = __fsub_rn(y2, y1);
    const float sx = __fmul_rn(pw, dx), sy = __fmul_rn(ph, dy);
    // torch.clamp propagates NaN; fminf/fmaxf would not
    dw = dw != dw ? dw : fminf(fmaxf(dw, -max_ratio), max_ratio);
    dh = dh != dh ? dh : fminf(fmaxf(dh, -max_ratio), max_ratio);
    const float gx = __fadd_rn(px, sx), gy = __fadd_rn(py, sy);
    const float gw = __fmul_rn(pw, expf(dw)), gh = __fmul_rn(ph, expf(dh));
    const float hw = __fmul_rn(gw, 0.5f), hh = __fmul_rn(gh, 0.5f);
    o[0] = __fsub_rn(gx, hw);
    o[1] = __fsub_rn(gy, hh);
    o[2] = __fadd_rn(gx, hw);
    o[3] = __fadd_rn(gy, hh);
    if (clamp) {
        o[0] = o[0] != o[0] ? o[0] : fminf(fmaxf(o[0], 0.f), max_w);
        o[2] = o[2] != o[2] ? o[2] : fminf(fmaxf(o[2], 0.f), max_w);
        o[1] = o[1] != o[1] ? o[1] : fminf(fmaxf(o[1], 0.f), max_h);
        o[3] = o[3] != o[3] ? o[3] : fminf(fmaxf(o[3], 0.f), max_h);
    }
}
