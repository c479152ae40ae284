// Mask -> contour for sm_100a: the first contour of cv2.findContours(mask, RETR_TREE, CHAIN_APPROX_SIMPLE), i.e.
// `mask2inst` of /root/reference/tools/infer_wsi.py:51-54, for a batch of bit-row masks, without the per-nucleus
// device->host copy + OpenCV call of the reference's tile loop (infer_wsi.py:528-533).
//
// OpenCV implements Suzuki & Abe border following; contour [0] of the RETR_TREE list is the LAST outer border found in
// raster order whose parent is the frame.  The parent of a new border depends only on (hole?, outer-border-of-a-top-level
// component?) of the last border met on the row, so instead of border ids the pixel labels carry that 2-bit descriptor:
//   0 background, 1 unvisited, +-(2 + hole + 2*top) visited (negative = its right neighbour was examined as background).
// One warp per mask: all lanes find the tight window of the mask, expand it into an int8 label plane with a 1-pixel
// apron in shared memory and precompute the left/right-edge pixels of every row (the only places a border can start) as
// bit words; lane 0 then walks those events in raster order and follows the borders exactly like the reference
// algorithm, emitting the points of every top-level outer border (the last one wins).  Nucleus masks have one ~100-pixel
// border, so the serial part is a few thousand instructions per mask and 32 masks are resident per SM.
#include "common.cuh"

namespace {

constexpr int kSmallWin = 64; // window side handled by the small pass (one event word per row)

__device__ __forceinline__ uint64_t win_word(const uint64_t *__restrict__ mb, int wpm, int y, int x0, int ww, int k, int evw) {
    // 64 window columns [64k, 64k+64) of mask row y, window column 0 = mask column x0; 0 outside the window
    if (k < 0 || k >= evw) return 0ull;
    const int xs = x0 + 64 * k;
    const int w0 = xs >> 6, sh = xs & 63;
    const uint64_t *row = mb + (size_t)y * wpm;
    uint64_t v = __ldg(row + w0) >> sh;
    if (sh && w0 + 1 < wpm) v |= __ldg(row + w0 + 1) << (64 - sh);
    const int rem = ww - 64 * k;
    if (rem < 64) v &= (1ull << rem) - 1ull;
    return v;
}

// pass: 0 = small windows, larger ones are an error (no large pass possible); 1 = small windows, larger ones skipped;
//       2 = only the larger ones
__global__ void __launch_bounds__(32) contour_kernel(const uint64_t *__restrict__ bits, const int32_t *__restrict__ bbox,
                                                     const uint8_t *__restrict__ select, int64_t N,
                                                     int h, int wpm, int pass,
                                                     int max_pts, int32_t *__restrict__ out_xy, int32_t *__restrict__ out_count,
                                                     int32_t *__restrict__ status) {
    extern __shared__ uint64_t smem64[];
    __shared__ int s_off[8];
    const int lane = threadIdx.x;
    for (int64_t m = blockIdx.x; m < N; m += gridDim.x) {
        const uint64_t *mb = bits + (size_t)m * h * wpm;
        if (select && !select[m]) { // not asked for (e.g. suppressed by the mask NMS): no contour
            if (pass != 2 && lane == 0) out_count[m] = 0;
            continue;
        }
        // ---- tight window of the mask
        int ymin = h, ymax = -1, xmin = wpm * 64, xmax = -1;
        if (bbox) { // tight box from the paste / pack kernels: x0, y0, x1, y1 (exclusive), all 0 for an empty mask
            const int4 bb = __ldg(reinterpret_cast<const int4 *>(bbox) + m);
            xmin = bb.x;
            ymin = bb.y;
            xmax = bb.z - 1;
            ymax = bb.w > bb.y ? bb.w - 1 : -1;
        }
        for (int i = lane; !bbox && i < h * wpm; i += 32) {
            const uint64_t v = __ldg(mb + i);
            if (v) {
                const int r = i / wpm, k = i - r * wpm;
                ymin = min(ymin, r);
                ymax = max(ymax, r);
                xmin = min(xmin, 64 * k + __ffsll((long long)v) - 1);
                xmax = max(xmax, 64 * k + 63 - __clzll((long long)v));
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
            ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
            xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
            xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        }
        if (ymax < 0) { // empty mask: findContours returns no contour
            if (pass != 2 && lane == 0) out_count[m] = 0;
            continue;
        }
        const int ww = xmax - xmin + 1, wh = ymax - ymin + 1;
        const bool big = ww > kSmallWin || wh > kSmallWin;
        if (big != (pass == 2)) {
            if (pass == 0 && lane == 0) {
                out_count[m] = 0;
                atomicExch(status, 2);
            }
            continue;
        }
        const int evw = (ww + 63) >> 6, S = ww + 2;
        uint64_t *ev = smem64;
        signed char *lab = reinterpret_cast<signed char *>(smem64 + (size_t)wh * evw);
        __syncwarp();
        if (lane < 8) {
            const int dx = (lane == 0 || lane == 1 || lane == 7) ? 1 : ((lane >= 3 && lane <= 5) ? -1 : 0);
            const int dy = (lane >= 1 && lane <= 3) ? -1 : ((lane >= 5) ? 1 : 0);
            s_off[lane] = dy * S + dx;
        }
        // apron rows / columns
        for (int i = lane; i < S; i += 32) {
            lab[i] = 0;
            lab[(wh + 1) * S + i] = 0;
        }
        for (int r = lane; r < wh; r += 32) {
            lab[(r + 1) * S] = 0;
            lab[(r + 1) * S + ww + 1] = 0;
        }
        // label plane + start events, one 64-column word per step
        for (int i = lane; i < wh * evw; i += 32) {
            const int r = i / evw, k = i - r * evw;
            const uint64_t b = win_word(mb, wpm, ymin + r, xmin, ww, k, evw);
            const uint64_t pv = win_word(mb, wpm, ymin + r, xmin, ww, k - 1, evw);
            const uint64_t nx = win_word(mb, wpm, ymin + r, xmin, ww, k + 1, evw);
            ev[i] = (b & ~((b << 1) | (pv >> 63))) | (b & ~((b >> 1) | (nx << 63)));
            signed char *dst = lab + (r + 1) * S + 1 + 64 * k;
            const int nb = min(64, ww - 64 * k);
            for (int c = 0; c < nb; ++c) dst[c] = (signed char)((b >> c) & 1ull);
        }
        __syncwarp();
        if (lane == 0) {
            int cnt = 0;
            int32_t *oxy = out_xy + (size_t)m * max_pts * 2;
            for (int r = 0; r < wh; ++r) {
                for (int k = 0; k < evw; ++k) {
                    uint64_t e = ev[r * evw + k];
                    while (e) {
                        const int c = 64 * k + __ffsll((long long)e) - 1;
                        e &= e - 1;
                        signed char *p0 = lab + (r + 1) * S + c + 1;
                        const int v = *p0;
                        int hole;
                        if (v == 1 && p0[-1] == 0) hole = 0;
                        else if (v >= 1 && p0[1] == 0) hole = 1;
                        else continue;
                        // descriptor of the last border met on this row: bit0 hole, bit1 top-level outer, bit2 frame
                        int ld = 5;
                        if (hole && v > 1) ld = v - 2;
                        else {
                            const signed char *rs = lab + (r + 1) * S;
                            for (const signed char *q = p0 - 1; q > rs; --q) {
                                const int t = *q;
                                if (t != 0 && t != 1) {
                                    ld = abs(t) - 2;
                                    break;
                                }
                            }
                        }
                        int top = 0;
                        if (!hole) top = (ld & 1) ? ((ld >> 2) & 1) : ((ld >> 1) & 1);
                        const signed char code = (signed char)(2 + hole + 2 * top);
                        const bool emit = top != 0;
                        if (emit) cnt = 0;
                        int x = xmin + c, y = ymin + r;
                        // ---- follow the border (Suzuki & Abe step 3)
                        int s_end = hole ? 0 : 4, s = s_end;
                        signed char *i1;
                        do {
                            s = (s - 1) & 7;
                            i1 = p0 + s_off[s];
                        } while (*i1 == 0 && s != s_end);
                        if (s == s_end) {
                            *p0 = (signed char)-code;
                            if (emit) {
                                if (cnt < max_pts) {
                                    oxy[0] = x;
                                    oxy[1] = y;
                                }
                                cnt = 1;
                            }
                            continue;
                        }
                        signed char *i3 = p0, *i4;
                        int prev_s = s ^ 4;
                        for (;;) {
                            for (;;) {
                                ++s;
                                i4 = i3 + s_off[s & 7];
                                if (*i4 != 0) break;
                            }
                            if (s > 8) *i3 = (signed char)-code; // direction 0 (right) was examined as background
                            else if (*i3 == 1) *i3 = code;
                            s &= 7;
                            if (emit && s != prev_s) {
                                if (cnt < max_pts) {
                                    oxy[2 * cnt] = x;
                                    oxy[2 * cnt + 1] = y;
                                }
                                ++cnt;
                            }
                            prev_s = s;
                            x += (s == 0 || s == 1 || s == 7) ? 1 : ((s >= 3 && s <= 5) ? -1 : 0);
                            y += (s >= 1 && s <= 3) ? -1 : ((s >= 5) ? 1 : 0);
                            if (i4 == p0 && i3 == i1) break;
                            i3 = i4;
                            s = (s + 4) & 7;
                        }
                    }
                }
            }
            out_count[m] = cnt;
            if (cnt > max_pts) atomicExch(status, 1);
        }
        __syncwarp();
    }
}

// ring vertices for the merge: contour points + the repeated first point, shifted by the tile origin, as fp64
__global__ void contour_rings_kernel(const int32_t *__restrict__ xy, const int32_t *__restrict__ count,
                                     const int64_t *__restrict__ voff, const int32_t *__restrict__ origin, int max_pts,
                                     double *__restrict__ out) {
    const int64_t m = blockIdx.x;
    const int64_t v0 = voff[m];
    const int n = (int)(voff[m + 1] - v0);
    if (n <= 0) return;
    const int c = min(count[m], max_pts);
    const double ox = origin ? (double)origin[2 * m] : 0.0, oy = origin ? (double)origin[2 * m + 1] : 0.0;
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
        const int src = p < c ? p : 0;
        const int32_t *s = xy + ((size_t)m * max_pts + src) * 2;
        out[(v0 + p) * 2] = (double)s[0] + ox;
        out[(v0 + p) * 2 + 1] = (double)s[1] + oy;
    }
}

} // namespace

NUHTC_API int nuhtc_mask_contours(const uint64_t *bits, const int32_t *bbox, const uint8_t *select, int64_t n, int h, int w, int max_pts, int32_t *out_xy,
                                  int32_t *out_count, int32_t *status, void *stream) {
    NUHTC_CHECK_ARG(n >= 0 && h >= 1 && w >= 1 && max_pts >= 1 && status != nullptr, "mask_contours: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    NUHTC_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    if (n == 0) return NUHTC_OK;
    NUHTC_CHECK_ARG(bits && out_xy && out_count, "mask_contours: null pointer");
    const int wpm = (w + 63) / 64;
    const size_t small_smem = align_up((size_t)kSmallWin * 8 + (size_t)(kSmallWin + 2) * (kSmallWin + 2), 16);
    const size_t large_smem = align_up((size_t)h * wpm * 8 + (size_t)(h + 2) * (w + 2), 16);
    const bool need_large = h > kSmallWin || w > kSmallWin;
    static int max_optin = -1;
    if (max_optin < 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    }
    const bool large_ok = need_large && large_smem <= (size_t)max_optin;
    const unsigned grid = (unsigned)(n < (int64_t)nuhtc_sm_count() * 32 ? n : (int64_t)nuhtc_sm_count() * 32);
    NUHTC_CHECK_ARG(((uintptr_t)bbox & 15) == 0, "mask_contours: bbox must be 16-byte aligned");
    contour_kernel<<<grid, 32, small_smem, st>>>(bits, bbox, select, n, h, wpm, need_large ? (large_ok ? 1 : 0) : 1, max_pts, out_xy, out_count, status);
    NUHTC_LAUNCH_CHECK();
    if (large_ok) {
        NUHTC_CUDA(cudaFuncSetAttribute(contour_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)large_smem));
        const unsigned g2 = (unsigned)(n < (int64_t)nuhtc_sm_count() * 2 ? n : (int64_t)nuhtc_sm_count() * 2);
        contour_kernel<<<g2, 32, large_smem, st>>>(bits, bbox, select, n, h, wpm, 2, max_pts, out_xy, out_count, status);
        NUHTC_LAUNCH_CHECK();
    }
    return NUHTC_OK;
}

NUHTC_API int nuhtc_contour_rings(const int32_t *xy, const int32_t *count, const int64_t *voff, const int32_t *origin,
                                  int64_t n, int max_pts, double *out, void *stream) {
    NUHTC_CHECK_ARG(n >= 0 && max_pts >= 1, "contour_rings: bad sizes");
    if (n == 0) return NUHTC_OK;
    NUHTC_CHECK_ARG(xy && count && voff && out, "contour_rings: null pointer");
    contour_rings_kernel<<<(unsigned)n, 64, 0, (cudaStream_t)stream>>>(xy, count, voff, origin, max_pts, out);
    NUHTC_LAUNCH_CHECK();
    return NUHTC_OK;
}
