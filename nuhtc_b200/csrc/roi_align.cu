// Multi-level RoIAlign forward for sm_100a.
//
// Replaces mmcv-full 1.7.2 `roi_align_forward` (avg pool) as driven per FPN level by
//   /root/reference/thirdparty/mmdetection/mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py:58-115
//   /root/reference/nuhtc/models/roi_extractors_cus.py:195-218,246
// with ONE launch that covers every level (per-RoI level routing, or level sum).
//
// Two kernels:
//  * roi_align_sep_kernel  -- the fast path.  RoIAlign's sample grid is a tensor product, so
//      out[ph][pw] = sum_y sum_x Wy[ph][y] * Wx[pw][x] * V[y][x]     (avg pool, bilinear)
//    with per-bin tap weights Wx/Wy that depend only on the RoI.  A CTA owns (RoI, channel chunk);
//    the tap tables are built once per RoI in shared memory ("staging of sampling coordinates"),
//    every thread then owns 4 contiguous channels (one 128-bit NHWC gather per tap) and one output
//    column pw, sweeps the window rows once, and keeps its P outputs in registers.  The [CC,P,P]
//    result is transposed through shared memory and leaves as coalesced 128-bit streaming stores
//    of the contiguous NCHW chunk.  Reads per RoI drop from 4*gh*gw*P*P to (rows * x-taps) per
//    column, which is what lets the kernel run at the HBM write rate instead of the L1 rate.
//  * roi_align_direct_kernel -- literal per-sample restatement (any layout / shape), same float op
//    order as the CPU reference; used as fallback and as an on-device cross-check.
#include "common.cuh"

struct RoiLevels {
    const float *data[NUHTC_MAX_LEVELS];
    int H[NUHTC_MAX_LEVELS];
    int W[NUHTC_MAX_LEVELS];
    float scale[NUHTC_MAX_LEVELS];
    int L;
};

// SingleRoIExtractor.map_roi_levels: floor(log2(sqrt(w*h)/finest + 1e-6)) clamped to [0, L-1].
// floor(log2 v) >= k  <=>  v >= 2^k, so the level is found by comparisons (no log2 rounding).
__device__ __forceinline__ int route_level(const float *roi, int L, float finest) {
    const float s = __fsqrt_rn(__fmul_rn(__fsub_rn(roi[3], roi[1]), __fsub_rn(roi[4], roi[2])));
    const float v = __fadd_rn(__fdiv_rn(s, finest), 1e-6f);
    int l = 0;
    float p2 = 2.0f;
    for (int k = 1; k < L; ++k) {
        if (v >= p2) l = k;
        p2 *= 2.0f;
    }
    return l;
}

struct RoiGeom {
    float start_w, start_h, bin_w, bin_h;
    int gw, gh;
    float count;
    int b;
};

__device__ __forceinline__ RoiGeom roi_geom(const float *roi, float scale, int PH, int PW, int sr, int aligned) {
    RoiGeom g;
    const float off = aligned ? 0.5f : 0.0f;
    g.start_w = __fsub_rn(__fmul_rn(roi[1], scale), off);
    g.start_h = __fsub_rn(__fmul_rn(roi[2], scale), off);
    const float ew = __fsub_rn(__fmul_rn(roi[3], scale), off);
    const float eh = __fsub_rn(__fmul_rn(roi[4], scale), off);
    float rw = __fsub_rn(ew, g.start_w), rh = __fsub_rn(eh, g.start_h);
    if (!aligned) {
        rw = fmaxf(rw, 1.0f);
        rh = fmaxf(rh, 1.0f);
    }
    g.bin_h = __fdiv_rn(rh, (float)PH);
    g.bin_w = __fdiv_rn(rw, (float)PW);
    g.gh = sr > 0 ? sr : (int)ceilf(__fdiv_rn(rh, (float)PH));
    g.gw = sr > 0 ? sr : (int)ceilf(__fdiv_rn(rw, (float)PW));
    int c = g.gh * g.gw;
    if (c < 1) c = 1;
    g.count = (float)c;
    g.b = (int)roi[0];
    return g;
}

// sample coordinate, same association as the CPU reference: (start + p*bin) + ((i+.5)*bin)/g
__device__ __forceinline__ float sample_coord(float start, float bin, int p, int i, int g) {
    return __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)),
                     __fdiv_rn(__fmul_rn((float)i + 0.5f, bin), (float)g));
}

// one axis of the bilinear sample: false if the sample is outside (-1, D)
__device__ __forceinline__ bool axis_sample(float c, int D, int &lo, int &hi, float &l, float &h) {
    if (c < -1.0f || c > (float)D) return false;
    if (c <= 0.f) c = 0.f;
    lo = (int)c;
    if (lo >= D - 1) {
        hi = lo = D - 1;
        c = (float)lo;
    } else {
        hi = lo + 1;
    }
    l = __fsub_rn(c, (float)lo);
    h = __fsub_rn(1.0f, l);
    return true;
}

// ---------------------------------------------------------------------------------------------
// literal kernel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float direct_one(const float *plane, long sy, long sx, int H, int W, const RoiGeom &g,
                                            int ph, int pw) {
    float acc = 0.f;
    for (int iy = 0; iy < g.gh; ++iy) {
        const float y = sample_coord(g.start_h, g.bin_h, ph, iy, g.gh);
        for (int ix = 0; ix < g.gw; ++ix) {
            const float x = sample_coord(g.start_w, g.bin_w, pw, ix, g.gw);
            int yl, yh, xl, xh;
            float ly, hy, lx, hx;
            // the reference rejects the sample if EITHER axis is out of range
            const bool oky = axis_sample(y, H, yl, yh, ly, hy);
            const bool okx = axis_sample(x, W, xl, xh, lx, hx);
            if (!(oky && okx)) continue;
            const float w1 = __fmul_rn(hy, hx), w2 = __fmul_rn(hy, lx), w3 = __fmul_rn(ly, hx), w4 = __fmul_rn(ly, lx);
            const float v1 = __ldg(plane + yl * sy + xl * sx), v2 = __ldg(plane + yl * sy + xh * sx);
            const float v3 = __ldg(plane + yh * sy + xl * sx), v4 = __ldg(plane + yh * sy + xh * sx);
            const float s = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, v1), __fmul_rn(w2, v2)), __fmul_rn(w3, v3)),
                                      __fmul_rn(w4, v4));
            acc = __fadd_rn(acc, s);
        }
    }
    return __fdiv_rn(acc, g.count);
}

__global__ void __launch_bounds__(256) roi_align_direct_kernel(RoiLevels lv, int C, int layout, const float *__restrict__ rois,
                                                               long total, int PH, int PW, int sr, int aligned, int mode,
                                                               float finest, float *__restrict__ out) {
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int pw = (int)(idx % PW);
        const int ph = (int)((idx / PW) % PH);
        const int c = (int)((idx / ((long)PW * PH)) % C);
        const long k = idx / ((long)PW * PH * C);
        const float *roi = rois + k * 5;
        float r = 0.f;
        int l0 = 0, l1 = lv.L;
        if (mode == NUHTC_ROI_ROUTE) {
            l0 = route_level(roi, lv.L, finest);
            l1 = l0 + 1;
        }
        for (int l = l0; l < l1; ++l) {
            const int H = lv.H[l], W = lv.W[l];
            const RoiGeom g = roi_geom(roi, lv.scale[l], PH, PW, sr, aligned);
            const float *plane;
            long sy, sx;
            if (layout == NUHTC_LAYOUT_NCHW) {
                plane = lv.data[l] + ((long)g.b * C + c) * H * W;
                sy = W;
                sx = 1;
            } else {
                plane = lv.data[l] + (long)g.b * H * W * C + c;
                sy = (long)W * C;
                sx = C;
            }
            const float f = direct_one(plane, sy, sx, H, W, g, ph, pw);
            r = (mode == NUHTC_ROI_ROUTE) ? f : __fadd_rn(r, f);
        }
        out[idx] = r;
    }
}

// ---------------------------------------------------------------------------------------------
// separable fast kernel (NHWC input)
// ---------------------------------------------------------------------------------------------
constexpr int kMaxTap = 8;   // taps per bin per axis held in the tables
constexpr int kMaxRows = 48; // window rows per RoI handled by the fast path

template <int P>
struct SepCfg {
    static constexpr int PP = P * P;
    // channel stride of the staging tile: == 1 (mod 8) makes the rotated column writes conflict-free
    static constexpr int S = (PP % 8 == 1) ? PP : (PP + ((9 - PP % 8) % 8));
};

// per-axis tap table for bin p: accumulated bilinear weights over the g samples of the bin
__device__ __forceinline__ bool build_axis_taps(float start, float bin, int g, int p, int D, float *w, int &first,
                                                int &n) {
#pragma unroll
    for (int j = 0; j < kMaxTap; ++j) w[j] = 0.f;
    int base = -1, last = -1;
    bool ok = true;
    for (int i = 0; i < g; ++i) {
        const float c = sample_coord(start, bin, p, i, g);
        int lo, hi;
        float l, h;
        if (!axis_sample(c, D, lo, hi, l, h)) continue;
        if (base < 0) base = lo;
        const int a = lo - base, b = hi - base;
        if (b >= kMaxTap) {
            ok = false;
            break;
        }
        w[a] += h;
        w[b] += l;
        last = b;
    }
    first = base < 0 ? 0 : base;
    n = last + 1;
    return ok;
}

// PHS: the P output rows are split over PHS thread groups (keeps the accumulators at P/PHS float4)
template <int P, int NQ, int PHS>
__global__ void __launch_bounds__(NQ *P *PHS) roi_align_sep_kernel(RoiLevels lv, int C, const float *__restrict__ rois, int sr,
                                                              int aligned, int mode, float finest,
                                                              float *__restrict__ out) {
    constexpr int CC = NQ * 4;
    constexpr int PP = SepCfg<P>::PP;
    constexpr int S = SepCfg<P>::S;
    constexpr int NT = NQ * P * PHS;
    constexpr int PB = P / PHS;
    static_assert(P % PHS == 0, "row split must divide P");
    extern __shared__ __align__(16) float smem[];
    float *s_tile = smem;                    // [CC][S]
    float *s_wx = s_tile + CC * S;           // [P][kMaxTap]
    float *s_wy = s_wx + P * kMaxTap;        // [P][kMaxTap]
    float *s_wyd = s_wy + P * kMaxTap;       // [P][kMaxRows] dense, already divided by count
    int *s_xs = (int *)(s_wyd + P * kMaxRows); // [P] first tap column
    int *s_nx = s_xs + P;
    int *s_ys = s_nx + P;
    int *s_ny = s_ys + P;
    int *s_ok = s_ny + P;            // [2P]
    int *s_rowmask = s_ok + 2 * P;   // [kMaxRows]

    const int k = blockIdx.x;
    const int c0 = blockIdx.y * CC;
    const int tid = threadIdx.x;
    const int q = tid % NQ;
    const int pw = (tid / NQ) % P;
    const int ph0 = (tid / (NQ * P)) * PB; // first output row owned by this thread
    const float *roi = rois + (size_t)k * 5;

    float4 acc[PB];
#pragma unroll
    for (int i = 0; i < PB; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    int l0 = 0, l1 = lv.L;
    if (mode == NUHTC_ROI_ROUTE) {
        l0 = route_level(roi, lv.L, finest);
        l1 = l0 + 1;
    }
    for (int l = l0; l < l1; ++l) {
        const int H = lv.H[l], W = lv.W[l];
        const RoiGeom g = roi_geom(roi, lv.scale[l], P, P, sr, aligned);
        if (l != l0) __syncthreads(); // previous level's tables are still being read
        // ---- phase A: tap tables (x bins on warp 0, y bins on warp 1)
        if (tid < P) {
            int f, n;
            const bool ok = build_axis_taps(g.start_w, g.bin_w, g.gw, tid, W, s_wx + tid * kMaxTap, f, n);
            s_xs[tid] = f;
            s_nx[tid] = n;
            s_ok[tid] = ok;
        } else if (tid >= 32 && tid < 32 + P) {
            const int p = tid - 32;
            int f, n;
            const bool ok = build_axis_taps(g.start_h, g.bin_h, g.gh, p, H, s_wy + p * kMaxTap, f, n);
            s_ys[p] = f;
            s_ny[p] = n;
            s_ok[P + p] = ok;
        }
        __syncthreads();
        bool fast = true;
        int ymin = 1 << 30, ymax = -1;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            fast = fast && s_ok[p] && s_ok[P + p];
            const int n = s_ny[p];
            if (n > 0) {
                ymin = min(ymin, s_ys[p]);
                ymax = max(ymax, s_ys[p] + n);
            }
        }
        const int nrows = ymax - ymin; // <= 0 when no valid row
        fast = fast && nrows <= kMaxRows;
        if (fast) {
            if (nrows > 0) {
                // ---- phase B: dense row weights and the row -> bins mask
                for (int i = tid; i < P * kMaxRows; i += NT) {
                    const int p = i / kMaxRows, r = i - p * kMaxRows;
                    const int j = r + ymin - s_ys[p];
                    float w = 0.f;
                    if (r < nrows && j >= 0 && j < s_ny[p]) w = __fdiv_rn(s_wy[p * kMaxTap + j], g.count);
                    s_wyd[i] = w;
                }
                if (tid < kMaxRows) {
                    int m = 0;
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const int j = tid + ymin - s_ys[p];
                        if (tid < nrows && j >= 0 && j < s_ny[p]) m |= 1 << p;
                    }
                    s_rowmask[tid] = m;
                }
                __syncthreads();
                // ---- phase C: one sweep over the window rows
                float wx[kMaxTap];
                const int nx = s_nx[pw];
#pragma unroll
                for (int j = 0; j < kMaxTap; ++j) wx[j] = s_wx[pw * kMaxTap + j];
                // rows touched by this thread's own output rows
                int r0 = nrows, r1 = 0;
#pragma unroll
                for (int i = 0; i < PB; ++i) {
                    const int n = s_ny[ph0 + i];
                    if (n > 0) {
                        r0 = min(r0, s_ys[ph0 + i] - ymin);
                        r1 = max(r1, s_ys[ph0 + i] + n - ymin);
                    }
                }
                const float *rowp =
                    lv.data[l] + (((size_t)g.b * H + ymin + r0) * W + s_xs[pw]) * (size_t)C + c0 + 4 * q;
                const size_t rstride = (size_t)W * C;
#pragma unroll 2
                for (int r = r0; r < r1; ++r, rowp += rstride) {
                    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < kMaxTap; ++j) {
                        if (j < nx) {
                            const float4 v = ldg_f4(rowp + (size_t)j * C);
                            t.x = fmaf(wx[j], v.x, t.x);
                            t.y = fmaf(wx[j], v.y, t.y);
                            t.z = fmaf(wx[j], v.z, t.z);
                            t.w = fmaf(wx[j], v.w, t.w);
                        }
                    }
                    const int m = s_rowmask[r] >> ph0;
#pragma unroll
                    for (int i = 0; i < PB; ++i) {
                        if ((m >> i) & 1) {
                            const float w = s_wyd[(ph0 + i) * kMaxRows + r];
                            acc[i].x = fmaf(w, t.x, acc[i].x);
                            acc[i].y = fmaf(w, t.y, acc[i].y);
                            acc[i].z = fmaf(w, t.z, acc[i].z);
                            acc[i].w = fmaf(w, t.w, acc[i].w);
                        }
                    }
                }
            }
        } else {
            // ---- oversized RoI: literal per-sample accumulation, same thread mapping
            const float *img = lv.data[l] + (size_t)g.b * H * W * C + c0 + 4 * q;
#pragma unroll 1
            for (int pi = 0; pi < PB; ++pi) {
                const int ph = ph0 + pi;
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int iy = 0; iy < g.gh; ++iy) {
                    int yl, yh;
                    float ly, hy;
                    const bool oky = axis_sample(sample_coord(g.start_h, g.bin_h, ph, iy, g.gh), H, yl, yh, ly, hy);
                    for (int ix = 0; ix < g.gw; ++ix) {
                        int xl, xh;
                        float lx, hx;
                        const bool okx = axis_sample(sample_coord(g.start_w, g.bin_w, pw, ix, g.gw), W, xl, xh, lx, hx);
                        if (!(oky && okx)) continue;
                        const float4 v1 = ldg_f4(img + ((size_t)yl * W + xl) * C), v2 = ldg_f4(img + ((size_t)yl * W + xh) * C);
                        const float4 v3 = ldg_f4(img + ((size_t)yh * W + xl) * C), v4 = ldg_f4(img + ((size_t)yh * W + xh) * C);
                        const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
                        a.x += w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x;
                        a.y += w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y;
                        a.z += w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z;
                        a.w += w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w;
                    }
                }
                // acc is indexed statically below: fold this bin into the matching register
#pragma unroll
                for (int pp = 0; pp < PB; ++pp) {
                    if (pp == pi) {
                        acc[pp].x += __fdiv_rn(a.x, g.count);
                        acc[pp].y += __fdiv_rn(a.y, g.count);
                        acc[pp].z += __fdiv_rn(a.z, g.count);
                        acc[pp].w += __fdiv_rn(a.w, g.count);
                    }
                }
            }
        }
    }

    // ---- transpose [4 ch of this thread][ph][pw] into the [CC][S] staging tile.
    // Lanes of a warp hold consecutive channel quads (stride 4*S words == 4 mod 32), so the four
    // channels are written in a lane-rotated order that spreads the 32 lanes over all 32 banks.
    const int lane = tid & 31;
    const int rot = (lane >> 3) - (lane / NQ);
#pragma unroll
    for (int i = 0; i < PB; ++i) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = (jj + rot) & 3;
            const float v = j == 0 ? acc[i].x : (j == 1 ? acc[i].y : (j == 2 ? acc[i].z : acc[i].w));
            s_tile[(4 * q + j) * S + (ph0 + i) * P + pw] = v;
        }
    }
    __syncthreads();
    float *outp = out + ((size_t)k * C + c0) * PP;
    constexpr int N4 = CC * PP / 4;
    for (int i = tid; i < N4; i += NT) {
        float4 v;
        if (S == PP) {
            v = *reinterpret_cast<const float4 *>(s_tile + 4 * i);
        } else {
            int c = (4 * i) / PP, e = 4 * i - c * PP;
            float t[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                t[u] = s_tile[c * S + e];
                if (++e == PP) {
                    e = 0;
                    ++c;
                }
            }
            v = make_float4(t[0], t[1], t[2], t[3]);
        }
        st_stream_f4(outp + 4 * i, v);
    }
}

template <int P, int NQ>
static size_t sep_smem_bytes() {
    return sizeof(float) * (NQ * 4 * SepCfg<P>::S + 2 * P * kMaxTap + P * kMaxRows) + sizeof(int) * (6 * P + kMaxRows);
}

template <int P, int NQ, int PHS>
static int launch_sep(const RoiLevels &lv, int C, const float *rois, int K, int sr, int aligned, int mode, float finest,
                      float *out, cudaStream_t st) {
    static bool attr_done = false;
    const size_t smem = sep_smem_bytes<P, NQ>();
    if (!attr_done) {
        NUHTC_CUDA(cudaFuncSetAttribute(roi_align_sep_kernel<P, NQ, PHS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    dim3 grid(K, C / (NQ * 4));
    roi_align_sep_kernel<P, NQ, PHS><<<grid, NQ * P * PHS, smem, st>>>(lv, C, rois, sr, aligned, mode, finest, out);
    NUHTC_LAUNCH_CHECK();
    return NUHTC_OK;
}

NUHTC_API int nuhtc_roi_align_fwd(const float *const *feats, const int *H, const int *W, const float *scale, int L, int B,
                                  int C, int layout, const float *rois, int K, int PH, int PW, int sampling_ratio,
                                  int aligned, int mode, float finest_scale, int impl, float *out, void *stream) {
    NUHTC_CHECK_ARG(L >= 1 && L <= NUHTC_MAX_LEVELS, "roi_align: L=%d out of range", L);
    NUHTC_CHECK_ARG(B >= 0 && C >= 1 && PH >= 1 && PW >= 1 && K >= 0, "roi_align: bad sizes B=%d C=%d PH=%d PW=%d K=%d", B, C,
                    PH, PW, K);
    NUHTC_CHECK_ARG(layout == NUHTC_LAYOUT_NCHW || layout == NUHTC_LAYOUT_NHWC, "roi_align: bad layout %d", layout);
    NUHTC_CHECK_ARG(mode == NUHTC_ROI_ROUTE || mode == NUHTC_ROI_SUM, "roi_align: bad mode %d", mode);
    NUHTC_CHECK_ARG(finest_scale > 0.f || L == 1 || mode == NUHTC_ROI_SUM, "roi_align: finest_scale must be > 0");
    if (K == 0) return NUHTC_OK;
    NUHTC_CHECK_ARG(feats && H && W && scale && rois && out, "roi_align: null pointer");
    RoiLevels lv;
    lv.L = L;
    for (int l = 0; l < L; ++l) {
        NUHTC_CHECK_ARG(feats[l] != nullptr && H[l] >= 1 && W[l] >= 1, "roi_align: bad level %d", l);
        lv.data[l] = feats[l];
        lv.H[l] = H[l];
        lv.W[l] = W[l];
        lv.scale[l] = scale[l];
    }
    cudaStream_t st = (cudaStream_t)stream;
    bool aligned16 = true;
    for (int l = 0; l < L; ++l) aligned16 = aligned16 && (((uintptr_t)feats[l]) % 16 == 0);
    aligned16 = aligned16 && (((uintptr_t)out) % 16 == 0);
    const bool fast_ok = impl == NUHTC_IMPL_AUTO && layout == NUHTC_LAYOUT_NHWC && PH == PW && (PH == 7 || PH == 14) &&
                         C % 64 == 0 && aligned16;
    if (fast_ok) {
        if (PH == 7) {
            if (C % 256 == 0) return launch_sep<7, 64, 1>(lv, C, rois, K, sampling_ratio, aligned, mode, finest_scale, out, st);
            if (C % 128 == 0) return launch_sep<7, 32, 1>(lv, C, rois, K, sampling_ratio, aligned, mode, finest_scale, out, st);
            return launch_sep<7, 16, 1>(lv, C, rois, K, sampling_ratio, aligned, mode, finest_scale, out, st);
        }
        return launch_sep<14, 16, 2>(lv, C, rois, K, sampling_ratio, aligned, mode, finest_scale, out, st);
    }
    const long total = (long)K * C * PH * PW;
    const int threads = 256;
    long blocks = (total + threads - 1) / threads;
    const long cap = (long)nuhtc_sm_count() * 32;
    if (blocks > cap) blocks = cap;
    roi_align_direct_kernel<<<(unsigned)blocks, threads, 0, st>>>(lv, C, layout, rois, total, PH, PW, sampling_ratio, aligned,
                                                                  mode, finest_scale, out);
    NUHTC_LAUNCH_CHECK();
    return NUHTC_OK;
}

// ---------------------------------------------------------------------------------------------
// NCHW -> NHWC (once per level per batch; HBM-bound transpose through a padded smem tile)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float *__restrict__ in, float *__restrict__ out, int C,
                                                           int HW) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y; // (32, 8)
    const float *src = in + (size_t)b * C * HW;
    float *dst = out + (size_t)b * C * HW;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int c = c0 + ty + i, hw = hw0 + tx;
        if (c < C && hw < HW) tile[ty + i][tx] = __ldg(src + (size_t)c * HW + hw);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int hw = hw0 + ty + i, c = c0 + tx;
        if (c < C && hw < HW) dst[(size_t)hw * C + c] = tile[tx][ty + i];
    }
}

NUHTC_API int nuhtc_nchw_to_nhwc(const float *in, float *out, int B, int C, int H, int W, void *stream) {
    NUHTC_CHECK_ARG(B >= 0 && C >= 1 && H >= 1 && W >= 1, "nchw_to_nhwc: bad sizes");
    if (B == 0) return NUHTC_OK;
    NUHTC_CHECK_ARG(in && out, "nchw_to_nhwc: null pointer");
    const int HW = H * W;
    dim3 grid((HW + 31) / 32, (C + 31) / 32, B);
    NUHTC_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "nchw_to_nhwc: C or B too large");
    nchw_to_nhwc_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(in, out, C, HW);
    NUHTC_LAUNCH_CHECK();
    return NUHTC_OK;
}
