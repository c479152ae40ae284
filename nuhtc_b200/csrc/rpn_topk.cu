// RPN proposal pre-selection for sm_100a: per (image, level) the nms_pre best-scoring anchors in score order, decoded.
// Replaces the per-level body of RPNHead._get_bboxes_single (mmdet/models/dense_heads/rpn_head.py:103-165: permute,
// sigmoid, `scores.sort(descending=True)`, `[:nms_pre]`, gathers) and the decode + min-size test of _bbox_post_process
// (:167-236) for a whole batch in ONE launch; what leaves is exactly what the grouped NMS reads.
//
// One CTA per (level, image).  The n = H*W*A scores of the level sit in shared memory as order-preserving 32-bit keys
// (n <= 49152 at a 512 px frame: 192 KB); a 4-pass radix select finds the k-th largest key, the survivors are compacted
// (ties on the threshold key by ascending anchor index, i.e. what a stable sort keeps) and a bitonic sort of <= 2048
// (key, ~index) pairs orders them; then each thread decodes its anchors.  Levels with n <= nms_pre keep their natural order
// (the reference does not sort them).  No global scratch, no device-wide sort, no host round trip.
#include "common.cuh"
#include "decode.cuh"

namespace {

constexpr int kRpnThreads = 1024;
constexpr int kRpnMaxK = 2048;
constexpr int kRpnSubHist = 8;

struct RpnLevel {
    const float *cls;      // [B, A, H, W] logits
    const float *reg;      // [B, 4A, H, W]
    const float *anchors;  // [H*W*A, 4] in (h, w, a) order
    int H, W, n, k, out_off;
};
struct RpnArgs {
    RpnLevel lv[NUHTC_MAX_LEVELS];
    int L, B, A, per_image, apply_sigmoid, clamp;
    float max_w, max_h, max_ratio, min_size;
    float *boxes;      // [B, per_image, 4]
    float *scores;     // [B, per_image]
    int64_t *labels;   // [B, per_image] level index
    int32_t *groups;   // [B, per_image] image index, -1 when the box fails the min-size test
};

__device__ __forceinline__ uint32_t key_of(float s) {
    const uint32_t u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float score_of(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
// ATen's sigmoid kernel: one / (one + std::exp(-a)) in fp32, IEEE division
__device__ __forceinline__ float sigmoid_aten(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

__global__ void __launch_bounds__(kRpnThreads, 1) rpn_topk_decode_kernel(const __grid_constant__ RpnArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const RpnLevel lv = a.lv[blockIdx.x];
    const int b = blockIdx.y, tid = threadIdx.x, l = blockIdx.x;
    const int n = lv.n, k = lv.k, A = a.A, HW = lv.H * lv.W;
    unsigned long long *s_sel = reinterpret_cast<unsigned long long *>(smem_raw);        // [kRpnMaxK]
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(s_sel + kRpnMaxK);                   // [kRpnSubHist][256]
    uint32_t *s_scan = s_hist + kRpnSubHist * 256;                                       // [32] warp totals + scalars
    uint32_t *s_keys = s_scan + 64;                                                      // [n]
    const float *cls = lv.cls + (size_t)b * A * HW;
    const bool select = k < n;

    if (select) {
        // ---- keys into shared memory.  Candidate i = (hw, a) reads cls[a][hw]: consecutive threads take consecutive hw of one a
        for (int j = tid; j < n; j += kRpnThreads) {
            const int aa = j / HW, hw = j - aa * HW;
            float s = cls[j];
            if (a.apply_sigmoid) s = sigmoid_aten(s);
            s_keys[hw * A + aa] = key_of(s);
        }
        __syncthreads();
        // ---- radix select of the k-th largest key, 8 bits per pass from the top
        uint32_t prefix = 0, pmask = 0;
        int need = k;
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int j = tid; j < kRpnSubHist * 256; j += kRpnThreads) s_hist[j] = 0;
            __syncthreads();
            uint32_t *h = s_hist + ((tid >> 5) % kRpnSubHist) * 256;
            for (int j = tid; j < n; j += kRpnThreads) {
                const uint32_t key = s_keys[j];
                if ((key & pmask) == prefix) atomicAdd(h + ((key >> shift) & 255u), 1u);
            }
            __syncthreads();
            if (tid < 256) {
                uint32_t c = 0;
#pragma unroll
                for (int s = 0; s < kRpnSubHist; ++s) c += s_hist[s * 256 + tid];
                s_hist[tid] = c;
            }
            __syncthreads();
            if (tid == 0) {
                int rem = need, bin = 255;
                for (; bin > 0; --bin) {
                    const int c = (int)s_hist[bin];
                    if (c >= rem) break;
                    rem -= c;
                }
                s_scan[40] = (uint32_t)bin;
                s_scan[41] = (uint32_t)rem;
            }
            __syncthreads();
            prefix |= s_scan[40] << shift;
            pmask |= 255u << shift;
            need = (int)s_scan[41];
            __syncthreads();
        }
        const uint32_t T = prefix;   // the k-th largest key; `need` of the candidates equal to it are taken, lowest index first
        // ---- compaction.  key > T: any slot (sorted afterwards).  key == T: ordered by index through a block scan over
        // contiguous index chunks.
        if (tid == 0) s_scan[42] = 0;
        for (int j = tid; j < kRpnMaxK; j += kRpnThreads) s_sel[j] = 0ull;
        __syncthreads();
        const int chunk = (n + kRpnThreads - 1) / kRpnThreads;
        const int j0 = min(tid * chunk, n), j1 = min(j0 + chunk, n);
        int eq = 0;
        for (int j = j0; j < j1; ++j) {
            const uint32_t key = s_keys[j];
            if (key > T) {
                const uint32_t slot = atomicAdd(&s_scan[42], 1u);
                s_sel[slot] = ((unsigned long long)key << 32) | (0xffffffffu - (uint32_t)j);
            } else if (key == T) {
                ++eq;
            }
        }
        // exclusive scan of eq over the block
        int incl = eq;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((tid & 31) >= o) incl += v;
        }
        if ((tid & 31) == 31) s_scan[tid >> 5] = (uint32_t)incl;
        __syncthreads();
        if (tid < 32) {
            int w = (int)s_scan[tid], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, wi, o);
                if (tid >= o) wi += v;
            }
            s_scan[tid] = (uint32_t)(wi - w);
        }
        __syncthreads();
        int rank = (int)s_scan[tid >> 5] + incl - eq;
        const int base = k - need;   // == number of keys > T
        if (eq > 0 && rank < need) {
            for (int j = j0; j < j1 && rank < need; ++j) {
                if (s_keys[j] == T) {
                    s_sel[base + rank] = ((unsigned long long)T << 32) | (0xffffffffu - (uint32_t)j);
                    ++rank;
                }
            }
        }
        __syncthreads();
        // ---- bitonic sort, descending, of the next power of two >= k entries (padding = 0 sorts last)
        int P2 = 1;
        while (P2 < k) P2 <<= 1;
        for (int size = 2; size <= P2; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int t = tid; t < (P2 >> 1); t += kRpnThreads) {
                    const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                    const bool desc = (lo & size) == 0;
                    const unsigned long long x = s_sel[lo], y = s_sel[hi];
                    if ((x < y) == desc) {
                        s_sel[lo] = y;
                        s_sel[hi] = x;
                    }
                }
                __syncthreads();
            }
        }
    }
    // ---- decode
    F4 means, stds;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        means.v[c] = 0.f;
        stds.v[c] = 1.f;
    }
    const float *reg = lv.reg + (size_t)b * 4 * A * HW;
    const size_t out0 = (size_t)b * a.per_image + lv.out_off;
    for (int j = tid; j < k; j += kRpnThreads) {
        int idx;
        float score;
        if (select) {
            const unsigned long long e = s_sel[j];
            idx = (int)(0xffffffffu - (uint32_t)(e & 0xffffffffull));
            score = score_of((uint32_t)(e >> 32));
        } else {
            idx = j;
            const int hw = j / A, aa = j - hw * A;
            score = cls[aa * HW + hw];
            if (a.apply_sigmoid) score = sigmoid_aten(score);
        }
        const int hw = idx / A, aa = idx - hw * A;
        const float4 dl = make_float4(reg[(size_t)(aa * 4 + 0) * HW + hw], reg[(size_t)(aa * 4 + 1) * HW + hw],
                                      reg[(size_t)(aa * 4 + 2) * HW + hw], reg[(size_t)(aa * 4 + 3) * HW + hw]);
        const float4 an = *reinterpret_cast<const float4 *>(lv.anchors + (size_t)idx * 4);
        float o[4];
        decode_box(an.x, an.y, an.z, an.w, dl, means, stds, a.max_ratio, a.clamp, a.max_w, a.max_h, o);
        bool ok = true;
        if (a.min_size >= 0.f) ok = (__fsub_rn(o[2], o[0]) > a.min_size) && (__fsub_rn(o[3], o[1]) > a.min_size);
        *reinterpret_cast<float4 *>(a.boxes + (out0 + j) * 4) = make_float4(o[0], o[1], o[2], o[3]);
        a.scores[out0 + j] = score;
        a.labels[out0 + j] = l;
        a.groups[out0 + j] = ok ? b : -1;
    }
}

size_t rpn_smem_bytes(int nmax) { return sizeof(unsigned long long) * kRpnMaxK + 4 * (kRpnSubHist * 256 + 64) + 4 * (size_t)nmax; }

} // namespace

NUHTC_API int nuhtc_rpn_topk_supported(const int *H, const int *W, int L, int A, int nms_pre) {
    if (!H || !W || L < 1 || L > NUHTC_MAX_LEVELS || A < 1) return 0;
    int nmax = 0;
    for (int l = 0; l < L; ++l) {
        const long n = (long)H[l] * W[l] * A;
        if (n > (1 << 24)) return 0;
        if (nms_pre > 0 && n > nms_pre) {               // this level needs the select
            if (nms_pre > kRpnMaxK) return 0;
            nmax = (int)n > nmax ? (int)n : nmax;
        }
    }
    return rpn_smem_bytes(nmax) <= 227 * 1024 ? 1 : 0;
}

NUHTC_API int nuhtc_rpn_topk_decode(const float *const *cls, const float *const *reg, const float *const *anchors, const int *H,
                                    const int *W, int L, int B, int A, int nms_pre, int apply_sigmoid, int max_h, int max_w,
                                    double wh_ratio_clip, float min_bbox_size, float *boxes, float *scores, int64_t *labels,
                                    int32_t *groups, void *stream) {
    NUHTC_CHECK_ARG(L >= 1 && L <= NUHTC_MAX_LEVELS && B >= 0 && A >= 1 && wh_ratio_clip > 0.0, "rpn_topk_decode: bad sizes");
    if (B == 0) return NUHTC_OK;
    NUHTC_CHECK_ARG(cls && reg && anchors && H && W && boxes && scores && labels && groups, "rpn_topk_decode: null pointer");
    NUHTC_CHECK_ARG(nuhtc_rpn_topk_supported(H, W, L, A, nms_pre), "rpn_topk_decode: level too large for the shared-memory select "
                    "(nms_pre <= %d and H*W*A*4 + 26 KB <= 227 KB); use the sort path", kRpnMaxK);
    NUHTC_CHECK_ARG(((uintptr_t)boxes & 15) == 0, "rpn_topk_decode: boxes must be 16-byte aligned");
    RpnArgs a;
    memset(&a, 0, sizeof a);
    int off = 0, nmax = 0;
    for (int l = 0; l < L; ++l) {
        NUHTC_CHECK_ARG(cls[l] && reg[l] && anchors[l] && H[l] >= 1 && W[l] >= 1 && ((uintptr_t)anchors[l] & 15) == 0, "rpn_topk_decode: bad level %d", l);
        RpnLevel &v = a.lv[l];
        v.cls = cls[l];
        v.reg = reg[l];
        v.anchors = anchors[l];
        v.H = H[l];
        v.W = W[l];
        v.n = H[l] * W[l] * A;
        v.k = (nms_pre > 0 && nms_pre < v.n) ? nms_pre : v.n;
        v.out_off = off;
        off += v.k;
        if (v.k < v.n && v.n > nmax) nmax = v.n;
    }
    a.L = L;
    a.B = B;
    a.A = A;
    a.per_image = off;
    a.apply_sigmoid = apply_sigmoid;
    a.clamp = max_h > 0 && max_w > 0;
    a.max_w = (float)max_w;
    a.max_h = (float)max_h;
    a.max_ratio = (float)fabs(log(wh_ratio_clip));
    a.min_size = min_bbox_size;
    a.boxes = boxes;
    a.scores = scores;
    a.labels = labels;
    a.groups = groups;
    const size_t smem = rpn_smem_bytes(nmax);
    static size_t attr[kNuhtcMaxDevices] = {0};
    const int dev = nuhtc_device();
    if (smem > attr[dev]) {
        NUHTC_CUDA(cudaFuncSetAttribute(rpn_topk_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr[dev] = smem;
    }
    rpn_topk_decode_kernel<<<dim3(L, B), kRpnThreads, smem, (cudaStream_t)stream>>>(a);
    NUHTC_LAUNCH_CHECK();
    return NUHTC_OK;
}
