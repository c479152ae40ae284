// Greedy box NMS for sm_100a, batched over independent groups (images), bit-exact with the
// mmcv-full 1.7.2 CPU kernel `nms_cpu` and the offset arithmetic of mmcv.ops.batched_nms.
//
// Reference call sites: /root/reference/nuhtc/models/bbox_head.py:93,208,
//   /root/reference/nuhtc/core/post_processing/bbox_nms.py:83,
//   /root/reference/thirdparty/mmdetection/mmdet/core/post_processing/bbox_nms.py:86,
//   /root/reference/thirdparty/mmdetection/mmdet/models/dense_heads/rpn_head.py:232.
//
// Pipeline (all on the caller's stream, no host sync):
//   prep    : 64-bit sort key (group | descending score), per-group box count and max coordinate
//   sort    : one radix sort of (key, original index)   [cub, stable => ties keep the lower index]
//   gather  : boxes into sorted order with the class offset label*(max+1) applied in fp32 exactly as
//             batched_nms does, plus areas.  Classes are NOT sorted into segments: a class is just a
//             predicate on the pair ("sort-free per-class segments").
//   mask    : warp-ballot IoU bitmask.  A CTA owns 64 sorted rows x 256 sorted columns of one
//             group; each warp keeps 64 column boxes in registers (2 per lane), walks the 64 row
//             boxes broadcast from shared memory and turns the 32 per-lane verdicts into mask words
//             with __ballot_sync.  Only the upper triangle is produced.
//   scan    : one CTA per group walks the rows in 64-row chunks: the 64x64 diagonal block is
//             resolved serially by one thread from registers, the surviving rows' mask words are
//             OR-ed into the shared "removed" bitset by the whole CTA.  Kept original indices are
//             emitted in score order.
#include <cub/cub.cuh>

#include <stdlib.h>

#include "common.cuh"
#include "greedy_scan.cuh"

namespace {

__device__ __forceinline__ int float_ordered_int(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_int_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void nms_init_kernel(int *cnt, int *gmax, int nseg, int ngroup, int32_t *status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nseg) cnt[i] = 0;
    if (i < ngroup) gmax[i] = float_ordered_int(-INFINITY);
    if (i == 0) *status = 0;
}

// class-segmented variant: the sort segment is (group, label); only the per-group max coordinate stays per group.
// With `use_smem` the segment counts and group maxima are accumulated per CTA in shared memory first (80 000 candidates hit
// only ~80 counters: per-element global atomics serialise on them), then flushed with one atomic per touched counter.
__global__ void __launch_bounds__(256) nms_prep_seg_kernel(const float4 *__restrict__ boxes, const float *__restrict__ scores,
                                                           const int64_t *__restrict__ labels, const int32_t *__restrict__ groups,
                                                           int64_t N, int G, int Cn, int need_max, int need_nonneg, int use_smem,
                                                           uint64_t *__restrict__ keys, int32_t *__restrict__ vals, int *cnt,
                                                           int *gmax, int32_t *status) {
    extern __shared__ int sh_prep[]; // [G*Cn + 1] counts, [G] maxima
    const int S1 = G * Cn + 1;
    const int neg_inf = float_ordered_int(-INFINITY);
    if (use_smem) {
        for (int t = threadIdx.x; t < S1 + G; t += blockDim.x) sh_prep[t] = t < S1 ? 0 : neg_inf;
        __syncthreads();
    }
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < N) {
        int g = groups ? groups[i] : 0;
        const long long lab = labels[i];
        int seg;
        if (g < 0) {
            seg = G * Cn; // not a candidate
        } else if (g >= G || lab < 0 || lab >= Cn) {
            atomicExch(status, 2);
            seg = G * Cn;
            g = -1;
        } else {
            seg = g * Cn + (int)lab;
        }
        keys[i] = ((uint64_t)(uint32_t)seg << 32) | float_desc_key(scores[i]);
        vals[i] = (int32_t)i;
        atomicAdd((use_smem ? sh_prep : cnt) + seg, 1);
        if (g >= 0 && (need_max || need_nonneg)) {
            const float4 b = boxes[i];
            if (need_max) atomicMax((use_smem ? sh_prep + S1 : gmax) + g, float_ordered_int(fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w))));
            // class segments are only equivalent to the all-pairs test on offset boxes when no coordinate is negative
            if (need_nonneg && fminf(fminf(b.x, b.y), fminf(b.z, b.w)) < 0.f) atomicExch(status, 3);
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int t = threadIdx.x; t < S1 + G; t += blockDim.x) {
            const int v = sh_prep[t];
            if (t < S1) {
                if (v) atomicAdd(cnt + t, v);
            } else if (v != neg_inf) {
                atomicMax(gmax + (t - S1), v);
            }
        }
    }
}

__global__ void __launch_bounds__(256) nms_prep_kernel(const float4 *__restrict__ boxes, const float *__restrict__ scores,
                                                       const int32_t *__restrict__ groups, int64_t N, int G, int need_max,
                                                       uint64_t *__restrict__ keys, int32_t *__restrict__ vals, int *cnt,
                                                       int *gmax, int32_t *status) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool live = i < N;
    int g = 0;
    int m = float_ordered_int(-INFINITY);
    if (live) {
        if (groups) g = groups[i];
        if (g < 0) {
            g = G; // negative group = "not a candidate": parked in a trash segment behind every real group
        } else if (g >= G) {
            atomicExch(status, 2);
            g = G;
        }
        keys[i] = ((uint64_t)(uint32_t)g << 32) | float_desc_key(scores[i]);
        vals[i] = (int32_t)i;
        if (need_max) {
            const float4 b = boxes[i];
            m = float_ordered_int(fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
        }
    }
    // warp-aggregate the two atomics when the whole warp sits in one group (the common case)
    const unsigned act = __ballot_sync(0xffffffffu, live);
    const int g0 = __shfl_sync(0xffffffffu, g, __ffs(act) - 1);
    const bool uniform = __all_sync(0xffffffffu, !live || g == g0);
    if (uniform) {
        const int wm = __reduce_max_sync(0xffffffffu, m);
        if ((threadIdx.x & 31) == __ffs(act) - 1 && act) {
            atomicAdd(cnt + g0, __popc(act));
            if (need_max) atomicMax(gmax + g0, wm);
        }
    } else if (live) {
        atomicAdd(cnt + g, 1);
        if (need_max) atomicMax(gmax + g, m);
    }
}

// ---- small segments (<= 2048 candidates each): counting-sort scatter + one in-shared-memory bitonic sort per segment
// instead of five device-wide radix passes.  Key = (descending score key << 32) | index: ascending order = score descending,
// ties lower index first, exactly the order of the stable radix sort it replaces.
__global__ void nms_scatter_kernel(const uint64_t *__restrict__ keys_in, int64_t N, const int *__restrict__ seg_start,
                                   int *__restrict__ cursor, uint64_t *__restrict__ skey) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const uint64_t k = keys_in[i];
    const int seg = (int)(k >> 32);
    const int pos = seg_start[seg] + atomicAdd(cursor + seg, 1);
    skey[pos] = (k << 32) | (uint64_t)(uint32_t)i;
}

template <int CAP>
__global__ void __launch_bounds__(256) nms_blocksort_kernel(uint64_t *__restrict__ skey, const int *__restrict__ seg_start, int64_t N,
                                                            int S, int32_t *__restrict__ vals_out) {
    __shared__ uint64_t s[CAP];
    const int t = blockIdx.x, tid = threadIdx.x;
    const int s0 = seg_start[t];
    const int cnt = (t < S ? seg_start[t + 1] : (int)N) - s0;
    const uint64_t hi = (uint64_t)(uint32_t)t << 32;
    if (t == S || cnt > CAP) { // the trash segment needs no order; an over-capacity segment is flagged in status
        for (int p = tid; p < cnt; p += 256) {
            const uint64_t k = skey[s0 + p];
            vals_out[s0 + p] = (int32_t)(uint32_t)k;
            skey[s0 + p] = hi | (k >> 32);
        }
        return;
    }
    int P2 = 2;
    while (P2 < cnt) P2 <<= 1;
    for (int p = tid; p < P2; p += 256) s[p] = p < cnt ? skey[s0 + p] : ~0ull;
    __syncthreads();
    for (int k = 2; k <= P2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int p = tid; p < P2; p += 256) {
                const int q = p ^ j;
                if (q > p) {
                    const uint64_t a = s[p], b = s[q];
                    if ((a > b) == ((p & k) == 0)) {
                        s[p] = b;
                        s[q] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int p = tid; p < cnt; p += 256) {
        vals_out[s0 + p] = (int32_t)(uint32_t)s[p];
        skey[s0 + p] = hi | (s[p] >> 32); // back to the (segment, score key) form the later kernels read
    }
}

__global__ void __launch_bounds__(256) nms_gather_kernel(const float4 *__restrict__ boxes, const int64_t *__restrict__ labels,
                                                         const uint64_t *__restrict__ skeys, const int32_t *__restrict__ svals,
                                                         const int *__restrict__ gmax, int64_t N, int mode, float fo, int seg_div,
                                                         float4 *__restrict__ sbox, float *__restrict__ sarea,
                                                         int32_t *__restrict__ slab) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= N) return;
    const int i = svals[p];
    float4 b = boxes[i];
    int lab = 0;
    if (mode != NUHTC_NMS_AGNOSTIC) lab = (int)labels[i];
    if (mode == NUHTC_NMS_OFFSET || mode == NUHTC_NMS_PERCLASS) {
        const int g = (int)(skeys[p] >> 32) / seg_div; // seg_div = classes per group when class-segmented, else 1
        // batched_nms: offsets = idxs.to(boxes) * (boxes.max() + 1); boxes_for_nms = boxes + offsets[:, None]
        const float off = __fmul_rn((float)labels[i], __fadd_rn(ordered_int_float(gmax[g]), 1.0f));
        b.x = __fadd_rn(b.x, off);
        b.y = __fadd_rn(b.y, off);
        b.z = __fadd_rn(b.z, off);
        b.w = __fadd_rn(b.w, off);
    }
    sbox[p] = b;
    sarea[p] = __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), fo), __fadd_rn(__fsub_rn(b.w, b.y), fo));
    slab[p] = lab;
}

// nms_cpu's test: inter / (area_i + area_j - inter) > thr, IEEE division, no contraction
__device__ __forceinline__ bool suppresses(const float4 a, const float aa, const float4 b, const float ab, const float fo,
                                           const float thr) {
    const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
    const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    const float w = fmaxf(0.f, __fadd_rn(__fsub_rn(xx2, xx1), fo));
    const float h = fmaxf(0.f, __fadd_rn(__fsub_rn(yy2, yy1), fo));
    const float inter = __fmul_rn(w, h);
    if (inter == 0.f && thr >= 0.f) return false; // 0/x (or 0/0 = NaN) is never > thr >= 0
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ab), inter)) > thr;
}

constexpr int kMaskWarps = 4; // column blocks (64 columns each) per CTA

__global__ void __launch_bounds__(kMaskWarps * 32) nms_mask_kernel(const float4 *__restrict__ sbox, const float *__restrict__ sarea,
                                                                   const int32_t *__restrict__ slab,
                                                                   const int *__restrict__ seg_start, int wpr, int mode,
                                                                   float fo, float thr, uint64_t *__restrict__ mask) {
    const int g = blockIdx.z;
    const int s0 = seg_start[g], n = min(seg_start[g + 1] - s0, wpr * 64); // an over-capacity group is flagged in status
    const int rb = blockIdx.y;
    const int cb = blockIdx.x * kMaskWarps + (threadIdx.x >> 5);
    if (rb * 64 >= n) return;
    if ((int)(blockIdx.x * kMaskWarps + kMaskWarps - 1) < rb) return; // whole CTA below the diagonal
    __shared__ float4 r_box[64];
    __shared__ float r_area[64];
    __shared__ int r_lab[64];
    __shared__ uint64_t words[64][kMaskWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nrow = min(64, n - rb * 64);
    if (tid < 64) {
        const int p = s0 + rb * 64 + min(tid, nrow - 1);
        r_box[tid] = sbox[p];
        r_area[tid] = sarea[p];
        r_lab[tid] = slab[p];
    }
    __syncthreads();
    uint64_t w0 = 0, w1 = 0; // mask words of rows lane and lane+32 for this warp's column block
    if (cb >= rb && cb * 64 < n) {
        const int c0 = cb * 64 + lane, c1 = c0 + 32;
        const bool v0 = c0 < n, v1 = c1 < n;
        const float4 b0 = sbox[s0 + (v0 ? c0 : 0)], b1 = sbox[s0 + (v1 ? c1 : 0)];
        const float a0 = sarea[s0 + (v0 ? c0 : 0)], a1 = sarea[s0 + (v1 ? c1 : 0)];
        const int l0 = slab[s0 + (v0 ? c0 : 0)], l1 = slab[s0 + (v1 ? c1 : 0)];
        const bool perclass = mode == NUHTC_NMS_PERCLASS || mode == NUHTC_NMS_PERCLASS_RAW;
        for (int r = 0; r < nrow; ++r) {
            const float4 rbx = r_box[r];
            const float ra = r_area[r];
            const int rl = r_lab[r];
            const int row = rb * 64 + r;
            // row is the higher-scoring box (i), column the later one (j): same operand order as nms_cpu
            bool p0 = v0 && c0 > row && (!perclass || rl == l0) && suppresses(rbx, ra, b0, a0, fo, thr);
            bool p1 = v1 && c1 > row && (!perclass || rl == l1) && suppresses(rbx, ra, b1, a1, fo, thr);
            const uint32_t lo = __ballot_sync(0xffffffffu, p0), hi = __ballot_sync(0xffffffffu, p1);
            const uint64_t word = ((uint64_t)hi << 32) | lo;
            if ((r & 31) == lane) {
                if (r < 32) w0 = word; else w1 = word;
            }
        }
    }
    words[lane][warp] = w0;
    words[lane + 32][warp] = w1;
    __syncthreads();
    // 64 rows x kMaskWarps words: one 32-byte sector per row
    if (tid < 64 && tid < nrow) {
        uint64_t *dst = mask + (size_t)(s0 + rb * 64 + tid) * wpr + (size_t)blockIdx.x * kMaskWarps;
#pragma unroll
        for (int k = 0; k < kMaskWarps; ++k)
            if ((int)(blockIdx.x * kMaskWarps + k) < wpr) dst[k] = words[tid][k];
    }
}

// class-segmented NMS leaves one score-sorted kept list per (group, class); the group's list is their merge.  The rank of
// a kept box inside its group = its rank in its own list + the number of boxes of the other classes that sort before it
// (binary search on (score key, index), the order of the radix sort).
__global__ void __launch_bounds__(256) nms_merge_segments_kernel(const uint64_t *__restrict__ skeys, const int *__restrict__ seg_start,
                                                                const int64_t *__restrict__ seg_count, const int64_t *__restrict__ keep_seg,
                                                                const uint64_t *__restrict__ kkeys, int64_t N, int G, int Cn,
                                                                int64_t *__restrict__ keep, int64_t *__restrict__ group_start,
                                                                int64_t *__restrict__ group_count) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p < G) {
        int64_t tot = 0;
        for (int c = 0; c < Cn; ++c) tot += seg_count[p * Cn + c];
        group_start[p] = seg_start[p * Cn];
        group_count[p] = tot;
    }
    if (p >= N) return;
    const int seg = (int)(skeys[p] >> 32);
    if (seg >= G * Cn) return;
    const int r = (int)(p - seg_start[seg]);
    if (r >= seg_count[seg]) return;
    const int s0 = seg_start[seg];
    const uint32_t key = (uint32_t)kkeys[s0 + r];
    const int64_t idx = keep_seg[s0 + r];
    const int g = seg / Cn;
    int rank = r;
    for (int c = 0; c < Cn; ++c) {
        const int o = g * Cn + c;
        if (o == seg) continue;
        const int b0 = seg_start[o];
        int lo = 0, hi = (int)seg_count[o];
        while (lo < hi) { // first element of list o that does not sort before (key, idx)
            const int mid = (lo + hi) >> 1;
            const uint32_t k2 = (uint32_t)kkeys[b0 + mid];
            const bool before = k2 < key || (k2 == key && keep_seg[b0 + mid] < idx);
            if (before) lo = mid + 1; else hi = mid;
        }
        rank += lo;
    }
    keep[seg_start[g * Cn] + rank] = idx;
}

struct NmsWs {
    uint64_t *keys_in, *keys_out;
    int32_t *vals_in, *vals_out;
    float4 *sbox;
    float *sarea;
    int32_t *slab;
    int *cnt, *gmax, *seg_start;
    uint64_t *mask;
    int64_t *keep_seg, *seg_count; // class-segmented path
    uint64_t *kkeys;
    void *cub_tmp;
    size_t cub_bytes;
    size_t total;
};

static size_t cub_sort_bytes(int64_t N) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const int32_t *)nullptr,
                                    (int32_t *)nullptr, N > 0 ? N : 1, 0, 64, (cudaStream_t)0);
    return bytes;
}

static NmsWs nms_layout(void *ws, int64_t N, int G, int64_t M, int Cn = 0) {
    NmsWs L;
    char *p = (char *)ws;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void *r = p ? p + off : nullptr;
        off += align_up(bytes, 256);
        return r;
    };
    const int64_t wpr = (M + 63) / 64;
    L.keys_in = (uint64_t *)take(sizeof(uint64_t) * N);
    L.keys_out = (uint64_t *)take(sizeof(uint64_t) * N);
    L.vals_in = (int32_t *)take(sizeof(int32_t) * N);
    L.vals_out = (int32_t *)take(sizeof(int32_t) * N);
    L.sbox = (float4 *)take(sizeof(float4) * N);
    L.sarea = (float *)take(sizeof(float) * N);
    L.slab = (int32_t *)take(sizeof(int32_t) * N);
    const int S = Cn > 0 ? G * Cn : G; // sort segments
    L.cnt = (int *)take(sizeof(int) * (S + 1));
    L.gmax = (int *)take(sizeof(int) * (G + 1));
    L.seg_start = (int *)take(sizeof(int) * (S + 1));
    L.mask = (uint64_t *)take(sizeof(uint64_t) * (size_t)N * wpr);
    L.keep_seg = (int64_t *)take(Cn > 0 ? sizeof(int64_t) * N : 8);
    L.seg_count = (int64_t *)take(sizeof(int64_t) * (S + 1));
    L.kkeys = (uint64_t *)take(Cn > 0 ? sizeof(uint64_t) * N : 8);
    L.cub_bytes = cub_sort_bytes(N);
    L.cub_tmp = take(L.cub_bytes);
    L.total = off;
    return L;
}

} // namespace

NUHTC_API size_t nuhtc_nms_workspace_bytes(int64_t N, int num_groups, int64_t max_group_size, int num_classes) {
    if (N <= 0 || num_groups <= 0) return 256;
    if (max_group_size > N) max_group_size = N;
    if (max_group_size < 1) max_group_size = 1;
    return nms_layout(nullptr, N, num_groups, max_group_size, num_classes > 0 ? num_classes : 0).total;
}

NUHTC_API int nuhtc_nms(const float *boxes, const float *scores, const int64_t *labels, const int32_t *groups, int64_t N,
                        int num_groups, int64_t max_group_size, float iou_thr, int offset, int mode, int num_classes, int64_t *keep,
                        int64_t *group_start, int64_t *group_count, int32_t *status, void *ws, size_t ws_bytes, void *stream) {
    NUHTC_CHECK_ARG(N >= 0 && N < (1ll << 31), "nms: N=%lld out of range", (long long)N);
    NUHTC_CHECK_ARG(num_groups >= 1 && num_groups <= 65535, "nms: num_groups=%d out of range", num_groups);
    NUHTC_CHECK_ARG(mode >= NUHTC_NMS_AGNOSTIC && mode <= NUHTC_NMS_PERCLASS_RAW, "nms: bad mode %d", mode);
    NUHTC_CHECK_ARG(offset == 0 || offset == 1, "nms: offset must be 0 or 1");
    NUHTC_CHECK_ARG(group_start && group_count && status, "nms: null output pointer");
    NUHTC_CHECK_ARG(mode == NUHTC_NMS_AGNOSTIC || labels != nullptr || N == 0, "nms: labels required for this mode");
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) {
        NUHTC_CUDA(cudaMemsetAsync(group_start, 0, sizeof(int64_t) * num_groups, st));
        NUHTC_CUDA(cudaMemsetAsync(group_count, 0, sizeof(int64_t) * num_groups, st));
        NUHTC_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
        return NUHTC_OK;
    }
    NUHTC_CHECK_ARG(boxes && scores && keep && ws, "nms: null pointer");
    NUHTC_CHECK_ARG(((uintptr_t)boxes) % 16 == 0, "nms: boxes must be 16-byte aligned");
    if (max_group_size > N) max_group_size = N;
    if (max_group_size < 1) max_group_size = 1;
    const int G = num_groups;
    const int Cn = (num_classes > 0 && mode != NUHTC_NMS_AGNOSTIC) ? num_classes : 0; // class segments
    NUHTC_CHECK_ARG(Cn == 0 || (long)G * Cn <= 65535, "nms: num_groups*num_classes=%ld too large", (long)G * Cn);
    NmsWs L = nms_layout(ws, N, G, max_group_size, Cn);
    if (L.total > ws_bytes) {
        nuhtc_set_error("nms: workspace %zu < required %zu", ws_bytes, L.total);
        return NUHTC_EWORKSPACE;
    }
    const int wpr = (int)((max_group_size + 63) / 64);
    NUHTC_CHECK_ARG(wpr <= 65535, "nms: max_group_size too large for one launch");
    NUHTC_CHECK_ARG((size_t)wpr * 8 <= 200 * 1024, "nms: group too large for the shared removed-set");
    const float fo = (float)offset;
    const int nb = (int)((N + 255) / 256);
    const int S = Cn > 0 ? G * Cn : G;
    const bool offs = mode == NUHTC_NMS_OFFSET || mode == NUHTC_NMS_PERCLASS;
    nms_init_kernel<<<(S + 256) / 256, 256, 0, st>>>(L.cnt, L.gmax, S + 1, G + 1, status);
    if (Cn > 0) {
        const size_t prep_smem = sizeof(int) * (size_t)(S + 1 + G);
        const int use_smem = prep_smem <= 32 * 1024;
        nms_prep_seg_kernel<<<nb, 256, use_smem ? prep_smem : 0, st>>>((const float4 *)boxes, scores, labels, groups, N, G, Cn, offs,
                                                                       mode == NUHTC_NMS_OFFSET, use_smem, L.keys_in, L.vals_in, L.cnt,
                                                                       L.gmax, status);
    } else
        nms_prep_kernel<<<nb, 256, 0, st>>>((const float4 *)boxes, scores, groups, N, G, offs, L.keys_in, L.vals_in, L.cnt, L.gmax, status);
    segments_kernel<int64_t><<<1, 256, 0, st>>>(L.cnt, S, max_group_size, L.seg_start, Cn > 0 ? L.seg_count : group_start, status);
    static const bool blocksort = !(getenv("NUHTC_NMS_BLOCKSORT") && getenv("NUHTC_NMS_BLOCKSORT")[0] == '0'); // A/B switch
    if (blocksort && max_group_size <= 2048) {
        NUHTC_CUDA(cudaMemsetAsync(L.cnt, 0, sizeof(int) * (S + 1), st)); // the counts are consumed: reused as scatter cursors
        nms_scatter_kernel<<<nb, 256, 0, st>>>(L.keys_in, N, L.seg_start, L.cnt, L.keys_out);
        if (max_group_size <= 512) nms_blocksort_kernel<512><<<S + 1, 256, 0, st>>>(L.keys_out, L.seg_start, N, S, L.vals_out);
        else if (max_group_size <= 1024) nms_blocksort_kernel<1024><<<S + 1, 256, 0, st>>>(L.keys_out, L.seg_start, N, S, L.vals_out);
        else nms_blocksort_kernel<2048><<<S + 1, 256, 0, st>>>(L.keys_out, L.seg_start, N, S, L.vals_out);
    } else {
        int gbits = 0;
        while ((1 << gbits) < S + 1) ++gbits;
        size_t cub_bytes = L.cub_bytes;
        NUHTC_CUDA(cub::DeviceRadixSort::SortPairs(L.cub_tmp, cub_bytes, L.keys_in, L.keys_out, L.vals_in, L.vals_out, N, 0,
                                                   32 + gbits, st));
    }
    nms_gather_kernel<<<nb, 256, 0, st>>>((const float4 *)boxes, labels, L.keys_out, L.vals_out, L.gmax, N, mode, fo, Cn > 0 ? Cn : 1,
                                          L.sbox, L.sarea, L.slab);
    dim3 mgrid((wpr + kMaskWarps - 1) / kMaskWarps, wpr, S);
    nms_mask_kernel<<<mgrid, kMaskWarps * 32, 0, st>>>(L.sbox, L.sarea, L.slab, L.seg_start, wpr, mode, fo, iou_thr, L.mask);
    NUHTC_LAUNCH_CHECK();
    if (Cn > 0) {
        int rc = launch_greedy_scan<int64_t>(L.mask, L.vals_out, L.seg_start, wpr, S, L.keep_seg, L.seg_count, st, L.keys_out, L.kkeys);
        if (rc) return rc;
        const long nt = N > G ? N : G;
        nms_merge_segments_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(L.keys_out, L.seg_start, L.seg_count, L.keep_seg, L.kkeys, N,
                                                                              G, Cn, keep, group_start, group_count);
        NUHTC_LAUNCH_CHECK();
        return NUHTC_OK;
    }
    return launch_greedy_scan<int64_t>(L.mask, L.vals_out, L.seg_start, wpr, G, keep, group_count, st);
}
