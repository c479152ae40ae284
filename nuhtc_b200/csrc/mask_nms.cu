// Per-tile mask NMS for sm_100a, bit-exact with tools/infer_wsi.py:60-84
// (pycocotools rleEncode + rleIou(iscrowd=0) + the greedy double loop), batched over tiles.
//
// Masks are held as bit rows (1 bit per pixel: 8 KB for a 256x256 tile instead of 64 KB), so the
// pairwise intersection is a popcount over the rows/words where the two tight boxes overlap;
// union = area_i + area_j - inter, IoU = (double)inter/(double)union (inter==0 -> 0), suppressed when
// IoU > thr -- integer pixel counts and one IEEE double division, exactly what rleIou produces.
// The suppression bitmask + greedy scan machinery is shared with the box NMS.
#include <cub/cub.cuh>

#include "common.cuh"
#include "greedy_scan.cuh"

namespace {

// 4 mask bytes -> 4 bits (nonzero test)
__device__ __forceinline__ uint32_t nz4(uint32_t x) {
    const uint32_t m = __vcmpne4(x, 0u);
    return (m & 1u) | ((m >> 7) & 2u) | ((m >> 14) & 4u) | ((m >> 21) & 8u);
}

__global__ void pack_init_kernel(int32_t *area, int32_t *bbox, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        if (area) area[i] = 0;
        if (bbox) {
            bbox[4 * i + 0] = 1 << 30;
            bbox[4 * i + 1] = 1 << 30;
            bbox[4 * i + 2] = 0;
            bbox[4 * i + 3] = 0;
        }
    }
}
__global__ void pack_fini_kernel(int32_t *bbox, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && bbox && bbox[4 * i + 2] == 0) bbox[4 * i + 0] = bbox[4 * i + 1] = bbox[4 * i + 3] = 0;
}

__device__ __forceinline__ void pack_stats(uint64_t word, int i, int y, int xw, int32_t *area, int32_t *bbox) {
    if (!word) return;
    if (area) atomicAdd(area + i, __popcll(word));
    if (bbox) {
        atomicMin(bbox + 4 * i + 0, xw + __ffsll((long long)word) - 1);
        atomicMin(bbox + 4 * i + 1, y);
        atomicMax(bbox + 4 * i + 2, xw + 64 - __clzll((long long)word));
        atomicMax(bbox + 4 * i + 3, y + 1);
    }
}

// w % 64 == 0 and 16-byte aligned: each lane converts 16 pixels, 4 lanes make a word
__global__ void __launch_bounds__(256) pack_vec_kernel(const uint4 *__restrict__ masks, int64_t nvec, int h, int wpm,
                                                      uint64_t *__restrict__ bits, int32_t *area, int32_t *bbox) {
    const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool live = v < nvec;
    uint4 x = make_uint4(0, 0, 0, 0);
    if (live) x = __ldg(masks + v);
    uint64_t nib = (uint64_t)(nz4(x.x) | (nz4(x.y) << 4) | (nz4(x.z) << 8) | (nz4(x.w) << 12));
    const int lane = threadIdx.x & 31;
    nib <<= 16 * (lane & 3);
    nib |= __shfl_xor_sync(0xffffffffu, nib, 1);
    nib |= __shfl_xor_sync(0xffffffffu, nib, 2);
    if (live && (lane & 3) == 0) {
        const int64_t word = v >> 2;
        bits[word] = nib;
        const int64_t row = word / wpm;
        pack_stats(nib, (int)(row / h), (int)(row % h), (int)(word % wpm) * 64, area, bbox);
    }
}

__global__ void __launch_bounds__(256) pack_generic_kernel(const uint8_t *__restrict__ masks, int64_t nwords, int h, int w, int wpm,
                                                          uint64_t *__restrict__ bits, int32_t *area, int32_t *bbox) {
    const int64_t word = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (word >= nwords) return;
    const int64_t row = word / wpm;
    const int xw = (int)(word % wpm) * 64;
    const uint8_t *src = masks + row * w + xw;
    uint64_t bitsw = 0;
    const int lim = min(64, w - xw);
    for (int k = 0; k < lim; ++k) bitsw |= (uint64_t)(src[k] != 0) << k;
    bits[word] = bitsw;
    pack_stats(bitsw, (int)(row / h), (int)(row % h), xw, area, bbox);
}

__global__ void mnms_init_kernel(int *cnt, int T, int32_t *status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < T) cnt[i] = 0;
    if (i == 0) *status = 0;
}

// input position p maps to index n-1-p so that the stable sort leaves equal scores in
// DESCENDING index order, which is what np.argsort(scores)[::-1] yields for a stable argsort
__global__ void mnms_prep_kernel(const float *__restrict__ scores, const int32_t *__restrict__ tile, int n, int T,
                                 uint64_t *__restrict__ keys, int32_t *__restrict__ vals, int *cnt, int32_t *status) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int i = n - 1 - p;
    int t = tile ? tile[i] : 0;
    if (t < 0) {
        t = T; // negative tile = filtered-out mask: parked in a trash segment behind every real tile
    } else if (t >= T) {
        atomicExch(status, 2);
        t = T;
    }
    keys[p] = ((uint64_t)(uint32_t)t << 32) | float_desc_key(scores[i]);
    vals[p] = i;
    atomicAdd(cnt + t, 1);
}

// ---- small tiles (<= 2048 masks each): one counting-sort scatter + one in-shared-memory bitonic sort per tile instead of
// five device-wide radix passes (a tile holds a few hundred masks; the passes are pure launch latency at that size).
// Key = (descending score key << 32) | (0xffffffff - index): ascending order = score descending, ties higher index first,
// the order the stable radix sort of the reversed input produces.
__global__ void mnms_scatter_kernel(const float *__restrict__ scores, const int32_t *__restrict__ tile, int n, int T,
                                    const int *__restrict__ seg_start, int *__restrict__ cursor, uint64_t *__restrict__ skey) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int t = tile ? tile[i] : 0;
    if (t < 0 || t >= T) t = T;
    const int pos = seg_start[t] + atomicAdd(cursor + t, 1);
    skey[pos] = ((uint64_t)float_desc_key(scores[i]) << 32) | (uint64_t)(0xffffffffu - (uint32_t)i);
}

template <int CAP>
__global__ void __launch_bounds__(256) mnms_blocksort_kernel(const uint64_t *__restrict__ skey, const int *__restrict__ seg_start,
                                                             int n, int T, int32_t *__restrict__ vals_out) {
    __shared__ uint64_t s[CAP];
    const int t = blockIdx.x, tid = threadIdx.x;
    const int s0 = seg_start[t];
    const int cnt = (t < T ? seg_start[t + 1] : n) - s0;
    if (t == T || cnt > CAP) { // the trash segment needs no order; an over-capacity tile is flagged in status by segments_kernel
        for (int p = tid; p < cnt; p += 256) vals_out[s0 + p] = (int32_t)(0xffffffffu - (uint32_t)skey[s0 + p]);
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
    for (int p = tid; p < cnt; p += 256) vals_out[s0 + p] = (int32_t)(0xffffffffu - (uint32_t)s[p]);
}

__global__ void mnms_gather_kernel(const int32_t *__restrict__ area, const int32_t *__restrict__ bbox,
                                   const int32_t *__restrict__ svals, int n, int32_t *__restrict__ sarea,
                                   int4 *__restrict__ sbbox) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int i = svals[p];
    sarea[p] = area[i];
    sbbox[p] = make_int4(bbox[4 * i], bbox[4 * i + 1], bbox[4 * i + 2], bbox[4 * i + 3]);
}

__device__ __forceinline__ bool mask_suppresses(const uint64_t *__restrict__ bits, int h, int wpm, int oi, int4 bi, int ai,
                                                int oj, int4 bj, int aj, double thr) {
    const int x0 = max(bi.x, bj.x), x1 = min(bi.z, bj.z), y0 = max(bi.y, bj.y), y1 = min(bi.w, bj.w);
    long long inter = 0;
    if (x1 > x0 && y1 > y0) { // rleIou's bbIou prefilter: only boxes with positive overlap are walked
        const int w0 = x0 >> 6, w1 = (x1 - 1) >> 6;
        const uint64_t *pi = bits + ((size_t)oi * h + y0) * wpm, *pj = bits + ((size_t)oj * h + y0) * wpm;
        // the rows are independent: 8 of them (16 loads) are kept in flight per lane, otherwise the walk is one L2 round
        // trip per row (integer sums: the order does not matter)
        for (int w = w0; w <= w1; ++w) {
            const uint64_t *a = pi + w, *b = pj + w;
            int y = y0, acc = 0;
            for (; y < y1; y += 8, a += 8 * (size_t)wpm, b += 8 * (size_t)wpm) {
                uint64_t va[8], vb[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const bool live = y + u < y1;
                    va[u] = live ? __ldg(a + (size_t)u * wpm) : 0ull;
                    vb[u] = live ? __ldg(b + (size_t)u * wpm) : 0ull;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) acc += __popcll(va[u] & vb[u]);
            }
            inter += acc;
        }
    }
    if (inter == 0) return 0.0 > thr; // rleIou: i==0 -> u=1 -> 0/1
    const long long uni = (long long)ai + (long long)aj - inter;
    return __ddiv_rn((double)inter, (double)uni) > thr;
}

constexpr int kMWarps = 8; // warps per CTA; a CTA owns one 64 x 64 block of the pair matrix, warp w its rows w, w+8, ...

// The pair test is a data-dependent loop over bit rows (most pairs are rejected on their boxes, overlapping ones read up to
// 2 x h words), so the kernel is latency bound: one 64 x 64 block per CTA with its rows spread over 8 warps keeps ~30 warps
// per SM busy for 16 tiles x 500 masks (4 warps walking 64 rows each left most SMs with 1-3 warps: 156 us -> see DESIGN).
__global__ void __launch_bounds__(kMWarps * 32) mnms_mask_kernel(const uint64_t *__restrict__ bits, int h, int wpm,
                                                                 const int32_t *__restrict__ svals, const int32_t *__restrict__ sarea,
                                                                 const int4 *__restrict__ sbbox, const int *__restrict__ seg_start,
                                                                 int wpr, double thr, uint64_t *__restrict__ mask) {
    const int g = blockIdx.z;
    const int s0 = seg_start[g], n = min(seg_start[g + 1] - s0, wpr * 64);
    const int rb = blockIdx.y, cb = blockIdx.x;
    if (rb * 64 >= n || cb < rb) return;
    __shared__ int4 r_bb[64];
    __shared__ int r_area[64], r_idx[64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nrow = min(64, n - rb * 64);
    if (tid < 64) {
        const int p = s0 + rb * 64 + min(tid, nrow - 1);
        r_bb[tid] = sbbox[p];
        r_area[tid] = sarea[p];
        r_idx[tid] = svals[p];
    }
    __syncthreads();
    if (cb * 64 >= n) { // no columns: the rows' words of this block are zero
        for (int r = tid; r < nrow; r += kMWarps * 32) mask[(size_t)(s0 + rb * 64 + r) * wpr + cb] = 0ull;
        return;
    }
    const int c0 = cb * 64 + lane, c1 = c0 + 32;
    const bool v0 = c0 < n, v1 = c1 < n;
    const int4 b0 = sbbox[s0 + (v0 ? c0 : 0)], b1 = sbbox[s0 + (v1 ? c1 : 0)];
    const int a0 = sarea[s0 + (v0 ? c0 : 0)], a1 = sarea[s0 + (v1 ? c1 : 0)];
    const int o0 = svals[s0 + (v0 ? c0 : 0)], o1 = svals[s0 + (v1 ? c1 : 0)];
    for (int r = warp; r < nrow; r += kMWarps) {
        const int row = rb * 64 + r;
        const bool p0 = v0 && c0 > row && mask_suppresses(bits, h, wpm, r_idx[r], r_bb[r], r_area[r], o0, b0, a0, thr);
        const bool p1 = v1 && c1 > row && mask_suppresses(bits, h, wpm, r_idx[r], r_bb[r], r_area[r], o1, b1, a1, thr);
        const uint32_t lo = __ballot_sync(0xffffffffu, p0), hi = __ballot_sync(0xffffffffu, p1);
        if (lane == 0) mask[(size_t)(s0 + row) * wpr + cb] = ((uint64_t)hi << 32) | lo;
    }
}

struct MnmsWs {
    uint64_t *keys_in, *keys_out;
    int32_t *vals_in, *vals_out, *sarea;
    int4 *sbbox;
    int *cnt, *seg_start;
    uint64_t *mask;
    void *cub_tmp;
    size_t cub_bytes, total;
};

static MnmsWs mnms_layout(void *ws, int n, int T, int M) {
    MnmsWs L;
    char *p = (char *)ws;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void *r = p ? p + off : nullptr;
        off += align_up(bytes, 256);
        return r;
    };
    const int wpr = (M + 63) / 64;
    L.keys_in = (uint64_t *)take(sizeof(uint64_t) * n);
    L.keys_out = (uint64_t *)take(sizeof(uint64_t) * n);
    L.vals_in = (int32_t *)take(sizeof(int32_t) * n);
    L.vals_out = (int32_t *)take(sizeof(int32_t) * n);
    L.sarea = (int32_t *)take(sizeof(int32_t) * n);
    L.sbbox = (int4 *)take(sizeof(int4) * n);
    L.cnt = (int *)take(sizeof(int) * (T + 1));
    L.seg_start = (int *)take(sizeof(int) * (T + 1));
    L.mask = (uint64_t *)take(sizeof(uint64_t) * (size_t)n * wpr);
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const int32_t *)nullptr,
                                    (int32_t *)nullptr, n > 0 ? n : 1, 0, 64, (cudaStream_t)0);
    L.cub_bytes = bytes;
    L.cub_tmp = take(bytes);
    L.total = off;
    return L;
}

} // namespace

NUHTC_API int nuhtc_pack_masks(const uint8_t *masks, int n, int h, int w, uint64_t *bits, int32_t *area, int32_t *bbox,
                               void *stream) {
    NUHTC_CHECK_ARG(n >= 0 && h >= 1 && w >= 1, "pack_masks: bad sizes");
    if (n == 0) return NUHTC_OK;
    NUHTC_CHECK_ARG(masks && bits, "pack_masks: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int wpm = (w + 63) / 64;
    if (area || bbox) pack_init_kernel<<<(n + 255) / 256, 256, 0, st>>>(area, bbox, n);
    const int64_t nwords = (int64_t)n * h * wpm;
    if (w % 64 == 0 && ((uintptr_t)masks) % 16 == 0) {
        const int64_t nvec = nwords * 4;
        pack_vec_kernel<<<(unsigned)((nvec + 255) / 256), 256, 0, st>>>((const uint4 *)masks, nvec, h, wpm, bits, area, bbox);
    } else {
        pack_generic_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, st>>>(masks, nwords, h, w, wpm, bits, area, bbox);
    }
    if (bbox) pack_fini_kernel<<<(n + 255) / 256, 256, 0, st>>>(bbox, n);
    NUHTC_LAUNCH_CHECK();
    return NUHTC_OK;
}

NUHTC_API size_t nuhtc_mask_nms_workspace_bytes(int n, int num_tiles, int max_tile_size) {
    if (n <= 0 || num_tiles <= 0) return 256;
    if (max_tile_size > n) max_tile_size = n;
    if (max_tile_size < 1) max_tile_size = 1;
    return mnms_layout(nullptr, n, num_tiles, max_tile_size).total;
}

NUHTC_API int nuhtc_mask_nms(const uint64_t *bits, const int32_t *area, const int32_t *bbox, const float *scores,
                             const int32_t *tile, int n, int num_tiles, int max_tile_size, int h, int w, double thr, int32_t *keep,
                             int32_t *tile_start, int32_t *tile_count, int32_t *status, void *ws, size_t ws_bytes, void *stream) {
    NUHTC_CHECK_ARG(n >= 0 && num_tiles >= 1 && num_tiles <= 65535 && h >= 1 && w >= 1, "mask_nms: bad sizes");
    NUHTC_CHECK_ARG(tile_start && tile_count && status, "mask_nms: null output pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        NUHTC_CUDA(cudaMemsetAsync(tile_start, 0, sizeof(int32_t) * num_tiles, st));
        NUHTC_CUDA(cudaMemsetAsync(tile_count, 0, sizeof(int32_t) * num_tiles, st));
        NUHTC_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
        return NUHTC_OK;
    }
    NUHTC_CHECK_ARG(bits && area && bbox && scores && keep && ws, "mask_nms: null pointer");
    if (max_tile_size > n) max_tile_size = n;
    if (max_tile_size < 1) max_tile_size = 1;
    const int T = num_tiles;
    MnmsWs L = mnms_layout(ws, n, T, max_tile_size);
    if (L.total > ws_bytes) {
        nuhtc_set_error("mask_nms: workspace %zu < required %zu", ws_bytes, L.total);
        return NUHTC_EWORKSPACE;
    }
    const int wpr = (max_tile_size + 63) / 64;
    const int wpm = (w + 63) / 64;
    NUHTC_CHECK_ARG(wpr <= 65535, "mask_nms: max_tile_size too large");
    const int nb = (n + 255) / 256;
    mnms_init_kernel<<<(T + 256) / 256, 256, 0, st>>>(L.cnt, T + 1, status);
    mnms_prep_kernel<<<nb, 256, 0, st>>>(scores, tile, n, T, L.keys_in, L.vals_in, L.cnt, status);
    segments_kernel<int32_t><<<1, 256, 0, st>>>(L.cnt, T, max_tile_size, L.seg_start, tile_start, status);
    if (max_tile_size <= 2048) {
        NUHTC_CUDA(cudaMemsetAsync(L.cnt, 0, sizeof(int) * (T + 1), st)); // the counts are consumed: reused as scatter cursors
        mnms_scatter_kernel<<<nb, 256, 0, st>>>(scores, tile, n, T, L.seg_start, L.cnt, L.keys_out);
        if (max_tile_size <= 512) mnms_blocksort_kernel<512><<<T + 1, 256, 0, st>>>(L.keys_out, L.seg_start, n, T, L.vals_out);
        else if (max_tile_size <= 1024) mnms_blocksort_kernel<1024><<<T + 1, 256, 0, st>>>(L.keys_out, L.seg_start, n, T, L.vals_out);
        else mnms_blocksort_kernel<2048><<<T + 1, 256, 0, st>>>(L.keys_out, L.seg_start, n, T, L.vals_out);
    } else {
        int tbits = 0;
        while ((1 << tbits) < T + 1) ++tbits;
        size_t cub_bytes = L.cub_bytes;
        NUHTC_CUDA(cub::DeviceRadixSort::SortPairs(L.cub_tmp, cub_bytes, L.keys_in, L.keys_out, L.vals_in, L.vals_out, n, 0,
                                                   32 + tbits, st));
    }
    mnms_gather_kernel<<<nb, 256, 0, st>>>(area, bbox, L.vals_out, n, L.sarea, L.sbbox);
    dim3 mgrid(wpr, wpr, T);
    mnms_mask_kernel<<<mgrid, kMWarps * 32, 0, st>>>(bits, h, wpm, L.vals_out, L.sarea, L.sbbox, L.seg_start, wpr, thr, L.mask);
    NUHTC_LAUNCH_CHECK();
    return launch_greedy_scan<int32_t>(L.mask, L.vals_out, L.seg_start, wpr, T, keep, tile_count, st);
}
