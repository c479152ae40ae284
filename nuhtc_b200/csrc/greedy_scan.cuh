// Greedy suppression scan over a precomputed upper-triangular suppression bitmask, shared by the
// box NMS (nms.cu) and the per-tile mask NMS (mask_nms.cu).
//   mask  [rows][wpr] uint64: bit j of word w of sorted row i set <=> row i suppresses row 64*w+j (> i),
//         rows and columns are segment-local sorted positions
//   one CTA per segment walks the rows in 64-row chunks: the 64x64 diagonal block is resolved
//   serially by one thread, the surviving rows' words are OR-ed into the shared "removed" bitset by
//   the whole CTA.  Kept original indices (svals) are emitted in sorted order.
#pragma once
#include <cub/block/block_scan.cuh>

#include "common.cuh"

// score -> radix key that sorts descending (stable sort => ties keep input order)
__device__ __forceinline__ uint32_t float_desc_key(float s) {
    uint32_t u = __float_as_uint(s);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u); // ascending-monotone
    return ~u;                                      // descending
}

// exclusive scan of the group counts (G is small: one CTA), capacity check
template <typename OutT>
__global__ void segments_kernel(const int *cnt, int G, int64_t max_group, int *seg_start, OutT *group_start,
                                int32_t *status) {
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    typedef cub::BlockScan<int, 256> Scan;
    __shared__ typename Scan::TempStorage tmp;
    for (int base = 0; base < G; base += 256) {
        const int i = base + threadIdx.x;
        const int c = i < G ? cnt[i] : 0;
        if (i < G && c > max_group) atomicExch(status, 1);
        int ex, tot;
        Scan(tmp).ExclusiveSum(c, ex, tot);
        if (i < G) {
            seg_start[i] = carry + ex;
            group_start[i] = (OutT)(carry + ex);
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) seg_start[G] = carry;
}


constexpr int kScanThreads = 1024;

template <typename OutT>
__global__ void __launch_bounds__(kScanThreads) greedy_scan_kernel(const uint64_t *__restrict__ mask, const int32_t *__restrict__ svals,
                                                                   const int *__restrict__ seg_start, int wpr,
                                                                   OutT *__restrict__ keep, OutT *__restrict__ group_count,
                                                                   const uint64_t *__restrict__ skeys, uint64_t *__restrict__ kkeys) {
    extern __shared__ unsigned long long removed[]; // [wpr]
    __shared__ unsigned long long s_diag[64];
    __shared__ unsigned long long s_keepbits;
    const int g = blockIdx.x;
    const int s0 = seg_start[g], n = min(seg_start[g + 1] - s0, wpr * 64);
    const int tid = threadIdx.x;
    const int nchunk = (n + 63) / 64;
    for (int w = tid; w < nchunk; w += kScanThreads) removed[w] = 0ull;
    int kept = 0;
    __syncthreads();
    for (int c = 0; c < nchunk; ++c) {
        const int nrow = min(64, n - c * 64);
        if (tid < 64) s_diag[tid] = tid < nrow ? mask[(size_t)(s0 + c * 64 + tid) * wpr + c] : 0ull;
        __syncthreads();
        if (tid == 0) {
            unsigned long long rem = removed[c];
            if (nrow < 64) rem |= ~0ull << nrow;
            unsigned long long kb = 0ull;
#pragma unroll
            for (int b = 0; b < 64; ++b) {
                const bool k = !((rem >> b) & 1ull);
                kb |= (unsigned long long)k << b;
                rem |= k ? s_diag[b] : 0ull;
            }
            s_keepbits = kb;
        }
        __syncthreads();
        const unsigned long long kb = s_keepbits;
        if (tid < 64 && ((kb >> tid) & 1ull)) {
            const int rank = __popcll(kb & ((1ull << tid) - 1ull));
            keep[s0 + kept + rank] = (OutT)svals[s0 + c * 64 + tid];
            if (kkeys) kkeys[s0 + kept + rank] = skeys[s0 + c * 64 + tid]; // sort key of the kept row (segment merge)
        }
        kept += __popcll(kb);
        // OR the kept rows into `removed` for the words to the right of the diagonal.
        // thread = (row slot rs of 16, word lane wl of 64)
        const int nw = nchunk - (c + 1);
        if (nw > 0 && kb) {
            const int wl = tid & 63, rs = tid >> 6;
            for (int wbase = 0; wbase < nw; wbase += 64) {
                const int w = c + 1 + wbase + wl;
                if (w < nchunk) {
                    unsigned long long acc = 0ull;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int b = rs * 4 + u;
                        if ((kb >> b) & 1ull) acc |= mask[(size_t)(s0 + c * 64 + b) * wpr + w];
                    }
                    if (acc) atomicOr(&removed[w], acc);
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0) group_count[g] = (OutT)kept;
}


template <typename OutT>
static int launch_greedy_scan(const uint64_t *mask, const int32_t *svals, const int *seg_start, int wpr, int G, OutT *keep,
                              OutT *group_count, cudaStream_t st, const uint64_t *skeys = nullptr, uint64_t *kkeys = nullptr) {
    static bool attr_done_dev[kNuhtcMaxDevices] = {false};
    bool &attr_done = attr_done_dev[nuhtc_device()];   // the attribute is per device
    if ((size_t)wpr * 8 > 200 * 1024) {
        nuhtc_set_error("greedy scan: segment too large for the shared removed-set");
        return NUHTC_EINVAL;
    }
    if (!attr_done) {
        NUHTC_CUDA(cudaFuncSetAttribute(greedy_scan_kernel<OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    greedy_scan_kernel<OutT><<<G, kScanThreads, (size_t)wpr * 8, st>>>(mask, svals, seg_start, wpr, keep, group_count, skeys, kkeys);
    NUHTC_LAUNCH_CHECK();
    return NUHTC_OK;
}
