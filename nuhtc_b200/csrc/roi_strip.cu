// Strip-shared RoIAlign forward for sm_100a -- the default path for 7x7 / 14x14 outputs (C % 32 == 0).
//
// Replaces mmcv-full 1.7.2 `roi_align_forward` (avg pool, aligned) as driven per FPN level by
//   /root/reference/thirdparty/mmdetection/mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py:58-115
// for ALL levels and RoIs in one launch sequence, like roi_align.cu, but organised around the FEATURE MAP instead of
// the RoI.  Round 1's kernel pulled every RoI's own window through L2 (16 000 windows x 121 KB = 1.9 GB for a 0.27 GB
// level: each cell crossed the L2->SM fabric ~7 times, profiles/r01_roialign_pipe.md) and was bound by the L2 slices.
// Here a CTA owns a vertical STRIP of one image for one group of 32 channels and marches down it once:
//
//   layout   levels are staged as [B][C/32][H][W][32] fp32 ("CG32": one 128-byte line per cell and channel group), so a
//            strip row is ONE contiguous segment and a CTA that owns 32 channels still moves whole lines;
//   prepass  (3 small launches) bins the RoIs by (level, image, strip, y-part), orders every bin by window top, and writes
//            one record per RoI -- x tap tables and dense per-row y weights in the reference's fp32 op order ("staging of
//            sampling coordinates") -- at its sorted position; windows that cannot be staged go to a leftover list;
//   ring     the producer warp streams the strip's rows top to bottom into a shared-memory ring with one TMA bulk copy per
//            row (cp.async.bulk, full/empty mbarriers); the ring runs on across unit boundaries, so the next unit's first
//            rows land while the teams finish the current one;
//   teams    NT teams of 8*P threads (thread = 4 channels x output column) pop RoIs in window-top order, wait for the rows
//            [y0, y0+hh) they need, run the separable pooling out of shared memory (LDS.128 + packed FFMA2), transpose
//            their [32][P][P] result through a team tile and send it off with one asynchronous bulk store; a row is
//            released to the producer once every team has moved past it.
// Each level row is fetched by (strips x channel groups) CTAs once (+ the x halo, ~1.4x at W = 128) instead of once per
// overlapping RoI.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"
#include "roi_common.cuh"
#include "roi_strip.cuh"

namespace {

constexpr int kCG = 32;                               // channels per group: 32 floats = one 128-byte line per cell
constexpr int kBW = 60;                               // cells per staged row (ring row = 7680 B)
constexpr int kXH = 18;                               // widest window staged on a level that needs several strips
constexpr int kCore = kBW - kXH + 1;                  // 43: strip s owns window origins [s*kCore, (s+1)*kCore)
constexpr int kRMax = 18;                             // tallest staged window
constexpr int kNU = 4;                                // unit descriptors in flight
constexpr int kRowFloats = kBW * kCG;
constexpr int kSmemMax = 232448;                      // 227 KB opt-in shared memory per CTA

template <int P> struct StripRec;
template <int P>
struct StripCfg {
    static constexpr int NW = 15;                     // consumer warps (+ the producer warp = 512 threads, 128 registers each)
    static constexpr int NTHREADS = 32 * (NW + 1);
    static constexpr int REC_BYTES = 96 + 2 * P * kMaxTap * 4;
    // ring rows: whatever the records, unit slots and barriers leave of the shared memory
    static constexpr int NR_RAW = (kSmemMax - NW * 2 * REC_BYTES - 1024 - 16 * NW) / (kRowFloats * 4 + 16);
    static constexpr int NR = NR_RAW > 30 ? 30 : NR_RAW;
    static_assert(NR >= kRMax + 2, "the ring must hold the tallest window plus some slack");
};

// One record per staged RoI, written by the prepass at the RoI's sorted position and fetched with one bulk copy:
// the per-bin tap tables of both axes in the reference's fp32 op order.  The consumers work BIN-MAJOR and y-first:
// for output row i they sweep only the ny[i] window rows that bin reaches (u[j] += wy * V[row][xs + j]) and then close
// the bin with the x taps (out = sum_j wx[j] * u[j]).  Everything is statically indexed and both passes are sparse:
// 2*NX*(ny + 1) packed FMAs per bin instead of the row-major form's 2*NX + 2*P per row, which spent two thirds of its
// FMAs on zero weights (profiles/r02_roialign_strip.md).
template <int P>
struct __align__(16) StripRec {
    int k, y0, hh, x0;
    unsigned char xs[16];   // per output column: first x tap - x0
    signed char nx[16];     // per output column: x taps (0: no valid sample)
    unsigned char ys[16];   // per output row: first y tap - y0
    signed char ny[16];     // per output row: y taps (0: no valid sample)
    int pad[4];
    float wx[P][kMaxTap];
    float wy[P][kMaxTap];   // already divided by the sample count
};
static_assert(sizeof(StripRec<7>) == StripCfg<7>::REC_BYTES && sizeof(StripRec<14>) == StripCfg<14>::REC_BYTES, "record size");

struct StripLevel {
    const float *data;
    int H, W;
    float scale;
    int nstrips, nyp, ypart_rows;
    int key0, bin0;
};

struct StripUnit {   // one per non-empty bin; the unit proper is (bin, channel group)
    int level, b, X0, BW, Y0, Y1, item_begin, item_count;
};

struct StripArgs {
    StripLevel lv[NUHTC_MAX_LEVELS];
    int L, B, C, K;
    int sr, aligned, mode;
    float finest;
    int nkeys, nbins;
    int dbg;           // measurement switches (NUHTC_RA_DBG): 1 skip the pooling arithmetic + stores, 2 skip the stores only
    const float *rois;
    float *out;
    const float *bias;
    // workspace
    int *item_key;     // [K] sort key of the item, -1: leftover
    int *hist;         // [nkeys + 1] -> exclusive offsets
    int *cursor;       // [nkeys]
    int *bin_ymax;     // [nbins]
    StripUnit *units;  // [nbins]
    int *counters;     // [0] nunits (bins x channel groups), [1] unit cursor, [2] leftover count
    int *leftover;     // [K]
    void *records;     // [K] StripRec<P>
};

// ---------------------------------------------------------------------------------------------
// prepass
// ---------------------------------------------------------------------------------------------
struct ItemGeom {
    int level, b;
    int x0, ww, y0, hh;   // window
    bool lit;             // a bin exceeds the tap table
    bool empty;           // no valid sample at all
};

// A warp evaluates the tap tables of one RoI: lanes [0, P) the x bins, lanes [16, 16 + P) the y bins.
// Returns the lane's table (w, first, n) and the window geometry (all lanes).
template <int P>
__device__ __forceinline__ ItemGeom item_tables(const StripArgs &a, int k, int lane, float (&w)[kMaxTap], int &first, int &n,
                                                float &count) {
    const float *roi = a.rois + (size_t)k * 5;
    ItemGeom g;
    g.level = a.mode == NUHTC_ROI_ROUTE ? route_level(roi, a.L, a.finest) : 0;
    const StripLevel &lv = a.lv[g.level];
    const RoiGeom rg = roi_geom(roi, lv.scale, P, P, a.sr, a.aligned);
    g.b = rg.b;
    count = rg.count;
    first = 0;
    n = 0;
    const bool xb = lane < P, yb = lane >= 16 && lane < 16 + P;
    if (xb) build_axis_taps(rg.start_w, rg.bin_w, rg.gw, lane, lv.W, 1.0f, w, &first, &n);
    else if (yb) build_axis_taps(rg.start_h, rg.bin_h, rg.gh, lane - 16, lv.H, rg.count, w, &first, &n);
    const bool has = (xb || yb) && n > 0, bad = (xb || yb) && n < 0;
    int lo = has ? first : (1 << 30), hi = has ? first + n : -1;
#pragma unroll
    for (int o = 8; o; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    const int x0 = __shfl_sync(0xffffffffu, lo, 0), x1 = __shfl_sync(0xffffffffu, hi, 0);
    const int y0 = __shfl_sync(0xffffffffu, lo, 16), y1 = __shfl_sync(0xffffffffu, hi, 16);
    g.lit = __any_sync(0xffffffffu, bad);
    g.empty = x1 <= x0 || y1 <= y0;
    g.x0 = g.empty ? 0 : x0;
    g.ww = g.empty ? 0 : x1 - x0;
    g.y0 = g.empty ? 0 : y0;
    g.hh = g.empty ? 0 : y1 - y0;
    return g;
}

__device__ __forceinline__ bool item_staged(const StripArgs &a, const ItemGeom &g) {
    if (g.lit || g.empty || g.hh > kRMax) return false;
    if (g.b < 0 || g.b >= a.B) return false;
    const StripLevel &lv = a.lv[g.level];
    return lv.nstrips == 1 ? lv.W <= kBW : g.ww <= kXH;
}

// K1: sort key of every RoI (warp per RoI), histogram of the keys, lowest window bottom of every bin
template <int P>
__global__ void __launch_bounds__(256) strip_keys_kernel(StripArgs a) {
    const int lane = threadIdx.x & 31;
    const int k = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    if (k >= a.K) return;
    float w[kMaxTap];
    int first, n;
    float count;
    const ItemGeom g = item_tables<P>(a, k, lane, w, first, n, count);
    if (lane != 0) return;
    int key = -1;
    if (item_staged(a, g)) {
        const StripLevel &lv = a.lv[g.level];
        const int s = lv.nstrips == 1 ? 0 : min(g.x0 / kCore, lv.nstrips - 1);
        // the last strip of a level is cut at the map edge: its windows reach at most W, which fits by construction
        key = lv.key0 + (g.b * lv.nstrips + s) * lv.H + g.y0;
        const int bin = lv.bin0 + (g.b * lv.nstrips + s) * lv.nyp + g.y0 / lv.ypart_rows;
        atomicAdd(a.hist + key, 1);
        atomicMax(a.bin_ymax + bin, g.y0 + g.hh);
    }
    a.item_key[k] = key;
}

// K2: one CTA.  Exclusive scan of the key histogram (-> sorted positions), then one descriptor per non-empty bin.
__global__ void __launch_bounds__(1024) strip_scan_kernel(StripArgs a) {
    __shared__ int s_warp[32];
    __shared__ int s_carry, s_units;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        s_carry = 0;
        s_units = 0;
    }
    __syncthreads();
    for (int base = 0; base < a.nkeys; base += 1024) {
        const int i = base + tid;
        const int v = i < a.nkeys ? a.hist[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int t = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += y;
            }
            s_warp[lane] = t;
        }
        __syncthreads();
        const int excl = s_carry + (wid ? s_warp[wid - 1] : 0) + x - v;
        if (i < a.nkeys) {
            a.hist[i] = excl;
            a.cursor[i] = excl;
        }
        __syncthreads();
        if (tid == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) a.hist[a.nkeys] = s_carry;
    __syncthreads();
    // bins -> units, in (level, image, strip, y-part) order so that neighbouring units share rows and records in L2
    for (int base = 0; base < a.nbins; base += 1024) {
        const int bin = base + tid;
        int cnt = 0, begin = 0, Y0 = 0, level = 0, b = 0, s = 0;
        if (bin < a.nbins) {
            level = a.L - 1;
            while (level > 0 && bin < a.lv[level].bin0) --level;
            const StripLevel &lv = a.lv[level];
            const int r = bin - lv.bin0;
            const int yp = r % lv.nyp, bs = r / lv.nyp;
            s = bs % lv.nstrips;
            b = bs / lv.nstrips;
            const int ka = lv.key0 + bs * lv.H + yp * lv.ypart_rows;
            const int kb = lv.key0 + bs * lv.H + min(lv.H, (yp + 1) * lv.ypart_rows);
            begin = a.hist[ka];
            cnt = a.hist[kb] - begin;
            if (cnt > 0) {
                int kk = ka;
                while (a.hist[kk + 1] == begin) ++kk;   // first non-empty key of the bin = smallest window top
                Y0 = kk - (lv.key0 + bs * lv.H);
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, cnt > 0);
        if (lane == 0) s_warp[wid] = __popc(m);
        __syncthreads();
        if (wid == 0) {
            int t = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += y;
            }
            s_warp[lane] = t;
        }
        __syncthreads();
        if (cnt > 0) {
            const int slot = s_units + (wid ? s_warp[wid - 1] : 0) + __popc(m & ((1u << lane) - 1));
            const StripLevel &lv = a.lv[level];
            StripUnit u;
            u.level = level;
            u.b = b;
            u.X0 = lv.nstrips == 1 ? 0 : s * kCore;
            u.BW = min(kBW, lv.W - u.X0);
            u.Y0 = Y0;
            u.Y1 = a.bin_ymax[bin];
            u.item_begin = begin;
            u.item_count = cnt;
            a.units[slot] = u;
        }
        __syncthreads();
        if (tid == 0) s_units += s_warp[31];
        __syncthreads();
    }
    if (tid == 0) a.counters[0] = s_units;
}

// K3: records at their sorted positions (warp per RoI); RoIs that cannot be staged go to the leftover list
template <int P>
__global__ void __launch_bounds__(256) strip_records_kernel(StripArgs a) {
    // (record layout: StripRec<P>)
    const int lane = threadIdx.x & 31;
    const int k = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    if (k >= a.K) return;
    const int key = a.item_key[k];
    if (key < 0) {
        if (lane == 0) a.leftover[atomicAdd(a.counters + 2, 1)] = k;
        return;
    }
    float w[kMaxTap];
    int first, n;
    float count;
    const ItemGeom g = item_tables<P>(a, k, lane, w, first, n, count);
    int pos = 0;
    if (lane == 0) pos = atomicAdd(a.cursor + key, 1);
    pos = __shfl_sync(0xffffffffu, pos, 0);
    StripRec<P> *rec = reinterpret_cast<StripRec<P> *>(a.records) + pos;
    if (lane == 0) {
        rec->k = k;
        rec->y0 = g.y0;
        rec->hh = g.hh;
        rec->x0 = g.x0;
    }
    if (lane < 16) {
        rec->xs[lane] = (unsigned char)((lane < P && n > 0) ? first - g.x0 : 0);
        rec->nx[lane] = (signed char)(lane < P ? n : 0);
    }
    if (lane < P) {
        *reinterpret_cast<float4 *>(&rec->wx[lane][0]) = make_float4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<float4 *>(&rec->wx[lane][4]) = make_float4(w[4], w[5], w[6], w[7]);
    }
    if (lane >= 16 && lane < 32) {
        const int p = lane - 16;
        rec->ys[p] = (unsigned char)((p < P && n > 0) ? first - g.y0 : 0);
        rec->ny[p] = (signed char)(p < P ? n : 0);
        if (p < P) {
            *reinterpret_cast<float4 *>(&rec->wy[p][0]) = make_float4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<float4 *>(&rec->wy[p][4]) = make_float4(w[4], w[5], w[6], w[7]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// strip kernel
// ---------------------------------------------------------------------------------------------
struct __align__(16) UnitSlot {
    StripUnit u;
    int cg;        // channel group
    int row_base;  // position of row Y0 in the ring's running row count
    int next_item; // shared pop counter of the consumer warps
    int valid;     // 0: no more units
};

// Consumer = ONE WARP per work item, no barrier between warps anywhere: lane = (channel-quad column q < 4, output column
// pw < 7), 28 active lanes; an item is (RoI, group of 7 output columns) -- P = 14 has two column groups per RoI.  A lane
// owns two float4 slices of the 32 channels (channels 4q.. and 16+4q..).  LDS.128 is served a quarter-warp at a time (8
// lanes = 2 output columns x 4 quads): lanes of an odd column load their HIGH slice first, so the two columns of a phase
// cover banks 0-15 and 16-31 and every tap load is conflict-free although the two columns read different cells.
// The bin loop is a real loop (tiny code, ~60 registers): each bin's 8 results per lane go straight to global memory
// (7 lanes of one channel write 28 contiguous bytes; L2 merges the partial sectors before they reach DRAM).
template <int P>
__global__ void __launch_bounds__(StripCfg<P>::NTHREADS, 1) roi_align_strip_kernel(const __grid_constant__ StripArgs a) {
    using Cfg = StripCfg<P>;
    using Rec = StripRec<P>;
    constexpr int PP = P * P, NW = Cfg::NW, NR = Cfg::NR, GROUPS = P / 7;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *s_ring = reinterpret_cast<float *>(smem_raw);                              // [NR][kBW][32]
    Rec *s_rec = reinterpret_cast<Rec *>(s_ring + (size_t)NR * kRowFloats);           // [NW][2]
    UnitSlot *s_unit = reinterpret_cast<UnitSlot *>(s_rec + NW * 2);                  // [kNU]
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_unit + kNU);                     // full[NR], empty[NR], ufull[kNU], uempty[kNU], rfull[NW*2]
    const uint32_t full0 = smem_u32(s_bar), empty0 = full0 + 8 * NR, ufull0 = empty0 + 8 * NR, uempty0 = ufull0 + 8 * kNU,
                   rfull0 = uempty0 + 8 * kNU;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < NR; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, NW);
        }
        for (int s = 0; s < kNU; ++s) {
            mbar_init(ufull0 + 8 * s, 1);
            mbar_init(uempty0 + 8 * s, NW);
        }
        for (int s = 0; s < NW * 2; ++s) mbar_init(rfull0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int ncg = a.C / kCG;

    if (tid < 32) {
        // =========================== producer warp: unit scheduler + row streamer ===========================
        if (tid != 0) return;
        const int nunits = a.counters[0] * ncg;
        unsigned g = 0;   // running row count of this CTA's ring
        for (unsigned uc = 0;; ++uc) {
            const unsigned us = uc % kNU;
            mbar_wait(uempty0 + 8 * us, ((uc / kNU) & 1) ^ 1);
            const int u = atomicAdd(a.counters + 1, 1);
            UnitSlot &S = s_unit[us];
            if (u >= nunits) {
                S.valid = 0;
                mbar_arrive(ufull0 + 8 * us);
                break;
            }
            const StripUnit d = a.units[u / ncg];
            S.u = d;
            S.cg = u % ncg;
            S.row_base = (int)g;
            S.next_item = 0;
            S.valid = 1;
            mbar_arrive(ufull0 + 8 * us);   // release: the slot's contents are visible to the warps that acquire the phase
            const StripLevel &lv = a.lv[d.level];
            const float *src = lv.data + ((((size_t)d.b * ncg + S.cg) * lv.H + d.Y0) * lv.W + d.X0) * kCG;
            const uint32_t bytes = (uint32_t)d.BW * kCG * 4;
            const size_t rstride = (size_t)lv.W * kCG;
            for (int y = d.Y0; y < d.Y1; ++y, ++g, src += rstride) {
                const unsigned slot = g % NR;
                mbar_wait(empty0 + 8 * slot, ((g / NR) & 1) ^ 1);
                mbar_arrive_expect_tx(full0 + 8 * slot, bytes);
                tma_bulk_g2s(smem_u32(s_ring + (size_t)slot * kRowFloats), src, bytes, full0 + 8 * slot);
            }
        }
        return;
    }

    // =========================== consumer warps ===========================
    const int cw = (tid >> 5) - 1, lane = tid & 31;
    const bool worker = lane < 28;
    const int q = lane & 3, pwl = worker ? lane >> 2 : 0;   // output column inside the item's column group
    const bool odd = (pwl & 1) != 0;                        // this lane keeps its HIGH slice in register set 0
    Rec *recs = s_rec + cw * 2;
    const uint32_t rf0 = rfull0 + 8 * (cw * 2);
    unsigned rec_uses[2] = {0u, 0u};
    // Rows [from, to) of the current unit are behind this warp: one arrival per row on its slot's `empty` barrier.  A
    // slot's barrier only moves on to its next use once ALL warps have arrived, and the producer refills the slot only
    // then -- so before arriving for a row the lane first sees that row's `full` phase complete: that proves the slot's
    // previous use has been released by everybody and this arrival is counted for the right phase (a warp that skips more
    // than NR rows would otherwise arrive twice in one phase; and a parity wait can tell the current phase from the next
    // one but not from the one after, so the warp must have seen row g - NR land before it may wait for row g).
    // 16 rows at a time: distinct slots per round.
    auto release_rows = [&](int from, int to, int row_base, int Y0) {
        for (int base = from; base < to; base += 16) {
            const int r = base + (lane & 15);
            if (lane < 16 && r < to) {
                const unsigned gg = (unsigned)(row_base + r - Y0);
                mbar_wait(full0 + 8 * (gg % NR), (gg / NR) & 1);
                mbar_arrive(empty0 + 8 * (gg % NR));
            }
            __syncwarp();
        }
    };

    for (unsigned uc = 0;; ++uc) {
        const unsigned us = uc % kNU;
        mbar_wait(ufull0 + 8 * us, (uc / kNU) & 1);
        UnitSlot &S = s_unit[us];
        if (!S.valid) break;
        const StripUnit d = S.u;
        const int cg = S.cg, row_base = S.row_base;
        const int nitems = d.item_count * GROUPS;
        const Rec *gtab = reinterpret_cast<const Rec *>(a.records) + d.item_begin;
        int passed = d.Y0;   // first row of this unit the warp has not released yet
        auto fetch = [&](int item, int buf) {   // lane 0 only: the whole fixed-size record with one bulk copy
            mbar_arrive_expect_tx(rf0 + 8 * buf, (uint32_t)sizeof(Rec));
            tma_bulk_g2s(smem_u32(&recs[buf]), gtab + item / GROUPS, (uint32_t)sizeof(Rec), rf0 + 8 * buf);
        };
        int cur = 0, curbuf = 0;
        if (lane == 0) {
            cur = atomicAdd(&S.next_item, 1);
            if (cur < nitems) fetch(cur, 0);
        }
        cur = __shfl_sync(0xffffffffu, cur, 0);
        while (cur < nitems) {
            // ---- next item: pop + prefetch its record into the other buffer
            int nxt = 0;
            if (lane == 0) {
                nxt = atomicAdd(&S.next_item, 1);
                if (nxt < nitems) fetch(nxt, curbuf ^ 1);
            }
            nxt = __shfl_sync(0xffffffffu, nxt, 0);
            // ---- this item's record
            mbar_wait(rf0 + 8 * curbuf, rec_uses[curbuf] & 1);
            ++rec_uses[curbuf];
            const Rec &R = recs[curbuf];
            const int k = R.k, y0 = R.y0, hh = R.hh;
            const int pw = pwl + 7 * (cur % GROUPS);
            // ---- release the rows above this window, then wait for the window's rows
            release_rows(passed, y0, row_base, d.Y0);
            passed = max(passed, y0);
            {
                const unsigned gg = (unsigned)(row_base + y0 - d.Y0) + (lane < hh ? lane : 0);
                mbar_wait(full0 + 8 * (gg % NR), (gg / NR) & 1);
                __syncwarp();
            }
            // ---- separable pooling out of the ring, bin-major and y-first (see StripRec)
            if (worker && !(a.dbg & 1)) {
                const int nx = R.nx[pw];
                const int xoff = R.x0 - d.X0 + (int)R.xs[pw];
                const float *wxp = R.wx[pw];
                const unsigned slot = (unsigned)(row_base + y0 - d.Y0) % NR;
                // register set 0 reads this float offset inside a cell, set 1 the other half of the 128-byte line
                const float *colp = s_ring + (size_t)xoff * kCG + (odd ? 16 : 0) + 4 * q;
                const int d1 = odd ? -16 : 16;
                // channel of set s, element 0: 16 * (s ^ odd) + 4q
                float *out0 = a.out + ((size_t)k * a.C + (size_t)cg * kCG + (odd ? 16 : 0) + 4 * q) * PP + pw;
                const int o1 = d1 * PP;
                float4 bz0 = make_float4(0.f, 0.f, 0.f, 0.f), bz1 = bz0;
                if (a.bias) {
                    const float *bp = a.bias + (size_t)k * a.C + cg * kCG + (odd ? 16 : 0) + 4 * q;
                    bz0 = ldg_f4(bp);
                    bz1 = ldg_f4(bp + d1);
                }
                auto bins = [&](auto NXC, int j0, bool first_pass) {   // x taps j0 .. j0 + NX - 1 of this lane's column
                    constexpr int NX = decltype(NXC)::value;
                    constexpr int NXR = NX > 0 ? NX : 1;
                    float wx[NXR];
#pragma unroll
                    for (int j = 0; j < NX; ++j) wx[j] = wxp[j0 + j];
#pragma unroll 1
                    for (int i = 0; i < P; ++i) {
                        const int nyi = R.ny[i];                      // uniform over the warp
                        unsigned sl = slot + R.ys[i];
                        if (sl >= NR) sl -= NR;
                        float2 u[2][NXR][2];
#pragma unroll
                        for (int s = 0; s < 2; ++s)
#pragma unroll
                            for (int j = 0; j < NXR; ++j) u[s][j][0] = u[s][j][1] = make_float2(0.f, 0.f);
#pragma unroll 1
                        for (int jj = 0; jj < nyi; ++jj) {
                            const float w = R.wy[i][jj];
                            const float *row = colp + (size_t)sl * kRowFloats + j0 * kCG;
                            float4 v[2][NXR];
#pragma unroll
                            for (int s = 0; s < 2; ++s)
#pragma unroll
                                for (int j = 0; j < NX; ++j) v[s][j] = *reinterpret_cast<const float4 *>(row + j * kCG + s * d1);
#pragma unroll
                            for (int s = 0; s < 2; ++s)
#pragma unroll
                                for (int j = 0; j < NX; ++j) {
                                    u[s][j][0] = ffma2(w, make_float2(v[s][j].x, v[s][j].y), u[s][j][0]);
                                    u[s][j][1] = ffma2(w, make_float2(v[s][j].z, v[s][j].w), u[s][j][1]);
                                }
                            if (++sl == NR) sl = 0;
                        }
                        float2 r[2][2];
#pragma unroll
                        for (int s = 0; s < 2; ++s) {
                            r[s][0] = r[s][1] = make_float2(0.f, 0.f);
#pragma unroll
                            for (int j = 0; j < NX; ++j) {
                                r[s][0] = ffma2(wx[j], u[s][j][0], r[s][0]);
                                r[s][1] = ffma2(wx[j], u[s][j][1], r[s][1]);
                            }
                        }
                        float *op = out0 + i * P;
                        if (a.dbg & 2) {
                            if (r[0][0].x + r[0][0].y + r[0][1].x + r[0][1].y + r[1][0].x + r[1][0].y + r[1][1].x + r[1][1].y == 12345.678f) op[0] = 1.f;
                        } else if (first_pass) {
                            __stcs(op, r[0][0].x + bz0.x);
                            __stcs(op + PP, r[0][0].y + bz0.y);
                            __stcs(op + 2 * PP, r[0][1].x + bz0.z);
                            __stcs(op + 3 * PP, r[0][1].y + bz0.w);
                            __stcs(op + o1, r[1][0].x + bz1.x);
                            __stcs(op + o1 + PP, r[1][0].y + bz1.y);
                            __stcs(op + o1 + 2 * PP, r[1][1].x + bz1.z);
                            __stcs(op + o1 + 3 * PP, r[1][1].y + bz1.w);
                        } else {   // second pass of a bin wider than 4 taps: add to what the first pass stored
                            op[0] += r[0][0].x;
                            op[PP] += r[0][0].y;
                            op[2 * PP] += r[0][1].x;
                            op[3 * PP] += r[0][1].y;
                            op[o1] += r[1][0].x;
                            op[o1 + PP] += r[1][0].y;
                            op[o1 + 2 * PP] += r[1][1].x;
                            op[o1 + 3 * PP] += r[1][1].y;
                        }
                    }
                };
                // exact tap counts only (no zero-weight padding: a padded tap would multiply whatever sits past the bin by
                // 0, and 0 * inf is NaN); nx == 0 (no valid sample in this column) stores the bias / zeros; bins wider than
                // 4 taps (large RoIs) take two passes of up to 4 taps, the second one read-modify-writes the lane's own
                // stores (same thread, same addresses: program order)
                auto taps = [&](int n, int j0, bool first_pass) {
                    switch (n) {
                        case 0: bins(std::integral_constant<int, 0>{}, j0, first_pass); break;
                        case 1: bins(std::integral_constant<int, 1>{}, j0, first_pass); break;
                        case 2: bins(std::integral_constant<int, 2>{}, j0, first_pass); break;
                        case 3: bins(std::integral_constant<int, 3>{}, j0, first_pass); break;
                        default: bins(std::integral_constant<int, 4>{}, j0, first_pass); break;
                    }
                };
                taps(nx < 4 ? (nx < 0 ? 0 : nx) : 4, 0, true);
                if (nx > 4) taps(nx - 4, 4, false);
            }
            __syncwarp();   // every lane is done with the ring rows and the record before the warp moves on
            cur = nxt;
            curbuf ^= 1;
        }
        // ---- unit done for this warp: release its remaining rows and the unit slot
        release_rows(passed, d.Y1, row_base, d.Y0);
        if (lane == 0) mbar_arrive(uempty0 + 8 * us);
    }
}

template <int P>
size_t strip_smem_bytes() {
    using Cfg = StripCfg<P>;
    return (size_t)Cfg::NR * kRowFloats * 4 + (size_t)Cfg::NW * 2 * sizeof(StripRec<P>) + kNU * sizeof(UnitSlot) +
           8 * (2 * Cfg::NR + 2 * kNU + 2 * Cfg::NW) + 128;
}

// ---------------------------------------------------------------------------------------------
// layout staging: NCHW / channels_last -> [B][C/32][H][W][32]
// ---------------------------------------------------------------------------------------------
// 32 channels x 64 pixels per CTA: reads 64 contiguous floats per channel, writes one contiguous 8 KB chunk
__global__ void __launch_bounds__(256) nchw_to_cg32_kernel(const float *__restrict__ in, float *__restrict__ out, int C, int HW) {
    __shared__ float tile[32][65];
    const int b = blockIdx.z, cg = blockIdx.y, hw0 = blockIdx.x * 64;
    const int t = threadIdx.x;
    const float *src = in + ((size_t)b * C + cg * 32) * HW;
    {   // thread = (channel r + 16k, pixel quad q): 128-bit loads
        const int r = t >> 4, q = (t & 15) * 4;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int c = r + 16 * k, hw = hw0 + q;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (hw + 3 < HW) v = ldg_f4(src + (size_t)c * HW + hw);
            else
                for (int e = 0; e < 4; ++e)
                    if (hw + e < HW) (&v.x)[e] = __ldg(src + (size_t)c * HW + hw + e);
            tile[c][q] = v.x; tile[c][q + 1] = v.y; tile[c][q + 2] = v.z; tile[c][q + 3] = v.w;
        }
    }
    __syncthreads();
    float *dst = out + (((size_t)b * (C / 32) + cg) * HW + hw0) * 32;
    {   // thread = (pixel p + 32k, channel quad q): 128-bit stores, a warp writes 4 pixels x 128 B
        const int p = t >> 3, q = (t & 7) * 4;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int px = p + 32 * k;
            if (hw0 + px < HW)
                *reinterpret_cast<float4 *>(dst + (size_t)px * 32 + q) = make_float4(tile[q][px], tile[q + 1][px], tile[q + 2][px], tile[q + 3][px]);
        }
    }
}

// channels_last source: [B][HW][C] -> [B][C/32][HW][32], a pure 128-byte-line permutation
__global__ void __launch_bounds__(256) nhwc_to_cg32_kernel(const float *__restrict__ in, float *__restrict__ out, int C, long HW, long total4) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total4; i += (long)gridDim.x * blockDim.x) {
        // i indexes float4s of the OUTPUT: ((b*CG + cg)*HW + hw)*8 + q
        const int q = (int)(i & 7);
        const long r = i >> 3;
        const long hw = r % HW;
        const long bc = r / HW;
        const int cg = (int)(bc % (C / 32));
        const long b = bc / (C / 32);
        const float4 v = ldg_f4(in + ((size_t)(b * HW + hw)) * C + cg * 32 + q * 4);
        *reinterpret_cast<float4 *>(out + (size_t)i * 4) = v;
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int strip_nyp_env() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("NUHTC_RA_YROWS");
        v = e ? atoi(e) : 32;
        if (v < 8) v = 8;
    }
    return v;
}

static void strip_plan(const int *H, const int *W, int L, int B, StripLevel *lv, int *nkeys, int *nbins) {
    int key = 0, bin = 0;
    for (int l = 0; l < L; ++l) {
        lv[l].H = H[l];
        lv[l].W = W[l];
        lv[l].nstrips = W[l] <= kBW ? 1 : (W[l] + kCore - 1) / kCore;
        lv[l].ypart_rows = strip_nyp_env();
        lv[l].nyp = (H[l] + lv[l].ypart_rows - 1) / lv[l].ypart_rows;
        lv[l].key0 = key;
        lv[l].bin0 = bin;
        key += B * lv[l].nstrips * H[l];
        bin += B * lv[l].nstrips * lv[l].nyp;
    }
    *nkeys = key;
    *nbins = bin;
}

struct StripWs {
    size_t item_key, hist, cursor, bin_ymax, units, counters, leftover, records, total;
};
static StripWs strip_ws_layout(int K, int nkeys, int nbins, int P) {
    StripWs w;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        const size_t at = o;
        o = align_up(o + bytes, 256);
        return at;
    };
    w.counters = take(64);
    w.hist = take(sizeof(int) * ((size_t)nkeys + 1));
    w.bin_ymax = take(sizeof(int) * (size_t)nbins);
    // everything up to here is zeroed by one memset per call
    w.cursor = take(sizeof(int) * (size_t)nkeys);
    w.item_key = take(sizeof(int) * (size_t)K);
    w.units = take(sizeof(StripUnit) * (size_t)nbins);
    w.leftover = take(sizeof(int) * (size_t)K);
    w.records = take((P == 7 ? sizeof(StripRec<7>) : sizeof(StripRec<14>)) * (size_t)K);
    w.total = o;
    return w;
}

size_t roi_strip_workspace_bytes(const int *H, const int *W, int L, int B, int K, int P) {
    StripLevel lv[NUHTC_MAX_LEVELS];
    int nkeys, nbins;
    strip_plan(H, W, L, B, lv, &nkeys, &nbins);
    return strip_ws_layout(K, nkeys, nbins, P).total;
}

bool roi_strip_supported(int C, int PH, int PW, int mode, int L) {
    return PH == PW && (PH == 7 || PH == 14) && C % 32 == 0 && C >= 32 && (mode == NUHTC_ROI_ROUTE || L == 1);
}

template <int P>
static int strip_launch(StripArgs &a, const RoiLevels &lv, cudaStream_t st, const StripWs &w, char *ws) {
    using Cfg = StripCfg<P>;
    static bool attr_done[kNuhtcMaxDevices] = {false};
    const int dev = nuhtc_device();
    const size_t smem = strip_smem_bytes<P>();
    if (!attr_done[dev]) {
        NUHTC_CUDA(cudaFuncSetAttribute(roi_align_strip_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done[dev] = true;
    }
    NUHTC_CUDA(cudaMemsetAsync(ws + w.counters, 0, w.cursor - w.counters, st));
    const int warps_per_block = 8;
    const unsigned gblocks = (unsigned)((a.K + warps_per_block - 1) / warps_per_block);
    strip_keys_kernel<P><<<gblocks, 256, 0, st>>>(a);
    NUHTC_LAUNCH_CHECK();
    strip_scan_kernel<<<1, 1024, 0, st>>>(a);
    NUHTC_LAUNCH_CHECK();
    strip_records_kernel<P><<<gblocks, 256, 0, st>>>(a);
    NUHTC_LAUNCH_CHECK();
    roi_align_strip_kernel<P><<<nuhtc_sm_count(), Cfg::NTHREADS, smem, st>>>(a);
    NUHTC_LAUNCH_CHECK();
    (void)lv;
    return NUHTC_OK;
}

// Fills `a` and runs prepass + strip kernel; the caller then runs its per-RoI kernel over the leftover list
// (*leftover, *leftover_count are device pointers into the workspace).
int roi_strip_forward(const RoiLevels &lv, int B, int C, const float *rois, int K, int P, int sr, int aligned, int mode,
                      float finest, float *out, const float *bias, void *ws_, size_t ws_bytes, cudaStream_t st,
                      const int **leftover, const int **leftover_count) {
    StripArgs a;
    memset(&a, 0, sizeof a);
    int H[NUHTC_MAX_LEVELS], W[NUHTC_MAX_LEVELS];
    for (int l = 0; l < lv.L; ++l) {
        H[l] = lv.H[l];
        W[l] = lv.W[l];
    }
    strip_plan(H, W, lv.L, B, a.lv, &a.nkeys, &a.nbins);
    for (int l = 0; l < lv.L; ++l) {
        a.lv[l].data = lv.data[l];
        a.lv[l].scale = lv.scale[l];
    }
    const StripWs w = strip_ws_layout(K, a.nkeys, a.nbins, P);
    if (ws_bytes < w.total || !ws_) {
        nuhtc_set_error("roi_align: workspace %zu < %zu bytes", ws_bytes, w.total);
        return NUHTC_EWORKSPACE;
    }
    char *ws = (char *)ws_;
    static const int dbg = getenv("NUHTC_RA_DBG") ? atoi(getenv("NUHTC_RA_DBG")) : 0;
    a.dbg = dbg;
    a.L = lv.L;
    a.B = B;
    a.C = C;
    a.K = K;
    a.sr = sr;
    a.aligned = aligned;
    a.mode = mode;
    a.finest = finest;
    a.rois = rois;
    a.out = out;
    a.bias = bias;
    a.item_key = (int *)(ws + w.item_key);
    a.hist = (int *)(ws + w.hist);
    a.cursor = (int *)(ws + w.cursor);
    a.bin_ymax = (int *)(ws + w.bin_ymax);
    a.units = (StripUnit *)(ws + w.units);
    a.counters = (int *)(ws + w.counters);
    a.leftover = (int *)(ws + w.leftover);
    a.records = ws + w.records;
    *leftover = a.leftover;
    *leftover_count = a.counters + 2;
    return P == 7 ? strip_launch<7>(a, lv, st, w, ws) : strip_launch<14>(a, lv, st, w, ws);
}

int roi_to_cg32(const float *in, float *out, int B, int C, int H, int W, int channels_last, cudaStream_t st) {
    const int HW = H * W;
    if (channels_last) {
        const long total4 = (long)B * C * HW / 4;
        long blocks = (total4 + 255) / 256;
        const long cap = (long)nuhtc_sm_count() * 16;
        if (blocks > cap) blocks = cap;
        nhwc_to_cg32_kernel<<<(unsigned)blocks, 256, 0, st>>>(in, out, C, HW, total4);
    } else {
        dim3 grid((HW + 63) / 64, C / 32, B);
        nchw_to_cg32_kernel<<<grid, 256, 0, st>>>(in, out, C, HW);
    }
    NUHTC_LAUNCH_CHECK();
    return NUHTC_OK;
}
