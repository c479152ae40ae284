// Cross-tile nucleus merge for sm_100a: greedy polygon-IoU suppression in score order.
//
// Replaces merge_overlap() of /root/reference/tools/nuclei_merge.py:62-174:
//   sort by score (desc) -> STRtree envelope candidates (:107,118) -> shapely
//   intersection().area / IoU (:132-133) -> greedy suppression in score order (:114-150) ->
//   kept rows re-indexed in score order (= nuclei_id, :167-174,201).
//
// B200 formulation (compile this file with -fmad=false: the double arithmetic must round exactly
// like the CPU oracle's):
//   stats   : per polygon envelope, signed area, global extent (one thread per polygon)
//   rank    : one stable radix sort of the scores (ties keep the lower index)
//   bin     : uniform grid keyed on the envelope's lower-left corner, cell >= largest envelope, so
//             all envelope-intersecting partners sit in the 3x3 neighbourhood (replaces the STRtree)
//   pairs   : count / scan / fill the candidate list (higher score -> lower score)
//   iou     : one warp per candidate pair; area(P∩Q) by the signed-trapezoid identity
//             1_P = o_P * sum_e -sgn(dx_e) 1_{T(e)}: lanes own edges of P, walk all edges of Q,
//             partial sums are combined by a fixed xor-butterfly (the oracle adds in the same order)
//   resolve : the greedy order is a DAG over "suppresses" edges; a nucleus is kept once every
//             higher-scoring suppressor is known suppressed, suppressed once one is known kept.
//             Rounds of that rule reach the same fixed point as the sequential loop.
//   emit    : kept nuclei compacted in rank order.
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

__device__ __forceinline__ unsigned long long dbl_desc_key(double s) {
    unsigned long long u = (unsigned long long)__double_as_longlong(s);
    u = (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
    return ~u;
}
__device__ __forceinline__ long long dbl_ordered(double d) {
    long long i = __double_as_longlong(d);
    return i >= 0 ? i : i ^ 0x7fffffffffffffffll;
}
static inline double ordered_dbl_host(long long i) {
    long long b = i >= 0 ? i : i ^ 0x7fffffffffffffffll;
    double d;
    memcpy(&d, &b, sizeof d);
    return d;
}

struct Extent { // device-side reduction cell: ordered-int encodings of doubles
    long long minx, miny, maxx, maxy, maxext;
};

__global__ void merge_init_kernel(Extent *ext, int64_t *counters, int32_t *status) {
    ext->minx = ext->miny = dbl_ordered(1e300);
    ext->maxx = ext->maxy = dbl_ordered(-1e300);
    ext->maxext = dbl_ordered(1.0);
    counters[0] = counters[1] = counters[2] = counters[3] = 0;
    *status = 0;
}

// twice the signed area, relative to the first vertex, ascending edge order (matches the oracle)
__device__ __forceinline__ double area2_of(const double *p, int V) {
    if (V < 3) return 0.0;
    const double ox = p[0], oy = p[1];
    double s = 0.0;
    for (int i = 0; i < V; ++i) {
        const int j = (i + 1 == V) ? 0 : i + 1;
        const double x0 = p[2 * i] - ox, y0 = p[2 * i + 1] - oy;
        const double x1 = p[2 * j] - ox, y1 = p[2 * j + 1] - oy;
        s += x0 * y1 - x1 * y0;
    }
    return s;
}

__global__ void __launch_bounds__(256) merge_stats_kernel(const double *__restrict__ xy, const int64_t *__restrict__ voff,
                                                          const double *__restrict__ score, int64_t N, double4 *__restrict__ env,
                                                          double *__restrict__ area2, unsigned long long *__restrict__ keys,
                                                          int32_t *__restrict__ vals, Extent *ext) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300, me = 1.0;
    if (i < N) {
        const double *p = xy + 2 * voff[i];
        const int V = (int)(voff[i + 1] - voff[i]);
        for (int v = 0; v < V; ++v) {
            const double x = p[2 * v], y = p[2 * v + 1];
            x0 = fmin(x0, x);
            x1 = fmax(x1, x);
            y0 = fmin(y0, y);
            y1 = fmax(y1, y);
        }
        if (V == 0) x0 = y0 = x1 = y1 = 0.0;
        env[i] = make_double4(x0, y0, x1, y1);
        area2[i] = area2_of(p, V);
        keys[i] = dbl_desc_key(score[i]);
        vals[i] = (int32_t)i;
        me = fmax(1.0, fmax(x1 - x0, y1 - y0));
    }
    // warp reduce, one atomic per warp
    for (int o = 16; o; o >>= 1) {
        x0 = fmin(x0, __shfl_xor_sync(0xffffffffu, x0, o));
        y0 = fmin(y0, __shfl_xor_sync(0xffffffffu, y0, o));
        x1 = fmax(x1, __shfl_xor_sync(0xffffffffu, x1, o));
        y1 = fmax(y1, __shfl_xor_sync(0xffffffffu, y1, o));
        me = fmax(me, __shfl_xor_sync(0xffffffffu, me, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&ext->minx, dbl_ordered(x0));
        atomicMin(&ext->miny, dbl_ordered(y0));
        atomicMax(&ext->maxx, dbl_ordered(x1));
        atomicMax(&ext->maxy, dbl_ordered(y1));
        atomicMax(&ext->maxext, dbl_ordered(me));
    }
}

__global__ void merge_rank_kernel(const int32_t *__restrict__ order, int64_t N, int32_t *__restrict__ rank) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r < N) rank[order[r]] = (int32_t)r;
}

struct Grid {
    double gx0, gy0, cell;
    int nx, ny;
};

__device__ __forceinline__ int cell_of(const Grid g, double x0, double y0, int &cx, int &cy) {
    cx = (int)floor((x0 - g.gx0) / g.cell);
    cy = (int)floor((y0 - g.gy0) / g.cell);
    cx = min(max(cx, 0), g.nx - 1);
    cy = min(max(cy, 0), g.ny - 1);
    return cy * g.nx + cx;
}

__global__ void merge_cell_kernel(const double4 *__restrict__ env, int64_t N, Grid g, uint32_t *__restrict__ cell,
                                  int32_t *__restrict__ idx) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    int cx, cy;
    cell[i] = (uint32_t)cell_of(g, env[i].x, env[i].y, cx, cy);
    idx[i] = (int32_t)i;
}

__global__ void merge_bounds_kernel(const uint32_t *__restrict__ scell, int64_t N, int32_t *__restrict__ cstart,
                                    int32_t *__restrict__ cend) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= N) return;
    const uint32_t c = scell[p];
    if (p == 0 || scell[p - 1] != c) cstart[c] = (int32_t)p;
    if (p == N - 1 || scell[p + 1] != c) cend[c] = (int32_t)p + 1;
}

// FILL=false: count the candidates of polygon i; FILL=true: write them at poff[i]
template <bool FILL>
__global__ void __launch_bounds__(128) merge_pairs_kernel(const double4 *__restrict__ env, const int32_t *__restrict__ rank,
                                                          const int32_t *__restrict__ sidx, const int32_t *__restrict__ cstart,
                                                          const int32_t *__restrict__ cend, int64_t N, Grid g,
                                                          int64_t *__restrict__ pcount, const int64_t *__restrict__ poff,
                                                          int64_t max_pairs, int2 *__restrict__ pairs) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double4 e = env[i];
    const int ri = rank[i];
    int cx, cy;
    cell_of(g, e.x, e.y, cx, cy);
    int64_t k = 0;
    const int64_t base = FILL ? poff[i] : 0;
    for (int yy = max(cy - 1, 0); yy <= min(cy + 1, g.ny - 1); ++yy)
        for (int xx = max(cx - 1, 0); xx <= min(cx + 1, g.nx - 1); ++xx) {
            const int c = yy * g.nx + xx;
            for (int q = cstart[c]; q < cend[c]; ++q) {
                const int j = sidx[q];
                if (j == (int)i) continue;
                const double4 f = env[j];
                // STRtree.query: envelopes intersect (touching counts)
                if (f.x > e.z || f.z < e.x || f.y > e.w || f.w < e.y) continue;
                if (rank[j] < ri) continue; // the pair is emitted by its higher-scoring member
                if (FILL) {
                    if (base + k < max_pairs) pairs[base + k] = make_int2((int)i, j);
                }
                ++k;
            }
        }
    if (!FILL) pcount[i] = k;
}

// integral over [xa,xb] of min(la, lb), the lines given by their end values
__device__ __forceinline__ double min_integral(double xa, double xb, double a0, double a1, double b0, double b1) {
    const double wdt = xb - xa;
    const double d0 = a0 - b0, d1 = a1 - b1;
    if (d0 <= 0.0 && d1 <= 0.0) return 0.5 * (a0 + a1) * wdt;
    if (d0 >= 0.0 && d1 >= 0.0) return 0.5 * (b0 + b1) * wdt;
    const double t = d0 / (d0 - d1);
    const double wc = wdt * t;
    const double ac = a0 + (a1 - a0) * t;
    if (d0 < 0.0) return 0.5 * (a0 + ac) * wc + 0.5 * (ac + b1) * (wdt - wc);
    return 0.5 * (b0 + ac) * wc + 0.5 * (ac + a1) * (wdt - wc);
}

constexpr int kIouWarps = 8;

__global__ void __launch_bounds__(kIouWarps * 32) merge_iou_kernel(const double *__restrict__ xy, const int64_t *__restrict__ voff,
                                                                   const double4 *__restrict__ env, const double *__restrict__ area2,
                                                                   const int2 *__restrict__ pairs, int64_t npairs, double thr,
                                                                   uint8_t *__restrict__ sup, int32_t *__restrict__ indeg) {
    const int64_t pid = blockIdx.x * (int64_t)kIouWarps + (threadIdx.x >> 5);
    if (pid >= npairs) return;
    const int lane = threadIdx.x & 31;
    const int a = pairs[pid].x, b = pairs[pid].y;
    const double *P = xy + 2 * voff[a], *Q = xy + 2 * voff[b];
    const int n = (int)(voff[a + 1] - voff[a]), m = (int)(voff[b + 1] - voff[b]);
    const double oP = area2[a], oQ = area2[b];
    double inter = 0.0;
    if (n >= 3 && m >= 3 && oP != 0.0 && oQ != 0.0) {
        const double ox = fmin(env[a].x, env[b].x), oy = fmin(env[a].y, env[b].y);
        double part = 0.0;
        for (int i = lane; i < n; i += 32) {
            const int i1 = (i + 1 == n) ? 0 : i + 1;
            double ex0 = P[2 * i] - ox, ey0 = P[2 * i + 1] - oy, ex1 = P[2 * i1] - ox, ey1 = P[2 * i1 + 1] - oy;
            if (ex0 == ex1) continue;
            double se = 1.0;
            if (ex0 > ex1) {
                double t = ex0; ex0 = ex1; ex1 = t;
                t = ey0; ey0 = ey1; ey1 = t;
                se = -1.0;
            }
            const double me = (ey1 - ey0) / (ex1 - ex0);
            double s_i = 0.0;
            for (int j = 0; j < m; ++j) {
                const int j1 = (j + 1 == m) ? 0 : j + 1;
                double fx0 = Q[2 * j] - ox, fy0 = Q[2 * j + 1] - oy, fx1 = Q[2 * j1] - ox, fy1 = Q[2 * j1 + 1] - oy;
                if (fx0 == fx1) continue;
                double sf = 1.0;
                if (fx0 > fx1) {
                    double t = fx0; fx0 = fx1; fx1 = t;
                    t = fy0; fy0 = fy1; fy1 = t;
                    sf = -1.0;
                }
                const double xa = ex0 > fx0 ? ex0 : fx0, xb = ex1 < fx1 ? ex1 : fx1;
                if (!(xb > xa)) continue;
                const double mf = (fy1 - fy0) / (fx1 - fx0);
                const double a0 = ey0 + me * (xa - ex0), a1 = ey0 + me * (xb - ex0);
                const double b0 = fy0 + mf * (xa - fx0), b1 = fy0 + mf * (xb - fx0);
                s_i += se * sf * min_integral(xa, xb, a0, a1, b0, b1);
            }
            part += s_i;
        }
        for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if ((oP < 0.0) != (oQ < 0.0)) part = -part;
        inter = part < 0.0 ? 0.0 : part;
    }
    if (lane == 0) {
        const double aP = fabs(oP) * 0.5, aQ = fabs(oQ) * 0.5;
        const double iou = inter / (aP + aQ - inter);
        const bool s = iou > thr;
        sup[pid] = s;
        if (s) atomicAdd(indeg + b, 1);
    }
}

__global__ void merge_fill_in_kernel(const int2 *__restrict__ pairs, const uint8_t *__restrict__ sup, int64_t npairs,
                                     const int32_t *__restrict__ in_off, int32_t *__restrict__ cursor,
                                     int32_t *__restrict__ in_list) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= npairs || !sup[p]) return;
    const int b = pairs[p].y;
    in_list[in_off[b] + atomicAdd(cursor + b, 1)] = pairs[p].x;
}

// state: 0 undecided, 1 kept (fires), 2 suppressed
__global__ void merge_state_init_kernel(const int32_t *__restrict__ indeg, int64_t N, uint8_t *__restrict__ state) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < N) state[i] = indeg[i] == 0 ? 1 : 0;
}

// `frozen` (nullable): nodes whose state is owned elsewhere (halo copies in the multi-GPU merge) are only read
__global__ void merge_round_kernel(const int32_t *__restrict__ in_off, const int32_t *__restrict__ indeg,
                                   const int32_t *__restrict__ in_list, int64_t N, const uint8_t *__restrict__ frozen,
                                   volatile uint8_t *state, int64_t *remaining) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N || state[i] != 0) return;
    if (frozen && frozen[i]) return;
    bool all_sup = true, any_kept = false;
    const int o = in_off[i];
    for (int k = 0; k < indeg[i]; ++k) {
        const uint8_t s = state[in_list[o + k]];
        if (s == 1) {
            any_kept = true;
            break;
        }
        if (s == 0) all_sup = false;
    }
    if (any_kept) state[i] = 2;
    else if (all_sup) state[i] = 1;
    else atomicAdd((unsigned long long *)remaining, 1ull);
}

// 'area' strategy (nuclei_merge.py:143-150): a firing nucleus q is replaced by the largest of the
// nuclei it suppressed first.  owner(c) = lowest-rank kept suppressor of c.
__global__ void merge_owner_kernel(const int32_t *__restrict__ in_off, const int32_t *__restrict__ indeg,
                                   const int32_t *__restrict__ in_list, const uint8_t *__restrict__ state,
                                   const int32_t *__restrict__ rank, const double *__restrict__ area2, int64_t N,
                                   int32_t *__restrict__ owner, unsigned long long *__restrict__ best_area) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    owner[i] = -1;
    if (state[i] != 2) return;
    int best = -1, br = 0x7fffffff;
    const int o = in_off[i];
    for (int k = 0; k < indeg[i]; ++k) {
        const int a = in_list[o + k];
        if (state[a] == 1 && rank[a] < br) {
            br = rank[a];
            best = a;
        }
    }
    owner[i] = best;
    if (best >= 0) atomicMax(best_area + best, (unsigned long long)__double_as_longlong(fabs(area2[i]) * 0.5));
}
__global__ void merge_pick_kernel(const int32_t *__restrict__ owner, const int32_t *__restrict__ rank,
                                  const double *__restrict__ area2, const unsigned long long *__restrict__ best_area, int64_t N,
                                  int32_t *__restrict__ pick_rank) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int q = owner[i];
    if (q < 0) return;
    if ((unsigned long long)__double_as_longlong(fabs(area2[i]) * 0.5) == best_area[q]) atomicMin(pick_rank + q, rank[i]);
}

// flags in rank order: 1 if the nucleus of that rank is in merged_idx
__global__ void merge_flags_kernel(const int32_t *__restrict__ order, const uint8_t *__restrict__ state,
                                   const int32_t *__restrict__ pick_rank, int strategy, int64_t N, uint8_t *__restrict__ flags) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= N) return;
    const int i = order[r];
    if (state[i] != 1) return;
    if (strategy == 1 && pick_rank[i] != 0x7fffffff) flags[pick_rank[i]] = 1;
    else flags[r] = 1;
}

struct MergeWs {
    Extent *ext;
    int64_t *counters; // [0] remaining, [1] total pairs, [2] kept
    double4 *env;
    double *area2;
    unsigned long long *keys_in, *keys_out;
    int32_t *vals_in, *order, *rank;
    uint32_t *cell, *scell;
    int32_t *cidx, *sidx, *cstart, *cend;
    int64_t *pcount, *poff;
    int2 *pairs;
    uint8_t *sup, *state, *flags;
    int32_t *indeg, *in_off, *cursor, *in_list, *owner, *pick_rank;
    unsigned long long *best_area;
    int64_t *order64;
    void *cub_tmp;
    size_t cub_bytes, total;
    int64_t ncell_cap;
};

static MergeWs merge_layout(void *ws, int64_t N, int64_t max_pairs) {
    MergeWs L;
    char *p = (char *)ws;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void *r = p ? p + off : nullptr;
        off += align_up(bytes, 256);
        return r;
    };
    L.ncell_cap = 4 * N + 1024 + 8;
    L.ext = (Extent *)take(sizeof(Extent));
    L.counters = (int64_t *)take(sizeof(int64_t) * 4);
    L.env = (double4 *)take(sizeof(double4) * N);
    L.area2 = (double *)take(sizeof(double) * N);
    L.keys_in = (unsigned long long *)take(8 * N);
    L.keys_out = (unsigned long long *)take(8 * N);
    L.vals_in = (int32_t *)take(4 * N);
    L.order = (int32_t *)take(4 * N);
    L.rank = (int32_t *)take(4 * N);
    L.cell = (uint32_t *)take(4 * N);
    L.scell = (uint32_t *)take(4 * N);
    L.cidx = (int32_t *)take(4 * N);
    L.sidx = (int32_t *)take(4 * N);
    L.cstart = (int32_t *)take(4 * L.ncell_cap);
    L.cend = (int32_t *)take(4 * L.ncell_cap);
    L.pcount = (int64_t *)take(8 * (N + 1));
    L.poff = (int64_t *)take(8 * (N + 1));
    L.pairs = (int2 *)take(sizeof(int2) * max_pairs);
    L.sup = (uint8_t *)take(max_pairs);
    L.state = (uint8_t *)take(N);
    L.flags = (uint8_t *)take(N);
    L.indeg = (int32_t *)take(4 * N);
    L.in_off = (int32_t *)take(4 * (N + 1));
    L.cursor = (int32_t *)take(4 * N);
    L.in_list = (int32_t *)take(4 * max_pairs);
    L.owner = (int32_t *)take(4 * N);
    L.pick_rank = (int32_t *)take(4 * N);
    L.best_area = (unsigned long long *)take(8 * N);
    L.order64 = (int64_t *)take(8 * N);
    size_t b1 = 0, b2 = 0, b3 = 0, b4 = 0, b5 = 0;
    const int64_t n1 = N > 0 ? N : 1;
    cub::DeviceRadixSort::SortPairs(nullptr, b1, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (const int32_t *)nullptr, (int32_t *)nullptr, n1, 0, 64, (cudaStream_t)0);
    cub::DeviceRadixSort::SortPairs(nullptr, b2, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const int32_t *)nullptr,
                                    (int32_t *)nullptr, n1, 0, 32, (cudaStream_t)0);
    cub::DeviceScan::ExclusiveSum(nullptr, b3, (const int64_t *)nullptr, (int64_t *)nullptr, n1 + 1, (cudaStream_t)0);
    cub::DeviceScan::ExclusiveSum(nullptr, b4, (const int32_t *)nullptr, (int32_t *)nullptr, n1 + 1, (cudaStream_t)0);
    cub::DeviceSelect::Flagged(nullptr, b5, (const int64_t *)nullptr, (const uint8_t *)nullptr, (int64_t *)nullptr,
                               (int64_t *)nullptr, n1, (cudaStream_t)0);
    L.cub_bytes = b1;
    if (b2 > L.cub_bytes) L.cub_bytes = b2;
    if (b3 > L.cub_bytes) L.cub_bytes = b3;
    if (b4 > L.cub_bytes) L.cub_bytes = b4;
    if (b5 > L.cub_bytes) L.cub_bytes = b5;
    L.cub_tmp = take(L.cub_bytes);
    L.total = off;
    return L;
}

__global__ void merge_fill_i32_kernel(int32_t *p, int64_t N, int32_t v) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < N) p[i] = v;
}

__global__ void merge_order64_kernel(const int32_t *__restrict__ order, int64_t N, int64_t *__restrict__ order64) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r < N) order64[r] = order[r];
}

} // namespace

// Builds the suppression graph of the N polygons: CSR by suppressed node (in_off/in_list = its higher-ranked
// suppressors), plus the rank order.  Graph arrays may live in the workspace (single-GPU merge) or be
// caller-provided (nuhtc_merge_graph).  Synchronises the stream twice (grid sizing, pair count).
static int build_graph(MergeWs &L, const double *xy, const int64_t *voff, const double *score, int64_t N, double thr,
                       int64_t max_pairs, int32_t *indeg, int32_t *in_off, int32_t *in_list, int64_t *npairs_out,
                       int32_t *status, cudaStream_t st) {
    const unsigned nb = (unsigned)((N + 255) / 256);
    merge_init_kernel<<<1, 1, 0, st>>>(L.ext, L.counters, status);
    merge_stats_kernel<<<nb, 256, 0, st>>>(xy, voff, score, N, L.env, L.area2, L.keys_in, L.vals_in, L.ext);
    size_t cb = L.cub_bytes;
    NUHTC_CUDA(cub::DeviceRadixSort::SortPairs(L.cub_tmp, cb, L.keys_in, L.keys_out, L.vals_in, L.order, N, 0, 64, st));
    merge_rank_kernel<<<nb, 256, 0, st>>>(L.order, N, L.rank);
    // ---- grid dimensions need the global extent on the host
    Extent hext;
    NUHTC_CUDA(cudaMemcpyAsync(&hext, L.ext, sizeof hext, cudaMemcpyDeviceToHost, st));
    NUHTC_CUDA(cudaStreamSynchronize(st));
    Grid g;
    g.gx0 = ordered_dbl_host(hext.minx);
    g.gy0 = ordered_dbl_host(hext.miny);
    const double gx1 = ordered_dbl_host(hext.maxx), gy1 = ordered_dbl_host(hext.maxy);
    double cell = ordered_dbl_host(hext.maxext);
    while ((floor((gx1 - g.gx0) / cell) + 1.0) * (floor((gy1 - g.gy0) / cell) + 1.0) > 4.0 * (double)N + 1024.0) cell *= 2.0;
    g.cell = cell;
    g.nx = (int)floor((gx1 - g.gx0) / cell) + 1;
    g.ny = (int)floor((gy1 - g.gy0) / cell) + 1;
    const int64_t ncell = (int64_t)g.nx * g.ny;
    merge_cell_kernel<<<nb, 256, 0, st>>>(L.env, N, g, L.cell, L.cidx);
    int cbits = 1;
    while ((1ll << cbits) < ncell) ++cbits;
    cb = L.cub_bytes;
    NUHTC_CUDA(cub::DeviceRadixSort::SortPairs(L.cub_tmp, cb, L.cell, L.scell, L.cidx, L.sidx, N, 0, cbits, st));
    NUHTC_CUDA(cudaMemsetAsync(L.cstart, 0, 4 * ncell, st));
    NUHTC_CUDA(cudaMemsetAsync(L.cend, 0, 4 * ncell, st));
    merge_bounds_kernel<<<nb, 256, 0, st>>>(L.scell, N, L.cstart, L.cend);
    // ---- candidate pairs
    const unsigned nb128 = (unsigned)((N + 127) / 128);
    merge_pairs_kernel<false><<<nb128, 128, 0, st>>>(L.env, L.rank, L.sidx, L.cstart, L.cend, N, g, L.pcount, nullptr, 0, nullptr);
    NUHTC_CUDA(cudaMemsetAsync(L.pcount + N, 0, 8, st));
    cb = L.cub_bytes;
    NUHTC_CUDA(cub::DeviceScan::ExclusiveSum(L.cub_tmp, cb, L.pcount, L.poff, N + 1, st));
    int64_t npairs = 0;
    NUHTC_CUDA(cudaMemcpyAsync(&npairs, L.poff + N, 8, cudaMemcpyDeviceToHost, st));
    NUHTC_CUDA(cudaStreamSynchronize(st));
    *npairs_out = npairs;
    if (npairs > max_pairs) {
        const int32_t one = 1;
        NUHTC_CUDA(cudaMemcpyAsync(status, &one, 4, cudaMemcpyHostToDevice, st));
        NUHTC_CUDA(cudaStreamSynchronize(st));
        nuhtc_set_error("merge: %lld candidate pairs exceed max_pairs=%lld", (long long)npairs, (long long)max_pairs);
        return NUHTC_EOVERFLOW;
    }
    NUHTC_CUDA(cudaMemsetAsync(indeg, 0, 4 * N, st));
    NUHTC_CUDA(cudaMemsetAsync(L.cursor, 0, 4 * N, st));
    if (npairs > 0) {
        merge_pairs_kernel<true><<<nb128, 128, 0, st>>>(L.env, L.rank, L.sidx, L.cstart, L.cend, N, g, nullptr, L.poff, max_pairs,
                                                         L.pairs);
        merge_iou_kernel<<<(unsigned)((npairs + kIouWarps - 1) / kIouWarps), kIouWarps * 32, 0, st>>>(xy, voff, L.env, L.area2, L.pairs,
                                                                                                     npairs, thr, L.sup, indeg);
    }
    NUHTC_CUDA(cudaMemsetAsync(in_off + N, 0, 4, st));
    NUHTC_CUDA(cudaMemcpyAsync(in_off, indeg, 4 * N, cudaMemcpyDeviceToDevice, st));
    cb = L.cub_bytes;
    NUHTC_CUDA(cub::DeviceScan::ExclusiveSum(L.cub_tmp, cb, in_off, in_off, N + 1, st));
    if (npairs > 0)
        merge_fill_in_kernel<<<(unsigned)((npairs + 255) / 256), 256, 0, st>>>(L.pairs, L.sup, npairs, in_off, L.cursor, in_list);
    NUHTC_LAUNCH_CHECK();
    return NUHTC_OK;
}

NUHTC_API size_t nuhtc_merge_workspace_bytes(int64_t N, int64_t sumV, int64_t max_pairs) {
    (void)sumV;
    if (N <= 0) return 256;
    if (max_pairs < 1) max_pairs = 1;
    return merge_layout(nullptr, N, max_pairs).total;
}

static int merge_check_args(const double *xy, const int64_t *voff, const double *score, int64_t N, double thr, void *ws) {
    NUHTC_CHECK_ARG(N >= 0 && N < (1ll << 31) - 1, "merge: N out of range");
    NUHTC_CHECK_ARG(thr >= 0.0, "merge: overlap_threshold must be >= 0");
    NUHTC_CHECK_ARG(N == 0 || (xy && voff && score && ws), "merge: null pointer");
    return NUHTC_OK;
}

NUHTC_API int nuhtc_merge_graph(const double *xy, const int64_t *voff, const double *score, int64_t N, int64_t sumV, double thr,
                                int64_t max_pairs, int32_t *indeg, int32_t *in_off, int32_t *in_list, int64_t *num_pairs,
                                int32_t *status, void *ws, size_t ws_bytes, void *stream) {
    (void)sumV;
    int rc = merge_check_args(xy, voff, score, N, thr, ws);
    if (rc) return rc;
    NUHTC_CHECK_ARG(num_pairs && status, "merge_graph: null output pointer");
    cudaStream_t st = (cudaStream_t)stream;
    *num_pairs = 0;
    if (N == 0) {
        NUHTC_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
        if (in_off) NUHTC_CUDA(cudaMemsetAsync(in_off, 0, 4, st));
        return NUHTC_OK;
    }
    NUHTC_CHECK_ARG(indeg && in_off && in_list, "merge_graph: null output pointer");
    if (max_pairs < 1) max_pairs = 1;
    MergeWs L = merge_layout(ws, N, max_pairs);
    if (L.total > ws_bytes) {
        nuhtc_set_error("merge: workspace %zu < required %zu", ws_bytes, L.total);
        return NUHTC_EWORKSPACE;
    }
    return build_graph(L, xy, voff, score, N, thr, max_pairs, indeg, in_off, in_list, num_pairs, status, st);
}

namespace {
// y extent of every ring (the multi-GPU merge classifies nuclei against the stripe extents with it): one thread per ring
__global__ void ring_yextent_kernel(const double *__restrict__ xy, const int64_t *__restrict__ voff, int64_t N,
                                    double *__restrict__ ymin, double *__restrict__ ymax) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    double lo = 1e300, hi = -1e300;
    for (int64_t v = voff[i]; v < voff[i + 1]; ++v) {
        const double y = xy[2 * v + 1];
        lo = fmin(lo, y);
        hi = fmax(hi, y);
    }
    ymin[i] = lo;
    ymax[i] = hi;
}

} // namespace

NUHTC_API int nuhtc_ring_yextent(const double *xy, const int64_t *voff, int64_t N, double *ymin, double *ymax, void *stream) {
    NUHTC_CHECK_ARG(N >= 0, "ring_yextent: bad size");
    if (N == 0) return NUHTC_OK;
    NUHTC_CHECK_ARG(xy && voff && ymin && ymax, "ring_yextent: null pointer");
    ring_yextent_kernel<<<(unsigned)((N + 127) / 128), 128, 0, (cudaStream_t)stream>>>(xy, voff, N, ymin, ymax);
    NUHTC_LAUNCH_CHECK();
    return NUHTC_OK;
}

NUHTC_API int nuhtc_merge_rounds(const int32_t *in_off, const int32_t *indeg, const int32_t *in_list, int64_t N,
                                 const uint8_t *frozen, uint8_t *state, int64_t *remaining, int rounds, void *stream) {
    NUHTC_CHECK_ARG(N >= 0 && rounds >= 1, "merge_rounds: bad sizes");
    NUHTC_CHECK_ARG(remaining != nullptr, "merge_rounds: null output pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) {
        NUHTC_CUDA(cudaMemsetAsync(remaining, 0, 8, st));
        return NUHTC_OK;
    }
    NUHTC_CHECK_ARG(in_off && indeg && in_list && state, "merge_rounds: null pointer");
    const unsigned nb = (unsigned)((N + 255) / 256);
    for (int r = 0; r < rounds; ++r) {
        if (r == rounds - 1) NUHTC_CUDA(cudaMemsetAsync(remaining, 0, 8, st)); // only the last round's count is meaningful
        merge_round_kernel<<<nb, 256, 0, st>>>(in_off, indeg, in_list, N, frozen, state, remaining);
    }
    NUHTC_LAUNCH_CHECK();
    return NUHTC_OK;
}

NUHTC_API int nuhtc_merge(const double *xy, const int64_t *voff, const double *score, int64_t N, int64_t sumV, double thr,
                          int strategy, int64_t max_pairs, int64_t *keep_ids, int64_t *num_keep, int32_t *status, void *ws,
                          size_t ws_bytes, void *stream) {
    (void)sumV;
    int rc = merge_check_args(xy, voff, score, N, thr, ws);
    if (rc) return rc;
    NUHTC_CHECK_ARG(strategy == 0 || strategy == 1, "merge: strategy must be 0 (probability) or 1 (area)");
    NUHTC_CHECK_ARG(num_keep && status, "merge: null output pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) {
        NUHTC_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int64_t), st));
        NUHTC_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
        return NUHTC_OK;
    }
    NUHTC_CHECK_ARG(keep_ids != nullptr, "merge: null pointer");
    if (max_pairs < 1) max_pairs = 1;
    MergeWs L = merge_layout(ws, N, max_pairs);
    if (L.total > ws_bytes) {
        nuhtc_set_error("merge: workspace %zu < required %zu", ws_bytes, L.total);
        return NUHTC_EWORKSPACE;
    }
    const unsigned nb = (unsigned)((N + 255) / 256);
    int64_t npairs = 0;
    rc = build_graph(L, xy, voff, score, N, thr, max_pairs, L.indeg, L.in_off, L.in_list, &npairs, status, st);
    if (rc == NUHTC_EOVERFLOW) {
        NUHTC_CUDA(cudaMemcpyAsync(num_keep, &npairs, 8, cudaMemcpyHostToDevice, st)); // tells the caller what to retry with
        NUHTC_CUDA(cudaStreamSynchronize(st));
    }
    if (rc) return rc;
    merge_state_init_kernel<<<nb, 256, 0, st>>>(L.indeg, N, L.state);
    // ---- resolve rounds
    for (int guard = 0; guard < (1 << 20); ++guard) {
        rc = nuhtc_merge_rounds(L.in_off, L.indeg, L.in_list, N, nullptr, L.state, L.counters, 4, stream);
        if (rc) return rc;
        int64_t remaining = 0;
        NUHTC_CUDA(cudaMemcpyAsync(&remaining, L.counters, 8, cudaMemcpyDeviceToHost, st));
        NUHTC_CUDA(cudaStreamSynchronize(st));
        if (remaining == 0) break;
    }
    // ---- emit
    NUHTC_CUDA(cudaMemsetAsync(L.flags, 0, N, st));
    merge_fill_i32_kernel<<<nb, 256, 0, st>>>(L.pick_rank, N, 0x7fffffff);
    if (strategy == 1) {
        NUHTC_CUDA(cudaMemsetAsync(L.best_area, 0, 8 * N, st));
        merge_owner_kernel<<<nb, 256, 0, st>>>(L.in_off, L.indeg, L.in_list, L.state, L.rank, L.area2, N, L.owner, L.best_area);
        merge_pick_kernel<<<nb, 256, 0, st>>>(L.owner, L.rank, L.area2, L.best_area, N, L.pick_rank);
    }
    merge_flags_kernel<<<nb, 256, 0, st>>>(L.order, L.state, L.pick_rank, strategy, N, L.flags);
    merge_order64_kernel<<<nb, 256, 0, st>>>(L.order, N, L.order64);
    size_t cb = L.cub_bytes;
    NUHTC_CUDA(cub::DeviceSelect::Flagged(L.cub_tmp, cb, L.order64, L.flags, keep_ids, num_keep, N, st));
    NUHTC_LAUNCH_CHECK();
    return NUHTC_OK;
}
