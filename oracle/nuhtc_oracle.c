/*
 * nuhtc_oracle.c -- CPU restatement of the NuHTC RoI-stage + merge arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under nuhtc_b200/ may import, link or
 * execute this file; it is the checker used by tests/, __graft_entry__.smoke()
 * and the cpu_baseline / --impl reference legs of bench.py.
 *
 * PARITY STATUS: the native arithmetic of the reference lives in un-vendored
 * third-party wheels (mmcv-full==1.7.2, pycocotools==2.0.7, shapely>=2.0 --
 * /root/reference/Dockerfile:46,60, requirements.txt:20) that are neither in
 * /root/reference nor installable here.  The functions below restate the
 * published algorithms of those packages:
 *   - oracle_roi_align_fwd : mmcv 1.7.2 ops/csrc/pytorch/cpu/roi_align.cpp
 *     (ROIAlignForward + pre_calc_for_bilinear_interpolate), reached from
 *     mmdet/models/roi_heads/roi_extractors/base_roi_extractor.py:54-60 and
 *     single_level_roi_extractor.py:79,96,103.  Pinned bit-for-bit against
 *     torchvision.ops.roi_align(aligned=True) CPU (same Detectron lineage) in
 *     tests/test_oracle_cpu.py; NOT pinned against mmcv itself ("parity
 *     unpinned" for the mmcv boundary).
 *   - oracle_nms           : mmcv 1.7.2 ops/csrc/pytorch/cpu/nms.cpp (nms_cpu),
 *     reached from nuhtc/models/bbox_head.py:93,208.  Pinned against
 *     torchvision.ops.nms CPU on distinct scores.
 *   - oracle_rle_* / oracle_mask_iou : pycocotools common/maskApi.c
 *     (rleEncode, rleArea, rleToBbox, bbIou, rleIou) as used by
 *     tools/infer_wsi.py:60-84.  Integer arithmetic; pinned against a
 *     brute-force numpy pixel count.
 *   - oracle_poly_* / oracle_merge : shapely `intersection().area` semantics
 *     used by tools/nuclei_merge.py:107-150 (exact area of the set
 *     intersection of two simple polygons, IoU in double, greedy in score
 *     order).  GEOS is absent, so the area is computed with the signed
 *     trapezoid identity; pinned against exact rational arithmetic and pixel
 *     counting (tests/test_oracle_cpu.py).  "parity unpinned" vs GEOS.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off; no fast-math so the
 * float op order written here is the op order executed).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define API __attribute__((visibility("default")))

/* minimal pthread parallel-for: items [0,n) handed out in blocks through an atomic cursor */
typedef void (*pf_body_t)(int64_t i, void *ctx);
typedef struct { pf_body_t body; void *ctx; int64_t n, blk; int64_t *cursor; } pf_job_t;
static void *pf_worker(void *arg) {
    pf_job_t *j = (pf_job_t *)arg;
    for (;;) {
        const int64_t b = __atomic_fetch_add(j->cursor, j->blk, __ATOMIC_RELAXED);
        if (b >= j->n) break;
        const int64_t e = b + j->blk < j->n ? b + j->blk : j->n;
        for (int64_t i = b; i < e; ++i) j->body(i, j->ctx);
    }
    return NULL;
}
static void parallel_for(int64_t n, int64_t blk, int nthreads, pf_body_t body, void *ctx) {
    int64_t cursor = 0;
    pf_job_t job = {body, ctx, n, blk, &cursor};
    if (nthreads <= 1 || n <= blk) { for (int64_t i = 0; i < n; ++i) body(i, ctx); return; }
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, pf_worker, &job);
    pf_worker(&job);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
}

/* ------------------------------------------------------------------------ */
/* A1. RoIAlign forward, avg pool, NCHW fp32.                                */
/* ------------------------------------------------------------------------ */
typedef struct {
    int p1, p2, p3, p4;
    float w1, w2, w3, w4;
} tap4_t;

static void roi_align_one(const float *in, int C, int H, int W, const float *roi, int PH, int PW,
                          float scale, int sr, int aligned, float *out /* [C,PH,PW] */) {
    const float off = aligned ? 0.5f : 0.0f;
    float sw = roi[1] * scale - off;
    float sh = roi[2] * scale - off;
    float ew = roi[3] * scale - off;
    float eh = roi[4] * scale - off;
    float rw = ew - sw;
    float rh = eh - sh;
    if (!aligned) {
        if (rw < 1.0f) rw = 1.0f;
        if (rh < 1.0f) rh = 1.0f;
    }
    const float bh = rh / (float)PH;
    const float bw = rw / (float)PW;
    const int gh = sr > 0 ? sr : (int)ceilf(rh / (float)PH);
    const int gw = sr > 0 ? sr : (int)ceilf(rw / (float)PW);
    int cnt_i = gh * gw;
    if (cnt_i < 1) cnt_i = 1;
    const float count = (float)cnt_i;
    const size_t ntap = (size_t)(gh > 0 ? gh : 0) * (size_t)(gw > 0 ? gw : 0) * PH * PW;
    tap4_t *taps = (tap4_t *)malloc((ntap ? ntap : 1) * sizeof(tap4_t));
    size_t t = 0;
    for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw)
            for (int iy = 0; iy < gh; ++iy) {
                const float yy = sh + (float)ph * bh + ((float)iy + .5f) * bh / (float)gh;
                for (int ix = 0; ix < gw; ++ix) {
                    const float xx = sw + (float)pw * bw + ((float)ix + .5f) * bw / (float)gw;
                    float x = xx, y = yy;
                    tap4_t tp;
                    if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) {
                        memset(&tp, 0, sizeof tp);
                        taps[t++] = tp;
                        continue;
                    }
                    if (y <= 0) y = 0;
                    if (x <= 0) x = 0;
                    int yl = (int)y, xl = (int)x, yh, xh;
                    if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
                    if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
                    const float ly = y - (float)yl, lx = x - (float)xl;
                    const float hy = 1.0f - ly, hx = 1.0f - lx;
                    tp.p1 = yl * W + xl; tp.p2 = yl * W + xh;
                    tp.p3 = yh * W + xl; tp.p4 = yh * W + xh;
                    tp.w1 = hy * hx; tp.w2 = hy * lx; tp.w3 = ly * hx; tp.w4 = ly * lx;
                    taps[t++] = tp;
                }
            }
    const int b = (int)roi[0];
    for (int c = 0; c < C; ++c) {
        const float *plane = in + ((size_t)b * C + c) * (size_t)H * W;
        t = 0;
        for (int ph = 0; ph < PH; ++ph)
            for (int pw = 0; pw < PW; ++pw) {
                float acc = 0.f;
                for (int iy = 0; iy < gh; ++iy)
                    for (int ix = 0; ix < gw; ++ix) {
                        const tap4_t tp = taps[t++];
                        acc += tp.w1 * plane[tp.p1] + tp.w2 * plane[tp.p2] + tp.w3 * plane[tp.p3] +
                               tp.w4 * plane[tp.p4];
                    }
                out[((size_t)c * PH + ph) * PW + pw] = acc / count;
            }
    }
    free(taps);
}

typedef struct { const float *in; int C, H, W; const float *rois; int PH, PW; float scale; int sr, aligned; float *out; } ra_ctx_t;
static void ra_body(int64_t k, void *p) {
    ra_ctx_t *c = (ra_ctx_t *)p;
    roi_align_one(c->in, c->C, c->H, c->W, c->rois + (size_t)k * 5, c->PH, c->PW, c->scale, c->sr, c->aligned,
                  c->out + (size_t)k * c->C * c->PH * c->PW);
}

/* in [B,C,H,W]; rois [K,5] (batch,x1,y1,x2,y2); out [K,C,PH,PW].  nthreads<=1: serial
 * (the mmcv CPU kernel is a serial loop over RoIs); >1: RoIs split over pthreads,
 * identical results (RoIs are independent). */
API int oracle_roi_align_fwd(const float *in, int B, int C, int H, int W, const float *rois, int K,
                             int PH, int PW, float scale, int sr, int aligned, float *out,
                             int nthreads) {
    (void)B;
    ra_ctx_t c = {in, C, H, W, rois, PH, PW, scale, sr, aligned, out};
    parallel_for(K, 16, nthreads, ra_body, &c);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* A2. NMS (mmcv nms_cpu).  order: score descending, ties -> lower index     */
/* first (mmcv's sort is unstable; ties are outside the parity contract).    */
/* ------------------------------------------------------------------------ */
typedef struct { float s; int64_t i; } sc_t;
static int sc_desc(const void *a, const void *b) {
    const sc_t *x = (const sc_t *)a, *y = (const sc_t *)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return x->i < y->i ? -1 : (x->i > y->i ? 1 : 0);
}

API int64_t oracle_nms(const float *boxes, const float *scores, int64_t N, float thr, int offset,
                       int64_t *keep) {
    if (N <= 0) return 0;
    sc_t *ord = (sc_t *)malloc(N * sizeof(sc_t));
    float *area = (float *)malloc(N * sizeof(float));
    uint8_t *sel = (uint8_t *)malloc(N);
    const float fo = (float)offset;
    for (int64_t i = 0; i < N; ++i) {
        ord[i].s = scores[i]; ord[i].i = i; sel[i] = 1;
        area[i] = (boxes[4 * i + 2] - boxes[4 * i] + fo) * (boxes[4 * i + 3] - boxes[4 * i + 1] + fo);
    }
    qsort(ord, N, sizeof(sc_t), sc_desc);
    int64_t k = 0;
    for (int64_t _i = 0; _i < N; ++_i) {
        if (!sel[_i]) continue;
        const int64_t i = ord[_i].i;
        keep[k++] = i;
        const float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2],
                    iy2 = boxes[4 * i + 3], ia = area[i];
        for (int64_t _j = _i + 1; _j < N; ++_j) {
            if (!sel[_j]) continue;
            const int64_t j = ord[_j].i;
            const float xx1 = fmaxf(ix1, boxes[4 * j]), yy1 = fmaxf(iy1, boxes[4 * j + 1]);
            const float xx2 = fminf(ix2, boxes[4 * j + 2]), yy2 = fminf(iy2, boxes[4 * j + 3]);
            const float w = fmaxf(0.f, xx2 - xx1 + fo), h = fmaxf(0.f, yy2 - yy1 + fo);
            const float inter = w * h;
            const float ovr = inter / (ia + area[j] - inter);
            if (ovr > thr) sel[_j] = 0;
        }
    }
    free(ord); free(area); free(sel);
    return k;
}

/* ------------------------------------------------------------------------ */
/* A4. pycocotools RLE: encode (column-major), area, bbox, IoU (iscrowd=0).  */
/* mask: uint8 [n,h,w] row-major (C order) as infer_wsi.py holds them; the   */
/* reference calls np.asfortranarray first, so runs are along columns.       */
/* ------------------------------------------------------------------------ */
/* counts buffer must hold h*w+1 entries; returns number of runs. */
API int oracle_rle_encode(const uint8_t *mask, int h, int w, uint32_t *cnts) {
    int m = 0;
    uint32_t run = 0;
    uint8_t prev = 0;
    for (int x = 0; x < w; ++x)
        for (int y = 0; y < h; ++y) {
            const uint8_t v = mask[(size_t)y * w + x] ? 1 : 0;
            if (v != prev) { cnts[m++] = run; run = 0; prev = v; }
            ++run;
        }
    cnts[m++] = run;
    return m;
}

static uint32_t rle_area(const uint32_t *c, int m) {
    uint32_t a = 0;
    for (int j = 1; j < m; j += 2) a += c[j];
    return a;
}

static void rle_bbox(const uint32_t *c, int m, int h, double *bb) {
    /* rleToBbox: x,y,w,h of the tight box, from run starts/ends */
    const int mm = (m / 2) * 2; /* a trailing zero-run carries no foreground */
    if (mm == 0) { bb[0] = bb[1] = bb[2] = bb[3] = 0; return; }
    const uint32_t hh = (uint32_t)h;
    uint32_t xs = 0xffffffffu, ys = hh, xe = 0, ye = 0, cc = 0, xp = 0;
    for (int j = 0; j < mm; ++j) {
        cc += c[j];
        const uint32_t t = cc - (uint32_t)(j % 2); /* run start (even j) / last pixel of run (odd j) */
        const uint32_t y = t % hh, x = (t - y) / hh;
        if (j % 2 == 0) xp = x;
        else if (xp < x) { ys = 0; ye = hh - 1; } /* a run that wraps a column spans all rows */
        if (x < xs) xs = x;
        if (x > xe) xe = x;
        if (y < ys) ys = y;
        if (y > ye) ye = y;
    }
    bb[0] = xs; bb[1] = ys; bb[2] = (double)(xe - xs + 1); bb[3] = (double)(ye - ys + 1);
}

/* masks [n,h,w] uint8; iou [n,n] double, iou[d*n+g] as maskUtils.iou(dt,gt) returns [m,n]. */
API int oracle_mask_iou(const uint8_t *masks, int n, int h, int w, double *iou) {
    uint32_t **cn = (uint32_t **)malloc(sizeof(uint32_t *) * (n ? n : 1));
    int *mm = (int *)malloc(sizeof(int) * (n ? n : 1));
    double *bb = (double *)malloc(sizeof(double) * 4 * (n ? n : 1));
    uint32_t *tmp = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)h * w + 1));
    for (int i = 0; i < n; ++i) {
        mm[i] = oracle_rle_encode(masks + (size_t)i * h * w, h, w, tmp);
        cn[i] = (uint32_t *)malloc(sizeof(uint32_t) * mm[i]);
        memcpy(cn[i], tmp, sizeof(uint32_t) * mm[i]);
        rle_bbox(cn[i], mm[i], h, bb + 4 * i);
    }
    for (int g = 0; g < n; ++g)
        for (int d = 0; d < n; ++d) {
            /* bbIou prefilter */
            const double *D = bb + 4 * d, *G = bb + 4 * g;
            double o = 0;
            double ww = fmin(D[2] + D[0], G[2] + G[0]) - fmax(D[0], G[0]);
            double hh = fmin(D[3] + D[1], G[3] + G[1]) - fmax(D[1], G[1]);
            if (ww > 0 && hh > 0) {
                double i = ww * hh, u = D[2] * D[3] + G[2] * G[3] - i;
                o = i / u;
            }
            if (o > 0) {
                /* run-merge walk over the two RLEs */
                const uint32_t *A = cn[d], *Bc = cn[g];
                int ka = mm[d], kb = mm[g], a = 1, b = 1, va = 0, vb = 0;
                uint32_t ca = A[0], cb = Bc[0], c, ct = 1, ii = 0, uu = 0;
                while (ct > 0) {
                    c = ca < cb ? ca : cb;
                    if (va || vb) { uu += c; if (va && vb) ii += c; }
                    ct = 0;
                    ca -= c; if (!ca && a < ka) { ca = A[a++]; va = !va; } ct += ca;
                    cb -= c; if (!cb && b < kb) { cb = Bc[b++]; vb = !vb; } ct += cb;
                }
                if (ii == 0) uu = 1;
                o = (double)ii / (double)uu;
            }
            iou[(size_t)d * n + g] = o;
        }
    for (int i = 0; i < n; ++i) free(cn[i]);
    free(cn); free(mm); free(bb); free(tmp);
    return 0;
}

API int oracle_mask_area(const uint8_t *masks, int n, int h, int w, int64_t *area) {
    uint32_t *tmp = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)h * w + 1));
    for (int i = 0; i < n; ++i) {
        int m = oracle_rle_encode(masks + (size_t)i * h * w, h, w, tmp);
        area[i] = rle_area(tmp, m);
    }
    free(tmp);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* A5. polygon area / intersection area / IoU, double.                       */
/* Rings: xy [V,2] doubles, closed or open (a repeated last vertex is a      */
/* zero-length edge and contributes nothing).                                */
/* ------------------------------------------------------------------------ */
API double oracle_poly_area2(const double *xy, int V) { /* twice the signed area */
    if (V < 3) return 0.0;
    const double ox = xy[0], oy = xy[1];
    double s = 0.0;
    for (int i = 0; i < V; ++i) {
        const int j = (i + 1 == V) ? 0 : i + 1;
        const double x0 = xy[2 * i] - ox, y0 = xy[2 * i + 1] - oy;
        const double x1 = xy[2 * j] - ox, y1 = xy[2 * j + 1] - oy;
        s += x0 * y1 - x1 * y0;
    }
    return s;
}

/* integral over [xa,xb] of min(la, lb), la/lb given by their end values */
static double min_integral(double xa, double xb, double a0, double a1, double b0, double b1) {
    const double wdt = xb - xa;
    const double d0 = a0 - b0, d1 = a1 - b1;
    if (d0 <= 0.0 && d1 <= 0.0) return 0.5 * (a0 + a1) * wdt;
    if (d0 >= 0.0 && d1 >= 0.0) return 0.5 * (b0 + b1) * wdt;
    const double t = d0 / (d0 - d1); /* crossing parameter in (0,1) */
    const double wc = wdt * t;
    const double ac = a0 + (a1 - a0) * t; /* value of both lines at the crossing */
    if (d0 < 0.0) /* a below first, then b */
        return 0.5 * (a0 + ac) * wc + 0.5 * (ac + b1) * (wdt - wc);
    return 0.5 * (b0 + ac) * wc + 0.5 * (ac + a1) * (wdt - wc);
}

/* Area of P∩Q for simple polygons via 1_P = o_P * sum_e -sgn(dx_e) 1_{T(e)},
 * T(e) the trapezoid between edge e and the baseline y=0 of a local frame in
 * which every y >= 0. */
API double oracle_poly_inter_area(const double *P, int n, const double *Q, int m) {
    if (n < 3 || m < 3) return 0.0;
    double ox = P[0], oy = P[1];
    for (int i = 0; i < n; ++i) { if (P[2 * i] < ox) ox = P[2 * i]; if (P[2 * i + 1] < oy) oy = P[2 * i + 1]; }
    for (int i = 0; i < m; ++i) { if (Q[2 * i] < ox) ox = Q[2 * i]; if (Q[2 * i + 1] < oy) oy = Q[2 * i + 1]; }
    const double oP = oracle_poly_area2(P, n), oQ = oracle_poly_area2(Q, m);
    if (oP == 0.0 || oQ == 0.0) return 0.0;
    /* summation order: edge i of P contributes s_i = sum_j term(i,j) (j ascending); the s_i are
       accumulated into 32 interleaved partials (i mod 32) that are combined by an xor-butterfly --
       the order a 32-lane evaluation produces, so the CUDA kernel can match this bit for bit */
    double part[32];
    for (int l = 0; l < 32; ++l) part[l] = 0.0;
    for (int i = 0; i < n; ++i) {
        const int i1 = (i + 1 == n) ? 0 : i + 1;
        double ex0 = P[2 * i] - ox, ey0 = P[2 * i + 1] - oy, ex1 = P[2 * i1] - ox, ey1 = P[2 * i1 + 1] - oy;
        if (ex0 == ex1) continue;
        double se = 1.0;
        if (ex0 > ex1) { double t; t = ex0; ex0 = ex1; ex1 = t; t = ey0; ey0 = ey1; ey1 = t; se = -1.0; }
        const double me = (ey1 - ey0) / (ex1 - ex0);
        double s_i = 0.0;
        for (int j = 0; j < m; ++j) {
            const int j1 = (j + 1 == m) ? 0 : j + 1;
            double fx0 = Q[2 * j] - ox, fy0 = Q[2 * j + 1] - oy, fx1 = Q[2 * j1] - ox, fy1 = Q[2 * j1 + 1] - oy;
            if (fx0 == fx1) continue;
            double sf = 1.0;
            if (fx0 > fx1) { double t; t = fx0; fx0 = fx1; fx1 = t; t = fy0; fy0 = fy1; fy1 = t; sf = -1.0; }
            const double xa = ex0 > fx0 ? ex0 : fx0, xb = ex1 < fx1 ? ex1 : fx1;
            if (!(xb > xa)) continue;
            const double mf = (fy1 - fy0) / (fx1 - fx0);
            const double a0 = ey0 + me * (xa - ex0), a1 = ey0 + me * (xb - ex0);
            const double b0 = fy0 + mf * (xa - fx0), b1 = fy0 + mf * (xb - fx0);
            s_i += se * sf * min_integral(xa, xb, a0, a1, b0, b1);
        }
        part[i & 31] += s_i;
    }
    for (int o = 16; o; o >>= 1) {
        double nxt[32];
        for (int l = 0; l < 32; ++l) nxt[l] = part[l] + part[l ^ o];
        for (int l = 0; l < 32; ++l) part[l] = nxt[l];
    }
    double acc = part[0];
    if ((oP < 0.0) != (oQ < 0.0)) acc = -acc;
    return acc < 0.0 ? 0.0 : acc;
}

API double oracle_poly_iou(const double *P, int n, const double *Q, int m) {
    const double inter = oracle_poly_inter_area(P, n, Q, m);
    const double aP = fabs(oracle_poly_area2(P, n)) * 0.5, aQ = fabs(oracle_poly_area2(Q, m)) * 0.5;
    return inter / (aP + aQ - inter);
}

/* ------------------------------------------------------------------------ */
/* SECOND, INDEPENDENT anchor for the polygon intersection area: a vertical    */
/* slab decomposition with the even-odd rule.  It shares nothing with          */
/* oracle_poly_inter_area above (no orientation, no signed trapezoids, no      */
/* butterfly summation): the x axis is cut at every vertex and every proper    */
/* edge crossing; inside a slab no two edges cross, so each polygon is a        */
/* sorted stack of y-intervals (consecutive pairs of its edges that span the   */
/* slab) bounded by lines, and the area of the overlap of two such intervals    */
/* is (width) x (overlap height at mid-slab), exact for linear bounds.         */
/* Used only by tests/test_oracle_cpu.py to check the trapezoid oracle on      */
/* >= 1e4 contour-shaped pairs ("bent to the kernel").             */
/* ------------------------------------------------------------------------ */
static int dbl_asc(const void *a, const void *b) {
    const double x = *(const double *)a, y = *(const double *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}
typedef struct { double ym, y0, y1; } slab_edge_t;
static int slab_edge_asc(const void *a, const void *b) {
    const slab_edge_t *x = (const slab_edge_t *)a, *y = (const slab_edge_t *)b;
    if (x->ym < y->ym) return -1;
    if (x->ym > y->ym) return 1;
    /* equal at mid-slab (two edges that meet only at a slab boundary cannot be; collinear ones can): any order */
    return 0;
}
static int slab_collect(const double *P, int n, double xa, double xb, slab_edge_t *out) {
    int k = 0;
    const double xm = 0.5 * (xa + xb);
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1 == n) ? 0 : i + 1;
        double x0 = P[2 * i], y0 = P[2 * i + 1], x1 = P[2 * j], y1 = P[2 * j + 1];
        if (x0 == x1) continue;
        if (x0 > x1) { double t; t = x0; x0 = x1; x1 = t; t = y0; y0 = y1; y1 = t; }
        if (x0 <= xa && x1 >= xb) {
            const double m = (y1 - y0) / (x1 - x0);
            out[k].ym = y0 + m * (xm - x0);
            out[k].y0 = y0 + m * (xa - x0);
            out[k].y1 = y0 + m * (xb - x0);
            ++k;
        }
    }
    qsort(out, k, sizeof(slab_edge_t), slab_edge_asc);
    return k;
}
API double oracle_poly_inter_area_slab(const double *P, int n, const double *Q, int m) {
    if (n < 3 || m < 3) return 0.0;
    int cap = n + m + n * m, nx = 0;
    double *xs = (double *)malloc(sizeof(double) * cap);
    for (int i = 0; i < n; ++i) xs[nx++] = P[2 * i];
    for (int j = 0; j < m; ++j) xs[nx++] = Q[2 * j];
    for (int i = 0; i < n; ++i) {
        const int i1 = (i + 1 == n) ? 0 : i + 1;
        const double ax = P[2 * i], ay = P[2 * i + 1], bx = P[2 * i1], by = P[2 * i1 + 1];
        for (int j = 0; j < m; ++j) {
            const int j1 = (j + 1 == m) ? 0 : j + 1;
            const double cx = Q[2 * j], cy = Q[2 * j + 1], dx = Q[2 * j1], dy = Q[2 * j1 + 1];
            const double den = (bx - ax) * (dy - cy) - (by - ay) * (dx - cx);
            if (den == 0.0) continue; /* parallel or collinear: no isolated crossing */
            const double t = ((cx - ax) * (dy - cy) - (cy - ay) * (dx - cx)) / den;
            const double u = ((cx - ax) * (by - ay) - (cy - ay) * (bx - ax)) / den;
            if (t > 0.0 && t < 1.0 && u > 0.0 && u < 1.0) xs[nx++] = ax + t * (bx - ax);
        }
    }
    qsort(xs, nx, sizeof(double), dbl_asc);
    slab_edge_t *ep = (slab_edge_t *)malloc(sizeof(slab_edge_t) * n), *eq = (slab_edge_t *)malloc(sizeof(slab_edge_t) * m);
    double area = 0.0;
    for (int s = 0; s + 1 < nx; ++s) {
        const double xa = xs[s], xb = xs[s + 1];
        if (!(xb > xa)) continue;
        const int kp = slab_collect(P, n, xa, xb, ep), kq = slab_collect(Q, m, xa, xb, eq);
        for (int a = 0; a + 1 < kp; a += 2)
            for (int b = 0; b + 1 < kq; b += 2) {
                /* interval a of P = [ep[a], ep[a+1]], interval b of Q likewise; their bounds are lines that do not
                   cross inside the slab, so max/min of the bounds is decided at mid-slab */
                const double lo = ep[a].ym > eq[b].ym ? ep[a].ym : eq[b].ym;
                const double hi = ep[a + 1].ym < eq[b + 1].ym ? ep[a + 1].ym : eq[b + 1].ym;
                if (hi > lo) area += (hi - lo) * (xb - xa);
            }
    }
    free(xs); free(ep); free(eq);
    return area;
}

/* Greedy cross-tile merge (tools/nuclei_merge.py:62-174).
 *   xy      : concatenated rings, doubles [sum V, 2]
 *   voff    : [N+1] ring offsets into xy (in vertices)
 *   score   : [N] doubles (properties.score)
 *   strategy: 0 'probability', 1 'area'
 *   out_ids : kept ORIGINAL indices, ascending score-rank order of the kept
 *             rows (row r of the returned frame, i.e. nuclei_id r)
 * returns number kept.  Rank order = score descending, ties -> lower index. */
typedef struct { double s; int64_t i; } dsc_t;
static int dsc_desc(const void *a, const void *b) {
    const dsc_t *x = (const dsc_t *)a, *y = (const dsc_t *)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return x->i < y->i ? -1 : (x->i > y->i ? 1 : 0);
}
static int i64_asc(const void *a, const void *b) {
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* optional probe of the decisive pairs (the pairs the greedy loop evaluates): min |IoU - thr| and the pairs closer to
 * the threshold than eps (original indices, query first) */
typedef struct { double eps; int64_t cap, n; int64_t *pairs; double min_margin; } merge_probe_t;
static int64_t merge_impl(const double *xy, const int64_t *voff, const double *score, int64_t N,
                          double thr, int strategy, int64_t *out_ids, int naive, int slab, merge_probe_t *probe) {
    if (probe) { probe->min_margin = 1e300; probe->n = 0; }
    if (N <= 0) return 0;
    dsc_t *ord = (dsc_t *)malloc(N * sizeof(dsc_t));
    for (int64_t i = 0; i < N; ++i) { ord[i].s = score[i]; ord[i].i = i; }
    qsort(ord, N, sizeof(dsc_t), dsc_desc);
    /* envelopes in rank order */
    double *bx0 = malloc(N * sizeof(double)), *by0 = malloc(N * sizeof(double));
    double *bx1 = malloc(N * sizeof(double)), *by1 = malloc(N * sizeof(double));
    double *ar = malloc(N * sizeof(double));
    double gx0 = 1e300, gy0 = 1e300, gx1 = -1e300, gy1 = -1e300, maxext = 1.0;
    for (int64_t r = 0; r < N; ++r) {
        const int64_t i = ord[r].i;
        const double *p = xy + 2 * voff[i];
        const int V = (int)(voff[i + 1] - voff[i]);
        double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300;
        for (int v = 0; v < V; ++v) {
            if (p[2 * v] < x0) x0 = p[2 * v]; if (p[2 * v] > x1) x1 = p[2 * v];
            if (p[2 * v + 1] < y0) y0 = p[2 * v + 1]; if (p[2 * v + 1] > y1) y1 = p[2 * v + 1];
        }
        if (V == 0) { x0 = y0 = x1 = y1 = 0; }
        bx0[r] = x0; by0[r] = y0; bx1[r] = x1; by1[r] = y1;
        ar[r] = fabs(oracle_poly_area2(p, V)) * 0.5;
        if (x0 < gx0) gx0 = x0; if (y0 < gy0) gy0 = y0; if (x1 > gx1) gx1 = x1; if (y1 > gy1) gy1 = y1;
        if (x1 - x0 > maxext) maxext = x1 - x0;
        if (y1 - y0 > maxext) maxext = y1 - y0;
    }
    /* uniform grid on envelope lower-left corner; cell >= max extent so all
       envelope-intersecting partners of a box lie in the 3x3 neighbourhood */
    double cell = maxext;
    while ((floor((gx1 - gx0) / cell) + 1.0) * (floor((gy1 - gy0) / cell) + 1.0) > 4.0 * (double)N + 1024.0)
        cell *= 2.0;
    const double cwx = cell, chy = cell;
    const int64_t nx = (int64_t)floor((gx1 - gx0) / cwx) + 1, ny = (int64_t)floor((gy1 - gy0) / chy) + 1;
    int64_t *cstart = calloc(nx * ny + 1, sizeof(int64_t));
    int64_t *cellof = malloc(N * sizeof(int64_t));
    for (int64_t r = 0; r < N; ++r) {
        int64_t cx = (int64_t)floor((bx0[r] - gx0) / cwx), cy = (int64_t)floor((by0[r] - gy0) / chy);
        if (cx >= nx) cx = nx - 1; if (cy >= ny) cy = ny - 1;
        cellof[r] = cy * nx + cx;
        cstart[cellof[r] + 1]++;
    }
    for (int64_t c = 0; c < nx * ny; ++c) cstart[c + 1] += cstart[c];
    int64_t *fill = malloc((nx * ny + 1) * sizeof(int64_t));
    memcpy(fill, cstart, (nx * ny + 1) * sizeof(int64_t));
    int64_t *items = malloc(N * sizeof(int64_t)); /* ranks, ascending within a cell */
    for (int64_t r = 0; r < N; ++r) items[fill[cellof[r]]++] = r;
    uint8_t *iter = calloc(N, 1); /* iterated_cells */
    int64_t *kept_rank = malloc(N * sizeof(int64_t));
    int64_t k = 0;
    for (int64_t r = 0; r < N; ++r) {
        if (iter[r]) continue;
        const int64_t qi = ord[r].i;
        const double *qp = xy + 2 * voff[qi];
        const int qV = (int)(voff[qi + 1] - voff[qi]);
        const int64_t cx = cellof[r] % nx, cy = cellof[r] / nx;
        int64_t best = -1; double best_area = -1.0;
#define MERGE_TRY(s_) do { const int64_t s = (s_);                                                                        \
            if (s == r || iter[s]) break;                                                                               \
            if (bx0[s] > bx1[r] || bx1[s] < bx0[r] || by0[s] > by1[r] || by1[s] < by0[r]) break;                        \
            const int64_t si = ord[s].i;                                                                                \
            const double inter = slab ? oracle_poly_inter_area_slab(qp, qV, xy + 2 * voff[si], (int)(voff[si + 1] - voff[si])) \
                                      : oracle_poly_inter_area(qp, qV, xy + 2 * voff[si], (int)(voff[si + 1] - voff[si]));    \
            const double iou = inter / (ar[r] + ar[s] - inter);                                                         \
            if (probe) {                                                                                                \
                const double mg = fabs(iou - thr);                                                                      \
                if (mg < probe->min_margin) probe->min_margin = mg;                                                     \
                if (mg < probe->eps && probe->n < probe->cap) {                                                         \
                    probe->pairs[2 * probe->n] = ord[r].i; probe->pairs[2 * probe->n + 1] = si; probe->n++; }           \
            }                                                                                                           \
            if (iou > thr) {                                                                                            \
                iter[s] = 1;                                                                                            \
                if (ar[s] > best_area || (ar[s] == best_area && s < best)) { best_area = ar[s]; best = s; }             \
            } } while (0)
        if (naive) { /* every polygon whose envelope intersects the query's: what STRtree.query returns, found the O(N) way */
            for (int64_t s2 = 0; s2 < N; ++s2) MERGE_TRY(s2);
        } else
        for (int64_t yy = cy - 1; yy <= cy + 1; ++yy) {
            if (yy < 0 || yy >= ny) continue;
            for (int64_t xx = cx - 1; xx <= cx + 1; ++xx) {
                if (xx < 0 || xx >= nx) continue;
                const int64_t c = yy * nx + xx;
                for (int64_t q = cstart[c]; q < cstart[c + 1]; ++q) MERGE_TRY(items[q]);
            }
        }
#undef MERGE_TRY
        kept_rank[k++] = (strategy == 1 && best >= 0) ? best : r;
        iter[r] = 1;
    }
    /* merged_cells.index.isin(merged_idx) keeps rank order; duplicates impossible */
    qsort(kept_rank, k, sizeof(int64_t), i64_asc);
    for (int64_t q = 0; q < k; ++q) out_ids[q] = ord[kept_rank[q]].i;
    free(ord); free(bx0); free(by0); free(bx1); free(by1); free(ar); free(cstart); free(cellof);
    free(fill); free(items); free(iter); free(kept_rank);
    return k;
}
API int64_t oracle_merge(const double *xy, const int64_t *voff, const double *score, int64_t N,
                         double thr, int strategy, int64_t *out_ids) {
    return merge_impl(xy, voff, score, N, thr, strategy, out_ids, 0, 0, NULL);
}
/* flags bit 0: O(N^2) candidate search (every envelope pair) instead of the grid; bit 1: slab-decomposition area.
 * min_margin (may be NULL): min |IoU - thr| over the pairs the greedy loop evaluated, i.e. the pairs whose decision
 * shapes the result -- pairs closer than 1e-9 to the threshold are outside the parity contract (an integer-vertex IoU
 * can equal 1/20 exactly).  Cross-checks of the fast oracle, used by tests/ only. */
API int64_t oracle_merge_check(const double *xy, const int64_t *voff, const double *score, int64_t N,
                               double thr, int strategy, int64_t *out_ids, int flags, double *min_margin,
                               double eps, int64_t *tie_pairs, int64_t tie_cap, int64_t *num_ties) {
    merge_probe_t pr = {eps, tie_pairs ? tie_cap : 0, 0, tie_pairs, 1e300};
    const int64_t k = merge_impl(xy, voff, score, N, thr, strategy, out_ids, flags & 1, (flags >> 1) & 1, &pr);
    if (min_margin) *min_margin = pr.min_margin;
    if (num_ties) *num_ties = pr.n;
    return k;
}

/* ------------------------------------------------------------------------ */
/* A3. paste: ATen grid_sampler_2d (bilinear, zeros, align_corners=False)    */
/* applied to the grid _do_paste_mask builds (fcn_mask_head.py:344-412),     */
/* full-image form.  probs [N,mh,mw] fp32, boxes [N,4] fp32 -> out [N,H,W].  */
/* Scalar restatement of the generic ATen formula; the Python oracle         */
/* (oracle/cpu.py:paste_masks) drives torch's own CPU grid_sample instead;   */
/* this C copy exists so that the CPU baseline has a torch-free timing.      */
/* ------------------------------------------------------------------------ */
typedef struct { const float *probs, *boxes; int mh, mw, H, W; float *out; } ps_ctx_t;
static void ps_body(int64_t n, void *p) {
    ps_ctx_t *c = (ps_ctx_t *)p;
    const float *probs = c->probs, *boxes = c->boxes;
    const int mh = c->mh, mw = c->mw, H = c->H, W = c->W;
    float *out = c->out;
    {
        const float x0 = boxes[4 * n], y0 = boxes[4 * n + 1], x1 = boxes[4 * n + 2], y1 = boxes[4 * n + 3];
        const float *m = probs + (size_t)n * mh * mw;
        float *o = out + (size_t)n * H * W;
        for (int y = 0; y < H; ++y) {
            float gy = ((float)y + 0.5f - y0) / (y1 - y0) * 2.0f - 1.0f;
            if (isinf(gy)) gy = 0.f;
            const float iy = ((gy + 1.f) * (float)mh - 1.f) / 2.f;
            const float fy = floorf(iy);
            const int iy0 = (int)fy, iy1 = iy0 + 1;
            const float wy1 = iy - fy, wy0 = (fy + 1.f) - iy;
            for (int x = 0; x < W; ++x) {
                float gx = ((float)x + 0.5f - x0) / (x1 - x0) * 2.0f - 1.0f;
                if (isinf(gx)) gx = 0.f;
                const float ix = ((gx + 1.f) * (float)mw - 1.f) / 2.f;
                const float fx = floorf(ix);
                const int ix0 = (int)fx, ix1 = ix0 + 1;
                const float wx1 = ix - fx, wx0 = (fx + 1.f) - ix;
                float v = 0.f;
                if (iy0 >= 0 && iy0 < mh && ix0 >= 0 && ix0 < mw) v += m[iy0 * mw + ix0] * (wx0 * wy0);
                if (iy0 >= 0 && iy0 < mh && ix1 >= 0 && ix1 < mw) v += m[iy0 * mw + ix1] * (wx1 * wy0);
                if (iy1 >= 0 && iy1 < mh && ix0 >= 0 && ix0 < mw) v += m[iy1 * mw + ix0] * (wx0 * wy1);
                if (iy1 >= 0 && iy1 < mh && ix1 >= 0 && ix1 < mw) v += m[iy1 * mw + ix1] * (wx1 * wy1);
                o[(size_t)y * W + x] = v;
            }
        }
    }
}
API int oracle_paste(const float *probs, const float *boxes, int N, int mh, int mw, int H, int W,
                     float *out, int nthreads) {
    ps_ctx_t c = {probs, boxes, mh, mw, H, W, out};
    parallel_for(N, 4, nthreads, ps_body, &c);
    return 0;
}

API int oracle_abi_version(void) { return 1; }

/* ------------------------------------------------------------------------------------------------------------------
 * mask2inst (/root/reference/tools/infer_wsi.py:51-54): cv2.findContours(mask, RETR_TREE, CHAIN_APPROX_SIMPLE)[0][0].
 * OpenCV (opencv-python, requirements.txt; not vendored) implements Suzuki & Abe 1985 "Topological structural analysis
 * of digitized binary images by border following"; this is a restatement of that published algorithm with explicit
 * border ids and parents:
 *   - raster scan, LNBD = last border met on the row; outer border starts at f==1 with a 0 on its left, hole border at
 *     f>=1 with a 0 on its right; parent from the (type of new border, type of LNBD) table of the paper;
 *   - border following: clockwise search for the first neighbour, then counter-clockwise sweeps; a pixel whose right
 *     neighbour was examined as 0 gets -NBD, otherwise NBD if still 1;
 *   - CHAIN_APPROX_SIMPLE keeps a border point only where the chain direction changes;
 *   - OpenCV returns the tree in pre-order with siblings in reverse order of discovery, so contour [0] is the LAST
 *     outer border found whose parent is the frame.
 * Pinned against cv2 itself (tests/golden/contours.npz made with cv2 4.13 here; live comparison when cv2 imports).
 * Returns the number of points of contour [0] (0 for an empty mask); writes at most cap (x, y) pairs. */
static const int C_DX[8] = {1, 1, 0, -1, -1, -1, 0, 1};
static const int C_DY[8] = {0, -1, -1, -1, 0, 1, 1, 1};

API int oracle_contour0(const uint8_t *mask, int h, int w, int approx_simple, int32_t *out_xy, int cap) {
    const int W2 = w + 2;
    int32_t *f = (int32_t *)calloc((size_t)(h + 2) * W2, sizeof(int32_t));
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) f[(y + 1) * W2 + x + 1] = mask[(size_t)y * w + x] ? 1 : 0;
    int nb_cap = 64, nbd = 1;
    uint8_t *is_hole = (uint8_t *)malloc(nb_cap);
    int32_t *parent = (int32_t *)malloc(sizeof(int32_t) * nb_cap);
    is_hole[1] = 1; parent[1] = 0;                   /* the frame acts as a hole border */
    int best_i = -1, best_j = -1;                    /* start of the last outer border whose parent is the frame */
    for (int i = 1; i <= h; ++i) {
        int lnbd = 1;
        for (int j = 1; j <= w; ++j) {
            int32_t *p0 = f + i * W2 + j;
            if (*p0 == 0) continue;
            int hole = -1;
            if (*p0 == 1 && p0[-1] == 0) hole = 0;
            else if (*p0 >= 1 && p0[1] == 0) { hole = 1; if (*p0 > 1) lnbd = *p0; }
            if (hole >= 0) {
                ++nbd;
                if (nbd >= nb_cap) {
                    nb_cap *= 2;
                    is_hole = (uint8_t *)realloc(is_hole, nb_cap);
                    parent = (int32_t *)realloc(parent, sizeof(int32_t) * nb_cap);
                }
                is_hole[nbd] = (uint8_t)hole;
                parent[nbd] = (is_hole[lnbd] == hole) ? parent[lnbd] : lnbd;
                if (!hole && parent[nbd] == 1) { best_i = i; best_j = j; }
                /* follow the border */
                int s_end = hole ? 0 : 4, s = s_end;
                int32_t *i1;
                do { s = (s - 1) & 7; i1 = p0 + C_DY[s] * W2 + C_DX[s]; } while (*i1 == 0 && s != s_end);
                if (s == s_end) *p0 = -nbd;
                else {
                    int32_t *i3 = p0, *i4;
                    for (;;) {
                        s_end = s;
                        for (;;) { ++s; i4 = i3 + C_DY[s & 7] * W2 + C_DX[s & 7]; if (*i4 != 0) break; }
                        if (s > 8 && s_end < 8) *i3 = -nbd;        /* direction 0 (right) was examined as a 0-pixel */
                        else if (*i3 == 1) *i3 = nbd;
                        if (i4 == p0 && i3 == i1) break;
                        i3 = i4; s = (s + 4) & 7;
                    }
                }
            }
            if (*p0 != 1) lnbd = *p0 < 0 ? -*p0 : *p0;
        }
    }
    int n = 0;
    if (best_i >= 0) {                               /* re-trace the chosen border and emit its points */
        for (size_t k = 0; k < (size_t)(h + 2) * W2; ++k) f[k] = f[k] != 0;
        int32_t *p0 = f + best_i * W2 + best_j, *i1;
        int s_end = 4, s = 4, x = best_j - 1, y = best_i - 1;
        do { s = (s - 1) & 7; i1 = p0 + C_DY[s] * W2 + C_DX[s]; } while (*i1 == 0 && s != s_end);
        if (s == s_end) { if (n < cap) { out_xy[0] = x; out_xy[1] = y; } n = 1; }
        else {
            int32_t *i3 = p0, *i4;
            int prev_s = s ^ 4;
            for (;;) {
                for (;;) { ++s; i4 = i3 + C_DY[s & 7] * W2 + C_DX[s & 7]; if (*i4 != 0) break; }
                s &= 7;
                if (s != prev_s || !approx_simple) { if (n < cap) { out_xy[2 * n] = x; out_xy[2 * n + 1] = y; } ++n; }
                prev_s = s;
                x += C_DX[s]; y += C_DY[s];
                if (i4 == p0 && i3 == i1) break;
                i3 = i4; s = (s + 4) & 7;
            }
        }
    }
    free(f); free(is_hole); free(parent);
    return n;
}
