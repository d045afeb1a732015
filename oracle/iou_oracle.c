/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Never imported, linked or executed by the product (mvxnet_makise_b200/).
 *
 * Plain-C restatement of the label-side native code of the reference extension:
 *   rotated-quad intersection          cpp/voxelutil.cpp:15-93   (sig, cross, area, lineCross, polygon_cut, intersectArea)
 *   bboxOverlap / bboxIntersection     cpp/voxelutil.cpp:95-139
 *   classifyAnchors                    cpp/voxelutil.cpp:141-316 (python caller modules/Calc.py:88-96)
 *
 * Arithmetic: every operation is a separately rounded fp32 operation in the reference's order (the reference is built
 * for baseline x86-64: SSE scalar floats, no FMA contraction) -> compile with -ffp-contract=off.
 *
 * Deliberate, documented differences from the reference text:
 *  1. bboxOverlap / bboxIntersection fill the second quad through `r2[j]` with j = the BOX index (voxelutil.cpp:108,129)
 *     instead of the corner index k: the quad that is evaluated is then a mix of stale corners, and for more than five boxes
 *     the write runs past the 5-element global. SURVEY.md §8f asks for the intended algorithm (corner k); that is what is
 *     restated. The intersection routine itself is pinned against the live reference through the one configuration where
 *     the indexing slip is harmless (tests/test_oracle_iou.py: one box in bboxes2 whose last corner equals the first corner of the
 *     quad left in r2 by a preceding classifyAnchors call).
 *  2. State kept in globals by the reference (r1, r2, the static scratch of polygon_cut) is local here. The only
 *     input-visible consequence: when lineCross meets parallel segments inside polygon_cut (cross products within 1e-6 of each other
 *     but on different sides of the +-1e-6 dead band) the reference consumes a stale scratch vertex; here that vertex is the
 *     slot's previous content within the same quad pair (zero-initialised per pair).
 *  3. classifyAnchors reads anchors(nl+h, nw, ..) unchecked; a ground truth whose start cell (nl, nw) lies outside the anchor
 *     grid is undefined behaviour there. Here such a ground truth yields no entries and is counted in `n_outside`.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct { float x, y; } pt_t;

static const float kEps = 1e-6f;   /* voxelutil.cpp:15 */

static int sig(float d) { return (d > kEps) - (d < -kEps); }   /* :18-20 */

static int pt_eq(pt_t a, pt_t b) { return sig(a.x - b.x) == 0 && sig(a.y - b.y) == 0; }   /* :24-26 */

static float cross3(pt_t o, pt_t a, pt_t b) {   /* :28-30 */
    return (a.x - o.x) * (b.y - o.y) - (b.x - o.x) * (a.y - o.y);
}

static float area_n(pt_t *ps, int n) {   /* :31-38; writes the closing vertex like the reference */
    ps[n] = ps[0];
    float res = 0;
    for (int i = 0; i < n; i++) res += ps[i].x * ps[i + 1].y - ps[i].y * ps[i + 1].x;
    return (float)(res / 2.0);
}

static int line_cross(pt_t a, pt_t b, pt_t c, pt_t d, pt_t *p) {   /* :39-48 */
    float s1 = cross3(a, b, c), s2 = cross3(a, b, d);
    if (sig(s1) == 0 && sig(s2) == 0) return 2;
    if (sig(s2 - s1) == 0) return 0;
    p->x = (c.x * s2 - d.x * s1) / (s2 - s1);
    p->y = (c.y * s2 - d.y * s1) / (s2 - s1);
    return 1;
}

static void polygon_cut(pt_t *p, int *n_io, pt_t a, pt_t b, pt_t *pp) {   /* :50-63 */
    int n = *n_io, m = 0;
    p[n] = p[0];
    for (int i = 0; i < n; i++) {
        if (sig(cross3(a, b, p[i])) > 0) pp[m++] = p[i];
        if (sig(cross3(a, b, p[i])) != sig(cross3(a, b, p[i + 1]))) line_cross(a, b, p[i], p[i + 1], &pp[m++]);
    }
    n = 0;
    for (int i = 0; i < m; i++)
        if (!i || !pt_eq(pp[i], pp[i - 1])) p[n++] = pp[i];
    while (n > 1 && pt_eq(p[n - 1], p[0])) n--;
    *n_io = n;
}

static float tri_intersect(pt_t a, pt_t b, pt_t c, pt_t d, pt_t *pp) {   /* :65-80: signed area of (o,a,b) ∩ (o,c,d) */
    pt_t o = {0.f, 0.f};
    int s1 = sig(cross3(o, a, b)), s2 = sig(cross3(o, c, d));
    if (s1 == 0 || s2 == 0) return 0.0f;
    if (s1 == -1) { pt_t t = a; a = b; b = t; }
    if (s2 == -1) { pt_t t = c; c = d; d = t; }
    pt_t p[10];
    memset(p, 0, sizeof p);
    p[0] = o; p[1] = a; p[2] = b;
    int n = 3;
    polygon_cut(p, &n, o, c, pp);
    polygon_cut(p, &n, c, d, pp);
    polygon_cut(p, &n, d, o, pp);
    float res = fabsf(area_n(p, n));
    if (s1 * s2 == -1) res = -res;
    return res;
}

static void reverse4(pt_t *ps, int n) {
    for (int i = 0, j = n - 1; i < j; i++, j--) { pt_t t = ps[i]; ps[i] = ps[j]; ps[j] = t; }
}

/* :82-93. ps1/ps2 have room for n+1 vertices and are re-oriented IN PLACE like the reference's globals. */
static float quad_intersect(pt_t *ps1, int n1, pt_t *ps2, int n2) {
    pt_t pp[20];
    memset(pp, 0, sizeof pp);
    if (area_n(ps1, n1) < 0) reverse4(ps1, n1);
    if (area_n(ps2, n2) < 0) reverse4(ps2, n2);
    ps1[n1] = ps1[0];
    ps2[n2] = ps2[0];
    float res = 0;
    for (int i = 0; i < n1; i++)
        for (int j = 0; j < n2; j++) res += tri_intersect(ps1[i], ps1[i + 1], ps2[j], ps2[j + 1], pp);
    return res;
}

static void load_quad(pt_t *r, const float *q) {
    for (int k = 0; k < 4; k++) { r[k].x = q[2 * k]; r[k].y = q[2 * k + 1]; }
}

/* bboxOverlap (mode 0, :96-116) / bboxIntersection (mode 1, :118-139), with the corner index fixed (see header). */
void iou_oracle_pairwise(const float *b1, int64_t n, const float *b2, int64_t m, int mode, float *out) {
    for (int64_t i = 0; i < n; i++) {
        pt_t r1[5];
        load_quad(r1, b1 + i * 8);
        float area1 = area_n(r1, 4);   /* signed, before any re-orientation (:104) */
        for (int64_t j = 0; j < m; j++) {
            pt_t r2[5];
            load_quad(r2, b2 + j * 8);
            float area2 = area_n(r2, 4);
            float inter = quad_intersect(r1, 4, r2, 4);   /* r1 keeps its re-orientation across j like the global */
            out[i * m + j] = mode == 0 ? inter / (area1 + area2 - inter) : inter;
        }
    }
}

typedef struct {
    int64_t *pos, *neg, *gi;   /* pos/neg: [cap][3] (l, w, z) rows; gi: [cap] */
    int64_t npos, nneg, cap;
} sink_t;

/* returns 0 = stop scanning (iou < 0.1), 1 = go on. :168-186 */
static int visit(sink_t *s, float iou, float neg_thr, float pos_thr, int64_t a, int64_t b, int64_t z, int64_t i) {
    if (iou < 0.1) return 0;
    if (iou >= pos_thr) {
        if (s->npos < s->cap) { s->pos[3 * s->npos] = a; s->pos[3 * s->npos + 1] = b; s->pos[3 * s->npos + 2] = z; s->gi[s->npos] = i; }
        s->npos++;
        if (s->nneg < s->cap) { s->neg[3 * s->nneg] = a; s->neg[3 * s->nneg + 1] = b; s->neg[3 * s->nneg + 2] = z; }
        s->nneg++;
    } else if (iou >= neg_thr) {
        if (s->nneg < s->cap) { s->neg[3 * s->nneg] = a; s->neg[3 * s->nneg + 1] = b; s->neg[3 * s->nneg + 2] = z; }
        s->nneg++;
    }
    return 1;
}

/* classifyAnchors (:141-316). gts (G,4,2), anchors (L,W,A,4,2), nls/nws int64[G]. counts[0..2] = npos, nneg, n_outside.
 * Entries beyond `cap` are counted, not stored. */
void iou_oracle_classify(const float *gts, int64_t G, const float *anchors, int64_t L, int64_t W, int64_t A, const int64_t *nls,
                         const int64_t *nws, float neg_thr, float pos_thr, int64_t *pos, int64_t *neg, int64_t *gi, int64_t cap,
                         int64_t *counts) {
    sink_t s = {pos, neg, gi, 0, 0, cap};
    int64_t outside = 0;
    pt_t r1[5], r2[5];
    load_quad(r2, anchors);
    const float anchor_area = area_n(r2, 4);   /* :155: the area of anchor (0,0,0) stands for every anchor */
#define ANCH(a, b, z) (anchors + ((((a) * W) + (b)) * A + (z)) * 8)
#define IOU_AT(a, b, z) (load_quad(r2, ANCH(a, b, z)), inter = quad_intersect(r1, 4, r2, 4), inter / (gt_area + anchor_area - inter))
    for (int64_t i = 0; i < G; i++) {
        const int64_t nl = nls[i], nw = nws[i];
        if (nl < 0 || nl >= L || nw < 0 || nw >= W) { outside++; continue; }
        load_quad(r1, gts + i * 8);
        const float gt_area = area_n(r1, 4);
        float inter;
        for (int64_t z = 0; z < A; z++) {
            for (int phase = 0; phase < 2; phase++) {   /* h = 0, 1, 2, ... then h = -1, -2, ... */
                for (int64_t h = phase == 0 ? 0 : -1; phase == 0 ? nl + h < L : nl + h >= 0; h += phase == 0 ? 1 : -1) {
                    if (!visit(&s, IOU_AT(nl + h, nw, z), neg_thr, pos_thr, nl + h, nw, z, i)) break;
                    for (int64_t v = 1; nw + v < W; v++)
                        if (!visit(&s, IOU_AT(nl + h, nw + v, z), neg_thr, pos_thr, nl + h, nw + v, z, i)) break;
                    for (int64_t v = -1; nw + v >= 0; v--)
                        if (!visit(&s, IOU_AT(nl + h, nw + v, z), neg_thr, pos_thr, nl + h, nw + v, z, i)) break;
                }
            }
        }
    }
#undef IOU_AT
#undef ANCH
    counts[0] = s.npos;
    counts[1] = s.nneg;
    counts[2] = outside;
}
