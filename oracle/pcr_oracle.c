/*
 * pcr_oracle.c -- CPU oracle for the pcrs KNN hot path.  TEST INFRASTRUCTURE ONLY
 * (see pcr_oracle.h for the rules and the parity status).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -msse2 -mfpmath=sse -shared -fPIC -pthread
 * (oracle/Makefile).  -ffp-contract=off matters: Rust never fuses a*b+c, so neither may we.
 */
#include "pcr_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* small helpers                                                                              */
/* ------------------------------------------------------------------------------------------ */

static inline uint32_t f32_bits(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}
static inline float bits_f32(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static inline int finite3(float a, float b, float c) { return isfinite(a) && isfinite(b) && isfinite(c); }

/* kiddo SquaredEuclidean on [f32;3]: sum over axes of (a-b)^2, accumulated x -> y -> z,
 * every operation rounded to f32, no FMA. */
static inline float dist2(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    float s = dx * dx;
    s = s + dy * dy;
    s = s + dz * dz;
    return s;
}

/* (d^2, index) total order packed in one u64: d^2 >= 0 so its bit pattern is monotone. */
static inline uint64_t make_key(float d2, uint32_t idx) { return ((uint64_t)f32_bits(d2) << 32) | idx; }
static inline float key_d2(uint64_t k) { return bits_f32((uint32_t)(k >> 32)); }
static inline uint32_t key_idx(uint64_t k) { return (uint32_t)k; }

/* bounded max-heap of keys (the k best so far) */
typedef struct {
    uint64_t *a;
    size_t n, cap;
} kheap;

static inline void heap_push(kheap *h, uint64_t key) {
    if (h->n < h->cap) {
        size_t i = h->n++;
        h->a[i] = key;
        while (i > 0) {
            size_t p = (i - 1) / 2;
            if (h->a[p] >= h->a[i]) break;
            uint64_t t = h->a[p];
            h->a[p] = h->a[i];
            h->a[i] = t;
            i = p;
        }
    } else if (key < h->a[0]) {
        size_t i = 0;
        h->a[0] = key;
        for (;;) {
            size_t l = 2 * i + 1, r = l + 1, m = i;
            if (l < h->n && h->a[l] > h->a[m]) m = l;
            if (r < h->n && h->a[r] > h->a[m]) m = r;
            if (m == i) break;
            uint64_t t = h->a[m];
            h->a[m] = h->a[i];
            h->a[i] = t;
            i = m;
        }
    }
}
/* worst accepted d^2 (INF while the heap is not full) */
static inline float heap_worst_d2(const kheap *h) { return h->n < h->cap ? INFINITY : key_d2(h->a[0]); }

static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}
static int cmp_u32(const void *a, const void *b) {
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* emit heap content ascending; dist = sqrt(d^2) (kdtree.rs:76) */
static size_t heap_emit(kheap *h, uint32_t *idx, float *dist) {
    qsort(h->a, h->n, sizeof(uint64_t), cmp_u64);
    for (size_t i = 0; i < h->n; i++) {
        if (idx) idx[i] = key_idx(h->a[i]);
        if (dist) dist[i] = sqrtf(key_d2(h->a[i]));
    }
    return h->n;
}

/* parallel-for over [0,n) in contiguous static chunks (mirrors rayon's role, not its schedule) */
typedef void (*range_fn)(size_t lo, size_t hi, void *ctx);
typedef struct {
    range_fn fn;
    void *ctx;
    size_t lo, hi;
} pf_task;
static void *pf_run(void *p) {
    pf_task *t = (pf_task *)p;
    t->fn(t->lo, t->hi, t->ctx);
    return NULL;
}
static void parallel_for(size_t n, int threads, range_fn fn, void *ctx) {
    if (threads <= 1 || n < 2) {
        fn(0, n, ctx);
        return;
    }
    if ((size_t)threads > n) threads = (int)n;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    pf_task *tk = (pf_task *)malloc(sizeof(pf_task) * threads);
    for (int i = 0; i < threads; i++) {
        tk[i].fn = fn;
        tk[i].ctx = ctx;
        tk[i].lo = n * (size_t)i / threads;
        tk[i].hi = n * (size_t)(i + 1) / threads;
        pthread_create(&th[i], NULL, pf_run, &tk[i]);
    }
    for (int i = 0; i < threads; i++) pthread_join(th[i], NULL);
    free(th);
    free(tk);
}

/* ------------------------------------------------------------------------------------------ */
/* kd-tree (stands in for kiddo ImmutableKdTree<f32,u32,3,32>; bucket size 32)                 */
/* ------------------------------------------------------------------------------------------ */

#define ORC_BUCKET 32

typedef struct {
    float split;
    int32_t axis;          /* -1 = leaf */
    uint32_t left, right;  /* children (inner) or [begin,end) into the point arrays (leaf) */
} orc_node;

struct orc_tree {
    size_t num_points; /* cloud length, kdtree.rs:16 */
    size_t m;          /* indexed (finite) points */
    float *px, *py, *pz;
    uint32_t *pidx;
    orc_node *nodes;
    size_t n_nodes, cap_nodes;
};

typedef struct {
    const float *c[3];
    uint32_t *perm;
} build_ctx;

static inline float coord(const build_ctx *b, uint32_t i, int axis) { return b->c[axis][i]; }

/* quickselect on perm[lo,hi) by coordinate (ties by index so the build is deterministic) */
static inline int pt_less(const build_ctx *b, int axis, uint32_t i, uint32_t j) {
    float a = coord(b, i, axis), c = coord(b, j, axis);
    return a < c || (a == c && i < j);
}
static void nth_element(const build_ctx *b, int axis, size_t lo, size_t hi, size_t nth) {
    uint32_t *p = b->perm;
    while (hi - lo > 1) {
        /* Hoare partition, pivot = middle element (never the last one, so j < hi-1) */
        uint32_t piv = p[lo + (hi - 1 - lo) / 2];
        ptrdiff_t i = (ptrdiff_t)lo - 1, j = (ptrdiff_t)hi;
        for (;;) {
            do i++; while (pt_less(b, axis, p[i], piv));
            do j--; while (pt_less(b, axis, piv, p[j]));
            if (i >= j) break;
            uint32_t t = p[i];
            p[i] = p[j];
            p[j] = t;
        }
        /* now [lo..j] <= piv <= [j+1..hi) */
        if ((ptrdiff_t)nth <= j) hi = (size_t)j + 1;
        else lo = (size_t)j + 1;
    }
}

static uint32_t new_node(orc_tree *t) {
    if (t->n_nodes == t->cap_nodes) {
        t->cap_nodes = t->cap_nodes ? t->cap_nodes * 2 : 64;
        t->nodes = (orc_node *)realloc(t->nodes, t->cap_nodes * sizeof(orc_node));
    }
    return (uint32_t)t->n_nodes++;
}

static uint32_t build_rec(orc_tree *t, const build_ctx *b, size_t lo, size_t hi) {
    uint32_t id = new_node(t);
    if (hi - lo <= ORC_BUCKET) {
        t->nodes[id].axis = -1;
        t->nodes[id].left = (uint32_t)lo;
        t->nodes[id].right = (uint32_t)hi;
        t->nodes[id].split = 0.f;
        return id;
    }
    /* widest axis */
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t i = lo; i < hi; i++)
        for (int a = 0; a < 3; a++) {
            float v = coord(b, b->perm[i], a);
            if (v < mn[a]) mn[a] = v;
            if (v > mx[a]) mx[a] = v;
        }
    int axis = 0;
    float ext = mx[0] - mn[0];
    for (int a = 1; a < 3; a++)
        if (mx[a] - mn[a] > ext) {
            ext = mx[a] - mn[a];
            axis = a;
        }
    size_t mid = lo + (hi - lo) / 2;
    nth_element(b, axis, lo, hi, mid);
    float split = coord(b, b->perm[mid], axis);
    uint32_t l = build_rec(t, b, lo, mid);
    uint32_t r = build_rec(t, b, mid, hi);
    /* left subtree: coords <= split, right subtree: coords >= split */
    t->nodes[id].axis = axis;
    t->nodes[id].split = split;
    t->nodes[id].left = l;
    t->nodes[id].right = r;
    return id;
}

orc_tree *orc_tree_build(const float *x, const float *y, const float *z, size_t n) {
    orc_tree *t = (orc_tree *)calloc(1, sizeof(orc_tree));
    t->num_points = n;
    uint32_t *perm = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    size_t m = 0;
    for (size_t i = 0; i < n; i++)
        if (finite3(x[i], y[i], z[i])) perm[m++] = (uint32_t)i;
    t->m = m;
    if (m > 0) {
        build_ctx b = {{x, y, z}, perm};
        build_rec(t, &b, 0, m);
    }
    t->px = (float *)malloc(sizeof(float) * (m ? m : 1));
    t->py = (float *)malloc(sizeof(float) * (m ? m : 1));
    t->pz = (float *)malloc(sizeof(float) * (m ? m : 1));
    t->pidx = perm;
    for (size_t i = 0; i < m; i++) {
        t->px[i] = x[perm[i]];
        t->py[i] = y[perm[i]];
        t->pz[i] = z[perm[i]];
    }
    return t;
}

void orc_tree_free(orc_tree *t) {
    if (!t) return;
    free(t->px);
    free(t->py);
    free(t->pz);
    free(t->pidx);
    free(t->nodes);
    free(t);
}

size_t orc_tree_len(const orc_tree *t) { return t->num_points; }

/* Exact pruning in f32: rounding is monotone, so every point p beyond the split plane has
 * computed d^2(p) >= fl(fl(q-split)^2).  A subtree is skipped only if that is strictly greater
 * than the current worst d^2 (an equal d^2 with a lower index would still win the tie-break). */
static void knn_rec(const orc_tree *t, uint32_t id, const float q[3], kheap *h) {
    const orc_node *nd = &t->nodes[id];
    if (nd->axis < 0) {
        for (uint32_t i = nd->left; i < nd->right; i++) {
            float d2 = dist2(q[0], q[1], q[2], t->px[i], t->py[i], t->pz[i]);
            heap_push(h, make_key(d2, t->pidx[i]));
        }
        return;
    }
    float diff = q[nd->axis] - nd->split;
    uint32_t near = diff < 0.f ? nd->left : nd->right;
    uint32_t far = diff < 0.f ? nd->right : nd->left;
    knn_rec(t, near, q, h);
    float pd2 = diff * diff;
    if (pd2 <= heap_worst_d2(h)) knn_rec(t, far, q, h);
}

size_t orc_tree_knn(const orc_tree *t, const float q[3], size_t k, uint32_t *idx, float *dist) {
    /* kdtree.rs:65 / :88 */
    if (k == 0 || t->num_points == 0 || !finite3(q[0], q[1], q[2]) || t->m == 0) return 0;
    size_t kk = k < t->m ? k : t->m;
    uint64_t stackbuf[64];
    kheap h = {kk <= 64 ? stackbuf : (uint64_t *)malloc(sizeof(uint64_t) * kk), 0, kk};
    knn_rec(t, 0, q, &h);
    size_t r = heap_emit(&h, idx, dist);
    if (h.a != stackbuf) free(h.a);
    return r;
}

size_t orc_knn_brute(const float *x, const float *y, const float *z, size_t n, const float q[3],
                     size_t k, uint32_t *idx, float *dist) {
    if (k == 0 || n == 0 || !finite3(q[0], q[1], q[2])) return 0;
    size_t m = 0;
    for (size_t i = 0; i < n; i++) m += finite3(x[i], y[i], z[i]) ? 1 : 0;
    if (m == 0) return 0;
    size_t kk = k < m ? k : m;
    kheap h = {(uint64_t *)malloc(sizeof(uint64_t) * kk), 0, kk};
    for (size_t i = 0; i < n; i++) {
        if (!finite3(x[i], y[i], z[i])) continue;
        heap_push(&h, make_key(dist2(q[0], q[1], q[2], x[i], y[i], z[i]), (uint32_t)i));
    }
    size_t r = heap_emit(&h, idx, dist);
    free(h.a);
    return r;
}

typedef struct {
    const orc_tree *t;
    const float *qx, *qy, *qz;
    size_t k;
    uint32_t *idx;
    float *dist;
    uint32_t *counts;
    float radius;
} knn_batch_ctx;

static void knn_batch_range(size_t lo, size_t hi, void *p) {
    knn_batch_ctx *c = (knn_batch_ctx *)p;
    for (size_t i = lo; i < hi; i++) {
        float q[3] = {c->qx[i], c->qy[i], c->qz[i]};
        uint32_t *ri = c->idx ? c->idx + i * c->k : NULL;
        float *rd = c->dist ? c->dist + i * c->k : NULL;
        size_t r = orc_tree_knn(c->t, q, c->k, ri, rd);
        for (size_t j = r; j < c->k; j++) {
            if (ri) ri[j] = UINT32_MAX;
            if (rd) rd[j] = INFINITY;
        }
        if (c->counts) c->counts[i] = (uint32_t)r;
    }
}

void orc_knn_batch(const orc_tree *t, const float *qx, const float *qy, const float *qz, size_t nq,
                   size_t k, uint32_t *idx, float *dist, uint32_t *counts, int threads) {
    knn_batch_ctx c = {t, qx, qy, qz, k, idx, dist, counts, 0.f};
    parallel_for(nq, threads, knn_batch_range, &c);
}

/* radius: kdtree.rs:114-127.  r^2 rounded in f32; kiddo's strict `<` on the epsilon-inflated
 * radius followed by `<= radius_sq` is exactly `d^2 <= radius_sq` (the inflation only widens). */
typedef struct {
    uint32_t *idx;
    size_t cap, n;
} rad_out;

static void radius_rec(const orc_tree *t, uint32_t id, const float q[3], float r2, rad_out *o) {
    const orc_node *nd = &t->nodes[id];
    if (nd->axis < 0) {
        for (uint32_t i = nd->left; i < nd->right; i++) {
            float d2 = dist2(q[0], q[1], q[2], t->px[i], t->py[i], t->pz[i]);
            if (d2 <= r2) {
                if (o->idx && o->n < o->cap) o->idx[o->n] = t->pidx[i];
                o->n++;
            }
        }
        return;
    }
    float diff = q[nd->axis] - nd->split;
    uint32_t near = diff < 0.f ? nd->left : nd->right;
    uint32_t far = diff < 0.f ? nd->right : nd->left;
    radius_rec(t, near, q, r2, o);
    if (diff * diff <= r2) radius_rec(t, far, q, r2, o);
}

static int radius_guard(const orc_tree *t, const float q[3], float radius) {
    /* kdtree.rs:106-112 */
    return !(t->num_points == 0 || !(radius > 0.0f) || !isfinite(radius) || !finite3(q[0], q[1], q[2]) ||
             t->m == 0);
}

size_t orc_tree_radius_search(const orc_tree *t, const float q[3], float radius, uint32_t *idx,
                              size_t cap) {
    if (!radius_guard(t, q, radius)) return 0;
    float r2 = radius * radius;
    rad_out o = {idx, cap, 0};
    radius_rec(t, 0, q, r2, &o);
    if (idx) qsort(idx, o.n < cap ? o.n : cap, sizeof(uint32_t), cmp_u32); /* kdtree.rs:132 */
    return o.n;
}

size_t orc_tree_radius_count(const orc_tree *t, const float q[3], float radius) {
    if (!radius_guard(t, q, radius)) return 0;
    rad_out o = {NULL, 0, 0};
    radius_rec(t, 0, q, radius * radius, &o);
    return o.n;
}

static void radius_count_range(size_t lo, size_t hi, void *p) {
    knn_batch_ctx *c = (knn_batch_ctx *)p;
    for (size_t i = lo; i < hi; i++) {
        float q[3] = {c->qx[i], c->qy[i], c->qz[i]};
        c->counts[i] = (uint32_t)orc_tree_radius_count(c->t, q, c->radius);
    }
}

void orc_radius_count_batch(const orc_tree *t, const float *qx, const float *qy, const float *qz,
                            size_t nq, float radius, uint32_t *counts, int threads) {
    knn_batch_ctx c = {t, qx, qy, qz, 0, NULL, NULL, counts, radius};
    parallel_for(nq, threads, radius_count_range, &c);
}

/* ------------------------------------------------------------------------------------------ */
/* statistical_outlier_removal  (crates/filters/src/statistical_outlier.rs:4-69)              */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    const orc_tree *t;
    const float *x, *y, *z;
    size_t k;
    float *mean_d;
} sor_ctx;

static void sor_range(size_t lo, size_t hi, void *p) {
    sor_ctx *c = (sor_ctx *)p;
    size_t kk = c->k + 1;
    float *d = (float *)malloc(sizeof(float) * kk);
    for (size_t i = lo; i < hi; i++) {
        float q[3] = {c->x[i], c->y[i], c->z[i]};
        if (!finite3(q[0], q[1], q[2])) { /* :22-24 */
            c->mean_d[i] = INFINITY;
            continue;
        }
        size_t r = orc_tree_knn(c->t, q, kk, NULL, d); /* :25 */
        const float *nd = d;
        size_t cnt = r;
        if (r > 1) { /* :28-32 drop the first (self) */
            nd = d + 1;
            cnt = r - 1;
        }
        if (cnt == 0) { /* :33-35 */
            c->mean_d[i] = INFINITY;
            continue;
        }
        float sum = 0.0f; /* :36 sequential f32, ascending distance order */
        for (size_t j = 0; j < cnt; j++) sum = sum + nd[j];
        c->mean_d[i] = sum / (float)cnt; /* :37 */
    }
    free(d);
}

size_t orc_sor(const float *x, const float *y, const float *z, size_t n, size_t k, float std_mul,
               uint8_t *keep, float *mean_d_out, float *stats, int threads) {
    if (stats) stats[0] = stats[1] = stats[2] = NAN;
    if (n == 0 || k == 0) { /* :5-7 -> empty cloud */
        for (size_t i = 0; i < n; i++) keep[i] = 0;
        return 0;
    }
    if (n == 1) { /* :10-12 -> clone */
        keep[0] = 1;
        if (mean_d_out) mean_d_out[0] = INFINITY;
        return 1;
    }
    orc_tree *t = orc_tree_build(x, y, z, n);
    float *md = mean_d_out ? mean_d_out : (float *)malloc(sizeof(float) * n);
    sor_ctx c = {t, x, y, z, k, md};
    parallel_for(n, threads, sor_range, &c);
    orc_tree_free(t);

    /* :43-60 sequential f32 folds over the finite subset, in index order */
    size_t nf = 0;
    float sum = 0.0f;
    for (size_t i = 0; i < n; i++)
        if (isfinite(md[i])) {
            sum = sum + md[i];
            nf++;
        }
    size_t kept = 0;
    if (nf == 0) { /* :49-51 */
        for (size_t i = 0; i < n; i++) keep[i] = 0;
    } else {
        float nn = (float)nf;
        float gmean = sum / nn;
        float var = 0.0f;
        for (size_t i = 0; i < n; i++)
            if (isfinite(md[i])) {
                float d = md[i] - gmean;
                var = var + d * d; /* powi(2) */
            }
        var = var / nn;
        float sd = sqrtf(var);
        float thr = gmean + std_mul * sd; /* :62 */
        if (stats) {
            stats[0] = gmean;
            stats[1] = sd;
            stats[2] = thr;
        }
        for (size_t i = 0; i < n; i++) { /* :64-66 */
            keep[i] = md[i] <= thr ? 1 : 0;
            kept += keep[i];
        }
    }
    if (!mean_d_out) free(md);
    return kept;
}

/* ------------------------------------------------------------------------------------------ */
/* radius_outlier_removal  (crates/filters/src/radius_outlier.rs:4-18)                        */
/* ------------------------------------------------------------------------------------------ */

size_t orc_ror(const float *x, const float *y, const float *z, size_t n, float radius,
               size_t min_neighbors, uint8_t *keep, int threads) {
    if (n == 0) return 0;
    orc_tree *t = orc_tree_build(x, y, z, n);
    uint32_t *cnt = (uint32_t *)malloc(sizeof(uint32_t) * n);
    orc_radius_count_batch(t, x, y, z, n, radius, cnt, threads);
    size_t kept = 0;
    for (size_t i = 0; i < n; i++) {
        keep[i] = (size_t)cnt[i] >= min_neighbors ? 1 : 0; /* :13 */
        kept += keep[i];
    }
    free(cnt);
    orc_tree_free(t);
    return kept;
}

/* ------------------------------------------------------------------------------------------ */
/* normals  (crates/normals/src/estimate.rs)                                                  */
/* ------------------------------------------------------------------------------------------ */

/* estimate.rs:139-238, f64 internally */
void orc_smallest_eigenvector_3x3(float fa00, float fa01, float fa02, float fa11, float fa12,
                                  float fa22, float out[3]) {
    double a00 = fa00, a01 = fa01, a02 = fa02, a11 = fa11, a12 = fa12, a22 = fa22;
    double m = (a00 + a11 + a22) / 3.0; /* :157 */
    double b00 = a00 - m, b11 = a11 - m, b22 = a22 - m;
    double q = (b00 * (b11 * b22 - a12 * a12) - a01 * (a01 * b22 - a12 * a02) +
                a02 * (a01 * a12 - b11 * a02)) /
               2.0; /* :165-167 */
    double p = (b00 * b00 + b11 * b11 + b22 * b22 + 2.0 * (a01 * a01 + a02 * a02 + a12 * a12)) / 6.0;
    double pp = p > 0.0 ? p : 0.0; /* :172 p.max(0.0) (NaN -> 0.0 like f64::max) */
    if (pp < 1e-30) {              /* :174-177 */
        out[0] = 0.f;
        out[1] = 0.f;
        out[2] = 1.f;
        return;
    }
    double det_ratio = q / (pp * sqrt(pp)); /* :180 */
    if (det_ratio < -1.0) det_ratio = -1.0; /* :181 clamp */
    if (det_ratio > 1.0) det_ratio = 1.0;
    double phi = acos(det_ratio) / 3.0;
    double sqrt_p = sqrt(pp);
    const double FRAC_PI_3 = 1.04719755119659774615421446109316763;
    double eig0 = m + 2.0 * sqrt_p * cos(phi + 2.0 * FRAC_PI_3); /* :186 */
    double eig2 = m + 2.0 * sqrt_p * cos(phi);                   /* :187 */
    double eig1 = 3.0 * m - eig0 - eig2;                         /* :188 */
    double lambda;                                               /* :191-197 smallest |eig| */
    if (fabs(eig0) <= fabs(eig1) && fabs(eig0) <= fabs(eig2)) lambda = eig0;
    else if (fabs(eig1) <= fabs(eig2)) lambda = eig1;
    else lambda = eig2;
    double r00 = a00 - lambda, r11 = a11 - lambda, r22 = a22 - lambda;
    double ex = a01 * a12 - r11 * a02; /* :206-208 row0 x row1 */
    double ey = a02 * a01 - a12 * r00;
    double ez = r00 * r11 - a01 * a01;
    double len2 = ex * ex + ey * ey + ez * ez;
    if (len2 < 1e-30) {
        ex = a01 * r22 - a12 * a02; /* :214-216 row0 x row2 */
        ey = a02 * a02 - r22 * r00;
        ez = r00 * a12 - a01 * a02;
        double len2b = ex * ex + ey * ey + ez * ez;
        if (len2b < 1e-30) {
            ex = r11 * r22 - a12 * a12; /* :221-223 row1 x row2 */
            ey = a12 * a02 - r22 * a01;
            ez = a01 * a12 - r11 * a02;
            double len2c = ex * ex + ey * ey + ez * ez;
            if (len2c < 1e-30) { /* :226-228 */
                out[0] = 0.f;
                out[1] = 0.f;
                out[2] = 1.f;
                return;
            }
            double inv = 1.0 / sqrt(len2c);
            out[0] = (float)(ex * inv);
            out[1] = (float)(ey * inv);
            out[2] = (float)(ez * inv);
            return;
        }
        double inv = 1.0 / sqrt(len2b);
        out[0] = (float)(ex * inv);
        out[1] = (float)(ey * inv);
        out[2] = (float)(ez * inv);
        return;
    }
    double inv = 1.0 / sqrt(len2);
    out[0] = (float)(ex * inv);
    out[1] = (float)(ey * inv);
    out[2] = (float)(ez * inv);
}

typedef struct {
    const orc_tree *t;
    const float *x, *y, *z;
    size_t k;
    const float *vp;
    float *nx, *ny, *nz;
} nrm_ctx;

static void normals_range(size_t lo, size_t hi, void *p) {
    nrm_ctx *c = (nrm_ctx *)p;
    uint32_t *idx = (uint32_t *)malloc(sizeof(uint32_t) * c->k);
    for (size_t i = lo; i < hi; i++) {
        float pt[3] = {c->x[i], c->y[i], c->z[i]};
        size_t r = orc_tree_knn(c->t, pt, c->k, idx, NULL); /* :45 */
        float count = (float)r;
        if (count < 1.0f) { /* :49-51 */
            c->nx[i] = 0.f;
            c->ny[i] = 0.f;
            c->nz[i] = 1.f;
            continue;
        }
        float cx = 0.f, cy = 0.f, cz = 0.f; /* :54-65 */
        for (size_t j = 0; j < r; j++) {
            cx += c->x[idx[j]];
            cy += c->y[idx[j]];
            cz += c->z[idx[j]];
        }
        cx /= count;
        cy /= count;
        cz /= count;
        float c00 = 0.f, c01 = 0.f, c02 = 0.f, c11 = 0.f, c12 = 0.f, c22 = 0.f; /* :68-84 */
        for (size_t j = 0; j < r; j++) {
            float dx = c->x[idx[j]] - cx, dy = c->y[idx[j]] - cy, dz = c->z[idx[j]] - cz;
            c00 += dx * dx;
            c01 += dx * dy;
            c02 += dx * dz;
            c11 += dy * dy;
            c12 += dy * dz;
            c22 += dz * dz;
        }
        float nrm[3];
        orc_smallest_eigenvector_3x3(c00, c01, c02, c11, c12, c22, nrm);
        float nnx = nrm[0], nny = nrm[1], nnz = nrm[2];
        float len = sqrtf(nnx * nnx + nny * nny + nnz * nnz); /* :91 */
        if (len > 1e-10f) {
            nnx /= len;
            nny /= len;
            nnz /= len;
        }
        float vx = c->vp[0] - pt[0], vy = c->vp[1] - pt[1], vz = c->vp[2] - pt[2]; /* :99-101 */
        float dot = nnx * vx + nny * vy + nnz * vz;
        if (dot < 0.0f) {
            nnx = -nnx;
            nny = -nny;
            nnz = -nnz;
        }
        c->nx[i] = nnx;
        c->ny[i] = nny;
        c->nz[i] = nnz;
    }
    free(idx);
}

void orc_normals(const float *x, const float *y, const float *z, size_t n, size_t k,
                 const float viewpoint[3], float *nx, float *ny, float *nz, int threads) {
    if (n == 0 || k == 0) return; /* :25-31 -> empty normals */
    orc_tree *t = orc_tree_build(x, y, z, n);
    nrm_ctx c = {t, x, y, z, k, viewpoint, nx, ny, nz};
    parallel_for(n, threads, normals_range, &c);
    orc_tree_free(t);
}

/* ------------------------------------------------------------------------------------------ */
/* registration  (crates/registration/src/{icp,icp_plane,correspondence}.rs)                  */
/* ------------------------------------------------------------------------------------------ */

static void transform_identity(orc_transform *t) {
    static const float I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    memcpy(t->rotation, I, sizeof(I));
    t->translation[0] = t->translation[1] = t->translation[2] = 0.f;
}

/* icp.rs:39-47: r0*x + r1*y + r2*z + t, left to right */
static inline void apply_point(const orc_transform *t, float x, float y, float z, float o[3]) {
    const float *r = t->rotation;
    o[0] = r[0] * x + r[1] * y + r[2] * z + t->translation[0];
    o[1] = r[3] * x + r[4] * y + r[5] * z + t->translation[1];
    o[2] = r[6] * x + r[7] * y + r[8] * z + t->translation[2];
}

void orc_apply_transform(const float *x, const float *y, const float *z, size_t n,
                         const orc_transform *t, float *ox, float *oy, float *oz) {
    for (size_t i = 0; i < n; i++) {
        float o[3];
        apply_point(t, x[i], y[i], z[i], o);
        ox[i] = o[0];
        oy[i] = o[1];
        oz[i] = o[2];
    }
}

static void mat3_mul(const float *a, const float *b, float *o) { /* o = a*b (nalgebra f32 product) */
    float r[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r[i * 3 + j] = a[i * 3 + 0] * b[0 * 3 + j] + a[i * 3 + 1] * b[1 * 3 + j] + a[i * 3 + 2] * b[2 * 3 + j];
    memcpy(o, r, sizeof(r));
}

/* icp.rs:52-73: R_new = other.R * self.R ; t_new = other.R * self.t + other.t */
void orc_compose(const orc_transform *self, const orc_transform *other, orc_transform *out) {
    orc_transform r;
    mat3_mul(other->rotation, self->rotation, r.rotation);
    for (int i = 0; i < 3; i++)
        r.translation[i] = (other->rotation[i * 3 + 0] * self->translation[0] + other->rotation[i * 3 + 1] * self->translation[1] +
                            other->rotation[i * 3 + 2] * self->translation[2]) +
                           other->translation[i];
    *out = r;
}

typedef struct {
    const float *sx, *sy, *sz;
    const orc_tree *t;
    uint32_t *tgt;
    float *dist;
} corr_ctx;

static void corr_range(size_t lo, size_t hi, void *p) {
    corr_ctx *c = (corr_ctx *)p;
    for (size_t i = lo; i < hi; i++) {
        float q[3] = {c->sx[i], c->sy[i], c->sz[i]};
        uint32_t ti;
        float d;
        size_t r = orc_tree_knn(c->t, q, 1, &ti, &d); /* correspondence.rs:25 */
        c->tgt[i] = r ? ti : UINT32_MAX;
        c->dist[i] = r ? d : INFINITY;
    }
}

size_t orc_find_correspondences(const float *sx, const float *sy, const float *sz, size_t ns,
                                const orc_tree *target, float max_distance, uint32_t *src_idx,
                                uint32_t *tgt_idx, float *dist, int threads) {
    uint32_t *tt = (uint32_t *)malloc(sizeof(uint32_t) * (ns ? ns : 1));
    float *dd = (float *)malloc(sizeof(float) * (ns ? ns : 1));
    corr_ctx c = {sx, sy, sz, target, tt, dd};
    parallel_for(ns, threads, corr_range, &c);
    size_t m = 0;
    for (size_t i = 0; i < ns; i++) { /* correspondence.rs:27-35, in source order */
        if (tt[i] == UINT32_MAX) continue;
        if (dd[i] <= max_distance) {
            src_idx[m] = (uint32_t)i;
            tgt_idx[m] = tt[i];
            dist[m] = dd[i];
            m++;
        }
    }
    free(tt);
    free(dd);
    return m;
}

/* icp.rs:273-282 */
static float compute_rmse(const float *dist, size_t m) {
    if (m == 0) return 0.0f;
    float s = 0.0f;
    for (size_t i = 0; i < m; i++) s = s + dist[i] * dist[i];
    return sqrtf(s / (float)m);
}

/* 3x3 SVD of a (row-major, f64): a = U diag(s) V^T, s descending.  nalgebra is not available;
 * this uses cyclic Jacobi on a^T a.  R = V U^T is unique for non-degenerate H, so the choice of
 * SVD algorithm only moves the result at rounding level. */
static void svd3(const double a[9], double U[9], double s[3], double V[9]) {
    double ata[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double t = 0;
            for (int k = 0; k < 3; k++) t += a[k * 3 + i] * a[k * 3 + j];
            ata[i * 3 + j] = t;
        }
    double v[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = fabs(ata[1]) + fabs(ata[2]) + fabs(ata[5]);
        double diag = fabs(ata[0]) + fabs(ata[4]) + fabs(ata[8]);
        if (off <= 1e-300 || off <= 1e-18 * diag) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                double apq = ata[p * 3 + q];
                if (fabs(apq) < 1e-300) continue;
                double app = ata[p * 3 + p], aqq = ata[q * 3 + q];
                double theta = (aqq - app) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 3; k++) { /* columns p,q of ata */
                    double akp = ata[k * 3 + p], akq = ata[k * 3 + q];
                    ata[k * 3 + p] = c * akp - sn * akq;
                    ata[k * 3 + q] = sn * akp + c * akq;
                }
                for (int k = 0; k < 3; k++) { /* rows p,q */
                    double apk = ata[p * 3 + k], aqk = ata[q * 3 + k];
                    ata[p * 3 + k] = c * apk - sn * aqk;
                    ata[q * 3 + k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < 3; k++) {
                    double vkp = v[k * 3 + p], vkq = v[k * 3 + q];
                    v[k * 3 + p] = c * vkp - sn * vkq;
                    v[k * 3 + q] = sn * vkp + c * vkq;
                }
            }
    }
    int ord[3] = {0, 1, 2};
    double ev[3] = {ata[0], ata[4], ata[8]};
    for (int i = 0; i < 2; i++)
        for (int j = i + 1; j < 3; j++)
            if (ev[ord[j]] > ev[ord[i]]) {
                int t = ord[i];
                ord[i] = ord[j];
                ord[j] = t;
            }
    for (int c = 0; c < 3; c++) {
        s[c] = sqrt(ev[ord[c]] > 0 ? ev[ord[c]] : 0);
        for (int r = 0; r < 3; r++) V[r * 3 + c] = v[r * 3 + ord[c]];
    }
    /* make V right-handed irrelevant; U columns = a v / s */
    double u[3][3];
    int have[3] = {0, 0, 0};
    for (int c = 0; c < 3; c++) {
        double col[3];
        for (int r = 0; r < 3; r++) col[r] = a[r * 3 + 0] * V[0 * 3 + c] + a[r * 3 + 1] * V[1 * 3 + c] + a[r * 3 + 2] * V[2 * 3 + c];
        double nrm = sqrt(col[0] * col[0] + col[1] * col[1] + col[2] * col[2]);
        if (s[0] > 0 && nrm > 1e-12 * s[0] && nrm > 1e-300) {
            for (int r = 0; r < 3; r++) u[c][r] = col[r] / nrm;
            have[c] = 1;
        }
    }
    /* complete a rank-deficient U to an orthonormal basis */
    for (int c = 0; c < 3; c++) {
        if (have[c]) continue;
        int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
        double w[3];
        if (have[c1] && have[c2]) {
            w[0] = u[c1][1] * u[c2][2] - u[c1][2] * u[c2][1];
            w[1] = u[c1][2] * u[c2][0] - u[c1][0] * u[c2][2];
            w[2] = u[c1][0] * u[c2][1] - u[c1][1] * u[c2][0];
        } else {
            /* pick any unit vector orthogonal to the ones we have */
            double best = -1;
            w[0] = w[1] = w[2] = 0;
            for (int e = 0; e < 3; e++) {
                double cand[3] = {0, 0, 0};
                cand[e] = 1;
                for (int o = 0; o < 3; o++)
                    if (have[o]) {
                        double d = cand[0] * u[o][0] + cand[1] * u[o][1] + cand[2] * u[o][2];
                        for (int r = 0; r < 3; r++) cand[r] -= d * u[o][r];
                    }
                double n2 = cand[0] * cand[0] + cand[1] * cand[1] + cand[2] * cand[2];
                if (n2 > best) {
                    best = n2;
                    memcpy(w, cand, sizeof(w));
                }
            }
        }
        double n = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
        for (int r = 0; r < 3; r++) u[c][r] = w[r] / n;
        have[c] = 1;
    }
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) U[r * 3 + c] = u[c][r];
}

static float det3f(const float *m) {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

/* icp.rs:210-270 */
static void rigid_transform_svd(const float *sx, const float *sy, const float *sz, const float *tx,
                                const float *ty, const float *tz, const uint32_t *si,
                                const uint32_t *ti, size_t m, orc_transform *out) {
    if (m == 0) {
        transform_identity(out);
        return;
    }
    float sc[3] = {0, 0, 0}, tc[3] = {0, 0, 0}; /* :221-233 sequential f32 */
    for (size_t c = 0; c < m; c++) {
        sc[0] += sx[si[c]];
        sc[1] += sy[si[c]];
        sc[2] += sz[si[c]];
        tc[0] += tx[ti[c]];
        tc[1] += ty[ti[c]];
        tc[2] += tz[ti[c]];
    }
    float nf = (float)m;
    for (int a = 0; a < 3; a++) {
        sc[a] /= nf;
        tc[a] /= nf;
    }
    float h[9] = {0}; /* :236-244 H = sum (s - sc)(t - tc)^T, f32 */
    for (size_t c = 0; c < m; c++) {
        float sp[3] = {sx[si[c]] - sc[0], sy[si[c]] - sc[1], sz[si[c]] - sc[2]};
        float tp[3] = {tx[ti[c]] - tc[0], ty[ti[c]] - tc[1], tz[ti[c]] - tc[2]};
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) h[i * 3 + j] += sp[i] * tp[j];
    }
    double hd[9], U[9], S[3], V[9];
    for (int i = 0; i < 9; i++) hd[i] = h[i];
    svd3(hd, U, S, V);
    float u[9], vt[9], v[9], ut[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            u[i * 3 + j] = (float)U[i * 3 + j];
            vt[i * 3 + j] = (float)V[j * 3 + i];
        }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            v[i * 3 + j] = vt[j * 3 + i];
            ut[i * 3 + j] = u[j * 3 + i];
        }
    float vut[9];
    mat3_mul(v, ut, vut);
    if (det3f(vut) < 0.0f) { /* :253-261 */
        vt[6] = -vt[6];
        vt[7] = -vt[7];
        vt[8] = -vt[8];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) v[i * 3 + j] = vt[j * 3 + i];
    }
    float R[9];
    mat3_mul(v, ut, R); /* :263 */
    for (int i = 0; i < 9; i++) out->rotation[i] = R[i];
    for (int i = 0; i < 3; i++) /* :264 */
        out->translation[i] = tc[i] - (R[i * 3 + 0] * sc[0] + R[i * 3 + 1] * sc[1] + R[i * 3 + 2] * sc[2]);
}

/* icp_plane.rs:131-236 */
static void solve_point_to_plane(const float *sx, const float *sy, const float *sz, const float *tx,
                                 const float *ty, const float *tz, const float *nx, const float *ny,
                                 const float *nz, const uint32_t *si, const uint32_t *ti, size_t m,
                                 orc_transform *out) {
    transform_identity(out);
    if (m == 0) return;
    double ata[36] = {0}, atb[6] = {0};
    for (size_t c = 0; c < m; c++) { /* :148-180 f64 */
        double s0 = sx[si[c]], s1 = sy[si[c]], s2 = sz[si[c]];
        double t0 = tx[ti[c]], t1 = ty[ti[c]], t2 = tz[ti[c]];
        double n0 = nx[ti[c]], n1 = ny[ti[c]], n2 = nz[ti[c]];
        double a[6] = {s1 * n2 - s2 * n1, s2 * n0 - s0 * n2, s0 * n1 - s1 * n0, n0, n1, n2};
        double b = (t0 - s0) * n0 + (t1 - s1) * n1 + (t2 - s2) * n2;
        for (int i = 0; i < 6; i++) {
            for (int j = 0; j < 6; j++) ata[i * 6 + j] += a[i] * a[j];
            atb[i] += a[i] * b;
        }
    }
    double diag_max = 0.0; /* :185-189 */
    for (int i = 0; i < 6; i++) diag_max = fmax(diag_max, fabs(ata[i * 6 + i]));
    double lambda = 1e-6 * fmax(diag_max, 1e-12);
    for (int i = 0; i < 6; i++) ata[i * 6 + i] += lambda;

    double xs[6];
    int solved = 0;
    { /* Cholesky (:193) */
        double L[36] = {0};
        int ok = 1;
        for (int j = 0; j < 6 && ok; j++) {
            double d = ata[j * 6 + j];
            for (int k = 0; k < j; k++) d -= L[j * 6 + k] * L[j * 6 + k];
            if (!(d > 0.0)) {
                ok = 0;
                break;
            }
            L[j * 6 + j] = sqrt(d);
            for (int i = j + 1; i < 6; i++) {
                double v = ata[i * 6 + j];
                for (int k = 0; k < j; k++) v -= L[i * 6 + k] * L[j * 6 + k];
                L[i * 6 + j] = v / L[j * 6 + j];
            }
        }
        if (ok) {
            double yv[6];
            for (int i = 0; i < 6; i++) {
                double v = atb[i];
                for (int k = 0; k < i; k++) v -= L[i * 6 + k] * yv[k];
                yv[i] = v / L[i * 6 + i];
            }
            for (int i = 5; i >= 0; i--) {
                double v = yv[i];
                for (int k = i + 1; k < 6; k++) v -= L[k * 6 + i] * xs[k];
                xs[i] = v / L[i * 6 + i];
            }
            solved = 1;
        }
    }
    if (!solved) { /* LU with partial pivoting (:194-197) */
        double A[36], b[6];
        memcpy(A, ata, sizeof(A));
        memcpy(b, atb, sizeof(b));
        int ok = 1;
        for (int c = 0; c < 6 && ok; c++) {
            int piv = c;
            for (int r = c + 1; r < 6; r++)
                if (fabs(A[r * 6 + c]) > fabs(A[piv * 6 + c])) piv = r;
            if (A[piv * 6 + c] == 0.0 || !isfinite(A[piv * 6 + c])) {
                ok = 0;
                break;
            }
            if (piv != c) {
                for (int k = 0; k < 6; k++) {
                    double t = A[c * 6 + k];
                    A[c * 6 + k] = A[piv * 6 + k];
                    A[piv * 6 + k] = t;
                }
                double t = b[c];
                b[c] = b[piv];
                b[piv] = t;
            }
            for (int r = c + 1; r < 6; r++) {
                double f = A[r * 6 + c] / A[c * 6 + c];
                for (int k = c; k < 6; k++) A[r * 6 + k] -= f * A[c * 6 + k];
                b[r] -= f * b[c];
            }
        }
        if (!ok) return; /* identity (:197) */
        for (int i = 5; i >= 0; i--) {
            double v = b[i];
            for (int k = i + 1; k < 6; k++) v -= A[i * 6 + k] * xs[k];
            xs[i] = v / A[i * 6 + i];
        }
    }
    float alpha = (float)xs[0], beta = (float)xs[1], gamma = (float)xs[2]; /* :201-206 */
    float ttx = (float)xs[3], tty = (float)xs[4], ttz = (float)xs[5];
    float angle = sqrtf(alpha * alpha + beta * beta + gamma * gamma); /* :209 */
    float *R = out->rotation;
    if (angle < 1e-10f) { /* :211-216 */
        R[0] = 1.0f; R[1] = -gamma; R[2] = beta;
        R[3] = gamma; R[4] = 1.0f; R[5] = -alpha;
        R[6] = -beta; R[7] = alpha; R[8] = 1.0f;
    } else { /* :217-230 Rodrigues, f32 */
        float ax = alpha / angle, ay = beta / angle, az = gamma / angle;
        float c = cosf(angle), s = sinf(angle), t = 1.0f - c;
        R[0] = t * ax * ax + c;      R[1] = t * ax * ay - s * az; R[2] = t * ax * az + s * ay;
        R[3] = t * ax * ay + s * az; R[4] = t * ay * ay + c;      R[5] = t * ay * az - s * ax;
        R[6] = t * ax * az - s * ay; R[7] = t * ay * az + s * ax; R[8] = t * az * az + c;
    }
    out->translation[0] = ttx;
    out->translation[1] = tty;
    out->translation[2] = ttz;
}

/* shared ICP loop: icp.rs:125-206 and icp_plane.rs:20-97 are the same loop around a different solve */
static void icp_loop(const float *sx, const float *sy, const float *sz, size_t ns, const float *tx,
                     const float *ty, const float *tz, size_t nt, const float *nx, const float *ny,
                     const float *nz, int plane, const orc_icp_params *p, orc_icp_result *out,
                     int threads) {
    transform_identity(&out->transform);
    out->fitness = 0.f;
    out->rmse = 0.f;
    out->converged = 0;
    out->num_iterations = 0;
    if (ns == 0 || nt == 0) { /* icp.rs:131-139 */
        out->converged = (ns == 0 && nt == 0);
        return;
    }
    orc_tree *tree = orc_tree_build(tx, ty, tz, nt);
    float *cx = (float *)malloc(sizeof(float) * ns), *cy = (float *)malloc(sizeof(float) * ns),
          *cz = (float *)malloc(sizeof(float) * ns);
    orc_transform ident;
    transform_identity(&ident);
    orc_apply_transform(sx, sy, sz, ns, &ident, cx, cy, cz); /* icp.rs:145 */
    orc_transform cumulative = ident;
    uint32_t *si = (uint32_t *)malloc(sizeof(uint32_t) * ns), *ti = (uint32_t *)malloc(sizeof(uint32_t) * ns);
    float *dd = (float *)malloc(sizeof(float) * ns);
    float prev_rmse = INFINITY, last_rmse = INFINITY, last_fitness = 0.0f;
    int converged = 0;
    size_t num_iterations = 0;
    for (size_t iter = 0; iter < p->max_iterations; iter++) {
        num_iterations = iter + 1;
        size_t m = orc_find_correspondences(cx, cy, cz, ns, tree, p->max_correspondence_distance, si, ti, dd, threads);
        if (m == 0) break;
        float rmse = compute_rmse(dd, m);
        last_rmse = rmse;
        last_fitness = (float)m / (float)ns;
        if (fabsf(prev_rmse - rmse) < p->tolerance) {
            converged = 1;
            break;
        }
        prev_rmse = rmse;
        orc_transform inc;
        if (plane) solve_point_to_plane(cx, cy, cz, tx, ty, tz, nx, ny, nz, si, ti, m, &inc);
        else rigid_transform_svd(cx, cy, cz, tx, ty, tz, si, ti, m, &inc);
        orc_compose(&cumulative, &inc, &cumulative);
        orc_apply_transform(cx, cy, cz, ns, &inc, cx, cy, cz);
    }
    if (num_iterations == 0) { /* icp.rs:190-197 */
        size_t m = orc_find_correspondences(cx, cy, cz, ns, tree, p->max_correspondence_distance, si, ti, dd, threads);
        if (m) {
            last_rmse = compute_rmse(dd, m);
            last_fitness = (float)m / (float)ns;
        }
    }
    out->transform = cumulative;
    out->fitness = last_fitness;
    out->rmse = last_rmse;
    out->converged = converged;
    out->num_iterations = num_iterations;
    free(cx); free(cy); free(cz); free(si); free(ti); free(dd);
    orc_tree_free(tree);
}

void orc_icp_point_to_point(const float *sx, const float *sy, const float *sz, size_t ns,
                            const float *tx, const float *ty, const float *tz, size_t nt,
                            const orc_icp_params *p, orc_icp_result *out, int threads) {
    icp_loop(sx, sy, sz, ns, tx, ty, tz, nt, NULL, NULL, NULL, 0, p, out, threads);
}

int orc_icp_point_to_plane(const float *sx, const float *sy, const float *sz, size_t ns,
                           const float *tx, const float *ty, const float *tz, size_t nt,
                           const float *nx, const float *ny, const float *nz, size_t nn,
                           const orc_icp_params *p, orc_icp_result *out, int threads) {
    if (nn != nt) return 1; /* icp_plane.rs:27-32 */
    icp_loop(sx, sy, sz, ns, tx, ty, tz, nt, nx, ny, nz, 1, p, out, threads);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* voxel_downsample  (crates/filters/src/voxel_downsample.rs:12-65)                           */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    int32_t k[3];
    uint32_t i;
} vox_key;

static int cmp_vox(const void *a, const void *b) {
    const vox_key *p = (const vox_key *)a, *q = (const vox_key *)b;
    for (int c = 0; c < 3; c++)
        if (p->k[c] != q->k[c]) return p->k[c] < q->k[c] ? -1 : 1;
    return p->i < q->i ? -1 : (p->i > q->i ? 1 : 0);
}

static int32_t sat_i32(float f) { /* Rust `as i32`: saturating, NaN -> 0 */
    if (isnan(f)) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}

size_t orc_voxel_downsample(const float *x, const float *y, const float *z, size_t n,
                            float voxel_size, float *ox, float *oy, float *oz) {
    if (!isfinite(voxel_size) || !(voxel_size > 0.0f)) return (size_t)-1;
    if (n == 0) return 0;
    vox_key *keys = (vox_key *)malloc(sizeof(vox_key) * n);
    size_t m = 0;
    for (size_t i = 0; i < n; i++) {
        if (!finite3(x[i], y[i], z[i])) continue; /* :28-30 */
        keys[m].k[0] = sat_i32(floorf(x[i] / voxel_size)); /* :32-36 */
        keys[m].k[1] = sat_i32(floorf(y[i] / voxel_size));
        keys[m].k[2] = sat_i32(floorf(z[i] / voxel_size));
        keys[m].i = (uint32_t)i;
        m++;
    }
    qsort(keys, m, sizeof(vox_key), cmp_vox); /* key order (:49-50); input order inside a voxel (:38-42) */
    size_t out = 0, a = 0;
    while (a < m) {
        size_t b = a;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        while (b < m && keys[b].k[0] == keys[a].k[0] && keys[b].k[1] == keys[a].k[1] && keys[b].k[2] == keys[a].k[2]) {
            s0 += x[keys[b].i];
            s1 += y[keys[b].i];
            s2 += z[keys[b].i];
            b++;
        }
        float denom = (float)(b - a);
        ox[out] = s0 / denom;
        oy[out] = s1 / denom;
        oz[out] = s2 / denom;
        out++;
        a = b;
    }
    free(keys);
    return out;
}

/* ------------------------------------------------------------------------------------------ */
/* euclidean_cluster  (crates/segmentation/src/euclidean_cluster.rs:96-187)                   */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    uint32_t *parent;
    uint8_t *rank;
} orc_uf;

static uint32_t uf_find(orc_uf *u, uint32_t x) { /* :20-28, path splitting */
    while (u->parent[x] != x) {
        uint32_t p = u->parent[x];
        u->parent[x] = u->parent[p];
        x = u->parent[x];
    }
    return x;
}

static void uf_union(orc_uf *u, uint32_t a, uint32_t b) { /* :30-46, union by rank */
    uint32_t ra = uf_find(u, a), rb = uf_find(u, b);
    if (ra == rb) return;
    if (u->rank[ra] < u->rank[rb]) u->parent[ra] = rb;
    else if (u->rank[ra] > u->rank[rb]) u->parent[rb] = ra;
    else {
        u->parent[rb] = ra;
        u->rank[ra]++;
    }
}

static int cmp_key3(const int32_t *a, const int32_t *b) {
    for (int c = 0; c < 3; c++)
        if (a[c] != b[c]) return a[c] < b[c] ? -1 : 1;
    return 0;
}

/* first entry of `keys` (sorted) whose key equals k, or -1 */
static long find_cell(const vox_key *keys, size_t m, const int32_t k[3]) {
    size_t lo = 0, hi = m;
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        if (cmp_key3(keys[mid].k, k) < 0) lo = mid + 1;
        else hi = mid;
    }
    return (lo < m && cmp_key3(keys[lo].k, k) == 0) ? (long)lo : -1;
}

typedef struct {
    uint32_t root, size, first;
} orc_comp;

static int cmp_comp(const void *a, const void *b) { /* :183-185: size descending, then lexicographic */
    const orc_comp *p = (const orc_comp *)a, *q = (const orc_comp *)b;
    if (p->size != q->size) return p->size > q->size ? -1 : 1;
    return p->first < q->first ? -1 : (p->first > q->first ? 1 : 0);
}

/* components of `parent` -> CSR in the reference's output order; returns the cluster count */
static size_t emit_clusters(orc_uf *u, size_t n, size_t min_size, size_t max_size, uint32_t *offsets, uint32_t *indices) {
    uint32_t *root = (uint32_t *)malloc(sizeof(uint32_t) * n), *size = (uint32_t *)calloc(n, sizeof(uint32_t));
    uint32_t *first = (uint32_t *)malloc(sizeof(uint32_t) * n), *slot = (uint32_t *)malloc(sizeof(uint32_t) * n);
    for (size_t i = 0; i < n; i++) {
        root[i] = uf_find(u, (uint32_t)i); /* :164-168 */
        if (size[root[i]]++ == 0) first[root[i]] = (uint32_t)i;
    }
    orc_comp *comps = (orc_comp *)malloc(sizeof(orc_comp) * n);
    size_t nc = 0;
    for (size_t r = 0; r < n; r++)
        if (size[r] && size[r] >= min_size && size[r] <= max_size) { /* :172-175 */
            comps[nc].root = (uint32_t)r;
            comps[nc].size = size[r];
            comps[nc].first = first[r];
            nc++;
        }
    qsort(comps, nc, sizeof(orc_comp), cmp_comp);
    for (size_t r = 0; r < n; r++) slot[r] = UINT32_MAX;
    offsets[0] = 0;
    for (size_t c = 0; c < nc; c++) {
        slot[comps[c].root] = offsets[c];
        offsets[c + 1] = offsets[c] + comps[c].size;
    }
    for (size_t i = 0; i < n; i++) /* ascending i => indices inside a cluster ascending (:177) */
        if (slot[root[i]] != UINT32_MAX) indices[slot[root[i]]++] = (uint32_t)i;
    free(root); free(size); free(first); free(slot); free(comps);
    return nc;
}

size_t orc_euclidean_cluster(const float *x, const float *y, const float *z, size_t n, float distance_threshold,
                             size_t min_size, size_t max_size, uint32_t *offsets, uint32_t *indices) {
    offsets[0] = 0;
    if (n == 0 || distance_threshold <= 0.0f || min_size == 0) return 0; /* :102-104 */
    const float inv_r = 1.0f / distance_threshold;               /* :107 */
    const float r2 = distance_threshold * distance_threshold;    /* :108 */
    vox_key *keys = (vox_key *)malloc(sizeof(vox_key) * n);
    size_t m = 0;
    for (size_t i = 0; i < n; i++) { /* :112-119, cell_key :55-61 */
        if (!finite3(x[i], y[i], z[i])) continue;
        keys[m].k[0] = sat_i32(floorf(x[i] * inv_r));
        keys[m].k[1] = sat_i32(floorf(y[i] * inv_r));
        keys[m].k[2] = sat_i32(floorf(z[i] * inv_r));
        keys[m].i = (uint32_t)i;
        m++;
    }
    qsort(keys, m, sizeof(vox_key), cmp_vox); /* cells = runs of equal keys, input order inside a cell */
    orc_uf u;
    u.parent = (uint32_t *)malloc(sizeof(uint32_t) * n);
    u.rank = (uint8_t *)calloc(n, 1);
    for (size_t i = 0; i < n; i++) u.parent[i] = (uint32_t)i;
    static const int half[14][3] = {{0, 0, 0},  {1, 0, 0},  {1, 1, 0},   {1, -1, 0}, {1, 0, 1}, {1, 0, -1}, {1, 1, 1},
                                    {1, 1, -1}, {1, -1, 1}, {1, -1, -1}, {0, 1, 0},  {0, 1, 1}, {0, 1, -1}, {0, 0, 1}}; /* :65-82 */
    size_t a0 = 0;
    while (a0 < m) {
        size_t a1 = a0;
        while (a1 < m && cmp_key3(keys[a1].k, keys[a0].k) == 0) a1++;
        for (int o = 0; o < 14; o++) { /* :131-157 */
            /* i32 wrap-around of cx + dx is a panic in debug Rust / wraps in release; never reached by finite data */
            int32_t nk[3] = {(int32_t)((int64_t)keys[a0].k[0] + half[o][0]), (int32_t)((int64_t)keys[a0].k[1] + half[o][1]),
                             (int32_t)((int64_t)keys[a0].k[2] + half[o][2])};
            long b0 = o == 0 ? (long)a0 : find_cell(keys, m, nk);
            if (b0 < 0) continue;
            size_t b1 = (size_t)b0;
            while (b1 < m && cmp_key3(keys[b1].k, nk) == 0) b1++;
            for (size_t ai = a0; ai < a1; ai++) {
                const uint32_t ia = keys[ai].i;
                size_t start = o == 0 ? ai + 1 : (size_t)b0; /* :145 */
                for (size_t bi = start; bi < b1; bi++) {
                    const uint32_t ib = keys[bi].i;
                    if (dist2(x[ia], y[ia], z[ia], x[ib], y[ib], z[ib]) <= r2) uf_union(&u, ia, ib); /* :147-151, :160-165 */
                }
            }
        }
        a0 = a1;
    }
    size_t nc = emit_clusters(&u, n, min_size, max_size, offsets, indices);
    free(keys); free(u.parent); free(u.rank);
    return nc;
}

/* tests/cluster_differential.rs:13-82: the reference's own O(n^2) checker */
size_t orc_euclidean_cluster_brute(const float *x, const float *y, const float *z, size_t n, float distance_threshold,
                                   size_t min_size, size_t max_size, uint32_t *offsets, uint32_t *indices) {
    offsets[0] = 0;
    if (n == 0 || distance_threshold <= 0.0f || min_size == 0) return 0;
    const float r2 = distance_threshold * distance_threshold;
    orc_uf u;
    u.parent = (uint32_t *)malloc(sizeof(uint32_t) * n);
    u.rank = (uint8_t *)calloc(n, 1);
    for (size_t i = 0; i < n; i++) u.parent[i] = (uint32_t)i;
    for (size_t i = 0; i < n; i++) {
        if (!finite3(x[i], y[i], z[i])) continue;
        for (size_t j = i + 1; j < n; j++) {
            if (!finite3(x[j], y[j], z[j])) continue;
            if (dist2(x[i], y[i], z[i], x[j], y[j], z[j]) <= r2) uf_union(&u, (uint32_t)i, (uint32_t)j);
        }
    }
    size_t nc = emit_clusters(&u, n, min_size, max_size, offsets, indices);
    free(u.parent); free(u.rank);
    return nc;
}

/* ------------------------------------------------------------------------------------------ */
/* ransac_plane_seeded  (crates/segmentation/src/ransac_plane.rs:56-129) for GIVEN samples     */
/* ------------------------------------------------------------------------------------------ */

/* :166-190; returns 0 if the three points are collinear */
static int fit_plane3(const float p0[3], const float p1[3], const float p2[3], float model[4]) {
    const float v1[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
    const float v2[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
    const float nx = v1[1] * v2[2] - v1[2] * v2[1];
    const float ny = v1[2] * v2[0] - v1[0] * v2[2];
    const float nz = v1[0] * v2[1] - v1[1] * v2[0];
    const float len = sqrtf(nx * nx + ny * ny + nz * nz);
    if (len < 1e-10f) return 0;
    model[0] = nx / len;
    model[1] = ny / len;
    model[2] = nz / len;
    model[3] = -(model[0] * p0[0] + model[1] * p0[1] + model[2] * p0[2]);
    return 1;
}

static inline float plane_dist(const float m[4], float x, float y, float z) { /* :17-20 */
    return fabsf(m[0] * x + m[1] * y + m[2] * z + m[3]);
}

/* The sampling (StdRng = ChaCha12, :75-78, :140-163) stays with the caller: `samples` holds m index
 * triples exactly as sample_three_distinct produced them.  model = {nx, ny, nz, d}; inliers sized n. */
size_t orc_ransac_plane_samples(const float *x, const float *y, const float *z, size_t n, float threshold,
                                const uint32_t *samples, size_t m, float model[4], uint32_t *inliers) {
    float best[4] = {0.f, 0.f, 1.f, 0.f}; /* PlaneModel::default(), :23-30 */
    memcpy(model, best, sizeof(best));
    if (n < 3) return 0; /* :64-66 */
    size_t best_count = 0;
    int have = 0;
    const int parallel = n >= 10000 && m >= 16; /* :80 */
    for (size_t it = 0; it < m; it++) {
        const uint32_t i0 = samples[3 * it], i1 = samples[3 * it + 1], i2 = samples[3 * it + 2];
        const float p0[3] = {x[i0], y[i0], z[i0]}, p1[3] = {x[i1], y[i1], z[i1]}, p2[3] = {x[i2], y[i2], z[i2]};
        float cand[4];
        if (!fit_plane3(p0, p1, p2, cand)) continue; /* :86 / :98-101 */
        size_t cnt = 0;
        for (size_t j = 0; j < n; j++) cnt += plane_dist(cand, x[j], y[j], z[j]) <= threshold ? 1 : 0; /* :132-137 */
        if (parallel) {
            /* :84-93: reduce_with(|a, b| if a.1 >= b.1 { a } else { b }) over an ordered iterator keeps the FIRST
             * hypothesis with the largest count; the default model is used only if no hypothesis is valid */
            if (!have || cnt > best_count) {
                have = 1;
                best_count = cnt;
                memcpy(best, cand, sizeof(best));
            }
        } else if (cnt > best_count) { /* :106 */
            best_count = cnt;
            memcpy(best, cand, sizeof(best));
            const double w = (double)best_count / (double)n; /* :110-117 adaptive early termination */
            if (w > 0.5) {
                const double needed = log(1.0 - 0.999) / log(1.0 - w * w * w);
                if ((double)it > needed) break;
            }
        }
    }
    memcpy(model, best, sizeof(best));
    size_t k = 0; /* :123-126 */
    for (size_t j = 0; j < n; j++)
        if (plane_dist(best, x[j], y[j], z[j]) <= threshold) inliers[k++] = (uint32_t)j;
    return k;
}

/* ------------------------------------------------------------------------------------------ */
/* ASCII PCD body reader (crates/io/src/pcd.rs:202-234)                                       */
/* ------------------------------------------------------------------------------------------ */

long orc_read_pcd_ascii(const char *path, float *x, float *y, float *z, size_t cap) {
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    char line[4096];
    int in_data = 0;
    long n = 0;
    while (fgets(line, sizeof(line), f)) {
        char *s = line;
        while (*s == ' ' || *s == '\t') s++;
        if (strncmp(s, "DATA", 4) == 0) {
            in_data = 1;
            continue;
        }
        if (!in_data || *s == '\n' || *s == '\r' || *s == 0 || *s == '#') continue;
        char *e1, *e2, *e3;
        float a = strtof(s, &e1);
        if (e1 == s) continue;
        float b = strtof(e1, &e2);
        if (e2 == e1) continue;
        float c = strtof(e2, &e3);
        if (e3 == e2) continue;
        if ((size_t)n < cap) {
            x[n] = a;
            y[n] = b;
            z[n] = c;
        }
        n++;
    }
    fclose(f);
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* Uniform-grid ring search MODEL (mirrors pointclouds_rs_b200/csrc search rule) */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    double o[3], h, inv_h;
    int dims[3];
    uint32_t *cell_start; /* dims0*dims1*dims2 + 1 */
    float *px, *py, *pz;
    uint32_t *pidx;
    size_t m;
} grid_model;

static inline int cell_of(const grid_model *g, float v, int a) {
    double t = floor(((double)v - g->o[a]) * g->inv_h);
    if (t < 0) return 0;
    if (t > g->dims[a] - 1) return g->dims[a] - 1;
    return (int)t;
}
static inline size_t cell_lin(const grid_model *g, int cx, int cy, int cz) {
    return ((size_t)cx * g->dims[1] + cy) * g->dims[2] + cz;
}

static void grid_scan_run(const grid_model *g, int cx, int cy, int z0, int z1, const float q[3], kheap *h, double *stats) {
    /* one contiguous run of points: cells (cx,cy,z0..z1) are adjacent in the sorted order */
    uint32_t b = g->cell_start[cell_lin(g, cx, cy, z0)], e = g->cell_start[cell_lin(g, cx, cy, z1) + 1];
    for (uint32_t i = b; i < e; i++)
        heap_push(h, make_key(dist2(q[0], q[1], q[2], g->px[i], g->py[i], g->pz[i]), g->pidx[i]));
    if (stats) {
        stats[0] += (double)(e - b);
        stats[2] += 1;
        stats[3] += (double)((e - b + 31) / 32);
    }
}

void orc_grid_knn_model(const float *x, const float *y, const float *z, size_t n, const float *qx,
                        const float *qy, const float *qz, size_t nq, size_t k, float cell,
                        uint32_t *idx, float *dist, uint32_t *counts, double *stats) {
    grid_model g;
    memset(&g, 0, sizeof(g));
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    size_t m = 0;
    for (size_t i = 0; i < n; i++) {
        if (!finite3(x[i], y[i], z[i])) continue;
        const float v[3] = {x[i], y[i], z[i]};
        for (int a = 0; a < 3; a++) {
            if (v[a] < mn[a]) mn[a] = v[a];
            if (v[a] > mx[a]) mx[a] = v[a];
        }
        m++;
    }
    g.m = m;
    g.h = (double)cell;
    g.inv_h = 1.0 / g.h;
    size_t ncell = 1;
    for (int a = 0; a < 3; a++) {
        g.o[a] = m ? (double)mn[a] : 0.0;
        g.dims[a] = m ? (int)floor(((double)mx[a] - g.o[a]) * g.inv_h) + 1 : 1;
        ncell *= (size_t)g.dims[a];
    }
    g.cell_start = (uint32_t *)calloc(ncell + 1, sizeof(uint32_t));
    g.px = (float *)malloc(sizeof(float) * (m ? m : 1));
    g.py = (float *)malloc(sizeof(float) * (m ? m : 1));
    g.pz = (float *)malloc(sizeof(float) * (m ? m : 1));
    g.pidx = (uint32_t *)malloc(sizeof(uint32_t) * (m ? m : 1));
    for (size_t i = 0; i < n; i++)
        if (finite3(x[i], y[i], z[i])) g.cell_start[cell_lin(&g, cell_of(&g, x[i], 0), cell_of(&g, y[i], 1), cell_of(&g, z[i], 2)) + 1]++;
    for (size_t c = 0; c < ncell; c++) g.cell_start[c + 1] += g.cell_start[c];
    uint32_t *fill = (uint32_t *)malloc(sizeof(uint32_t) * (ncell ? ncell : 1));
    memcpy(fill, g.cell_start, sizeof(uint32_t) * ncell);
    for (size_t i = 0; i < n; i++)
        if (finite3(x[i], y[i], z[i])) {
            uint32_t p = fill[cell_lin(&g, cell_of(&g, x[i], 0), cell_of(&g, y[i], 1), cell_of(&g, z[i], 2))]++;
            g.px[p] = x[i];
            g.py[p] = y[i];
            g.pz[p] = z[i];
            g.pidx[p] = (uint32_t)i;
        }
    free(fill);

    uint64_t *hb = (uint64_t *)malloc(sizeof(uint64_t) * (k ? k : 1));
    for (size_t qi = 0; qi < nq; qi++) {
        float q[3] = {qx[qi], qy[qi], qz[qi]};
        size_t r = 0;
        uint32_t *ri = idx ? idx + qi * k : NULL;
        float *rd = dist ? dist + qi * k : NULL;
        if (k > 0 && m > 0 && finite3(q[0], q[1], q[2])) {
            size_t kk = k < m ? k : m;
            kheap h = {hb, 0, kk};
            int c[3] = {cell_of(&g, q[0], 0), cell_of(&g, q[1], 1), cell_of(&g, q[2], 2)};
            for (int R = 0;; R++) {
                if (stats) stats[1] += 1;
                /* shell R: Chebyshev distance exactly R, clipped to the grid */
                for (int dx = -R; dx <= R; dx++) {
                    int cx = c[0] + dx;
                    if (cx < 0 || cx >= g.dims[0]) continue;
                    for (int dy = -R; dy <= R; dy++) {
                        int cy = c[1] + dy;
                        if (cy < 0 || cy >= g.dims[1]) continue;
                        int full = (abs(dx) == R || abs(dy) == R);
                        if (full) {
                            int z0 = c[2] - R < 0 ? 0 : c[2] - R;
                            int z1 = c[2] + R >= g.dims[2] ? g.dims[2] - 1 : c[2] + R;
                            grid_scan_run(&g, cx, cy, z0, z1, q, &h, stats);
                        } else {
                            if (c[2] - R >= 0) grid_scan_run(&g, cx, cy, c[2] - R, c[2] - R, q, &h, stats);
                            if (R > 0 && c[2] + R < g.dims[2]) grid_scan_run(&g, cx, cy, c[2] + R, c[2] + R, q, &h, stats);
                        }
                    }
                }
                /* termination: lower bound on the distance to anything not yet scanned.  Beyond an
                 * open face of axis a everything is at least dist_a away along a AND at least gap_b
                 * away along every other axis b on which the query lies outside the grid. */
                double gap[3], g2 = 0.0;
                for (int a = 0; a < 3; a++) {
                    double lo_ext = g.o[a], hi_ext = g.o[a] + (double)g.dims[a] * g.h;
                    gap[a] = (double)q[a] < lo_ext ? lo_ext - (double)q[a] : ((double)q[a] > hi_ext ? (double)q[a] - hi_ext : 0.0);
                    g2 += gap[a] * gap[a];
                }
                double bound2 = INFINITY;
                int open = 0;
                for (int a = 0; a < 3; a++) {
                    int lo = c[a] - R, hi = c[a] + R;
                    if (lo > 0) {
                        open = 1;
                        double d = (double)q[a] - (g.o[a] + (double)lo * g.h);
                        double b2 = d > 0.0 ? d * d + (g2 - gap[a] * gap[a]) : 0.0;
                        if (b2 < bound2) bound2 = b2;
                    }
                    if (hi < g.dims[a] - 1) {
                        open = 1;
                        double d = (g.o[a] + (double)(hi + 1) * g.h) - (double)q[a];
                        double b2 = d > 0.0 ? d * d + (g2 - gap[a] * gap[a]) : 0.0;
                        if (b2 < bound2) bound2 = b2;
                    }
                }
                if (!open) break;
                if (h.n == h.cap && bound2 > 0.0 && (double)key_d2(h.a[0]) < bound2 * (1.0 - 1e-6)) break;
            }
            r = heap_emit(&h, ri, rd);
        }
        for (size_t j = r; j < k; j++) {
            if (ri) ri[j] = UINT32_MAX;
            if (rd) rd[j] = INFINITY;
        }
        if (counts) counts[qi] = (uint32_t)r;
    }
    free(hb);
    free(g.cell_start);
    free(g.px);
    free(g.py);
    free(g.pz);
    free(g.pidx);
}
