/*
 * oracle/cl2_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked, imported or called by the product path).
 *
 * Sequential CPU restatement of the cLoops2 loop-calling hot path, used as the parity checker for the
 * CUDA library (tests/, __graft_entry__.smoke()) and as the `cpu_baseline` / `--impl reference` leg of
 * bench.py.  Pinned against the live Python reference in the build container (tests/test_oracle_vs_reference.py,
 * oracle/gen_golden.py -> tests/golden/).  Reference files are cited as path:line under /root/reference.
 *
 *   orc_block_dbscan   follows cLoops2/blockDBSCAN.py:13-234 step by step (dict order == first appearance,
 *                      stable sort by cell population, BFS with the same claim / overwrite rules).
 *   orc_cdbscan2       restates cLoops2/cDBSCAN2.py:13-383: 45-degree rotated grid,
 *                      axis-only neighbour tests, first-come border claims, release of < minPts clusters.
 *   orc_loop_counts    set-based counts of cLoops2/ds.py:165-198 (XY._query/queryPeak/queryLoop) as used by
 *                      callCisLoops.py:173-354 (estLoopSig, getPerRegions, estAnchorSig) and
 *                      callTransLoops.py:106-193.
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/_build/libcl2_oracle.so oracle/cl2_oracle.c   (oracle/Makefile)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* ------------------------------------------------------------------ hash map: cell key -> cell index */
typedef struct {
    uint64_t *keys;
    int32_t *vals;
    uint64_t mask;
} hmap;

static uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
static int hmap_init(hmap *h, int64_t n) {
    uint64_t cap = 16;
    while (cap < (uint64_t)n * 2 + 2) cap <<= 1;
    h->keys = (uint64_t *)malloc(cap * sizeof(uint64_t));
    h->vals = (int32_t *)malloc(cap * sizeof(int32_t));
    if (!h->keys || !h->vals) return -1;
    memset(h->vals, 0xff, cap * sizeof(int32_t));
    h->mask = cap - 1;
    return 0;
}
static void hmap_free(hmap *h) { free(h->keys); free(h->vals); }
static int32_t hmap_get(const hmap *h, uint64_t key) {
    uint64_t s = mix64(key) & h->mask;
    while (h->vals[s] >= 0) {
        if (h->keys[s] == key) return h->vals[s];
        s = (s + 1) & h->mask;
    }
    return -1;
}
/* returns existing index or inserts `next` */
static int32_t hmap_put(hmap *h, uint64_t key, int32_t next) {
    uint64_t s = mix64(key) & h->mask;
    while (h->vals[s] >= 0) {
        if (h->keys[s] == key) return h->vals[s];
        s = (s + 1) & h->mask;
    }
    h->keys[s] = key; h->vals[s] = next;
    return next;
}
static uint64_t cellkey(int64_t nx, int64_t ny) {
    return ((uint64_t)(uint32_t)(int32_t)nx << 32) | (uint64_t)(uint32_t)(int32_t)ny;
}

/* neighbour order of getNearbyGrids / getNearbyCells (blockDBSCAN.py:54-55, cDBSCAN2.py:42-43) */
static const int NBR_DX[8] = {0, 0, -1, 1, -1, -1, 1, 1};
static const int NBR_DY[8] = {-1, 1, 0, 0, -1, 1, -1, 1};

/* ================================================================== blockDBSCAN */
typedef struct {
    int64_t n, eps;
    int minPts;
    const int64_t *xy;
    int32_t ncell;
    int64_t *cx, *cy;      /* cell grid coords */
    int32_t *cstart;       /* points of cell c: plist[cstart[c] .. cstart[c+1]) in row order */
    int32_t *plist;
    int32_t *nbr;          /* [ncell*8] cell index or -1 */
    uint8_t *alive;
    double *mx, *my;       /* centroids (blockDBSCAN.py:124-139) */
    int32_t *label;        /* -1 unclassified, -2 noise, >=0 cluster */
    int8_t *adj;           /* [ncell*8] memoised regionQuery edge: -1 unknown, 0 no, 1 yes */
} bctx;

static int b_cnt(const bctx *b, int32_t c) { return b->cstart[c + 1] - b->cstart[c]; }

/* getGridDist, blockDBSCAN.py:198-208 (early return on first close pair) */
static int b_griddist(const bctx *b, int32_t ka, int32_t kb) {
    for (int32_t i = b->cstart[ka]; i < b->cstart[ka + 1]; i++) {
        int64_t px = b->xy[2 * (int64_t)b->plist[i]], py = b->xy[2 * (int64_t)b->plist[i] + 1];
        for (int32_t j = b->cstart[kb]; j < b->cstart[kb + 1]; j++) {
            int64_t qx = b->xy[2 * (int64_t)b->plist[j]], qy = b->xy[2 * (int64_t)b->plist[j] + 1];
            int64_t d = llabs(px - qx) + llabs(py - qy);
            if (d <= b->eps) return 1;
        }
    }
    return 0;
}

/* regionQuery, blockDBSCAN.py:210-234: result list (self first) and population sum */
static int b_region(bctx *b, int32_t k, int32_t *res, int64_t *psum) {
    int nres = 0;
    res[nres++] = k;
    int64_t s = b_cnt(b, k);
    for (int d = 0; d < 8; d++) {
        int32_t q = b->nbr[8 * (int64_t)k + d];
        if (q < 0 || !b->alive[q]) continue;
        int8_t a = b->adj[8 * (int64_t)k + d];
        if (a < 0) {
            double dist = fabs(b->mx[k] - b->mx[q]) + fabs(b->my[k] - b->my[q]);
            if (dist <= (double)b->eps) a = 1;
            else a = (int8_t)b_griddist(b, k, q);
            b->adj[8 * (int64_t)k + d] = a;
        }
        if (a) { res[nres++] = q; s += b_cnt(b, q); }
    }
    *psum = s;
    return nres;
}

typedef struct { int32_t cnt; int32_t idx; } sortrec;
static int cmp_cnt_desc(const void *a, const void *b) {
    const sortrec *x = (const sortrec *)a, *y = (const sortrec *)b;
    if (x->cnt != y->cnt) return x->cnt > y->cnt ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);   /* stable: dict order */
}

int orc_block_dbscan(const int64_t *xy, int64_t n, int64_t eps, int minPts, int32_t *labels,
                     int32_t *n_clusters) {
    *n_clusters = 0;
    for (int64_t i = 0; i < n; i++) labels[i] = -1;
    if (n == 0) return 0;
    if (eps <= 0) return -1;
    bctx b; memset(&b, 0, sizeof(b));
    b.n = n; b.eps = eps; b.minPts = minPts; b.xy = xy;
    /* buildGrids, blockDBSCAN.py:69-86 */
    int64_t minX = xy[0], minY = xy[1];
    for (int64_t i = 0; i < n; i++) {
        if (xy[2 * i] < minX) minX = xy[2 * i];
        if (xy[2 * i + 1] < minY) minY = xy[2 * i + 1];
    }
    hmap h;
    if (hmap_init(&h, n)) return -2;
    int32_t *cell_of = (int32_t *)malloc(n * sizeof(int32_t));
    b.cx = (int64_t *)malloc(n * sizeof(int64_t));
    b.cy = (int64_t *)malloc(n * sizeof(int64_t));
    int32_t nc = 0;
    for (int64_t i = 0; i < n; i++) {
        int64_t nx = (xy[2 * i] - minX) / eps + 1, ny = (xy[2 * i + 1] - minY) / eps + 1;
        int32_t c = hmap_put(&h, cellkey(nx, ny), nc);
        if (c == nc) { b.cx[nc] = nx; b.cy[nc] = ny; nc++; }
        cell_of[i] = c;
    }
    b.ncell = nc;
    b.cstart = (int32_t *)calloc((size_t)nc + 1, sizeof(int32_t));
    for (int64_t i = 0; i < n; i++) b.cstart[cell_of[i] + 1]++;
    for (int32_t c = 0; c < nc; c++) b.cstart[c + 1] += b.cstart[c];
    b.plist = (int32_t *)malloc(n * sizeof(int32_t));
    {
        int32_t *fill = (int32_t *)malloc((size_t)nc * sizeof(int32_t));
        memcpy(fill, b.cstart, (size_t)nc * sizeof(int32_t));
        for (int64_t i = 0; i < n; i++) b.plist[fill[cell_of[i]]++] = (int32_t)i;
        free(fill);
    }
    /* buildGridNeighbors, blockDBSCAN.py:88-99 */
    b.nbr = (int32_t *)malloc((size_t)nc * 8 * sizeof(int32_t));
    int64_t *s3 = (int64_t *)malloc((size_t)nc * sizeof(int64_t));
    for (int32_t c = 0; c < nc; c++) {
        int64_t s = b_cnt(&b, c);
        for (int d = 0; d < 8; d++) {
            int32_t q = hmap_get(&h, cellkey(b.cx[c] + NBR_DX[d], b.cy[c] + NBR_DY[d]));
            b.nbr[8 * (int64_t)c + d] = q;
            if (q >= 0) s += b_cnt(&b, q);
        }
        s3[c] = s;
    }
    /* removeNoiseGrids, blockDBSCAN.py:101-122 */
    b.alive = (uint8_t *)malloc(nc);
    for (int32_t c = 0; c < nc; c++) {
        int dead = 0;
        if (s3[c] < minPts) {
            dead = 1;
            for (int d = 0; d < 8; d++) {
                int32_t q = b.nbr[8 * (int64_t)c + d];
                if (q >= 0 && !(s3[q] < minPts)) { dead = 0; break; }
            }
        }
        b.alive[c] = !dead;
    }
    /* centerGrids, blockDBSCAN.py:124-139 */
    b.mx = (double *)malloc((size_t)nc * sizeof(double));
    b.my = (double *)malloc((size_t)nc * sizeof(double));
    b.label = (int32_t *)malloc((size_t)nc * sizeof(int32_t));
    b.adj = (int8_t *)malloc((size_t)nc * 8);
    memset(b.adj, 0xff, (size_t)nc * 8);
    int32_t nalive = 0;
    for (int32_t c = 0; c < nc; c++) {
        b.label[c] = -1;
        if (!b.alive[c]) continue;
        nalive++;
        int64_t sx = 0, sy = 0;
        for (int32_t i = b.cstart[c]; i < b.cstart[c + 1]; i++) {
            sx += xy[2 * (int64_t)b.plist[i]]; sy += xy[2 * (int64_t)b.plist[i] + 1];
        }
        b.mx[c] = (double)sx / (double)b_cnt(&b, c);
        b.my[c] = (double)sy / (double)b_cnt(&b, c);
    }
    /* callClusters, blockDBSCAN.py:141-152 */
    sortrec *ord = (sortrec *)malloc(((size_t)nalive + 1) * sizeof(sortrec));
    {
        int32_t k = 0;
        for (int32_t c = 0; c < nc; c++) if (b.alive[c]) { ord[k].cnt = b_cnt(&b, c); ord[k].idx = c; k++; }
        qsort(ord, nalive, sizeof(sortrec), cmp_cnt_desc);
    }
    int32_t *queue = (int32_t *)malloc(((size_t)nalive * 9 + 16) * sizeof(int32_t));
    int32_t clusterId = 0;
    int32_t res[9];
    for (int32_t oi = 0; oi < nalive; oi++) {
        int32_t key = ord[oi].idx;
        if (b.label[key] != -1) continue;
        /* expandCluster, blockDBSCAN.py:170-196 */
        int64_t near;
        int nres = b_region(&b, key, res, &near);
        if (near < minPts) { b.label[key] = -2; continue; }
        int64_t qh = 0, qt = 0;
        for (int i = 0; i < nres; i++) { b.label[res[i]] = clusterId; queue[qt++] = res[i]; }
        while (qh < qt) {
            int32_t cur = queue[qh++];
            nres = b_region(&b, cur, res, &near);
            if (near < minPts) continue;
            for (int i = 1; i < nres; i++) {
                if (b.label[res[i]] == -1 || b.label[res[i]] == -2) {
                    b.label[res[i]] = clusterId;
                    queue[qt++] = res[i];
                }
            }
        }
        clusterId++;
    }
    /* getLabels, blockDBSCAN.py:154-168 */
    for (int64_t i = 0; i < n; i++) {
        int32_t c = cell_of[i];
        if (b.alive[c] && b.label[c] >= 0) labels[i] = b.label[c];
    }
    *n_clusters = clusterId;
    free(queue); free(ord); free(b.adj); free(b.label); free(b.mx); free(b.my); free(b.alive);
    free(s3); free(b.nbr); free(b.plist); free(b.cstart); free(b.cx); free(b.cy); free(cell_of);
    hmap_free(&h);
    return 0;
}

/* blockDBSCAN's removeNoiseGrids alone (blockDBSCAN.py:69-122): keep[i]=1 iff the point survives.
 * Used to check the first kernels of the CUDA path in isolation. */
int orc_block_noise(const int64_t *xy, int64_t n, int64_t eps, int minPts, uint8_t *keep) {
    if (n == 0) return 0;
    int64_t minX = xy[0], minY = xy[1];
    for (int64_t i = 0; i < n; i++) {
        if (xy[2 * i] < minX) minX = xy[2 * i];
        if (xy[2 * i + 1] < minY) minY = xy[2 * i + 1];
    }
    hmap h;
    if (hmap_init(&h, n)) return -2;
    int32_t *cell_of = (int32_t *)malloc(n * sizeof(int32_t));
    int64_t *cx = (int64_t *)malloc(n * sizeof(int64_t)), *cy = (int64_t *)malloc(n * sizeof(int64_t));
    int64_t *cnt = (int64_t *)calloc(n, sizeof(int64_t));
    int32_t nc = 0;
    for (int64_t i = 0; i < n; i++) {
        int64_t nx = (xy[2 * i] - minX) / eps + 1, ny = (xy[2 * i + 1] - minY) / eps + 1;
        int32_t c = hmap_put(&h, cellkey(nx, ny), nc);
        if (c == nc) { cx[nc] = nx; cy[nc] = ny; nc++; }
        cell_of[i] = c; cnt[c]++;
    }
    int64_t *s3 = (int64_t *)malloc((size_t)nc * sizeof(int64_t));
    for (int32_t c = 0; c < nc; c++) {
        int64_t s = cnt[c];
        for (int d = 0; d < 8; d++) {
            int32_t q = hmap_get(&h, cellkey(cx[c] + NBR_DX[d], cy[c] + NBR_DY[d]));
            if (q >= 0) s += cnt[q];
        }
        s3[c] = s;
    }
    uint8_t *alive = (uint8_t *)malloc(nc);
    for (int32_t c = 0; c < nc; c++) {
        int dead = 0;
        if (s3[c] < minPts) {
            dead = 1;
            for (int d = 0; d < 8; d++) {
                int32_t q = hmap_get(&h, cellkey(cx[c] + NBR_DX[d], cy[c] + NBR_DY[d]));
                if (q >= 0 && !(s3[q] < minPts)) { dead = 0; break; }
            }
        }
        alive[c] = !dead;
    }
    for (int64_t i = 0; i < n; i++) keep[i] = alive[cell_of[i]];
    free(alive); free(s3); free(cnt); free(cx); free(cy); free(cell_of); hmap_free(&h);
    return 0;
}

/* ================================================================== cDBSCAN2 */
typedef struct {
    int64_t eps;
    const int64_t *u, *v;
    const int32_t *cstart, *plist, *nbr;
} cctx;

/* neighbour test of point p (in cell c) against q in the cell at offset d of c:
 * binSearchAdjPt (cDBSCAN2.py:364-378): delta=+1 -> pos <= p+eps ; delta=-1 -> pos >= p-eps; axis-only for
 * edge-adjacent cells (:320-327), both axes for corner cells (:328-332). d<0 means same cell. */
static inline int c_near(const cctx *c, int d, int32_t p, int32_t q) {
    if (d < 0) return 1;
    int dx = NBR_DX[d], dy = NBR_DY[d];
    if (dx == 1 && !(c->u[q] <= c->u[p] + c->eps)) return 0;
    if (dx == -1 && !(c->u[q] >= c->u[p] - c->eps)) return 0;
    if (dy == 1 && !(c->v[q] <= c->v[p] + c->eps)) return 0;
    if (dy == -1 && !(c->v[q] >= c->v[p] - c->eps)) return 0;
    return 1;
}

static int32_t uf_find(int32_t *par, int32_t x) {
    while (par[x] != x) { par[x] = par[par[x]]; x = par[x]; }
    return x;
}
static void uf_union(int32_t *par, int32_t a, int32_t b) {
    a = uf_find(par, a); b = uf_find(par, b);
    if (a == b) return;
    if (a < b) par[b] = a; else par[a] = b;
}

typedef struct { int32_t key; int32_t root; } comprec;
static int cmp_comp(const void *a, const void *b) {
    const comprec *x = (const comprec *)a, *y = (const comprec *)b;
    return x->key < y->key ? -1 : (x->key > y->key ? 1 : 0);
}

int orc_cdbscan2(const int64_t *xy, int64_t n, int64_t eps, int minPts, int32_t *labels,
                 int32_t *n_clusters) {
    *n_clusters = 0;
    for (int64_t i = 0; i < n; i++) labels[i] = -1;
    if (n == 0) return 0;
    if (eps <= 0) return -1;
    /* buildGrid, cDBSCAN2.py:55-75: rotate, truncating division, dict order = first appearance */
    int64_t *u = (int64_t *)malloc(n * sizeof(int64_t)), *v = (int64_t *)malloc(n * sizeof(int64_t));
    hmap h;
    if (hmap_init(&h, n)) return -2;
    int32_t *cell_of = (int32_t *)malloc(n * sizeof(int32_t));
    int64_t *cx = (int64_t *)malloc(n * sizeof(int64_t)), *cy = (int64_t *)malloc(n * sizeof(int64_t));
    int32_t nc = 0;
    for (int64_t i = 0; i < n; i++) {
        u[i] = xy[2 * i] - xy[2 * i + 1];
        v[i] = xy[2 * i] + xy[2 * i + 1];
        int64_t nx = u[i] / eps + 1, ny = v[i] / eps + 1;   /* C division truncates toward zero == int() */
        int32_t c = hmap_put(&h, cellkey(nx, ny), nc);
        if (c == nc) { cx[nc] = nx; cy[nc] = ny; nc++; }
        cell_of[i] = c;
    }
    int32_t *cstart = (int32_t *)calloc((size_t)nc + 1, sizeof(int32_t));
    for (int64_t i = 0; i < n; i++) cstart[cell_of[i] + 1]++;
    for (int32_t c = 0; c < nc; c++) cstart[c + 1] += cstart[c];
    int32_t *plist = (int32_t *)malloc(n * sizeof(int32_t));
    {
        int32_t *fill = (int32_t *)malloc((size_t)nc * sizeof(int32_t));
        memcpy(fill, cstart, (size_t)nc * sizeof(int32_t));
        for (int64_t i = 0; i < n; i++) plist[fill[cell_of[i]]++] = (int32_t)i;
        free(fill);
    }
    int32_t *nbr = (int32_t *)malloc((size_t)nc * 8 * sizeof(int32_t));
    for (int32_t c = 0; c < nc; c++)
        for (int d = 0; d < 8; d++)
            nbr[8 * (int64_t)c + d] = hmap_get(&h, cellkey(cx[c] + NBR_DX[d], cy[c] + NBR_DY[d]));
    cctx cc = {eps, u, v, cstart, plist, nbr};
    /* core points: crowded cells (:80-83) are all core; otherwise count neighbours (:333-334) */
    uint8_t *core = (uint8_t *)calloc(n, 1);
    for (int32_t c = 0; c < nc; c++) {
        int32_t cnt = cstart[c + 1] - cstart[c];
        if (cnt >= minPts) {
            for (int32_t i = cstart[c]; i < cstart[c + 1]; i++) core[plist[i]] = 1;
            continue;
        }
        for (int32_t i = cstart[c]; i < cstart[c + 1]; i++) {
            int32_t p = plist[i];
            int64_t deg = cnt;
            for (int d = 0; d < 8 && deg < minPts; d++) {
                int32_t q = nbr[8 * (int64_t)c + d];
                if (q < 0) continue;
                for (int32_t j = cstart[q]; j < cstart[q + 1]; j++)
                    if (c_near(&cc, d, p, plist[j])) deg++;
            }
            if (deg >= minPts) core[p] = 1;
        }
    }
    /* components of core points */
    int32_t *par = (int32_t *)malloc(n * sizeof(int32_t));
    for (int64_t i = 0; i < n; i++) par[i] = (int32_t)i;
    for (int32_t c = 0; c < nc; c++) {
        int32_t firstcore = -1;
        for (int32_t i = cstart[c]; i < cstart[c + 1]; i++) {
            int32_t p = plist[i];
            if (!core[p]) continue;
            if (firstcore < 0) firstcore = p; else uf_union(par, firstcore, p);   /* same cell -> always */
        }
        if (firstcore < 0) continue;
        for (int32_t i = cstart[c]; i < cstart[c + 1]; i++) {
            int32_t p = plist[i];
            if (!core[p]) continue;
            for (int d = 0; d < 8; d++) {
                int32_t q = nbr[8 * (int64_t)c + d];
                if (q < 0 || q < c) continue;   /* symmetric relation: handle each cell pair once */
                for (int32_t j = cstart[q]; j < cstart[q + 1]; j++) {
                    int32_t qq = plist[j];
                    if (core[qq] && c_near(&cc, d, p, qq)) uf_union(par, p, qq);
                }
            }
        }
    }
    /* order components by the first cell (dict order) holding one of their core points (:117-140) */
    int32_t *ckey = (int32_t *)malloc(n * sizeof(int32_t));
    for (int64_t i = 0; i < n; i++) ckey[i] = INT32_MAX;
    int32_t ncomp = 0;
    for (int64_t i = 0; i < n; i++) {
        if (!core[i]) continue;
        int32_t r = uf_find(par, (int32_t)i);
        if (ckey[r] == INT32_MAX) ncomp++;
        if (cell_of[i] < ckey[r]) ckey[r] = cell_of[i];
    }
    comprec *comps = (comprec *)malloc(((size_t)ncomp + 1) * sizeof(comprec));
    {
        int32_t k = 0;
        for (int64_t i = 0; i < n; i++)
            if (core[i] && par[i] == (int32_t)i) { comps[k].key = ckey[i]; comps[k].root = (int32_t)i; k++; }
        qsort(comps, ncomp, sizeof(comprec), cmp_comp);
    }
    /* member lists of components (core points), bucketed by root */
    int32_t *comp_of_root = (int32_t *)malloc(n * sizeof(int32_t));
    for (int32_t k = 0; k < ncomp; k++) comp_of_root[comps[k].root] = k;
    int32_t *mstart = (int32_t *)calloc((size_t)ncomp + 1, sizeof(int32_t));
    for (int64_t i = 0; i < n; i++) if (core[i]) mstart[comp_of_root[uf_find(par, (int32_t)i)] + 1]++;
    for (int32_t k = 0; k < ncomp; k++) mstart[k + 1] += mstart[k];
    int32_t *mlist = (int32_t *)malloc(((size_t)mstart[ncomp] + 1) * sizeof(int32_t));
    {
        int32_t *fill = (int32_t *)malloc(((size_t)ncomp + 1) * sizeof(int32_t));
        memcpy(fill, mstart, (size_t)ncomp * sizeof(int32_t));
        for (int64_t i = 0; i < n; i++) if (core[i]) mlist[fill[comp_of_root[uf_find(par, (int32_t)i)]]++] = (int32_t)i;
        free(fill);
    }
    /* process in order: claim unlabelled border points, release small clusters (:130-185) */
    int32_t *claimed = (int32_t *)malloc(n * sizeof(int32_t));
    int32_t cid = 0;
    for (int32_t k = 0; k < ncomp; k++) {
        int64_t nmem = mstart[k + 1] - mstart[k];
        int64_t nclaim = 0;
        for (int32_t m = mstart[k]; m < mstart[k + 1]; m++) {
            int32_t p = mlist[m];
            int32_t c = cell_of[p];
            for (int d = -1; d < 8; d++) {
                int32_t q = d < 0 ? c : nbr[8 * (int64_t)c + d];
                if (q < 0) continue;
                for (int32_t j = cstart[q]; j < cstart[q + 1]; j++) {
                    int32_t qq = plist[j];
                    if (core[qq] || labels[qq] != -1) continue;
                    if (c_near(&cc, d, p, qq)) { labels[qq] = -3; claimed[nclaim++] = qq; }
                }
            }
        }
        if (nmem + nclaim < minPts) {
            for (int64_t i = 0; i < nclaim; i++) labels[claimed[i]] = -1;
        } else {
            for (int64_t i = 0; i < nclaim; i++) labels[claimed[i]] = cid;
            for (int32_t m = mstart[k]; m < mstart[k + 1]; m++) labels[mlist[m]] = cid;
            cid++;
        }
    }
    *n_clusters = cid;
    free(claimed); free(mlist); free(mstart); free(comp_of_root); free(comps); free(ckey); free(par);
    free(core); free(nbr); free(plist); free(cstart); free(cx); free(cy); free(cell_of); hmap_free(&h);
    free(u); free(v);
    return 0;
}

/* ================================================================== counting (XY index) */
typedef struct {
    int64_t n;
    int64_t *xs, *ys;       /* sorted coordinate values (ds.py:161-162) */
    int32_t *xrow, *yrow;   /* row index of each sorted entry (stands in for x2i / y2i, ds.py:155-159) */
    int32_t *stampA, *stampB;
    int32_t stamp;
} xyidx;

typedef struct { int64_t v; int32_t r; } vr;
static int cmp_vr(const void *a, const void *b) {
    const vr *x = (const vr *)a, *y = (const vr *)b;
    if (x->v != y->v) return x->v < y->v ? -1 : 1;
    return x->r < y->r ? -1 : (x->r > y->r ? 1 : 0);
}

void *orc_index_build(const int64_t *xy, int64_t n) {
    xyidx *ix = (xyidx *)calloc(1, sizeof(xyidx));
    ix->n = n;
    ix->xs = (int64_t *)malloc((n + 1) * sizeof(int64_t)); ix->ys = (int64_t *)malloc((n + 1) * sizeof(int64_t));
    ix->xrow = (int32_t *)malloc((n + 1) * sizeof(int32_t)); ix->yrow = (int32_t *)malloc((n + 1) * sizeof(int32_t));
    ix->stampA = (int32_t *)calloc(n + 1, sizeof(int32_t)); ix->stampB = (int32_t *)calloc(n + 1, sizeof(int32_t));
    vr *tmp = (vr *)malloc((n + 1) * sizeof(vr));
    for (int64_t i = 0; i < n; i++) { tmp[i].v = xy[2 * i]; tmp[i].r = (int32_t)i; }
    qsort(tmp, n, sizeof(vr), cmp_vr);
    for (int64_t i = 0; i < n; i++) { ix->xs[i] = tmp[i].v; ix->xrow[i] = tmp[i].r; }
    for (int64_t i = 0; i < n; i++) { tmp[i].v = xy[2 * i + 1]; tmp[i].r = (int32_t)i; }
    qsort(tmp, n, sizeof(vr), cmp_vr);
    for (int64_t i = 0; i < n; i++) { ix->ys[i] = tmp[i].v; ix->yrow[i] = tmp[i].r; }
    free(tmp);
    return ix;
}
void orc_index_free(void *p) {
    xyidx *ix = (xyidx *)p;
    if (!ix) return;
    free(ix->xs); free(ix->ys); free(ix->xrow); free(ix->yrow); free(ix->stampA); free(ix->stampB); free(ix);
}

/* np.searchsorted(cor, left, "left") / (cor, right, "right") with float bounds, ds.py:170-171 */
static int64_t ss_left(const int64_t *a, int64_t n, double x) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if ((double)a[m] < x) lo = m + 1; else hi = m; }
    return lo;
}
static int64_t ss_right(const int64_t *a, int64_t n, double x) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if ((double)a[m] <= x) lo = m + 1; else hi = m; }
    return lo;
}

/* queryPeak (ds.py:176-182) into a stamp array; returns the set size */
static int64_t q_mark(xyidx *ix, double l, double r, int32_t *stamp, int32_t s) {
    int64_t c = 0;
    int64_t a = ss_left(ix->xs, ix->n, l), b = ss_right(ix->xs, ix->n, r);
    for (int64_t i = a; i < b; i++) if (stamp[ix->xrow[i]] != s) { stamp[ix->xrow[i]] = s; c++; }
    a = ss_left(ix->ys, ix->n, l); b = ss_right(ix->ys, ix->n, r);
    for (int64_t i = a; i < b; i++) if (stamp[ix->yrow[i]] != s) { stamp[ix->yrow[i]] = s; c++; }
    return c;
}
/* |queryPeak(l,r) & markedA| where A was marked with stamp sa; also returns |queryPeak(l,r)| */
static int64_t q_inter(xyidx *ix, double l, double r, int32_t sa, int64_t *size_b) {
    int32_t sb = ++ix->stamp;
    int64_t c = 0, nb = 0;
    int64_t a = ss_left(ix->xs, ix->n, l), b = ss_right(ix->xs, ix->n, r);
    for (int64_t i = a; i < b; i++) {
        int32_t row = ix->xrow[i];
        if (ix->stampB[row] != sb) { ix->stampB[row] = sb; nb++; if (ix->stampA[row] == sa) c++; }
    }
    a = ss_left(ix->ys, ix->n, l); b = ss_right(ix->ys, ix->n, r);
    for (int64_t i = a; i < b; i++) {
        int32_t row = ix->yrow[i];
        if (ix->stampB[row] != sb) { ix->stampB[row] = sb; nb++; if (ix->stampA[row] == sa) c++; }
    }
    if (size_b) *size_b = nb;
    return c;
}

static double dmax(double a, double b) { return a > b ? a : b; }

/*
 * Per-loop raw counts.  mode 0 = cis (callCisLoops.py:250-354), 1 = trans (callTransLoops.py:106-193; only the
 * P2LL window differs, :142-145).  All windows use the reference's float arithmetic.
 *   loops    [L,4] x_start,x_end,y_start,y_end
 *   core     [L,4] ra, rb, rab, lower_rab                      (queryLoop :287-289, P2LL :302-306)
 *   na, nb   [L,2*win] |Na_i|, |Nb_j|                          (getPerRegions :173-198)
 *   nab      [L,(2*win)^2] |Na_i & Nb_j|, a-major              (getLoopNearbyPETs :207-223)
 *   anchors  [L,4] count_x, wide_x, count_y, wide_y            (estAnchorSig :226-247, ext=5)
 */
int orc_loop_counts(void *pidx, const int64_t *loops, int64_t L, int win, int mode, int64_t *core,
                    int64_t *na, int64_t *nb, int64_t *nab, int64_t *anchors) {
    xyidx *ix = (xyidx *)pidx;
    int W = 2 * win;
    int32_t *sav = (int32_t *)malloc((size_t)W * sizeof(int32_t));
    for (int64_t k = 0; k < L; k++) {
        double xs = (double)loops[4 * k], xe = (double)loops[4 * k + 1];
        double ys = (double)loops[4 * k + 2], ye = (double)loops[4 * k + 3];
        /* queryLoop */
        int32_t sa = ++ix->stamp;
        int64_t ra = q_mark(ix, xs, xe, ix->stampA, sa);
        int64_t rb = 0;
        int64_t rab = q_inter(ix, ys, ye, sa, &rb);
        core[4 * k] = ra; core[4 * k + 1] = rb; core[4 * k + 2] = rab;
        /* P2LL window */
        double al, ar, bl, br;
        if (mode == 0) { al = xe; ar = (xe - xs) + xe; bl = ys - (ye - ys); br = ys; }
        else { al = xs - (xe - xs); ar = xs; bl = ys - (ye - ys); br = ys; }
        sa = ++ix->stamp;
        q_mark(ix, al, ar, ix->stampA, sa);
        core[4 * k + 3] = q_inter(ix, bl, br, sa, NULL);
        /* permuted windows, getPerRegions */
        double ca = (xs + xe) / 2, cb = (ys + ye) / 2;
        double sA = (xe - xs) / 2, sB = (ye - ys) / 2;
        double step = (sA + sB) / 2;
        int t = 0;
        /* nab[i][j] = |Na_i & Nb_j|: mark each Na_i in its own stamp plane is costly; do it a-major */
        for (int i = -win; i <= win; i++) {
            if (i == 0) continue;
            double l = dmax(0, ca + i * step - sA), r = dmax(0, ca + i * step + sA);
            int32_t s = ++ix->stamp;
            na[(int64_t)W * k + t] = q_mark(ix, l, r, ix->stampA, s);
            int tj = 0;
            for (int j = -win; j <= win; j++) {
                if (j == 0) continue;
                double l2 = dmax(0, cb + j * step - sB), r2 = dmax(0, cb + j * step + sB);
                int64_t szb = 0;
                nab[(int64_t)W * W * k + (int64_t)t * W + tj] = q_inter(ix, l2, r2, s, &szb);
                nb[(int64_t)W * k + tj] = szb;
                tj++;
            }
            t++;
        }
        /* estAnchorSig */
        for (int a = 0; a < 2; a++) {
            double left = a == 0 ? xs : ys, right = a == 0 ? xe : ye;
            double m = (left + right) / 2, length = right - left;
            int32_t s = ++ix->stamp;
            anchors[4 * k + 2 * a] = q_mark(ix, left, right, ix->stampA, s);
            s = ++ix->stamp;
            anchors[4 * k + 2 * a + 1] = q_mark(ix, dmax(0, m - 5 * length), m + 5 * length, ix->stampA, s);
        }
    }
    free(sav);
    return 0;
}

/* cluster -> bounding boxes (callCisLoops.py:92-105): bbox[c] = xmin,xmax,ymin,ymax ; size[c] */
int orc_cluster_bbox(const int64_t *xy, int64_t n, const int32_t *labels, int32_t ncl, int64_t *bbox,
                     int64_t *size) {
    for (int32_t c = 0; c < ncl; c++) {
        bbox[4 * c] = INT64_MAX; bbox[4 * c + 1] = INT64_MIN; bbox[4 * c + 2] = INT64_MAX; bbox[4 * c + 3] = INT64_MIN;
        size[c] = 0;
    }
    for (int64_t i = 0; i < n; i++) {
        int32_t c = labels[i];
        if (c < 0) continue;
        if (c >= ncl) return -1;
        int64_t x = xy[2 * i], y = xy[2 * i + 1];
        if (x < bbox[4 * c]) bbox[4 * c] = x;
        if (x > bbox[4 * c + 1]) bbox[4 * c + 1] = x;
        if (y < bbox[4 * c + 2]) bbox[4 * c + 2] = y;
        if (y > bbox[4 * c + 3]) bbox[4 * c + 3] = y;
        size[c]++;
    }
    return 0;
}
