// cl2_pre.cu -- the de-duplication step of `cLoops2 pre` (getUniqueTxt, cLoops2/io.py:40-52): the reference keeps a
// Python set of the text rows of one chromosome pair; here the rows are sorted by the packed (X, Y) key with the
// path's radix sort and the first row of every run of equal keys is kept.  The result comes back ascending by
// (X, Y), which also makes the row order of the .ixy file deterministic (the reference's is the set's).
#include "cl2_points.cuh"

__global__ void __launch_bounds__(256) k_uniq_flags(const int2 *__restrict__ xy, const uint32_t *__restrict__ row,
                                                     int32_t *__restrict__ f, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int2 p = xy[row[i]];
    int2 q = i > 0 ? xy[row[i - 1]] : make_int2(-1, -1);
    f[i] = (i == 0 || p.x != q.x || p.y != q.y) ? 1 : 0;
}
__global__ void __launch_bounds__(256) k_uniq_write(const int2 *__restrict__ xy, const uint32_t *__restrict__ row,
                                                     const int32_t *__restrict__ pos, int64_t n, int64_t total,
                                                     longlong2 *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t p = pos[i];
    const bool head = (i + 1 < n) ? (pos[i + 1] != p) : ((int64_t)p != total);
    if (head) {
        int2 v = xy[row[i]];
        out[p] = make_longlong2(v.x, v.y);
    }
}

extern "C" int cl2_points_unique(cl2_ctx *ctx, cl2_points *pts, int64_t *xy_out, int64_t *n_out) {
    if (!ctx || !pts || !n_out) return cl2_fail(ctx, CL2_ERR_ARG, "cl2_points_unique: bad argument");
    *n_out = 0;
    const int64_t n = pts->n;
    if (n == 0) return CL2_OK;
    if (!xy_out) return cl2_fail(ctx, CL2_ERR_ARG, "cl2_points_unique: no output buffer");
    CL2_CUDA(ctx, cudaSetDevice(ctx->device));
    Scratch sc(ctx);
    uint64_t *k0, *k1, *ks;
    uint32_t *v0, *v1, *vs;
    int32_t *f;
    int64_t *d_total;
    longlong2 *d_out;
    CL2_TRY(sc.get(&k0, (size_t)n));
    CL2_TRY(sc.get(&k1, (size_t)n));
    CL2_TRY(sc.get(&v0, (size_t)n));
    CL2_TRY(sc.get(&v1, (size_t)n));
    CL2_TRY(sc.get(&f, (size_t)n));
    CL2_TRY(sc.get(&d_total, 1));
    // packed key (X - minX) << by | (Y - minY): the blockDBSCAN cell key with eps = 1
    cl2_keygen gen;
    gen.kind = CL2_KEY_BLOCK;
    gen.xy = pts->xy;
    gen.minx = (int)pts->ext[0];
    gen.miny = (int)pts->ext[2];
    gen.eps = 1;
    gen.by = bits_for((uint64_t)(pts->ext[3] - pts->ext[2]));
    const int bx = bits_for((uint64_t)(pts->ext[1] - pts->ext[0]));
    CL2_TRY(cl2_radix_sort_gen(ctx, gen, k0, k1, v0, v1, n, bx + gen.by, &ks, &vs));
    {
        FamTimer t(ctx, FAM_MISC, 16.0 * (double)n);
        k_uniq_flags<<<cl2_grid(n, 256), 256, 0, ctx->stream>>>(pts->xy, vs, f, n);
    }
    CL2_TRY(cl2_exclusive_scan_i32(ctx, f, n, d_total));
    int64_t total = 0;
    CL2_TRY(cl2_read_i64(ctx, d_total, &total));
    CL2_TRY(sc.get(&d_out, (size_t)total));
    {
        FamTimer t(ctx, FAM_MISC, 16.0 * (double)n + 16.0 * (double)total);
        k_uniq_write<<<cl2_grid(n, 256), 256, 0, ctx->stream>>>(pts->xy, vs, f, n, total, d_out);
    }
    CL2_CUDA(ctx, cudaMemcpyAsync(xy_out, d_out, (size_t)total * 16, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->d2h_bytes += total * 16;
    CL2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    CL2_CUDA(ctx, cudaGetLastError());
    *n_out = total;
    return CL2_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Host side of `pre`: BEDPE text -> numeric records (the field handling of PET.__init__, cLoops2/ds.py:42-57).
// The reference tokenises every line in Python; this is the same tokeniser as a single pass over the buffer.
#include <string>
#include <unordered_map>

struct cl2_bedpe {
    std::unordered_map<std::string, int32_t> ids;
    std::vector<std::string> names;      // id -> chromosome name, first-seen order
    std::string joined;                  // scratch for cl2_bedpe_names
};

extern "C" int cl2_bedpe_create(cl2_bedpe **out) {
    if (!out) return CL2_ERR_ARG;
    *out = new (std::nothrow) cl2_bedpe();
    return *out ? CL2_OK : CL2_ERR_NOMEM;
}
extern "C" void cl2_bedpe_destroy(cl2_bedpe *p) { delete p; }

// int(field) of Python for the plain cases: optional blanks, optional sign, decimal digits, optional blanks
static bool parse_int_field(const char *b, const char *e, int64_t *out) {
    while (b < e && (*b == ' ' || *b == '\r' || *b == '\f' || *b == '\v')) b++;
    while (e > b && (e[-1] == ' ' || e[-1] == '\r' || e[-1] == '\f' || e[-1] == '\v')) e--;
    if (b == e) return false;
    bool neg = false;
    if (*b == '+' || *b == '-') { neg = *b == '-'; b++; }
    if (b == e || e - b > 18) return false;
    int64_t v = 0;
    for (; b < e; b++) {
        if (*b < '0' || *b > '9') return false;
        v = v * 10 + (*b - '0');
    }
    *out = neg ? -v : v;
    return true;
}

// Parses the whole lines of buf[0, len).  For every accepted line i (in file order) writes chrom[2i], chrom[2i+1] =
// ids of fields 0 and 3, pos[4i..4i+3] = fields 1, 2, 4, 5 and mapq[i] = field 7 (255 when it is not an integer).
// A line needs at least 10 fields (ds.py:49,53 read fields 8 and 9) and integer coordinates, otherwise it is skipped
// exactly as the reference's `try: PET(line) except: continue` (io.py:97-100) skips it.  Returns the number of
// accepted lines, or a negative error code; *n_lines (optional) receives the number of lines seen.
extern "C" int64_t cl2_bedpe_parse(cl2_bedpe *p, const char *buf, int64_t len, int64_t cap, int32_t *chrom, int64_t *pos,
                                   int32_t *mapq, int64_t *n_lines) {
    if (!p || (len > 0 && !buf) || !chrom || !pos || !mapq) return CL2_ERR_ARG;
    int64_t n = 0, lines = 0;
    const char *cur = buf, *end = buf + len;
    std::string key;
    while (cur < end) {
        const char *nl = (const char *)memchr(cur, '\n', (size_t)(end - cur));
        const char *le = nl ? nl : end;
        lines++;
        const char *fb[10], *fe[10];
        int nf = 0;
        const char *q = cur;
        while (nf < 10) {
            const char *tab = (const char *)memchr(q, '\t', (size_t)(le - q));
            fb[nf] = q;
            fe[nf] = tab ? tab : le;
            nf++;
            if (!tab) break;
            q = tab + 1;
        }
        cur = nl ? nl + 1 : end;
        if (nf < 10) continue;
        int64_t v[4];
        if (!parse_int_field(fb[1], fe[1], &v[0]) || !parse_int_field(fb[2], fe[2], &v[1]) ||
            !parse_int_field(fb[4], fe[4], &v[2]) || !parse_int_field(fb[5], fe[5], &v[3]))
            continue;
        if (n >= cap) return CL2_ERR_RANGE;
        int64_t mq;
        mapq[n] = parse_int_field(fb[7], fe[7], &mq) ? (int32_t)(mq > INT32_MAX ? INT32_MAX : (mq < INT32_MIN ? INT32_MIN : mq)) : 255;
        for (int s = 0; s < 2; s++) {
            const int f = s == 0 ? 0 : 3;
            key.assign(fb[f], (size_t)(fe[f] - fb[f]));
            auto it = p->ids.find(key);
            int32_t id;
            if (it == p->ids.end()) {
                id = (int32_t)p->names.size();
                p->ids.emplace(key, id);
                p->names.push_back(key);
            } else {
                id = it->second;
            }
            chrom[2 * n + s] = id;
        }
        for (int j = 0; j < 4; j++) pos[4 * n + j] = v[j];
        n++;
    }
    if (n_lines) *n_lines = lines;
    return n;
}

// chromosome names seen so far, '\n'-joined, id order; the pointer stays valid until the next call on `p`
extern "C" const char *cl2_bedpe_names(cl2_bedpe *p, int32_t *count) {
    if (!p) return "";
    p->joined.clear();
    for (size_t i = 0; i < p->names.size(); i++) {
        if (i) p->joined.push_back('\n');
        p->joined += p->names[i];
    }
    if (count) *count = (int32_t)p->names.size();
    return p->joined.c_str();
}

// .pairs text (https://pairtools.readthedocs.io formats) -> records: the line handling of parsePairs and Pair.__init__
// (cLoops2/io.py:189-207, ds.py:111-137).  A record line has exactly 7 or 8 tab-separated fields; with 8 the pair type
// (field 7) must be one of UU, MR, MU, RU, UR; fields 2 and 4 must be integers; lines starting with '#' are headers.
// For the i-th accepted line: chrom[2i], chrom[2i+1] = ids of fields 1 and 3, pos[2i], pos[2i+1] = fields 2 and 4.
extern "C" int64_t cl2_pairs_parse(cl2_bedpe *p, const char *buf, int64_t len, int64_t cap, int32_t *chrom, int64_t *pos,
                                   int64_t *n_lines) {
    if (!p || (len > 0 && !buf) || !chrom || !pos) return CL2_ERR_ARG;
    int64_t n = 0, lines = 0;
    const char *cur = buf, *end = buf + len;
    std::string key;
    while (cur < end) {
        const char *nl = (const char *)memchr(cur, '\n', (size_t)(end - cur));
        const char *le = nl ? nl : end;
        lines++;
        const char *fb[9], *fe[9];
        int nf = 0;
        const char *q = cur;
        const bool header = cur < le && *cur == '#';
        while (!header && nf < 9) {                           // a ninth field already disqualifies the line
            const char *tab = (const char *)memchr(q, '\t', (size_t)(le - q));
            fb[nf] = q;
            fe[nf] = tab ? tab : le;
            nf++;
            if (!tab) break;
            q = tab + 1;
        }
        cur = nl ? nl + 1 : end;
        if (header || (nf != 7 && nf != 8)) continue;
        if (nf == 8) {
            const char *t = fb[7];
            const bool ok = fe[7] - t == 2 && ((t[0] == 'U' && t[1] == 'U') || (t[0] == 'M' && t[1] == 'R') ||
                                               (t[0] == 'M' && t[1] == 'U') || (t[0] == 'R' && t[1] == 'U') ||
                                               (t[0] == 'U' && t[1] == 'R'));
            if (!ok) continue;
        }
        int64_t v[2];
        if (!parse_int_field(fb[2], fe[2], &v[0]) || !parse_int_field(fb[4], fe[4], &v[1])) continue;
        if (n >= cap) return CL2_ERR_RANGE;
        for (int s = 0; s < 2; s++) {
            const int f = s == 0 ? 1 : 3;
            key.assign(fb[f], (size_t)(fe[f] - fb[f]));
            auto it = p->ids.find(key);
            int32_t id;
            if (it == p->ids.end()) {
                id = (int32_t)p->names.size();
                p->ids.emplace(key, id);
                p->names.push_back(key);
            } else {
                id = it->second;
            }
            chrom[2 * n + s] = id;
        }
        pos[2 * n] = v[0];
        pos[2 * n + 1] = v[1];
        n++;
    }
    if (n_lines) *n_lines = lines;
    return n;
}
