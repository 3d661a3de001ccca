/*
 * oracle/raster_ref.c -- CPU restatement of the tile rasterizer MANUS calls.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Not linked into the product.
 *
 * PARITY UNPINNED: the algorithm lives in a third-party dependency that is absent from
 * /root/reference -- graphdeco-inria/diff-gaussian-rasterization, cloned at an un-pinned HEAD by
 * /root/reference/setup_env.sh:6,9-10 (API window: 12-field GaussianRasterizationSettings, 2-tuple
 * return; the only call site is /root/reference/src/utils/gaussian_utils.py:378-416).  This file restates
 * its published algorithm as recorded in SURVEY.md Appendix A (A.1 preprocess, A.2 binning, A.3 blend,
 * A.4 backward, including the five places where the hand-written backward differs from autograd).
 * It is cross-checked by oracle/raster_autograd.py (independent differentiable restatement) and by the
 * analytic known-answer tests in tests/test_oracle_raster.py, not by upstream outputs.
 *
 * Build (oracle/Makefile):  REAL=float  -> libraster_ref_f32.so   (the checker; same precision as the GPU path)
 *                           REAL=double -> libraster_ref_f64.so   (error yard-stick)
 * All arithmetic is done in `real`; inputs/outputs are float arrays.  Matrices are the 16 floats exactly as
 * MANUS hands them over (row-vector convention, src/utils/cam_utils.py:58-64), read column-major m[4*col+row].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#ifndef REAL
#define REAL float
#endif
typedef REAL real;

#define TILE 16
#define NCH 3

static const real SH0 = (real)0.28209479177387814;
static const real SH1 = (real)0.4886025119029199;
static const real SH2[5] = {(real)1.0925484305920792, (real)-1.0925484305920792, (real)0.31539156525252005,
                            (real)-1.0925484305920792, (real)0.5462742152960396};
static const real SH3[7] = {(real)-0.5900435899266435, (real)2.890611442640554, (real)-0.4570457994644658,
                            (real)0.3731763325901154, (real)-0.4570457994644658, (real)1.445305721320277,
                            (real)-0.5900435899266435};

typedef struct {
    int N, W, H, gx, gy, M, deg;
    int has_sh, has_scale_rot;
    real view[16], proj[16], campos[3], bg[3];
    real tanx, tany, focx, focy, scale_mod;
    /* per-Gaussian state kept for backward */
    real *mean, *cov6, *rgb, *opac, *xy, *depth, *conic;
    int *radii, *clamped, *tiles;
    /* binning */
    int64_t D;
    uint64_t *keys;
    int32_t *list;
    int64_t *range; /* gx*gy*2 */
    /* image state */
    real *finalT;
    int32_t *ncontrib;
} ctx_t;

static real *ralloc(size_t n) { return (real *)calloc(n ? n : 1, sizeof(real)); }

void raster_ref_free(ctx_t *c) {
    if (!c) return;
    free(c->mean); free(c->cov6); free(c->rgb); free(c->opac); free(c->xy); free(c->depth); free(c->conic);
    free(c->radii); free(c->clamped); free(c->tiles); free(c->keys); free(c->list); free(c->range);
    free(c->finalT); free(c->ncontrib); free(c);
}

int raster_ref_real_bytes(void) { return (int)sizeof(real); }

/* ---- A.1 helpers ------------------------------------------------------------------------------- */
static void quat_to_R(const real q[4], real R[9]) { /* row-major standard rotation, q = (r,x,y,z), not normalised */
    real r = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - r * z);     R[2] = 2 * (x * z + r * y);
    R[3] = 2 * (x * y + r * z);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - r * x);
    R[6] = 2 * (x * z - r * y);     R[7] = 2 * (y * z + r * x);     R[8] = 1 - 2 * (x * x + y * y);
}

static void cov_from_scale_rot(const real s[3], real mod, const real q[4], real c6[6]) {
    real R[9], L[9];
    quat_to_R(q, R);
    for (int r = 0; r < 3; r++)
        for (int k = 0; k < 3; k++) L[3 * r + k] = R[3 * r + k] * (mod * s[k]); /* L = R diag(s) */
    real S[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            real a = 0;
            for (int k = 0; k < 3; k++) a += L[3 * i + k] * L[3 * j + k];
            S[3 * i + j] = a;
        }
    c6[0] = S[0]; c6[1] = S[1]; c6[2] = S[2]; c6[3] = S[4]; c6[4] = S[5]; c6[5] = S[8];
}

/* colour from SH, direction = mean - campos (world space); returns clamp flags */
static void sh_to_rgb(int deg, int M, const real *sh /* M*3 */, const real p[3], const real cam[3], real out[3], int cl[3]) {
    real d[3] = {p[0] - cam[0], p[1] - cam[1], p[2] - cam[2]};
    real inv = 1 / (real)sqrt((double)(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]));
    real x = d[0] * inv, y = d[1] * inv, z = d[2] * inv;
    (void)M;
    for (int c = 0; c < 3; c++) {
#define SHC(k) sh[(k) * 3 + c]
        real r = SH0 * SHC(0);
        if (deg > 0) {
            r = r - SH1 * y * SHC(1) + SH1 * z * SHC(2) - SH1 * x * SHC(3);
            if (deg > 1) {
                real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                r = r + SH2[0] * xy * SHC(4) + SH2[1] * yz * SHC(5) + SH2[2] * (2 * zz - xx - yy) * SHC(6) +
                    SH2[3] * xz * SHC(7) + SH2[4] * (xx - yy) * SHC(8);
                if (deg > 2) {
                    r = r + SH3[0] * y * (3 * xx - yy) * SHC(9) + SH3[1] * xy * z * SHC(10) +
                        SH3[2] * y * (4 * zz - xx - yy) * SHC(11) + SH3[3] * z * (2 * zz - 3 * xx - 3 * yy) * SHC(12) +
                        SH3[4] * x * (4 * zz - xx - yy) * SHC(13) + SH3[5] * z * (xx - yy) * SHC(14) +
                        SH3[6] * x * (xx - 3 * yy) * SHC(15);
                }
            }
        }
#undef SHC
        r += (real)0.5;
        cl[c] = r < 0;
        out[c] = r < 0 ? 0 : r;
    }
}

static int msb_bits(uint32_t n) { /* number of bits needed to index n values: position of highest set bit of n */
    int b = 0;
    while (n >> b) b++;
    return b;
}

/* stable LSD radix sort of (key, val) pairs on the low `bits` bits */
static void radix_sort_pairs(uint64_t *k, int32_t *v, int64_t n, int bits) {
    uint64_t *k2 = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
    int32_t *v2 = (int32_t *)malloc((n ? n : 1) * sizeof(int32_t));
    for (int sh = 0; sh < bits; sh += 8) {
        int64_t cnt[257] = {0};
        for (int64_t i = 0; i < n; i++) cnt[((k[i] >> sh) & 255) + 1]++;
        for (int b = 0; b < 256; b++) cnt[b + 1] += cnt[b];
        for (int64_t i = 0; i < n; i++) {
            int64_t d = cnt[(k[i] >> sh) & 255]++;
            k2[d] = k[i]; v2[d] = v[i];
        }
        uint64_t *tk = k; k = k2; k2 = tk;
        int32_t *tv = v; v = v2; v2 = tv;
    }
    /* after an odd number of passes the data sits in the scratch buffers: copy back */
    if (((bits + 7) / 8) & 1) { memcpy(k2, k, n * sizeof(uint64_t)); memcpy(v2, v, n * sizeof(int32_t)); free(k); free(v); }
    else { free(k2); free(v2); }
}

static void tile_rect(const ctx_t *c, real px, real py, int rad, int rmin[2], int rmax[2]) {
    int ax = (int)((px - rad) / TILE), ay = (int)((py - rad) / TILE);
    int bx = (int)((px + rad + TILE - 1) / TILE), by = (int)((py + rad + TILE - 1) / TILE);
    rmin[0] = ax < 0 ? 0 : (ax > c->gx ? c->gx : ax);
    rmin[1] = ay < 0 ? 0 : (ay > c->gy ? c->gy : ay);
    rmax[0] = bx < 0 ? 0 : (bx > c->gx ? c->gx : bx);
    rmax[1] = by < 0 ? 0 : (by > c->gy ? c->gy : by);
}

/* ---- A.3 blend of one tile ------------------------------------------------------------------------ */
static void blend_tile(ctx_t *c, int tl, float *out_color) {
    const int W = c->W, H = c->H;
    int tx0 = (tl % c->gx) * TILE, ty0 = (tl / c->gx) * TILE;
    int64_t r0 = c->range[2 * tl], r1 = c->range[2 * tl + 1];
    for (int ly = 0; ly < TILE; ly++)
        for (int lx = 0; lx < TILE; lx++) {
            int x = tx0 + lx, y = ty0 + ly;
            if (x >= W || y >= H) continue;
            real T = 1, C[3] = {0, 0, 0};
            int contributor = 0, last = 0;
            for (int64_t k = r0; k < r1; k++) {
                int g = c->list[k];
                contributor++;
                real dx = c->xy[2 * g] - (real)x, dy = c->xy[2 * g + 1] - (real)y;
                real power = (real)-0.5 * (c->conic[3 * g] * dx * dx + c->conic[3 * g + 2] * dy * dy) - c->conic[3 * g + 1] * dx * dy;
                if (power > 0) continue;
                real al = c->opac[g] * (real)(sizeof(real) == 4 ? expf((float)power) : exp((double)power));
                if (al > (real)0.99) al = (real)0.99;
                if (al < (real)(1.0 / 255.0)) continue;
                real Tn = T * (1 - al);
                if (Tn < (real)0.0001) break;
                for (int ch = 0; ch < 3; ch++) C[ch] += c->rgb[3 * (size_t)g + ch] * al * T;
                T = Tn; last = contributor;
            }
            size_t pix = (size_t)y * W + x;
            c->finalT[pix] = T; c->ncontrib[pix] = last;
            for (int ch = 0; ch < 3; ch++) out_color[(size_t)ch * W * H + pix] = (float)(C[ch] + T * c->bg[ch]);
        }
}

/* CPU-baseline timing only: blend just the tiles with (tile % stride) == offset (bench.py reports the sample). */
static int g_tile_stride = 1, g_tile_offset = 0;
void raster_ref_set_tile_sampling(int stride, int offset) { g_tile_stride = stride < 1 ? 1 : stride; g_tile_offset = offset; }
static int tile_selected(int t) { return g_tile_stride == 1 || (t % g_tile_stride) == g_tile_offset; }

typedef struct bwd_acc { real *gm2, *gcon, *gop, *gcol; } bwd_acc_t;
static void bwd_tile(const ctx_t *c, int tl, const float *dL_dout, bwd_acc_t *a);

typedef struct {
    ctx_t *c; float *out;                         /* forward */
    const float *dL_dout; bwd_acc_t *acc;         /* backward: one private accumulator set per thread */
    int next, ntile, backward, nthreads_started;
    pthread_mutex_t mu;
} tile_job_t;

static void *tile_worker(void *arg) {
    tile_job_t *j = (tile_job_t *)arg;
    pthread_mutex_lock(&j->mu);
    const int me = j->nthreads_started++;
    pthread_mutex_unlock(&j->mu);
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int t0 = j->next; j->next += 8;
        pthread_mutex_unlock(&j->mu);
        if (t0 >= j->ntile) break;
        for (int t = t0; t < t0 + 8 && t < j->ntile; t++) {
            if (!tile_selected(t)) continue;
            if (j->backward) bwd_tile(j->c, t, j->dL_dout, &j->acc[me]);
            else blend_tile(j->c, t, j->out);
        }
    }
    return NULL;
}

static void run_job(tile_job_t *j, int nthreads) {
    j->next = 0; j->nthreads_started = 0; j->ntile = j->c->gx * j->c->gy;
    pthread_mutex_init(&j->mu, NULL);
    if (nthreads <= 1) tile_worker(j);
    else {
        pthread_t th[256];
        for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, tile_worker, j);
        for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
    }
    pthread_mutex_destroy(&j->mu);
}

static void run_tiles(ctx_t *c, float *out_color, int nthreads) {
    tile_job_t j; memset(&j, 0, sizeof(j)); j.c = c; j.out = out_color;
    run_job(&j, nthreads > 256 ? 256 : nthreads);
}

/* ---- forward ----------------------------------------------------------------------------------- */
ctx_t *raster_ref_forward(int N, int W, int H, const float *means3D, const float *cov3D_precomp, const float *scales,
                          const float *rots, float scale_modifier, const float *colors_precomp, const float *shs,
                          int sh_degree, int M, const float *opacities, const float *viewmatrix, const float *projmatrix,
                          const float *campos, float tanfovx, float tanfovy, const float *bg, float *out_color /*3*H*W*/,
                          int32_t *out_radii /*N*/, int64_t *out_num_rendered, int nthreads) {
    ctx_t *c = (ctx_t *)calloc(1, sizeof(ctx_t));
    c->N = N; c->W = W; c->H = H; c->gx = (W + TILE - 1) / TILE; c->gy = (H + TILE - 1) / TILE;
    c->M = M; c->deg = sh_degree; c->has_sh = shs != NULL; c->has_scale_rot = cov3D_precomp == NULL;
    for (int i = 0; i < 16; i++) { c->view[i] = viewmatrix[i]; c->proj[i] = projmatrix[i]; }
    for (int i = 0; i < 3; i++) { c->campos[i] = campos[i]; c->bg[i] = bg[i]; }
    c->tanx = tanfovx; c->tany = tanfovy; c->scale_mod = scale_modifier;
    c->focx = W / (2 * c->tanx); c->focy = H / (2 * c->tany);
    c->mean = ralloc(3 * (size_t)N); c->cov6 = ralloc(6 * (size_t)N); c->rgb = ralloc(3 * (size_t)N); c->opac = ralloc(N);
    c->xy = ralloc(2 * (size_t)N); c->depth = ralloc(N); c->conic = ralloc(3 * (size_t)N);
    c->radii = (int *)calloc(N ? N : 1, sizeof(int)); c->clamped = (int *)calloc(3 * (size_t)N + 1, sizeof(int));
    c->tiles = (int *)calloc(N ? N : 1, sizeof(int));
    const real *v = c->view, *p = c->proj;
    const real limx = (real)1.3 * c->tanx, limy = (real)1.3 * c->tany;

    for (int i = 0; i < N; i++) {
        real m[3] = {means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]};
        c->mean[3 * i] = m[0]; c->mean[3 * i + 1] = m[1]; c->mean[3 * i + 2] = m[2];
        c->opac[i] = opacities[i];
        /* A.1 step 2: view-space point, near cull */
        real t[3] = {v[0] * m[0] + v[4] * m[1] + v[8] * m[2] + v[12], v[1] * m[0] + v[5] * m[1] + v[9] * m[2] + v[13],
                     v[2] * m[0] + v[6] * m[1] + v[10] * m[2] + v[14]};
        if (t[2] <= (real)0.2) continue;
        /* step 3 */
        real hx = p[0] * m[0] + p[4] * m[1] + p[8] * m[2] + p[12], hy = p[1] * m[0] + p[5] * m[1] + p[9] * m[2] + p[13];
        real hw = p[3] * m[0] + p[7] * m[1] + p[11] * m[2] + p[15];
        real pw = 1 / (hw + (real)0.0000001);
        real ndcx = hx * pw, ndcy = hy * pw;
        /* step 4 */
        real *c6 = c->cov6 + 6 * (size_t)i;
        if (cov3D_precomp) for (int k = 0; k < 6; k++) c6[k] = cov3D_precomp[6 * (size_t)i + k];
        else {
            real s[3] = {scales[3 * i], scales[3 * i + 1], scales[3 * i + 2]};
            real q[4] = {rots[4 * i], rots[4 * i + 1], rots[4 * i + 2], rots[4 * i + 3]};
            cov_from_scale_rot(s, c->scale_mod, q, c6);
        }
        /* step 5: EWA projection with the clamped view-space point */
        real tx = t[0] / t[2], ty = t[1] / t[2];
        tx = (tx < -limx ? -limx : (tx > limx ? limx : tx)) * t[2];
        ty = (ty < -limy ? -limy : (ty > limy ? limy : ty)) * t[2];
        real J00 = c->focx / t[2], J02 = -(c->focx * tx) / (t[2] * t[2]);
        real J11 = c->focy / t[2], J12 = -(c->focy * ty) / (t[2] * t[2]);
        real M0[3], M1[3]; /* rows of J.R, R[r][k] = v[4k + r] */
        for (int k = 0; k < 3; k++) { M0[k] = J00 * v[4 * k + 0] + J02 * v[4 * k + 2]; M1[k] = J11 * v[4 * k + 1] + J12 * v[4 * k + 2]; }
        real S[9] = {c6[0], c6[1], c6[2], c6[1], c6[3], c6[4], c6[2], c6[4], c6[5]};
        real SM0[3], SM1[3];
        for (int k = 0; k < 3; k++) { SM0[k] = S[3 * k] * M0[0] + S[3 * k + 1] * M0[1] + S[3 * k + 2] * M0[2];
                                      SM1[k] = S[3 * k] * M1[0] + S[3 * k + 1] * M1[1] + S[3 * k + 2] * M1[2]; }
        real a = M0[0] * SM0[0] + M0[1] * SM0[1] + M0[2] * SM0[2] + (real)0.3;
        real b = M0[0] * SM1[0] + M0[1] * SM1[1] + M0[2] * SM1[2];
        real cc = M1[0] * SM1[0] + M1[1] * SM1[1] + M1[2] * SM1[2] + (real)0.3;
        /* step 6-7 */
        real det = a * cc - b * b;
        if (det == 0) continue;
        real di = 1 / det;
        real mid = (real)0.5 * (a + cc);
        real disc = mid * mid - det; if (disc < (real)0.1) disc = (real)0.1;
        real sq = (real)sqrt((double)disc);
        real l1 = mid + sq, l2 = mid - sq;
        int rad = (int)ceil((double)((real)3 * (real)sqrt((double)(l1 > l2 ? l1 : l2))));
        /* step 8-9 */
        real px = ((ndcx + 1) * W - 1) * (real)0.5, py = ((ndcy + 1) * H - 1) * (real)0.5;
        int rmin[2], rmax[2];
        tile_rect(c, px, py, rad, rmin, rmax);
        int area = (rmax[0] - rmin[0]) * (rmax[1] - rmin[1]);
        if (area == 0) continue;
        /* step 10 */
        if (shs) {
            real shr[48 * 3];
            for (int k = 0; k < M * 3; k++) shr[k] = shs[(size_t)i * M * 3 + k];
            sh_to_rgb(sh_degree, M, shr, m, c->campos, c->rgb + 3 * (size_t)i, c->clamped + 3 * (size_t)i);
        } else for (int k = 0; k < 3; k++) c->rgb[3 * (size_t)i + k] = colors_precomp[3 * (size_t)i + k];
        c->depth[i] = t[2]; c->radii[i] = rad; c->xy[2 * i] = px; c->xy[2 * i + 1] = py;
        c->conic[3 * i] = cc * di; c->conic[3 * i + 1] = -b * di; c->conic[3 * i + 2] = a * di;
        c->tiles[i] = area;
    }
    for (int i = 0; i < N; i++) out_radii[i] = c->radii[i];

    /* A.2 binning */
    int64_t D = 0;
    for (int i = 0; i < N; i++) D += c->tiles[i];
    c->D = D; *out_num_rendered = D;
    c->keys = (uint64_t *)malloc((D ? D : 1) * sizeof(uint64_t));
    c->list = (int32_t *)malloc((D ? D : 1) * sizeof(int32_t));
    int64_t o = 0;
    for (int i = 0; i < N; i++) {
        if (c->radii[i] <= 0) continue;
        int rmin[2], rmax[2];
        tile_rect(c, c->xy[2 * i], c->xy[2 * i + 1], c->radii[i], rmin, rmax);
        float df = (float)c->depth[i]; uint32_t dbits; memcpy(&dbits, &df, 4);
        for (int y = rmin[1]; y < rmax[1]; y++)
            for (int x = rmin[0]; x < rmax[0]; x++) { c->keys[o] = ((uint64_t)(y * c->gx + x) << 32) | dbits; c->list[o] = i; o++; }
    }
    radix_sort_pairs(c->keys, c->list, D, 32 + msb_bits((uint32_t)(c->gx * c->gy)));
    int ntile = c->gx * c->gy;
    c->range = (int64_t *)calloc(2 * (size_t)ntile, sizeof(int64_t));
    for (int64_t i = 0; i < D; i++) {
        uint32_t tl = (uint32_t)(c->keys[i] >> 32);
        if (i == 0 || tl != (uint32_t)(c->keys[i - 1] >> 32)) c->range[2 * tl] = i;
        if (i == D - 1 || tl != (uint32_t)(c->keys[i + 1] >> 32)) c->range[2 * tl + 1] = i + 1;
    }

    /* A.3 blend (tiles are independent: spread over host threads for the CPU-baseline timing) */
    c->finalT = ralloc((size_t)W * H);
    c->ncontrib = (int32_t *)calloc((size_t)W * H, sizeof(int32_t));
    run_tiles(c, out_color, nthreads);
    return c;
}

void raster_ref_get_state(const ctx_t *c, float *xy, float *depth, float *conic, float *rgb, int32_t *tiles, float *finalT,
                          int32_t *ncontrib, int32_t *list, int64_t *range) {
    for (int i = 0; i < c->N; i++) {
        if (xy) { xy[2 * i] = (float)c->xy[2 * i]; xy[2 * i + 1] = (float)c->xy[2 * i + 1]; }
        if (depth) depth[i] = (float)c->depth[i];
        if (conic) for (int k = 0; k < 3; k++) conic[3 * i + k] = (float)c->conic[3 * i + k];
        if (rgb) for (int k = 0; k < 3; k++) rgb[3 * i + k] = (float)c->rgb[3 * i + k];
        if (tiles) tiles[i] = c->tiles[i];
    }
    size_t P = (size_t)c->W * c->H;
    if (finalT) for (size_t i = 0; i < P; i++) finalT[i] = (float)c->finalT[i];
    if (ncontrib) memcpy(ncontrib, c->ncontrib, P * sizeof(int32_t));
    if (list) memcpy(list, c->list, c->D * sizeof(int32_t));
    if (range) memcpy(range, c->range, 2 * (size_t)c->gx * c->gy * sizeof(int64_t));
}

/* ---- backward (A.4) ------------------------------------------------------------------------------ */
static void bwd_tile(const ctx_t *c, int tl, const float *dL_dout, bwd_acc_t *a) {
    const int W = c->W, H = c->H;
    int tx0 = (tl % c->gx) * TILE, ty0 = (tl / c->gx) * TILE;
    int64_t r0 = c->range[2 * tl], r1 = c->range[2 * tl + 1];
    if (r1 <= r0) return;
    for (int ly = 0; ly < TILE; ly++)
        for (int lx = 0; lx < TILE; lx++) {
            int x = tx0 + lx, y = ty0 + ly;
            if (x >= W || y >= H) continue;
            size_t pix = (size_t)y * W + x;
            const real Tfinal = c->finalT[pix];
            real T = Tfinal;
            const int last_contributor = c->ncontrib[pix];
            real dpix[3] = {dL_dout[pix], dL_dout[(size_t)W * H + pix], dL_dout[2 * (size_t)W * H + pix]};
            real accum[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0;
            int contributor = (int)(r1 - r0);
            for (int64_t k = r1 - 1; k >= r0; k--) {
                contributor--;
                if (contributor >= last_contributor) continue;
                int g = c->list[k];
                real dx = c->xy[2 * g] - (real)x, dy = c->xy[2 * g + 1] - (real)y;
                real cx = c->conic[3 * g], cy = c->conic[3 * g + 1], cz = c->conic[3 * g + 2], op = c->opac[g];
                real power = (real)-0.5 * (cx * dx * dx + cz * dy * dy) - cy * dx * dy;
                if (power > 0) continue;
                real G = (real)(sizeof(real) == 4 ? expf((float)power) : exp((double)power));
                real al = op * G; if (al > (real)0.99) al = (real)0.99;
                if (al < (real)(1.0 / 255.0)) continue;
                T = T / (1 - al);
                real dch = al * T, dL_dalpha = 0;
                for (int ch = 0; ch < 3; ch++) {
                    real col = c->rgb[3 * (size_t)g + ch];
                    accum[ch] = last_alpha * last_color[ch] + (1 - last_alpha) * accum[ch];
                    last_color[ch] = col;
                    dL_dalpha += (col - accum[ch]) * dpix[ch];
                    a->gcol[3 * (size_t)g + ch] += dch * dpix[ch];
                }
                dL_dalpha *= T;
                last_alpha = al;
                real bgdot = c->bg[0] * dpix[0] + c->bg[1] * dpix[1] + c->bg[2] * dpix[2];
                dL_dalpha += (-Tfinal / (1 - al)) * bgdot;
                real dL_dG = op * dL_dalpha; /* quirk 1: 0.99 clamp not masked */
                real gdx = G * dx, gdy = G * dy;
                real dG_ddelx = -gdx * cx - gdy * cy, dG_ddely = -gdy * cz - gdx * cy;
                a->gm2[2 * g] += dL_dG * dG_ddelx * ((real)0.5 * W);
                a->gm2[2 * g + 1] += dL_dG * dG_ddely * ((real)0.5 * H);
                a->gcon[3 * g] += (real)-0.5 * gdx * dx * dL_dG;
                a->gcon[3 * g + 1] += (real)-0.5 * gdx * dy * dL_dG;
                a->gcon[3 * g + 2] += (real)-0.5 * gdy * dy * dL_dG;
                a->gop[g] += G * dL_dalpha;
            }
        }
}

/* nthreads <= 1: single thread, tile order then pixel order (deterministic -- the checker).  nthreads > 1 (CPU-baseline
 * timing): tiles are spread over threads with private accumulators that are summed at the end. */
void raster_ref_backward(const ctx_t *c, const float *dL_dout /*3*H*W*/, const float *shs, const float *scales,
                         const float *rots, float *dL_dmean2D /*N*3*/, float *dL_dcolors /*N*3*/, float *dL_dopacity /*N*/,
                         float *dL_dmean3D /*N*3*/, float *dL_dcov3D /*N*6*/, float *dL_dsh /*N*M*3*/, float *dL_dscales /*N*3*/,
                         float *dL_drots /*N*4*/, float *dL_dconic_out /*N*3 (x,y,w) optional*/, int nthreads) {
    const int N = c->N;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    bwd_acc_t acc[64];
    for (int t = 0; t < nthreads; t++) {
        acc[t].gm2 = ralloc(2 * (size_t)N); acc[t].gcon = ralloc(3 * (size_t)N); acc[t].gop = ralloc(N); acc[t].gcol = ralloc(3 * (size_t)N);
    }
    tile_job_t j; memset(&j, 0, sizeof(j)); j.c = (ctx_t *)c; j.backward = 1; j.dL_dout = dL_dout; j.acc = acc;
    run_job(&j, nthreads);
    real *gm2 = acc[0].gm2, *gcon = acc[0].gcon, *gop = acc[0].gop, *gcol = acc[0].gcol;
    for (int t = 1; t < nthreads; t++) {
        for (size_t i = 0; i < 2 * (size_t)N; i++) gm2[i] += acc[t].gm2[i];
        for (size_t i = 0; i < 3 * (size_t)N; i++) { gcon[i] += acc[t].gcon[i]; gcol[i] += acc[t].gcol[i]; }
        for (size_t i = 0; i < (size_t)N; i++) gop[i] += acc[t].gop[i];
        free(acc[t].gm2); free(acc[t].gcon); free(acc[t].gop); free(acc[t].gcol);
    }

    const real *v = c->view, *p = c->proj;
    const real limx = (real)1.3 * c->tanx, limy = (real)1.3 * c->tany;
    for (int i = 0; i < N; i++) {
        real gmean[3] = {0, 0, 0}, gcov[6] = {0, 0, 0, 0, 0, 0};
        for (int k = 0; k < 3; k++) dL_dmean2D[3 * i + k] = 0;
        for (int k = 0; k < 3; k++) dL_dcolors[3 * i + k] = 0;
        dL_dopacity[i] = 0;
        if (dL_dscales) for (int k = 0; k < 3; k++) dL_dscales[3 * i + k] = 0;
        if (dL_drots) for (int k = 0; k < 4; k++) dL_drots[4 * i + k] = 0;
        if (dL_dsh) for (int k = 0; k < c->M * 3; k++) dL_dsh[(size_t)i * c->M * 3 + k] = 0;
        if (dL_dconic_out) for (int k = 0; k < 3; k++) dL_dconic_out[3 * i + k] = (float)gcon[3 * i + k];
        if (c->radii[i] > 0) {
            const real *m = c->mean + 3 * (size_t)i, *c6 = c->cov6 + 6 * (size_t)i;
            /* cov2D backward */
            real t[3] = {v[0] * m[0] + v[4] * m[1] + v[8] * m[2] + v[12], v[1] * m[0] + v[5] * m[1] + v[9] * m[2] + v[13],
                         v[2] * m[0] + v[6] * m[1] + v[10] * m[2] + v[14]};
            real txtz = t[0] / t[2], tytz = t[1] / t[2];
            real xm = (txtz < -limx || txtz > limx) ? 0 : 1, ym = (tytz < -limy || tytz > limy) ? 0 : 1;
            real tx = (txtz < -limx ? -limx : (txtz > limx ? limx : txtz)) * t[2];
            real ty = (tytz < -limy ? -limy : (tytz > limy ? limy : tytz)) * t[2];
            real J00 = c->focx / t[2], J02 = -(c->focx * tx) / (t[2] * t[2]);
            real J11 = c->focy / t[2], J12 = -(c->focy * ty) / (t[2] * t[2]);
            real M0[3], M1[3];
            for (int k = 0; k < 3; k++) { M0[k] = J00 * v[4 * k + 0] + J02 * v[4 * k + 2]; M1[k] = J11 * v[4 * k + 1] + J12 * v[4 * k + 2]; }
            real S[9] = {c6[0], c6[1], c6[2], c6[1], c6[3], c6[4], c6[2], c6[4], c6[5]};
            real SM0[3], SM1[3];
            for (int k = 0; k < 3; k++) { SM0[k] = S[3 * k] * M0[0] + S[3 * k + 1] * M0[1] + S[3 * k + 2] * M0[2];
                                          SM1[k] = S[3 * k] * M1[0] + S[3 * k + 1] * M1[1] + S[3 * k + 2] * M1[2]; }
            real a = M0[0] * SM0[0] + M0[1] * SM0[1] + M0[2] * SM0[2] + (real)0.3;
            real b = M0[0] * SM1[0] + M0[1] * SM1[1] + M0[2] * SM1[2];
            real cc = M1[0] * SM1[0] + M1[1] * SM1[1] + M1[2] * SM1[2] + (real)0.3;
            real denom = a * cc - b * b;
            real d2 = 1 / (denom * denom + (real)0.0000001); /* quirk 3 */
            real gx = gcon[3 * i], gy = gcon[3 * i + 1], gz = gcon[3 * i + 2];
            real dL_da = 0, dL_db = 0, dL_dc = 0;
            if (d2 != 0) {
                dL_da = d2 * (-cc * cc * gx + 2 * b * cc * gy + (denom - a * cc) * gz);
                dL_dc = d2 * (-a * a * gz + 2 * a * b * gy + (denom - a * cc) * gx);
                dL_db = d2 * 2 * (b * cc * gx - (denom + 2 * b * b) * gy + a * b * gz);
                gcov[0] = M0[0] * M0[0] * dL_da + M0[0] * M1[0] * dL_db + M1[0] * M1[0] * dL_dc;
                gcov[3] = M0[1] * M0[1] * dL_da + M0[1] * M1[1] * dL_db + M1[1] * M1[1] * dL_dc;
                gcov[5] = M0[2] * M0[2] * dL_da + M0[2] * M1[2] * dL_db + M1[2] * M1[2] * dL_dc;
                gcov[1] = 2 * M0[0] * M0[1] * dL_da + (M0[0] * M1[1] + M0[1] * M1[0]) * dL_db + 2 * M1[0] * M1[1] * dL_dc;
                gcov[2] = 2 * M0[0] * M0[2] * dL_da + (M0[0] * M1[2] + M0[2] * M1[0]) * dL_db + 2 * M1[0] * M1[2] * dL_dc;
                gcov[4] = 2 * M0[2] * M0[1] * dL_da + (M0[1] * M1[2] + M0[2] * M1[1]) * dL_db + 2 * M1[1] * M1[2] * dL_dc;
            }
            real dM0[3], dM1[3];
            for (int k = 0; k < 3; k++) { dM0[k] = 2 * SM0[k] * dL_da + SM1[k] * dL_db; dM1[k] = 2 * SM1[k] * dL_dc + SM0[k] * dL_db; }
            real dJ00 = v[0] * dM0[0] + v[4] * dM0[1] + v[8] * dM0[2];
            real dJ02 = v[2] * dM0[0] + v[6] * dM0[1] + v[10] * dM0[2];
            real dJ11 = v[1] * dM1[0] + v[5] * dM1[1] + v[9] * dM1[2];
            real dJ12 = v[2] * dM1[0] + v[6] * dM1[1] + v[10] * dM1[2];
            real tz = 1 / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
            real dtx = xm * -c->focx * tz2 * dJ02, dty = ym * -c->focy * tz2 * dJ12; /* quirk 2 */
            real dtz = -c->focx * tz2 * dJ00 - c->focy * tz2 * dJ11 + (2 * c->focx * tx) * tz3 * dJ02 + (2 * c->focy * ty) * tz3 * dJ12;
            gmean[0] = v[0] * dtx + v[1] * dty + v[2] * dtz;
            gmean[1] = v[4] * dtx + v[5] * dty + v[6] * dtz;
            gmean[2] = v[8] * dtx + v[9] * dty + v[10] * dtz;
            /* projection backward */
            real hx = p[0] * m[0] + p[4] * m[1] + p[8] * m[2] + p[12], hy = p[1] * m[0] + p[5] * m[1] + p[9] * m[2] + p[13];
            real hw = p[3] * m[0] + p[7] * m[1] + p[11] * m[2] + p[15];
            real mw = 1 / (hw + (real)0.0000001);
            real mul1 = hx * mw * mw, mul2 = hy * mw * mw;
            real g2x = gm2[2 * i], g2y = gm2[2 * i + 1];
            gmean[0] += (p[0] * mw - p[3] * mul1) * g2x + (p[1] * mw - p[3] * mul2) * g2y;
            gmean[1] += (p[4] * mw - p[7] * mul1) * g2x + (p[5] * mw - p[7] * mul2) * g2y;
            gmean[2] += (p[8] * mw - p[11] * mul1) * g2x + (p[9] * mw - p[11] * mul2) * g2y;
            dL_dmean2D[3 * i] = (float)g2x; dL_dmean2D[3 * i + 1] = (float)g2y;
            dL_dopacity[i] = (float)gop[i];
            real gc[3] = {gcol[3 * i], gcol[3 * i + 1], gcol[3 * i + 2]};

            if (c->has_sh && shs && dL_dsh) {
                /* exact derivative of sh_to_rgb; clamped channels get no gradient */
                real d[3] = {m[0] - c->campos[0], m[1] - c->campos[1], m[2] - c->campos[2]};
                real nn = (real)sqrt((double)(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]));
                real inv = 1 / nn, x = d[0] * inv, y = d[1] * inv, z = d[2] * inv;
                real gd[3] = {0, 0, 0}; /* dL/d(unit dir) */
                const int deg = c->deg, Mc = c->M;
                for (int ch = 0; ch < 3; ch++) {
                    real go = c->clamped[3 * i + ch] ? 0 : gc[ch];
                    float *o = dL_dsh + (size_t)i * Mc * 3;
#define SHC(k) ((real)shs[((size_t)i * Mc + (k)) * 3 + ch])
#define DSH(k, val) o[(k) * 3 + ch] = (float)((val) * go)
                    DSH(0, SH0);
                    if (deg > 0) {
                        DSH(1, -SH1 * y); DSH(2, SH1 * z); DSH(3, -SH1 * x);
                        real dx_ = -SH1 * SHC(3), dy_ = -SH1 * SHC(1), dz_ = SH1 * SHC(2);
                        if (deg > 1) {
                            real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                            DSH(4, SH2[0] * xy); DSH(5, SH2[1] * yz); DSH(6, SH2[2] * (2 * zz - xx - yy));
                            DSH(7, SH2[3] * xz); DSH(8, SH2[4] * (xx - yy));
                            dx_ += SH2[0] * y * SHC(4) + SH2[2] * 2 * -x * SHC(6) + SH2[3] * z * SHC(7) + SH2[4] * 2 * x * SHC(8);
                            dy_ += SH2[0] * x * SHC(4) + SH2[1] * z * SHC(5) + SH2[2] * 2 * -y * SHC(6) + SH2[4] * 2 * -y * SHC(8);
                            dz_ += SH2[1] * y * SHC(5) + SH2[2] * 2 * 2 * z * SHC(6) + SH2[3] * x * SHC(7);
                            if (deg > 2) {
                                DSH(9, SH3[0] * y * (3 * xx - yy)); DSH(10, SH3[1] * xy * z);
                                DSH(11, SH3[2] * y * (4 * zz - xx - yy)); DSH(12, SH3[3] * z * (2 * zz - 3 * xx - 3 * yy));
                                DSH(13, SH3[4] * x * (4 * zz - xx - yy)); DSH(14, SH3[5] * z * (xx - yy));
                                DSH(15, SH3[6] * x * (xx - 3 * yy));
                                dx_ += SH3[0] * SHC(9) * 3 * 2 * xy + SH3[1] * SHC(10) * yz + SH3[2] * SHC(11) * -2 * xy +
                                       SH3[3] * SHC(12) * -3 * 2 * xz + SH3[4] * SHC(13) * (-3 * xx + 4 * zz - yy) +
                                       SH3[5] * SHC(14) * 2 * xz + SH3[6] * SHC(15) * 3 * (xx - yy);
                                dy_ += SH3[0] * SHC(9) * 3 * (xx - yy) + SH3[1] * SHC(10) * xz + SH3[2] * SHC(11) * (-3 * yy + 4 * zz - xx) +
                                       SH3[3] * SHC(12) * -3 * 2 * yz + SH3[4] * SHC(13) * -2 * xy + SH3[5] * SHC(14) * -2 * yz +
                                       SH3[6] * SHC(15) * -3 * 2 * xy;
                                dz_ += SH3[1] * SHC(10) * xy + SH3[2] * SHC(11) * 4 * 2 * yz + SH3[3] * SHC(12) * 3 * (2 * zz - xx - yy) +
                                       SH3[4] * SHC(13) * 4 * 2 * xz + SH3[5] * SHC(14) * (xx - yy);
                            }
                        }
                        gd[0] += dx_ * go; gd[1] += dy_ * go; gd[2] += dz_ * go;
                    }
#undef SHC
#undef DSH
                }
                /* through the normalisation d/|d| */
                real dot = (x * gd[0] + y * gd[1] + z * gd[2]);
                gmean[0] += (gd[0] - x * dot) * inv; gmean[1] += (gd[1] - y * dot) * inv; gmean[2] += (gd[2] - z * dot) * inv;
            } else {
                for (int k = 0; k < 3; k++) dL_dcolors[3 * i + k] = (float)gc[k];
            }

            if (c->has_scale_rot && scales && rots && dL_dscales && dL_drots) {
                /* Sigma = L L^T, L = R diag(mod*s): exact derivative w.r.t. s and raw q */
                real s[3] = {c->scale_mod * scales[3 * i], c->scale_mod * scales[3 * i + 1], c->scale_mod * scales[3 * i + 2]};
                real q[4] = {rots[4 * i], rots[4 * i + 1], rots[4 * i + 2], rots[4 * i + 3]};
                real R[9]; quat_to_R(q, R);
                /* symmetric gradient matrix: dL/dSigma_full, off-diagonals split in half */
                real Gm[9] = {gcov[0], (real)0.5 * gcov[1], (real)0.5 * gcov[2], (real)0.5 * gcov[1], gcov[3], (real)0.5 * gcov[4],
                              (real)0.5 * gcov[2], (real)0.5 * gcov[4], gcov[5]};
                /* dL/dL = 2 Gm L */
                real L[9], dLm[9];
                for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) L[3 * r + k] = R[3 * r + k] * s[k];
                for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) {
                    real acc = 0; for (int j = 0; j < 3; j++) acc += Gm[3 * r + j] * L[3 * j + k];
                    dLm[3 * r + k] = 2 * acc;
                }
                real dR[9];
                for (int k = 0; k < 3; k++) {
                    real acc = 0; for (int r = 0; r < 3; r++) acc += dLm[3 * r + k] * R[3 * r + k];
                    dL_dscales[3 * i + k] = (float)(acc * c->scale_mod);
                    for (int r = 0; r < 3; r++) dR[3 * r + k] = dLm[3 * r + k] * s[k];
                }
                real r_ = q[0], x = q[1], y = q[2], z = q[3];
                real dq0 = 2 * (-z * dR[1] + y * dR[2] + z * dR[3] - x * dR[5] - y * dR[6] + x * dR[7]);
                real dq1 = 2 * (y * dR[1] + z * dR[2] + y * dR[3] - 2 * x * dR[4] - r_ * dR[5] + z * dR[6] + r_ * dR[7] - 2 * x * dR[8]);
                real dq2 = 2 * (-2 * y * dR[0] + x * dR[1] + r_ * dR[2] + x * dR[3] + z * dR[5] - r_ * dR[6] + z * dR[7] - 2 * y * dR[8]);
                real dq3 = 2 * (-2 * z * dR[0] - r_ * dR[1] + x * dR[2] + r_ * dR[3] - 2 * z * dR[4] + y * dR[5] + x * dR[6] + y * dR[7]);
                dL_drots[4 * i] = (float)dq0; dL_drots[4 * i + 1] = (float)dq1; dL_drots[4 * i + 2] = (float)dq2; dL_drots[4 * i + 3] = (float)dq3;
            }
        }
        for (int k = 0; k < 3; k++) dL_dmean3D[3 * i + k] = (float)gmean[k];
        for (int k = 0; k < 6; k++) dL_dcov3D[6 * i + k] = (float)gcov[k];
    }
    free(gm2); free(gcon); free(gop); free(gcol);
}

/* checkFrustum / markVisible: in front of the near plane (p_view.z > 0.2) */
void raster_ref_mark_visible(int N, const float *means3D, const float *viewmatrix, const float *projmatrix, uint8_t *out) {
    (void)projmatrix;
    for (int i = 0; i < N; i++) {
        real m[3] = {means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]};
        real z = (real)viewmatrix[2] * m[0] + (real)viewmatrix[6] * m[1] + (real)viewmatrix[10] * m[2] + (real)viewmatrix[14];
        out[i] = z > (real)0.2;
    }
}
