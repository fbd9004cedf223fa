/*
 * densecrf_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * PARITY UNPINNED: the arithmetic this file restates lives in the third-party
 * package `pydensecrf` (git+https://github.com/lucasb-eyer/pydensecrf.git, un-pinned in the
 * reference's requirements.txt:11; it wraps Kraehenbuehl's densecrf v2 C++/Eigen).  That package
 * is not vendored under /root/reference, is not installed in this image and cannot be fetched, and
 * the reference ships no test / golden vector for this path.  This file is
 * therefore a from-scratch restatement of the *published* algorithm (Adams, Baek & Davis 2010,
 * "Fast high-dimensional filtering using the permutohedral lattice"; Kraehenbuehl & Koltun 2011,
 * "Efficient inference in fully connected CRFs with Gaussian edge potentials") following the
 * arithmetic specification in SURVEY.md Appendix A, anchored on the reference's own call sites:
 *     03c_hsn/utilities.py:427-443          (DenseCRF2D / setUnaryEnergy / addPairwise* / inference)
 *     03a_sec-dsrg/SEC.py:270-280           (crf_inference batch closure)
 *     03b_irn/step/cam_to_ir_label.py:35-67 (crf_inference_label)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (wsss_analysis_b200) never does.
 *
 * Build:  gcc -O3 -ffp-contract=off -fno-fast-math -fPIC -shared  (see oracle/Makefile)
 * Float order is defined: no FMA contraction, SSE2 scalar float math (x86-64 default).
 *
 * Layout conventions (Appendix A.1):
 *   - pixel index p = y*W + x, N = W*H
 *   - Python-side unary / Q are row-major (L, N)
 *   - internal matrices are "pixel-major": L contiguous labels per pixel (Eigen column-major (L,N))
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* enums: same numeric values as the product C-ABI (include/dcrf_b200.h)                       */
/* ------------------------------------------------------------------------------------------ */
enum { ORC_CONST_KERNEL = 0, ORC_DIAG_KERNEL = 1, ORC_FULL_KERNEL = 2 };
enum { ORC_NO_NORMALIZATION = 0, ORC_NORMALIZE_BEFORE = 1, ORC_NORMALIZE_AFTER = 2,
       ORC_NORMALIZE_SYMMETRIC = 3 };
enum { ORC_COMPAT_POTTS = 0, ORC_COMPAT_DIAGONAL = 1, ORC_COMPAT_MATRIX = 2 };

/* ------------------------------------------------------------------------------------------ */
/* Sequential hash of int16 keys.  Appendix A.3 step 8: only "id = rank of the key's first      */
/* occurrence in p-major / remainder-minor scan order" is observable, not the hash function.    */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int key_size;
    size_t filled, capacity; /* capacity is a power of two */
    int16_t *keys;           /* filled * key_size */
    size_t keys_cap;         /* in keys */
    int32_t *table;          /* capacity, -1 = empty */
} orc_hash;

static size_t orc_hash_fn(const orc_hash *h, const int16_t *k) {
    uint64_t r = 1469598103934665603ull;
    for (int i = 0; i < h->key_size; i++) {
        r ^= (uint16_t)k[i];
        r *= 1099511628211ull;
    }
    r ^= r >> 29;
    return (size_t)r & (h->capacity - 1);
}

static void orc_hash_init(orc_hash *h, int key_size, size_t n_expected) {
    h->key_size = key_size;
    h->filled = 0;
    h->capacity = 16;
    while (h->capacity < 2 * n_expected) h->capacity *= 2;
    h->keys_cap = n_expected > 16 ? n_expected : 16;
    h->keys = (int16_t *)malloc(h->keys_cap * (size_t)key_size * sizeof(int16_t));
    h->table = (int32_t *)malloc(h->capacity * sizeof(int32_t));
    for (size_t i = 0; i < h->capacity; i++) h->table[i] = -1;
}

static void orc_hash_free(orc_hash *h) {
    free(h->keys);
    free(h->table);
}

static void orc_hash_grow(orc_hash *h) {
    size_t old_cap = h->capacity;
    int32_t *old = h->table;
    h->capacity *= 2;
    h->table = (int32_t *)malloc(h->capacity * sizeof(int32_t));
    for (size_t i = 0; i < h->capacity; i++) h->table[i] = -1;
    for (size_t i = 0; i < old_cap; i++) {
        int32_t e = old[i];
        if (e >= 0) {
            size_t s = orc_hash_fn(h, h->keys + (size_t)e * h->key_size);
            while (h->table[s] >= 0) s = (s + 1) & (h->capacity - 1);
            h->table[s] = e;
        }
    }
    free(old);
}

/* returns the id of key k; when absent returns -1, or inserts it with id = #keys so far */
static int32_t orc_hash_find(orc_hash *h, const int16_t *k, int create) {
    if (create && 2 * h->filled >= h->capacity) orc_hash_grow(h);
    size_t s = orc_hash_fn(h, k);
    for (;;) {
        int32_t e = h->table[s];
        if (e < 0) {
            if (!create) return -1;
            if (h->filled == h->keys_cap) {
                h->keys_cap *= 2;
                h->keys = (int16_t *)realloc(h->keys, h->keys_cap * (size_t)h->key_size * sizeof(int16_t));
            }
            memcpy(h->keys + h->filled * h->key_size, k, (size_t)h->key_size * sizeof(int16_t));
            h->table[s] = (int32_t)h->filled;
            return (int32_t)h->filled++;
        }
        if (memcmp(h->keys + (size_t)e * h->key_size, k, (size_t)h->key_size * sizeof(int16_t)) == 0)
            return e;
        s = (s + 1) & (h->capacity - 1);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Permutohedral lattice (Appendix A.3 / A.4)                                                   */
/* ------------------------------------------------------------------------------------------ */
typedef struct orc_lattice {
    int N, d, M;
    int32_t *offset;      /* N*(d+1) vertex ids */
    int16_t *rank;        /* N*(d+1) */
    float *barycentric;   /* N*(d+1) */
    int32_t *neigh;       /* (d+1)*M*2 : [j*M+i] -> (n1, n2), -1 = absent */
    int16_t *keys;        /* M*d */
} orc_lattice;

/* Permutohedral::init [EXT], Appendix A.3 steps 1-9; reached from addPairwiseGaussian / Bilateral
 * (/root/reference/03c_hsn/utilities.py:435, :439-440).
 * feature: N x d, pixel-major (Eigen (d,N) column-major) */
orc_lattice *orc_lattice_create(const float *feature, int N, int d) {
    orc_lattice *lat = (orc_lattice *)calloc(1, sizeof(orc_lattice));
    lat->N = N;
    lat->d = d;
    const int d1 = d + 1;
    lat->offset = (int32_t *)malloc((size_t)N * d1 * sizeof(int32_t) + 4);
    lat->rank = (int16_t *)malloc((size_t)N * d1 * sizeof(int16_t) + 4);
    lat->barycentric = (float *)malloc((size_t)N * d1 * sizeof(float) + 4);

    orc_hash ht;
    orc_hash_init(&ht, d, (size_t)N * d1);

    float *scale_factor = (float *)malloc(sizeof(float) * (d > 0 ? d : 1));
    float *elevated = (float *)malloc(sizeof(float) * d1);
    float *rem0 = (float *)malloc(sizeof(float) * d1);
    float *bary = (float *)malloc(sizeof(float) * (d + 2));
    short *rank = (short *)malloc(sizeof(short) * d1);
    short *canonical = (short *)malloc(sizeof(short) * d1 * d1);
    int16_t *key = (int16_t *)malloc(sizeof(int16_t) * d1);

    /* A.3 step 7: canonical simplex */
    for (int i = 0; i <= d; i++) {
        for (int j = 0; j <= d - i; j++) canonical[i * d1 + j] = (short)i;
        for (int j = d - i + 1; j <= d; j++) canonical[i * d1 + j] = (short)(i - d1);
    }
    /* A.3 step 1: double math, stored as float */
    float inv_std_dev = (float)(sqrt(2.0 / 3.0) * (double)d1);
    for (int i = 0; i < d; i++)
        scale_factor[i] = (float)(1.0 / sqrt((double)((i + 2) * (i + 1))) * (double)inv_std_dev);

    for (int k = 0; k < N; k++) {
        const float *f = feature + (size_t)k * d;
        /* A.3 step 2: elevate */
        float sm = 0;
        for (int j = d; j > 0; j--) {
            float cf = f[j - 1] * scale_factor[j - 1];
            elevated[j] = sm - (float)j * cf;
            sm += cf;
        }
        elevated[0] = sm;

        /* A.3 step 3: nearest remainder-0 point; `sum` is an int accumulator of float terms */
        float down_factor = 1.0f / (float)d1;
        float up_factor = (float)d1;
        int sum = 0;
        for (int i = 0; i <= d; i++) {
            int rd2;
            float v = down_factor * elevated[i];
            float up = ceilf(v) * up_factor;
            float down = floorf(v) * up_factor;
            if (up - elevated[i] < elevated[i] - down) rd2 = (short)up;
            else rd2 = (short)down;
            rem0[i] = (float)rd2;
            sum = (int)((float)sum + (float)rd2 * down_factor);
        }

        /* A.3 step 4: rank */
        for (int i = 0; i <= d; i++) rank[i] = 0;
        for (int i = 0; i < d; i++) {
            double di = (double)(elevated[i] - rem0[i]);
            for (int j = i + 1; j <= d; j++) {
                if (di < (double)(elevated[j] - rem0[j])) rank[i]++;
                else rank[j]++;
            }
        }

        /* A.3 step 5: re-project onto the plane */
        for (int i = 0; i <= d; i++) {
            rank[i] = (short)(rank[i] + sum);
            if (rank[i] < 0) {
                rank[i] = (short)(rank[i] + d1);
                rem0[i] += (float)d1;
            } else if (rank[i] > d) {
                rank[i] = (short)(rank[i] - d1);
                rem0[i] -= (float)d1;
            }
        }

        /* A.3 step 6: barycentric weights */
        for (int i = 0; i <= d + 1; i++) bary[i] = 0;
        for (int i = 0; i <= d; i++) {
            float v = (elevated[i] - rem0[i]) * down_factor;
            bary[d - rank[i]] += v;
            bary[d - rank[i] + 1] -= v;
        }
        bary[0] = (float)((double)bary[0] + (1.0 + (double)bary[d + 1]));

        /* A.3 step 7/8: the d+1 simplex vertices, ids by first insertion */
        for (int r = 0; r <= d; r++) {
            for (int i = 0; i < d; i++)
                key[i] = (int16_t)(rem0[i] + (float)canonical[r * d1 + rank[i]]);
            lat->offset[(size_t)k * d1 + r] = orc_hash_find(&ht, key, 1);
            lat->rank[(size_t)k * d1 + r] = rank[r];
            lat->barycentric[(size_t)k * d1 + r] = bary[r];
        }
    }

    /* A.3 step 9: blur neighbours */
    const int M = (int)ht.filled;
    lat->M = M;
    lat->neigh = (int32_t *)malloc((size_t)d1 * (M > 0 ? M : 1) * 2 * sizeof(int32_t));
    lat->keys = (int16_t *)malloc((size_t)(M > 0 ? M : 1) * (d > 0 ? d : 1) * sizeof(int16_t));
    memcpy(lat->keys, ht.keys, (size_t)M * d * sizeof(int16_t));
    int16_t *n1 = (int16_t *)malloc(sizeof(int16_t) * d1);
    int16_t *n2 = (int16_t *)malloc(sizeof(int16_t) * d1);
    for (int j = 0; j <= d; j++) {
        for (int i = 0; i < M; i++) {
            const int16_t *kk = lat->keys + (size_t)i * d;
            for (int k = 0; k < d; k++) {
                n1[k] = (int16_t)(kk[k] - 1);
                n2[k] = (int16_t)(kk[k] + 1);
            }
            if (j < d) { /* for j == d no stored coordinate carries the exception */
                n1[j] = (int16_t)(kk[j] + d);
                n2[j] = (int16_t)(kk[j] - d);
            }
            lat->neigh[((size_t)j * M + i) * 2 + 0] = orc_hash_find(&ht, n1, 0);
            lat->neigh[((size_t)j * M + i) * 2 + 1] = orc_hash_find(&ht, n2, 0);
        }
    }
    free(n1); free(n2);
    free(scale_factor); free(elevated); free(rem0); free(bary); free(rank); free(canonical); free(key);
    orc_hash_free(&ht);
    return lat;
}

void orc_lattice_free(orc_lattice *lat) {
    if (!lat) return;
    free(lat->offset); free(lat->rank); free(lat->barycentric); free(lat->neigh); free(lat->keys);
    free(lat);
}

int orc_lattice_M(const orc_lattice *lat) { return lat->M; }
int orc_lattice_d(const orc_lattice *lat) { return lat->d; }
int orc_lattice_N(const orc_lattice *lat) { return lat->N; }

/* any output pointer may be NULL */
void orc_lattice_export(const orc_lattice *lat, int16_t *keys, int32_t *offsets, float *bary,
                        int32_t *neigh, int16_t *rank) {
    const size_t E = (size_t)lat->N * (lat->d + 1);
    if (keys) memcpy(keys, lat->keys, (size_t)lat->M * lat->d * sizeof(int16_t));
    if (offsets) memcpy(offsets, lat->offset, E * sizeof(int32_t));
    if (bary) memcpy(bary, lat->barycentric, E * sizeof(float));
    if (neigh) memcpy(neigh, lat->neigh, (size_t)(lat->d + 1) * lat->M * 2 * sizeof(int32_t));
    if (rank) memcpy(rank, lat->rank, E * sizeof(int16_t));
}

/*
 * Permutohedral::compute [EXT] (splat / blur / slice), run twice per iteration inside
 * `d.inference(n)` (/root/reference/03c_hsn/utilities.py:442).
 * A.4 filter.  in/out: N x value_size pixel-major (may alias).
 * Two association variants exist upstream: a scalar path used when value_size <= 2 (blur through a
 * double 0.5, slice as (w*v)*alpha) and a 4-wide float path used otherwise (blur in float,
 * slice as (w*alpha)*v).  Both are restated; `value_size <= 2` selects, as upstream does.
 */
void orc_lattice_compute(const orc_lattice *lat, float *out, const float *in, int value_size,
                         int reverse) {
    const int N = lat->N, d = lat->d, M = lat->M, d1 = d + 1, vs = value_size;
    const int seq = (vs <= 2);
    float *buf_a = (float *)calloc((size_t)(M + 2) * vs, sizeof(float));
    float *buf_b = (float *)calloc((size_t)(M + 2) * vs, sizeof(float));
    float *values = buf_a, *new_values = buf_b;

    /* splat: pixel order, remainder order */
    for (int i = 0; i < N; i++) {
        for (int j = 0; j <= d; j++) {
            int o = lat->offset[(size_t)i * d1 + j] + 1;
            float w = lat->barycentric[(size_t)i * d1 + j];
            float *v = values + (size_t)o * vs;
            const float *src = in + (size_t)i * vs;
            for (int k = 0; k < vs; k++) v[k] += w * src[k];
        }
    }
    /* blur: Jacobi within a direction, sequential across directions */
    for (int j = reverse ? d : 0; j <= d && j >= 0; reverse ? j-- : j++) {
        for (int i = 0; i < M; i++) {
            const float *old_val = values + (size_t)(i + 1) * vs;
            float *new_val = new_values + (size_t)(i + 1) * vs;
            int n1 = lat->neigh[((size_t)j * M + i) * 2 + 0] + 1;
            int n2 = lat->neigh[((size_t)j * M + i) * 2 + 1] + 1;
            const float *n1_val = values + (size_t)n1 * vs;
            const float *n2_val = values + (size_t)n2 * vs;
            if (seq) {
                for (int k = 0; k < vs; k++)
                    new_val[k] = (float)((double)old_val[k] + 0.5 * (double)(n1_val[k] + n2_val[k]));
            } else {
                for (int k = 0; k < vs; k++)
                    new_val[k] = old_val[k] + 0.5f * (n1_val[k] + n2_val[k]);
            }
        }
        float *t = values; values = new_values; new_values = t;
    }
    /* slice */
    float alpha = 1.0f / (1.0f + powf(2.0f, (float)-d));
    float *acc = (float *)malloc(sizeof(float) * (vs > 0 ? vs : 1));
    for (int i = 0; i < N; i++) {
        for (int k = 0; k < vs; k++) acc[k] = 0;
        for (int j = 0; j <= d; j++) {
            int o = lat->offset[(size_t)i * d1 + j] + 1;
            float w = lat->barycentric[(size_t)i * d1 + j];
            const float *v = values + (size_t)o * vs;
            if (seq) {
                for (int k = 0; k < vs; k++) acc[k] += w * v[k] * alpha;
            } else {
                float wa = w * alpha;
                for (int k = 0; k < vs; k++) acc[k] += wa * v[k];
            }
        }
        memcpy(out + (size_t)i * vs, acc, sizeof(float) * vs);
    }
    free(acc);
    free(buf_a);
    free(buf_b);
}

/* ------------------------------------------------------------------------------------------ */
/* Dense kernel = lattice + normalisation (Appendix A.5), pairwise potential = kernel + compat */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    orc_lattice *lat;
    float *norm; /* N */
    int ntype, ktype;
    int compat_kind;
    float *compat; /* 1 (Potts), L (diagonal) or L*L row-major symmetrised (matrix) */
} orc_pairwise;

typedef struct orc_crf {
    int N, L;
    float *unary; /* N x L pixel-major */
    int n_pair, cap_pair;
    orc_pairwise *pair;
} orc_crf;

/* `dcrf.DenseCRF2D(w, h, nlabels)` / `DenseCRF(nvar, nlabels)` (/root/reference/03c_hsn/utilities.py:427) */
orc_crf *orc_crf_create(int N, int L) {
    orc_crf *c = (orc_crf *)calloc(1, sizeof(orc_crf));
    c->N = N;
    c->L = L;
    c->unary = (float *)calloc((size_t)(N > 0 ? N : 1) * (L > 0 ? L : 1), sizeof(float));
    return c;
}

void orc_crf_free(orc_crf *c) {
    if (!c) return;
    for (int k = 0; k < c->n_pair; k++) {
        orc_lattice_free(c->pair[k].lat);
        free(c->pair[k].norm);
        free(c->pair[k].compat);
    }
    free(c->pair);
    free(c->unary);
    free(c);
}

/* U: row-major (L, N) as handed over by Python (03c_hsn/utilities.py:431-432) */
void orc_crf_set_unary(orc_crf *c, const float *U) {
    for (int l = 0; l < c->L; l++)
        for (int p = 0; p < c->N; p++) c->unary[(size_t)p * c->L + l] = U[(size_t)l * c->N + p];
}

/* feature_nd: N x d pixel-major.  compat per compat_kind.  Returns the pairwise index. */
int orc_crf_add_pairwise_nd(orc_crf *c, const float *feature_nd, int d, int compat_kind,
                            const float *compat, int ktype, int ntype) {
    if (c->n_pair == c->cap_pair) {
        c->cap_pair = c->cap_pair ? 2 * c->cap_pair : 4;
        c->pair = (orc_pairwise *)realloc(c->pair, sizeof(orc_pairwise) * c->cap_pair);
    }
    orc_pairwise *pw = &c->pair[c->n_pair];
    const int N = c->N, L = c->L;
    pw->ktype = ktype; /* CONST/DIAG/FULL all act as identity feature scaling at default params */
    pw->ntype = ntype;
    pw->lat = orc_lattice_create(feature_nd, N, d);
    /* A.5: norm = filter(ones) */
    pw->norm = (float *)malloc(sizeof(float) * (N > 0 ? N : 1));
    for (int i = 0; i < N; i++) pw->norm[i] = 1.0f;
    orc_lattice_compute(pw->lat, pw->norm, pw->norm, 1, 0);
    if (ntype == ORC_NO_NORMALIZATION) {
        float mean_norm = 0;
        for (int i = 0; i < N; i++) mean_norm += pw->norm[i];
        mean_norm = (float)N / mean_norm;
        for (int i = 0; i < N; i++) pw->norm[i] = mean_norm;
    } else if (ntype == ORC_NORMALIZE_SYMMETRIC) {
        for (int i = 0; i < N; i++) pw->norm[i] = (float)(1.0 / sqrt((double)pw->norm[i] + 1e-20));
    } else {
        for (int i = 0; i < N; i++) pw->norm[i] = (float)(1.0 / ((double)pw->norm[i] + 1e-20));
    }
    pw->compat_kind = compat_kind;
    if (compat_kind == ORC_COMPAT_POTTS) {
        pw->compat = (float *)malloc(sizeof(float));
        pw->compat[0] = compat[0];
    } else if (compat_kind == ORC_COMPAT_DIAGONAL) {
        pw->compat = (float *)malloc(sizeof(float) * L);
        memcpy(pw->compat, compat, sizeof(float) * L);
    } else {
        pw->compat = (float *)malloc(sizeof(float) * L * L);
        for (int a = 0; a < L; a++)
            for (int b = 0; b < L; b++)
                pw->compat[a * L + b] = 0.5f * (compat[a * L + b] + compat[b * L + a]);
    }
    return c->n_pair++;
}

/* feature_dN: row-major (d, N) as handed over by Python addPairwiseEnergy */
int orc_crf_add_pairwise(orc_crf *c, const float *feature_dN, int d, int compat_kind,
                         const float *compat, int ktype, int ntype) {
    const int N = c->N;
    float *f = (float *)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1) * (d > 0 ? d : 1));
    for (int j = 0; j < d; j++)
        for (int p = 0; p < N; p++) f[(size_t)p * d + j] = feature_dN[(size_t)j * N + p];
    int r = orc_crf_add_pairwise_nd(c, f, d, compat_kind, compat, ktype, ntype);
    free(f);
    return r;
}

/* `d.addPairwiseGaussian(sxy=..., compat=...)` (/root/reference/03c_hsn/utilities.py:435).
 * A.2 features: float32 true division of an integer by a float32 parameter */
int orc_crf_add_gaussian_2d(orc_crf *c, int W, int H, float sx, float sy, int compat_kind,
                            const float *compat, int ktype, int ntype) {
    float *f = (float *)malloc(sizeof(float) * (size_t)(W * H > 0 ? W * H : 1) * 2);
    for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++) {
            f[((size_t)j * W + i) * 2 + 0] = (float)i / sx;
            f[((size_t)j * W + i) * 2 + 1] = (float)j / sy;
        }
    int r = orc_crf_add_pairwise_nd(c, f, 2, compat_kind, compat, ktype, ntype);
    free(f);
    return r;
}

/* `d.addPairwiseBilateral(sxy=..., srgb=..., rgbim=..., compat=...)` (/root/reference/03c_hsn/utilities.py:439-440) */
int orc_crf_add_bilateral_2d(orc_crf *c, int W, int H, float sx, float sy, float sr, float sg,
                             float sb, const uint8_t *im, int compat_kind, const float *compat,
                             int ktype, int ntype) {
    float *f = (float *)malloc(sizeof(float) * (size_t)(W * H > 0 ? W * H : 1) * 5);
    for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++) {
            size_t p = (size_t)j * W + i;
            f[p * 5 + 0] = (float)i / sx;
            f[p * 5 + 1] = (float)j / sy;
            f[p * 5 + 2] = (float)im[p * 3 + 0] / sr;
            f[p * 5 + 3] = (float)im[p * 3 + 1] / sg;
            f[p * 5 + 4] = (float)im[p * 3 + 2] / sb;
        }
    int r = orc_crf_add_pairwise_nd(c, f, 5, compat_kind, compat, ktype, ntype);
    free(f);
    return r;
}

int orc_crf_num_pairwise(const orc_crf *c) { return c->n_pair; }
const orc_lattice *orc_crf_lattice(const orc_crf *c, int k) { return c->pair[k].lat; }
void orc_crf_norm(const orc_crf *c, int k, float *out) {
    memcpy(out, c->pair[k].norm, sizeof(float) * c->N);
}

/* DenseKernel::filter + LabelCompatibility::apply [EXT] (A.5 + A.6), once per pairwise term and
 * iteration of `d.inference(n)` (/root/reference/03c_hsn/utilities.py:442):
 * out = compat( norm (.) K( norm (.) Q ) ), pixel-major N x L */
static void orc_pairwise_apply(const orc_crf *c, const orc_pairwise *pw, float *out, const float *Q,
                               int transpose) {
    const int N = c->N, L = c->L;
    const int nt = pw->ntype;
    const int pre = (nt == ORC_NORMALIZE_SYMMETRIC) || (nt == ORC_NORMALIZE_BEFORE && !transpose) ||
                    (nt == ORC_NORMALIZE_AFTER && transpose);
    /* NO_NORMALIZATION: upstream computes the scalar N / sum(norm) into norm_ but its filter()
       applies norm_ on neither side for that mode; restated as "no scaling at all". */
    const int post = (nt == ORC_NORMALIZE_SYMMETRIC) || (nt == ORC_NORMALIZE_BEFORE && transpose) ||
                     (nt == ORC_NORMALIZE_AFTER && !transpose);
    for (int p = 0; p < N; p++)
        for (int l = 0; l < L; l++)
            out[(size_t)p * L + l] = pre ? Q[(size_t)p * L + l] * pw->norm[p] : Q[(size_t)p * L + l];
    orc_lattice_compute(pw->lat, out, out, L, transpose);
    if (post)
        for (int p = 0; p < N; p++)
            for (int l = 0; l < L; l++) out[(size_t)p * L + l] *= pw->norm[p];
    if (pw->compat_kind == ORC_COMPAT_POTTS) {
        const float w = -pw->compat[0];
        for (size_t i = 0; i < (size_t)N * L; i++) out[i] = w * out[i];
    } else if (pw->compat_kind == ORC_COMPAT_DIAGONAL) {
        for (int p = 0; p < N; p++)
            for (int l = 0; l < L; l++) out[(size_t)p * L + l] *= pw->compat[l];
    } else {
        float *tmp = (float *)malloc(sizeof(float) * L);
        for (int p = 0; p < N; p++) {
            float *o = out + (size_t)p * L;
            for (int a = 0; a < L; a++) {
                float s = 0;
                for (int b = 0; b < L; b++) s += pw->compat[a * L + b] * o[b];
                tmp[a] = s;
            }
            memcpy(o, tmp, sizeof(float) * L);
        }
        free(tmp);
    }
}

/* A.7: per-pixel max-subtracted softmax */
static void orc_exp_and_normalize(float *out, const float *in, int N, int L) {
    for (int p = 0; p < N; p++) {
        const float *b = in + (size_t)p * L;
        float *o = out + (size_t)p * L;
        float mx = b[0];
        for (int l = 1; l < L; l++) mx = b[l] > mx ? b[l] : mx;
        float s = 0;
        for (int l = 0; l < L; l++) {
            o[l] = expf(b[l] - mx);
            s += o[l];
        }
        for (int l = 0; l < L; l++) o[l] = o[l] / s;
    }
}

/* Q, tmp1, tmp2: N x L pixel-major work buffers owned by the caller */
void orc_crf_start_inference_pm(const orc_crf *c, float *Q) {
    const size_t n = (size_t)c->N * c->L;
    float *neg = (float *)malloc(sizeof(float) * (n ? n : 1));
    for (size_t i = 0; i < n; i++) neg[i] = -c->unary[i];
    orc_exp_and_normalize(Q, neg, c->N, c->L);
    free(neg);
}

void orc_crf_step_inference_pm(const orc_crf *c, float *Q, float *tmp1, float *tmp2) {
    const size_t n = (size_t)c->N * c->L;
    for (size_t i = 0; i < n; i++) tmp1[i] = -c->unary[i];
    for (int k = 0; k < c->n_pair; k++) {
        orc_pairwise_apply(c, &c->pair[k], tmp2, Q, 0);
        for (size_t i = 0; i < n; i++) tmp1[i] -= tmp2[i];
    }
    orc_exp_and_normalize(Q, tmp1, c->N, c->L);
}

static void orc_pm_to_ln(const orc_crf *c, const float *pm, float *ln) {
    for (int p = 0; p < c->N; p++)
        for (int l = 0; l < c->L; l++) ln[(size_t)l * c->N + p] = pm[(size_t)p * c->L + l];
}
static void orc_ln_to_pm(const orc_crf *c, const float *ln, float *pm) {
    for (int p = 0; p < c->N; p++)
        for (int l = 0; l < c->L; l++) pm[(size_t)p * c->L + l] = ln[(size_t)l * c->N + p];
}

/* `Q = d.inference(n_infer)` (/root/reference/03c_hsn/utilities.py:442), DenseCRF::inference [EXT] A.7.
 * Q_out: row-major (L, N) -- what np.array(Q) yields (03c_hsn/utilities.py:443) */
void orc_crf_inference(const orc_crf *c, int n_iter, float *Q_out) {
    const size_t n = (size_t)c->N * c->L;
    float *Q = (float *)malloc(sizeof(float) * (n ? n : 1));
    float *t1 = (float *)malloc(sizeof(float) * (n ? n : 1));
    float *t2 = (float *)malloc(sizeof(float) * (n ? n : 1));
    orc_crf_start_inference_pm(c, Q);
    for (int it = 0; it < n_iter; it++) orc_crf_step_inference_pm(c, Q, t1, t2);
    orc_pm_to_ln(c, Q, Q_out);
    free(Q); free(t1); free(t2);
}

/* startInference / stepInference on row-major (L,N) buffers (pydensecrf-style stepping) */
void orc_crf_start_inference(const orc_crf *c, float *Q_LN) {
    const size_t n = (size_t)c->N * c->L;
    float *Q = (float *)malloc(sizeof(float) * (n ? n : 1));
    orc_crf_start_inference_pm(c, Q);
    orc_pm_to_ln(c, Q, Q_LN);
    free(Q);
}
void orc_crf_step_inference(const orc_crf *c, float *Q_LN) {
    const size_t n = (size_t)c->N * c->L;
    float *Q = (float *)malloc(sizeof(float) * (n ? n : 1));
    float *t1 = (float *)malloc(sizeof(float) * (n ? n : 1));
    float *t2 = (float *)malloc(sizeof(float) * (n ? n : 1));
    orc_ln_to_pm(c, Q_LN, Q);
    orc_crf_step_inference_pm(c, Q, t1, t2);
    orc_pm_to_ln(c, Q, Q_LN);
    free(Q); free(t1); free(t2);
}

/* KL(Q || P) up to the log-partition constant: entropy + unary + pairwise terms, double accumulate */
double orc_crf_kl_divergence(const orc_crf *c, const float *Q_LN) {
    const size_t n = (size_t)c->N * c->L;
    float *Q = (float *)malloc(sizeof(float) * (n ? n : 1));
    float *tmp = (float *)malloc(sizeof(float) * (n ? n : 1));
    orc_ln_to_pm(c, Q_LN, Q);
    double kl = 0;
    for (size_t i = 0; i < n; i++) {
        float q = Q[i] > 1e-20f ? Q[i] : 1e-20f;
        kl += (double)Q[i] * log((double)q);
    }
    for (size_t i = 0; i < n; i++) kl += (double)c->unary[i] * (double)Q[i];
    for (int k = 0; k < c->n_pair; k++) {
        orc_pairwise_apply(c, &c->pair[k], tmp, Q, 0);
        double s = 0;
        for (size_t i = 0; i < n; i++) s += (double)(Q[i] * tmp[i]);
        kl += s;
    }
    free(Q); free(tmp);
    return kl;
}

/* ------------------------------------------------------------------------------------------ */
/* Brute-force O(N^2) Gaussian filter used only to sanity-check the lattice approximation:     */
/* out_i = sum_j exp(-0.5 |f_i - f_j|^2) in_j.  feature: N x d pixel-major, in/out: N x vs.     */
/* ------------------------------------------------------------------------------------------ */
void orc_bruteforce_gaussian(const float *feature, int N, int d, const float *in, float *out,
                             int vs) {
    for (int i = 0; i < N; i++) {
        double *acc = (double *)calloc((size_t)vs, sizeof(double));
        for (int j = 0; j < N; j++) {
            double d2 = 0;
            for (int k = 0; k < d; k++) {
                double t = (double)feature[(size_t)i * d + k] - (double)feature[(size_t)j * d + k];
                d2 += t * t;
            }
            double w = exp(-0.5 * d2);
            for (int k = 0; k < vs; k++) acc[k] += w * (double)in[(size_t)j * vs + k];
        }
        for (int k = 0; k < vs; k++) out[(size_t)i * vs + k] = (float)acc[k];
        free(acc);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Confusion matrix, chainercv convention (Appendix A.8; 03b_irn/step/eval_sem_seg.py:41):      */
/* conf[(C+1) x C] int64, row = GT class, col = prediction, gt < 0 or gt >= C -> row C          */
/* (ignored; the caller drops that row).  pred outside [0,C) is counted nowhere and reported.   */
/* ------------------------------------------------------------------------------------------ */
int64_t orc_confusion_accumulate(const int32_t *gt, const int32_t *pred, int64_t n, int C,
                                 int64_t *conf) {
    int64_t bad = 0;
    for (int64_t i = 0; i < n; i++) {
        int g = gt[i], p = pred[i];
        if (p < 0 || p >= C) { bad++; continue; }
        int row = (g >= 0 && g < C) ? g : C;
        conf[(size_t)row * C + p]++;
    }
    return bad;
}
