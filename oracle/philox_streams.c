/*
 * philox_streams.c -- CPU restatement of scorus_b200's DEVICE random-stream
 * generator ("stream spec v1").
 *
 * TEST INFRASTRUCTURE ONLY.  This is NOT reference behaviour: the reference
 * consumes a caller-supplied rand::Rng sequentially, which a
 * GPU cannot reproduce in parallel.  The product draws its partner / z / flag
 * / accept streams from counter-based Philox4x32-10 keyed by (seed, step,
 * walker).  This file regenerates exactly those streams on the CPU so that a
 * Philox-mode GPU step can be replayed through orc_sample_pt_fixed /
 * orc_swap_walkers_fixed and compared bit for bit.  It is written
 * independently of scorus_b200/csrc/rng.cuh; the two must agree.
 */
#include "scorus_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

enum {
    ST_ZU = 1,        /* idx = walker, sub = 0: (r0,r1) -> z uniform, (r2,r3) -> accept uniform */
    ST_FLAGS = 2,     /* idx = walker, sub = attempt*nblk + blk */
    ST_PERM = 3,      /* idx = rung, sub = 0,1: keys of the pairing bijection (npb > 1024) */
    ST_SWAP_PERM = 4, /* idx = hot rung, sub = 0,1: keys of the jvec bijection (npb > 1024) */
    ST_SWAP_U = 5,    /* idx = cold slot j2, sub = hot rung */
    ST_PERM_KEY = 6,  /* idx = walker: 64-bit sort key of the exact shuffle (npb <= 1024) */
    ST_SWAP_KEY = 7,  /* idx = hot walker slot: 64-bit sort key of the exact jvec (npb <= 1024) */
    ST_INIT = 8       /* idx = walker, sub = d/2, step = 0: box-uniform start of coordinates d, d+1 */
};

#define EXACT_SHUFFLE_MAX 1024u
/* Prob(p) flags: after this many all-false draws one coordinate is set, uniformly placed (tile_common.cuh) */
#define MAX_FLAG_ATTEMPTS 64u

static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    philox4x32_10(ctr, key, out);
}

static void draw(uint64_t seed, uint64_t step, uint32_t stream, uint32_t idx, uint32_t sub,
                 uint32_t out[4])
{
    uint32_t ctr[4] = {idx, sub, (uint32_t)step,
                       ((uint32_t)(step >> 32) & 0x00FFFFFFu) | (stream << 24)};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    philox4x32_10(ctr, key, out);
}

static double u53(uint32_t lo, uint32_t hi)
{
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (double)(v >> 11) * 0x1.0p-53;
}

/* keyed bijection on [0, n): 4 add / multiply / xorshift rounds on b bits, cycle-walked */
static uint32_t bij(uint32_t x, uint32_t n, const uint32_t k[8])
{
    static const uint32_t C[4] = {0x9E3779B1u, 0x85EBCA6Bu, 0xC2B2AE35u, 0x27D4EB2Fu};
    uint32_t b = 1;
    while (b < 32 && (1ull << b) < n) ++b;
    uint32_t mask = b >= 32 ? 0xFFFFFFFFu : ((1u << b) - 1u);
    uint32_t s1 = (b + 1) / 2, s2 = b / 3;
    if (s1 < 1) s1 = 1;
    if (s2 < 1) s2 = 1;
    do {
        for (int r = 0; r < 4; ++r) {
            x = (x + k[2 * r]) & mask;
            x = (x * (k[2 * r + 1] | 1u)) & mask;
            x ^= x >> s1;
            x = (x * C[r]) & mask;
            x ^= x >> s2;
        }
    } while (x >= n);
    return x;
}

typedef struct { uint64_t key; uint32_t id; } keyed;
static int cmp_keyed(const void *a, const void *b)
{
    const keyed *x = (const keyed *)a, *y = (const keyed *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->id < y->id ? -1 : (x->id > y->id);
}

/* order[j] = walker at matching position j; positions 2k and 2k+1 are partners.  Every rung is cut into
 * `nblocks` contiguous blocks that are matched separately (nblocks = 1: the reference's whole-rung matching,
 * ensemble_sample.rs:135-148; nblocks > 1: the shard-local steps of the k-schedule).
 * Block c of rung r draws its bijection keys at (idx = r, sub = 2c, 2c + 1); blocks of up to 1024 walkers are
 * shuffled exactly by ranking their walkers' 64-bit keys. */
static void pairing_order_blocked(uint64_t seed, uint64_t step, size_t n, size_t nbeta, size_t nblocks, uint32_t *order)
{
    size_t npb = n / nbeta, bs = npb / nblocks;
    if (bs <= EXACT_SHUFFLE_MAX) {
        keyed *kv = (keyed *)malloc(bs * sizeof(keyed));
        for (size_t blk = 0; blk < nbeta * nblocks; ++blk) {
            for (size_t k = 0; k < bs; ++k) {
                uint32_t o[4];
                uint32_t w = (uint32_t)(blk * bs + k);
                draw(seed, step, ST_PERM_KEY, w, 0, o);
                kv[k].key = ((uint64_t)o[1] << 32) | o[0];
                kv[k].id = w;
            }
            qsort(kv, bs, sizeof(keyed), cmp_keyed);
            for (size_t k = 0; k < bs; ++k) order[blk * bs + k] = kv[k].id;
        }
        free(kv);
    } else {
        for (size_t r = 0; r < nbeta; ++r)
            for (size_t c = 0; c < nblocks; ++c) {
                uint32_t k[8];
                draw(seed, step, ST_PERM, (uint32_t)r, (uint32_t)(2 * c), k);
                draw(seed, step, ST_PERM, (uint32_t)r, (uint32_t)(2 * c + 1), k + 4);
                const size_t base = r * npb + c * bs;
                for (size_t j = 0; j < bs; ++j) order[base + j] = (uint32_t)base + bij((uint32_t)j, (uint32_t)bs, k);
            }
    }
}

static void pairing_order(uint64_t seed, uint64_t step, size_t n, size_t nbeta, uint32_t *order)
{
    pairing_order_blocked(seed, step, n, nbeta, 1, order);
}

/* Prob(p) flags of walker w at `step` into f[dim] (the device's gen_flags_prob, tile_common.cuh): bpf bits per
 * coordinate, flag = value < thr, redrawn until one is set; bounded by MAX_FLAG_ATTEMPTS */
/* 32 bits per coordinate (a p that is not a multiple of 2^-16): no redraw loop (gen_flags_first, tile_common.cuh).  The
 * position J of the first set flag given that one is set is truncated-geometric; it is read off 32-bit thresholds
 * cdf[j] = (1 - (1-p)^(j+1)) / (1 - (1-p)^D) * 2^32 with the random word of coordinate 0; coordinates before J are clear,
 * J is set, the ones after J are iid Bernoulli from their own words. */
static void prob_flags_first(uint64_t seed, uint64_t step, uint32_t w, size_t dim, double prob, uint64_t thr, uint8_t *f)
{
    const double q = fmin(prob, 1.0), lq = log1p(-q), den = expm1((double)dim * lq);
    uint32_t o[4];
    draw(seed, step, ST_FLAGS, w, 0, o);
    size_t J = 0;
    while (J + 1 < dim) {
        const double t = expm1((double)(J + 1) * lq) / den * 4294967296.0;
        const long long v = llrint(t);
        const uint32_t c = v < 0 ? 0u : (v > 0xFFFFFFFFll ? 0xFFFFFFFFu : (uint32_t)v);
        if (o[0] >= c) ++J;
        else break;
    }
    const uint32_t thr32 = thr > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)thr;
    for (size_t d = 0; d < dim; ++d) {
        if (d % 4 == 0 && d) draw(seed, step, ST_FLAGS, w, (uint32_t)(d / 4), o);
        f[d] = d < J ? 0 : (d == J ? 1 : (uint8_t)(o[d % 4] < thr32));
    }
}

static void prob_flags(uint64_t seed, uint64_t step, uint32_t w, size_t dim, double prob, uint32_t bpf, uint64_t thr,
                       uint8_t *f)
{
    if (bpf == 32) { prob_flags_first(seed, step, w, dim, prob, thr, f); return; }
    const uint32_t fpc = 128 / bpf;
    const uint32_t nblk = (uint32_t)((dim + fpc - 1) / fpc);
    const uint64_t vmask = bpf == 32 ? 0xFFFFFFFFull : ((1ull << bpf) - 1);
    uint32_t o[4];
    for (uint32_t attempt = 0; attempt < MAX_FLAG_ATTEMPTS; ++attempt) {
        int any = 0;
        for (uint32_t blk = 0; blk < nblk; ++blk) {
            draw(seed, step, ST_FLAGS, w, attempt * nblk + blk, o);
            for (uint32_t q = 0; q < fpc; ++q) {
                const size_t d = (size_t)blk * fpc + q;
                if (d >= dim) break;
                const uint32_t bit = q * bpf;
                const uint64_t v = (o[bit / 32] >> (bit % 32)) & vmask;
                f[d] = v < thr;
                any |= f[d];
            }
        }
        if (any) return;
    }
    draw(seed, step, ST_FLAGS, w, MAX_FLAG_ATTEMPTS * nblk, o);
    f[(size_t)(((uint64_t)o[0] * (uint64_t)dim) >> 32)] = 1;
}

/* threshold on bpf random bits: prob * 2^bpf (rounded to nearest for 32 bits), at least 1 */
static uint64_t flag_threshold(double prob, uint32_t bpf)
{
    const double q = fmin(prob, 1.0);
    uint64_t thr = bpf == 32 ? (uint64_t)llrint(q * 4294967296.0) : (uint64_t)(q * (double)(1u << bpf));
    return thr ? thr : 1;
}

/* bits per Bernoulli flag: the smallest of 1,2,4,8,16 for which prob*2^bits is an integer, else 32 */
static uint32_t flag_bits(double prob)
{
    static const uint32_t cand[5] = {1, 2, 4, 8, 16};
    for (int i = 0; i < 5; ++i) {
        double t = prob * (double)(1u << cand[i]);
        if (t == floor(t)) return cand[i];
    }
    return 32;
}

/* The device streams of one sample_pt step.  walkers == NULL: all n walkers (m ignored), outputs indexed by walker.
 * walkers != NULL: only the m listed walkers (outputs indexed 0..m-1; partner[] still holds global ids) -- streams
 * are keyed by walker, so a sample of a 2^22-walker step is replayed at the cost of the sample (plus one pass
 * over the pairing bijection). */
int orc_philox_sample_streams_ex(uint64_t seed, uint64_t step, size_t n, size_t dim, size_t nbeta, double a,
                                 int ufs_kind, double prob, uint64_t onehot_counter, size_t nblocks,
                                 const uint32_t *walkers, size_t m, uint32_t *partner, double *z, uint8_t *flags,
                                 double *u)
{
    if (nbeta == 0 || n == 0 || n % nbeta || nblocks == 0) return ORC_ERR_INVALID_ARG;
    const size_t npb = n / nbeta;
    if (npb % nblocks || (npb / nblocks) % 2) return ORC_ERR_INVALID_ARG;
    if (!walkers) m = n;
    uint32_t *order = (uint32_t *)malloc(n * sizeof(uint32_t));
    pairing_order_blocked(seed, step, n, nbeta, nblocks, order);
    if (!walkers) {
        for (size_t j = 0; j < n; j += 2) {
            partner[order[j]] = order[j + 1];
            partner[order[j + 1]] = order[j];
        }
    } else {
        uint32_t *mate = (uint32_t *)malloc(n * sizeof(uint32_t));
        for (size_t j = 0; j < n; j += 2) {
            mate[order[j]] = order[j + 1];
            mate[order[j + 1]] = order[j];
        }
        for (size_t i = 0; i < m; ++i) {
            if (walkers[i] >= n) { free(mate); free(order); return ORC_ERR_INVALID_ARG; }
            partner[i] = mate[walkers[i]];
        }
        free(mate);
    }
    free(order);

    double sqrt_a = sqrt(a);
    double inv_sqrt_a = 1.0 / sqrt_a;
    double range = 2.0 * (sqrt_a - inv_sqrt_a);
    for (size_t i = 0; i < m; ++i) {
        uint32_t o[4];
        draw(seed, step, ST_ZU, walkers ? walkers[i] : (uint32_t)i, 0, o);
        double p = u53(o[0], o[1]) * range;
        double y = inv_sqrt_a + p / 2.0;
        z[i] = y * y;
        u[i] = u53(o[2], o[3]);
    }

    if (ufs_kind == ORC_UFS_ALL) {
        memset(flags, 1, m * dim);
    } else if (ufs_kind == ORC_UFS_ONE_HOT_CYCLE) {
        for (size_t i = 0; i < m; ++i) {
            size_t hot = (size_t)((onehot_counter + 1 + (walkers ? walkers[i] : i)) % dim);
            for (size_t k = 0; k < dim; ++k) flags[i * dim + k] = (k == hot);
        }
    } else if (ufs_kind == ORC_UFS_PROB) {
        if (!(prob > 0.0)) return ORC_ERR_INVALID_ARG;
        const uint32_t bpf = flag_bits(fmin(prob, 1.0));
        const uint64_t thr = flag_threshold(prob, bpf);
        for (size_t i = 0; i < m; ++i)
            prob_flags(seed, step, walkers ? walkers[i] : (uint32_t)i, dim, prob, bpf, thr, flags + i * dim);
    } else {
        return ORC_ERR_INVALID_ARG;
    }
    return ORC_OK;
}

/* The same streams by RANGES, so that a threaded caller (the flat CPU baseline, ref_structure_bench.c) can generate
 * them on all cores: partners of the matching positions [pos_lo, pos_hi) (even bounds; large rungs only -- the keyed
 * bijection is a pure function of the position), and z / flags / u of the walkers [w_lo, w_hi). */
int orc_philox_partner_range(uint64_t seed, uint64_t step, size_t n, size_t nbeta, size_t pos_lo, size_t pos_hi,
                             uint32_t *partner)
{
    const size_t npb = n / nbeta;
    if (npb <= EXACT_SHUFFLE_MAX || (pos_lo & 1) || (pos_hi & 1)) return ORC_ERR_INVALID_ARG;
    size_t r_cur = (size_t)-1;
    uint32_t k[8];
    for (size_t j = pos_lo; j < pos_hi; j += 2) {
        const size_t r = j / npb;
        if (r != r_cur) {
            draw(seed, step, ST_PERM, (uint32_t)r, 0, k);
            draw(seed, step, ST_PERM, (uint32_t)r, 1, k + 4);
            r_cur = r;
        }
        const uint32_t wa = (uint32_t)(r * npb) + bij((uint32_t)(j - r * npb), (uint32_t)npb, k);
        const uint32_t wb = (uint32_t)(r * npb) + bij((uint32_t)(j + 1 - r * npb), (uint32_t)npb, k);
        partner[wa] = wb;
        partner[wb] = wa;
    }
    return ORC_OK;
}

int orc_philox_walker_range(uint64_t seed, uint64_t step, size_t dim, double a, int ufs_kind, double prob,
                            uint64_t onehot_counter, size_t w_lo, size_t w_hi, double *z, uint8_t *flags, double *u)
{
    const double sqrt_a = sqrt(a), inv_sqrt_a = 1.0 / sqrt_a, range = 2.0 * (sqrt_a - inv_sqrt_a);
    const uint32_t bpf = ufs_kind == ORC_UFS_PROB ? flag_bits(fmin(prob, 1.0)) : 1;
    const uint64_t thr = ufs_kind == ORC_UFS_PROB ? flag_threshold(prob, bpf) : 0;
    if (ufs_kind == ORC_UFS_PROB && !(prob > 0.0)) return ORC_ERR_INVALID_ARG;
    for (size_t i = w_lo; i < w_hi; ++i) {
        uint32_t o[4];
        draw(seed, step, ST_ZU, (uint32_t)i, 0, o);
        const double p = u53(o[0], o[1]) * range, y = inv_sqrt_a + p / 2.0;
        z[i] = y * y;
        u[i] = u53(o[2], o[3]);
        if (ufs_kind == ORC_UFS_ALL) {
            memset(flags + i * dim, 1, dim);
        } else if (ufs_kind == ORC_UFS_ONE_HOT_CYCLE) {
            const size_t hot = (size_t)((onehot_counter + 1 + i) % dim);
            for (size_t d = 0; d < dim; ++d) flags[i * dim + d] = (d == hot);
        } else if (ufs_kind == ORC_UFS_PROB) {
            prob_flags(seed, step, (uint32_t)i, dim, prob, bpf, thr, flags + i * dim);
        } else {
            return ORC_ERR_INVALID_ARG;
        }
    }
    return ORC_OK;
}

int orc_philox_sample_streams(uint64_t seed, uint64_t step, size_t n, size_t dim, size_t nbeta,
                              double a, int ufs_kind, double prob, uint64_t onehot_counter,
                              uint32_t *partner, double *z, uint8_t *flags, double *u)
{
    if (nbeta == 0 || n == 0 || n % nbeta || (n / nbeta) % 2) return ORC_ERR_INVALID_ARG;
    return orc_philox_sample_streams_ex(seed, step, n, dim, nbeta, a, ufs_kind, prob, onehot_counter, 1, NULL, 0,
                                        partner, z, flags, u);
}

int orc_philox_swap_streams(uint64_t seed, uint64_t swap_call, size_t n, size_t nbeta,
                            uint32_t *jvec, double *r)
{
    if (nbeta == 0 || n % nbeta) return ORC_ERR_INVALID_ARG;
    size_t npb = n / nbeta;
    keyed *kv = (keyed *)malloc((npb ? npb : 1) * sizeof(keyed));
    for (size_t s = 0; s + 1 < nbeta; ++s) {
        size_t i = nbeta - 1 - s;
        if (npb <= EXACT_SHUFFLE_MAX) {
            for (size_t k = 0; k < npb; ++k) {
                uint32_t o[4];
                draw(seed, swap_call, ST_SWAP_KEY, (uint32_t)(i * npb + k), 0, o);
                kv[k].key = ((uint64_t)o[1] << 32) | o[0];
                kv[k].id = (uint32_t)k;
            }
            qsort(kv, npb, sizeof(keyed), cmp_keyed);
            for (size_t k = 0; k < npb; ++k) jvec[s * npb + k] = kv[k].id;
        } else {
            uint32_t k[8];
            draw(seed, swap_call, ST_SWAP_PERM, (uint32_t)i, 0, k);
            draw(seed, swap_call, ST_SWAP_PERM, (uint32_t)i, 1, k + 4);
            /* the bijection maps the hot slot j1 to the cold slot j2, i.e. jvec[j2] = j1 */
            for (size_t j1 = 0; j1 < npb; ++j1) jvec[s * npb + bij((uint32_t)j1, (uint32_t)npb, k)] = (uint32_t)j1;
        }
        for (size_t j2 = 0; j2 < npb; ++j2) {
            uint32_t o[4];
            draw(seed, swap_call, ST_SWAP_U, (uint32_t)j2, (uint32_t)i, o);
            r[s * npb + j2] = u53(o[0], o[1]);
        }
    }
    free(kv);
    return ORC_OK;
}

/* the device's box-uniform init (scorus_init_ensemble): x[w][d] = lo[d] + u * (hi[d] - lo[d]) */
void orc_philox_init_ensemble(uint64_t seed, size_t n, size_t dim, uint64_t walker_offset, const double *lo,
                              const double *hi, double *ens)
{
    for (size_t w = 0; w < n; ++w)
        for (size_t c = 0; 2 * c < dim; ++c) {
            uint32_t o[4];
            draw(seed, 0, ST_INIT, (uint32_t)(w + walker_offset), (uint32_t)c, o);
            size_t d0 = 2 * c;
            ens[w * dim + d0] = lo[d0] + u53(o[0], o[1]) * (hi[d0] - lo[d0]);
            if (d0 + 1 < dim) ens[w * dim + d0 + 1] = lo[d0 + 1] + u53(o[2], o[3]) * (hi[d0 + 1] - lo[d0 + 1]);
        }
}

/* ---- t-walk half-pass streams (scorus_b200/csrc/twalk_kernel.cuh) ----
 * The matching is drawn at order_step (shared by both half-passes); everything else at `step`, keyed by the
 * MOVING walker: ST_ZU (r0,r1) -> kernel uniform scaled by fw[1], (r2,r3) -> accept uniform; ST_TW_B = 9
 * (r0,r1) -> x, (r2,r3) -> v of sim_b; ST_FLAGS as Prob(pphi); ST_TW_WALK = 10, sub = d/2 -> the sim_walk
 * uniforms of coordinates d, d+1.  Outputs are per pair k in matching order, as orc_twalk_half_fixed takes
 * them.  b goes through this host's pow(): a device run is replayed with the device's own b values
 * (scorus_twalk_draw_b), which agree with these to a few ulp. */
enum { ST_TW_B = 9, ST_TW_WALK = 10 };

int orc_philox_twalk_streams(uint64_t seed, uint64_t step, uint64_t order_step, size_t n, size_t dim, size_t nbeta,
                             const orc_twalk_params *tw, int half, uint32_t *order, uint8_t *kernel, double *b,
                             uint8_t *flags, double *walk_u, double *u)
{
    if (nbeta == 0 || n == 0 || n % nbeta || (n / nbeta) % 2) return ORC_ERR_INVALID_ARG;
    if (!(tw->pphi > 0.0)) return ORC_ERR_INVALID_ARG;
    pairing_order(seed, order_step, n, nbeta, order);
    const uint32_t bpf = flag_bits(fmin(tw->pphi, 1.0));
    const uint64_t thr = flag_threshold(tw->pphi, bpf);
    for (size_t k = 0; k < n / 2; ++k) {
        const uint32_t w = order[2 * k + half];
        uint32_t o[4];
        draw(seed, step, ST_ZU, w, 0, o);
        kernel[k] = u53(o[0], o[1]) * tw->fw[1] < tw->fw[0] ? 0 : 1;
        u[k] = u53(o[2], o[3]);
        b[k] = 1.0;
        if (kernel[k] == 1) {
            draw(seed, step, ST_TW_B, w, 0, o);
            b[k] = orc_twalk_b_from_uniforms(u53(o[0], o[1]), u53(o[2], o[3]), tw->at);
        }
        prob_flags(seed, step, w, dim, tw->pphi, bpf, thr, flags + k * dim);
        for (size_t d = 0; d < dim; d += 2) {
            draw(seed, step, ST_TW_WALK, w, (uint32_t)(d / 2), o);
            walk_u[k * dim + d] = u53(o[0], o[1]);
            if (d + 1 < dim) walk_u[k * dim + d + 1] = u53(o[2], o[3]);
        }
    }
    return ORC_OK;
}
