/*
 * kge_oracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C) of the reference's
 * knowledge-graph-embedding hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this file's library; the product path
 * (openkeonspark_b200/) never does.
 *
 * What is restated, and from where (paths under /root/reference):
 *   LCG streams ............... base/Random.h:9-34
 *   index build ............... base/Reader.h:103-177 (train), 259-291 (test/valid), 302-365 (types)
 *   corrupt_head/tail/rel ..... base/Corrupt.h:7-101
 *   _find ..................... base/Corrupt.h:104-115
 *   getBatch / sampling ....... base/Base.cpp:74-172
 *   testHead / testTail ....... base/Test.h:31-249
 *   getBestThreshold .......... base/Test.h:304-341
 *   test_triple_classification  base/Test.h:347-387
 *   model scores (predict_def)  TransE.py:11-15,53-58  TransH.py:12-20,72-82
 *                               TransR.py:16-23,77-87  TransD.py:23-31,86-98
 *
 * Pinning: the integer half is pinned against the reference itself — oracle/_ref/Base.so is
 * compiled from /root/reference/base/Base.cpp (oracle/Makefile) and tests/test_oracle_vs_ref.py
 * compares every function here with it bit-for-bit; tests/golden/ holds vectors generated from
 * that library.  The floating-point half (scores) restates TensorFlow 1.x graph semantics; TF is
 * not installable here and the reference ships no golden vectors, so for that half: PARITY
 * UNPINNED (checked against an fp64 autograd shadow only).
 *
 * Canonical fp32 evaluation order for scores (shared with the CUDA kernels so that scores, and
 * therefore ranks, are bit-identical): every reduction over the embedding dimension runs
 * sequentially d = 0..D-1, every multiply and add is individually rounded (no FMA contraction:
 * build with -ffp-contract=off), 1/sqrt is an IEEE sqrt followed by an IEEE divide.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;
typedef uint64_t u64;

typedef struct { i64 h, r, t; } trip;

typedef struct {
    i64 E, R;
    /* train side */
    i64 n_raw;          /* rows of train2id.txt, duplicates kept (trainTotal_) */
    i64 n;              /* deduplicated count (trainTotal) */
    i64 new_batch;      /* newBatchTotal (batch2id.txt), 0 if none */
    trip *raw;          /* file order */
    trip *by_h, *by_t, *by_ht;   /* sorted (h,r,t) / (t,r,h) / (h,t,r) */
    i64 *lef_h, *rig_h, *lef_t, *rig_t, *lef_ht, *rig_ht;
    float *tph, *hpt;   /* left_mean / right_mean */
    /* test side */
    i64 n_test, n_valid, n_all;
    trip *test, *valid, *all;    /* test/valid sorted (r,h,t); all sorted (h,r,t), dups kept */
    i64 *test_lef, *test_rig, *valid_lef, *valid_rig;
    /* type constraints: per relation sorted id lists */
    i64 *head_lef, *head_rig, *tail_lef, *tail_rig, *head_type, *tail_type;
    /* ontology: per entity sorted lists */
    i64 *sup_lef, *sup_rig, *sub_lef, *sub_rig, *sup_type, *sub_type;
    /* sampler state */
    i64 W;
    u64 *state;
    int bern;
} orc;

/* ---------------------------------------------------------------- comparators (Triple.h:18-32) */
static int cmp_hrt(const void *a, const void *b) {
    const trip *x = a, *y = b;
    if (x->h != y->h) return x->h < y->h ? -1 : 1;
    if (x->r != y->r) return x->r < y->r ? -1 : 1;
    if (x->t != y->t) return x->t < y->t ? -1 : 1;
    return 0;
}
static int cmp_trh(const void *a, const void *b) {
    const trip *x = a, *y = b;
    if (x->t != y->t) return x->t < y->t ? -1 : 1;
    if (x->r != y->r) return x->r < y->r ? -1 : 1;
    if (x->h != y->h) return x->h < y->h ? -1 : 1;
    return 0;
}
static int cmp_htr(const void *a, const void *b) {
    const trip *x = a, *y = b;
    if (x->h != y->h) return x->h < y->h ? -1 : 1;
    if (x->t != y->t) return x->t < y->t ? -1 : 1;
    if (x->r != y->r) return x->r < y->r ? -1 : 1;
    return 0;
}
static int cmp_rht(const void *a, const void *b) {
    const trip *x = a, *y = b;
    if (x->r != y->r) return x->r < y->r ? -1 : 1;
    if (x->h != y->h) return x->h < y->h ? -1 : 1;
    if (x->t != y->t) return x->t < y->t ? -1 : 1;
    return 0;
}
static int cmp_i64(const void *a, const void *b) {
    i64 x = *(const i64 *)a, y = *(const i64 *)b;
    return x < y ? -1 : x > y;
}

orc *orc_new(void) { return (orc *)calloc(1, sizeof(orc)); }

/* ---------------------------------------------------------------- Reader.h:81-177 */
/* hs/ts/rs: the rows of train2id.txt in file order. */
void orc_load_train(orc *o, i64 E, i64 R, const i64 *hs, const i64 *ts, const i64 *rs, i64 n_raw,
                    i64 new_batch) {
    o->E = E; o->R = R; o->n_raw = n_raw; o->new_batch = new_batch;
    o->raw = malloc(sizeof(trip) * (n_raw ? n_raw : 1));
    trip *tmp = malloc(sizeof(trip) * (n_raw ? n_raw : 1));
    for (i64 i = 0; i < n_raw; i++) {
        o->raw[i].h = hs[i]; o->raw[i].t = ts[i]; o->raw[i].r = rs[i];
        tmp[i] = o->raw[i];
    }
    qsort(tmp, n_raw, sizeof(trip), cmp_hrt);
    i64 *freq_rel = calloc(R ? R : 1, sizeof(i64));
    i64 n = 0;
    for (i64 i = 0; i < n_raw; i++)
        if (i == 0 || cmp_hrt(&tmp[i], &tmp[i - 1]) != 0) { tmp[n++] = tmp[i]; freq_rel[tmp[n - 1].r]++; }
    o->n = n;
    o->by_h = malloc(sizeof(trip) * (n ? n : 1));
    o->by_t = malloc(sizeof(trip) * (n ? n : 1));
    o->by_ht = malloc(sizeof(trip) * (n ? n : 1));
    memcpy(o->by_h, tmp, sizeof(trip) * n);
    memcpy(o->by_t, tmp, sizeof(trip) * n);
    memcpy(o->by_ht, tmp, sizeof(trip) * n);
    free(tmp);
    qsort(o->by_t, n, sizeof(trip), cmp_trh);
    qsort(o->by_ht, n, sizeof(trip), cmp_htr);

    i64 **lr[6] = {&o->lef_h, &o->rig_h, &o->lef_t, &o->rig_t, &o->lef_ht, &o->rig_ht};
    for (int k = 0; k < 6; k++) {
        *lr[k] = malloc(sizeof(i64) * (E ? E : 1));
        for (i64 e = 0; e < E; e++) (*lr[k])[e] = (k & 1) ? -1 : 0;   /* lef = 0, rig = -1 when absent */
    }
    for (i64 i = 0; i < n; i++) {
        if (i == 0 || o->by_h[i].h != o->by_h[i - 1].h) o->lef_h[o->by_h[i].h] = i;
        if (i == n - 1 || o->by_h[i].h != o->by_h[i + 1].h) o->rig_h[o->by_h[i].h] = i;
        if (i == 0 || o->by_t[i].t != o->by_t[i - 1].t) o->lef_t[o->by_t[i].t] = i;
        if (i == n - 1 || o->by_t[i].t != o->by_t[i + 1].t) o->rig_t[o->by_t[i].t] = i;
        if (i == 0 || o->by_ht[i].h != o->by_ht[i - 1].h) o->lef_ht[o->by_ht[i].h] = i;
        if (i == n - 1 || o->by_ht[i].h != o->by_ht[i + 1].h) o->rig_ht[o->by_ht[i].h] = i;
    }
    /* Reader.h:160-177: tph[r] = freq[r] / #distinct heads of r, hpt[r] = freq[r] / #distinct tails */
    o->tph = calloc(R ? R : 1, sizeof(float));
    o->hpt = calloc(R ? R : 1, sizeof(float));
    for (i64 i = 0; i < n; i++) {
        if (i == 0 || o->by_h[i].h != o->by_h[i - 1].h || o->by_h[i].r != o->by_h[i - 1].r)
            o->tph[o->by_h[i].r] += 1.0f;
        if (i == 0 || o->by_t[i].t != o->by_t[i - 1].t || o->by_t[i].r != o->by_t[i - 1].r)
            o->hpt[o->by_t[i].r] += 1.0f;
    }
    for (i64 r = 0; r < R; r++) {
        o->tph[r] = (float)freq_rel[r] / o->tph[r];
        o->hpt[r] = (float)freq_rel[r] / o->hpt[r];
    }
    free(freq_rel);
}

/* ---------------------------------------------------------------- Reader.h:225-291 */
static void rel_ranges(const trip *lst, i64 n, i64 R, i64 **lef, i64 **rig) {
    *lef = malloc(sizeof(i64) * (R ? R : 1));
    *rig = malloc(sizeof(i64) * (R ? R : 1));
    for (i64 r = 0; r < R; r++) (*lef)[r] = (*rig)[r] = -1;
    for (i64 i = 0; i < n; i++) {
        if (i == 0 || lst[i].r != lst[i - 1].r) (*lef)[lst[i].r] = i;
        if (i == n - 1 || lst[i].r != lst[i + 1].r) (*rig)[lst[i].r] = i;
    }
}

/* test/valid rows in file order; the train rows are taken from orc_load_train's raw copy. */
void orc_load_test(orc *o, const i64 *th, const i64 *tt, const i64 *tr, i64 n_test,
                   const i64 *vh, const i64 *vt, const i64 *vr, i64 n_valid) {
    o->n_test = n_test; o->n_valid = n_valid; o->n_all = n_test + o->n_raw + n_valid;
    o->test = malloc(sizeof(trip) * (n_test ? n_test : 1));
    o->valid = malloc(sizeof(trip) * (n_valid ? n_valid : 1));
    o->all = malloc(sizeof(trip) * (o->n_all ? o->n_all : 1));
    i64 k = 0;
    for (i64 i = 0; i < n_test; i++) { trip x = {th[i], tr[i], tt[i]}; o->test[i] = x; o->all[k++] = x; }
    for (i64 i = 0; i < o->n_raw; i++) o->all[k++] = o->raw[i];
    for (i64 i = 0; i < n_valid; i++) { trip x = {vh[i], vr[i], vt[i]}; o->valid[i] = x; o->all[k++] = x; }
    qsort(o->all, o->n_all, sizeof(trip), cmp_hrt);
    qsort(o->test, n_test, sizeof(trip), cmp_rht);
    qsort(o->valid, n_valid, sizeof(trip), cmp_rht);
    rel_ranges(o->test, n_test, o->R, &o->test_lef, &o->test_rig);
    rel_ranges(o->valid, n_valid, o->R, &o->valid_lef, &o->valid_rig);
}

/* ---------------------------------------------------------------- Reader.h:302-365, 376-449 */
/* CSR input: for relation rel[i], ids flat[off[i]..off[i+1]).  Lists are sorted here. */
static void csr_lists(i64 n_keys, const i64 *keys, const i64 *off, const i64 *flat, i64 n_slots,
                      i64 **lef, i64 **rig, i64 **out) {
    *lef = calloc(n_slots ? n_slots : 1, sizeof(i64));
    *rig = calloc(n_slots ? n_slots : 1, sizeof(i64));
    i64 tot = n_keys ? off[n_keys] : 0;
    *out = malloc(sizeof(i64) * (tot ? tot : 1));
    memcpy(*out, flat, sizeof(i64) * tot);
    for (i64 i = 0; i < n_keys; i++) {
        (*lef)[keys[i]] = off[i];
        (*rig)[keys[i]] = off[i + 1];
        qsort(*out + off[i], off[i + 1] - off[i], sizeof(i64), cmp_i64);
    }
}
void orc_load_types(orc *o, i64 n_rel, const i64 *rels, const i64 *hoff, const i64 *hflat,
                    const i64 *toff, const i64 *tflat) {
    csr_lists(n_rel, rels, hoff, hflat, o->R, &o->head_lef, &o->head_rig, &o->head_type);
    csr_lists(n_rel, rels, toff, tflat, o->R, &o->tail_lef, &o->tail_rig, &o->tail_type);
}
void orc_load_ontology(orc *o, i64 n_ent, const i64 *ents, const i64 *supoff, const i64 *supflat,
                       const i64 *suboff, const i64 *subflat) {
    csr_lists(n_ent, ents, supoff, supflat, o->E, &o->sup_lef, &o->sup_rig, &o->sup_type);
    csr_lists(n_ent, ents, suboff, subflat, o->E, &o->sub_lef, &o->sub_rig, &o->sub_type);
}

/* ---------------------------------------------------------------- Random.h:9-34 */
void orc_set_streams(orc *o, i64 W, const u64 *seeds, int bern) {
    free(o->state);
    o->W = W; o->bern = bern;
    o->state = malloc(sizeof(u64) * W);
    memcpy(o->state, seeds, sizeof(u64) * W);
}
/* Random.h:12 seeds each stream from libc rand(); exposed so that a test can draw the same
 * process-global sequence for the oracle and for the library under test. */
void orc_libc_seeds(u64 *out, i64 W) { for (i64 i = 0; i < W; i++) out[i] = (u64)rand(); }
void orc_get_streams(const orc *o, u64 *out) { memcpy(out, o->state, sizeof(u64) * o->W); }

static u64 lcg(orc *o, i64 id) {
    o->state[id] = o->state[id] * 25214903917ULL + 11ULL;
    return o->state[id];
}
static i64 draw_below(orc *o, i64 id, i64 x) { return (i64)(lcg(o, id) % (u64)x); }

/* ---------------------------------------------------------------- Corrupt.h:7-101 */
/* One routine serves all three corruptions: `lst` restricted to [lo,hi] (the rows of one primary
 * key) is sorted by a secondary key sec(), then by the value val().  Find the run [ll,rr] with
 * sec == key, draw tmp in [0, total - run), return the tmp-th id NOT among the run's values. */
#define SEC_R 0
#define SEC_T 1
#define VAL_T 0
#define VAL_H 1
#define VAL_R 2
static i64 fld(const trip *x, int which_sec, int which_val, int want_val) {
    if (!want_val) return which_sec == SEC_R ? x->r : x->t;
    return which_val == VAL_T ? x->t : (which_val == VAL_H ? x->h : x->r);
}
static i64 kth_absent(orc *o, i64 id, const trip *lst, i64 lo, i64 hi, i64 key, i64 total, int sec, int val) {
    i64 lef = lo - 1, rig = hi, mid;
    while (lef + 1 < rig) {
        mid = (lef + rig) >> 1;
        if (fld(&lst[mid], sec, val, 0) >= key) rig = mid; else lef = mid;
    }
    i64 ll = rig;
    lef = lo; rig = hi + 1;
    while (lef + 1 < rig) {
        mid = (lef + rig) >> 1;
        if (fld(&lst[mid], sec, val, 0) <= key) lef = mid; else rig = mid;
    }
    i64 rr = lef;
    i64 tmp = draw_below(o, id, total - (rr - ll + 1));
    if (tmp < fld(&lst[ll], sec, val, 1)) return tmp;
    if (tmp > fld(&lst[rr], sec, val, 1) - rr + ll - 1) return tmp + rr - ll + 1;
    lef = ll; rig = rr + 1;
    while (lef + 1 < rig) {
        mid = (lef + rig) >> 1;
        if (fld(&lst[mid], sec, val, 1) - mid + ll - 1 < tmp) lef = mid; else rig = mid;
    }
    return tmp + lef - ll + 1;
}
/* keeps (h,r), returns a new TAIL (the reference calls this corrupt_head) */
i64 orc_new_tail(orc *o, i64 id, i64 h, i64 r) {
    return kth_absent(o, id, o->by_h, o->lef_h[h], o->rig_h[h], r, o->E, SEC_R, VAL_T);
}
/* keeps (t,r), returns a new HEAD (reference: corrupt_tail) */
i64 orc_new_head(orc *o, i64 id, i64 t, i64 r) {
    return kth_absent(o, id, o->by_t, o->lef_t[t], o->rig_t[t], r, o->E, SEC_R, VAL_H);
}
/* keeps (h,t), returns a new RELATION (reference: corrupt_rel) */
i64 orc_new_rel(orc *o, i64 id, i64 h, i64 t) {
    return kth_absent(o, id, o->by_ht, o->lef_ht[h], o->rig_ht[h], t, o->R, SEC_T, VAL_R);
}

/* ---------------------------------------------------------------- Base.cpp:74-172 */
void orc_sampling(orc *o, i64 *bh, i64 *bt, i64 *br, float *by, i64 B, i64 k, i64 kr) {
    for (i64 id = 0; id < o->W; id++) {       /* the reference's threads touch disjoint state: order-free */
        i64 per = B / o->W + (B % o->W ? 1 : 0);
        i64 lef = id * per, rig = (id + 1) * per;
        if (rig > B) rig = B;
        float prob = 500;
        for (i64 b = lef; b < rig; b++) {
            i64 i = o->new_batch > 0
                ? (i64)(lcg(o, id) % (u64)o->new_batch) + (o->n_raw - o->new_batch)   /* Base.cpp:101-103 */
                : draw_below(o, id, o->n_raw);
            trip p = o->raw[i];
            bh[b] = p.h; bt[b] = p.t; br[b] = p.r; by[b] = 1;
            i64 at = b + B;
            for (i64 m = 0; m < k; m++, at += B) {
                if (o->bern) prob = 1000 * o->hpt[p.r] / (o->hpt[p.r] + o->tph[p.r]);
                if ((float)(lcg(o, id) % 1000) < prob) {
                    bh[at] = p.h; bt[at] = orc_new_tail(o, id, p.h, p.r); br[at] = p.r;
                } else {
                    bh[at] = orc_new_head(o, id, p.t, p.r); bt[at] = p.t; br[at] = p.r;
                }
                by[at] = -1;
            }
            for (i64 m = 0; m < kr; m++, at += B) {
                bh[at] = p.h; bt[at] = p.t; br[at] = orc_new_rel(o, id, p.h, p.t); by[at] = -1;
            }
        }
    }
}

/* ---------------------------------------------------------------- Corrupt.h:104-115 */
int orc_find(const orc *o, i64 h, i64 t, i64 r) {
    trip key = {h, r, t};
    i64 lef = 0, rig = o->n_all - 1;
    while (lef + 1 < rig) {
        i64 mid = (lef + rig) >> 1;
        if (cmp_hrt(&o->all[mid], &key) < 0) lef = mid; else rig = mid;
    }
    return cmp_hrt(&o->all[lef], &key) == 0 || cmp_hrt(&o->all[rig], &key) == 0;
}

/* ---------------------------------------------------------------- Corrupt.h:118-137 */
/* Type-constrained negative tail for triple classification; draws from libc rand(). */
i64 orc_tc_negative_tail(orc *o, i64 h, i64 r) {
    i64 ll = o->tail_lef[r], rr = o->tail_rig[r];
    for (int tries = 0; tries < 1000; tries++) {
        i64 t = o->tail_type[(rand() % (rr - ll)) + ll];
        if (!orc_find(o, h, t, r)) return t;
    }
    return orc_new_tail(o, 0, h, r);
}
/* Test.h:258-300.  which = 0 test list, 1 valid list. */
void orc_tc_batch(orc *o, int which, i64 *ph, i64 *pt, i64 *pr, i64 *nh, i64 *nt, i64 *nr) {
    const trip *lst = which ? o->valid : o->test;
    i64 n = which ? o->n_valid : o->n_test;
    for (i64 i = 0; i < n; i++) {
        ph[i] = nh[i] = lst[i].h; pr[i] = nr[i] = lst[i].r; pt[i] = lst[i].t;
    }
    for (i64 i = 0; i < n; i++) nt[i] = orc_tc_negative_tail(o, lst[i].h, lst[i].r);
}
void orc_get_list(const orc *o, int which, i64 *h, i64 *t, i64 *r) {
    const trip *lst = which ? o->valid : o->test;
    i64 n = which ? o->n_valid : o->n_test;
    for (i64 i = 0; i < n; i++) { h[i] = lst[i].h; t[i] = lst[i].t; r[i] = lst[i].r; }
}

/* ---------------------------------------------------------------- Test.h:31-249 */
/* side 0 = head replaced (testHead), 1 = tail replaced (testTail).  out[8] as the reference:
 * raw, filter, raw_constrain, filter_constrain better-than counts; then the 4 argmin ids mapped
 * to 0 ok / 1 generalisation / 2 specialisation / 3 misclassification. */
void orc_rank(const orc *o, int side, i64 index, const float *score, i64 *out) {
    trip q = o->test[index];
    i64 target = side ? q.t : q.h;
    float ref = score[target];
    i64 cnt[4] = {0, 0, 0, 0};
    i64 arg[4] = {target, target, target, target};
    float best[4] = {ref, ref, ref, ref};
    const i64 *types = side ? o->tail_type : o->head_type;
    i64 cur = side ? o->tail_lef[q.r] : o->head_lef[q.r];
    i64 end = side ? o->tail_rig[q.r] : o->head_rig[q.r];
    for (i64 j = 0; j < o->E; j++) {
        if (j == target) continue;
        float v = score[j];
        int better = v < ref;
        int known = 0;
        if (better) {
            cnt[0]++;
            if (v < best[0]) { best[0] = v; arg[0] = j; }
            known = side ? orc_find(o, q.h, j, q.r) : orc_find(o, j, q.t, q.r);
            if (!known) cnt[1]++;
            /* Test.h:69-74: on the head side the argmin update is NOT guarded by the filter
             * (missing braces), so filter-argmin tracks the raw argmin; the tail side guards it. */
            if ((side == 0 || !known) && v < best[1]) { best[1] = v; arg[1] = j; }
        }
        while (cur < end && types[cur] < j) cur++;
        if (cur < end && types[cur] == j && better) {
            cnt[2]++;
            if (v < best[2]) { best[2] = v; arg[2] = j; }
            if (!known) {
                cnt[3]++;
                if (v < best[3]) { best[3] = v; arg[3] = j; }
            }
        }
    }
    for (int k = 0; k < 4; k++) out[k] = cnt[k];
    /* Test.h:109-132 / 223-246: cursors are shared by the four lookups and never rewound. */
    i64 a = o->sup_lef ? o->sup_lef[target] : 0, ae = o->sup_rig ? o->sup_rig[target] : 0;
    i64 b = o->sub_lef ? o->sub_lef[target] : 0, be = o->sub_rig ? o->sub_rig[target] : 0;
    for (int k = 0; k < 4; k++) {
        i64 id = arg[k];
        if (id == target) { out[4 + k] = 0; continue; }
        while (a < ae && o->sup_type[a] < id) a++;
        if (a < ae && o->sup_type[a] == id) { out[4 + k] = 1; continue; }
        while (b < be && o->sub_type[b] < id) b++;
        if (b < be && o->sub_type[b] == id) { out[4 + k] = 2; continue; }
        out[4 + k] = 3;
    }
}

/* ---------------------------------------------------------------- Test.h:304-341 */
void orc_best_threshold(const orc *o, float *thresh, const float *pos, const float *neg) {
    const float step = 0.01f;      /* Setting.h:118 */
    for (i64 r = 0; r < o->R; r++) {
        i64 lo = o->valid_lef[r], hi = o->valid_rig[r];
        if (lo == -1) continue;
        i64 total = (hi - lo + 1) * 2;
        float mn = pos[lo], mx = pos[lo];
        for (i64 i = lo; i <= hi; i++) {
            if (pos[i] < mn) mn = pos[i];
            if (pos[i] > mx) mx = pos[i];
            if (neg[i] < mn) mn = neg[i];
            if (neg[i] > mx) mx = neg[i];
        }
        i64 n_int = (i64)((mx - mn) / step);
        float best_t = 0, best_a = 0;
        for (i64 i = 0; i <= n_int; i++) {
            float th = mn + i * step;
            i64 ok = 0;
            for (i64 j = lo; j <= hi; j++) { ok += pos[j] <= th; ok += neg[j] > th; }
            float acc = 1.0 * ok / total;
            if (i == 0 || acc > best_a) { best_a = acc; best_t = th; }
        }
        thresh[r] = best_t;
    }
}

/* ---------------------------------------------------------------- Test.h:347-387 */
/* counts[4] = TP, TN, FP, FN; returns accuracy as the reference stores it (float). */
float orc_tc_eval(const orc *o, const float *thresh, const float *pos, const float *neg, i64 *counts) {
    i64 TP = 0, TN = 0, FP = 0, FN = 0;
    for (i64 r = 0; r < o->R; r++) {
        if (o->valid_lef[r] == -1 || o->test_lef[r] == -1) continue;
        for (i64 i = o->test_lef[r]; i <= o->test_rig[r]; i++) {
            if (pos[i] <= thresh[r]) TP++; else FN++;
            if (neg[i] > thresh[r]) TN++; else FP++;
        }
    }
    counts[0] = TP; counts[1] = TN; counts[2] = FP; counts[3] = FN;
    return (float)(1.0 * (TP + TN) / (TP + TN + FP + FN));
}

/* ================================================================ model scores (canonical order) */
static void unit(const float *x, int D, float *y) {
    /* tf.nn.l2_normalize: x * rsqrt(max(sum x^2, 1e-12)) */
    float ss = 0.0f;
    for (int d = 0; d < D; d++) { float sq = x[d] * x[d]; ss = ss + sq; }
    float inv = 1.0f / sqrtf(ss > 1e-12f ? ss : 1e-12f);
    for (int d = 0; d < D; d++) y[d] = x[d] * inv;
}
static float dotseq(const float *a, const float *b, int D) {
    float s = 0.0f;
    for (int d = 0; d < D; d++) { float p = a[d] * b[d]; s = s + p; }
    return s;
}
static float l1_of(const float *h, const float *r, const float *t, int D) {
    /* _calc: abs(h + r - t), association (h + r) - t as written (TransE.py:15) */
    float s = 0.0f;
    for (int d = 0; d < D; d++) { float a = h[d] + r[d]; float b = a - t[d]; s = s + fabsf(b); }
    return s;
}

#define MAXD 2048
/* model: 0 TransE, 1 TransH, 2 TransR, 3 TransD.  Tables are row-major fp32.
 *   ent [E,De], rel [R,Dr]
 *   aux_ent: TransD ent_transfer [E,De]
 *   aux_rel: TransH normal_vectors [R,D]; TransD rel_transfer [R,Dr]; TransR transfer_matrix [R,De*Dr]
 * out[i] = predict for triple i: TransE = mean over d (TransE.py:58), others = sum (TransH.py:82 ...).
 * TransR uses the matrix of r[0] for every row (TransR.py:83). */
void orc_predict(int model, int De, int Dr, const float *ent, const float *rel, const float *aux_ent,
                 const float *aux_rel, const i64 *h, const i64 *t, const i64 *r, i64 n, float *out) {
    float hv[MAXD], tv[MAXD], rv[MAXD], nv[MAXD], hp[MAXD], tp[MAXD];
    for (i64 i = 0; i < n; i++) {
        const float *eh = ent + h[i] * De, *et = ent + t[i] * De, *er = rel + r[i] * Dr;
        if (model == 0) {
            unit(eh, De, hv); unit(et, De, tv); unit(er, Dr, rv);
            out[i] = l1_of(hv, rv, tv, De) / (float)De;
        } else if (model == 1) {
            unit(aux_rel + r[i] * De, De, nv);
            float dh = dotseq(eh, nv, De), dt = dotseq(et, nv, De);
            for (int d = 0; d < De; d++) { float a = dh * nv[d]; hp[d] = eh[d] - a; float b = dt * nv[d]; tp[d] = et[d] - b; }
            unit(hp, De, hv); unit(tp, De, tv); unit(er, De, rv);
            out[i] = l1_of(hv, rv, tv, De);
        } else if (model == 3) {
            const float *rt = aux_rel + r[i] * Dr;
            float ch = dotseq(eh, aux_ent + h[i] * De, De), ct = dotseq(et, aux_ent + t[i] * De, De);
            for (int d = 0; d < De; d++) { float a = ch * rt[d]; hp[d] = eh[d] + a; float b = ct * rt[d]; tp[d] = et[d] + b; }
            unit(hp, De, hv); unit(tp, De, tv); unit(er, Dr, rv);
            out[i] = l1_of(hv, rv, tv, Dr);
        } else {
            const float *M = aux_rel + r[0] * (i64)De * Dr;      /* row-major [De,Dr] */
            for (int k = 0; k < Dr; k++) {
                float a = 0.0f, b = 0.0f;
                for (int d = 0; d < De; d++) { float p = eh[d] * M[d * Dr + k]; a = a + p; float q = et[d] * M[d * Dr + k]; b = b + q; }
                hp[k] = a; tp[k] = b;
            }
            unit(hp, Dr, hv); unit(tp, Dr, tv); unit(er, Dr, rv);
            out[i] = l1_of(hv, rv, tv, Dr);
        }
    }
}

/* accessors used by the tests to compare index structures */
i64 orc_n_dedup(const orc *o) { return o->n; }
void orc_means(const orc *o, float *tph, float *hpt) {
    memcpy(tph, o->tph, sizeof(float) * o->R); memcpy(hpt, o->hpt, sizeof(float) * o->R);
}
