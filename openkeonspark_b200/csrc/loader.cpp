// Host-side dataset reader and index builder (C++), the native counterpart of the reference's
// base/Reader.h.  It parses the reference's text files (or takes arrays), builds the sorted copies
// and per-entity ranges the sampler and the ranker need, and uploads them once to HBM as int32
// structure-of-arrays.  Differences from the reference that do not change any result:
//   * ids are int32 on the device (the reference uses 3x int64 AoS, Triple.h:5-7);
//   * for every training row the runs of (h,r,*) / (*,r,t) / (h,t,*) in the sorted copies are
//     found HERE, once (the reference binary-searches them for every negative, Corrupt.h:9-24);
//   * the known-true sets used for filtered ranking are materialised per test triple instead of
//     being probed by _find() per candidate (Corrupt.h:104-115) — same membership relation.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "okb_internal.h"

// ------------------------------------------------------------------------------------------ DevBuf
int DevBuf::ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (external) return 1;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    if (cudaMalloc(&p, want) != cudaSuccess) { p = nullptr; return 1; }
    cap = want;
    return 0;
}
void DevBuf::release() { if (p && !external) cudaFree(p); p = nullptr; cap = 0; external = false; }

template <class T> static int upload(okb_ctx *c, T *&dst, const std::vector<T> &src) {
    if (dst) { cudaFree(dst); dst = nullptr; }
    size_t bytes = sizeof(T) * (src.empty() ? 1 : src.size());
    OKB_CUDA(c, cudaMalloc((void **)&dst, bytes));
    if (!src.empty()) OKB_CUDA(c, cudaMemcpy(dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice));
    return 0;
}

// ------------------------------------------------------------------------------------------ parsing
static bool slurp(const std::string &path, std::vector<char> &buf) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize(sz + 1);
    size_t got = fread(buf.data(), 1, sz, f);
    fclose(f);
    buf[got] = 0;
    buf.resize(got + 1);
    return true;
}
struct Tok {                                          // whitespace-separated integers
    const char *p, *e;
    explicit Tok(const std::vector<char> &b) : p(b.data()), e(b.data() + b.size() - 1) {}
    bool next(i64 &v) {
        while (p < e && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++;
        if (p >= e) return false;
        bool neg = false;
        if (*p == '-') { neg = true; p++; }
        if (p >= e || *p < '0' || *p > '9') return false;
        i64 x = 0;
        while (p < e && *p >= '0' && *p <= '9') x = x * 10 + (*p++ - '0');
        v = neg ? -x : x;
        return true;
    }
};
static bool read_count(const std::string &path, i64 &v) {
    std::vector<char> b;
    if (!slurp(path, b)) return false;
    Tok t(b);
    return t.next(v);
}
static int read_triples(okb_ctx *c, const std::string &path, std::vector<i64> &h, std::vector<i64> &t,
                        std::vector<i64> &r) {
    std::vector<char> b;
    if (!slurp(path, b)) { printf("`%s` does not exist\n", path.c_str()); OKB_FAIL(c, OKB_ERR_IO, path + " does not exist"); }
    Tok tk(b);
    i64 n = 0;
    if (!tk.next(n) || n < 0) OKB_FAIL(c, OKB_ERR_IO, path + ": bad header");
    h.resize(n); t.resize(n); r.resize(n);
    for (i64 i = 0; i < n; i++)
        if (!tk.next(h[i]) || !tk.next(t[i]) || !tk.next(r[i])) OKB_FAIL(c, OKB_ERR_IO, path + ": truncated");
    return 0;
}

// ------------------------------------------------------------------------------------------ packed keys
struct Packer {
    int be, br;
    bool ok;
    Packer(i64 E, i64 R) : be(bits_for(E > 1 ? E : 2)), br(bits_for(R > 1 ? R : 2)) { ok = 2 * be + br <= 63; }
    u64 ere(i64 a, i64 r, i64 b) const { return ((u64)a << (br + be)) | ((u64)r << be) | (u64)b; }   // (ent, rel, ent)
    u64 eer(i64 a, i64 b, i64 r) const { return ((u64)a << (be + br)) | ((u64)b << br) | (u64)r; }   // (ent, ent, rel)
    u64 ree(i64 r, i64 a, i64 b) const { return ((u64)r << (2 * be)) | ((u64)a << be) | (u64)b; }    // (rel, ent, ent)
    i64 hi(u64 k, int lo_bits) const { return (i64)(k >> lo_bits); }
};

template <class F> static void parallel_for(i64 n, F f) {
    unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (n < (1 << 16)) nt = 1;
    std::vector<std::thread> th;
    i64 per = (n + nt - 1) / nt;
    for (unsigned k = 0; k < nt; k++) {
        i64 lo = k * per, hi = std::min(n, lo + per);
        if (lo >= hi) break;
        th.emplace_back([=] { f(lo, hi); });
    }
    for (auto &x : th) x.join();
}

static void ranges_by_first(const std::vector<u64> &keys, int lo_bits, i64 slots, std::vector<i32> &lef,
                            std::vector<i32> &rig, i32 lef_default) {
    lef.assign(slots, lef_default);
    rig.assign(slots, -1);
    i64 n = keys.size();
    for (i64 i = 0; i < n; i++) {
        i64 a = (i64)(keys[i] >> lo_bits);
        if (i == 0 || (i64)(keys[i - 1] >> lo_bits) != a) lef[a] = (i32)i;
        if (i == n - 1 || (i64)(keys[i + 1] >> lo_bits) != a) rig[a] = (i32)i;
    }
}

// ------------------------------------------------------------------------------------------ train index
static int build_train(okb_ctx *c, const i64 *h, const i64 *t, const i64 *r, i64 n_raw) {
    const i64 E = c->E, R = c->R;
    Packer pk(E, R);
    if (!pk.ok) OKB_FAIL(c, OKB_ERR_ARG, "entity/relation id space too large for 63-bit packed keys");
    if (n_raw <= 0) OKB_FAIL(c, OKB_ERR_ARG, "empty training set");
    for (i64 i = 0; i < n_raw; i++)
        if (h[i] < 0 || h[i] >= E || t[i] < 0 || t[i] >= E || r[i] < 0 || r[i] >= R)
            OKB_FAIL(c, OKB_ERR_ARG, "train triple id out of range");
    c->n_raw = n_raw;
    c->raw_h.resize(n_raw); c->raw_t.resize(n_raw); c->raw_r.resize(n_raw);
    std::vector<u64> k_hrt(n_raw);
    for (i64 i = 0; i < n_raw; i++) {
        c->raw_h[i] = (i32)h[i]; c->raw_t[i] = (i32)t[i]; c->raw_r[i] = (i32)r[i];
        k_hrt[i] = pk.ere(h[i], r[i], t[i]);
    }
    std::sort(k_hrt.begin(), k_hrt.end());
    k_hrt.erase(std::unique(k_hrt.begin(), k_hrt.end()), k_hrt.end());     // Reader.h:103-123
    const i64 n = c->n = (i64)k_hrt.size();
    const u64 me = (1ull << pk.be) - 1, mr = (1ull << pk.br) - 1;
    std::vector<u64> k_trh(n), k_htr(n);
    std::vector<i64> freq(R, 0);
    c->byh_r.resize(n); c->byh_t.resize(n);
    for (i64 i = 0; i < n; i++) {
        i64 hh = (i64)(k_hrt[i] >> (pk.br + pk.be)), rr = (i64)((k_hrt[i] >> pk.be) & mr), tt = (i64)(k_hrt[i] & me);
        c->byh_r[i] = (i32)rr; c->byh_t[i] = (i32)tt;
        k_trh[i] = pk.ere(tt, rr, hh);
        k_htr[i] = pk.eer(hh, tt, rr);
        freq[rr]++;
    }
    std::thread s1([&] { std::sort(k_trh.begin(), k_trh.end()); });          // Reader.h:125-127
    std::sort(k_htr.begin(), k_htr.end());
    s1.join();
    c->byt_r.resize(n); c->byt_h.resize(n); c->byht_t.resize(n); c->byht_r.resize(n);
    for (i64 i = 0; i < n; i++) {
        c->byt_r[i] = (i32)((k_trh[i] >> pk.be) & mr); c->byt_h[i] = (i32)(k_trh[i] & me);
        c->byht_t[i] = (i32)((k_htr[i] >> pk.br) & me); c->byht_r[i] = (i32)(k_htr[i] & mr);
    }
    ranges_by_first(k_hrt, pk.br + pk.be, E, c->lef_h, c->rig_h, 0);        // Reader.h:130-158
    ranges_by_first(k_trh, pk.br + pk.be, E, c->lef_t, c->rig_t, 0);
    ranges_by_first(k_htr, pk.br + pk.be, E, c->lef_ht, c->rig_ht, 0);
    // Reader.h:160-177
    c->tph.assign(R, 0.f); c->hpt.assign(R, 0.f);
    for (i64 i = 0; i < n; i++) {
        if (i == 0 || (k_hrt[i] >> pk.be) != (k_hrt[i - 1] >> pk.be)) c->tph[c->byh_r[i]] += 1.0f;
        if (i == 0 || (k_trh[i] >> pk.be) != (k_trh[i - 1] >> pk.be)) c->hpt[c->byt_r[i]] += 1.0f;
    }
    for (i64 r_ = 0; r_ < R; r_++) {
        c->tph[r_] = (float)freq[r_] / c->tph[r_];
        c->hpt[r_] = (float)freq[r_] / c->hpt[r_];
    }
    {   // hub statistics: how long can a gradient-row segment get?  (picks the pre-reduction path of the update)
        i64 mr = 0, me_ = 0;
        for (i64 r_ = 0; r_ < R; r_++) mr = std::max(mr, freq[r_]);
        std::vector<i32> deg(E, 0);
        for (i64 i = 0; i < n; i++) { deg[(i64)(k_hrt[i] >> (pk.br + pk.be))]++; deg[c->byh_t[i]]++; }
        for (i64 e = 0; e < E; e++) me_ = std::max<i64>(me_, deg[e]);
        c->max_rel_share = (double)mr / (double)n;
        c->max_ent_share = (double)me_ / (double)(2 * n);
    }
    return okb_upload_train(c);
}

int okb_upload_train(okb_ctx *c) {
    const i64 n_raw = c->n_raw;
    std::vector<int4> raw(n_raw), run(n_raw);
    std::vector<int2> run_ht(n_raw);
    // For each file row the three runs it will be corrupted against (Corrupt.h:9-24, 41-56, 73-88).
    parallel_for(n_raw, [&](i64 lo, i64 hi) {
        for (i64 i = lo; i < hi; i++) {
            i32 h = c->raw_h[i], t = c->raw_t[i], r = c->raw_r[i];
            raw[i] = make_int4(h, t, r, 0);
            const i32 *b = c->byh_r.data();
            i32 ll = (i32)(std::lower_bound(b + c->lef_h[h], b + c->rig_h[h] + 1, r) - b);
            i32 rr = (i32)(std::upper_bound(b + c->lef_h[h], b + c->rig_h[h] + 1, r) - b) - 1;
            const i32 *b2 = c->byt_r.data();
            i32 ll2 = (i32)(std::lower_bound(b2 + c->lef_t[t], b2 + c->rig_t[t] + 1, r) - b2);
            i32 rr2 = (i32)(std::upper_bound(b2 + c->lef_t[t], b2 + c->rig_t[t] + 1, r) - b2) - 1;
            run[i] = make_int4(ll, rr, ll2, rr2);
            const i32 *b3 = c->byht_t.data();
            i32 ll3 = (i32)(std::lower_bound(b3 + c->lef_ht[h], b3 + c->rig_ht[h] + 1, t) - b3);
            i32 rr3 = (i32)(std::upper_bound(b3 + c->lef_ht[h], b3 + c->rig_ht[h] + 1, t) - b3) - 1;
            run_ht[i] = make_int2(ll3, rr3);
        }
    });
    std::vector<float> prob(c->R);
    for (i64 r = 0; r < c->R; r++) prob[r] = 1000 * c->hpt[r] / (c->hpt[r] + c->tph[r]);     // Base.cpp:117
    if (upload(c, c->d_raw, raw) || upload(c, c->d_run, run) || upload(c, c->d_run_ht, run_ht) ||
        upload(c, c->d_byh_t, c->byh_t) || upload(c, c->d_byt_h, c->byt_h) || upload(c, c->d_byht_r, c->byht_r) ||
        upload(c, c->d_prob, prob))
        return OKB_ERR_CUDA;
    return 0;
}

// ------------------------------------------------------------------------------------------ test index
static void sort_rht(const Packer &pk, const i64 *h, const i64 *t, const i64 *r, i64 n, std::vector<i32> &oh,
                     std::vector<i32> &ot, std::vector<i32> &orr) {
    std::vector<u64> k(n);
    for (i64 i = 0; i < n; i++) k[i] = pk.ree(r[i], h[i], t[i]);
    std::sort(k.begin(), k.end());                                          // Reader.h:260-261 (cmp_rel2)
    const u64 me = (1ull << pk.be) - 1;
    oh.resize(n); ot.resize(n); orr.resize(n);
    for (i64 i = 0; i < n; i++) { orr[i] = (i32)(k[i] >> (2 * pk.be)); oh[i] = (i32)((k[i] >> pk.be) & me); ot[i] = (i32)(k[i] & me); }
}
static void rel_ranges(const std::vector<i32> &r, i64 R, std::vector<i32> &lef, std::vector<i32> &rig) {
    lef.assign(R, -1); rig.assign(R, -1);                                   // Reader.h:267-291
    i64 n = r.size();
    for (i64 i = 0; i < n; i++) {
        if (i == 0 || r[i] != r[i - 1]) lef[r[i]] = (i32)i;
        if (i == n - 1 || r[i] != r[i + 1]) rig[r[i]] = (i32)i;
    }
}

static int build_test(okb_ctx *c, const i64 *th, const i64 *tt, const i64 *tr, i64 n_test, const i64 *vh,
                      const i64 *vt, const i64 *vr, i64 n_valid) {
    if (c->n_raw == 0) OKB_FAIL(c, OKB_ERR_STATE, "import the training files first");
    const i64 E = c->E, R = c->R;
    Packer pk(E, R);
    for (i64 i = 0; i < n_test; i++)
        if (th[i] < 0 || th[i] >= E || tt[i] < 0 || tt[i] >= E || tr[i] < 0 || tr[i] >= R) OKB_FAIL(c, OKB_ERR_ARG, "test triple id out of range");
    for (i64 i = 0; i < n_valid; i++)
        if (vh[i] < 0 || vh[i] >= E || vt[i] < 0 || vt[i] >= E || vr[i] < 0 || vr[i] >= R) OKB_FAIL(c, OKB_ERR_ARG, "valid triple id out of range");
    c->n_test = n_test; c->n_valid = n_valid; c->n_all = n_test + c->n_raw + n_valid;   // Reader.h:229
    sort_rht(pk, th, tt, tr, n_test, c->test_h, c->test_t, c->test_r);
    sort_rht(pk, vh, vt, vr, n_valid, c->valid_h, c->valid_t, c->valid_r);
    rel_ranges(c->test_r, R, c->test_lef, c->test_rig);
    rel_ranges(c->valid_r, R, c->valid_lef, c->valid_rig);
    c->tc_ranges_ready = false;                            // tc.cu re-uploads the ranges
    c->all_hrt.resize(c->n_all);
    i64 k = 0;
    for (i64 i = 0; i < n_test; i++) c->all_hrt[k++] = pk.ere(th[i], tr[i], tt[i]);
    for (i64 i = 0; i < c->n_raw; i++) c->all_hrt[k++] = pk.ere(c->raw_h[i], c->raw_r[i], c->raw_t[i]);
    for (i64 i = 0; i < n_valid; i++) c->all_hrt[k++] = pk.ere(vh[i], vr[i], vt[i]);
    std::sort(c->all_hrt.begin(), c->all_hrt.end());                        // Reader.h:259
    return okb_upload_test(c);
}

int okb_upload_test(okb_ctx *c) {
    Packer pk(c->E, c->R);
    const u64 me = (1ull << pk.be) - 1, mr = (1ull << pk.br) - 1;
    // unique known triples, in (h,r,t) and (t,r,h) order
    std::vector<u64> u(c->all_hrt);
    u.erase(std::unique(u.begin(), u.end()), u.end());
    std::vector<u64> v(u.size());
    std::vector<i32> known_t(u.size()), known_h(u.size());
    for (size_t i = 0; i < u.size(); i++) {
        i64 h = (i64)(u[i] >> (pk.br + pk.be)), r = (i64)((u[i] >> pk.be) & mr), t = (i64)(u[i] & me);
        known_t[i] = (i32)t;
        v[i] = pk.ere(t, r, h);
    }
    std::sort(v.begin(), v.end());
    for (size_t i = 0; i < v.size(); i++) known_h[i] = (i32)(v[i] & me);
    std::vector<int4> trun(c->n_test);
    parallel_for(c->n_test, [&](i64 lo, i64 hi) {
        for (i64 i = lo; i < hi; i++) {
            u64 p = pk.ere(c->test_h[i], c->test_r[i], 0), q = pk.ere(c->test_t[i], c->test_r[i], 0);
            i32 a = (i32)(std::lower_bound(u.begin(), u.end(), p) - u.begin());
            i32 b = (i32)(std::lower_bound(u.begin(), u.end(), p + (1ull << pk.be)) - u.begin());
            i32 a2 = (i32)(std::lower_bound(v.begin(), v.end(), q) - v.begin());
            i32 b2 = (i32)(std::lower_bound(v.begin(), v.end(), q + (1ull << pk.be)) - v.begin());
            trun[i] = make_int4(a, b, a2, b2);
        }
    });
    c->grp_rel.clear(); c->grp_lo.clear(); c->grp_hi.clear();
    for (i64 i = 0; i < c->n_test; i++)
        if (i == 0 || c->test_r[i] != c->test_r[i - 1]) {
            c->grp_rel.push_back(c->test_r[i]); c->grp_lo.push_back((i32)i);
            c->grp_hi.push_back(c->test_rig[c->test_r[i]] + 1);
        }
    if (upload(c, c->d_test_h, c->test_h) || upload(c, c->d_test_t, c->test_t) || upload(c, c->d_test_r, c->test_r) ||
        upload(c, c->d_known_t, known_t) || upload(c, c->d_known_h, known_h) || upload(c, c->d_test_run, trun))
        return OKB_ERR_CUDA;
    return 0;
}

bool okb_host_find(const okb_ctx *c, i64 h, i64 t, i64 r) {                // Corrupt.h:104-115
    Packer pk(c->E, c->R);
    return std::binary_search(c->all_hrt.begin(), c->all_hrt.end(), pk.ere(h, r, t));
}

// ------------------------------------------------------------------------------------------ id lists
static int read_lists(okb_ctx *c, const std::string &path, i64 slots, bool header_is_count, Lists &A, Lists &B) {
    std::vector<char> buf;
    if (!slurp(path, buf)) { printf("`%s` does not exist\n", path.c_str()); OKB_FAIL(c, OKB_ERR_IO, path + " does not exist"); }
    Tok tk(buf);
    i64 n = 0;
    tk.next(n);
    // type_constrain.txt: the reference ignores the header and reads relationTotal entries
    // (Reader.h:318-331); a file with fewer entries makes it re-use stale values.  We read the
    // entries that are present.
    (void)header_is_count;
    std::vector<std::vector<i32>> la(slots), lb(slots);
    for (i64 i = 0;; i++) {
        i64 key, cnt, x;
        if (!tk.next(key) || !tk.next(cnt)) break;
        if (key < 0 || key >= slots) OKB_FAIL(c, OKB_ERR_IO, path + ": key out of range");
        la[key].clear();
        for (i64 j = 0; j < cnt; j++) { if (!tk.next(x)) OKB_FAIL(c, OKB_ERR_IO, path + ": truncated"); la[key].push_back((i32)x); }
        if (!tk.next(key) || !tk.next(cnt)) OKB_FAIL(c, OKB_ERR_IO, path + ": truncated");
        if (key < 0 || key >= slots) OKB_FAIL(c, OKB_ERR_IO, path + ": key out of range");
        lb[key].clear();
        for (i64 j = 0; j < cnt; j++) { if (!tk.next(x)) OKB_FAIL(c, OKB_ERR_IO, path + ": truncated"); lb[key].push_back((i32)x); }
    }
    auto pack = [&](std::vector<std::vector<i32>> &l, Lists &L) {
        L.lef.assign(slots, 0); L.rig.assign(slots, 0); L.ids.clear();
        for (i64 k = 0; k < slots; k++) {
            std::sort(l[k].begin(), l[k].end());
            L.lef[k] = (i32)L.ids.size();
            L.ids.insert(L.ids.end(), l[k].begin(), l[k].end());
            L.rig[k] = (i32)L.ids.size();
        }
    };
    pack(la, A); pack(lb, B);
    return 0;
}
static int upload_lists_one(okb_ctx *c, Lists &L, i64 slots) {
    if (L.lef.empty()) { L.lef.assign(slots, 0); L.rig.assign(slots, 0); }
    return upload(c, L.d_lef, L.lef) || upload(c, L.d_rig, L.rig) || upload(c, L.d_ids, L.ids);
}
int okb_upload_lists(okb_ctx *c) {
    if (upload_lists_one(c, c->head_type, c->R) || upload_lists_one(c, c->tail_type, c->R) ||
        upload_lists_one(c, c->sup, c->E) || upload_lists_one(c, c->sub, c->E))
        return OKB_ERR_CUDA;
    return 0;
}

// ------------------------------------------------------------------------------------------ C ABI
extern "C" {

int okb_set_in_path(okb_ctx *c, const char *path) {
    c->in_path = path ? path : "";
    if (!c->in_path.empty() && c->in_path.back() != '/') c->in_path += '/';
    return 0;
}
int okb_set_bern(okb_ctx *c, INT flag) { c->bern = flag; return 0; }
int okb_set_work_threads(okb_ctx *c, INT w) {
    if (w < 1) OKB_FAIL(c, OKB_ERR_ARG, "workThreads must be >= 1");
    c->W = w;
    return 0;
}

int okb_import_train_arrays(okb_ctx *c, INT n_ent, INT n_rel, const INT *h, const INT *t, const INT *r, INT n,
                            INT new_batch_total) {
    c->E = n_ent; c->R = n_rel; c->new_batch = new_batch_total;
    return build_train(c, h, t, r, n);
}
int okb_import_train_files(okb_ctx *c) {
    printf("The toolkit is importing datasets.\n");
    i64 v;
    if (!read_count(c->in_path + "relation2id.txt", v)) { printf("`%srelation2id.txt` does not exist\n", c->in_path.c_str()); OKB_FAIL(c, OKB_ERR_IO, "relation2id.txt missing"); }
    c->R = v;
    printf("The total of relations is %ld.\n", (long)c->R);
    if (!read_count(c->in_path + "entity2id.txt", v)) { printf("`%sentity2id.txt` does not exist\n", c->in_path.c_str()); OKB_FAIL(c, OKB_ERR_IO, "entity2id.txt missing"); }
    c->E = v;
    printf("The total of entities is %ld.\n", (long)c->E);
    c->new_batch = 0;
    if (read_count(c->in_path + "batch2id.txt", v)) {                        // Reader.h:61-67
        c->new_batch = v;
        printf("`%sbatch2id.txt` founded!\nThe total number of new batch triples is: %ld\n", c->in_path.c_str(), (long)v);
    }
    std::vector<i64> h, t, r;
    int rc = read_triples(c, c->in_path + "train2id.txt", h, t, r);
    if (rc) return rc;
    printf("The total of train triples is %ld.\n", (long)h.size());
    return build_train(c, h.data(), t.data(), r.data(), (i64)h.size());
}
int okb_import_test_arrays(okb_ctx *c, const INT *th, const INT *tt, const INT *tr, INT n_test, const INT *vh,
                           const INT *vt, const INT *vr, INT n_valid) {
    return build_test(c, th, tt, tr, n_test, vh, vt, vr, n_valid);
}
int okb_import_test_files(okb_ctx *c) {
    std::vector<i64> th, tt, tr, vh, vt, vr;
    int rc = read_triples(c, c->in_path + "test2id.txt", th, tt, tr);
    if (rc) return rc;
    rc = read_triples(c, c->in_path + "valid2id.txt", vh, vt, vr);
    if (rc) return rc;
    rc = build_test(c, th.data(), tt.data(), tr.data(), (i64)th.size(), vh.data(), vt.data(), vr.data(), (i64)vh.size());
    if (rc) return rc;
    printf("The total of test triples is %ld.\n", (long)c->n_test);
    printf("The total of valid triples is %ld.\n", (long)c->n_valid);
    return 0;
}
int okb_import_type_files(okb_ctx *c) {
    int rc = read_lists(c, c->in_path + "type_constrain.txt", c->R, false, c->head_type, c->tail_type);
    if (rc) return rc;
    c->have_types = true;
    return okb_upload_lists(c);
}
// Type constraints derived from the loaded graph exactly as the reference's n_n() writes them
// (main_spark.py:209-290): per relation, the set of heads / tails seen in train + valid + test.
int okb_build_type_constraints(okb_ctx *c) {
    if (c->all_hrt.empty()) OKB_FAIL(c, OKB_ERR_STATE, "import train and test data first");
    Packer pk(c->E, c->R);
    const u64 me = (1ull << pk.be) - 1, mr = (1ull << pk.br) - 1;
    std::vector<u64> hk, tk;                              // (r, h) and (r, t) pairs
    hk.reserve(c->all_hrt.size()); tk.reserve(c->all_hrt.size());
    for (u64 k : c->all_hrt) {
        const u64 h = k >> (pk.br + pk.be), r = (k >> pk.be) & mr, t = k & me;
        hk.push_back((r << pk.be) | h); tk.push_back((r << pk.be) | t);
    }
    auto build = [&](std::vector<u64> &v, Lists &L) {
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        L.lef.assign(c->R, 0); L.rig.assign(c->R, 0); L.ids.resize(v.size());
        size_t i = 0;
        for (i64 r = 0; r < c->R; r++) {
            L.lef[r] = (i32)i;
            while (i < v.size() && (i64)(v[i] >> pk.be) == r) { L.ids[i] = (i32)(v[i] & me); i++; }
            L.rig[r] = (i32)i;
        }
    };
    build(hk, c->head_type); build(tk, c->tail_type);
    c->have_types = true;
    return okb_upload_lists(c);
}
int okb_import_ontology_files(okb_ctx *c) {
    printf("Reading %sontology_constrain.txt\n", c->in_path.c_str());
    auto clear = [&](Lists &L) { L.lef.assign(c->E, 0); L.rig.assign(c->E, 0); L.ids.clear(); };
    clear(c->sup); clear(c->sub);
    int rc = read_lists(c, c->in_path + "ontology_constrain.txt", c->E, true, c->sup, c->sub);
    if (rc == OKB_ERR_IO) { clear(c->sup); clear(c->sub); }   // Reader.h:384-388: a missing file leaves empty lists (every class = 3)
    else if (rc) return rc;
    c->have_onto = true;
    return okb_upload_lists(c);
}
INT okb_total(okb_ctx *c, int what) {
    switch (what) {
        case 0: return c->E; case 1: return c->R; case 2: return c->n_raw; case 3: return c->n;
        case 4: return c->n_test; case 5: return c->n_valid; case 6: return c->n_all; case 7: return c->new_batch;
    }
    return -1;
}
int okb_test_list(okb_ctx *c, int which, INT *h, INT *t, INT *r) {
    const std::vector<i32> &H = which ? c->valid_h : c->test_h, &T = which ? c->valid_t : c->test_t, &Rr = which ? c->valid_r : c->test_r;
    for (size_t i = 0; i < H.size(); i++) { h[i] = H[i]; t[i] = T[i]; r[i] = Rr[i]; }
    return 0;
}

// ---------------------------------------------------------------- triple classification (host, tiny)
// Corrupt.h:118-137: a type-constrained tail drawn with libc rand(), rejected while known, with the
// filtered sampler on stream 0 as the 1000-tries fallback.
static i64 tc_negative_tail(okb_ctx *c, i64 h, i64 r) {
    i64 ll = c->tail_type.lef[r], rr = c->tail_type.rig[r];
    if (rr > ll)
        for (int tries = 0; tries < 1000; tries++) {
            i64 t = c->tail_type.ids[(rand() % (rr - ll)) + ll];
            if (!okb_host_find(c, h, t, r)) return t;
        }
    return okb_host_new_tail(c, h, r);
}
int okb_tc_batch(okb_ctx *c, int which, INT *ph, INT *pt, INT *pr, INT *nh, INT *nt, INT *nr) {
    if (!c->have_types || c->all_hrt.empty()) OKB_FAIL(c, OKB_ERR_STATE, "import test and type files first");
    const std::vector<i32> &H = which ? c->valid_h : c->test_h, &T = which ? c->valid_t : c->test_t, &Rr = which ? c->valid_r : c->test_r;
    std::vector<i32> &neg = which ? c->neg_valid_t : c->neg_test_t;
    neg.resize(H.size());
    for (size_t i = 0; i < H.size(); i++) neg[i] = (i32)tc_negative_tail(c, H[i], Rr[i]);   // Test.h:258-274
    if (ph)
        for (size_t i = 0; i < H.size(); i++) {
            ph[i] = nh[i] = H[i]; pr[i] = nr[i] = Rr[i]; pt[i] = T[i]; nt[i] = neg[i];
        }
    return 0;
}
static bool score_range(const okb_ctx *c, i64 r, const REAL *pos, const REAL *neg, float &mn, float &mx) {
    i64 lo = c->valid_lef[r], hi = c->valid_rig[r];
    if (lo == -1) return false;
    mn = mx = pos[lo];
    for (i64 i = lo; i <= hi; i++) {
        if (pos[i] < mn) mn = pos[i];
        if (pos[i] > mx) mx = pos[i];
        if (neg[i] < mn) mn = neg[i];
        if (neg[i] > mx) mx = neg[i];
    }
    return true;
}
static const float kInterval = 0.01f;     // Setting.h:118
// Host-pointer entry points (the reference's ctypes calls pass numpy buffers): the scores are staged on the device and
// the threshold search / counting run in tc.cu's kernels; nothing is computed on the CPU.
int okb_tc_thresholds_dev(okb_ctx *c, const float *pos, const float *neg, float *thresh, void *stream);
int okb_tc_counts_dev(okb_ctx *c, const float *thresh, const float *pos, const float *neg, int on_valid, INT *tp_tn_fp_fn, REAL *acc, void *stream);
static int tc_stage(okb_ctx *c, const REAL *thresh, const REAL *pos, const REAL *neg, i64 n, float *&d_th, float *&d_pos, float *&d_neg) {
    if (c->valid_lef.empty()) OKB_FAIL(c, OKB_ERR_STATE, "import test files first");
    if (c->tc_io.ensure(sizeof(float) * (size_t)(c->R + 2 * std::max<i64>(n, 1)))) OKB_FAIL(c, OKB_ERR_CUDA, "out of device memory");
    d_th = c->tc_io.as<float>(); d_pos = d_th + c->R; d_neg = d_pos + n;
    OKB_CUDA(c, cudaMemcpy(d_th, thresh, sizeof(float) * c->R, cudaMemcpyHostToDevice));
    OKB_CUDA(c, cudaMemcpy(d_pos, pos, sizeof(float) * n, cudaMemcpyHostToDevice));
    OKB_CUDA(c, cudaMemcpy(d_neg, neg, sizeof(float) * n, cudaMemcpyHostToDevice));
    return 0;
}
int okb_best_threshold(okb_ctx *c, REAL *thresh, const REAL *pos, const REAL *neg) {   // Test.h:304-341
    float *d_th, *d_pos, *d_neg;
    int rc = tc_stage(c, thresh, pos, neg, c->n_valid, d_th, d_pos, d_neg);
    if (rc) return rc;
    if ((rc = okb_tc_thresholds_dev(c, d_pos, d_neg, d_th, nullptr))) return rc;
    OKB_CUDA(c, cudaMemcpy(thresh, d_th, sizeof(float) * c->R, cudaMemcpyDeviceToHost));
    return 0;
}
int okb_tc_eval(okb_ctx *c, const REAL *thresh, const REAL *pos, const REAL *neg, INT *cnt, REAL *acc) {   // Test.h:347-387
    float *d_th, *d_pos, *d_neg;
    int rc = tc_stage(c, thresh, pos, neg, c->n_test, d_th, d_pos, d_neg);
    if (rc) return rc;
    return okb_tc_counts_dev(c, d_th, d_pos, d_neg, 0, cnt, acc, nullptr);
}
// Accuracy of the thresholds on the VALID triples (early stopping, distribute_training.py:299-316).  The reference calls
// test_triple_classification with the valid scores there, i.e. it indexes arrays of validTotal scores with the TEST
// ranges (Test.h:355-366) — an out-of-bounds read whenever testTotal > validTotal.  The evident intent, accuracy over
// the valid set with the thresholds just fitted on it, is what this entry point computes.
int okb_tc_eval_valid(okb_ctx *c, const REAL *thresh, const REAL *pos, const REAL *neg, INT *cnt, REAL *acc) {
    float *d_th, *d_pos, *d_neg;
    int rc = tc_stage(c, thresh, pos, neg, c->n_valid, d_th, d_pos, d_neg);
    if (rc) return rc;
    return okb_tc_counts_dev(c, d_th, d_pos, d_neg, 1, cnt, acc, nullptr);
}
INT okb_n_interval(okb_ctx *c, INT r, const REAL *pos, const REAL *neg) {
    float mn, mx;
    if (c->valid_lef.empty() || !score_range(c, r, pos, neg, mn, mx)) return 0;
    return (i64)((mx - mn) / kInterval);
}
INT *okb_tpfp(okb_ctx *c, INT r, const REAL *pos, const REAL *neg, const REAL *pos_test, const REAL *neg_test) {
    float mn, mx;
    if (c->valid_lef.empty() || !score_range(c, r, pos, neg, mn, mx)) return nullptr;
    i64 n_int = (i64)((mx - mn) / kInterval);
    c->tpfp.assign((n_int + 1) * 2, 0);
    for (i64 i = 0; i <= n_int; i++) {
        float th = mn + i * kInterval;
        i64 TP = 0, FP = 0;
        for (i64 j = c->test_lef[r]; j <= c->test_rig[r] && c->test_lef[r] != -1; j++) { TP += pos_test[j] <= th; FP += neg_test[j] <= th; }
        c->tpfp[i] = TP; c->tpfp[i + n_int + 1] = FP;
    }
    return c->tpfp.data();
}

}  // extern "C"
