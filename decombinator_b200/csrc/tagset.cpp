// tagset.cpp -- host builder of the flattened tag tables (see dcb_tables.h).
//
// Replaces get_v_tags/get_j_tags and the AcoraBuilder().add/.build() calls of import_tcr_info
// (/root/reference/src/decombinator/decombine.py:698-746, 820-866).
#include "dcb_internal.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace {

inline int base_code(char c) {
    switch (c) {
        case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3;
        default: return -1;
    }
}

bool pack64(const std::string& s, size_t from, size_t n, uint32_t& lo, uint32_t& hi) {
    uint64_t v = 0;
    lo = hi = 0;
    for (size_t i = 0; i < n; i++) {
        int c = base_code(s[from + i]);
        if (c < 0) return false;
        v |= (uint64_t)c << (2 * i);
    }
    lo = (uint32_t)v; hi = (uint32_t)(v >> 32);
    return true;
}

uint32_t pow2_at_least(uint32_t x) {
    uint32_t p = 16;
    while (p < x) p <<= 1;
    return p;
}

struct Blob {
    std::vector<uint32_t> w;
    int32_t reserve(size_t n_words) {
        int32_t off = (int32_t)w.size();
        w.resize(w.size() + n_words, 0u);
        return off;
    }
    void align4() { while (w.size() % 4) w.push_back(0u); }  // 16-byte granularity for bulk copies
};

// 2-choice cuckoo table: a key lives in slot h1(key) or h2(key).
struct Cuckoo {
    uint32_t c1 = 1, c2 = 1;
    int bits = 4;
    std::vector<uint32_t> key, val;
    std::vector<uint8_t> used;
    uint32_t h(uint32_t k, int which) const { return (k * (which ? c2 : c1)) >> (32 - bits); }
    bool build(const std::vector<std::pair<uint32_t, uint32_t>>& items, int bits_, uint32_t c1_, uint32_t c2_) {
        bits = bits_; c1 = c1_ | 1u; c2 = c2_ | 1u;
        const size_t n = (size_t)1 << bits;
        key.assign(n, 0); val.assign(n, 0); used.assign(n, 0);
        for (auto it : items) {
            uint32_t k = it.first, v = it.second;
            int which = 0;
            bool placed = false;
            for (int kick = 0; kick < 500; kick++) {
                uint32_t s = h(k, which);
                if (!used[s]) { used[s] = 1; key[s] = k; val[s] = v; placed = true; break; }
                uint32_t s2 = h(k, which ^ 1);
                if (!used[s2]) { used[s2] = 1; key[s2] = k; val[s2] = v; placed = true; break; }
                std::swap(k, key[s]); std::swap(v, val[s]);  // evict the occupant of the first choice
                which = (h(k, 0) == s) ? 1 : 0;              // and send it to its other slot
            }
            if (!placed) return false;
        }
        return true;
    }
};

bool build_cuckoo(Cuckoo& ck, const std::vector<std::pair<uint32_t, uint32_t>>& items) {
    int bits = 4;
    while (((size_t)1 << bits) < 2 * items.size() + 2) bits++;
    uint64_t rng = 0x2545F4914F6CDD1Dull;
    for (int grow = 0; grow < 6; grow++, bits++) {
        for (int attempt = 0; attempt < 64; attempt++) {
            rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
            if (ck.build(items, bits, (uint32_t)rng, (uint32_t)(rng >> 32))) return true;
        }
    }
    return false;
}

// One keyword set -> bitmap + hash + DcbKw array + tag list (general kernel).
void build_kwset(Blob& b, DcbKwSet& ks, const std::vector<std::string>& list) {
    std::vector<std::string> kws;
    std::vector<std::vector<int>> members;
    std::map<std::string, int> seen;
    for (size_t i = 0; i < list.size(); i++) {
        if (list[i].empty()) continue;  // AcoraBuilder drops empty keywords
        auto it = seen.find(list[i]);
        if (it == seen.end()) {
            seen[list[i]] = (int)kws.size();
            kws.push_back(list[i]);
            members.push_back({(int)i});
        } else {
            members[it->second].push_back((int)i);
        }
    }
    int min_len = 1 << 30, max_len = 0;
    for (auto& k : kws) { min_len = std::min(min_len, (int)k.size()); max_len = std::max(max_len, (int)k.size()); }
    if (kws.empty()) { min_len = 1; max_len = 1; }
    ks.n_kw = (int)kws.size();
    ks.min_len = min_len; ks.max_len = max_len;
    ks.kq = std::min(min_len, 6);   // 4^6 bits = 512 B per set: the scans only visit marked positions, a sharper bitmap buys nothing
    struct Ent { uint32_t key; int len; int id; };
    std::vector<Ent> ents;
    for (size_t i = 0; i < kws.size(); i++) {
        uint32_t lo, hi;
        pack64(kws[i], kws[i].size() - ks.kq, ks.kq, lo, hi);
        ents.push_back({lo, (int)kws[i].size(), (int)i});
    }
    std::stable_sort(ents.begin(), ents.end(), [](const Ent& a, const Ent& c) {
        if (a.key != c.key) return a.key < c.key;
        return a.len > c.len;  // longest first = findall() order at equal end position
    });
    size_t bitmap_words = ((size_t)1 << (2 * ks.kq)) / 32;
    if (bitmap_words == 0) bitmap_words = 1;
    ks.bitmap_off = b.reserve(bitmap_words);
    size_t n_groups = 0;
    for (size_t i = 0; i < ents.size(); i++) if (i == 0 || ents[i].key != ents[i - 1].key) n_groups++;
    uint32_t hsize = pow2_at_least((uint32_t)(2 * n_groups + 1));
    ks.hash_mask = (int32_t)(hsize - 1);
    ks.hash_off = b.reserve(hsize);
    for (uint32_t i = 0; i < hsize; i++) b.w[ks.hash_off + i] = DCB_HASH_EMPTY;
    ks.kw_off = b.reserve(4 * ents.size());
    size_t total_tags = 0;
    for (auto& m : members) total_tags += m.size();
    ks.taglist_off = b.reserve((total_tags + 3) / 4 + 1);
    uint8_t* taglist = reinterpret_cast<uint8_t*>(&b.w[ks.taglist_off]);
    size_t tl = 0;
    for (size_t i = 0; i < ents.size(); i++) {
        const std::string& s = kws[ents[i].id];
        DcbKw kw;
        std::memset(&kw, 0, sizeof(kw));
        pack64(s, 0, s.size(), kw.bits_lo, kw.bits_hi);
        kw.len = (uint8_t)s.size();
        kw.first_tag = (uint8_t)members[ents[i].id][0];
        kw.n_tags = (uint8_t)members[ents[i].id].size();
        kw.tags_off = (uint8_t)tl;
        for (int t : members[ents[i].id]) taglist[tl++] = (uint8_t)t;
        std::memcpy(&b.w[ks.kw_off + 4 * i], &kw, sizeof(kw));
        b.w[ks.bitmap_off + (ents[i].key >> 5)] |= 1u << (ents[i].key & 31);
        if (i == 0 || ents[i].key != ents[i - 1].key) {
            size_t cnt = 1;
            while (i + cnt < ents.size() && ents[i + cnt].key == ents[i].key) cnt++;
            uint32_t h = dcb_hash32(ents[i].key) & (hsize - 1);
            while (b.w[ks.hash_off + h] != DCB_HASH_EMPTY) h = (h + 1) & (hsize - 1);
            b.w[ks.hash_off + h] = (ents[i].key << 16) | ((uint32_t)i << 8) | (uint32_t)cnt;
        }
    }
}

// Hash-and-displace table over the indexed q-mers (see DcbSeedIndex): buckets are placed largest first, each with the
// smallest displacement that drops all its keys into free slots.
struct Chd {
    uint32_t m1 = 1, m2 = 1;
    int b1 = 1, b2 = 1;
    std::vector<uint16_t> disp, slot;
    bool build(const std::vector<std::pair<uint32_t, uint32_t>>& items, int b1_, int b2_, uint32_t m1_, uint32_t m2_) {
        b1 = b1_; b2 = b2_; m1 = m1_; m2 = m2_;
        const size_t n1 = (size_t)1 << b1, n2 = (size_t)1 << b2;
        std::vector<std::vector<size_t>> bucket(n1);
        for (size_t i = 0; i < items.size(); i++) bucket[(items[i].first * m1) >> (32 - b1)].push_back(i);
        std::vector<size_t> order(n1);
        for (size_t i = 0; i < n1; i++) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t c) { return bucket[a].size() > bucket[c].size(); });
        disp.assign(n1, 0); slot.assign(n2, 0);
        std::vector<uint8_t> used(n2, 0);
        std::vector<uint32_t> base, at;
        for (size_t bi : order) {
            const auto& bk = bucket[bi];
            if (bk.empty()) break;
            base.clear();
            for (size_t i : bk) base.push_back((items[i].first * m2) >> (32 - b2));
            for (size_t i = 0; i < base.size(); i++)
                for (size_t k = i + 1; k < base.size(); k++)
                    if (base[i] == base[k]) return false;           // collide under every displacement
            bool placed = false;
            for (uint32_t d = 0; d < n2 && d < 65536u && !placed; d++) {
                bool ok = true;
                for (uint32_t x : base) if (used[(x + d) & (n2 - 1)]) { ok = false; break; }
                if (!ok) continue;
                for (size_t i = 0; i < bk.size(); i++) {
                    const uint32_t s = (base[i] + d) & (uint32_t)(n2 - 1);
                    used[s] = 1; slot[s] = (uint16_t)items[bk[i]].second;
                }
                disp[bi] = (uint16_t)d;
                placed = true;
            }
            if (!placed) return false;
        }
        return true;
    }
};

bool build_chd(Chd& chd, const std::vector<std::pair<uint32_t, uint32_t>>& items, int q) {
    if (2 * q > 30) return false;
    int b2 = 4;
    while (((size_t)1 << b2) < 2 * items.size() + 2) b2++;
    int b1 = b2 > 4 ? b2 - 3 : 1;                                   // ~2 keys per bucket on average at load <= 0.5
    if (b2 > 2 * q) return false;
    uint64_t rng = 0xD1B54A32D192ED03ull;
    for (int attempt = 0; attempt < 4096; attempt++) {
        rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
        const uint32_t m1 = ((uint32_t)rng | 1u) << (32 - 2 * q), m2 = ((uint32_t)(rng >> 32) | 1u) << (32 - 2 * q);
        if (chd.build(items, b1, b2, m1, m2)) return true;
    }
    return false;
}

int seed_q(int lmin) { return lmin >= 18 ? 9 : lmin >= 12 ? 8 : lmin >= 8 ? 6 : lmin; }

}  // namespace

// Seed index over the tags of one gene (the other pointer null) or of both genes of a chain (same lmin).
// wbits: log2(words) of the seed filter.
bool dcb_build_seed_index(const std::vector<std::string>* gene_v, const std::vector<std::string>* gene_j, int lmin,
                          int wbits, std::vector<uint32_t>& out) {
    Blob b;
    DcbSeedIndex idx;
    std::memset(&idx, 0, sizeof(idx));
    b.reserve((sizeof(DcbSeedIndex) + 3) / 4);
    b.align4();
    idx.q = seed_q(lmin);
    idx.lmin = lmin;
    idx.max_off = DCB_IDX_MAXOFF(lmin, idx.q);
    idx.stride = DCB_IDX_STRIDE(lmin, idx.q);
    idx.wlead = DCB_IDX_WLEAD(lmin, idx.q);
    idx.wbits = wbits;
    idx.span = DCB_IDX_SPAN(lmin, idx.q);            // two classes of offsets: [0, span) and [span, max_off]
    idx.k = DCB_IDX_K(lmin, idx.q);
    if (idx.k < 1 || idx.k > 15) return false;
    if (idx.q < 3 || idx.q > 9 || lmin > 32 || wbits < 2 || wbits > 2 * idx.q - 5 || wbits > 27) return false;
    idx.bmul = DCB_BLOOM_MUL(idx.q);

    // combined tag list: V first
    std::vector<const std::string*> all;
    idx.n_v = gene_v ? (int32_t)gene_v->size() : 0;
    if (gene_v) for (auto& s : *gene_v) all.push_back(&s);
    if (gene_j) for (auto& s : *gene_j) all.push_back(&s);
    idx.n_tags = (int32_t)all.size();
    if (all.empty() || all.size() > 510) return false;

    std::map<uint32_t, uint32_t> seeds;                           // indexed q-mers (the filter's keys)
    std::map<uint32_t, uint32_t> classkeys;                       // class << 30 | k-mer -> set of offsets
    std::map<std::pair<uint32_t, uint32_t>, int> prefixes;        // lmin-prefix -> first ctag
    std::vector<int> next_same(all.size(), 0x1FF);
    for (size_t t = 0; t < all.size(); t++) {
        const std::string& s = *all[t];
        if ((int)s.size() < lmin || s.size() > DCB_FAST_MAX_TAG_LEN) return false;
        uint32_t lo, hi;
        for (int o = 0; o <= idx.max_off; o++) {
            if (!pack64(s, o, idx.q, lo, hi)) return false;
            seeds[lo] |= 1u << o;
            const int c = o / idx.span;
            pack64(s, o - c * idx.span, idx.k, lo, hi);               // the k-mer starts c*span before the seed
            classkeys[((uint32_t)c << 30) | lo] |= 1u << o;
        }
        pack64(s, 0, lmin, lo, hi);
        auto key = std::make_pair(lo, hi);
        auto it = prefixes.find(key);
        if (it == prefixes.end()) prefixes[key] = (int)t;
        else {                                                    // chain the tags that share a prefix, ascending
            int k = it->second;
            while (next_same[k] != 0x1FF) k = next_same[k];
            next_same[k] = (int)t;
        }
    }
    {   // class keys
        std::vector<std::pair<uint32_t, uint32_t>> items(classkeys.begin(), classkeys.end());
        Cuckoo ck;
        if (!build_cuckoo(ck, items)) return false;
        idx.c1 = ck.c1; idx.c2 = ck.c2; idx.cshift = 32 - ck.bits;
        const size_t slots = (size_t)1 << ck.bits;
        idx.ck_off = b.reserve(slots);
        for (size_t i = 0; i < slots; i++) {
            if (!ck.used[i]) { b.w[idx.ck_off + i] = 0u; continue; }
            // the fingerprint comes from the product that did NOT choose this slot
            const uint32_t other = ck.h(ck.key[i], 0) == i ? ck.key[i] * ck.c2 : ck.key[i] * ck.c1;
            b.w[idx.ck_off + i] = (other & DCB_CK_FPMASK) | ck.val[i];
        }
    }
    {   // lmin-prefixes: search a multiplier that spreads them over the slots without a collision
        std::vector<std::pair<uint32_t, int>> items;
        for (auto& kv : prefixes) items.emplace_back(dcb_fold64(kv.first.first, kv.first.second), kv.second);
        for (size_t i = 0; i < items.size(); i++)
            for (size_t k = i + 1; k < items.size(); k++)
                if (items[i].first == items[k].first) return false;   // two prefixes fold to the same word
        uint64_t rng = 0x9E3779B97F4A7C15ull;
        bool done = false;
        for (int bits = 9; bits <= 14 && !done; bits++) {
            if (((size_t)1 << bits) < 2 * items.size()) continue;
            std::vector<int> slot((size_t)1 << bits);
            for (int attempt = 0; attempt < 20000 && !done; attempt++) {
                rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
                const uint32_t m = (uint32_t)rng | 1u;
                std::fill(slot.begin(), slot.end(), -1);
                bool ok = true;
                for (auto& it : items) {
                    int& sl = slot[(it.first * m) >> (32 - bits)];
                    if (sl >= 0) { ok = false; break; }
                    sl = it.second;
                }
                if (!ok) continue;
                idx.t1 = m; idx.tshift = 32 - bits;
                idx.tk_off = b.reserve(((size_t)1 << bits) / 2);
                uint16_t* tk = reinterpret_cast<uint16_t*>(&b.w[idx.tk_off]);
                for (size_t i = 0; i < slot.size(); i++) tk[i] = slot[i] >= 0 ? (uint16_t)slot[i] : (uint16_t)0x1FF;
                done = true;
            }
        }
        if (!done) return false;
    }
    b.align4();
    idx.utag_off = b.reserve(4 * all.size());
    bool any_chain = false;
    for (size_t t = 0; t < all.size(); t++) {
        const std::string& s = *all[t];
        DcbUTag u;
        pack64(s, 0, s.size(), u.bits_lo, u.bits_hi);
        const uint64_t m = (1ull << (2 * s.size())) - 1ull;
        u.mask_lo = (uint32_t)m;
        u.mask_hi_len = (uint32_t)(m >> 32) | ((uint32_t)s.size() << 24);
        std::memcpy(&b.w[idx.utag_off + 4 * t], &u, sizeof(u));
        if (next_same[t] != 0x1FF) any_chain = true;
    }
    // chains of tags sharing an lmin-prefix (rare): one 16-bit successor per tag, 0x1FF = none
    idx.chain_off = 0;
    if (any_chain) {
        idx.chain_off = b.reserve((all.size() + 1) / 2);
        uint16_t* ch = reinterpret_cast<uint16_t*>(&b.w[idx.chain_off]);
        for (size_t t = 0; t < all.size(); t++) ch[t] = (uint16_t)next_same[t];
    }
    b.align4();
    idx.head_words = (int32_t)b.w.size();
    const size_t words = (size_t)1 << wbits;
    idx.bloom_off = b.reserve(words);
    for (auto& kv : seeds)
        b.w[idx.bloom_off + DCB_BLOOM_WORD(kv.first, idx.bmul, wbits)] |= 1u << DCB_BLOOM_BIT(kv.first);
    b.align4();
    idx.legacy_words = (int32_t)b.w.size();
    {   // flat kernel: qq-mer -> offset set by hash-and-displace, and the byte filter
        idx.qq = DCB_QQ(lmin, idx.q);
        idx.qstride = lmin - idx.qq + 1;
        if (idx.qstride > 12) { idx.qq = idx.q; idx.qstride = idx.stride; }
        std::map<uint32_t, uint32_t> seeds2;
        for (size_t t = 0; t < all.size(); t++) {
            uint32_t lo, hi;
            for (int o = 0; o < idx.qstride; o++) {
                if (!pack64(*all[t], o, idx.qq, lo, hi)) return false;
                seeds2[lo] |= 1u << o;
            }
        }
        std::vector<std::pair<uint32_t, uint32_t>> items(seeds2.begin(), seeds2.end());
        Chd chd;
        if (!build_chd(chd, items, idx.qq)) return false;
        idx.m1 = chd.m1; idx.m2 = chd.m2; idx.b1 = chd.b1; idx.b2 = chd.b2;
        const size_t n1 = (size_t)1 << chd.b1, n2 = (size_t)1 << chd.b2;
        idx.qtab_off = b.reserve((n1 + n2 + 1) / 2);
        uint16_t* t16 = reinterpret_cast<uint16_t*>(&b.w[idx.qtab_off]);
        for (size_t i = 0; i < n1; i++) t16[i] = chd.disp[i];
        for (size_t i = 0; i < n2; i++) t16[n1 + i] = chd.slot[i];
        b.align4();
        // tag slots: perfect hash over the lmin-prefixes straight to {prefix_lo, meta} records
        idx.tq_off = 0; idx.tq_bits = 0; idx.ta = idx.tb = 1;
        if (lmin >= 17 && lmin <= 23 && all.size() <= 255) {
            const uint32_t himask = (1u << DCB_TQ_HIBITS(lmin)) - 1u;
            uint64_t rng2 = 0xA0761D6478BD642Full;
            bool done = false;
            for (int bits = 8; bits <= 12 && !done; bits++) {
                if (((size_t)1 << bits) < 4 * prefixes.size()) continue;
                std::vector<int> slot((size_t)1 << bits);
                for (int attempt = 0; attempt < 4000 && !done; attempt++) {
                    rng2 ^= rng2 << 13; rng2 ^= rng2 >> 7; rng2 ^= rng2 << 17;
                    const uint32_t ta = (uint32_t)rng2 | 1u, tb = (uint32_t)(rng2 >> 32) | 1u;
                    std::fill(slot.begin(), slot.end(), -1);
                    bool ok = true;
                    for (auto& kv : prefixes) {
                        int& sl = slot[(kv.first.first * ta + (kv.first.second & himask) * tb) >> (32 - bits)];
                        if (sl >= 0) { ok = false; break; }
                        sl = kv.second;
                    }
                    if (!ok) continue;
                    idx.ta = ta; idx.tb = tb; idx.tq_bits = bits;
                    idx.tq_off = b.reserve((size_t)2 << bits);
                    for (size_t i = 0; i < slot.size(); i++) {
                        uint32_t lo = 0, meta = DCB_TQ_FREE(lmin);
                        if (slot[i] >= 0) {
                            const std::string& t = *all[slot[i]];
                            uint32_t hi;
                            pack64(t, 0, lmin, lo, hi);
                            meta = DCB_TQ_META(hi & himask, slot[i], t.size(), (int)t.size() > lmin || next_same[slot[i]] != 0x1FF);
                        }
                        b.w[idx.tq_off + 2 * i] = lo; b.w[idx.tq_off + 2 * i + 1] = meta;
                    }
                    done = true;
                }
            }
        }
        b.align4();
        idx.qwlead = idx.qstride - 1;
        idx.qtab_words = (int32_t)b.w.size() - idx.qtab_off;
        idx.fbits = DCB_FBITS;
        idx.fmul = DCB_BLOOM_MUL(idx.qq);
        if (idx.fbits > 2 * idx.qq) idx.fbits = 2 * idx.qq;
        idx.bfilter_off = b.reserve(((size_t)1 << idx.fbits) / 4);
        uint8_t* f8 = reinterpret_cast<uint8_t*>(&b.w[idx.bfilter_off]);
        for (auto& kv : seeds2) f8[DCB_FSLOT(kv.first, idx.fmul, idx.fbits)] = 1;
    }
    b.align4();
    idx.n_words = (int32_t)b.w.size();
    std::memcpy(&b.w[0], &idx, sizeof(idx));
    out = std::move(b.w);
    return true;
}

// Sampled half-tag index over the four half keyword sets of a chain (see DcbHalfIndex).  false: the chain has a half
// keyword shorter than q + stride - 1 bases (the half-tag kernel then does not run for it).
bool dcb_build_half_index(const dcb_tagset* v, const dcb_tagset* j, std::vector<uint32_t>& out) {
    DcbHalfIndex hx;
    std::memset(&hx, 0, sizeof(hx));
    hx.q = DCB_HALF_Q; hx.stride = DCB_HALF_STRIDE; hx.kmin = hx.q + hx.stride - 1;
    hx.n_v = v->n_tags; hx.n_tags = v->n_tags + j->n_tags;
    hx.v_split = v->split; hx.j_split = j->split;
    struct Kw { std::string s; int set; std::vector<int> tags; };
    std::vector<Kw> kws;                                            // distinct keywords per set, in order of first use
    hx.j_ok = 1;
    for (int t = 0; t < j->n_tags; t++)
        if (j->split < hx.kmin || (int)j->tags[t].size() - j->split < hx.kmin) hx.j_ok = 0;
    // J halves below kmin (the 6-base halves of the 12-nt J tags) are not in the sampled index; when they have at least
    // DCB_HALF_JQ bases they get a table of their own, probed at EVERY base of the reads that need it (j_short)
    hx.j_short = hx.j_ok ? 0 : 1;
    for (int t = 0; t < j->n_tags; t++)
        if (j->split < DCB_HALF_JQ || (int)j->tags[t].size() - j->split < DCB_HALF_JQ) hx.j_short = 0;
    for (int gi = 0; gi < 2; gi++) {
        const dcb_tagset* ts = gi ? j : v;
        if (gi == 1 && !hx.j_ok && !hx.j_short) break;              // J halves too short: the V side only
        const int kmin = (gi == 1 && !hx.j_ok) ? DCB_HALF_JQ : hx.kmin;
        for (int half = 0; half < 2; half++) {
            std::map<std::string, size_t> seen;
            for (int t = 0; t < ts->n_tags; t++) {
                const std::string h = half ? ts->tags[t].substr(ts->split) : ts->tags[t].substr(0, ts->split);
                if ((int)h.size() < kmin || h.size() > 31 || ts->tags[t].size() > 31) return false;
                auto it = seen.find(h);
                if (it == seen.end()) { seen[h] = kws.size(); kws.push_back({h, 2 * gi + half, {t}}); }
                else kws[it->second].tags.push_back(t);             // ascending tag index
            }
        }
    }
    if (kws.empty() || kws.size() > 255) return false;
    Blob b;
    b.reserve((sizeof(DcbHalfIndex) + 3) / 4);
    b.align4();
    hx.t_off = b.reserve(((size_t)1 << (2 * hx.q)) / 2);
    uint16_t* t16 = reinterpret_cast<uint16_t*>(&b.w[hx.t_off]);
    std::map<uint32_t, std::vector<int>> by_prefix;                 // set << 28 | kmin-prefix -> keyword record numbers
    std::map<uint32_t, std::vector<int>> by_six;                    // j_short: first DCB_HALF_JQ bases -> J keyword record numbers (both halves)
    std::vector<uint8_t> tag_ids;
    std::vector<DcbHalfKw> recs(kws.size());
    for (size_t i = 0; i < kws.size(); i++) {
        const Kw& k = kws[i];
        uint32_t lo, hi;
        if (k.set >= 2 && hx.j_short) {
            if (!pack64(k.s, 0, DCB_HALF_JQ, lo, hi)) return false;
            by_six[lo].push_back((int)i);
        } else {
            for (int o = 0; o < hx.stride; o++) {
                if (!pack64(k.s, o, hx.q, lo, hi)) return false;
                t16[lo] |= (uint16_t)(1u << (4 * k.set + o));
            }
            pack64(k.s, 0, hx.kmin, lo, hi);
            by_prefix[((uint32_t)k.set << 28) | lo].push_back((int)i);
        }
        if (!pack64(k.s, 0, k.s.size(), lo, hi)) return false;
        DcbHalfKw& r = recs[i];
        std::memset(&r, 0, sizeof(r));
        pack64(k.s, 0, k.s.size(), r.bits_lo, r.bits_hi);
        const dcb_tagset* ts = k.set >= 2 ? j : v;
        r.len = (uint8_t)k.s.size();
        r.first_len = (uint8_t)ts->tags[k.tags[0]].size();          // len(seqs[halfN_seqs.index(keyword)])
        if (k.tags.size() > 15) return false;                       // the kernel's work items count them in four bits
        r.n_tags = (uint8_t)k.tags.size();
        r.set = (uint8_t)k.set;
        if (tag_ids.size() + k.tags.size() > 65535) return false;
        r.tags_off = (uint16_t)tag_ids.size();
        for (int t : k.tags) tag_ids.push_back((uint8_t)t);
    }
    std::vector<uint8_t> ids;
    std::vector<std::pair<uint32_t, uint32_t>> items;
    for (auto& kv : by_prefix) {
        if (kv.second.size() > 255 || ids.size() > 255) return false;
        // longest keyword first: the order acora reports keywords that END together in is irrelevant here (they start
        // together), the hit list is sorted by (end, length) afterwards
        items.emplace_back(kv.first, (uint32_t)ids.size() | ((uint32_t)kv.second.size() << 8));
        for (int i : kv.second) ids.push_back((uint8_t)i);
    }
    std::vector<uint16_t> jt;
    if (hx.j_short) {
        jt.assign((size_t)1 << (2 * DCB_HALF_JQ), 0);
        for (auto& kv : by_six) {
            if (kv.second.size() > 255 || ids.size() > 255) return false;
            jt[kv.first] = (uint16_t)(ids.size() | (kv.second.size() << 8));     // same layout as a prefix-table value
            for (int i : kv.second) ids.push_back((uint8_t)i);
        }
    }
    Cuckoo ck;
    if (!build_cuckoo(ck, items)) return false;
    hx.c1 = ck.c1; hx.c2 = ck.c2; hx.hshift = 32 - ck.bits;
    const size_t slots = (size_t)1 << ck.bits;
    b.align4();
    hx.h_off = b.reserve(2 * slots);
    for (size_t i = 0; i < slots; i++) {
        b.w[hx.h_off + 2 * i] = ck.used[i] ? ck.key[i] : DCB_HALF_FREE;
        b.w[hx.h_off + 2 * i + 1] = ck.used[i] ? ck.val[i] : 0u;
    }
    hx.ids_off = b.reserve((ids.size() + 3) / 4 + 1);
    std::memcpy(&b.w[hx.ids_off], ids.data(), ids.size());
    b.align4();
    hx.n_kw = (int32_t)recs.size();
    hx.kw_off = b.reserve(4 * recs.size());
    std::memcpy(&b.w[hx.kw_off], recs.data(), sizeof(DcbHalfKw) * recs.size());
    hx.tags_off = b.reserve((tag_ids.size() + 3) / 4 + 1);
    std::memcpy(&b.w[hx.tags_off], tag_ids.data(), tag_ids.size());
    if (hx.j_short) {
        hx.jt_off = b.reserve(jt.size() / 2);
        std::memcpy(&b.w[hx.jt_off], jt.data(), 2 * jt.size());
    }
    b.align4();
    hx.n_words = (int32_t)b.w.size();
    std::memcpy(&b.w[0], &hx, sizeof(hx));
    out = std::move(b.w);
    return true;
}

extern "C" {

dcb_tagset* dcb_tagset_build(const char* const* tags, const int32_t* jumps, const char* const* regions,
                             int n, int half_split, int is_v) {
    if (!tags || !jumps || !regions || n < 1) { dcb_set_error("dcb_tagset_build: null/empty input"); return nullptr; }
    if (n > DCB_MAX_TAGS) { dcb_set_error("dcb_tagset_build: more than 255 tags"); return nullptr; }
    if (half_split < 1 || half_split > 16) { dcb_set_error("dcb_tagset_build: half_split out of range"); return nullptr; }
    std::vector<std::string> full, h1, h2, reg;
    int lmin = 1 << 30;
    for (int i = 0; i < n; i++) {
        std::string t(tags[i]), r(regions[i]);
        if ((int)t.size() <= half_split || t.size() > DCB_MAX_TAG_LEN || t.size() < 4) {
            dcb_set_error("dcb_tagset_build: tag %d has unsupported length %zu (need split < len <= 32)", i, t.size());
            return nullptr;
        }
        for (char c : t) if (base_code(c) < 0) { dcb_set_error("dcb_tagset_build: tag %d is not pure ACGT", i); return nullptr; }
        for (char c : r) if (base_code(c) < 0) { dcb_set_error("dcb_tagset_build: region %d is not pure ACGT", i); return nullptr; }
        if (r.size() > 32000 || jumps[i] > 32000 || jumps[i] < -32000) { dcb_set_error("dcb_tagset_build: region/jump too large"); return nullptr; }
        full.push_back(t);
        h1.push_back(t.substr(0, half_split));
        h2.push_back(t.substr(half_split));
        reg.push_back(r);
        lmin = std::min(lmin, (int)t.size());
    }

    dcb_tagset* ts = new dcb_tagset();
    ts->n_tags = n; ts->split = half_split; ts->is_v = is_v; ts->lmin = lmin;
    ts->tags = full;

    for (int which = 0; which < 2; which++) {  // 0: general blob (tags + keyword sets + regions), 1: core blob (tags only)
        Blob b;
        b.reserve((sizeof(DcbGene) + 3) / 4);
        DcbGene g;
        std::memset(&g, 0, sizeof(g));
        g.n_tags = n; g.split = half_split; g.is_v = is_v; g.lmin = lmin;
        b.align4();                                   // tag records are read with 128-bit loads
        g.tag_off = b.reserve(DCB_TAG_WORDS * (size_t)n);
        std::vector<DcbTag> trec(n);
        for (int i = 0; i < n; i++) {
            DcbTag& t = trec[i];
            std::memset(&t, 0, sizeof(t));
            pack64(full[i], 0, full[i].size(), t.bits_lo, t.bits_hi);
            t.len = (uint8_t)full[i].size();
            {
                const uint64_t m = full[i].size() >= 32 ? ~0ull : ((1ull << (2 * full[i].size())) - 1ull);
                t.mask_lo = (uint32_t)m; t.mask_hi = (uint32_t)(m >> 32);
            }
            t.jump = (int16_t)jumps[i];
            t.region_len = (int16_t)reg[i].size();
            t.edge_ok = reg[i].size() >= 32;
            if (t.edge_ok) pack64(reg[i], is_v ? reg[i].size() - 32 : 0, 32, t.edge_lo, t.edge_hi);
            t.next_same_prefix = 0xFF;
            for (int k2 = i + 1; k2 < n; k2++)
                if (full[k2].compare(0, lmin, full[i], 0, lmin) == 0) { t.next_same_prefix = (uint8_t)k2; break; }
        }
        if (which == 0) {
            build_kwset(b, g.full, full);
            build_kwset(b, g.half1, h1);
            build_kwset(b, g.half2, h2);
            g.full.set_id = is_v ? 0 : 3; g.half1.set_id = is_v ? 1 : 4; g.half2.set_id = is_v ? 2 : 5;
            for (int i = 0; i < n; i++) {
                size_t nw = (reg[i].size() + 15) / 16;
                trec[i].region_off = b.reserve(nw + 1);
                for (size_t p = 0; p < reg[i].size(); p++)
                    b.w[trec[i].region_off + p / 16] |= (uint32_t)base_code(reg[i][p]) << (2 * (p % 16));
            }
        }
        b.align4();
        g.n_words = (int32_t)b.w.size();
        std::memcpy(&b.w[g.tag_off], trec.data(), sizeof(DcbTag) * n);
        std::memcpy(&b.w[0], &g, sizeof(g));
        (which == 0 ? ts->general : ts->core) = std::move(b.w);
    }
    if (!dcb_build_seed_index(is_v ? &ts->tags : nullptr, is_v ? nullptr : &ts->tags, lmin, DCB_WBITS_SINGLE, ts->index)) {
        dcb_set_error("dcb_tagset_build: seed index construction failed");
        delete ts;
        return nullptr;
    }
    return ts;
}

void dcb_tagset_free(dcb_tagset* ts) { delete ts; }

size_t dcb_tagset_table_bytes(const dcb_tagset* ts) {
    return ts ? 4 * (ts->core.size() + ts->index.size() + ts->general.size()) : 0;
}

int dcb_tagset_blob(const dcb_tagset* ts, int which, const uint32_t** words, size_t* n_words) {
    if (!ts || !words || !n_words || which < 0 || which > 2) return DCB_EINVAL;
    const std::vector<uint32_t>& v = which == 0 ? ts->general : which == 1 ? ts->core : ts->index;
    *words = v.data(); *n_words = v.size();
    return DCB_OK;
}

int dcb_tagset_suffix_filter(const dcb_tagset* v, const dcb_tagset* j, uint32_t* out, size_t cap, size_t* n_words) {
    if (!v || !j || !n_words) return DCB_EINVAL;
    std::vector<std::string> kws;
    int kq = 1 << 30;
    for (const dcb_tagset* ts : {v, j})
        for (const std::string& t : ts->tags) {
            kws.push_back(t);
            kws.push_back(t.substr(0, ts->split));
            kws.push_back(t.substr(ts->split));
        }
    for (auto& k : kws) kq = std::min(kq, (int)k.size());
    if (kq < 1) { dcb_set_error("dcb_tagset_suffix_filter: empty keyword"); return DCB_EUNSUPPORTED; }
    if (kq > 16) kq = 16;
    DcbSuffixFilter h;
    h.kq = kq;
    h.fbits = 2 * kq < 16 ? (2 * kq < 5 ? 5 : 2 * kq) : 16;
    // short keywords index the filter directly (exact membership), longer ones through a multiplicative hash; the
    // prober masks the window to the kq-mer's own 2 kq bits first
    h.fmul = 2 * kq <= h.fbits ? (1u << (32 - h.fbits)) : ((0x2C1B3C6Du | 1u) << (32 - 2 * kq));
    const size_t words = ((size_t)1 << h.fbits) / 32;
    h.n_words = (int32_t)(DCB_SFILTER_HEAD + words);
    std::vector<uint32_t> blob(((size_t)h.n_words + 3) & ~(size_t)3, 0u);
    std::memcpy(blob.data(), &h, sizeof(h));
    for (auto& k : kws) {
        uint32_t lo, hi;
        if (!pack64(k, k.size() - kq, kq, lo, hi)) { dcb_set_error("dcb_tagset_suffix_filter: non-ACGT keyword"); return DCB_EUNSUPPORTED; }
        const uint32_t sl = DCB_SFSLOT(lo, h.fmul, h.fbits);
        blob[DCB_SFILTER_HEAD + (sl >> 5)] |= 1u << (sl & 31);
    }
    *n_words = blob.size();
    if (out && cap >= blob.size()) std::memcpy(out, blob.data(), blob.size() * 4);
    return DCB_OK;
}

int dcb_tagset_half_index(const dcb_tagset* v, const dcb_tagset* j, uint32_t* out, size_t cap, size_t* n_words) {
    if (!v || !j || !n_words) return DCB_EINVAL;
    std::vector<uint32_t> u;
    if (!dcb_build_half_index(v, j, u)) { dcb_set_error("dcb_tagset_half_index: the chain has half tags shorter than %d bases", DCB_HALF_Q + DCB_HALF_STRIDE - 1); return DCB_EUNSUPPORTED; }
    *n_words = u.size();
    if (out && cap >= u.size()) std::memcpy(out, u.data(), u.size() * 4);
    return DCB_OK;
}

int dcb_tagset_union_index(const dcb_tagset* v, const dcb_tagset* j, uint32_t* out, size_t cap, size_t* n_words) {
    if (!v || !j || !n_words) return DCB_EINVAL;
    if (v->lmin != j->lmin) { dcb_set_error("dcb_tagset_union_index: V and J tags have different minimum lengths"); return DCB_EUNSUPPORTED; }
    std::vector<uint32_t> u;
    if (!dcb_build_seed_index(&v->tags, &j->tags, v->lmin, DCB_WBITS_UNION, u)) { dcb_set_error("dcb_tagset_union_index: construction failed"); return DCB_EUNSUPPORTED; }
    *n_words = u.size();
    if (out && cap >= u.size()) std::memcpy(out, u.data(), u.size() * 4);
    return DCB_OK;
}

}  // extern "C"
