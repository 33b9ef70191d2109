// synth.cpp -- deterministic synthetic TCR read generator (host side).
//
// Produces the workloads SURVEY.md section 8(d) defines for BASELINE.json's configs:
// a molecule is  V_region[:len-vdel] + insert + J_region[jdel:] + C  (C = fixed pseudo
// constant region), R1 is the reverse complement of the L nt ending 20..60 nt into C
// (so the reference's default `-or reverse` finds it), R2 carries the M13 barcode layout
// of collapse.py:176 (spacer + N6 + spacer + N6 + filler).  Every read is a pure function
// of (seed, read index): counter-based splitmix64, so any shard of the stream can be
// regenerated anywhere (host threads here; the same arithmetic is cheap to restate on device).
//
// This is workload generation, not part of the matching path.
#include "dcb.h"

#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    inline uint64_t next() {
        s += 0x9E3779B97F4A7C15ull;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    inline uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
};

const char kBases[4] = {'A', 'C', 'G', 'T'};
const int kVdel[10] = {0, 0, 1, 2, 3, 4, 5, 6, 8, 10};
const int kJdel[9] = {0, 0, 1, 2, 3, 4, 5, 6, 8};
const int kIns[11] = {0, 1, 2, 3, 4, 5, 6, 8, 10, 12, 15};
const char kSpacer1[] = "GTCGTGACTGGGAAAACCCTGG";
const char kSpacer2[] = "GTCGTGAT";

inline char comp(char c) {
    switch (c) {
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
        default: return c;
    }
}

struct GeneSet { std::vector<std::string> v, j; };

struct Synth {
    dcb_synth_params p;
    std::vector<GeneSet> sets;
    std::string constant;  // pseudo constant region, fixed per seed
};

void make_read(const Synth& S, uint64_t index, char* r1, char* r2) {
    const dcb_synth_params& p = S.p;
    Rng rng(p.seed ^ (index * 0x9E3779B97F4A7C15ull));
    // with a UMI pool, read `index` is a copy of molecule index % umi_pool: the rearrangement comes from the molecule's
    // own stream, sequencing errors from the read's
    Rng pool_rng((p.seed + 0x5DEECE66Dull) ^ ((index % (p.umi_pool ? p.umi_pool : 1)) * 0xC2B2AE3D27D4EB4Full));
    Rng& mrng = p.umi_pool ? pool_rng : rng;
    const int L = (int)p.read_len;
    bool junk = p.junk_rate && (uint32_t)(mrng.next() >> 32) < p.junk_rate;
    std::string mol;
    int e = 0;
    if (!junk) {
        const GeneSet& g = S.sets[index % S.sets.size()];
        const std::string& V = g.v[mrng.below((uint32_t)g.v.size())];
        const std::string& J = g.j[mrng.below((uint32_t)g.j.size())];
        int vdel = kVdel[mrng.below(10)], jdel = kJdel[mrng.below(9)], nins = kIns[mrng.below(11)];
        if (vdel > (int)V.size()) vdel = (int)V.size();
        if (jdel > (int)J.size()) jdel = (int)J.size();
        mol.reserve(V.size() + J.size() + 512);
        mol.append(V, 0, V.size() - vdel);
        for (int i = 0; i < nins; i++) mol.push_back(kBases[mrng.below(4)]);
        mol.append(J, jdel, std::string::npos);
        e = (int)mol.size() + 20 + (int)mrng.below(41);
        mol.append(S.constant);
    }
    // window = mol[e-L : e], left-padded with random bases when the molecule is short
    for (int i = 0; i < L; i++) {
        int src = e - 1 - i;  // reverse complement on the fly: r1[i] = comp(window[L-1-i])
        char c;
        if (junk || src < 0) c = kBases[rng.below(4)];
        else c = comp(mol[src]);
        r1[i] = c;
    }
    if (p.sub_rate || p.n_rate) {
        for (int i = 0; i < L; i++) {
            uint64_t r = rng.next();
            uint32_t a = (uint32_t)(r >> 32), b = (uint32_t)r;
            if (p.sub_rate && a < p.sub_rate) {
                int cur = r1[i] == 'A' ? 0 : r1[i] == 'C' ? 1 : r1[i] == 'G' ? 2 : 3;
                r1[i] = kBases[(cur + 1 + (b % 3)) & 3];
            }
            if (p.n_rate && (uint32_t)(b * 2654435761u) < p.n_rate) r1[i] = 'N';
        }
    }
    if (r2) {
        const int L2 = (int)p.read2_len;
        int k = 0;
        auto put = [&](char c) { if (k < L2) r2[k++] = c; };
        // UMI: either per-read random, or drawn from a pool of `umi_pool` molecules (copies share a UMI)
        Rng urng = p.umi_pool ? Rng(p.seed * 0xD1342543DE82EF95ull + (index % p.umi_pool)) : rng;
        for (const char* s = kSpacer1; *s; s++) put(*s);
        for (int i = 0; i < 6; i++) put(kBases[urng.below(4)]);
        for (const char* s = kSpacer2; *s; s++) put(*s);
        for (int i = 0; i < 6; i++) put(kBases[urng.below(4)]);
        while (k < L2) put(kBases[rng.below(4)]);
        if (p.sub_rate2) {   // sequencing errors in the barcode read (fuzzy spacers, UMI neighbours)
            for (int i = 0; i < L2; i++) {
                uint64_t r = rng.next();
                if ((uint32_t)(r >> 32) < p.sub_rate2) {
                    int cur = r2[i] == 'A' ? 0 : r2[i] == 'C' ? 1 : r2[i] == 'G' ? 2 : 3;
                    r2[i] = kBases[(cur + 1 + ((uint32_t)r % 3)) & 3];
                }
            }
        }
    }
}

}  // namespace

extern "C" {

dcb_synth* dcb_synth_create(const dcb_synth_params* p, int n_sets, const char* const* const* v_regions,
                            const int* n_v, const char* const* const* j_regions, const int* n_j) {
    if (!p || n_sets < 1 || p->read_len == 0) return nullptr;
    Synth* S = new Synth();
    S->p = *p;
    for (int s = 0; s < n_sets; s++) {
        GeneSet g;
        for (int i = 0; i < n_v[s]; i++) g.v.emplace_back(v_regions[s][i]);
        for (int i = 0; i < n_j[s]; i++) g.j.emplace_back(j_regions[s][i]);
        if (g.v.empty() || g.j.empty()) { delete S; return nullptr; }
        S->sets.push_back(std::move(g));
    }
    Rng crng(p->seed ^ 0xC0FFEE123456789ull);
    for (int i = 0; i < 400; i++) S->constant.push_back(kBases[crng.below(4)]);
    return reinterpret_cast<dcb_synth*>(S);
}

void dcb_synth_destroy(dcb_synth* h) { delete reinterpret_cast<Synth*>(h); }

int dcb_synth_reads(const dcb_synth* h, uint64_t first_index, uint64_t n, char* r1, char* r2, int n_threads) {
    const Synth* S = reinterpret_cast<const Synth*>(h);
    if (!S || !r1) return -1;
    if (n_threads < 1) n_threads = 1;
    const size_t L = S->p.read_len, L2 = S->p.read2_len;
    auto work = [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; i++) make_read(*S, first_index + i, r1 + i * L, r2 ? r2 + i * L2 : nullptr);
    };
    if (n_threads == 1 || n < 1024) { work(0, n); return 0; }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(work, n * t / n_threads, n * (t + 1) / n_threads);
    for (auto& t : th) t.join();
    return 0;
}

}  // extern "C"
