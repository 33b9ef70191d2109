// lev_core.cuh -- bit-parallel Levenshtein distance, written once for host and device.
//
// Replaces the two native distance routines collapse reaches through third-party wheels
// (/root/reference/src/decombinator/collapse.py):
//   * polyleven.levenshtein(seq1, seq2)                       collapse.py:360, 364  (are_seqs_equivalent /
//     are_barcodes_equivalent)
//   * the Levenshtein.distance calls inside pyrepseq.nn.symdel collapse.py:735-740  (UMI neighbour search)
// Both are plain unit-cost edit distances; the algorithm here is Myers' bit-vector method in Hyyro's
// block formulation (one machine word per 32 / 64 pattern symbols, horizontal carries between words).
// Functions are __host__ __device__ so tests/sim can check them against a textbook DP without a GPU; the
// library exports no CPU compute path.
#ifndef DCB_LEV_CORE_CUH
#define DCB_LEV_CORE_CUH

#include <stdint.h>

#if defined(__CUDACC__)
#define LEV_HD __host__ __device__ __forceinline__
#else
#define LEV_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define LEV_UNROLL _Pragma("unroll")
#else
#define LEV_UNROLL
#endif

// ------------------------------------------------------------------------------------------------
// UMI codes: up to 19 symbols of 3 bits (symbol k in bits [3k, 3k+3)), length in bits [58, 64).
// ------------------------------------------------------------------------------------------------
#define DCB_UMI_MAX_LEN 19
LEV_HD int umi_len(uint64_t c) { return (int)(c >> 58); }
LEV_HD uint32_t umi_sym(uint64_t c, int k) { return (uint32_t)(c >> (3 * k)) & 7u; }

// Match masks of a UMI used as the Myers pattern: peq[s] has bit k set iff symbol k == s.
struct UmiPattern {
    uint32_t peq[8];
    int len;
};
LEV_HD void umi_pattern(uint64_t c, UmiPattern& p) {
    p.len = umi_len(c);
    LEV_UNROLL
    for (int s = 0; s < 8; s++) p.peq[s] = 0;
    for (int k = 0; k < p.len; k++) {
        const uint32_t s = umi_sym(c, k);
        LEV_UNROLL
        for (int t = 0; t < 8; t++) p.peq[t] |= (t == (int)s ? 1u : 0u) << k;   // no dynamic register indexing
    }
}

// Levenshtein(pattern, text) for a pattern of at most 32 symbols (one word).
LEV_HD int umi_distance(const UmiPattern& p, uint64_t text) {
    const int n = umi_len(text), m = p.len;
    if (m == 0) return n;
    uint32_t vp = m >= 32 ? 0xFFFFFFFFu : ((1u << m) - 1u), vn = 0;
    const uint32_t last = 1u << (m - 1);
    int score = m;
    for (int j = 0; j < n; j++) {
        const uint32_t s = umi_sym(text, j);
        uint32_t eq = 0;
        LEV_UNROLL
        for (int t = 0; t < 8; t++) eq |= (t == (int)s) ? p.peq[t] : 0u;
        const uint32_t x = eq | vn;
        const uint32_t d0 = (((x & vp) + vp) ^ vp) | x;
        uint32_t hp = vn | ~(d0 | vp);
        uint32_t hn = d0 & vp;
        score += (hp & last) ? 1 : 0;
        score -= (hn & last) ? 1 : 0;
        hp = (hp << 1) | 1u;
        hn <<= 1;
        vp = hn | ~(d0 | hp);
        vn = hp & d0;
    }
    return score;
}

// Pigeonhole prefilter: if Levenshtein(a, b) <= k then, cutting a into k+1 contiguous blocks, at least one block
// survives the edits verbatim and sits in b shifted by at most k positions.  Returns false only when the
// distance is certainly > k.
LEV_HD bool umi_may_be_within(uint64_t a, uint64_t b, int k) {
    const int m = umi_len(a), n = umi_len(b);
    const int diff = m > n ? m - n : n - m;
    if (diff > k) return false;
    if (m < k + 1) return true;          // blocks would be empty: no filtering power
    const uint64_t body_b = b & ((1ull << 58) - 1ull);
    int start = 0;
    for (int blk = 0; blk <= k; blk++) {
        const int len = (m - start) / (k + 1 - blk);          // remaining symbols spread evenly
        const uint64_t mask = (1ull << (3 * len)) - 1ull;
        const uint64_t want = (a >> (3 * start)) & mask;
        for (int d = -k; d <= k; d++) {
            const int pos = start + d;
            if (pos < 0 || pos + len > n) continue;
            if (((body_b >> (3 * pos)) & mask) == want) return true;
        }
        start += len;
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
// Sequences of small codes (one byte per symbol), pattern up to 64*W symbols.
// The pattern is kept as NP bit planes per word instead of one match mask per symbol, so that the whole
// state stays in registers: eq(c) = the positions whose NP code bits all agree with c.  NP = 3 for codes 0..7
// (the usual case: A C G T N and a few more), NP = 8 for arbitrary bytes (lower case, the whole IUPAC alphabet).
// ------------------------------------------------------------------------------------------------
template <int W, int NP = 3>
struct SeqPattern {
    uint64_t p[NP][W], valid[W];
    int len;
};

template <int W, int NP>
LEV_HD void seq_pattern(const uint8_t* s, int m, SeqPattern<W, NP>& p) {
    p.len = m;
    LEV_UNROLL
    for (int w = 0; w < W; w++) {
        uint64_t pl[NP], v = 0;
        LEV_UNROLL
        for (int b = 0; b < NP; b++) pl[b] = 0;
        for (int k = 0; k < 64; k++) {
            const int i = 64 * w + k;
            if (i < m) {
                const uint64_t x = s[i];
                LEV_UNROLL
                for (int b = 0; b < NP; b++) pl[b] |= ((x >> b) & 1ull) << k;
                v |= 1ull << k;
            }
        }
        LEV_UNROLL
        for (int b = 0; b < NP; b++) p.p[b][w] = pl[b];
        p.valid[w] = v;
    }
}

// Levenshtein(pattern, text[0:n])
template <int W, int NP>
LEV_HD int seq_distance(const SeqPattern<W, NP>& p, const uint8_t* text, int n) {
    const int m = p.len;
    if (m == 0) return n;
    uint64_t vp[W], vn[W];
    LEV_UNROLL
    for (int w = 0; w < W; w++) { vp[w] = ~0ull; vn[w] = 0ull; }
    const int lw = (m - 1) >> 6;
    const uint64_t last = 1ull << ((m - 1) & 63);
    int score = m;
    for (int j = 0; j < n; j++) {
        const uint64_t x = text[j];
        uint64_t mk[NP];
        LEV_UNROLL
        for (int b = 0; b < NP; b++) mk[b] = ((x >> b) & 1ull) ? ~0ull : 0ull;
        uint64_t hp_carry = 1ull, hn_carry = 0ull;
    LEV_UNROLL
        for (int w = 0; w < W; w++) {
            if (w > lw) break;
            uint64_t eq = p.valid[w];
            LEV_UNROLL
            for (int b = 0; b < NP; b++) eq &= ~(p.p[b][w] ^ mk[b]);
            const uint64_t xx = eq | hn_carry;
            const uint64_t d0 = ((((xx | vn[w]) & vp[w]) + vp[w]) ^ vp[w]) | xx | vn[w];
            uint64_t hp = vn[w] | ~(d0 | vp[w]);
            uint64_t hn = d0 & vp[w];
            if (w == lw) {
                score += (hp & last) ? 1 : 0;
                score -= (hn & last) ? 1 : 0;
            }
            const uint64_t hp_out = hp >> 63, hn_out = hn >> 63;
            hp = (hp << 1) | hp_carry;
            hn = (hn << 1) | hn_carry;
            hp_carry = hp_out; hn_carry = hn_out;
            vp[w] = hn | ~(d0 | hp);
            vn[w] = hp & d0;
        }
    }
    return score;
}

#endif  // DCB_LEV_CORE_CUH
