// dcb_tables.h -- flattened tag tables shared by the host builder (tagset.cpp) and the kernels.
//
// One gene (V or J) becomes ONE blob of 32-bit words that a thread block stages in shared memory.
// It replaces the six acora automata per chain of the reference (decombine.py:722-746):
//
//   * three keyword sets (full tags, half1 = tag[:split], half2 = tag[split:]), each searchable at
//     every END position through a suffix-k-mer bitmap + a small open-addressing hash that lists the
//     distinct keywords ending in that k-mer (longest first = findall() order at equal end);
//   * a sampled-seed index for the exact-tag fast path (DcbSeedIndex below): only every stride-th
//     position of a read is probed;
//   * per-tag records (packed tag, length, jump, first-tag-with-same-half lengths, the last / first
//     32 germline bases for the bit-parallel deletion walk) and the 2-bit packed germline regions.
//
// All offsets are in 32-bit words from the start of the blob.
#ifndef DCB_TABLES_H
#define DCB_TABLES_H

#include <stdint.h>

#define DCB_MAX_TAG_LEN 32
#define DCB_MAX_TAGS 255
#define DCB_MAX_READ_LEN 4096
#define DCB_HASH_EMPTY 0xFFFFFFFFu

// One distinct keyword of a keyword set (16 bytes = 4 words).
struct DcbKw {
    uint32_t bits_lo, bits_hi;  // 2-bit packed keyword, base 0 in bits [0,2)
    uint8_t len;                // bases
    uint8_t first_tag;          // list.index(keyword): first tag whose full/half1/half2 equals it
    uint8_t n_tags;             // tags sharing the keyword (ascending in the tag list)
    uint8_t tags_off;           // offset of those tag ids in the set's byte list
    uint32_t pad;
};

// One keyword set == one acora automaton of the reference.
struct DcbKwSet {
    int32_t n_kw;        // distinct keywords
    int32_t kq;          // suffix key length in bases (min(min_len, 8))
    int32_t bitmap_off;  // 4^kq bits
    int32_t hash_off;    // hash_size slots: (key << 16) | (first_kw << 8) | count ; DCB_HASH_EMPTY
    int32_t hash_mask;   // hash_size - 1
    int32_t kw_off;      // DcbKw[n_kw] grouped by suffix key, longest first inside a group
    int32_t taglist_off; // bytes: tag ids
    int32_t min_len, max_len;
};

// Per-tag record (48 bytes = 12 words; the first 16 bytes are read with one 128-bit load).
#define DCB_TAG_WORDS 12
struct DcbTag {
    uint32_t bits_lo, bits_hi;   // packed full tag
    uint32_t mask_lo, mask_hi;   // 2*len low bits set
    uint32_t edge_lo, edge_hi;   // V: last 32 germline bases (region[m-32:m]); J: first 32 (region[0:32])
    int16_t jump;                // jump_to_end_v / jump_to_start_j
    int16_t region_len;          // m
    uint8_t len;                 // tag length
    uint8_t edge_ok;             // region_len >= 32
    uint8_t next_same_prefix;    // next tag of the same gene sharing this tag's lmin-prefix, 0xFF = none
    uint8_t pad8;
    int32_t region_off;          // packed region words (general blob)
    uint32_t pad[3];
};

struct DcbGene {
    int32_t n_tags, split, is_v;
    int32_t tag_off;             // DcbTag[n_tags]
    DcbKwSet full, half1, half2; // general blob only
    int32_t lmin;                // shortest full tag
    int32_t n_words;             // blob size
};

// Seed index of the exact-tag fast path (one per gene, or ONE for both genes of a chain when their seed
// geometry agrees).  Every occurrence of a full tag (length >= lmin) contains a q-mer that starts at a multiple
// of stride = lmin-q+1 at tag offset o <= lmin-q, so only every stride-th position is probed:
//   1. seed bitmap (4^q bits): is the q-mer at the sampled position p part of any tag at an offset <= lmin-q?
//   2. on a hit, the unknown offset o falls in class c = o / span (two classes); the k-mer at [p - c*span, +k)
//      then lies inside the tag for every o of that class, and a 2-choice cuckoo table (two slot reads, no
//      probing loop) maps (c, k-mer) to the set of offsets it occurs at -- almost always exactly one;
//   3. for each offset the tag would start at P = p - o: its lmin-prefix is looked up in a second cuckoo table
//      (all tags of the index, keyed on the folded prefix) and the whole tag is compared with the read.
struct DcbSeedIndex {
    int32_t q, stride;
    int32_t max_off;             // lmin - q: largest indexed tag offset
    int32_t k, span;             // class key length (<= 15 bases) and offsets per class
    int32_t wlead;               // the verification window starts at p - wlead (max_off + 1)
    int32_t ck_off;              // 2^bits slots of 2 words: [class << 31 | key  (DCB_HASH_EMPTY if free), mask of offsets]
    uint32_t c1, c2;             // h(x) = (x * c) >> shift
    int32_t shift;
    int32_t tk_off;              // 2^bits slots: fingerprint << 9 | gene << 8 | tag  (gene 0 = V, 1 = J) or DCB_HASH_EMPTY;
                                 // fingerprint = top 23 bits of the folded prefix
    uint32_t t1, t2;             // h(f) = (f * t) >> tshift, f = dcb_fold64(lmin-prefix)
    int32_t tshift;
    int32_t seedmap_off;         // the bitmap comes last
    int32_t n_words;
};

// Geometry of a seed index as a function of (lmin, q): shared by the host builder and the kernels, whose
// specialisations evaluate these at compile time.
#define DCB_IDX_STRIDE(lmin, q) ((lmin) - (q) + 1)
#define DCB_IDX_MAXOFF(lmin, q) ((lmin) - (q))
#define DCB_IDX_SPAN(lmin, q) ((DCB_IDX_MAXOFF(lmin, q) + 2) / 2)
#define DCB_IDX_WLEAD(lmin, q) (DCB_IDX_MAXOFF(lmin, q) + 1)
#define DCB_IDX_K0(lmin, q) (((lmin) - DCB_IDX_SPAN(lmin, q) + 1) < 15 ? ((lmin) - DCB_IDX_SPAN(lmin, q) + 1) : 15)
#define DCB_IDX_K(lmin, q) ((DCB_IDX_WLEAD(lmin, q) + DCB_IDX_K0(lmin, q)) > 32 ? (32 - DCB_IDX_WLEAD(lmin, q)) : DCB_IDX_K0(lmin, q))

// Seed bitmap addressing: the low 5 bits of a q-mer key select the bit -- stored MSB-first, so that
// `word << (key & 31)` moves it to bit 31, from where one funnel shift appends it to a hit mask -- and the
// remaining high bits select the word.
#define DCB_SEEDMAP_WORD(key, q) ((key) >> 5)
#define DCB_SEEDMAP_BIT(key, q) (31u - ((key) & 31u))

#if defined(__CUDACC__)
#define DCB_HD __host__ __device__ __forceinline__
#else
#define DCB_HD inline
#endif

DCB_HD uint32_t dcb_fold64(uint32_t lo, uint32_t hi) {
    uint32_t k = lo * 0x9E3779B1u + hi * 0x85EBCA77u;
    return k ^ (k >> 13);
}
#define DCB_TK_FP(f) ((f) & 0xFFFFFE00u)   // the fingerprint bits of a tag-prefix slot
DCB_HD uint32_t dcb_hash32(uint32_t k) {
    k *= 0x9E3779B1u;
    return k ^ (k >> 15);
}

#endif
