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
    int32_t kq;          // suffix key length in bases (min(min_len, 6))
    int32_t bitmap_off;  // 4^kq bits
    int32_t hash_off;    // hash_size slots: (key << 16) | (first_kw << 8) | count ; DCB_HASH_EMPTY
    int32_t hash_mask;   // hash_size - 1
    int32_t kw_off;      // DcbKw[n_kw] grouped by suffix key, longest first inside a group
    int32_t taglist_off; // bytes: tag ids
    int32_t min_len, max_len;
    int32_t set_id;      // 0 full, 1 first halves, 2 second halves; + 3 for the J gene (tags the entries of a hit list)
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
// of stride = max_off + 1 at a tag offset o <= max_off = min(lmin - q, 11), so only every stride-th position of
// a read is probed:
//   1. seed filter: a one-hash Bloom filter over the indexed q-mers (2^wbits words).  (Two bits per q-mer were
//      measured: 3 more instructions per probe bought 7 % fewer verification trips -- no net gain -- because most
//      extra hits are real q-mers of homologous tags, not filter noise.)  The kernels replicate it
//      32x in shared memory, one private copy per bank, so the 32 lanes of a warp -- each probing for its own
//      read -- never conflict: a probe is ONE shared-memory wavefront.
//   2. on a filter hit the unknown offset o falls in class c = o / span (two classes); the k-mer at
//      [p - c*span, +k) then lies inside the tag for every o of that class, and a 2-choice cuckoo table (two slot
//      reads per class, no probing loop) maps (c, k-mer) to the set of offsets it occurs at -- almost always one;
//   3. an offset o puts the tag start at P = p - o: the lmin-prefix at P is looked up in a perfect-hash table
//      (one slot read) that names the one tag with that prefix;
//   4. the whole tag (16-byte record: bits, mask, length) is compared with the read.
// Tags of both genes are numbered together ("ctag"): V tags first, then J tags.
struct DcbSeedIndex {
    int32_t q, stride;
    int32_t max_off;             // largest indexed tag offset
    int32_t wlead;               // the verification window starts at p - wlead (max_off + 1)
    int32_t lmin;                // shortest tag of the index
    int32_t wbits;               // log2(words) of the seed filter
    uint32_t bmul;               // word = (window * bmul) >> (32 - wbits); bmul = odd << (32 - 2q), so only the q-mer's
                                 // own 2q bits reach the product;  bit = 31 - (q-mer & 31)  (MSB-first)
    int32_t k, span;             // class key length (<= 15 bases) and offsets per class
    int32_t ck_off;              // 2^cbits slots: fingerprint << 12 | offset set.  For key x = class << 30 | k-mer the
    uint32_t c1, c2;             // two candidate slots are (x * c1) >> cshift and (x * c2) >> cshift; an entry sitting in
    int32_t cshift;              // its c1-slot carries the top 20 bits of x * c2 as fingerprint and vice versa (the slot
                                 // index already pins the top bits of its own product).  A free slot is 0 (empty offset
                                 // set); a lookup ORs the sets of BOTH slots whose fingerprint matches, so a chance
                                 // fingerprint match only adds offsets to try
    int32_t tk_off;              // 2^tbits 16-bit slots of a PERFECT hash over the lmin-prefixes: the first ctag with that
                                 // prefix, 0x1FF = free;  slot = (dcb_fold64(prefix) * t1) >> tshift.  No fingerprint: the
    uint32_t t1;                 // tag found is compared with the read as a whole anyway.
    int32_t tshift;
    int32_t utag_off;            // DcbUTag[n_tags]
    int32_t n_v, n_tags;         // ctag >= n_v is J tag ctag - n_v
    int32_t chain_off;           // 0, or uint16[n_tags]: next tag sharing this tag's lmin-prefix (0x1FF = none)
    int32_t head_words;          // words before the filter (what the specialised kernels stage verbatim)
    int32_t bloom_off;           // the bit filter follows the head
    int32_t n_words;
    // --- tables of the flat kernel (dcb_exact_kernel_flat), stored behind the bit filter -------------------------------
    //   * byte filter: ONE byte per slot (1 = an indexed seed hashes here), slot = DCB_FSLOT(seed).  A probe is IMAD (hash),
    //     LEA.HI (slot + the filter's fixed shared-memory address), LDS.U8 and one IMAD that appends the byte to the hit
    //     mask: no bit extraction, half of the probe's instructions on the FMA pipe;
    //   * q-mer -> offset set, hash-and-displace (CHD): bucket = (x * m1) >> (32 - b1), slot = (((x * m2) >> (32 - b2))
    //     + disp[bucket]) & (2^b2 - 1); the 16-bit slot holds the set of tag offsets the q-mer occurs at (bit o).  No
    //     fingerprint: every candidate is compared with the read as a whole, a filter false positive just finds nothing.
    //     m1 and m2 are odd << (32 - 2q), so only the q-mer's own bits of a wider window reach the products.
    //     The flat kernel's tables have their OWN seed geometry (qq, qstride): longer seeds sampled more densely
    //     (20-nt tags: 13-mers at every 8th position) hit fewer homologous tags, which is what its trip count pays for.
    int32_t legacy_words;        // words up to the end of the bit filter (what the other exact kernels stage)
    int32_t qq, qstride;         // seed length / sampling stride of the tables below (qstride = lmin - qq + 1 = their wlead)
    int32_t fbits;
    uint32_t fmul;
    uint32_t m1, m2;
    int32_t b1, b2;
    int32_t qtab_off;            // uint16 disp[2^b1] then uint16 offsets[2^b2]
    int32_t qtab_words;
    int32_t bfilter_off;         // 2^fbits bytes
    //   * tag slots: a PERFECT hash over the lmin-prefixes straight to 8-byte records (inside the qtab block):
    //     slot = (prefix_lo * ta + prefix_hi * tb) >> (32 - tq_bits); a record is {prefix_lo, DCB_TQ_META(...)} so a 20-nt
    //     tag is confirmed by ONE 64-bit read and two compares.  DCB_TQ_MORE marks tags longer than lmin or sharing their
    //     prefix with another tag: those go on to the whole-tag compare / chain walk of the other kernels.  A free slot
    //     has length 255 (never fits a read).
    int32_t tq_off, tq_bits;
    uint32_t ta, tb;
    int32_t qwlead;              // the flat kernel's verification window starts at p - qwlead (= qstride - 1)
};
// meta word of a tag slot: prefix bits 32.. (2 * lmin - 32 <= 14 of them) and one guard bit above them that only a free
// slot sets (so a free slot equals no read) | ctag << 16 (8 bits) | DCB_TQ_MORE | length << 25
#define DCB_TQ_MORE (1u << 24)
#define DCB_TQ_META(prefix_hi, ctag, len, more) ((uint32_t)(prefix_hi) | ((uint32_t)(ctag) << 16) | ((more) ? DCB_TQ_MORE : 0u) | ((uint32_t)(len) << 25))
#define DCB_TQ_LEN(meta) ((meta) >> 25)
#if defined(__CUDA_ARCH__)
#define DCB_TQ_CTAG(meta) __byte_perm((meta), 0u, 0x4442)           // byte 2, one PRMT
#else
#define DCB_TQ_CTAG(meta) (((meta) >> 16) & 0xFFu)
#endif
#define DCB_TQ_HIBITS(lmin) (2 * (lmin) - 32)                       // prefix bits held in the meta word
#define DCB_TQ_CMPMASK(lmin) ((2u << DCB_TQ_HIBITS(lmin)) - 1u)     // those bits and the guard bit
#define DCB_TQ_FREE(lmin) DCB_TQ_META(1u << DCB_TQ_HIBITS(lmin), 0xFF, 127, 0)
#ifndef DCB_FBITS
#define DCB_FBITS 16             // byte filter of the flat kernel: 64 KB
#endif
#define DCB_FSLOT(key, fmul, fbits) (((uint32_t)(key) * (fmul)) >> (32 - (fbits)))
// seed length of the flat kernel's tables: stride 8 where the tags allow it (2 * qq <= 30 bits of key)
#define DCB_QQ(lmin, q) (((lmin) >= 20 && (lmin) <= 22) ? (lmin) - 7 : (q))

#define DCB_CK_FPMASK 0xFFFFF000u                               // fingerprint bits of a slot / of a product
#define DCB_CK_OFFMASK(e) ((e) & 0xFFFu)

// Compact tag record of the fast path (16 bytes, one 128-bit load).
struct alignas(16) DcbUTag {
    uint32_t bits_lo, bits_hi;   // packed tag
    uint32_t mask_lo;            // low word of the 2*len-bit mask
    uint32_t mask_hi_len;        // high word of the mask (24 bits: len <= 28) | len << 24
};
#define DCB_FAST_MAX_TAG_LEN 28
// filter sizes the library builds: 2^12 words when one index serves both genes, 2^11 words per gene otherwise
// (DCB_BLOOM_COPIES private copies of all filters of a chain must fit in shared memory: 128 KB either way)
#define DCB_WBITS_UNION 12
#define DCB_WBITS_SINGLE 11
// private copies of a filter in shared memory: lane l probes copy l % DCB_BLOOM_COPIES; word i of copy c sits at word
// i * DCB_BLOOM_COPIES + c, i.e. in bank 8 * (i % 4) + c, so lanes of different copies never conflict and the four
// lanes that share a copy conflict only when their words fall in the same quarter (measured in profiles/)
#define DCB_BLOOM_COPIES 8

// Geometry of a seed index as a function of (lmin, q): shared by the host builder and the kernels, whose
// specialisations evaluate these at compile time.
#define DCB_IDX_MAXOFF(lmin, q) (((lmin) - (q)) < 11 ? ((lmin) - (q)) : 11)
#define DCB_IDX_STRIDE(lmin, q) (DCB_IDX_MAXOFF(lmin, q) + 1)
#define DCB_IDX_WLEAD(lmin, q) (DCB_IDX_MAXOFF(lmin, q) + 1)
#define DCB_IDX_SPAN(lmin, q) ((DCB_IDX_MAXOFF(lmin, q) + 2) / 2)
#define DCB_IDX_K0(lmin, q) (((lmin) - DCB_IDX_SPAN(lmin, q) + 1) < 15 ? ((lmin) - DCB_IDX_SPAN(lmin, q) + 1) : 15)
#define DCB_IDX_K(lmin, q) ((DCB_IDX_WLEAD(lmin, q) + DCB_IDX_K0(lmin, q)) > 32 ? (32 - DCB_IDX_WLEAD(lmin, q)) : DCB_IDX_K0(lmin, q))

// Seed filter addressing: the low 5 bits of a q-mer select the bit -- stored MSB-first, so that
// `word << (q-mer & 31)` moves it to bit 31, from where one funnel shift appends it to a hit mask.
#define DCB_BLOOM_MUL(q) ((0x2C1B3C6Du | 1u) << (32 - 2 * (q)))
#define DCB_BLOOM_BIT(key) (31u - ((key) & 31u))
#define DCB_BLOOM_WORD(win, bmul, wbits) (((uint32_t)(win) * (bmul)) >> (32 - (wbits)))

// Union suffix filter of the general kernel: one bit per slot over the last `kq` bases of EVERY keyword of the six
// keyword sets of a chain (full tags, first halves, second halves of both genes), kq = the shortest keyword.  A keyword
// can only end where the kq-mer in front of that position is in the filter, so the general kernel marks those
// positions once per read (one probe per base) and its six findall scans then visit only the marked ones.
// slot = DCB_SFSLOT(kq-mer); fmul = odd << (32 - 2 kq) (or 1 << (32 - 2 kq) when 2 kq <= fbits: the kq-mer itself).
struct DcbSuffixFilter {
    int32_t kq, fbits;
    uint32_t fmul;
    int32_t n_words;             // header + 2^fbits / 32 filter words
};
#define DCB_SFILTER_HEAD 4       // words in front of the bits
#define DCB_SFSLOT(key, fmul, fbits) (((uint32_t)(key) * (fmul)) >> (32 - (fbits)))

// Sampled half-tag index of the half-tag kernel (dcb_halftag_kernel): finds every occurrence of a half keyword (the four
// sets V half1, V half2, J half1, J half2 = set 0..3) without walking an automaton over every base.  Every keyword has
// at least kmin = q + stride - 1 bases, so each occurrence contains the q-mer that starts at the first multiple of
// `stride` at or behind its start (keyword offset o < stride): only every stride-th position of a read is probed.
//   * probe table: DIRECT-indexed by the q-mer (4^q 16-bit entries, q = 7: 32 KB): bit 4 * set + o says that the q-mer
//     occurs at offset o of a keyword of that set;
//   * keyword table: 2-choice cuckoo over (set, kmin-prefix), 8-byte slots {key, position | count << 8}: the keywords
//     of the set that start with this prefix (almost always one), as a run of `ids`, each naming a DcbHalfKw record:
//     the whole keyword to compare with the read, the length of the FIRST tag that has this half (the reference's
//     length guard, decombine.py:302-307) and the tags that share it, ascending.
// Built for chains whose V half keywords all have >= 10 bases (every shipped set: the V split is 10); the J sets are
// left out (j_ok = 0) when a J half is shorter (the 6-base halves of the `original` J sets).
struct alignas(16) DcbHalfKw {
    uint32_t bits_lo, bits_hi;   // packed keyword
    uint8_t len, first_len, n_tags, set;
    uint16_t tags_off, pad;      // tag ids (inside the gene) in the tag-id list
};
struct DcbHalfIndex {
    int32_t q, stride, kmin;
    int32_t n_v, n_tags;
    int32_t v_split, j_split;
    int32_t t_off;               // uint16[4^q]
    int32_t h_off, hshift;       // 2^(32 - hshift) slots of {uint32 key, uint32 meta}; key = set << 28 | kmin-prefix; free = ~0
    uint32_t c1, c2;
    int32_t ids_off;             // uint8: keyword record numbers, grouped by (set, prefix)
    int32_t kw_off, n_kw;        // DcbHalfKw[n_kw]
    int32_t tags_off;            // uint8 tag ids
    int32_t n_words;
    int32_t j_ok;                // 0: the J half keywords are too short for the index (6-base halves of the `original` J sets): only the
                                 // V side is searched; a read whose V is assigned but whose J tag is missing goes on to the general kernel
    int32_t j_short;             // j_ok == 0 and every J half has at least DCB_HALF_JQ bases: jt_off is a direct table over the first
    int32_t jt_off;              // DCB_HALF_JQ bases of the J half keywords, uint16[4^JQ] of (first | count << 8) into ids; a read whose
                                 // V is assigned and whose J is missing is probed with it at every base
};
#define DCB_HALF_JQ 6
#define DCB_HALF_Q 7
#define DCB_HALF_STRIDE 4
#define DCB_HALF_FREE 0xFFFFFFFFu

#if defined(__CUDACC__)
#define DCB_HD __host__ __device__ __forceinline__
#else
#define DCB_HD inline
#endif

DCB_HD uint32_t dcb_fold64(uint32_t lo, uint32_t hi) {
    uint32_t k = lo * 0x9E3779B1u + hi * 0x85EBCA77u;
    return k ^ (k >> 13);
}
DCB_HD uint32_t dcb_hash32(uint32_t k) {
    k *= 0x9E3779B1u;
    return k ^ (k >> 15);
}

#endif
