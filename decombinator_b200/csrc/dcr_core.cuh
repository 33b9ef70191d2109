// dcr_core.cuh -- the per-read decombine logic, written once for the device.
//
// Every function is __host__ __device__ so that tests/sim can compile the SAME code with g++ and
// check its logic against the oracle on a machine without a GPU.  The host build is test
// scaffolding only: libdcb.so exports no CPU compute path.
//
// Behavioural spec: /root/reference/src/decombinator/decombine.py (cited per function as :line).
// The matching itself is NOT the reference's algorithm (six byte-wise Aho-Corasick scans per read);
// see dcb_tables.h for the table design.  What is preserved is the observable contract: the
// findall() hit order, the candidate / guard / Hamming / deletion-walk sequence with Python's slice
// semantics, every counter increment, and the seven values dcr() returns.
#ifndef DCR_CORE_CUH
#define DCR_CORE_CUH

#include "dcb_tables.h"
#include "../../include/dcb.h"

#if defined(__CUDA_ARCH__)
#define DCB_COUNT(C, id) atomicAdd(&(C)[id], 1u)
#define DCB_POPC(x) __popc(x)
#define DCB_FUNNEL_R(lo, hi, sh) __funnelshift_r((lo), (hi), (sh))
#define DCB_BREV(x) __brev(x)
#define DCB_CLZ(x) __clz(x)
#define DCB_FFS(x) __ffs(x)
#else
#define DCB_COUNT(C, id) ((C)[id] += 1u)
#define DCB_POPC(x) __builtin_popcount(x)
static inline uint32_t dcb_funnel_r_host(uint32_t lo, uint32_t hi, int sh) {
    sh &= 31;
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}
static inline uint32_t dcb_brev_host(uint32_t x) {
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
}
#define DCB_FUNNEL_R(lo, hi, sh) dcb_funnel_r_host((lo), (hi), (sh))
#define DCB_BREV(x) dcb_brev_host(x)
#define DCB_CLZ(x) ((x) ? __builtin_clz(x) : 32)
#define DCB_FFS(x) __builtin_ffs(x)
#endif

typedef uint32_t dcb_cnt_t;  // per-block (device) or per-call (sim) counter slots

// ------------------------------------------------------------------------------------------------
// A read as the kernels see it: 2-bit words, strided (shared memory is laid out [word][thread] so a
// dynamic word index never causes a bank conflict), plus the sparse non-ACGT information.
// ------------------------------------------------------------------------------------------------
struct ReadView {
    const uint32_t* w;      // word i at w[i * stride]
    const uint32_t* inv;    // invalid-base bitmask (bit b of word i <-> base 32 i + b), same stride; null = none
    int stride;
    int n;                  // bases
    int nw;                 // words available (reads past them yield 0)
    // exception list of this read in the frame being analysed (positions already mapped to the frame)
    const uint16_t* exc_pos;
    const uint8_t* exc_kind;
    int e0, e1;
    int mirror;             // frame 1: position p of the frame is exception position n-1-p
    // candidate keyword positions (general kernel): bit b of word i <-> the cand_kq-mer that STARTS at base 32 i + b may
    // be the tail of a keyword (union suffix filter); same stride; null = not marked, scans visit every position
    const uint32_t* cand;
    int cand_kq;
    // hit list (general kernel): every keyword occurrence of all six sets, found in ONE pass over the marked positions
    // (hits_build) in findall order; entry = start | keyword << 16 | set_id << 24, same stride; null = not built, each
    // findall scans by itself
    const uint32_t* hits;
    int n_hits;
};
#define DCB_HITS_CAP 16

DCB_HD uint32_t rd_word(const ReadView& r, int i) {
    return ((unsigned)i < (unsigned)r.nw) ? r.w[i * r.stride] : 0u;
}
// 16 bases starting at base p (p may be negative or run past the end: missing bases read as 0)
DCB_HD uint32_t rd_win16(const ReadView& r, int p) {
    int wi = p >> 4;
    return DCB_FUNNEL_R(rd_word(r, wi), rd_word(r, wi + 1), (p & 15) * 2);
}
DCB_HD void rd_win32(const ReadView& r, int p, uint32_t& lo, uint32_t& hi) {
    int wi = p >> 4, sh = (p & 15) * 2;
    uint32_t a = rd_word(r, wi), b = rd_word(r, wi + 1), c = rd_word(r, wi + 2);
    lo = DCB_FUNNEL_R(a, b, sh);
    hi = DCB_FUNNEL_R(b, c, sh);
}
DCB_HD bool rd_inv_at(const ReadView& r, int p) {
    return r.inv && ((r.inv[(p >> 5) * r.stride] >> (p & 31)) & 1u);
}
// any invalid base in [a, b), 0 <= a, b <= n
DCB_HD bool rd_inv_any(const ReadView& r, int a, int b) {
    if (!r.inv || a >= b) return false;
    if (b - a <= 32) {                        // one funnel shift over two mask words (every caller but the N filter)
        const int wi = a >> 5, nwm = (r.nw + 1) / 2;
        const uint32_t w0 = r.inv[wi * r.stride], w1 = wi + 1 < nwm ? r.inv[(wi + 1) * r.stride] : 0u;
        const uint32_t x = DCB_FUNNEL_R(w0, w1, a & 31);
        return (x & (b - a == 32 ? 0xFFFFFFFFu : ((1u << (b - a)) - 1u))) != 0u;
    }
    for (int wi = a >> 5; wi <= ((b - 1) >> 5); wi++) {
        int lo = a > wi * 32 ? a - wi * 32 : 0;
        int hi = b < wi * 32 + 32 ? b - wi * 32 : 32;
        uint32_t mask = (hi - lo == 32) ? 0xFFFFFFFFu : (((1u << (hi - lo)) - 1u) << lo);
        if (r.inv[wi * r.stride] & mask) return true;
    }
    return false;
}
// "N" in read[a:b]  (only exceptions of kind 1 are the letter N)
DCB_HD bool rd_has_N(const ReadView& r, int a, int b) {
    for (int e = r.e0; e < r.e1; e++) {
        if (r.exc_kind[e] != 1) continue;
        int p = r.mirror ? r.n - 1 - (int)r.exc_pos[e] : (int)r.exc_pos[e];
        if (p >= a && p < b) return true;
    }
    return false;
}

DCB_HD uint32_t mask2(int nbases) {  // low 2*nbases bits, nbases in [0,16]
    return nbases >= 16 ? 0xFFFFFFFFu : ((1u << (2 * nbases)) - 1u);
}

// read[s:s+L] == keyword (0 <= s, s+L <= n, L <= 32); invalid bases never match
DCB_HD bool rd_equals(const ReadView& r, int s, int L, uint32_t klo, uint32_t khi) {
    uint32_t lo, hi;
    rd_win32(r, s, lo, hi);
    uint32_t mlo = mask2(L), mhi = L > 16 ? mask2(L - 16) : 0u;
    if (((lo ^ klo) & mlo) | ((hi ^ khi) & mhi)) return false;
    return !rd_inv_any(r, s, s + L);
}

// lev.hamming(tag, read[s:s+L]) <= 1   (decombine.py:309, 359, 436, 493)
DCB_HD bool rd_hamming_le1(const ReadView& r, int s, int L, uint32_t klo, uint32_t khi) {
    uint32_t lo, hi;
    rd_win32(r, s, lo, hi);
    uint32_t xlo = (lo ^ klo) & mask2(L), xhi = L > 16 ? ((hi ^ khi) & mask2(L - 16)) : 0u;
    uint32_t nlo = (xlo | (xlo >> 1)) & 0x55555555u, nhi = (xhi | (xhi >> 1)) & 0x55555555u;
    if (!r.inv || !rd_inv_any(r, s, s + L)) return DCB_POPC(nlo) + DCB_POPC(nhi) <= 1;
    int d = 0;  // rare: the window holds non-ACGT symbols, each one is a mismatch
    for (int i = 0; i < L; i++) {
        uint32_t nq = i < 16 ? (nlo >> (2 * i)) & 1u : (nhi >> (2 * (i - 16))) & 1u;
        d += (rd_inv_at(r, s + i) || nq) ? 1 : 0;
    }
    return d <= 1;
}

// ------------------------------------------------------------------------------------------------
// Candidate keyword positions: one probe of the union suffix filter (DcbSuffixFilter) per base of the read, 32 start
// positions per output word.  Non-ACGT symbols are packed as base 0 and probed as such: a keyword needs valid bases, so
// marking by the packed bits can only mark too much, never too little.  cand_col: this read's words, stride r.stride.
// ------------------------------------------------------------------------------------------------
DCB_HD void cand_build(ReadView& r, const uint32_t* sf, uint32_t* cand_col) {
    const DcbSuffixFilter& h = *reinterpret_cast<const DcbSuffixFilter*>(sf);
    const uint32_t* bits = sf + DCB_SFILTER_HEAD;
    const uint32_t kmask = h.kq >= 16 ? 0xFFFFFFFFu : ((1u << (2 * h.kq)) - 1u);
    const int sh = 32 - h.fbits;
    const int nout = (r.nw + 1) / 2;
    for (int m = 0; m < nout; m++) {
        uint32_t acc = 0;
        for (int half = 0; half < 2; half++) {
            const int wi = 2 * m + half;
            const uint32_t a = rd_word(r, wi), b = rd_word(r, wi + 1);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int j = 0; j < 16; j++) {
                const uint32_t key = DCB_FUNNEL_R(a, b, 2 * j) & kmask;
                const uint32_t sl = (key * h.fmul) >> sh;
                acc |= ((bits[sl >> 5] >> (sl & 31)) & 1u) << (16 * half + j);
            }
        }
        cand_col[m * r.stride] = acc;
    }
    r.cand = cand_col;
    r.cand_kq = h.kq;
}
// Smallest marked start position >= p, or a value >= 32 * words when there is none.
DCB_HD int cand_next(const ReadView& r, int p) {
    const int nout = (r.nw + 1) / 2;
    if (p < 0) p = 0;
    int m = p >> 5;
    if (m >= nout) return 32 * nout;
    uint32_t wv = r.cand[m * r.stride] & (0xFFFFFFFFu << (p & 31));
    while (!wv) {
        if (++m >= nout) return 32 * nout;
        wv = r.cand[m * r.stride];
    }
    return 32 * m + DCB_FFS(wv) - 1;
}

// ------------------------------------------------------------------------------------------------
// Germline regions (2-bit packed in the blob)
// ------------------------------------------------------------------------------------------------
DCB_HD uint32_t rg_win16(const uint32_t* blob, const DcbTag& t, int p) {
    int nw = (t.region_len + 15) >> 4;
    int wi = p >> 4;
    uint32_t a = ((unsigned)wi < (unsigned)nw) ? blob[t.region_off + wi] : 0u;
    uint32_t b = ((unsigned)(wi + 1) < (unsigned)nw) ? blob[t.region_off + wi + 1] : 0u;
    return DCB_FUNNEL_R(a, b, (p & 15) * 2);
}

// Python s[start:stop] bounds on a sequence of length len
struct Span { int a, b; };
DCB_HD Span py_slice(int len, int start, int stop) {
    if (start < 0) { start += len; if (start < 0) start = 0; } else if (start > len) start = len;
    if (stop < 0) { stop += len; if (stop < 0) stop = 0; } else if (stop > len) stop = len;
    if (stop < start) stop = start;
    Span s; s.a = start; s.b = stop;
    return s;
}

// region[a] == read[b] as Python strings (both at most 10 long here)
DCB_HD bool slices_equal(const uint32_t* blob, const DcbTag& t, Span a, const ReadView& r, Span b) {
    int la = a.b - a.a;
    if (la != b.b - b.a) return false;
    if (la == 0) return true;
    uint32_t x = (rg_win16(blob, t, a.a) ^ rd_win16(r, b.a)) & mask2(la);
    if (x) return false;
    return !rd_inv_any(r, b.a, b.b);
}

DCB_HD const DcbTag& gene_tag(const uint32_t* blob, const DcbGene& g, int k) {
    return reinterpret_cast<const DcbTag*>(blob + g.tag_off)[k];
}

// Bit-parallel "10 consecutive equal bases": x = xor of two 32-base windows; returns a word pair where
// bit 2i is set iff bases i..i+9 are all equal (i <= 22).
DCB_HD void run10(uint32_t xlo, uint32_t xhi, uint32_t& rlo, uint32_t& rhi) {
    uint64_t x = ((uint64_t)xhi << 32) | xlo;
    uint64_t eq = ~(x | (x >> 1)) & 0x5555555555555555ull;
    uint64_t r2 = eq & (eq >> 2);
    uint64_t r4 = r2 & (r2 >> 4);
    uint64_t r8 = r4 & (r4 >> 8);
    uint64_t r10 = r8 & (r2 >> 16);
    rlo = (uint32_t)r10; rhi = (uint32_t)(r10 >> 32);
}

// 32 germline bases from p (0 <= p, p + 32 <= region_len)
DCB_HD void rg_win32(const uint32_t* blob, const DcbTag& t, int p, uint32_t& lo, uint32_t& hi) {
    const int wi = p >> 4, sh = (p & 15) * 2;
    const uint32_t a = blob[t.region_off + wi], b = blob[t.region_off + wi + 1];
    const uint32_t c = sh ? blob[t.region_off + wi + 2] : 0u;      // the packer leaves one zero word behind every region
    lo = DCB_FUNNEL_R(a, b, sh);
    hi = DCB_FUNNEL_R(b, c, sh);
}

// get_v_deletions (decombine.py:749-785), literal walk with Python slice semantics
DCB_HD bool v_deletions_general(const ReadView& r, const uint32_t* blob, const DcbTag& t, int temp_end_v,
                                int& end_v, int& dels, dcb_cnt_t* C) {
    const int n = r.n, m = t.region_len;
    if (temp_end_v >= n) {                                   // :760-762
        DCB_COUNT(C, DCB_C_v_del_failed_tag_at_end);
        return false;
    }
    int f = temp_end_v + 1, pos = m - 10, nd = 0;            // :754-765
    // Interior fast-forward: while the next 23 steps compare plain 10-base slices (no wrap, no truncation, no non-ACGT
    // symbol), they are 23 alignments of ONE 32-base window of the read against one of the region -- a run of 10 equal
    // bases found bit-parallel.  A substitution in the true junction makes this walk run to the start of the read
    // (v_del_failed); stepping it base by base was 80 % of the general kernel's time.
    while (f < n && f >= 32 && pos >= 22 && pos + 10 <= m && !rd_inv_any(r, f - 32, f)) {
        uint32_t lo, hi, glo, ghi, rl, rh;
        rd_win32(r, f - 32, lo, hi);
        rg_win32(blob, t, pos - 22, glo, ghi);
        run10(lo ^ glo, hi ^ ghi, rl, rh);
        rh &= (1u << 14) - 1u;                               // window index i <= 22 <-> step 22 - i
        if (rl | rh) {
            const int top = rh ? 63 - DCB_CLZ(rh) : 31 - DCB_CLZ(rl);
            nd += 22 - (top >> 1);
            dels = nd;
            end_v = temp_end_v - nd;
            return true;
        }
        f -= 23; pos -= 23; nd += 23;
    }
    while (0 <= f && f < n) {                                // :767
        if (slices_equal(blob, t, py_slice(m, pos, pos + 10), r, py_slice(n, f - 10, f))) {  // :769-772
            dels = nd;                                       // :774-775
            end_v = temp_end_v - nd;
            return true;
        }
        pos--; nd++; f--;                                    // :777-779
    }
    DCB_COUNT(C, DCB_C_v_del_failed);                        // :784
    return false;
}

// get_j_deletions (decombine.py:788-817)
DCB_HD bool j_deletions_general(const ReadView& r, const uint32_t* blob, const DcbTag& t, int temp_start_j,
                                int end_of_v, int& start_j, int& dels, dcb_cnt_t* C) {
    const int n = r.n, m = t.region_len;
    int f = temp_start_j, pos = 0;
    // Interior fast-forward, as in v_deletions_general: past end_of_v, 23 steps at a time.
    if (f >= 0 && f < end_of_v && end_of_v + 2 < n) { pos += end_of_v - f; f = end_of_v; }   // :798-800, junction bases
    while (f >= 0 && f >= end_of_v && f + 32 <= n && pos + 32 <= m && !rd_inv_any(r, f, f + 32)) {
        uint32_t lo, hi, glo, ghi, rl, rh;
        rd_win32(r, f, lo, hi);
        rg_win32(blob, t, pos, glo, ghi);
        run10(lo ^ glo, hi ^ ghi, rl, rh);
        rh &= (1u << 14) - 1u;                               // window index i <= 22 <-> step i
        if (rl | rh) {
            const int low = rl ? DCB_FFS(rl) - 1 : 32 + DCB_FFS(rh) - 1;
            dels = pos + (low >> 1);
            start_j = f + (low >> 1);
            return true;
        }
        f += 23; pos += 23;
    }
    while (0 <= f + 2 && f + 2 < n) {                        // :795
        if (f < end_of_v) {                                  // :798-800
            pos++; f++;
        } else if (slices_equal(blob, t, py_slice(m, pos, pos + 10), r, py_slice(n, f, f + 10))) {  // :802-805
            dels = pos;                                      // :807-808
            start_j = f;
            return true;
        } else {
            pos++; f++;                                      // :810-811
        }
    }
    DCB_COUNT(C, DCB_C_j_del_failed);                        // :816
    return false;
}

// ------------------------------------------------------------------------------------------------
// findall() of one keyword set as a resumable generator: hits come out ordered by END position,
// longest keyword first at equal end -- the order acora reports them in.
// ------------------------------------------------------------------------------------------------
struct KwScan { int e, ci, cend, hi; };

DCB_HD void kw_scan_init(KwScan& s, const DcbKwSet& ks) {
    s.e = ks.kq - 1;
    s.ci = s.cend = 0;
    s.hi = 0;
}

DCB_HD const DcbKw& kwset_kw(const uint32_t* blob, const DcbKwSet& ks, int c) {
    return reinterpret_cast<const DcbKw*>(blob + ks.kw_off)[c];
}

DCB_HD bool kw_scan_next(const ReadView& r, const uint32_t* blob, const DcbKwSet& ks, KwScan& s, int& kw, int& start) {
    if (r.hits) {            // the occurrences are already listed, in this order
        while (s.hi < r.n_hits) {
            const uint32_t e = r.hits[s.hi++ * r.stride];
            if ((int)(e >> 24) == ks.set_id) { kw = (int)((e >> 16) & 255u); start = (int)(e & 0xFFFFu); return true; }
        }
        return false;
    }
    const uint32_t kmask = mask2(ks.kq);
    for (;;) {
        while (s.ci < s.cend) {
            int c = s.ci++;
            const DcbKw& k = kwset_kw(blob, ks, c);
            int st = s.e - (int)k.len;
            if (st < 0) continue;
            if (rd_equals(r, st, k.len, k.bits_lo, k.bits_hi)) { kw = c; start = st; return true; }
        }
        s.e++;
        if (r.cand) {   // skip to the next end position whose cand_kq-mer tail is in the union suffix filter
            const int pmin = s.e - r.cand_kq;
            s.e = cand_next(r, pmin) + r.cand_kq;
        }
        if (s.e > r.n) return false;
        uint32_t key = rd_win16(r, s.e - ks.kq) & kmask;
        if (!((blob[ks.bitmap_off + (key >> 5)] >> (key & 31)) & 1u)) continue;
        uint32_t h = dcb_hash32(key) & (uint32_t)ks.hash_mask;
        for (;;) {
            uint32_t slot = blob[ks.hash_off + h];
            if (slot == DCB_HASH_EMPTY) break;
            if ((slot >> 16) == key) {
                s.ci = (int)((slot >> 8) & 255u);
                s.cend = s.ci + (int)(slot & 255u);
                break;
            }
            h = (h + 1) & (uint32_t)ks.hash_mask;
        }
    }
}

// All keyword occurrences of the six sets of a chain in one pass over the marked positions (cand_build), appended per
// set in exactly the order kw_scan_next reports them.  One converged loop per warp instead of up to six scans that
// every lane enters and leaves at its own time (measured: 4.6 of 32 lanes active in the scans).  More than
// DCB_HITS_CAP occurrences: the list is dropped and the scans run as before.
DCB_HD void hits_build(ReadView& r, const uint32_t* vblob, const uint32_t* jblob, uint32_t* hits_col) {
    const DcbGene& gv = *reinterpret_cast<const DcbGene*>(vblob);
    const DcbGene& gj = *reinterpret_cast<const DcbGene*>(jblob);
    int n = 0;
    bool ok = true;
    r.hits = nullptr; r.n_hits = 0;
    const int limit = 32 * ((r.nw + 1) / 2);
    for (int p = cand_next(r, 0); p < limit; p = cand_next(r, p + 1)) {
        const int e = p + r.cand_kq;                     // candidate end position (exclusive)
        if (e > r.n) break;
        // first the six cheap bitmap tests, then the (few) sets that passed: lanes walk their own set bits together,
        // whichever sets they are, instead of all waiting while one lane confirms a keyword of one set
        uint32_t pass = 0;
        for (int si = 0; si < 6; si++) {
            const uint32_t* blob = si < 3 ? vblob : jblob;
            const DcbGene& g = si < 3 ? gv : gj;
            const DcbKwSet& ks = (si % 3) == 0 ? g.full : (si % 3) == 1 ? g.half1 : g.half2;
            if (e < ks.kq) continue;
            const uint32_t key = rd_win16(r, e - ks.kq) & mask2(ks.kq);
            pass |= ((blob[ks.bitmap_off + (key >> 5)] >> (key & 31)) & 1u) << si;
        }
        while (pass) {
            const int si = DCB_FFS(pass) - 1;
            pass &= pass - 1;
            const uint32_t* blob = si < 3 ? vblob : jblob;
            const DcbGene& g = si < 3 ? gv : gj;
            const DcbKwSet& ks = (si % 3) == 0 ? g.full : (si % 3) == 1 ? g.half1 : g.half2;
            const uint32_t key = rd_win16(r, e - ks.kq) & mask2(ks.kq);
            uint32_t h = dcb_hash32(key) & (uint32_t)ks.hash_mask;
            int ci = 0, cend = 0;
            for (;;) {
                const uint32_t slot = blob[ks.hash_off + h];
                if (slot == DCB_HASH_EMPTY) break;
                if ((slot >> 16) == key) { ci = (int)((slot >> 8) & 255u); cend = ci + (int)(slot & 255u); break; }
                h = (h + 1) & (uint32_t)ks.hash_mask;
            }
            for (int c = ci; c < cend; c++) {
                const DcbKw& k = kwset_kw(blob, ks, c);
                const int st = e - (int)k.len;
                if (st < 0 || !rd_equals(r, st, k.len, k.bits_lo, k.bits_hi)) continue;
                if (n < DCB_HITS_CAP) hits_col[n * r.stride] = (uint32_t)st | ((uint32_t)c << 16) | ((uint32_t)si << 24);
                else ok = false;
                n++;
            }
        }
    }
    if (ok) { r.hits = hits_col; r.n_hits = n; }
}

struct VJ { int idx, pos, dels, seqpos; };  // (match, end_v | start_j, deletions, v_seq_start | j_seq_end)

// vanalysis (decombine.py:273-394) and janalysis (decombine.py:397-531): same skeleton, mirrored arithmetic.
template <bool IS_V>
DCB_HD bool analyse_general(const ReadView& r, const uint32_t* blob, int end_of_v, VJ& out, dcb_cnt_t* C) {
    const DcbGene& g = *reinterpret_cast<const DcbGene*>(blob);
    const uint8_t* taglist;
    KwScan sc;
    int kw, p;

    // ---- full tags: findall, >1 hit rejects (:275-290 / :399-418)
    kw_scan_init(sc, g.full);
    int nh = 0, kw0 = 0, p0 = 0;
    while (nh < 2 && kw_scan_next(r, blob, g.full, sc, kw, p)) {
        if (nh == 0) { kw0 = kw; p0 = p; }
        nh++;
    }
    if (nh) {
        if (nh > 1) { DCB_COUNT(C, IS_V ? DCB_C_multiple_v_matches : DCB_C_multiple_j_matches); return false; }
        const DcbKw& k = kwset_kw(blob, g.full, kw0);
        int idx = k.first_tag;
        const DcbTag& t = gene_tag(blob, g, idx);
        if (IS_V) {
            int temp_end_v = p0 + t.jump - 1;
            int end_v, dels;
            if (!v_deletions_general(r, blob, t, temp_end_v, end_v, dels, C)) return false;
            out.idx = idx; out.pos = end_v; out.dels = dels; out.seqpos = p0;
        } else {
            int temp_start_j = p0 - t.jump;
            int start_j, dels;
            if (!j_deletions_general(r, blob, t, temp_start_j, end_of_v, start_j, dels, C)) return false;
            out.idx = idx; out.pos = start_j; out.dels = dels; out.seqpos = p0 + (int)k.len;
        }
        return true;
    }

    // ---- half1 (:294-335 / :422-470), then half2 only if half1 never hit (:339-390 / :473-527)
    for (int half = 1; half <= 2; half++) {
        const DcbKwSet& ks = half == 1 ? g.half1 : g.half2;
        taglist = reinterpret_cast<const uint8_t*>(blob + ks.taglist_off);
        kw_scan_init(sc, ks);
        bool any = false;
        while (kw_scan_next(r, blob, ks, sc, kw, p)) {
            any = true;
            const DcbKw& k = kwset_kw(blob, ks, kw);
            const int L0 = gene_tag(blob, g, k.first_tag).len;       // len(seqs[halfN_seqs.index(hit)])
            const int s0 = half == 1 ? p : p - g.split;              // slice start as the reference writes it
            for (int ti = 0; ti < (int)k.n_tags; ti++) {
                const int kk = taglist[k.tags_off + ti];             // ascending tag index
                const DcbTag& t = gene_tag(blob, g, kk);
                Span gs = py_slice(r.n, s0, s0 + L0);                // length guard (:302-307 etc.)
                if ((int)t.len != gs.b - gs.a) continue;
                Span hs = py_slice(r.n, s0, s0 + (int)t.len);        // Hamming operand (:311-314 etc.)
                if (hs.b - hs.a != (int)t.len) continue;
                if (!rd_hamming_le1(r, hs.a, t.len, t.bits_lo, t.bits_hi)) continue;
                if (IS_V) {
                    DCB_COUNT(C, half == 1 ? DCB_C_verr2 : DCB_C_verr1);          // :318 / :370
                    int temp_end_v = half == 1 ? p + t.jump - 1 : p + t.jump - g.split - 1;
                    int end_v, dels;
                    if (v_deletions_general(r, blob, t, temp_end_v, end_v, dels, C)) {
                        out.idx = kk; out.pos = end_v; out.dels = dels; out.seqpos = s0;
                        return true;
                    }
                } else {
                    DCB_COUNT(C, half == 1 ? DCB_C_jerr2 : DCB_C_jerr1);          // :445 / :504
                    int temp_start_j = half == 1 ? p - t.jump : p - t.jump - g.split;
                    int j_seq_end = half == 1 ? p + (int)k.len + g.split : p + (int)k.len;  // :450-454 / :511
                    int start_j, dels;
                    if (j_deletions_general(r, blob, t, temp_start_j, end_of_v, start_j, dels, C)) {
                        out.idx = kk; out.pos = start_j; out.dels = dels; out.seqpos = j_seq_end;
                        return true;
                    }
                }
            }
        }
        if (any) {
            // :334 / :389 / :469 / :526 -- the J half2 failure bumps foundv2notv1 in the reference; preserved
            if (IS_V) DCB_COUNT(C, half == 1 ? DCB_C_foundv1notv2 : DCB_C_foundv2notv1);
            else DCB_COUNT(C, half == 1 ? DCB_C_foundj1notj2 : DCB_C_foundv2notv1);
            return false;
        }
    }
    DCB_COUNT(C, IS_V ? DCB_C_no_vtags_found : DCB_C_no_j_assigned);  // :393 / :530
    return false;
}

struct DcrParams { int allow_ns, lenthreshold; };

// The four filters and the result of dcr() (decombine.py:553-581), shared by both kernels.
// Returns -1 when the rearrangement is accepted (out filled), else the counter the reference bumps.
DCB_HD int dcr_finish(const ReadView& r, int vjump, int vlen, int jjump, int jlen, const VJ& v, const VJ& j,
                      const DcrParams& prm, dcb_result& out) {
    if (!prm.allow_ns && r.e1 > r.e0) {
        Span it = py_slice(r.n, v.seqpos, j.seqpos);
        if (rd_has_N(r, it.a, it.b)) return DCB_C_dcrfilter_intertagN;                                   // :553-556
    }
    if ((v.seqpos - j.seqpos) >= prm.lenthreshold) return DCB_C_dcrfilter_toolong_intertag;              // :557-560
    if (v.dels > (vjump - vlen) || j.dels > jjump) return DCB_C_dcrfilter_imposs_deletion;               // :561-565
    if ((v.seqpos + vlen) > (j.seqpos + jlen)) return DCB_C_dcrfilter_tag_overlap;                       // :566-569
    out.status = 1;                                                                                      // :572-581
    out.frame = 0;
    out.v = (uint8_t)v.idx; out.j = (uint8_t)j.idx;
    out.vdel = (uint16_t)v.dels; out.jdel = (uint16_t)j.dels;
    out.ins_start = (uint16_t)(v.pos + 1); out.ins_end = (uint16_t)j.pos;
    out.v_seq_start = (uint16_t)v.seqpos; out.j_seq_end = (uint16_t)j.seqpos;
    return -1;
}

// dcr(read, inputargs) (decombine.py:534-585), general path: any read, any edge case.
DCB_HD bool dcr_general(const ReadView& r, const uint32_t* vblob, const uint32_t* jblob, const DcrParams& prm,
                        dcb_result& out, dcb_cnt_t* C) {
    VJ v, j;
    if (!analyse_general<true>(r, vblob, 0, v, C)) return false;                 // :542-545
    if (!analyse_general<false>(r, jblob, v.pos + 1, j, C)) {                    // :547-548, :583-585
        DCB_COUNT(C, DCB_C_VJ_assignment_failed);
        return false;
    }
    const DcbGene& gv = *reinterpret_cast<const DcbGene*>(vblob);
    const DcbGene& gj = *reinterpret_cast<const DcbGene*>(jblob);
    const DcbTag& vt = gene_tag(vblob, gv, v.idx);
    const DcbTag& jt = gene_tag(jblob, gj, j.idx);
    int c = dcr_finish(r, vt.jump, vt.len, jt.jump, jt.len, v, j, prm, out);
    if (c >= 0) { DCB_COUNT(C, c); return false; }
    return true;
}

// ------------------------------------------------------------------------------------------------
// Reverse complement of a packed read (replaces Bio.Seq.reverse_complement, decombine.py:182-184,
// for the second try of `-or both`): output word ow holds bases [16 ow, 16 ow + 16) of the result.
// ------------------------------------------------------------------------------------------------
DCB_HD uint32_t revcomp_word(const ReadView& r, int ow) {
    int p = r.n - 16 * ow - 16;            // source bases [p, p+16) reversed
    uint32_t x = rd_win16(r, p);
    uint32_t y = DCB_BREV(x);              // reverses bit order: 2-bit groups reversed, bits inside swapped
    y = ((y & 0xAAAAAAAAu) >> 1) | ((y & 0x55555555u) << 1);
    y = ~y;                                // complement: A<->T (0<->3), C<->G (1<->2)
    int valid = r.n - 16 * ow;             // bases of this output word that exist
    if (valid <= 0) return 0u;
    return y & mask2(valid < 16 ? valid : 16);
}

// ------------------------------------------------------------------------------------------------
// Exact-tag fast path.
// ------------------------------------------------------------------------------------------------
// Result of the sampled-seed search for full tags: number of distinct occurrences (saturating at 2)
// and the first one.
struct FullHit {
    int count;        // distinct occurrences, saturating at 2
    uint32_t code;    // first occurrence: tag << 16 | position
};
DCB_HD int fullhit_tag(const FullHit& fh) { return (int)(fh.code >> 16); }
DCB_HD int fullhit_pos(const FullHit& fh) { return (int)(fh.code & 0xFFFFu); }

// Record one confirmed full-tag occurrence (the same one may be reported twice).
DCB_HD void fullhit_add(FullHit& fh, int tag, int pos) {
    const uint32_t code = ((uint32_t)tag << 16) | (uint32_t)pos;
    const bool first = fh.count == 0;
    fh.count = first ? 1 : (code != fh.code ? 2 : fh.count);
    fh.code = first ? code : fh.code;
}

// A DcbSeedIndex with its hot fields in registers.  The specialised kernels overwrite the geometry fields with
// compile-time constants, which the optimiser then folds through the (force-inlined) functions below.
struct SeedIdxView {
    const uint32_t* ck;        // class-key cuckoo slots
    const uint16_t* tk;        // tag-prefix perfect-hash slots
    const DcbUTag* utag;       // compact tag records
    const uint16_t* chain;     // successor of a tag among those sharing its lmin-prefix, or null
    const uint32_t* bloom;     // compact seed filter
    uint32_t c1, c2, t1, bmul;
    int cshift, tshift, wbits;
    int q, stride, lmin, wlead, span, k, n_v;
};
DCB_HD SeedIdxView seed_idx_view(const uint32_t* ib) {
    const DcbSeedIndex& ix = *reinterpret_cast<const DcbSeedIndex*>(ib);
    SeedIdxView v;
    v.ck = ib + ix.ck_off;
    v.tk = reinterpret_cast<const uint16_t*>(ib + ix.tk_off);
    v.utag = reinterpret_cast<const DcbUTag*>(ib + ix.utag_off);
    v.chain = ix.chain_off ? reinterpret_cast<const uint16_t*>(ib + ix.chain_off) : nullptr;
    v.bloom = ib + ix.bloom_off;
    v.c1 = ix.c1; v.c2 = ix.c2; v.cshift = ix.cshift;
    v.t1 = ix.t1; v.tshift = ix.tshift;
    v.bmul = ix.bmul; v.wbits = ix.wbits;
    v.q = ix.q; v.stride = ix.stride; v.lmin = ix.lmin; v.wlead = ix.wlead; v.span = ix.span; v.k = ix.k; v.n_v = ix.n_v;
    return v;
}
DCB_HD const DcbTag* gene_tags(const uint32_t* core) {
    return core ? reinterpret_cast<const DcbTag*>(core + reinterpret_cast<const DcbGene*>(core)->tag_off) : nullptr;
}

// Class-key lookup for a filter hit at sampled position p; (wlo, whi) are the 32 bases starting at p - wlead.
// Returns the set of candidate offsets (bit o <=> a tag may start at p - o); empty for a filter false positive.
DCB_HD uint32_t fast_class_lookup(const SeedIdxView& ix, uint32_t wlo, uint32_t whi) {
    uint32_t offs = 0;
    const uint32_t kmask = mask2(ix.k);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int c = 0; c < 2; c++) {
        const int sh = 2 * (ix.wlead - c * ix.span);   // < 32: wlead <= 12
        const uint32_t x = ((uint32_t)c << 30) | (DCB_FUNNEL_R(wlo, whi, sh) & kmask);
        const uint32_t p1 = x * ix.c1, p2 = x * ix.c2;
        const uint32_t e1 = ix.ck[p1 >> ix.cshift], e2 = ix.ck[p2 >> ix.cshift];
        offs |= (((e1 ^ p2) & DCB_CK_FPMASK) == 0 ? e1 : 0u) | (((e2 ^ p1) & DCB_CK_FPMASK) == 0 ? e2 : 0u);
    }
    return DCB_CK_OFFMASK(offs);
}

// One candidate offset o of a seed hit at p: the perfect-hash prefix table names the tag whose lmin-prefix starts at
// P = p - o; it (and the rare tags chained to it) is compared with the read as a whole and recorded.
DCB_HD void fast_check_offset(const ReadView& r, const SeedIdxView& ix, int p, int o, uint32_t wlo, uint32_t whi,
                              FullHit& vh, FullHit& jh) {
    const int P = p - o;
    const int sh = 2 * (ix.wlead - o);                 // 2 <= sh <= 2 * wlead <= 24
    const uint32_t lo = DCB_FUNNEL_R(wlo, whi, sh), hi = whi >> sh;   // the 32 - (wlead - o) >= 20 bases from P on
    const uint32_t f = dcb_fold64(lo & mask2(ix.lmin), ix.lmin > 16 ? (hi & mask2(ix.lmin - 16)) : 0u);
    uint32_t ctag = ix.tk[(f * ix.t1) >> ix.tshift];
    if (P < 0) return;
    while (ctag != 0x1FFu) {
        const DcbUTag u = ix.utag[ctag];
        const int L = (int)(u.mask_hi_len >> 24);
        if (P + L <= r.n) {
            uint32_t tlo = lo, thi = hi;
            if (ix.wlead - o + L > 32) rd_win32(r, P, tlo, thi);   // long tag: past the window in registers
            if (!(((tlo ^ u.bits_lo) & u.mask_lo) | ((thi ^ u.bits_hi) & (u.mask_hi_len & 0x00FFFFFFu)))) {
                if ((int)ctag >= ix.n_v) fullhit_add(jh, (int)ctag - ix.n_v, P);
                else fullhit_add(vh, (int)ctag, P);
            }
        }
        ctag = ix.chain ? ix.chain[ctag] : 0x1FFu;
    }
}

// The second 16 bytes of a DcbTag: all the exact-tag path needs after the match itself, in one 128-bit load.
struct alignas(16) DcbTagFin {
    uint32_t edge_lo, edge_hi;
    int16_t jump, region_len;
    uint8_t len, edge_ok, next_same_prefix, pad8;
};
DCB_HD DcbTagFin tag_fin(const DcbTag* tags, int k) {
    return *reinterpret_cast<const DcbTagFin*>(reinterpret_cast<const uint32_t*>(tags + k) + 4);
}

// 32 bases from p.  PADDED: the caller guarantees -16 <= p and (p >> 4) + 2 <= nw with a zero row before and behind the
// read's words (the specialised kernel's shared-memory columns), so no bounds checks are needed.
template <bool PADDED>
DCB_HD void rd_win32x(const ReadView& r, int p, uint32_t& lo, uint32_t& hi) {
    if (PADDED) {
        const uint32_t* c0 = r.w + (p >> 4) * r.stride;
        const uint32_t a = c0[0], b = c0[r.stride], c = c0[2 * r.stride];
        const int sh = (p & 15) * 2;
        lo = DCB_FUNNEL_R(a, b, sh);
        hi = DCB_FUNNEL_R(b, c, sh);
    } else {
        rd_win32(r, p, lo, hi);
    }
}

// Fast V: exactly the interior case of get_v_deletions (decombine.py:749-785).  Returns
//   1 handled (end_v / dels set), 0 defer to the general kernel.
// ri: the read's invalid-base column (01 per non-ACGT symbol, the read's own layout) or null: an invalid base never matches.
template <bool PADDED>
DCB_HD int fast_v_deletions(const ReadView& r, const DcbTagFin& t, int temp_end_v, int& end_v, int& dels, const ReadView* ri = nullptr) {
    const int f0 = temp_end_v + 1;
    if (!t.edge_ok || f0 >= r.n || f0 < 32) return 0;
    uint32_t lo, hi, rl, rh;
    rd_win32x<PADDED>(r, f0 - 32, lo, hi);
    lo ^= t.edge_lo; hi ^= t.edge_hi;
    if (ri) { uint32_t ilo, ihi; rd_win32x<PADDED>(*ri, f0 - 32, ilo, ihi); lo |= ilo; hi |= ihi; }
    run10(lo, hi, rl, rh);
    // window index i <-> deletions nd = 22 - i; want the smallest nd, i.e. the highest i <= 22 (bits 2i, i <= 22)
    rh &= (1u << 14) - 1u;
    if (!(rl | rh)) return 0;
    const int top = rh ? 63 - DCB_CLZ(rh) : 31 - DCB_CLZ(rl);
    dels = 22 - (top >> 1);
    end_v = temp_end_v - dels;
    return 1;
}

// Fast J: the interior case of get_j_deletions (decombine.py:788-817).
template <bool PADDED>
DCB_HD int fast_j_deletions(const ReadView& r, const DcbTagFin& t, int temp_start_j, int end_of_v, int& start_j, int& dels,
                            const ReadView* ri = nullptr) {
    if (!t.edge_ok || temp_start_j < 0) return 0;
    if (PADDED && temp_start_j >= 16 * (r.nw - 1)) return 0;   // the window would leave the padded columns
    int pos0 = end_of_v - temp_start_j;
    if (pos0 < 0) pos0 = 0;
    // the 10-mer must lie inside the read: temp_start_j + i + 10 <= n
    int imax = r.n - 10 - temp_start_j;
    if (pos0 > 22 || imax < 0) return 0;
    if (imax > 22) imax = 22;
    uint32_t lo, hi, rl, rh;
    rd_win32x<PADDED>(r, temp_start_j, lo, hi);
    lo ^= t.edge_lo; hi ^= t.edge_hi;
    if (ri) { uint32_t ilo, ihi; rd_win32x<PADDED>(*ri, temp_start_j, ilo, ihi); lo |= ilo; hi |= ihi; }
    run10(lo, hi, rl, rh);
    uint64_t r10 = ((uint64_t)rh << 32) | rl;
    r10 &= ~((1ull << (2 * pos0)) - 1);                      // pos >= pos0
    r10 &= (1ull << (2 * imax + 2)) - 1;                     // pos <= imax (<= 22)
    if (!r10) return 0;
    const int low = (uint32_t)r10 ? DCB_FFS((uint32_t)r10) - 1 : 32 + DCB_FFS((uint32_t)(r10 >> 32)) - 1;
    dels = low >> 1;
    start_j = temp_start_j + dels;
    return 1;
}

// Sampled-seed search through one index over a read held in the view: probe the seed filter at every
// multiple of `stride`, 32 probes at a time into a hit mask, then confirm the (rare) hits in a second loop
// so the lanes of a warp stay converged during the probes.  (The kernels have a register-resident unrolled
// specialisation of the probing, over bank-private copies of the filter; this is the generic form.)
DCB_HD void fast_find(const ReadView& r, const uint32_t* ib, FullHit& vh, FullHit& jh, bool stop_at_two_v) {
    const SeedIdxView ix = seed_idx_view(ib);
    const int last = r.n - ix.q;  // last start position of a whole q-mer
    for (int base = 0; base <= last; base += 32 * ix.stride) {
        uint32_t hits = 0;
        for (int i = 0; i < 32; i++) {
            const int p = base + i * ix.stride;
            if (p > last) break;
            const uint32_t win = rd_win16(r, p);
            const uint32_t word = ix.bloom[DCB_BLOOM_WORD(win, ix.bmul, ix.wbits)];
            hits |= ((word >> DCB_BLOOM_BIT(win)) & 1u) << i;
        }
        while (hits) {
            const int i = DCB_FFS(hits) - 1;
            hits &= hits - 1;
            const int p = base + i * ix.stride;
            uint32_t wlo, whi;
            rd_win32(r, p - ix.wlead, wlo, whi);
            for (uint32_t offs = fast_class_lookup(ix, wlo, whi); offs; offs &= offs - 1)
                fast_check_offset(r, ix, p, DCB_FFS(offs) - 1, wlo, whi, vh, jh);
            if (stop_at_two_v && vh.count >= 2) return;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Flat kernel (dcb_exact_kernel_flat): byte filter + hash-and-displace offset table (DcbSeedIndex, second half).
// A full-tag occurrence is kept as ONE word per gene: 0 = none, DCB_HIT_MULTI = two or more distinct occurrences,
// else DCB_HIT_ONE | tag << 16 | position.
// ------------------------------------------------------------------------------------------------
#define DCB_HIT_ONE 0x80000000u
#define DCB_HIT_MULTI 0xFFFFFFFFu
#define DCB_HIT_UNKNOWN 0xFFFFFFFEu     // hand-over only: the gene was not searched (J, when the read has no single full V tag)
// add an occurrence (or merge the state another lane collected) into a hit word
DCB_HD uint32_t hit_merge(uint32_t cur, uint32_t c) {
    return cur == 0u ? c : ((c == 0u || c == cur) ? cur : DCB_HIT_MULTI);
}
DCB_HD FullHit hit_decode(uint32_t h) {
    FullHit fh;
    fh.count = h == 0u ? 0 : (h == DCB_HIT_MULTI ? 2 : 1);
    fh.code = h & 0x7FFFFFFFu;
    return fh;
}

// Hand-over word of the exact-tag kernels for one gene: 0 = no full tag in the read, DCB_HIT_MULTI = several,
// else DCB_HIT_ONE | tag << 16 | position (tag numbered inside its gene).
DCB_HD uint32_t half_word_of(const FullHit& fh) {
    return fh.count == 0 ? 0u : (fh.count > 1 ? DCB_HIT_MULTI : (DCB_HIT_ONE | fh.code));
}

struct alignas(8) DcbTq { uint32_t x, y; };   // one tag slot: prefix_lo, meta
struct QIdxView {
    const uint16_t* disp;      // 2^b1 displacements
    const uint16_t* offs;      // 2^b2 offset sets
    const DcbTq* tq;           // tag slots: {prefix_lo, meta}
    const DcbUTag* utag;
    const uint16_t* chain;
    uint32_t m1, m2, ta, tb, mask2;
    int s1, s2, tqshift;
    int q, stride, lmin, wlead, n_v;
};
// ib: the index blob; qtab: where its qtab_words were staged (ib + qtab_off when the blob is used in place)
DCB_HD QIdxView q_idx_view(const uint32_t* ib, const uint32_t* qtab) {
    const DcbSeedIndex& ix = *reinterpret_cast<const DcbSeedIndex*>(ib);
    QIdxView v;
    v.disp = reinterpret_cast<const uint16_t*>(qtab);
    v.offs = v.disp + ((size_t)1 << ix.b1);
    v.tq = reinterpret_cast<const DcbTq*>(qtab + (ix.tq_off - ix.qtab_off));
    v.utag = reinterpret_cast<const DcbUTag*>(ib + ix.utag_off);
    v.chain = ix.chain_off ? reinterpret_cast<const uint16_t*>(ib + ix.chain_off) : nullptr;
    v.m1 = ix.m1; v.m2 = ix.m2; v.ta = ix.ta; v.tb = ix.tb; v.mask2 = (1u << ix.b2) - 1u;
    v.s1 = 32 - ix.b1; v.s2 = 32 - ix.b2; v.tqshift = 32 - ix.tq_bits;
    v.q = ix.qq; v.stride = ix.qstride; v.lmin = ix.lmin; v.wlead = ix.qwlead; v.n_v = ix.n_v;
    return v;
}
// Offsets the q-mer in the low 2q bits of x occurs at in some tag (bit o); anything for a q-mer that is not indexed.
DCB_HD uint32_t q_offsets(const QIdxView& ix, uint32_t x) {
    const uint32_t d = ix.disp[(x * ix.m1) >> ix.s1];
    return ix.offs[(((x * ix.m2) >> ix.s2) + d) & ix.mask2];
}
// One candidate: a tag starting at P = p - o, where (wlo, whi) are the 32 bases from p - wlead.  Calls
// sink(ctag, P) for every tag found there (one, unless tags share their lmin-prefix).  The lmin-prefix is looked up in
// the tag slots: one 64-bit read confirms a tag of minimum length; longer tags and tags sharing a prefix (DCB_TQ_MORE)
// go on to the whole-tag compare.
template <bool PADDED, class Sink>
DCB_HD void q_check_offset(const ReadView& r, const QIdxView& ix, int p, int o, uint32_t wlo, uint32_t whi, Sink& sink) {
    const int P = p - o;
    const int sh = 2 * (ix.wlead - o);                 // 0 <= sh <= 2 * wlead <= 22
    const uint32_t lo = DCB_FUNNEL_R(wlo, whi, sh), hi = whi >> sh;   // the 32 - (wlead - o) >= lmin + 5 bases from P on
    const uint32_t hp = hi & ((1u << DCB_TQ_HIBITS(ix.lmin)) - 1u);
    const DcbTq e = ix.tq[(lo * ix.ta + hp * ix.tb) >> ix.tqshift];
    if (e.x != lo || ((hp ^ e.y) & DCB_TQ_CMPMASK(ix.lmin)) != 0u || P < 0) return;
    uint32_t ctag = DCB_TQ_CTAG(e.y);
    if (!(e.y & DCB_TQ_MORE)) {
        if (P + (int)DCB_TQ_LEN(e.y) <= r.n) sink(ctag, P);
        return;
    }
    while (ctag != 0x1FFu) {
        const DcbUTag u = ix.utag[ctag];
        const int L = (int)(u.mask_hi_len >> 24);
        if (P + L <= r.n) {
            uint32_t tlo = lo, thi = hi;
            if (ix.wlead - o + L > 32) rd_win32x<PADDED>(r, P, tlo, thi);   // long tag: past the window in registers
            if (!(((tlo ^ u.bits_lo) & u.mask_lo) | ((thi ^ u.bits_hi) & (u.mask_hi_len & 0x00FFFFFFu)))) sink(ctag, P);
        }
        ctag = ix.chain ? ix.chain[ctag] : 0x1FFu;
    }
}
// Collects occurrences into two hit words (V, J).
struct HitWords {
    uint32_t v, j;             // hit words over the combined tag numbering (J tags still carry + n_v)
    int n_v;
    DCB_HD void operator()(uint32_t ctag, int P) {
        const bool is_j = (int)ctag >= n_v;
        const uint32_t c = DCB_HIT_ONE | (ctag << 16) | (uint32_t)P;
        const uint32_t m = hit_merge(is_j ? j : v, c);
        v = is_j ? v : m;
        j = is_j ? m : j;
    }
    DCB_HD void decode(FullHit& vh, FullHit& jh) const {
        vh = hit_decode(v);
        jh = hit_decode(j);
        jh.code -= (uint32_t)n_v << 16;          // meaningful only when jh.count == 1
    }
};
// The non-ACGT symbols of ONE read, in registers: up to four positions (16 bits each, 0xFFFF = none), fetched once per
// read (the kernel does it warp-wide: the 32 lanes of a warp own 32 consecutive reads, i.e. one run of the sorted list).
// A read with more of them is not searched by the exact-tag kernel (over: it goes to the general kernel).
struct ExcProbe {
    uint32_t p01, p23;
    uint32_t e0;                // the read's first entry in the exception list
    bool over;
};
DCB_HD ExcProbe exc_probe_none() {
    ExcProbe x;
    x.p01 = x.p23 = 0xFFFFFFFFu; x.e0 = 0; x.over = false;
    return x;
}
DCB_HD void exc_probe_add(ExcProbe& x, int slot, uint32_t pos) {   // slot 0..3
    const uint32_t sh = (slot & 1) * 16, m = 0xFFFFu << sh;
    uint32_t& w = slot < 2 ? x.p01 : x.p23;
    w = (w & ~m) | (pos << sh);
}
// The read's entries from the sorted list (tests/sim, and the shape the kernel's warp-wide fetch reproduces):
// index[k] = first entry whose read is >= 32 k; the list ends with read = 0xFFFFFFFF.  Entries of kind 3 are real bases
// in this frame.
DCB_HD ExcProbe exc_probe_load(const uint32_t* read, const uint16_t* pos, const uint8_t* kind, const uint32_t* index, uint32_t ri) {
    ExcProbe x = exc_probe_none();
    uint32_t e = index[ri >> 5];
    while (read[e] < ri) e++;
    x.e0 = e;
    int n = 0;
    for (; read[e] == ri; e++) {
        if (kind[e] == 3) continue;
        if (n < 4) exc_probe_add(x, n, pos[e]);
        n++;
    }
    x.over = n > 4;
    return x;
}
DCB_HD bool exc_in_span(const ExcProbe& x, int lo, int hi) {
    const int a = (int)(x.p01 & 0xFFFFu), b = (int)(x.p01 >> 16), c = (int)(x.p23 & 0xFFFFu), d = (int)(x.p23 >> 16);
    return (a >= lo && a < hi) || (b >= lo && b < hi) || (c >= lo && c < hi) || (d >= lo && d < hi);
}
// The same for a read with non-ACGT symbols (packed as base 0, which can fake an 'A'): an occurrence that covers such a
// symbol is no occurrence.  The exception list is only consulted when it matters -- when a SECOND distinct occurrence
// turns up (is one of the two a fake?) and once at the end for the single occurrence kept (finish) -- so the common
// path through the confirmation loop is the plain one.  After finish() the hit words are exact for these reads too.
struct HitWordsX {
    HitWords hw;
    const ExcProbe* xp;        // null: the read has no such symbols
    const DcbUTag* utag;
    DCB_HD bool fake(uint32_t c) const {                             // c: DCB_HIT_ONE | ctag << 16 | position
        const int P = (int)(c & 0xFFFFu);
        return exc_in_span(*xp, P, P + (int)(utag[(c >> 16) & 0x7FFFu].mask_hi_len >> 24));
    }
    DCB_HD void operator()(uint32_t ctag, int P) {
        if (xp) {
            const uint32_t c = DCB_HIT_ONE | (ctag << 16) | (uint32_t)P;
            uint32_t& cur = (int)ctag >= hw.n_v ? hw.j : hw.v;
            if (cur != 0u && cur != c && cur != DCB_HIT_MULTI) {     // a second occurrence: drop whichever is a fake
                if (fake(c)) return;
                if (fake(cur)) { cur = c; return; }
            }
        }
        hw(ctag, P);
    }
    DCB_HD void finish() {
        if (!xp) return;
        if (hw.v != 0u && hw.v != DCB_HIT_MULTI && fake(hw.v)) hw.v = 0u;
        if (hw.j != 0u && hw.j != DCB_HIT_MULTI && fake(hw.j)) hw.j = 0u;
    }
};
// The whole search for one read, serially (tests/sim and nothing else: the kernel spreads this work over a warp).
DCB_HD void q_find(const ReadView& r, const uint32_t* ib, FullHit& vh, FullHit& jh, const ExcProbe* xp = nullptr) {
    const DcbSeedIndex& hd = *reinterpret_cast<const DcbSeedIndex*>(ib);
    const QIdxView ix = q_idx_view(ib, ib + hd.qtab_off);
    const uint8_t* filt = reinterpret_cast<const uint8_t*>(ib + hd.bfilter_off);
    HitWordsX hw;
    hw.hw.v = 0; hw.hw.j = 0; hw.hw.n_v = ix.n_v; hw.xp = xp; hw.utag = ix.utag;
    for (int p = 0; p + ix.q <= r.n; p += ix.stride) {
        const uint32_t win = rd_win16(r, p);
        if (!filt[DCB_FSLOT(win, hd.fmul, hd.fbits)]) continue;
        uint32_t wlo, whi;
        rd_win32(r, p - ix.wlead, wlo, whi);
        for (uint32_t offs = q_offsets(ix, DCB_FUNNEL_R(wlo, whi, 2 * ix.wlead)); offs; offs &= offs - 1)
            q_check_offset<false>(r, ix, p, DCB_FFS(offs) - 1, wlo, whi, hw);
    }
    hw.finish();
    hw.hw.decode(vh, jh);
}

// Outcome of the fast path for one read.
enum { FAST_DONE = 0, FAST_DEFER = 1 };

// dcr() for the common case: read without exceptions, exactly one full V tag and one full J tag whose
// deletion walks stay in the interior.  Anything else is deferred UNCOUNTED to the general kernel,
// except the two outcomes that are final by themselves (multiple V / multiple J matches).
// When both_frames is set a failed first frame must be retried, so every non-success defers.
// Reads with non-ACGT symbols (packed as base 0, which can fake an 'A'): the exact-tag search drops every occurrence that
// covers such a symbol (HitWordsX), so its hit words are exact, and its outcome stands whenever no symbol lies in anything
// else the reference looks at for this read -- the two deletion windows, the inter-tag span.  xp names the read's entries of
// the sparse exception list; anything else about such a read is deferred.
template <bool PADDED>
DCB_HD int dcr_fast_from_hits(const ReadView& r, const DcbTag* vtags, const DcbTag* jtags, const FullHit& vh,
                              const FullHit& jh, const DcrParams& prm, int both_frames, dcb_result& out,
                              dcb_cnt_t* C, const bool use_xp = false, const ExcProbe xp = exc_probe_none()) {
    if (vh.count == 0) return FAST_DEFER;
    if (vh.count > 1) {
        if (both_frames) return FAST_DEFER;
        DCB_COUNT(C, DCB_C_multiple_v_matches);
        return FAST_DONE;
    }
    const DcbTagFin vt = tag_fin(vtags, fullhit_tag(vh));
    VJ v, j;
    v.idx = fullhit_tag(vh); v.seqpos = fullhit_pos(vh);
    if (!fast_v_deletions<PADDED>(r, vt, v.seqpos + vt.jump - 1, v.pos, v.dels)) return FAST_DEFER;
    if (use_xp) {   // a symbol in the V deletion window: the walk above compared a fake base
        const int f0 = v.seqpos + vt.jump;
        if (exc_in_span(xp, f0 - 32, f0)) return FAST_DEFER;
    }
    if (jh.count == 0) return FAST_DEFER;
    if (jh.count > 1) {
        if (both_frames) return FAST_DEFER;
        DCB_COUNT(C, DCB_C_multiple_j_matches);
        DCB_COUNT(C, DCB_C_VJ_assignment_failed);
        return FAST_DONE;
    }
    const DcbTagFin jt = tag_fin(jtags, fullhit_tag(jh));
    j.idx = fullhit_tag(jh); j.seqpos = fullhit_pos(jh) + (int)jt.len;
    if (!fast_j_deletions<PADDED>(r, jt, fullhit_pos(jh) - jt.jump, v.pos + 1, j.pos, j.dels)) return FAST_DEFER;
    if (use_xp) {
        const int f0 = v.seqpos + vt.jump, tsj = fullhit_pos(jh) - jt.jump;       // V window [f0 - 32, f0), J window [tsj, tsj + 32)
        const int lo = v.seqpos < f0 - 32 ? v.seqpos : f0 - 32;
        int hi = j.seqpos > tsj + 32 ? j.seqpos : tsj + 32;
        if (f0 > hi) hi = f0;
        if (exc_in_span(xp, lo, hi)) return FAST_DEFER;
    }
    // filters: a failed filter is final unless the other frame still has to be tried
    int c = dcr_finish(r, vt.jump, vt.len, jt.jump, jt.len, v, j, prm, out);
    if (c >= 0) {
        if (both_frames) return FAST_DEFER;
        DCB_COUNT(C, c);
    }
    return FAST_DONE;
}

// ------------------------------------------------------------------------------------------------
// Per-read drivers shared by the kernels (and by tests/sim): everything between "the packed words of
// read ri are in w[]" and "here is its 16-byte record".
// ------------------------------------------------------------------------------------------------
struct ExcList {
    const uint32_t* read;   // sorted read indices
    const uint16_t* pos;
    const uint8_t* kind;    // 1 'N', 2 other, 3 valid in the packed frame but not in its reverse complement
    uint32_t n;             // end of the range that can hold a read's entries (the kernels: the end of its 32-read group's run)
    uint32_t lo = 0;        // start of that range
};

DCB_HD uint32_t exc_lower_bound(const ExcList& ex, uint32_t key) {
    uint32_t lo = ex.lo, hi = ex.n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (ex.read[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Exact-tag kernel body for one read whose words are already in r.w.  vidx is the V seed index -- or the
// union index of both genes when jidx is null.  Returns FAST_DONE / FAST_DEFER.
DCB_HD int dcr_exact_read(const ReadView& r, bool flagged, const uint32_t* vcore, const uint32_t* jcore,
                          const uint32_t* vidx, const uint32_t* jidx, const DcrParams& prm, int both_frames,
                          dcb_result& out, dcb_cnt_t* C, bool use_q = false, const ExcProbe* xp = nullptr,
                          uint32_t* hand = nullptr) {
    // a read with non-ACGT symbols is not searched here unless the caller handed its symbols over (the flat kernel, up to
    // four of them): behind the bit-filter kernels the half-tag path searches both genes itself ("unknown"); a read the flat
    // kernel skips is passed on by the half-tag path ("several")
    if (hand) { hand[0] = hand[1] = use_q ? DCB_HIT_MULTI : DCB_HIT_UNKNOWN; }
    if (flagged && (!xp || xp->over)) return FAST_DEFER;
    if (!flagged) xp = nullptr;
    FullHit vh, jh;
    vh.count = 0; vh.code = 0;
    jh.count = 0; jh.code = 0;
    if (use_q) {          // the flat kernel's tables (union index only)
        q_find(r, vidx, vh, jh, xp);
    } else if (!jidx) {
        fast_find(r, vidx, vh, jh, true);
    } else {
        fast_find(r, vidx, vh, jh, true);
        if (vh.count == 1) fast_find(r, jidx, vh, jh, false);
    }
    if (hand) { hand[0] = half_word_of(vh); hand[1] = (!use_q && jidx && vh.count != 1) ? DCB_HIT_UNKNOWN : half_word_of(jh); }
    return dcr_fast_from_hits<false>(r, gene_tags(vcore), gene_tags(jcore), vh, jh, prm, both_frames, out, C, xp != nullptr,
                                     xp ? *xp : exc_probe_none());
}

// General kernel body for one read: r has w/stride/n/nw set; inv0, rd1, inv1 are this thread's scratch
// columns (same stride as r): inv0/inv1 hold (nw+1)/2 words, rd1 holds nw words (only used with both_frames).
// The general path in two steps, so that a thread block can regroup its reads in between (the kernel sorts them by what
// dcr_general_prepare found, then every warp runs mostly one path of the analysis):
//   dcr_general_prepare  candidate marks, invalid-base mask, hit list of the first frame; returns the read's class
//   dcr_general_run      the analysis of both frames on a prepared view
// sf / cand0 / hits0: the chain's union suffix filter, (nw+1)/2 words for the marks and DCB_HITS_CAP words for the hit
// list -- or null: the scans visit every position.
DCB_HD int dcr_general_prepare(ReadView& r, uint32_t ri, bool flagged, const ExcList& ex, uint32_t* inv0,
                               const uint32_t* vblob, const uint32_t* jblob, const uint32_t* sf, uint32_t* cand0,
                               uint32_t* hits0) {
    const int nwi = (r.nw + 1) / 2;
    r.inv = nullptr; r.exc_pos = ex.pos; r.exc_kind = ex.kind; r.e0 = r.e1 = 0; r.mirror = 0;
    r.cand = nullptr; r.cand_kq = 0; r.hits = nullptr; r.n_hits = 0;
    if (sf && cand0) cand_build(r, sf, cand0);
    if (flagged) {
        uint32_t e0 = exc_lower_bound(ex, ri), e1 = e0;
        while (e1 < ex.n && ex.read[e1] == ri) e1++;
        r.e0 = (int)e0; r.e1 = (int)e1;
        bool any = false;
        for (int k = 0; k < nwi; k++) inv0[k * r.stride] = 0;
        for (uint32_t e = e0; e < e1; e++) {
            if (ex.kind[e] == 3) continue;  // a real base in this frame
            uint32_t p = ex.pos[e];
            inv0[(p >> 5) * r.stride] |= 1u << (p & 31);
            any = true;
        }
        if (any) r.inv = inv0;
    }
    if (r.cand && hits0) hits_build(r, vblob, jblob, hits0);      // after r.inv: an occurrence needs valid bases
    // class: which paths of the analysis the read will take (full V tag found? full J tag found? non-ACGT symbols?)
    int nv = 0, nj = 0;
    if (r.hits)
        for (int i = 0; i < r.n_hits; i++) {
            const uint32_t set = r.hits[i * r.stride] >> 24;
            nv += set == 0u; nj += set == 3u;
        }
    return (r.hits ? 0 : 8) | (nv == 0 ? 1 : 0) | (nj == 0 ? 2 : 0) | (r.inv ? 4 : 0);
}

DCB_HD void dcr_general_run(ReadView r, const ExcList& ex, uint32_t* rd1, uint32_t* inv1, const uint32_t* vblob,
                            const uint32_t* jblob, const DcrParams& prm, int both_frames, dcb_result& out, dcb_cnt_t* C,
                            const uint32_t* sf, uint32_t* cand0, uint32_t* hits0) {
    const int nwi = (r.nw + 1) / 2;
    // One copy of the analysis code for both frames (the general kernel is instruction-fetch bound: 24 warps at
    // different places of ~90 KB of code): frame 1 re-enters the same loop body with the mirrored read.
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int frame = 0; frame < (both_frames ? 2 : 1); frame++) {
        if (frame == 1) {                                                  // decombine.py:1005-1010
            for (int k = 0; k < r.nw; k++) rd1[k * r.stride] = revcomp_word(r, k);
            const ReadView r0 = r;
            r.w = rd1; r.inv = nullptr; r.mirror = 1;
            if (r0.e1 > r0.e0) {
                for (int k = 0; k < nwi; k++) inv1[k * r.stride] = 0;
                for (int e = r0.e0; e < r0.e1; e++) {
                    uint32_t p = (uint32_t)(r.n - 1) - ex.pos[e];
                    inv1[(p >> 5) * r.stride] |= 1u << (p & 31);
                }
                r.inv = inv1;
            }
            r.hits = nullptr; r.n_hits = 0;
            if (sf && cand0) cand_build(r, sf, cand0);                    // the first frame is done with its marks
            if (r.cand && hits0) hits_build(r, vblob, jblob, hits0);
        }
        dcb_result o;
        o.status = 0; o.frame = 0; o.v = o.j = 0; o.vdel = o.jdel = 0;
        o.ins_start = o.ins_end = o.v_seq_start = o.j_seq_end = 0;
        if (dcr_general(r, vblob, jblob, prm, o, C)) { o.frame = (uint8_t)frame; out = o; break; }
    }
}

// ------------------------------------------------------------------------------------------------
// Half-tag path (dcb_halftag_kernel): reads the exact-tag kernel could not finish -- almost always because a
// substitution or an N sits in one of the two tags.  The exact-tag kernel hands over what it found (one hit word per
// gene); the half keywords of the gene(s) still missing are found through the sampled index (DcbHalfIndex), every
// occurrence is expanded into its (tag, start) candidates, kept sorted in the order the reference tries them
// (decombine.py:294-390, 422-527: half1 hits in findall order, each with its tags ascending; half2 only when half1
// never hit), and the candidates are then tried one by one: length guard, Hamming <= 1, counter, deletion walk.
// Only the INTERIOR case is decided here (tag windows inside the read, deletion walks that end inside one 32-base
// window); everything else -- and nothing has been counted by then: counters are kept pending in a bit mask -- is
// passed on to the general kernel.  Non-ACGT symbols (packed as base 0) are honoured through a second column in the
// read's own layout (01 per invalid base) that is OR-ed into every comparison: an invalid base never matches.
// ------------------------------------------------------------------------------------------------
struct HalfView {
    const uint16_t* t;
    const uint32_t* h;
    const uint8_t* ids;
    const DcbHalfKw* kw;
    const uint8_t* tags;
    const uint16_t* jt;        // j_short: direct table over the first DCB_HALF_JQ bases of the J half keywords
    uint32_t c1, c2;
    int hshift, v_split, j_split, j_ok, j_short;
};
DCB_HD HalfView half_view(const uint32_t* hb) {
    const DcbHalfIndex& hx = *reinterpret_cast<const DcbHalfIndex*>(hb);
    HalfView v;
    v.t = reinterpret_cast<const uint16_t*>(hb + hx.t_off);
    v.h = hb + hx.h_off;
    v.ids = reinterpret_cast<const uint8_t*>(hb + hx.ids_off);
    v.kw = reinterpret_cast<const DcbHalfKw*>(hb + hx.kw_off);
    v.tags = reinterpret_cast<const uint8_t*>(hb + hx.tags_off);
    v.jt = hx.j_short ? reinterpret_cast<const uint16_t*>(hb + hx.jt_off) : nullptr;
    v.c1 = hx.c1; v.c2 = hx.c2; v.hshift = hx.hshift; v.v_split = hx.v_split; v.j_split = hx.j_split; v.j_ok = hx.j_ok;
    v.j_short = hx.j_short;
    return v;
}
// A candidate: gene << 31 | kind << 29 (0 full tag, 1 half1, 2 half2) | end of the keyword occurrence << 19 |
// (31 - keyword length) << 14 | tag << 6.  Only (occurrence, tag) pairs that pass the length guard and Hamming <= 1 become
// candidates; an occurrence that yields none is remembered as a DCB_HALF_SEEN bit of the read's candidate word (it still
// says that its half keyword DID occur).  Ascending integer order = V before J, full tag, then half1 hits by
// (end, longest first) with their tags ascending, then half2 hits likewise: the reference's order.
#define DCB_HC_GENE(e) ((e) >> 31)
#define DCB_HC_KIND(e) (((e) >> 29) & 3u)
#define DCB_HC_END(e) (((e) >> 19) & 1023u)
#define DCB_HC_KWLEN(e) (31u - (((e) >> 14) & 31u))
#define DCB_HC_TAG(e) (((e) >> 6) & 255u)
#define DCB_HC_DEAD(e) (((e) >> 5) & 1u)
#define DCB_HC_MAKE(gene, kind, end, kwlen, tag, dead) \
    (((uint32_t)(gene) << 31) | ((uint32_t)(kind) << 29) | ((uint32_t)(end) << 19) | ((31u - (uint32_t)(kwlen)) << 14) | ((uint32_t)(tag) << 6) | ((uint32_t)(dead) << 5))
#define DCB_HALF_MAX_READ 1008     // `end` has 10 bits
// A read's candidate word: the number of candidates appended (12 bits) | which half keyword sets occurred WITHOUT
// yielding a candidate (length guard or Hamming <= 1 failed for every tag: DCB_HALF_SEEN, one bit per gene and half) |
// DCB_HALF_BAIL added any number of times: pass the read on.
#define DCB_HALF_COUNT(n) ((n) & 0xFFFu)
#define DCB_HALF_SEEN(gene, half2) (0x1000u << (2 * (gene) + (half2)))
#define DCB_HALF_BAIL 0x10000u
// Candidates are appended in any order (in the kernel by whichever lane confirmed the occurrence: the count is bumped
// atomically) and sorted once they are complete.
#if defined(__CUDA_ARCH__)
#define DCB_SLOT_TAKE(p, k) atomicAdd((p), (k))
#define DCB_SLOT_OR(p, k) atomicOr((p), (k))
#else
#define DCB_SLOT_TAKE(p, k) ((*(p) += (k)) - (k))
#define DCB_SLOT_OR(p, k) (*(p) |= (k))
#endif
DCB_HD void half_append(uint32_t* cand, int stride, int cap, uint32_t* n, uint32_t e) {
    const uint32_t k = DCB_HALF_COUNT(DCB_SLOT_TAKE(n, 1u));
    if (k < (uint32_t)cap) cand[k * stride] = e;
}
DCB_HD void half_sort(uint32_t* cand, int stride, int n) {
    for (int i = 1; i < n; i++) {
        const uint32_t e = cand[i * stride];
        int k = i;
        while (k > 0 && cand[(k - 1) * stride] > e) { cand[k * stride] = cand[(k - 1) * stride]; k--; }
        cand[k * stride] = e;
    }
}
// The view of the invalid-base column: same geometry as the read.
DCB_HD ReadView half_inv_view(const ReadView& r, const uint32_t* inv2) {
    ReadView ri = r;
    ri.w = inv2;
    return ri;
}
// One probe hit: a keyword of `set` may start at P.  half_prefix looks its kmin-prefix up: 0 when no keyword of the set
// starts like that, else the keyword list (position | count << 8).  half_keywords then compares every keyword of the list
// with the read as a whole: sink(keyword record number, number of tags that have this half) per occurrence.
template <bool PADDED>
DCB_HD uint32_t half_prefix(const ReadView& r, const HalfView& hx, int set, int P) {
    if (P < 0) return 0u;
    constexpr int KMIN = DCB_HALF_Q + DCB_HALF_STRIDE - 1;
    uint32_t lo, hi;
    rd_win32x<PADDED>(r, P, lo, hi);
    const uint32_t key = ((uint32_t)set << 28) | (lo & mask2(KMIN));
    const uint32_t s1 = (key * hx.c1) >> hx.hshift, s2 = (key * hx.c2) >> hx.hshift;
    const uint32_t k1 = hx.h[2 * s1], k2 = hx.h[2 * s2];
    if (k1 != key && k2 != key) return 0u;
    return hx.h[2 * (k1 == key ? s1 : s2) + 1] | 0x80000000u;
}
template <bool PADDED, class Sink>
DCB_HD void half_keywords(const ReadView& r, const uint32_t* inv2, const HalfView& hx, uint32_t meta, int P, Sink& sink) {
    const int first = (int)(meta & 255u), cnt = (int)((meta >> 8) & 255u);
    uint32_t lo, hi, ilo = 0, ihi = 0;
    rd_win32x<PADDED>(r, P, lo, hi);
    if (inv2) rd_win32x<PADDED>(half_inv_view(r, inv2), P, ilo, ihi);
    for (int i = 0; i < cnt; i++) {
        const int id = hx.ids[first + i];
        const DcbHalfKw k = hx.kw[id];
        const int len = k.len;
        if (P + len > r.n) continue;
        const uint32_t xlo = (lo ^ k.bits_lo) | ilo, xhi = (hi ^ k.bits_hi) | ihi;      // an occurrence needs valid bases
        if ((xlo & mask2(len)) | (len > 16 ? (xhi & mask2(len - 16)) : 0u)) continue;
        sink(id, (int)k.n_tags);
    }
}
template <bool PADDED, class Sink>
DCB_HD void half_lookup(const ReadView& r, const uint32_t* inv2, const HalfView& hx, int set, int P, Sink& sink) {
    const uint32_t meta = half_prefix<PADDED>(r, hx, set, P);
    if (meta) half_keywords<PADDED>(r, inv2, hx, meta, P, sink);
}
// One (occurrence, tag) pair: keyword record `id` occurs at P, `ti` counts the tags that have this half.  Appends the
// candidate with the reference's length guard (decombine.py:302-307) and lev.hamming(tag, window) <= 1 (:309) already
// evaluated.  A tag window that is not inside the read (the reference's slices then wrap or truncate) passes the read on.
template <bool PADDED>
DCB_HD void half_candidate(const ReadView& r, const uint32_t* inv2, const HalfView& hx, const DcbTag* vtags, const DcbTag* jtags,
                           int id, int ti, int P, uint32_t* cand, int cap, uint32_t* n) {
    const DcbHalfKw k = hx.kw[id];
    const int gene = k.set >> 1, half2 = k.set & 1;
    const int s0 = half2 ? P - (gene ? hx.j_split : hx.v_split) : P;   // where the whole tag would start (decombine.py:311, 361)
    const int kk = hx.tags[k.tags_off + ti];
    const DcbTag& t = (gene ? jtags : vtags)[kk];
    const int tlen = t.len;
    const int span = tlen > (int)k.first_len ? tlen : (int)k.first_len;
    if (s0 >= 0 && s0 + span > r.n && tlen == (int)k.first_len) {
        // the tag window runs past the end of the read: the reference's slice comes out short, the length guard fails
        // (:302-307) -- the keyword still DID occur
        (void)DCB_SLOT_OR(n, DCB_HALF_SEEN(gene, half2));
        return;
    }
    if (s0 < 0 || s0 + span > r.n) { (void)DCB_SLOT_TAKE(n, DCB_HALF_BAIL); return; }
    uint32_t lo, hi;
    rd_win32x<PADDED>(r, s0, lo, hi);
    lo ^= t.bits_lo; hi ^= t.bits_hi;
    if (inv2) {
        uint32_t ilo, ihi;
        rd_win32x<PADDED>(half_inv_view(r, inv2), s0, ilo, ihi);
        lo |= ilo; hi |= ihi;
    }
    lo &= t.mask_lo; hi &= t.mask_hi;
    const bool dead = tlen != (int)k.first_len || DCB_POPC((lo | (lo >> 1)) & 0x55555555u) + DCB_POPC((hi | (hi >> 1)) & 0x55555555u) > 1;
    // a tag that fails still says that its half keyword DID occur (half2 is only tried when half1 never hit)
    if (dead) (void)DCB_SLOT_OR(n, DCB_HALF_SEEN(gene, half2));
    else half_append(cand, r.stride, cap, n, DCB_HC_MAKE(gene, 1 + half2, P + (int)k.len, k.len, kk, 0));
}
// J half keywords below the sampled index's kmin (j_short: the 6-base halves of the 12-nt J tags) are found with a direct
// table over their first DCB_HALF_JQ bases, probed at EVERY base -- only in reads that need it: V assigned (a full V tag
// or a half-tag candidate that passed), J missing.  When the exact-tag kernel did not search J at all (full: it only
// does for reads with one full V tag), the full J tags are found here too: an occurrence starts with its first half.
// half_jshort_ok: does the read need the scan?  cand / n: the candidates so far (V side).
DCB_HD bool half_jshort_ok(const HalfView& hx, uint32_t hv, uint32_t hj, const uint32_t* cand, int stride, int cap, uint32_t n) {
    if (!hx.j_short || !(hj == 0u || hj == DCB_HIT_UNKNOWN) || n >= DCB_HALF_BAIL || DCB_HALF_COUNT(n) > (uint32_t)cap) return false;
    if (hv != 0u && hv != DCB_HIT_UNKNOWN) return true;
    for (uint32_t i = 0; i < DCB_HALF_COUNT(n); i++) {
        const uint32_t e = cand[i * stride];
        if (DCB_HC_GENE(e) == 0u) return true;
    }
    return false;          // no V: janalysis is never reached (decombine.py:542-545)
}
// A full tag starts with its first half: (occurrence of a half1 keyword at P, tag ti of that keyword) -> the whole tag
// compared, an occurrence appended as a candidate of kind 0.  fu: the genes whose full tags the exact-tag kernel did not
// search (bit 0 V, bit 1 J).
#define DCB_FU_OF(hv, hj) (((hv) == DCB_HIT_UNKNOWN ? 1u : 0u) | ((hj) == DCB_HIT_UNKNOWN ? 2u : 0u))
template <bool PADDED>
DCB_HD void half_full_candidate(const ReadView& r, const uint32_t* inv2, const HalfView& hx, const DcbTag* vtags, const DcbTag* jtags,
                                int id, int ti, int P, uint32_t fu, uint32_t* cand, int cap, uint32_t* n) {
    const DcbHalfKw k = hx.kw[id];
    if ((k.set & 1) || !((fu >> (k.set >> 1)) & 1u)) return;
    const int gene = k.set >> 1;
    const int kk = hx.tags[k.tags_off + ti];
    const DcbTag& t = (gene ? jtags : vtags)[kk];
    if (P + (int)t.len > r.n) return;
    uint32_t lo, hi;
    rd_win32x<PADDED>(r, P, lo, hi);
    lo ^= t.bits_lo; hi ^= t.bits_hi;
    if (inv2) {
        uint32_t ilo, ihi;
        rd_win32x<PADDED>(half_inv_view(r, inv2), P, ilo, ihi);
        lo |= ilo; hi |= ihi;
    }
    if (!((lo & t.mask_lo) | (hi & t.mask_hi))) half_append(cand, r.stride, cap, n, DCB_HC_MAKE(gene, 0, P + (int)t.len, t.len, kk, 0));
}
// One base of the scan, serially (tests/sim; the kernel pools the 6-mer hits of a warp): `six` = the DCB_HALF_JQ bases at P.
template <bool PADDED>
DCB_HD void half_jshort_at(const ReadView& r, const uint32_t* inv2, const HalfView& hx, const DcbTag* vtags, const DcbTag* jtags,
                           uint32_t six, int P, bool full, uint32_t* cand, int cap, uint32_t* n) {
    const uint32_t meta = hx.jt[six];
    if (!meta) return;
    struct Sink {
        const ReadView& r; const uint32_t* inv2; const HalfView& hx; const DcbTag* vtags; const DcbTag* jtags;
        int P; bool full; uint32_t* cand; int cap; uint32_t* n;
        DCB_HD void operator()(int id, int n_tags) {
            for (int ti = 0; ti < n_tags; ti++) {
                half_candidate<PADDED>(r, inv2, hx, vtags, jtags, id, ti, P, cand, cap, n);
                if (full) half_full_candidate<PADDED>(r, inv2, hx, vtags, jtags, id, ti, P, 2u, cand, cap, n);
            }
        }
    } sink{r, inv2, hx, vtags, jtags, P, full, cand, cap, n};
    half_keywords<PADDED>(r, inv2, hx, meta, P, sink);
}
// Set up the read: the invalid-base column of a flagged read (exception entries from e0 on), the hand-over words checked
// against it (a tag over a symbol packed as base 0 is no occurrence), which half sets have to be found.
// false: pass the read on (several full-tag candidates that cannot be told apart here, or a read too long).
DCB_HD bool half_begin(ReadView& r, const uint32_t*& inv2, bool flagged, const ExcList& ex, uint32_t e0, uint32_t* inv2col,
                       const DcbTag* vtags, const DcbTag* jtags, uint32_t& hv, uint32_t& hj, uint32_t& need, bool j_ok = true) {
    r.inv = nullptr; r.exc_pos = ex.pos; r.exc_kind = ex.kind; r.e0 = r.e1 = 0; r.mirror = 0;
    r.cand = nullptr; r.cand_kq = 0; r.hits = nullptr; r.n_hits = 0;
    inv2 = nullptr;
    need = 0;
    if (hv == DCB_HIT_MULTI || hj == DCB_HIT_MULTI || r.n > DCB_HALF_MAX_READ) return false;
    if (flagged) {
        uint32_t e1 = e0;
        bool any = false;
        for (int k = 0; k < r.nw; k++) inv2col[k * r.stride] = 0;
        for (; e1 < ex.n && ex.read[e1] == ex.read[e0]; e1++) {
            if (ex.kind[e1] == 3) continue;  // a real base in this frame
            const uint32_t p = ex.pos[e1];
            inv2col[(p >> 4) * r.stride] |= 1u << (2 * (p & 15));
            any = true;
        }
        r.e0 = (int)e0; r.e1 = (int)e1;
        if (any) inv2 = inv2col;
    }
    if (inv2) {
        const ReadView ri = half_inv_view(r, inv2);
        for (int g = 0; g < 2; g++) {
            uint32_t& h = g ? hj : hv;
            if (!h || h == DCB_HIT_UNKNOWN) continue;
            const int P = (int)(h & 0xFFFFu), L = (g ? jtags : vtags)[(h >> 16) & 0x7FFFu].len;
            uint32_t ilo, ihi;
            rd_win32(ri, P, ilo, ihi);
            if ((ilo & mask2(L)) | (L > 16 ? (ihi & mask2(L - 16)) : 0u)) h = 0u;
        }
    }
    need = ((hv && hv != DCB_HIT_UNKNOWN) ? 0u : 0x00FFu) | (((hj && hj != DCB_HIT_UNKNOWN) || !j_ok) ? 0u : 0xFF00u);
    return true;
}
// vanalysis / janalysis over the sorted candidates of one gene, from entry `i` on (decombine.py:273-394, 397-531): the
// candidate the reference would accept -- the full tag, else the first half1 candidate that passed the guard and
// Hamming <= 1, half2 candidates only when half1 never hit.  Returns its entry, or 0 with the failure counter pending.
// i is left at the first entry of the next gene.
template <bool IS_V>
DCB_HD uint32_t half_select(const uint32_t* cand, int stride, int n, int& i, uint32_t& pend, uint32_t seen) {
    const uint32_t gene = IS_V ? 0u : 1u;
    // the first keyword set that hit: full tags, else half1, else half2 -- by a candidate or by an occurrence that yielded none
    const uint32_t kind_seen = (seen & DCB_HALF_SEEN(gene, 0)) ? 1u : ((seen & DCB_HALF_SEEN(gene, 1)) ? 2u : 3u);
    const uint32_t kind_cand = (i < n && DCB_HC_GENE(cand[i * stride]) == gene) ? DCB_HC_KIND(cand[i * stride]) : 3u;
    const uint32_t kind0 = kind_cand < kind_seen ? kind_cand : kind_seen;
    if (kind0 == 3u) {
        pend |= 1u << (IS_V ? DCB_C_no_vtags_found : DCB_C_no_j_assigned);                  // :393 / :530
        return 0u;
    }
    uint32_t pick = 0u;
    int n_full = 0;
    for (; i < n; i++) {
        const uint32_t e = cand[i * stride];
        if (DCB_HC_GENE(e) != gene) break;
        if (!pick && DCB_HC_KIND(e) == kind0) pick = e;                                     // half2 only when half1 never hit (:339 / :473)
        n_full += DCB_HC_KIND(e) == 0u ? 1 : 0;
    }
    if (n_full > 1) {          // several occurrences of full tags (only the J scan of half_jshort_at lists them one by one): :278 / :402
        pend |= 1u << (IS_V ? DCB_C_multiple_v_matches : DCB_C_multiple_j_matches);
        return 0u;
    }
    if (!pick)  // :334 / :389 / :469 / :526 -- the J half2 failure bumps foundv2notv1 in the reference; preserved
        pend |= 1u << (IS_V ? (kind0 == 1u ? DCB_C_foundv1notv2 : DCB_C_foundv2notv1) : (kind0 == 1u ? DCB_C_foundj1notj2 : DCB_C_foundv2notv1));
    else if (kind0)
        pend |= 1u << (IS_V ? (kind0 == 1u ? DCB_C_verr2 : DCB_C_verr1) : (kind0 == 1u ? DCB_C_jerr2 : DCB_C_jerr1));  // :318 :370 :445 :504
    return pick;
}
// The deletion walk of the picked candidate (interior case).  false: pass the read on.
template <bool IS_V, bool PADDED>
DCB_HD bool half_walk(const ReadView& r, const uint32_t* inv2, const HalfView& hx, const DcbTag* tags, uint32_t e, int end_of_v, VJ& out) {
    const uint32_t kind = DCB_HC_KIND(e);
    const int split = IS_V ? hx.v_split : hx.j_split;
    const int kk = (int)DCB_HC_TAG(e), kwlen = (int)DCB_HC_KWLEN(e);
    const int P = (int)DCB_HC_END(e) - kwlen;                                               // start of the keyword occurrence
    const int s0 = kind == 2u ? P - split : P;                                              // start of the tag
    const DcbTagFin tf = tag_fin(tags, kk);
    const ReadView ri = half_inv_view(r, inv2);
    out.idx = kk;
    if (IS_V) {
        out.seqpos = s0;
        return fast_v_deletions<PADDED>(r, tf, s0 + tf.jump - 1, out.pos, out.dels, inv2 ? &ri : nullptr) != 0;
    }
    out.seqpos = kind == 1u ? P + 2 * split : s0 + (int)tf.len;                             // :450-454 (half1), :411 / :511
    return fast_j_deletions<PADDED>(r, tf, s0 - tf.jump, end_of_v, out.pos, out.dels, inv2 ? &ri : nullptr) != 0;
}
// The candidates are complete (n of them appended; more than cap, or DCB_HALF_BAIL added: pass the read on): add the
// full-tag occurrences, sort, run dcr() (decombine.py:534-585).  false: pass the read on to the general kernel (pend is
// then void).  Written as a sequence of short predicated steps so that the lanes of a warp walk it together.
template <bool PADDED>
DCB_HD bool half_run(const ReadView& r, const uint32_t* inv2, const HalfView& hx, const DcbTag* vtags, const DcbTag* jtags,
                     uint32_t hv, uint32_t hj, uint32_t* cand, int cap, uint32_t n, const DcrParams& prm, dcb_result& out, uint32_t& pend,
                     int* why = nullptr) {
    if (why) *why = n >= DCB_HALF_BAIL ? 2 : 3;
    const uint32_t seen = n;
    if (n >= DCB_HALF_BAIL) return false;
    n = DCB_HALF_COUNT(n);
    if (n > (uint32_t)cap) return false;
    for (int g = 0; g < 2; g++) {                                                           // the full-tag occurrences handed over
        const uint32_t h = g ? hj : hv;
        if (!h || h == DCB_HIT_UNKNOWN) continue;
        const int t = (int)((h >> 16) & 0x7FFFu), P = (int)(h & 0xFFFFu), L = (g ? jtags : vtags)[t].len;
        if (n < (uint32_t)cap) cand[n * r.stride] = DCB_HC_MAKE(g, 0, P + L, L, t, 0);
        n++;
    }
    if (n > (uint32_t)cap) return false;
    half_sort(cand, r.stride, (int)n);
    VJ v, j;
    v.idx = v.pos = v.dels = v.seqpos = 0;
    j = v;
    int i = 0;
    bool ok = true;
    const uint32_t ev = half_select<true>(cand, r.stride, (int)n, i, pend, seen);
    if (ev) ok = half_walk<true, PADDED>(r, inv2, hx, vtags, ev, 0, v);
    if (why) *why = 4;
    // V is assigned: now J.  A J gene that was not searched for full tags, or whose half tags this index does not hold
    // while the full tag is missing, is the general kernel's business.
    if (ev && ok && !hx.j_short && !hx.j_ok && (hj == DCB_HIT_UNKNOWN || hj == 0u)) return false;
    uint32_t ej = 0u;
    if (ev && ok) {                                                                         // :542-548
        ej = half_select<false>(cand, r.stride, (int)n, i, pend, seen);
        if (!ej) pend |= 1u << DCB_C_VJ_assignment_failed;                                  // :583-585
    }
    if (ej) ok = half_walk<false, PADDED>(r, inv2, hx, jtags, ej, v.pos + 1, j);
    if (why && ej) *why = 5;
    if (ej && ok) {
        const DcbTagFin vt = tag_fin(vtags, v.idx), jt = tag_fin(jtags, j.idx);
        const int c = dcr_finish(r, vt.jump, vt.len, jt.jump, jt.len, v, j, prm, out);
        if (c >= 0) pend |= 1u << c;
    }
    return ok;
}
DCB_HD void half_commit(uint32_t pend, dcb_cnt_t* C) {
    for (; pend; pend &= pend - 1) DCB_COUNT(C, DCB_FFS(pend) - 1);
}
// The whole path on one thread (tests/sim; the kernel probes from registers and pools the confirmations of a warp).
// false: pass the read on, nothing counted.
DCB_HD bool dcr_half_read(ReadView r, bool flagged, const ExcList& ex, uint32_t e0, uint32_t hv, uint32_t hj, uint32_t* inv2col,
                          uint32_t* cand, int cap, const uint32_t* vcore, const uint32_t* jcore, const uint32_t* hb,
                          const DcrParams& prm, dcb_result& out, dcb_cnt_t* C, int* why = nullptr) {
    const DcbTag* vtags = gene_tags(vcore);
    const DcbTag* jtags = gene_tags(jcore);
    const uint32_t* inv2;
    uint32_t need, pend = 0;
    if (why) *why = 1;
    const HalfView hx = half_view(hb);
    if (!half_begin(r, inv2, flagged, ex, e0, inv2col, vtags, jtags, hv, hj, need, hx.j_ok != 0)) return false;
    uint32_t n = 0;
    struct Sink {
        const ReadView& r; const uint32_t* inv2; const HalfView& hx; const DcbTag* vtags; const DcbTag* jtags;
        uint32_t* cand; int cap; uint32_t* n; int P;
        uint32_t fu;
        DCB_HD void operator()(int id, int n_tags) {
            for (int ti = 0; ti < n_tags; ti++) {
                half_candidate<false>(r, inv2, hx, vtags, jtags, id, ti, P, cand, cap, n);
                if (fu) half_full_candidate<false>(r, inv2, hx, vtags, jtags, id, ti, P, fu, cand, cap, n);
            }
        }
    } sink{r, inv2, hx, vtags, jtags, cand, cap, &n, 0, DCB_FU_OF(hv, hj)};
    if (need)
        for (int p = 0; p + DCB_HALF_Q <= r.n; p += DCB_HALF_STRIDE) {
            uint32_t e = hx.t[rd_win16(r, p) & mask2(DCB_HALF_Q)] & need;
            for (; e; e &= e - 1) {
                const int b = DCB_FFS(e) - 1;
                sink.P = p - (b & 3);
                half_lookup<false>(r, inv2, hx, b >> 2, sink.P, sink);
            }
        }
    if (half_jshort_ok(hx, hv, hj, cand, r.stride, cap, n))
        for (int p = 0; p + DCB_HALF_JQ <= r.n; p++)
            half_jshort_at<false>(r, inv2, hx, vtags, jtags, rd_win16(r, p) & mask2(DCB_HALF_JQ), p, hj == DCB_HIT_UNKNOWN, cand, cap, &n);
    dcb_result o;
    o.status = 0; o.frame = 0; o.v = o.j = 0; o.vdel = o.jdel = 0;
    o.ins_start = o.ins_end = o.v_seq_start = o.j_seq_end = 0;
    if (!half_run<false>(r, inv2, hx, vtags, jtags, hv, hj, cand, cap, n, prm, o, pend, why)) return false;
    out = o;
    half_commit(pend, C);
    return true;
}

// Both steps on one thread (tests/sim; the kernel regroups in between).
DCB_HD void dcr_general_read(ReadView r, uint32_t ri, bool flagged, const ExcList& ex, uint32_t* inv0, uint32_t* rd1,
                             uint32_t* inv1, const uint32_t* vblob, const uint32_t* jblob, const DcrParams& prm,
                             int both_frames, dcb_result& out, dcb_cnt_t* C, const uint32_t* sf = nullptr,
                             uint32_t* cand0 = nullptr, uint32_t* hits0 = nullptr) {
    dcr_general_prepare(r, ri, flagged, ex, inv0, vblob, jblob, sf, cand0, hits0);
    dcr_general_run(r, ex, rd1, inv1, vblob, jblob, prm, both_frames, out, C, sf, cand0, hits0);
}

#endif  // DCR_CORE_CUH
