// fastq.cpp -- record index of a FASTQ text held in memory (host side of the ingest).
//
// Replaces the per-record work of the reference's parser (readfq, decombine.py:228-265: a Python generator, 30 % of
// the reference's non-automaton time) for files in the layout sequencers write: four lines per record, '\n' line
// ends, ASCII.  The index is exactly what readfq would yield for such a file -- name = header after '@' up to the
// first SPACE (decombine.py:243), sequence and quality = the whole second and fourth line -- as (offset, length)
// pairs into the caller's buffer, so the sequences go to dcb_pack_reads without being copied.  Anything else
// (multi-line records, '>' records, '\r', a missing final newline, a quality shorter than its sequence, non-ASCII
// bytes, a partial last record) is reported as "not strict" and the caller uses the general parser, whose quirks
// (the last character of EVERY line is dropped, decombine.py:239-256) only matter there.
#include "dcb_internal.h"

#include <zlib.h>

#include <algorithm>
#include <cstdlib>
#include <chrono>
#include <sys/mman.h>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

template <class F>
void parallel_for(int n_threads, uint64_t n, F f) {
    if (n_threads <= 1 || n < 4096) { f(0, (uint64_t)0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++)
        th.emplace_back([=] { f(t, n * (uint64_t)t / n_threads, n * (uint64_t)(t + 1) / n_threads); });
    for (auto& x : th) x.join();
}

}  // namespace

namespace {
// A line of the text as the scan records it: where it starts (48 bits) | its first byte << 48 | the offset of its first
// SPACE << 56 (255: none within the first 255 bytes) -- all the record pass needs, so that it never touches the text again.
inline uint64_t line_entry(const unsigned char* p, uint64_t pos, uint64_t n_bytes) {
    return pos | ((uint64_t)(pos < n_bytes ? p[pos] : 0) << 48) | (255ull << 56);
}
inline void line_space(std::vector<uint64_t>& v, uint64_t at) {          // a space at `at`: the first of the current line?
    if (v.empty()) return;                                               // in a line that started in the chunk before
    uint64_t& e = v.back();
    if ((e >> 56) != 255) return;
    const uint64_t k = at - (e & 0xFFFFFFFFFFFFull);
    if (k < 255) e = (e & ~(255ull << 56)) | (k << 56);
}
#if defined(__x86_64__)
const bool kAvx2 = __builtin_cpu_supports("avx2");
// Lines of text[i .. b) in steps of 32 bytes (i is left at the first byte not looked at); returns non-zero when a
// '\r' or a non-ASCII byte was seen.
__attribute__((target("avx2"))) unsigned char scan_lines_avx2(const unsigned char* p, uint64_t& i, uint64_t b, uint64_t n_bytes, std::vector<uint64_t>& v) {
    const __m256i nl = _mm256_set1_epi8('\n'), cr = _mm256_set1_epi8('\r'), sp = _mm256_set1_epi8(' ');
    __m256i acc = _mm256_setzero_si256();
    for (; i + 32 <= b; i += 32) {
        const __m256i x = _mm256_loadu_si256((const __m256i*)(p + i));
        acc = _mm256_or_si256(acc, _mm256_or_si256(x, _mm256_cmpeq_epi8(x, cr)));
        const uint32_t m_nl = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, nl));
        const uint32_t m_sp = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, sp));
        for (uint32_t m = m_nl | m_sp; m; m &= m - 1u) {
            const uint32_t k = (uint32_t)__builtin_ctz(m);
            if ((m_nl >> k) & 1u) v.push_back(line_entry(p, i + k + 1, n_bytes));
            else line_space(v, i + k);
        }
    }
    return _mm256_movemask_epi8(acc) != 0 ? 1 : 0;
}
#endif
template <class F>
void run_threads(int nt, F f) {          // f(t) for t in [0, nt), each on a thread of its own
    if (nt <= 1) { f(0); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([=] { f(t); });
    for (auto& x : th) x.join();
}
}  // namespace

extern "C" {

int dcb_fastq_index_build(const char* text, uint64_t n_bytes, int n_threads, dcb_fastq_index** out) {
    if (!out || (n_bytes && !text)) { dcb_set_error("dcb_fastq_index_build: null argument"); return DCB_EINVAL; }
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    dcb_fastq_index* ix = (dcb_fastq_index*)std::calloc(1, sizeof(dcb_fastq_index));
    if (!ix) { dcb_set_error("dcb_fastq_index_build: out of memory"); return DCB_ENOMEM; }
    *out = ix;
    ix->strict = 0;
    if (n_bytes == 0 || text[n_bytes - 1] != '\n') return DCB_OK;          // empty, or the last line is not terminated

    const auto T0 = std::chrono::steady_clock::now();
    // pass 1 (the only one over the whole text): every thread lists the line starts of its chunk and checks its bytes --
    // a '\r' or a non-ASCII byte anywhere ends the strict path
    std::vector<std::vector<uint64_t>> found(n_threads);
    std::vector<int> bad(n_threads, 0);
    const int nt1 = (n_bytes < 4096) ? 1 : n_threads;
    parallel_for(nt1, n_bytes, [&](int t, uint64_t a, uint64_t b) {
        std::vector<uint64_t>& v = found[t];
        v.reserve((size_t)((b - a) / 48 + 16));
        const unsigned char* p = (const unsigned char*)text;
        unsigned char any = 0;
        uint64_t i = a;
        if (a == 0) v.push_back(line_entry(p, 0, n_bytes));                  // the first line starts the text
#if defined(__x86_64__)
        if (kAvx2) { any = scan_lines_avx2(p, i, b, n_bytes, v); }
#endif
        for (; i < b; i++) {
            const unsigned char ch = p[i];
            any |= (unsigned char)((ch & 0x80) | (ch == '\r' ? 0x80 : 0));
            if (ch == '\n') v.push_back(line_entry(p, i + 1, n_bytes));
            else if (ch == ' ') line_space(v, i);
        }
        // the first space of the chunk's last line may lie in the next chunk
        if (!v.empty() && (v.back() >> 56) == 255)
            for (uint64_t k = b; k < n_bytes && p[k] != '\n' && k - (v.back() & 0xFFFFFFFFFFFFull) < 255; k++)
                if (p[k] == ' ') { line_space(v, k); break; }
        bad[t] = any != 0;
    });
    const auto T1 = std::chrono::steady_clock::now();
    std::vector<uint64_t> nl(n_threads + 1, 0);
    for (int t = 0; t < n_threads; t++) { if (bad[t]) return DCB_OK; nl[t + 1] = nl[t] + found[t].size(); }
    const uint64_t n_lines = nl[n_threads] - 1;                             // entries: every line start + the end of the text
    if (n_bytes >= (1ull << 48)) return DCB_OK;
    if (n_lines == 0 || n_lines % 4 != 0) return DCB_OK;
    const uint64_t n = n_lines / 4;
    if (n >= 0xFFFFFFFFull) return DCB_OK;

    // where every line starts (line k + 1 starts behind the k-th newline): the threads' lists, end to end
    uint64_t* start = (uint64_t*)std::malloc(sizeof(uint64_t) * (n_lines + 1));
    if (!start) { dcb_set_error("dcb_fastq_index_build: out of memory"); return DCB_ENOMEM; }
    {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++)
            if (!found[t].empty())
                th.emplace_back([&, t] { std::memcpy(start + nl[t], found[t].data(), sizeof(uint64_t) * found[t].size()); });
        for (auto& x : th) x.join();
    }
    found.clear();
    const auto T2 = std::chrono::steady_clock::now();

    // pass 3: the records
    ix->name_off = (uint64_t*)std::malloc(sizeof(uint64_t) * n); ix->name_len = (uint32_t*)std::malloc(sizeof(uint32_t) * n);
    ix->seq_off = (uint64_t*)std::malloc(sizeof(uint64_t) * n);  ix->seq_len = (uint32_t*)std::malloc(sizeof(uint32_t) * n);
    ix->qual_off = (uint64_t*)std::malloc(sizeof(uint64_t) * n); ix->qual_len = (uint32_t*)std::malloc(sizeof(uint32_t) * n);
    if (!ix->name_off || !ix->name_len || !ix->seq_off || !ix->seq_len || !ix->qual_off || !ix->qual_len) {
        std::free(start);
        dcb_set_error("dcb_fastq_index_build: out of memory");
        return DCB_ENOMEM;
    }
    std::vector<int> irregular(n_threads, 0);
    parallel_for(n_threads, n, [&](int t, uint64_t a, uint64_t b) {
        for (uint64_t r = a; r < b; r++) {
            const uint64_t M = 0xFFFFFFFFFFFFull;
            const uint64_t eh = start[4 * r], es = start[4 * r + 1], ep = start[4 * r + 2];
            const uint64_t h = eh & M, s = es & M, p = ep & M, q = start[4 * r + 3] & M, e = start[4 * r + 4] & M;
            const uint64_t hl = s - h - 1, sl = p - s - 1, ql = e - q - 1;            // line lengths without the '\n'
            const unsigned char ch = (unsigned char)(eh >> 48), cs = (unsigned char)(es >> 48), cp = (unsigned char)(ep >> 48);
            if (ch != '@' || cp != '+') { irregular[t] = 1; return; }
            if (sl > 0 && (cs == '@' || cs == '+' || cs == '>')) { irregular[t] = 1; return; }   // a marker to readfq
            if (ql < sl || sl > 0xFFFFFFu || ql > 0xFFFFFFu) { irregular[t] = 1; return; }     // multi-line quality / absurd
            uint64_t name_len = hl == 0 ? 0 : hl - 1;                                 // up to the first space behind the '@'
            const uint64_t k = eh >> 56;
            if (k != 255) { if (k >= 1 && k <= hl) name_len = k - 1; }
            else if (hl > 255) {
                const char* sp = (const char*)std::memchr(text + h + 255, ' ', (size_t)(hl - 255));
                if (sp) name_len = (uint64_t)(sp - (text + h + 1));
            }
            ix->name_off[r] = h + 1;
            ix->name_len[r] = (uint32_t)name_len;
            ix->seq_off[r] = s; ix->seq_len[r] = (uint32_t)sl;
            ix->qual_off[r] = q; ix->qual_len[r] = (uint32_t)ql;
        }
    });
    std::free(start);
    if (std::getenv("DCB_TIMING")) {
        const auto T3 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "\t[timing] fastq index: scan %.3f s, line starts %.3f s, records %.3f s (%d threads)\n", std::chrono::duration<double>(T1 - T0).count(),
                     std::chrono::duration<double>(T2 - T1).count(), std::chrono::duration<double>(T3 - T2).count(), n_threads);
    }
    for (int t = 0; t < n_threads; t++) if (irregular[t]) return DCB_OK;
    ix->n_records = n;
    ix->strict = 1;
    return DCB_OK;
}

void dcb_fastq_index_free(dcb_fastq_index* ix) {
    if (!ix) return;
    std::free(ix->name_off); std::free(ix->name_len); std::free(ix->seq_off); std::free(ix->seq_len);
    std::free(ix->qual_off); std::free(ix->qual_len);
    std::free(ix);
}

/* Number of the n byte ranges (off[i], len[i]) of text that contain the byte `symbol`. */
uint64_t dcb_count_ranges_with(const char* text, const uint64_t* off, const uint32_t* len, uint64_t n, int symbol, int n_threads) {
    if (!text || !off || !len || n == 0) return 0;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    std::vector<uint64_t> part(n_threads, 0);
    parallel_for(n_threads, n, [&](int t, uint64_t a, uint64_t b) {
        uint64_t c = 0;
        for (uint64_t i = a; i < b; i++) c += len[i] && std::memchr(text + off[i], symbol, len[i]) != nullptr;
        part[t] += c;
    });
    uint64_t total = 0;
    for (auto v : part) total += v;
    return total;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Row assembly: the ten strings the reference builds per decombined read
// (decombine.py:1015-1039) -- v, j, vdel, jdel, insert, read id, tcrseq, tcrQ, barcode, barcode quality
// (+ the sampling column) -- formatted for ALL hits of a batch into one text buffer, fields joined by `sep`,
// one row per line.  With sep = ", " the buffer IS the .n12 text of write_out_intermediate (io.py:480-513);
// with a separator that cannot occur in the data, Python splits it into the list of rows at C speed.
// ------------------------------------------------------------------------------------------------
namespace {

struct Comp {
    unsigned char t[256];
    Comp() {   // Bio.Seq.reverse_complement's table (decombine.py:182-184), as in pack.cpp
        const char* from = "ACGTUMRWSYKVHDBNacgtumrwsykvhdbn";
        const char* to = "TGCAAKYWSRMBDHVNtgcaakywsrmbdhvn";
        for (int i = 0; i < 256; i++) t[i] = (unsigned char)i;
        for (int i = 0; from[i]; i++) t[(unsigned char)from[i]] = (unsigned char)to[i];
    }
};
const Comp kComp;

// dst[k] = src[n - 1 - k] for k in [a, b): a reversed slice, 16 bytes at a time where the CPU has SSSE3
#if defined(__x86_64__)
#define DCB_HAVE_SSSE3_PATH 1
const bool kSsse3 = __builtin_cpu_supports("ssse3");
__attribute__((target("ssse3"))) inline void rev_copy_ssse3(char* dst, const char* src, int64_t n, int64_t a, int64_t b) {
    const __m128i flip = _mm_set_epi8(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    int64_t k = a;
    for (; k + 16 <= b; k += 16)         // dst[k .. k+16) = reversed src[n-16-k .. n-k)
        _mm_storeu_si128((__m128i*)(dst + (k - a)), _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(src + n - 16 - k)), flip));
    for (; k < b; k++) dst[k - a] = src[n - 1 - k];
}
// The same with the complement of A C G T N; false (nothing usable written) when another symbol turns up: the caller
// then takes the table (Bio.Seq's whole IUPAC alphabet, lower case).
__attribute__((target("ssse3"))) inline bool revcomp_copy_ssse3(char* dst, const unsigned char* src, int64_t n, int64_t a, int64_t b) {
    const __m128i flip = _mm_set_epi8(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    // by low nibble: 'A' 0x41 -> 1, 'C' 0x43 -> 3, 'T' 0x54 -> 4, 'G' 0x47 -> 7, 'N' 0x4E -> 14
    const __m128i comp = _mm_setr_epi8(0, 'T', 0, 'G', 'A', 0, 0, 'C', 0, 0, 0, 0, 0, 0, 'N', 0);
    const __m128i self = _mm_setr_epi8(-1, 'A', -1, 'C', 'T', -1, -1, 'G', -1, -1, -1, -1, -1, -1, 'N', -1);
    const __m128i low = _mm_set1_epi8(0x0F);
    int64_t k = a;
    int ok = 0xFFFF;
    for (; k + 16 <= b; k += 16) {
        const __m128i v = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(src + n - 16 - k)), flip);
        const __m128i nib = _mm_and_si128(v, low);
        ok &= _mm_movemask_epi8(_mm_cmpeq_epi8(_mm_shuffle_epi8(self, nib), v));
        _mm_storeu_si128((__m128i*)(dst + (k - a)), _mm_shuffle_epi8(comp, nib));
    }
    if (ok != 0xFFFF) return false;
    for (; k < b; k++) {
        const unsigned char ch = src[n - 1 - k];
        char o;
        switch (ch) { case 'A': o = 'T'; break; case 'C': o = 'G'; break; case 'G': o = 'C'; break; case 'T': o = 'A'; break; case 'N': o = 'N'; break; default: return false; }
        dst[k - a] = o;
    }
    return true;
}
#endif

inline void clip(int64_t len, int64_t& a, int64_t& b) {          // Python s[a:b] for a, b >= 0
    if (a > len) a = len;
    if (b > len) b = len;
    if (b < a) b = a;
}
inline int dec_len(unsigned v) { return v >= 10000 ? 5 : v >= 1000 ? 4 : v >= 100 ? 3 : v >= 10 ? 2 : 1; }
inline char* put_dec(char* p, unsigned v) {
    const int n = dec_len(v);
    for (int i = n - 1; i >= 0; i--) { p[i] = (char)('0' + v % 10); v /= 10; }
    return p + n;
}

struct RowCtx {
    const dcb_result* res; int packed_rc;
    const dcb_column *ids, *vdj, *qual, *bc, *bcq, *tail;
    const char* sep; size_t sep_len;
    int mode;        // 0: the row, fields joined by sep; 1: what collapse builds per row, three lines: tcrseq, str(row[:5]),
                     //    "|".join((str(row[:5]), tcrseq, tcrQ, read id)) (collapse.py:560-590)
};

// length of the row of read i (it decombined), or its bytes when dst != nullptr
inline size_t row_emit(const RowCtx& c, uint64_t i, char* dst) {
    const dcb_result& r = c.res[i];
    const bool rev = (c.packed_rc != 0) != (r.frame != 0);
    const int64_t n = c.vdj->len[i], nq = c.qual->len[i];
    int64_t ia = r.ins_start, ib = r.ins_end, sa = r.v_seq_start, sb = r.j_seq_end, qa = r.v_seq_start, qb = r.j_seq_end;
    clip(n, ia, ib); clip(n, sa, sb); clip(nq, qa, qb);
    const size_t fixed = (size_t)dec_len(r.v) + dec_len(r.j) + dec_len(r.vdel) + dec_len(r.jdel);
    if (c.mode == 1) {
        const size_t dcr = fixed + (size_t)(ib - ia) + 20, sq = (size_t)(sb - sa);     // ['v', 'j', 'vdel', 'jdel', 'insert']
        const size_t total = sq + 1 + dcr + 1 + dcr + 1 + sq + 1 + (size_t)(qb - qa) + 1 + c.ids->len[i] + 1;
        if (!dst) return total;
        char* p = dst;
        const unsigned char* s = (const unsigned char*)c.vdj->text + c.vdj->off[i];
        const char* q = c.qual->text + c.qual->off[i];
        auto seq = [&](int64_t a, int64_t b) {
            if (rev) {
#ifdef DCB_HAVE_SSSE3_PATH
                if (kSsse3 && b - a >= 16 && revcomp_copy_ssse3(p, s, n, a, b)) { p += b - a; return; }
#endif
                for (int64_t k = a; k < b; k++) *p++ = (char)kComp.t[s[n - 1 - k]];
            } else { std::memcpy(p, s + a, (size_t)(b - a)); p += b - a; }
        };
        auto lit = [&](const char* t, size_t l) { std::memcpy(p, t, l); p += l; };
        char* const seq_at = p;
        seq(sa, sb); *p++ = '\n';
        char* const dcr_at = p;
        lit("['", 2); p = put_dec(p, r.v); lit("', '", 4); p = put_dec(p, r.j); lit("', '", 4); p = put_dec(p, r.vdel); lit("', '", 4);
        p = put_dec(p, r.jdel); lit("', '", 4); seq(ia, ib); lit("']", 2);
        *p++ = '\n';
        std::memcpy(p, dcr_at, dcr); p += dcr; *p++ = '|';
        std::memcpy(p, seq_at, sq); p += sq; *p++ = '|';
        if (rev) {
#ifdef DCB_HAVE_SSSE3_PATH
            if (kSsse3) { rev_copy_ssse3(p, q, nq, qa, qb); p += qb - qa; }
            else
#endif
            for (int64_t k = qa; k < qb; k++) *p++ = q[nq - 1 - k];
        } else { std::memcpy(p, q + qa, (size_t)(qb - qa)); p += qb - qa; }
        *p++ = '|';
        std::memcpy(p, c.ids->text + c.ids->off[i], c.ids->len[i]); p += c.ids->len[i];
        *p++ = '\n';
        return (size_t)(p - dst);
    }
    const int nf = c.tail ? 11 : 10;
    const size_t total = fixed + (size_t)(ib - ia) + c.ids->len[i] + (size_t)(sb - sa) + (size_t)(qb - qa) + c.bc->len[i] +
                         c.bcq->len[i] + (c.tail ? c.tail->len[i] : 0) + (size_t)(nf - 1) * c.sep_len + 1;
    if (!dst) return total;
    char* p = dst;
    auto sep = [&] { std::memcpy(p, c.sep, c.sep_len); p += c.sep_len; };
    auto raw = [&](const dcb_column* col) { std::memcpy(p, col->text + col->off[i], col->len[i]); p += col->len[i]; };
    const unsigned char* s = (const unsigned char*)c.vdj->text + c.vdj->off[i];
    const char* q = c.qual->text + c.qual->off[i];
    auto seq = [&](int64_t a, int64_t b) {                       // oriented[a:b]
        if (rev) {
#ifdef DCB_HAVE_SSSE3_PATH
            if (kSsse3 && b - a >= 16 && revcomp_copy_ssse3(p, s, n, a, b)) { p += b - a; return; }
#endif
            for (int64_t k = a; k < b; k++) *p++ = (char)kComp.t[s[n - 1 - k]];
        } else { std::memcpy(p, s + a, (size_t)(b - a)); p += b - a; }
    };
    p = put_dec(p, r.v); sep(); p = put_dec(p, r.j); sep(); p = put_dec(p, r.vdel); sep(); p = put_dec(p, r.jdel); sep();
    seq(ia, ib); sep();
    raw(c.ids); sep();
    seq(sa, sb); sep();
    if (rev) {
#ifdef DCB_HAVE_SSSE3_PATH
        if (kSsse3) { rev_copy_ssse3(p, q, nq, qa, qb); p += qb - qa; }
        else
#endif
        for (int64_t k = qa; k < qb; k++) *p++ = q[nq - 1 - k];
    } else { std::memcpy(p, q + qa, (size_t)(qb - qa)); p += qb - qa; }
    sep();
    raw(c.bc); sep();
    raw(c.bcq);
    if (c.tail) { sep(); raw(c.tail); }
    *p++ = '\n';
    return (size_t)(p - dst);
}

}  // namespace

extern "C" {

static int format_impl(const RowCtx& c, const dcb_result* res, uint64_t n, int n_threads, char** out, uint64_t* out_bytes, uint64_t* n_rows);

int dcb_format_rows(const dcb_result* res, uint64_t n, int packed_revcomp, const dcb_column* ids, const dcb_column* vdj,
                    const dcb_column* vdjqual, const dcb_column* bc, const dcb_column* bcq, const dcb_column* v_tail,
                    const char* sep, int n_threads, char** out, uint64_t* out_bytes, uint64_t* n_rows) {
    if (!out || !out_bytes || !n_rows || !sep || (n && (!res || !ids || !vdj || !vdjqual || !bc || !bcq))) {
        dcb_set_error("dcb_format_rows: null argument");
        return DCB_EINVAL;
    }
    RowCtx c;
    c.res = res; c.packed_rc = packed_revcomp; c.ids = ids; c.vdj = vdj; c.qual = vdjqual; c.bc = bc; c.bcq = bcq; c.tail = v_tail;
    c.sep = sep; c.sep_len = std::strlen(sep); c.mode = 0;
    return format_impl(c, res, n, n_threads, out, out_bytes, n_rows);
}

int dcb_format_collapse_rows(const dcb_result* res, uint64_t n, int packed_revcomp, const dcb_column* ids, const dcb_column* vdj,
                             const dcb_column* vdjqual, int n_threads, char** out, uint64_t* out_bytes, uint64_t* n_rows) {
    if (!out || !out_bytes || !n_rows || (n && (!res || !ids || !vdj || !vdjqual))) {
        dcb_set_error("dcb_format_collapse_rows: null argument");
        return DCB_EINVAL;
    }
    RowCtx c;
    c.res = res; c.packed_rc = packed_revcomp; c.ids = ids; c.vdj = vdj; c.qual = vdjqual; c.bc = nullptr; c.bcq = nullptr; c.tail = nullptr;
    c.sep = ""; c.sep_len = 0; c.mode = 1;
    return format_impl(c, res, n, n_threads, out, out_bytes, n_rows);
}

static int format_impl(const RowCtx& c, const dcb_result* res, uint64_t n, int n_threads, char** out, uint64_t* out_bytes, uint64_t* n_rows) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    const auto T0 = std::chrono::steady_clock::now();
    // contiguous read ranges per thread: sizes, then bytes, rows staying in read order
    std::vector<uint64_t> bytes(n_threads + 1, 0), rows(n_threads + 1, 0);
    const int nt = (n < 4096) ? 1 : n_threads;
    auto range = [&](int t, uint64_t& a, uint64_t& b) { a = n * (uint64_t)t / nt; b = n * (uint64_t)(t + 1) / nt; };
    {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++)
            th.emplace_back([&, t] {
                uint64_t a, b, sz = 0, nr = 0;
                range(t, a, b);
                for (uint64_t i = a; i < b; i++) if (res[i].status) { sz += row_emit(c, i, nullptr); nr++; }
                bytes[t + 1] = sz; rows[t + 1] = nr;
            });
        for (auto& x : th) x.join();
    }
    for (int t = 0; t < nt; t++) { bytes[t + 1] += bytes[t]; rows[t + 1] += rows[t]; }
    const auto T1 = std::chrono::steady_clock::now();
    // a large buffer is first touched here: huge pages (where the system hands them out) cut the page faults 512-fold
    char* buf = nullptr;
    if (bytes[nt] >= ((size_t)8 << 20) && !std::getenv("DCB_NO_HUGE")) {
        const size_t huge = (size_t)2 << 20, want = (bytes[nt] + 1 + huge - 1) / huge * huge;
        buf = (char*)std::aligned_alloc(huge, want);
#ifdef MADV_HUGEPAGE
        if (buf) madvise(buf, want, MADV_HUGEPAGE);
#endif
    }
    if (!buf) buf = (char*)std::malloc(bytes[nt] + 1);
    if (!buf) { dcb_set_error("dcb_format_rows: out of memory"); return DCB_ENOMEM; }
    {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++)
            th.emplace_back([&, t] {
                uint64_t a, b;
                range(t, a, b);
                char* p = buf + bytes[t];
                for (uint64_t i = a; i < b; i++) if (res[i].status) p += row_emit(c, i, p);
            });
        for (auto& x : th) x.join();
    }
    buf[bytes[nt]] = 0;
    if (std::getenv("DCB_TIMING")) {
        const auto T2 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "\t[timing] format_rows: sizes %.3f s, emit %.3f s (%.2f GB)\n", std::chrono::duration<double>(T1 - T0).count(),
                     std::chrono::duration<double>(T2 - T1).count(), bytes[nt] / 1e9);
    }
    *out = buf; *out_bytes = bytes[nt]; *n_rows = rows[nt];
    return DCB_OK;
}

void dcb_buffer_free(char* p) { std::free(p); }

// ------------------------------------------------------------------------------------------------
// BGZF blocks (gzip members of at most 64 KB that carry their own size: bgzip, Illumina's converters) inflated side by side:
// block k is raw deflate data raw[start[k], end[k]) of isize[k] bytes, followed by its CRC-32, and goes to out + out_off[k].
// ------------------------------------------------------------------------------------------------
int dcb_bgzf_inflate(const unsigned char* raw, const uint64_t* start, const uint64_t* end, const uint32_t* isize, const uint64_t* out_off,
                     uint64_t n_blocks, unsigned char* out, int n_threads) {
    if (n_blocks && (!raw || !start || !end || !isize || !out_off || !out)) { dcb_set_error("dcb_bgzf_inflate: null argument"); return DCB_EINVAL; }
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    if ((uint64_t)n_threads > n_blocks) n_threads = n_blocks ? (int)n_blocks : 1;
    std::vector<long long> bad(n_threads, -1);
    run_threads(n_threads, [&](int t) {
        z_stream z;
        std::memset(&z, 0, sizeof(z));
        if (inflateInit2(&z, -15) != Z_OK) { bad[t] = (long long)(n_blocks * (uint64_t)t / n_threads); return; }
        const uint64_t lo = n_blocks * (uint64_t)t / n_threads, hi = n_blocks * (uint64_t)(t + 1) / n_threads;
        for (uint64_t k = lo; k < hi; k++) {
            inflateReset(&z);
            z.next_in = const_cast<unsigned char*>(raw + start[k]); z.avail_in = (uInt)(end[k] - start[k]);
            z.next_out = out + out_off[k]; z.avail_out = isize[k];
            const int rc = inflate(&z, Z_FINISH);
            uint32_t want;
            std::memcpy(&want, raw + end[k], 4);
            if (rc != Z_STREAM_END || z.avail_out != 0 || (uint32_t)crc32(0L, out + out_off[k], isize[k]) != want) { bad[t] = (long long)k; break; }
        }
        inflateEnd(&z);
    });
    for (int t = 0; t < n_threads; t++)
        if (bad[t] >= 0) { dcb_set_error("dcb_bgzf_inflate: block %lld does not inflate to its size and CRC", bad[t]); return DCB_EINVAL; }
    return DCB_OK;
}

// ------------------------------------------------------------------------------------------------
// .n12 text (the `collapse` command's input, written by write_out_intermediate, io.py:480-513): where the ten fields
// of every row are, and the strings collapse files a row under, built from those fields.
// ------------------------------------------------------------------------------------------------
/* Index of an .n12 text: rows of exactly ten fields joined by ", ", every row ended by '\n'.  off / len: n_rows x 10,
 * row-major, malloc'ed (free with dcb_buffer_free).  *n_rows = 0 with DCB_OK when the text is not of that shape (a row
 * with another number of fields, a missing final newline, a '\r'): the caller splits the lines itself. */
int dcb_n12_index(const char* text, uint64_t n_bytes, int n_threads, uint64_t** off_out, uint32_t** len_out, uint64_t* n_rows) {
    if (!off_out || !len_out || !n_rows || (n_bytes && !text)) { dcb_set_error("dcb_n12_index: null argument"); return DCB_EINVAL; }
    *off_out = nullptr; *len_out = nullptr; *n_rows = 0;
    if (n_bytes == 0 || text[n_bytes - 1] != '\n') return DCB_OK;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    if (n_bytes < 65536) n_threads = 1;
    // thread t owns the rows that START in its byte range; a range begins at the first row start at or behind its nominal start
    std::vector<uint64_t> cut(n_threads + 1, n_bytes);
    cut[0] = 0;
    for (int t = 1; t < n_threads; t++) {
        uint64_t p = n_bytes * (uint64_t)t / n_threads;
        const char* q = (const char*)std::memchr(text + p, '\n', (size_t)(n_bytes - p));
        cut[t] = q ? (uint64_t)(q - text) + 1 : n_bytes;
    }
    std::vector<std::vector<uint64_t>> offs(n_threads);
    std::vector<std::vector<uint32_t>> lens(n_threads);
    std::vector<int> bad(n_threads, 0);
    run_threads(n_threads, [&](int t) {
        {
            std::vector<uint64_t>& o = offs[t];
            std::vector<uint32_t>& l = lens[t];
            uint64_t p = cut[t];
            const uint64_t end = cut[t + 1];
            o.reserve((size_t)((end - p) / 400 + 16) * 10); l.reserve((size_t)((end - p) / 400 + 16) * 10);
            while (p < end) {
                const char* nl = (const char*)std::memchr(text + p, '\n', (size_t)(n_bytes - p));
                const uint64_t e = (uint64_t)(nl - text);                       // the text ends with '\n': nl is never null
                int nf = 0;
                uint64_t f = p;
                for (uint64_t i = p; i + 1 < e && nf < 11; i++)
                    if (text[i] == ',' && text[i + 1] == ' ') { if (nf < 10) { o.push_back(f); l.push_back((uint32_t)(i - f)); } nf++; f = i + 2; i++; }
                if (nf != 9 || e - f > 0xFFFFFFFFull || std::memchr(text + p, '\r', (size_t)(e - p))) { bad[t] = 1; return; }
                o.push_back(f); l.push_back((uint32_t)(e - f));
                p = e + 1;
            }
        }
    });
    uint64_t total = 0;
    for (int t = 0; t < n_threads; t++) { if (bad[t]) return DCB_OK; total += offs[t].size(); }
    if (total == 0 || total % 10) return DCB_OK;
    uint64_t* O = (uint64_t*)std::malloc(sizeof(uint64_t) * total);
    uint32_t* Ln = (uint32_t*)std::malloc(sizeof(uint32_t) * total);
    if (!O || !Ln) { std::free(O); std::free(Ln); dcb_set_error("dcb_n12_index: out of memory"); return DCB_ENOMEM; }
    uint64_t at = 0;
    for (int t = 0; t < n_threads; t++) {
        if (!offs[t].empty()) {
            std::memcpy(O + at, offs[t].data(), sizeof(uint64_t) * offs[t].size());
            std::memcpy(Ln + at, lens[t].data(), sizeof(uint32_t) * lens[t].size());
        }
        at += offs[t].size();
    }
    *off_out = O; *len_out = Ln; *n_rows = total / 10;
    return DCB_OK;
}

/* The same three lines per row as dcb_format_collapse_rows, for the rows of an .n12 text selected by keep[i] != 0:
 * tcrseq (field 6); "['f0', 'f1', 'f2', 'f3', 'f4']" (str(row[:5]) -- fields with a quote or a backslash cannot be
 * written that way: *n_rows = UINT64_MAX, the caller builds the strings itself); that + "|" + field 6 + "|" + field 7 + "|" + field 5. */
int dcb_n12_collapse_rows(const char* text, const uint64_t* off, const uint32_t* len, uint64_t n, const uint8_t* keep, int n_threads,
                          char** out, uint64_t* out_bytes, uint64_t* n_rows) {
    if (!out || !out_bytes || !n_rows || (n && (!text || !off || !len || !keep))) { dcb_set_error("dcb_n12_collapse_rows: null argument"); return DCB_EINVAL; }
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    const int nt = n < 4096 ? 1 : n_threads;
    std::vector<uint64_t> bytes(nt + 1, 0), rows(nt + 1, 0);
    std::vector<int> odd(nt, 0);
    auto range = [&](int t, uint64_t& a, uint64_t& b) { a = n * (uint64_t)t / nt; b = n * (uint64_t)(t + 1) / nt; };
    auto dcr_len = [&](uint64_t i) { size_t s = 20; for (int k = 0; k < 5; k++) s += len[10 * i + k]; return s; };
    run_threads(nt, [&](int t) {
        {
            uint64_t a, b, sz = 0, nr = 0;
            range(t, a, b);
            for (uint64_t i = a; i < b; i++) {
                if (!keep[i]) continue;
                for (int k = 0; k < 5; k++) {
                    const char* f = text + off[10 * i + k];
                    for (uint32_t c = 0; c < len[10 * i + k]; c++) if (f[c] == '\'' || f[c] == '\\' || f[c] == '\n' || (unsigned char)f[c] < 32) odd[t] = 1;
                }
                const size_t d = dcr_len(i), sq = len[10 * i + 6];
                sz += sq + 1 + d + 1 + d + 1 + sq + 1 + len[10 * i + 7] + 1 + len[10 * i + 5] + 1;
                nr++;
            }
            bytes[t + 1] = sz; rows[t + 1] = nr;
        }
    });
    for (int t = 0; t < nt; t++) { if (odd[t]) { *out = nullptr; *out_bytes = 0; *n_rows = ~0ull; return DCB_OK; } bytes[t + 1] += bytes[t]; rows[t + 1] += rows[t]; }
    char* buf = (char*)std::malloc(bytes[nt] + 1);
    if (!buf) { dcb_set_error("dcb_n12_collapse_rows: out of memory"); return DCB_ENOMEM; }
    run_threads(nt, [&](int t) {
        {
            uint64_t a, b;
            range(t, a, b);
            char* p = buf + bytes[t];
            for (uint64_t i = a; i < b; i++) {
                if (!keep[i]) continue;
                auto field = [&](int k) { std::memcpy(p, text + off[10 * i + k], len[10 * i + k]); p += len[10 * i + k]; };
                auto lit = [&](const char* s, size_t l) { std::memcpy(p, s, l); p += l; };
                field(6); *p++ = '\n';
                char* const dcr_at = p;
                lit("['", 2);
                for (int k = 0; k < 5; k++) { field(k); if (k < 4) lit("', '", 4); }
                lit("']", 2);
                const size_t d = (size_t)(p - dcr_at);
                *p++ = '\n';
                std::memcpy(p, dcr_at, d); p += d; *p++ = '|';
                field(6); *p++ = '|'; field(7); *p++ = '|'; field(5); *p++ = '\n';
            }
        }
    });
    buf[bytes[nt]] = 0;
    *out = buf; *out_bytes = bytes[nt]; *n_rows = rows[nt];
    return DCB_OK;
}

}  // extern "C"
