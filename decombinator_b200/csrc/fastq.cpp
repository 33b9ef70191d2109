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

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

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

    // pass 1: newlines per chunk; '\r' or a non-ASCII byte anywhere ends the strict path
    std::vector<uint64_t> nl(n_threads + 1, 0);
    std::vector<int> bad(n_threads, 0);
    parallel_for(n_threads, n_bytes, [&](int t, uint64_t a, uint64_t b) {
        uint64_t c = 0;
        unsigned char any = 0;
        const unsigned char* p = (const unsigned char*)text;
        for (uint64_t i = a; i < b; i++) { c += p[i] == '\n'; any |= (unsigned char)((p[i] & 0x80) | (p[i] == '\r' ? 0x80 : 0)); }
        nl[t + 1] = c; bad[t] = any != 0;
    });
    for (int t = 0; t < n_threads; t++) { if (bad[t]) return DCB_OK; nl[t + 1] += nl[t]; }
    const uint64_t n_lines = nl[n_threads];
    if (n_lines == 0 || n_lines % 4 != 0) return DCB_OK;
    const uint64_t n = n_lines / 4;
    if (n >= 0xFFFFFFFFull) return DCB_OK;

    // pass 2: where every line starts (line k + 1 starts behind the k-th newline)
    uint64_t* start = (uint64_t*)std::malloc(sizeof(uint64_t) * (n_lines + 1));
    if (!start) { dcb_set_error("dcb_fastq_index_build: out of memory"); return DCB_ENOMEM; }
    start[0] = 0;
    parallel_for(n_threads, n_bytes, [&](int t, uint64_t a, uint64_t b) {
        uint64_t k = nl[t];
        const char* p = text + a;
        const char* end = text + b;
        while (p < end) {
            const char* q = (const char*)std::memchr(p, '\n', (size_t)(end - p));
            if (!q) break;
            start[++k] = (uint64_t)(q - text) + 1;
            p = q + 1;
        }
    });

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
            const uint64_t h = start[4 * r], s = start[4 * r + 1], p = start[4 * r + 2], q = start[4 * r + 3], e = start[4 * r + 4];
            const uint64_t hl = s - h - 1, sl = p - s - 1, ql = e - q - 1;            // line lengths without the '\n'
            if (text[h] != '@' || text[p] != '+') { irregular[t] = 1; return; }
            if (sl > 0 && (text[s] == '@' || text[s] == '+' || text[s] == '>')) { irregular[t] = 1; return; }   // a marker to readfq
            if (ql < sl || sl > 0xFFFFFFu || ql > 0xFFFFFFu) { irregular[t] = 1; return; }     // multi-line quality / absurd
            const char* sp = (const char*)std::memchr(text + h + 1, ' ', (size_t)(hl - 1 + (hl == 0)));
            ix->name_off[r] = h + 1;
            ix->name_len[r] = hl == 0 ? 0u : (uint32_t)(sp ? (uint64_t)(sp - (text + h + 1)) : hl - 1);
            ix->seq_off[r] = s; ix->seq_len[r] = (uint32_t)sl;
            ix->qual_off[r] = q; ix->qual_len[r] = (uint32_t)ql;
        }
    });
    std::free(start);
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
};

// length of the row of read i (it decombined), or its bytes when dst != nullptr
inline size_t row_emit(const RowCtx& c, uint64_t i, char* dst) {
    const dcb_result& r = c.res[i];
    const bool rev = (c.packed_rc != 0) != (r.frame != 0);
    const int64_t n = c.vdj->len[i], nq = c.qual->len[i];
    int64_t ia = r.ins_start, ib = r.ins_end, sa = r.v_seq_start, sb = r.j_seq_end, qa = r.v_seq_start, qb = r.j_seq_end;
    clip(n, ia, ib); clip(n, sa, sb); clip(nq, qa, qb);
    const size_t fixed = (size_t)dec_len(r.v) + dec_len(r.j) + dec_len(r.vdel) + dec_len(r.jdel);
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
        if (rev) for (int64_t k = a; k < b; k++) *p++ = (char)kComp.t[s[n - 1 - k]];
        else { std::memcpy(p, s + a, (size_t)(b - a)); p += b - a; }
    };
    p = put_dec(p, r.v); sep(); p = put_dec(p, r.j); sep(); p = put_dec(p, r.vdel); sep(); p = put_dec(p, r.jdel); sep();
    seq(ia, ib); sep();
    raw(c.ids); sep();
    seq(sa, sb); sep();
    if (rev) for (int64_t k = qa; k < qb; k++) *p++ = q[nq - 1 - k];
    else { std::memcpy(p, q + qa, (size_t)(qb - qa)); p += qb - qa; }
    sep();
    raw(c.bc); sep();
    raw(c.bcq);
    if (c.tail) { sep(); raw(c.tail); }
    *p++ = '\n';
    return (size_t)(p - dst);
}

}  // namespace

extern "C" {

int dcb_format_rows(const dcb_result* res, uint64_t n, int packed_revcomp, const dcb_column* ids, const dcb_column* vdj,
                    const dcb_column* vdjqual, const dcb_column* bc, const dcb_column* bcq, const dcb_column* v_tail,
                    const char* sep, int n_threads, char** out, uint64_t* out_bytes, uint64_t* n_rows) {
    if (!out || !out_bytes || !n_rows || !sep || (n && (!res || !ids || !vdj || !vdjqual || !bc || !bcq))) {
        dcb_set_error("dcb_format_rows: null argument");
        return DCB_EINVAL;
    }
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    RowCtx c;
    c.res = res; c.packed_rc = packed_revcomp; c.ids = ids; c.vdj = vdj; c.qual = vdjqual; c.bc = bc; c.bcq = bcq; c.tail = v_tail;
    c.sep = sep; c.sep_len = std::strlen(sep);
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
    char* buf = (char*)std::malloc(bytes[nt] + 1);
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
    *out = buf; *out_bytes = bytes[nt]; *n_rows = rows[nt];
    return DCB_OK;
}

void dcb_buffer_free(char* p) { std::free(p); }

}  // extern "C"
