// pack.cpp -- FASTQ sequence strings -> 2-bit packed, page-locked read slots (host side of the ingest).
//
// Replaces the per-read string handling in front of dcr() in the reference's main loop:
// `vdj = record1[1]` (decombine.py:965-977) and revcomp(vdj) (decombine.py:182-184, 1000), i.e.
// Bio.Seq.reverse_complement: complement through the IUPAC table (case preserving), then reverse.
// Only A/C/G/T can ever match a tag or a germline region (both are upper-case ACGT,
// decombine.py:695), so every other symbol becomes an entry of the sparse exception list.
#include "dcb_internal.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

struct Tables {
    unsigned char comp[256];
    signed char code[256];  // 0..3 for ACGT, -1 otherwise
    Tables() {
        const char* from = "ACGTUMRWSYKVHDBNacgtumrwsykvhdbn";
        const char* to = "TGCAAKYWSRMBDHVNtgcaakywsrmbdhvn";
        for (int i = 0; i < 256; i++) { comp[i] = (unsigned char)i; code[i] = -1; }
        for (int i = 0; from[i]; i++) comp[(unsigned char)from[i]] = (unsigned char)to[i];
        code['A'] = 0; code['C'] = 1; code['G'] = 2; code['T'] = 3;
    }
};
const Tables kT;

struct Owner {
    bool pinned = false;
    std::vector<std::pair<void*, bool>> blocks;  // (pointer, page-locked?)
};

void* host_alloc(Owner* o, size_t bytes) {
    if (bytes == 0) bytes = 16;
    void* p = nullptr;
    bool locked = false;
    if (o->pinned) {
        if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess) locked = true;
        else { (void)cudaGetLastError(); p = nullptr; }
    }
    if (!p && posix_memalign(&p, 256, (bytes + 255) & ~(size_t)255) != 0) return nullptr;
    o->blocks.emplace_back(p, locked);
    return p;
}

struct Exc { uint32_t read; uint16_t pos; uint8_t kind; };

#if defined(__x86_64__)
const bool kAvx2 = __builtin_cpu_supports("avx2");
// 32 bases -> 64 packed bits.  v: the 32 symbols in output order (already reversed for the reverse strand).  `bad` collects
// the symbols that are not A / C / G / T (upper case; everything else is the byte loop's business).
__attribute__((target("avx2"))) inline uint64_t pack32_avx2(__m256i v, bool comp, __m256i& bad) {
    // by low nibble: 'A' 0x41 -> 1, 'C' 0x43 -> 3, 'T' 0x54 -> 4, 'G' 0x47 -> 7
    const __m256i self = _mm256_setr_epi8(-1, 'A', -1, 'C', 'T', -1, -1, 'G', -1, -1, -1, -1, -1, -1, -1, -1,
                                          -1, 'A', -1, 'C', 'T', -1, -1, 'G', -1, -1, -1, -1, -1, -1, -1, -1);
    const __m256i low = _mm256_set1_epi8(0x0F), three = _mm256_set1_epi8(3);
    const __m256i nib = _mm256_and_si256(v, low);
    // an ASCII symbol whose low nibble names one of the four and that IS that one; bytes >= 0x80 shuffle to 0 and differ
    bad = _mm256_or_si256(bad, _mm256_xor_si256(_mm256_cmpeq_epi8(_mm256_shuffle_epi8(self, nib), v), _mm256_set1_epi8(-1)));
    // code = ((c >> 1) ^ (c >> 2)) & 3 maps A, C, G, T -> 0, 1, 2, 3 (16-bit shifts: the bits that cross a byte are masked off)
    __m256i c = _mm256_and_si256(_mm256_xor_si256(_mm256_srli_epi16(v, 1), _mm256_srli_epi16(v, 2)), three);
    if (comp) c = _mm256_xor_si256(c, three);
    // four codes per byte: pairs (b0 + 4 b1) as 16-bit lanes, then pairs of those (w0 + 16 w1) as 32-bit lanes
    const __m256i p16 = _mm256_maddubs_epi16(c, _mm256_set1_epi16(0x0401));
    const __m256i p32 = _mm256_madd_epi16(p16, _mm256_set1_epi32(0x00100001));
    // the low byte of every 32-bit lane, four per 128-bit half
    const __m256i pick = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                          0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    const __m256i q = _mm256_shuffle_epi8(p32, pick);
    return (uint64_t)(uint32_t)_mm256_extract_epi32(q, 0) | ((uint64_t)(uint32_t)_mm256_extract_epi32(q, 4) << 32);
}
// One read of A / C / G / T into its slot (all sw words written).  false: another symbol turned up (the slot's content is
// then void and the caller packs the read with the byte loop).
__attribute__((target("avx2"))) inline bool pack_read_avx2(const unsigned char* s, uint32_t L, bool revcomp, uint32_t* w, size_t sw) {
    const __m256i flip = _mm256_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
    __m256i bad = _mm256_setzero_si256();
    uint32_t i = 0;
    for (; i + 32 <= L; i += 32) {
        __m256i v;
        if (revcomp) {   // output bases i .. i+31 are the input bytes L-32-i .. L-1-i backwards
            v = _mm256_loadu_si256((const __m256i*)(s + (L - 32 - i)));
            v = _mm256_permute4x64_epi64(_mm256_shuffle_epi8(v, flip), 0x4E);
        } else {
            v = _mm256_loadu_si256((const __m256i*)(s + i));
        }
        const uint64_t bits = pack32_avx2(v, revcomp, bad);
        std::memcpy(w + (i >> 4), &bits, 8);
    }
    if (i < L && L >= 32) {
        // the last, partial block: the 32 symbols that END the output (they overlap the block before, whose bits are the
        // same, so they are OR-ed in), shifted to their place
        __m256i v;
        if (revcomp) v = _mm256_permute4x64_epi64(_mm256_shuffle_epi8(_mm256_loadu_si256((const __m256i*)s), flip), 0x4E);
        else v = _mm256_loadu_si256((const __m256i*)(s + (L - 32)));
        const uint64_t bits = pack32_avx2(v, revcomp, bad);
        const uint32_t b0 = L - 32, u = b0 >> 5, sh = 2 * (b0 & 31);
        uint64_t lo;
        std::memcpy(&lo, w + 2 * u, 8);
        lo |= bits << sh;
        std::memcpy(w + 2 * u, &lo, 8);
        if (sh) { const uint64_t hi = bits >> (64 - sh); std::memcpy(w + 2 * u + 2, &hi, 8); }
        i = 32 * (u + (sh ? 2 : 1));
    } else if (i < L) {  // a read of fewer than 32 symbols: through a buffer padded with 'A' (code 0; 'T' on the reverse strand, 3 ^ 3)
        alignas(32) unsigned char tmp[32];
        const uint32_t rem = L - i;
        std::memset(tmp, revcomp ? 'T' : 'A', 32);
        if (revcomp) for (uint32_t k = 0; k < rem; k++) tmp[k] = s[rem - 1 - k];
        else std::memcpy(tmp, s + i, rem);
        const uint64_t bits = pack32_avx2(_mm256_load_si256((const __m256i*)tmp), revcomp, bad);
        if ((i >> 4) + 1 < sw) std::memcpy(w + (i >> 4), &bits, 8);
        else { const uint32_t lo = (uint32_t)bits; std::memcpy(w + (i >> 4), &lo, 4); }
        i += 32;
    }
    for (size_t k = (size_t)(i >> 4); k < sw; k++) w[k] = 0;
    return _mm256_testz_si256(bad, bad) != 0;
}
// The same with AVX-512 (VBMI for the byte reversal): 64 bases -> 128 packed bits per step.  bad: bit k set <=> symbol k of
// some block was not A / C / G / T.
#define DCB_AVX512 __attribute__((target("avx512f,avx512bw,avx512vbmi,avx512vl")))
const bool kAvx512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vbmi") &&
                     __builtin_cpu_supports("avx512vl") && !std::getenv("DCB_NO_AVX512");
// How a packed slot is written: 2 = assembled in a local buffer, then streaming stores (measured on the GPU box with the
// copy engine reading the same memory: dcb_decombine_ascii with every chunk packed by the host 271 -> 307 M reads/s, shared
// with the device 345 -> 354); 0 = stored directly, 1 = local buffer and ordinary stores (DCB_PACK_STORE, measurements).
const int kPackStore = std::getenv("DCB_PACK_STORE") ? std::atoi(std::getenv("DCB_PACK_STORE")) : 2;
DCB_AVX512 inline __m128i pack64_avx512(__m512i v, bool comp, __mmask64& bad) {
    const __m512i self = _mm512_broadcast_i32x4(_mm_setr_epi8(-1, 'A', -1, 'C', 'T', -1, -1, 'G', -1, -1, -1, -1, -1, -1, -1, -1));
    const __m512i three = _mm512_set1_epi8(3);
    // a symbol whose low nibble names one of the four and that IS that one (bytes >= 0x80 shuffle to 0 and differ)
    bad |= _mm512_cmpneq_epi8_mask(_mm512_shuffle_epi8(self, _mm512_and_si512(v, _mm512_set1_epi8(0x0F))), v);
    __m512i c = _mm512_and_si512(_mm512_xor_si512(_mm512_srli_epi16(v, 1), _mm512_srli_epi16(v, 2)), three);
    if (comp) c = _mm512_xor_si512(c, three);
    const __m512i p16 = _mm512_maddubs_epi16(c, _mm512_set1_epi16(0x0401));
    const __m512i p32 = _mm512_madd_epi16(p16, _mm512_set1_epi32(0x00100001));
    return _mm512_cvtepi32_epi8(p32);                    // the low byte of every 32-bit lane: four bases each
}
DCB_AVX512 inline bool pack_read_avx512(const unsigned char* s, uint32_t L, bool revcomp, uint32_t* w, size_t sw) {
    const __m512i flip = _mm512_set_epi8(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31,
                                         32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63);
    __mmask64 bad = 0;
    uint32_t i = 0;
    if (kPackStore && sw <= 32 && ((uintptr_t)w & 15) == 0) {
        // The usual slot (reads of up to 512 bases): assembled in a local buffer and written with streaming stores -- the
        // slot is not read again by this CPU (the copy engine or the kernels' host takes it from memory), so the lines
        // need not be fetched before they are overwritten.  The caller fences (pack_fence) before the words are handed on.
        alignas(64) uint64_t t[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (; i + 64 <= L; i += 64) {
            __m512i v;
            if (revcomp) v = _mm512_permutexvar_epi8(flip, _mm512_loadu_si512(s + (L - 64 - i)));
            else v = _mm512_loadu_si512(s + i);
            _mm_store_si128((__m128i*)(t + (i >> 5)), pack64_avx512(v, revcomp, bad));
        }
        if (i < L) {
            __m512i v;
            if (revcomp) v = _mm512_permutexvar_epi8(flip, _mm512_loadu_si512(s));
            else v = _mm512_loadu_si512(s + (L - 64));
            alignas(16) uint64_t bits[2];
            _mm_store_si128((__m128i*)bits, pack64_avx512(v, revcomp, bad));
            const uint32_t b0 = L - 64, u = b0 >> 5, sh = 2 * (b0 & 31);
            if (sh == 0) { t[u] = bits[0]; t[u + 1] = bits[1]; }
            else { t[u] |= bits[0] << sh; t[u + 1] = (bits[0] >> (64 - sh)) | (bits[1] << sh); t[u + 2] = bits[1] >> (64 - sh); }
        }
        if (kPackStore == 2 && ((uintptr_t)w & 63) == 0 && (sw & 15) == 0)
            for (size_t k = 0; k < sw / 16; k++) _mm512_stream_si512((__m512i*)w + k, _mm512_load_si512((const __m512i*)t + k));
        else
            for (size_t k = 0; k < sw / 4; k++) _mm_store_si128((__m128i*)w + k, _mm_load_si128((const __m128i*)t + k));
        return bad == 0;
    }
    for (; i + 64 <= L; i += 64) {
        __m512i v;
        if (revcomp) v = _mm512_permutexvar_epi8(flip, _mm512_loadu_si512(s + (L - 64 - i)));   // output bases i .. i+63: the input bytes L-64-i .. L-1-i backwards
        else v = _mm512_loadu_si512(s + i);
        _mm_storeu_si128((__m128i*)(w + (i >> 4)), pack64_avx512(v, revcomp, bad));
    }
    uint32_t next = i >> 4;                                  // first word not written yet
    if (i < L) {
        // the last, partial block: the 64 symbols that END the output (they overlap the block before, whose bits are the
        // same, so they are OR-ed in), shifted to their place
        __m512i v;
        if (revcomp) v = _mm512_permutexvar_epi8(flip, _mm512_loadu_si512(s));
        else v = _mm512_loadu_si512(s + (L - 64));
        alignas(16) uint64_t bits[2];
        _mm_store_si128((__m128i*)bits, pack64_avx512(v, revcomp, bad));
        const uint32_t b0 = L - 64, u = b0 >> 5, sh = 2 * (b0 & 31);
        if (sh == 0) { std::memcpy(w + 2 * u, bits, 16); next = 2 * (u + 2); }
        else {
            uint64_t out[3];
            std::memcpy(&out[0], w + 2 * u, 8);
            out[0] |= bits[0] << sh;
            out[1] = (bits[0] >> (64 - sh)) | (bits[1] << sh);
            out[2] = bits[1] >> (64 - sh);
            std::memcpy(w + 2 * u, out, 24);
            next = 2 * (u + 3);
        }
    }
    for (size_t k = next; k < sw; k++) w[k] = 0;
    return bad == 0;
}
// One read of nothing but A / C / G / T into its slot with the widest unit the CPU has; false: another symbol (or no unit).
inline void pack_fence() { _mm_sfence(); }               // after a thread's last read: its streaming stores are visible
inline bool pack_read_simd(const unsigned char* s, uint32_t L, bool revcomp, uint32_t* w, size_t sw) {
    if (kAvx512 && L >= 64) return pack_read_avx512(s, L, revcomp, w, sw);
    return kAvx2 && pack_read_avx2(s, L, revcomp, w, sw);
}
#endif

}  // namespace

extern "C" {

int dcb_pack_reads(const char* ascii, const uint64_t* off, const uint32_t* len, uint64_t n, int revcomp,
                   int n_threads, dcb_packed** out) {
    if (!out || (n && (!ascii || !off || !len))) { dcb_set_error("dcb_pack_reads: null argument"); return DCB_EINVAL; }
    if (n >= 0xFFFFFFFFull) { dcb_set_error("dcb_pack_reads: at most 2^32-2 reads per batch"); return DCB_EINVAL; }
    uint32_t max_len = 0, min_len = 0xFFFFFFFFu;
    for (uint64_t i = 0; i < n; i++) { max_len = std::max(max_len, len[i]); min_len = std::min(min_len, len[i]); }
    if (max_len > DCB_MAX_READ_LEN) {
        dcb_set_error("dcb_pack_reads: read of %u nt exceeds the supported maximum of %d", max_len, DCB_MAX_READ_LEN);
        return DCB_EUNSUPPORTED;
    }
    Owner* o = new Owner();
    int dev_count = 0;
    o->pinned = cudaGetDeviceCount(&dev_count) == cudaSuccess && dev_count > 0;
    if (!o->pinned) (void)cudaGetLastError();
    dcb_packed* P = new dcb_packed();
    std::memset(P, 0, sizeof(*P));
    P->owner = o;
    P->n_reads = n;
    P->max_len = max_len;
    P->slot_words = std::max<uint32_t>(4u, ((max_len + 63) / 64) * 4);
    P->uniform_len = (n && min_len == max_len) ? max_len : 0;
    const size_t sw = P->slot_words;
    P->words = (uint32_t*)host_alloc(o, n * sw * 4);
    P->lens = (uint16_t*)host_alloc(o, n * 2);
    P->flags = (uint32_t*)host_alloc(o, ((n + 31) / 32) * 4);
    if (!P->words || !P->lens || !P->flags) { dcb_packed_free(P); dcb_set_error("dcb_pack_reads: out of memory"); return DCB_ENOMEM; }
    std::memset(P->flags, 0, ((n + 31) / 32) * 4);

    if (n_threads < 1) n_threads = 1;
    uint64_t chunks = (n + 31) / 32;  // threads own whole flag words
    if ((uint64_t)n_threads > chunks) n_threads = chunks ? (int)chunks : 1;
    std::vector<std::vector<Exc>> excs(n_threads);
    auto work = [&](int t) {
        uint64_t lo = (chunks * t / n_threads) * 32, hi = std::min<uint64_t>(n, (chunks * (t + 1) / n_threads) * 32);
        std::vector<Exc>& ex = excs[t];
        for (uint64_t r = lo; r < hi; r++) {
            const unsigned char* s = (const unsigned char*)ascii + off[r];
            const uint32_t L = len[r];
            uint32_t* w = P->words + r * sw;
            P->lens[r] = (uint16_t)L;
            bool flagged = false;
            uint32_t acc = 0;
            uint32_t i = 0;
#if defined(__x86_64__)
            if (pack_read_simd(s, L, revcomp != 0, w, sw)) continue;     // nothing but A / C / G / T: done
#endif
            // eight bases at a time while they are all A/C/G/T (SWAR): code = ((c >> 1) ^ (c >> 2)) & 3 maps
            // A,C,G,T -> 0,1,2,3; the reverse complement reads the bytes from the end (byte swap) and flips the code.
            // Any other symbol in a group hands the rest of the read to the byte loop below.
            for (; i + 8 <= L; i += 8) {
                uint64_t x;
                if (revcomp) { std::memcpy(&x, s + (L - 8 - i), 8); x = __builtin_bswap64(x); }
                else std::memcpy(&x, s + i, 8);
                const uint64_t k01 = 0x0101010101010101ull, k7f = 0x7F7F7F7F7F7F7F7Full, k80 = 0x8080808080808080ull;
                // 0x80 in exactly the bytes of v that are 0 (no carry crosses a byte: exact per byte)
                auto zero_bytes = [&](uint64_t v) { return ~(((v & k7f) + k7f) | v | k7f); };
                const uint64_t ok = zero_bytes(x ^ (k01 * 'A')) | zero_bytes(x ^ (k01 * 'C')) | zero_bytes(x ^ (k01 * 'G')) |
                                    zero_bytes(x ^ (k01 * 'T'));
                if (ok != k80) break;        // another symbol in this group: the byte loop below takes over
                uint64_t v = ((x >> 1) ^ (x >> 2)) & (k01 * 3);
                if (revcomp) v ^= k01 * 3;
                v = (v | (v >> 6)) & 0x000F000F000F000Full;
                v = (v | (v >> 12)) & 0x000000FF000000FFull;
                v = (v | (v >> 24)) & 0xFFFFull;
                acc |= (uint32_t)v << (2 * (i & 15));
                if ((i & 15) == 8) { w[i >> 4] = acc; acc = 0; }
            }
            for (; i < L; i++) {
                unsigned char c = revcomp ? kT.comp[s[L - 1 - i]] : s[i];
                int code = kT.code[c];
                if (code < 0) {
                    ex.push_back({(uint32_t)r, (uint16_t)i, (uint8_t)(c == 'N' ? 1 : 2)});
                    flagged = true;
                    code = 0;
                } else if (revcomp && s[L - 1 - i] == 'U') {
                    // complement('U') = 'A' is a real base in this frame but the forward frame of
                    // `-or both` sees 'U': kind 3 = valid here, invalid in the other frame
                    ex.push_back({(uint32_t)r, (uint16_t)i, 3});
                    flagged = true;
                }
                acc |= (uint32_t)code << (2 * (i & 15));
                if ((i & 15) == 15) { w[i >> 4] = acc; acc = 0; }
            }
            if (i & 15) w[i >> 4] = acc;
            for (uint32_t k = (L + 15) / 16; k < sw; k++) w[k] = 0;
            if (flagged) P->flags[r >> 5] |= 1u << (r & 31);
        }
#if defined(__x86_64__)
        pack_fence();
#endif
    };
    if (n_threads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++) th.emplace_back(work, t);
        for (auto& t : th) t.join();
    }
    size_t n_exc = 0;
    for (auto& e : excs) n_exc += e.size();
    if (n_exc >= 0xFFFFFFFFull) { dcb_packed_free(P); dcb_set_error("dcb_pack_reads: too many non-ACGT symbols"); return DCB_EUNSUPPORTED; }
    P->n_exc = (uint32_t)n_exc;
    P->exc_read = (uint32_t*)host_alloc(o, (n_exc + 1) * 4);
    P->exc_pos = (uint16_t*)host_alloc(o, (n_exc + 1) * 2);
    P->exc_kind = (uint8_t*)host_alloc(o, n_exc + 1);
    if (!P->exc_read || !P->exc_pos || !P->exc_kind) { dcb_packed_free(P); dcb_set_error("dcb_pack_reads: out of memory"); return DCB_ENOMEM; }
    size_t k = 0;
    for (auto& e : excs)
        for (auto& x : e) { P->exc_read[k] = x.read; P->exc_pos[k] = x.pos; P->exc_kind[k] = x.kind; k++; }
    *out = P;
    return DCB_OK;
}

// Reads of nothing but A / C / G / T, packed into the caller's buffer (count * slot_words words; page-locked when it is the
// staging buffer of dcb_decombine_ascii).  *clean = 0 when another symbol turned up or the CPU has no AVX2: the buffer's
// content is then void and the caller takes another path (the device packer, dcb_pack_reads).
int dcb_pack_words(const char* ascii, const uint64_t* off, const uint32_t* len, uint64_t first, uint64_t count, uint32_t uniform_len,
                   int revcomp, uint32_t slot_words, uint32_t* words, int n_threads, int* clean) {
    if (!clean || (count && (!ascii || !words)) || (!off && !uniform_len) || (!len && !uniform_len)) {
        dcb_set_error("dcb_pack_words: null argument");
        return DCB_EINVAL;
    }
    *clean = 0;
#if defined(__x86_64__)
    if (!kAvx2) return DCB_OK;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    if (count < 4096) n_threads = 1;
    std::vector<int> dirty(n_threads, 0);
    auto work = [&](int t) {
        const uint64_t lo = count * (uint64_t)t / n_threads, hi = count * (uint64_t)(t + 1) / n_threads;
        for (uint64_t i = lo; i < hi; i++) {
            const uint64_t r = first + i;
            const uint32_t L = uniform_len ? uniform_len : len[r];
            const unsigned char* s = (const unsigned char*)ascii + (off ? off[r] : r * (uint64_t)uniform_len);
            if ((L + 15) / 16 > slot_words || !pack_read_simd(s, L, revcomp != 0, words + i * slot_words, slot_words)) { dirty[t] = 1; pack_fence(); return; }
        }
        pack_fence();
    };
    if (n_threads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++) th.emplace_back(work, t);
        for (auto& t : th) t.join();
    }
    int any = 0;
    for (int d : dirty) any |= d;
    *clean = !any;
#endif
    return DCB_OK;
}

// An empty dcb_packed of the given shape (the device packer's output is copied into it); free with dcb_packed_free.
int dcb_packed_alloc(uint64_t n, uint32_t slot_words, uint32_t n_exc, dcb_packed** out) {
    Owner* o = new Owner();
    int dev_count = 0;
    o->pinned = cudaGetDeviceCount(&dev_count) == cudaSuccess && dev_count > 0;
    if (!o->pinned) (void)cudaGetLastError();
    dcb_packed* P = new dcb_packed();
    std::memset(P, 0, sizeof(*P));
    P->owner = o;
    P->n_reads = n; P->slot_words = slot_words; P->n_exc = n_exc;
    P->words = (uint32_t*)host_alloc(o, n * slot_words * 4);
    P->lens = (uint16_t*)host_alloc(o, n * 2);
    P->flags = (uint32_t*)host_alloc(o, ((n + 31) / 32) * 4);
    P->exc_read = (uint32_t*)host_alloc(o, ((size_t)n_exc + 1) * 4);
    P->exc_pos = (uint16_t*)host_alloc(o, ((size_t)n_exc + 1) * 2);
    P->exc_kind = (uint8_t*)host_alloc(o, (size_t)n_exc + 1);
    if (!P->words || !P->lens || !P->flags || !P->exc_read || !P->exc_pos || !P->exc_kind) {
        dcb_packed_free(P);
        dcb_set_error("dcb_packed_alloc: out of memory");
        return DCB_ENOMEM;
    }
    *out = P;
    return DCB_OK;
}

void dcb_packed_free(dcb_packed* P) {
    if (!P) return;
    Owner* o = (Owner*)P->owner;
    if (o) {
        for (auto& b : o->blocks) {
            if (b.second) cudaFreeHost(b.first);
            else std::free(b.first);
        }
        delete o;
    }
    delete P;
}

int dcb_unpack_read(const dcb_packed* P, uint64_t i, char* dst, uint32_t cap) {
    if (!P || !dst || i >= P->n_reads) return DCB_EINVAL;
    uint32_t L = P->lens[i];
    if (cap < L) return DCB_EINVAL;
    const uint32_t* w = P->words + i * P->slot_words;
    for (uint32_t p = 0; p < L; p++) dst[p] = "ACGT"[(w[p >> 4] >> (2 * (p & 15))) & 3];
    if ((P->flags[i >> 5] >> (i & 31)) & 1u) {
        const uint32_t* b = std::lower_bound(P->exc_read, P->exc_read + P->n_exc, (uint32_t)i);
        for (size_t e = b - P->exc_read; e < P->n_exc && P->exc_read[e] == i; e++)
            if (P->exc_kind[e] != 3) dst[P->exc_pos[e]] = P->exc_kind[e] == 1 ? 'N' : '?';
    }
    return (int)L;
}

}  // extern "C"
