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
