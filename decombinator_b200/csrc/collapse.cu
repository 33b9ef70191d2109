// collapse.cu -- the distance primitives of the collapse step on the GPU (sm_100a).
//
//   dcb_umi_pairs  all unordered pairs of UMIs within Levenshtein distance max_edits.  Replaces
//                  prsnn.symdel(umi_list, max_edits=bcthreshold, output_type="coo_matrix") + sparse.triu +
//                  sum_duplicates  (/root/reference/src/decombinator/collapse.py:735-742): output is the sorted
//                  list of (row, col), row < col, i.e. the COO entries in the order make_clusters walks them.
//   dcb_lev_leq    batch of are_seqs_equivalent(seq1, seq2, frac) verdicts (collapse.py:355-360):
//                  polyleven.levenshtein(a, b) <= len(shorter) * frac, compared in double like the reference.
//
// Both are integer-pipe bound (no tensor cores: nothing is a dense contraction).  The pair search is a tiled
// all-pairs sweep -- tile of 256 UMIs staged in shared memory and broadcast to 256 threads that each keep one
// UMI's match masks in registers -- with a pigeonhole prefilter in front of the bit-parallel verifier, a
// warp-aggregated append of the surviving pairs and a device radix sort of the 64-bit (row << 32 | col) keys.
#include "dcb_internal.h"
#include "lev_core.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <vector>

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            dcb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return DCB_ENOGPU;                                                                  \
        }                                                                                       \
    } while (0)

static constexpr int kPairTile = 256;

// grid.x enumerates the tile pairs (bi <= bj) of the upper triangle
__global__ void __launch_bounds__(kPairTile)
dcb_umi_pairs_kernel(const uint64_t* __restrict__ codes, uint32_t n, int k, uint32_t n_tiles,
                     unsigned long long* __restrict__ keys, unsigned long long cap, unsigned long long* __restrict__ count) {
    __shared__ uint64_t s_j[kPairTile];
    // decode the linear block index into (bi, bj), bi <= bj: row bi of the triangle starts at bi*n_tiles - bi*(bi-1)/2
    unsigned long long lin = blockIdx.x;
    uint32_t bi = 0;
    {
        // largest bi with start(bi) <= lin, by a float estimate corrected with integer steps
        const double nt = (double)n_tiles;
        double est = nt + 0.5 - sqrt((nt + 0.5) * (nt + 0.5) - 2.0 * (double)lin);
        long long b = (long long)est;
        if (b < 0) b = 0;
        if (b >= (long long)n_tiles) b = n_tiles - 1;
        auto start = [&](long long x) { return (unsigned long long)x * n_tiles - (unsigned long long)(x * (x - 1) / 2); };
        while (b > 0 && start(b) > lin) b--;
        while (b + 1 < (long long)n_tiles && start(b + 1) <= lin) b++;
        bi = (uint32_t)b;
        lin -= start(b);
    }
    const uint32_t bj = bi + (uint32_t)lin;
    const uint32_t i = bi * kPairTile + threadIdx.x;
    const uint32_t j0 = bj * kPairTile;
    s_j[threadIdx.x] = (j0 + threadIdx.x < n) ? codes[j0 + threadIdx.x] : 0ull;
    __syncthreads();
    const bool live = i < n;
    const uint64_t ci = live ? codes[i] : 0ull;
    UmiPattern pat;
    umi_pattern(ci, pat);
    const uint32_t jn = min((uint32_t)kPairTile, n - j0);
    const int lane = threadIdx.x & 31;
    for (uint32_t t = 0; t < jn; t++) {
        const uint32_t j = j0 + t;
        const uint64_t cj = s_j[t];
        bool hit = live && j > i && umi_may_be_within(ci, cj, k);
        if (hit) hit = umi_distance(pat, cj) <= k;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, hit);
        if (m) {
            const int leader = __ffs(m) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(count, (unsigned long long)__popc(m));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (hit) {
                const unsigned long long slot = base + __popc(m & ((1u << lane) - 1u));
                if (slot < cap) keys[slot] = ((unsigned long long)i << 32) | j;
            }
        }
    }
}

// one thread per pair; the shorter sequence is the pattern
template <int W>
__device__ __forceinline__ int lev_pair(const uint8_t* pa, int la, const uint8_t* pb, int lb) {
    SeqPattern<W> pat;
    seq_pattern<W>(pa, la, pat);
    return seq_distance<W>(pat, pb, lb);
}

__global__ void __launch_bounds__(128)
dcb_lev_leq_kernel(const uint8_t* __restrict__ sym, const uint64_t* __restrict__ off, const uint32_t* __restrict__ len,
                   const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, unsigned long long n_pairs, double frac,
                   uint8_t* __restrict__ verdict) {
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pairs) return;
    uint32_t ia = a[t], ib = b[t];
    int la = (int)len[ia], lb = (int)len[ib];
    if (la > lb) { const uint32_t x = ia; ia = ib; ib = x; const int y = la; la = lb; lb = y; }
    const uint8_t* pa = sym + off[ia];
    const uint8_t* pb = sym + off[ib];
    int d;
    if (la <= 64) d = lev_pair<1>(pa, la, pb, lb);
    else if (la <= 128) d = lev_pair<2>(pa, la, pb, lb);
    else if (la <= 192) d = lev_pair<3>(pa, la, pb, lb);
    else if (la <= 256) d = lev_pair<4>(pa, la, pb, lb);
    else d = lev_pair<8>(pa, la, pb, lb);
    // threshold = len(min(seq1, seq2, key=len)) * lev_threshold_fraction ; return distance <= threshold  (collapse.py:359-360)
    verdict[t] = ((double)d <= (double)la * frac) ? 1 : 0;
}

struct dcb_dist {
    int device = 0;
    cudaStream_t stream = nullptr;
    unsigned long long* d_keys = nullptr;      // sorted pair keys of the last dcb_umi_pairs
    unsigned long long n_keys = 0;
    double last_ms = 0.0;
};

extern "C" {

dcb_dist* dcb_dist_create(int device) {
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        (void)cudaGetLastError();
        dcb_set_error("dcb_dist_create: no CUDA device available (there is no CPU fallback)");
        return nullptr;
    }
    if (device < 0 || device >= n_dev || cudaSetDevice(device) != cudaSuccess) {
        dcb_set_error("dcb_dist_create: device %d not usable", device);
        return nullptr;
    }
    dcb_dist* d = new dcb_dist();
    d->device = device;
    if (cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess) {
        dcb_set_error("dcb_dist_create: cudaStreamCreate failed");
        delete d;
        return nullptr;
    }
    return d;
}

void dcb_dist_destroy(dcb_dist* d) {
    if (!d) return;
    cudaSetDevice(d->device);
    if (d->stream) { cudaStreamSynchronize(d->stream); cudaStreamDestroy(d->stream); }
    cudaFree(d->d_keys);
    delete d;
}

int dcb_umi_pairs(dcb_dist* d, const uint64_t* codes, uint32_t n, int max_edits, uint64_t* keys, uint64_t cap,
                  uint64_t* n_pairs) {
    if (!d || !n_pairs) { dcb_set_error("dcb_umi_pairs: null argument"); return DCB_EINVAL; }
    CUDA_TRY(cudaSetDevice(d->device));
    cudaStream_t s = d->stream;
    if (codes) {   // compute (and cache) the sorted pair list
        if (max_edits < 0 || max_edits > DCB_UMI_MAX_LEN) { dcb_set_error("dcb_umi_pairs: max_edits out of range"); return DCB_EINVAL; }
        for (uint32_t i = 0; i < n; i++)
            if ((codes[i] >> 58) > DCB_UMI_MAX_LEN) { dcb_set_error("dcb_umi_pairs: UMI %u longer than %d symbols", i, DCB_UMI_MAX_LEN); return DCB_EUNSUPPORTED; }
        cudaFree(d->d_keys); d->d_keys = nullptr; d->n_keys = 0;
        *n_pairs = 0;
        if (n < 2) return DCB_OK;
        uint64_t* d_codes = nullptr;
        unsigned long long* d_count = nullptr;
        unsigned long long* d_raw = nullptr;
        CUDA_TRY(cudaMalloc((void**)&d_codes, (size_t)n * 8));
        CUDA_TRY(cudaMalloc((void**)&d_count, 8));
        CUDA_TRY(cudaMemcpyAsync(d_codes, codes, (size_t)n * 8, cudaMemcpyHostToDevice, s));
        const uint32_t n_tiles = (n + kPairTile - 1) / kPairTile;
        const unsigned long long n_blocks = (unsigned long long)n_tiles * (n_tiles + 1) / 2;
        if (n_blocks > 0x7FFFFFFFull) { cudaFree(d_codes); cudaFree(d_count); dcb_set_error("dcb_umi_pairs: too many UMIs for one launch (%u)", n); return DCB_EUNSUPPORTED; }
        unsigned long long capacity = std::max<unsigned long long>(1ull << 20, 8ull * n);
        unsigned long long found = 0;
        cudaEvent_t e0, e1;
        CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
        for (int attempt = 0; attempt < 2; attempt++) {
            CUDA_TRY(cudaMalloc((void**)&d_raw, capacity * 8));
            CUDA_TRY(cudaMemsetAsync(d_count, 0, 8, s));
            CUDA_TRY(cudaEventRecord(e0, s));
            dcb_umi_pairs_kernel<<<(unsigned)n_blocks, kPairTile, 0, s>>>(d_codes, n, max_edits, n_tiles, d_raw, capacity, d_count);
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaEventRecord(e1, s));
            CUDA_TRY(cudaMemcpyAsync(&found, d_count, 8, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaStreamSynchronize(s));
            if (found <= capacity) break;
            cudaFree(d_raw); d_raw = nullptr;      // the list did not fit: size it exactly and sweep again
            capacity = found;
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        d->last_ms = ms;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        cudaFree(d_codes); cudaFree(d_count);
        if (found) {
            unsigned long long* d_sorted = nullptr;
            CUDA_TRY(cudaMalloc((void**)&d_sorted, found * 8));
            size_t tmp_bytes = 0;
            CUDA_TRY(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, d_raw, d_sorted, (size_t)found, 0, 64, s));
            void* d_tmp = nullptr;
            CUDA_TRY(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 16));
            CUDA_TRY(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, d_raw, d_sorted, (size_t)found, 0, 64, s));
            CUDA_TRY(cudaStreamSynchronize(s));
            cudaFree(d_tmp);
            d->d_keys = d_sorted;
        }
        cudaFree(d_raw);
        d->n_keys = found;
    }
    *n_pairs = d->n_keys;
    if (keys) {
        if (cap < d->n_keys) { dcb_set_error("dcb_umi_pairs: output holds %llu pairs, %llu needed", (unsigned long long)cap, d->n_keys); return DCB_ENOMEM; }
        if (d->n_keys) {
            CUDA_TRY(cudaMemcpyAsync(keys, d->d_keys, d->n_keys * 8, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaStreamSynchronize(s));
        }
    }
    return DCB_OK;
}

int dcb_lev_leq(dcb_dist* d, const uint8_t* symbols, const uint64_t* off, const uint32_t* len, uint32_t n_seqs,
                const uint32_t* a, const uint32_t* b, uint64_t n_pairs, double frac, uint8_t* verdict) {
    if (!d || (n_pairs && (!symbols || !off || !len || !a || !b || !verdict))) { dcb_set_error("dcb_lev_leq: null argument"); return DCB_EINVAL; }
    if (n_pairs == 0) return DCB_OK;
    CUDA_TRY(cudaSetDevice(d->device));
    cudaStream_t s = d->stream;
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_seqs; i++) {
        if (len[i] > 512) { dcb_set_error("dcb_lev_leq: sequence %u has %u symbols (limit 512)", i, len[i]); return DCB_EUNSUPPORTED; }
        total = std::max<uint64_t>(total, off[i] + len[i]);
    }
    for (uint64_t t = 0; t < total; t++)
        if (symbols[t] > 7) { dcb_set_error("dcb_lev_leq: symbol code %u at %llu (codes are 0..7)", symbols[t], (unsigned long long)t); return DCB_EINVAL; }
    for (uint64_t t = 0; t < n_pairs; t++)
        if (a[t] >= n_seqs || b[t] >= n_seqs) { dcb_set_error("dcb_lev_leq: pair %llu names a sequence out of range", (unsigned long long)t); return DCB_EINVAL; }
    uint8_t *d_sym = nullptr, *d_ver = nullptr;
    uint64_t* d_off = nullptr;
    uint32_t *d_len = nullptr, *d_a = nullptr, *d_b = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d_sym, total + 16));
    CUDA_TRY(cudaMalloc((void**)&d_off, (size_t)n_seqs * 8));
    CUDA_TRY(cudaMalloc((void**)&d_len, (size_t)n_seqs * 4));
    CUDA_TRY(cudaMalloc((void**)&d_a, n_pairs * 4));
    CUDA_TRY(cudaMalloc((void**)&d_b, n_pairs * 4));
    CUDA_TRY(cudaMalloc((void**)&d_ver, n_pairs));
    CUDA_TRY(cudaMemcpyAsync(d_sym, symbols, total, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_off, off, (size_t)n_seqs * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_len, len, (size_t)n_seqs * 4, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_a, a, n_pairs * 4, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_b, b, n_pairs * 4, cudaMemcpyHostToDevice, s));
    const unsigned long long blocks = (n_pairs + 127) / 128;
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    CUDA_TRY(cudaEventRecord(e0, s));
    dcb_lev_leq_kernel<<<(unsigned)blocks, 128, 0, s>>>(d_sym, d_off, d_len, d_a, d_b, n_pairs, frac, d_ver);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(e1, s));
    CUDA_TRY(cudaMemcpyAsync(verdict, d_ver, n_pairs, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) d->last_ms = ms;   // kernel time of this call (dcb_dist_last_ms)
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    cudaFree(d_sym); cudaFree(d_off); cudaFree(d_len); cudaFree(d_a); cudaFree(d_b); cudaFree(d_ver);
    return DCB_OK;
}

int dcb_dist_last_ms(dcb_dist* d, double* ms) {
    if (!d || !ms) return DCB_EINVAL;
    *ms = d->last_ms;
    return DCB_OK;
}

}  // extern "C"
