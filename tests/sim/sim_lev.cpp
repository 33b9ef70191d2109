// sim_lev.cpp -- TEST SCAFFOLDING: the collapse distance code (csrc/lev_core.cuh) compiled for the host by g++, so the
// bit-parallel Levenshtein logic can be checked against a textbook DP on a machine without a GPU.  Never linked into
// libdcb.so; the product path is the CUDA kernels of csrc/collapse.cu, which call the same functions.
#include "../../decombinator_b200/csrc/lev_core.cuh"

extern "C" int sim_umi_distance(uint64_t a, uint64_t b) {
    UmiPattern p;
    umi_pattern(a, p);
    return umi_distance(p, b);
}
extern "C" int sim_umi_may_be_within(uint64_t a, uint64_t b, int k) { return umi_may_be_within(a, b, k) ? 1 : 0; }
template <int NP>
static int seq_dist(const uint8_t* a, int la, const uint8_t* b, int lb) {
    if (la > lb) { const uint8_t* t = a; a = b; b = t; int x = la; la = lb; lb = x; }
    if (la <= 64) { SeqPattern<1, NP> p; seq_pattern<1, NP>(a, la, p); return seq_distance<1, NP>(p, b, lb); }
    if (la <= 128) { SeqPattern<2, NP> p; seq_pattern<2, NP>(a, la, p); return seq_distance<2, NP>(p, b, lb); }
    if (la <= 192) { SeqPattern<3, NP> p; seq_pattern<3, NP>(a, la, p); return seq_distance<3, NP>(p, b, lb); }
    if (la <= 256) { SeqPattern<4, NP> p; seq_pattern<4, NP>(a, la, p); return seq_distance<4, NP>(p, b, lb); }
    SeqPattern<8, NP> p; seq_pattern<8, NP>(a, la, p); return seq_distance<8, NP>(p, b, lb);
}
extern "C" int sim_seq_distance(const uint8_t* a, int la, const uint8_t* b, int lb) { return seq_dist<3>(a, la, b, lb); }
extern "C" int sim_seq_distance_bytes(const uint8_t* a, int la, const uint8_t* b, int lb) { return seq_dist<8>(a, la, b, lb); }
