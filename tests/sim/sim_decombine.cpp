// sim_decombine.cpp -- TEST SCAFFOLDING: runs the kernels' per-read device code (dcr_core.cuh,
// compiled for the host by g++) over a packed batch, one "thread" at a time, so that the matching
// logic can be checked against the oracle on a machine without a GPU.  Never linked into libdcb.so
// and never used by the product: the shipped path is the CUDA kernels in csrc/decombine.cu, which
// call exactly the same dcr_exact_read / dcr_general_read functions.
#include "../../decombinator_b200/csrc/dcr_core.cuh"

#include <cstring>
#include <vector>

// why the half-tag path passed reads on (diagnostics): 1 several full-tag candidates / read too long, 2 tag window outside the
// read, 3 too many candidates, 4 V deletion walk not interior, 5 J deletion walk not interior
extern "C" { uint64_t sim_half_pass_reasons[8] = {0}; }
// vidx: the V seed index, or the union index of both genes when jidx is null
extern "C" int sim_decombine(const dcb_packed* P, const uint32_t* vgen, const uint32_t* jgen, const uint32_t* vcore,
                             const uint32_t* jcore, const uint32_t* vidx, const uint32_t* jidx, int both_frames,
                             int allow_ns, int lenthreshold, int mode,
                             dcb_result* out, uint64_t* counters, uint64_t* n_deferred, const uint32_t* sfilt,
                             const uint32_t* half, uint64_t* n_deferred2) {
    DcrParams prm;
    prm.allow_ns = allow_ns; prm.lenthreshold = lenthreshold;
    const int nw = (int)P->slot_words, nwi = (nw + 1) / 2;
    std::vector<uint32_t> inv0(nwi + 1), inv1(nwi + 1), rd1(nw + 1), cand(nwi + 1), hits(DCB_HITS_CAP), inv2(nw + 1);
    dcb_cnt_t cnt[DCB_NCOUNTERS];
    std::memset(cnt, 0, sizeof(cnt));
    ExcList ex;
    ex.read = P->exc_read; ex.pos = P->exc_pos; ex.kind = P->exc_kind; ex.n = P->n_exc;
    uint64_t deferred = 0, deferred2 = 0;
    // what the context uploads beside the exception list: its end marker and the per-32-reads entry index (flat kernel)
    std::vector<uint32_t> xread(P->exc_read, P->exc_read + P->n_exc), xindex((P->n_reads + 31) / 32 + 1, P->n_exc);
    xread.push_back(0xFFFFFFFFu);
    {
        uint32_t e = 0;
        for (size_t k = 0; k < xindex.size(); k++) {
            while (e < P->n_exc && P->exc_read[e] < 32 * k) e++;
            xindex[k] = e;
        }
    }
    for (uint64_t ri = 0; ri < P->n_reads; ri++) {
        ReadView r;
        std::memset(&r, 0, sizeof(r));
        r.w = P->words + ri * P->slot_words; r.stride = 1;
        r.n = P->uniform_len ? (int)P->uniform_len : (int)P->lens[ri];
        r.nw = nw;
        const bool flagged = P->n_exc && ((P->flags[ri >> 5] >> (ri & 31)) & 1u);
        dcb_result o;
        std::memset(&o, 0, sizeof(o));
        int action = FAST_DEFER;
        const ExcProbe xp = flagged ? exc_probe_load(xread.data(), P->exc_pos, P->exc_kind, xindex.data(), (uint32_t)ri) : exc_probe_none();
        uint32_t hand[2] = {DCB_HIT_MULTI, DCB_HIT_MULTI};
        if (mode == 0 || mode == 2)
            action = dcr_exact_read(r, flagged, vcore, jcore, vidx, jidx, prm, both_frames, o, cnt, mode == 2,
                                    mode == 2 && P->n_exc ? &xp : nullptr, hand);   // the flat kernel also takes reads with non-ACGT symbols
        if (action == FAST_DEFER && half && !both_frames) {   // the half-tag path (dcb_halftag_kernel)
            deferred++;
            std::memset(&o, 0, sizeof(o));
            const uint32_t e0 = xp.e0;   // handed over by the exact-tag kernel
            int why = 0;
            if (dcr_half_read(r, flagged, ex, e0, hand[0], hand[1], inv2.data(), hits.data(), 12, vcore, jcore, half, prm, o, cnt, &why))
                action = FAST_DONE;
            else { deferred--; if (why >= 0 && why < 8) sim_half_pass_reasons[why]++; }
        }
        if (action == FAST_DEFER) {
            deferred++; deferred2++;
            std::memset(&o, 0, sizeof(o));
            dcr_general_read(r, (uint32_t)ri, flagged, ex, inv0.data(), rd1.data(), inv1.data(), vgen, jgen, prm,
                             both_frames, o, cnt, sfilt, sfilt ? cand.data() : nullptr, sfilt ? hits.data() : nullptr);
        }
        out[ri] = o;
    }
    for (int i = 0; i < DCB_NCOUNTERS; i++) counters[i] += cnt[i];
    if (n_deferred) *n_deferred = deferred;
    if (n_deferred2) *n_deferred2 = deferred2;
    return 0;
}

// Diagnostic (tools/defer_breakdown.py): why the flat kernel's per-read logic defers a read.  cls[ri] = 0 finished,
// else bit 0: no full V tag, bit 1: no full J tag, bit 2: read has non-ACGT symbols, bit 3: several V / J occurrences,
// bit 4: both tags found but a deletion walk / the exception probe sent it on.
extern "C" int sim_defer_classes(const dcb_packed* P, const uint32_t* vcore, const uint32_t* jcore, const uint32_t* uidx,
                                 int both_frames, int allow_ns, int lenthreshold, uint8_t* cls) {
    DcrParams prm;
    prm.allow_ns = allow_ns; prm.lenthreshold = lenthreshold;
    dcb_cnt_t cnt[DCB_NCOUNTERS];
    std::memset(cnt, 0, sizeof(cnt));
    std::vector<uint32_t> xread(P->exc_read, P->exc_read + P->n_exc), xindex((P->n_reads + 31) / 32 + 1, P->n_exc);
    xread.push_back(0xFFFFFFFFu);
    {
        uint32_t e = 0;
        for (size_t k = 0; k < xindex.size(); k++) {
            while (e < P->n_exc && P->exc_read[e] < 32 * k) e++;
            xindex[k] = e;
        }
    }
    for (uint64_t ri = 0; ri < P->n_reads; ri++) {
        ReadView r;
        std::memset(&r, 0, sizeof(r));
        r.w = P->words + ri * P->slot_words; r.stride = 1;
        r.n = P->uniform_len ? (int)P->uniform_len : (int)P->lens[ri];
        r.nw = (int)P->slot_words;
        const bool flagged = P->n_exc && ((P->flags[ri >> 5] >> (ri & 31)) & 1u);
        const ExcProbe xp = flagged ? exc_probe_load(xread.data(), P->exc_pos, P->exc_kind, xindex.data(), (uint32_t)ri) : exc_probe_none();
        FullHit vh, jh;
        q_find(r, uidx, vh, jh, flagged ? &xp : nullptr);
        dcb_result o;
        std::memset(&o, 0, sizeof(o));
        const int action = dcr_fast_from_hits<false>(r, gene_tags(vcore), gene_tags(jcore), vh, jh, prm, both_frames, o, cnt,
                                                     flagged, xp);
        uint8_t c = 0;
        if (action == FAST_DEFER) {
            if (vh.count == 0) c |= 1;
            if (jh.count == 0) c |= 2;
            if (flagged) c |= 4;
            if (vh.count > 1 || jh.count > 1) c |= 8;
            if (vh.count == 1 && jh.count == 1) c |= 16;
        }
        cls[ri] = c;
    }
    return 0;
}
