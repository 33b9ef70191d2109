/*
 * dcr_oracle.h -- CPU restatement of decombinator's per-read `dcr()` path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under decombinator_b200/ may include,
 * link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / `--impl reference` legs use it, and only as the checker or
 * the reported CPU baseline.
 *
 * Parity pinning: this restatement is checked (tests/test_oracle_golden.py)
 * against (1) the reference's own golden files dcr_TINY_1_{alpha,beta}.n12
 * (reference tests/test_pipeline.py:63-84) and (2) fixtures recorded by
 * running the UNMODIFIED reference source in the build container
 * (oracle/make_golden.py -> tests/golden/dcr_cases_*.json.gz).
 *
 * Every function cites the reference file:line it follows; paths are
 * relative to /root/reference/src/decombinator/.
 */
#ifndef DCR_ORACLE_H
#define DCR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Counter slots, named after the keys of the reference's `counts` Counter
 * (decombine.py:279..584; summary at decombine.py:1131-1195). */
enum {
    ORC_verr1 = 0,
    ORC_verr2,
    ORC_jerr1,
    ORC_jerr2,
    ORC_dcrfilter_intertagN,
    ORC_dcrfilter_toolong_intertag,
    ORC_dcrfilter_imposs_deletion,
    ORC_dcrfilter_tag_overlap,
    ORC_multiple_v_matches,
    ORC_v_del_failed_tag_at_end,
    ORC_v_del_failed,
    ORC_foundv1notv2,
    ORC_foundv2notv1,
    ORC_no_vtags_found,
    ORC_multiple_j_matches,
    ORC_j_del_failed,
    ORC_foundj1notj2,
    ORC_foundj2notj1,
    ORC_no_j_assigned,
    ORC_VJ_assignment_failed,
    ORC_NCOUNTERS
};

/* What dcr() returns (decombine.py:572-581), with the insert kept as the two
 * slice bounds of read[end_v+1 : start_j]; ok==0 means dcr() returned None. */
typedef struct orc_result {
    int32_t ok;
    int32_t v, j, vdel, jdel;
    int32_t ins_start, ins_end;     /* read[ins_start:ins_end] (python slice) */
    int32_t v_seq_start, j_seq_end; /* recom[5], recom[6] */
    int32_t frame;                  /* 0 = reverse, 1 = forward (decombine.py:999-1010) */
} orc_result;

typedef struct orc_ctx orc_ctx;

orc_ctx* orc_create(const char* const* v_tags, const int32_t* v_jumps, const char* const* v_regions, int nv,
                    const char* const* j_tags, const int32_t* j_jumps, const char* const* j_regions, int nj,
                    int v_half_split, int j_half_split);
void orc_destroy(orc_ctx*);

/* One call of dcr(read, inputargs) (decombine.py:534-585). */
void orc_dcr(const orc_ctx*, const char* read, int n, int allowNs, int lenthreshold,
             orc_result* out, uint64_t* counters);

/* Bio.Seq reverse_complement as used by revcomp() (decombine.py:182-184). dst may not alias src. */
void orc_revcomp(const char* src, int n, char* dst);

/* The orientation logic of the main loop (decombine.py:999-1010) over a batch of
 * ASCII reads; orientation: 0 reverse, 1 forward, 2 both.  nthreads>1 shards the
 * batch over pthreads (reads are independent; counters are summed). */
void orc_decombine(const orc_ctx*, const char* ascii, const uint64_t* off, const uint32_t* len, uint64_t n_reads,
                   int orientation, int allowNs, int lenthreshold, orc_result* results, uint64_t* counters,
                   int nthreads);

/* findall() of the acora automaton built from one keyword list (decombine.py:722-746):
 * which: 0 v_key, 1 half1_v_key, 2 half2_v_key, 3 j_key, 4 half1_j_key, 5 half2_j_key.
 * Writes up to cap (keyword-id-in-list-of-first-occurrence, start) pairs, returns the count. */
int orc_findall(const orc_ctx*, int which, const char* read, int n, int32_t* kw_first_index, int32_t* start, int cap);

#ifdef __cplusplus
}
#endif
#endif
