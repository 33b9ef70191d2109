/*
 * dcr_oracle.c -- CPU restatement of decombinator's `dcr()` hot path (see dcr_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: the checker for the CUDA path and the reported CPU baseline.
 * It works on ASCII strings exactly like the reference does and re-implements Python's
 * slice semantics literally, because the deletion walks rely on them (negative starts wrap,
 * truncated slices compare equal when their contents are equal).
 *
 * The multi-keyword search lives in a third-party wheel absent from /root/reference
 * (acora==2.4, pyproject.toml:10).  Its published algorithm is Aho-Corasick; findall()
 * reports every occurrence ordered by end position, which is what is restated here.
 * Levenshtein.hamming (Levenshtein==0.25.1, pyproject.toml:20) is a per-position inequality count.
 * Bio.Seq.reverse_complement (biopython==1.84, pyproject.toml:11) complements IUPAC codes
 * case-preservingly and reverses.
 */
#include "dcr_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Aho-Corasick automaton over bytes (acora stand-in; call sites decombine.py:722-746, 275, 294,
 * 339, 399, 422, 473).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int n_states;
    int32_t* next;     /* n_states * 256 : full DFA (goto + failure folded in) */
    int32_t* term;     /* keyword slot ending at this state, or -1 */
    int32_t* dict;     /* longest proper suffix state that is terminal, or 0 */
    int32_t* kw_len;   /* per keyword slot */
    int32_t* kw_first; /* per keyword slot: first index in the source list holding this string */
    int n_kw;
} ac_t;

static void ac_build(ac_t* ac, const char* const* kws, const int* lens, int n) {
    int cap = 1;
    for (int i = 0; i < n; i++) cap += lens[i];
    int32_t* child = (int32_t*)malloc((size_t)cap * 256 * sizeof(int32_t));
    for (size_t i = 0; i < (size_t)cap * 256; i++) child[i] = -1;
    ac->term = (int32_t*)malloc(cap * sizeof(int32_t));
    for (int i = 0; i < cap; i++) ac->term[i] = -1;
    ac->kw_len = (int32_t*)malloc((n + 1) * sizeof(int32_t));
    ac->kw_first = (int32_t*)malloc((n + 1) * sizeof(int32_t));
    ac->n_kw = 0;
    int ns = 1;
    for (int i = 0; i < n; i++) {
        if (lens[i] == 0) continue; /* AcoraBuilder ignores empty keywords */
        int s = 0;
        for (int p = 0; p < lens[i]; p++) {
            unsigned char c = (unsigned char)kws[i][p];
            if (child[(size_t)s * 256 + c] < 0) child[(size_t)s * 256 + c] = ns++;
            s = child[(size_t)s * 256 + c];
        }
        if (ac->term[s] < 0) { /* duplicate keywords collapse to one */
            ac->term[s] = ac->n_kw;
            ac->kw_len[ac->n_kw] = lens[i];
            ac->kw_first[ac->n_kw] = i;
            ac->n_kw++;
        }
    }
    ac->n_states = ns;
    ac->next = (int32_t*)malloc((size_t)ns * 256 * sizeof(int32_t));
    ac->dict = (int32_t*)calloc(ns, sizeof(int32_t));
    int32_t* fail = (int32_t*)calloc(ns, sizeof(int32_t));
    int32_t* queue = (int32_t*)malloc(ns * sizeof(int32_t));
    int qh = 0, qt = 0;
    for (int c = 0; c < 256; c++) {
        int t = child[c];
        if (t < 0) ac->next[c] = 0;
        else { ac->next[c] = t; fail[t] = 0; queue[qt++] = t; }
    }
    while (qh < qt) {
        int s = queue[qh++];
        int f = fail[s];
        ac->dict[s] = (ac->term[f] >= 0) ? f : ac->dict[f];
        for (int c = 0; c < 256; c++) {
            int t = child[(size_t)s * 256 + c];
            if (t < 0) ac->next[(size_t)s * 256 + c] = ac->next[(size_t)f * 256 + c];
            else {
                ac->next[(size_t)s * 256 + c] = t;
                fail[t] = ac->next[(size_t)f * 256 + c];
                queue[qt++] = t;
            }
        }
    }
    free(child); free(fail); free(queue);
}

static void ac_free(ac_t* ac) {
    free(ac->next); free(ac->term); free(ac->dict); free(ac->kw_len); free(ac->kw_first);
}

typedef struct { int32_t kw_first; int32_t start; int32_t len; } hit_t;

/* findall(): every occurrence, ascending end position; at equal end the longer keyword first. */
static int ac_findall(const ac_t* ac, const char* s, int n, hit_t* hits, int cap) {
    int st = 0, nh = 0;
    for (int i = 0; i < n; i++) {
        st = ac->next[(size_t)st * 256 + (unsigned char)s[i]];
        for (int t = (ac->term[st] >= 0) ? st : ac->dict[st]; t; t = ac->dict[t]) {
            int k = ac->term[t];
            if (nh < cap) {
                hits[nh].kw_first = ac->kw_first[k];
                hits[nh].len = ac->kw_len[k];
                hits[nh].start = i + 1 - ac->kw_len[k];
            }
            nh++;
        }
    }
    return nh;
}

/* ------------------------------------------------------------------------------------------
 * Tag data for one gene (what import_tcr_info leaves in module globals, decombine.py:681-746)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int n;
    char** seqs;   int* seq_len;    /* v_seqs / j_seqs                       decombine.py:826-835 */
    char** half1;  int* half1_len;  /* tag[:split]                           decombine.py:840-842 */
    char** half2;  int* half2_len;  /* tag[split:]                                                */
    int* jump;                      /* jump_to_end_v / jump_to_start_j                            */
    char** regions; int* region_len;/* upper-cased FASTA records             decombine.py:690-696 */
    int n_regions;
    int split;
    ac_t key, half1_key, half2_key;
} gene_t;

struct orc_ctx { gene_t v, j; };

static char* dupn(const char* s, int n) {
    char* r = (char*)malloc(n + 1);
    memcpy(r, s, n); r[n] = 0;
    return r;
}

static void gene_init(gene_t* g, const char* const* tags, const int32_t* jumps, const char* const* regions,
                      int n, int split) {
    g->n = n; g->n_regions = n; g->split = split;
    g->seqs = (char**)malloc(n * sizeof(char*)); g->seq_len = (int*)malloc(n * sizeof(int));
    g->half1 = (char**)malloc(n * sizeof(char*)); g->half1_len = (int*)malloc(n * sizeof(int));
    g->half2 = (char**)malloc(n * sizeof(char*)); g->half2_len = (int*)malloc(n * sizeof(int));
    g->jump = (int*)malloc(n * sizeof(int));
    g->regions = (char**)malloc(n * sizeof(char*)); g->region_len = (int*)malloc(n * sizeof(int));
    for (int i = 0; i < n; i++) {
        int L = (int)strlen(tags[i]);
        g->seqs[i] = dupn(tags[i], L); g->seq_len[i] = L;
        int h1 = L < split ? L : split; /* python tag[0:split] / tag[split:] */
        g->half1[i] = dupn(tags[i], h1); g->half1_len[i] = h1;
        g->half2[i] = dupn(tags[i] + h1, L - h1); g->half2_len[i] = L - h1;
        g->jump[i] = jumps[i];
        int R = (int)strlen(regions[i]);
        g->regions[i] = dupn(regions[i], R); g->region_len[i] = R;
    }
    ac_build(&g->key, (const char* const*)g->seqs, g->seq_len, n);
    ac_build(&g->half1_key, (const char* const*)g->half1, g->half1_len, n);
    ac_build(&g->half2_key, (const char* const*)g->half2, g->half2_len, n);
}

static void gene_free(gene_t* g) {
    for (int i = 0; i < g->n; i++) { free(g->seqs[i]); free(g->half1[i]); free(g->half2[i]); free(g->regions[i]); }
    free(g->seqs); free(g->seq_len); free(g->half1); free(g->half1_len); free(g->half2); free(g->half2_len);
    free(g->jump); free(g->regions); free(g->region_len);
    ac_free(&g->key); ac_free(&g->half1_key); ac_free(&g->half2_key);
}

orc_ctx* orc_create(const char* const* v_tags, const int32_t* v_jumps, const char* const* v_regions, int nv,
                    const char* const* j_tags, const int32_t* j_jumps, const char* const* j_regions, int nj,
                    int v_half_split, int j_half_split) {
    orc_ctx* c = (orc_ctx*)calloc(1, sizeof(orc_ctx));
    gene_init(&c->v, v_tags, v_jumps, v_regions, nv, v_half_split);
    gene_init(&c->j, j_tags, j_jumps, j_regions, nj, j_half_split);
    return c;
}

void orc_destroy(orc_ctx* c) {
    if (!c) return;
    gene_free(&c->v); gene_free(&c->j); free(c);
}

/* ------------------------------------------------------------------------------------------
 * Python slice semantics: s[start:stop] on a string of length len -> [a, b)
 * ---------------------------------------------------------------------------------------- */
typedef struct { int a, b; } span_t;

static span_t py_slice(int len, long start, long stop) {
    if (start < 0) { start += len; if (start < 0) start = 0; } else if (start > len) start = len;
    if (stop < 0) { stop += len; if (stop < 0) stop = 0; } else if (stop > len) stop = len;
    if (stop < start) stop = start;
    span_t r = { (int)start, (int)stop };
    return r;
}

static int slices_equal(const char* s1, span_t a, const char* s2, span_t b) {
    int la = a.b - a.a, lb = b.b - b.a;
    if (la != lb) return 0;
    return memcmp(s1 + a.a, s2 + b.a, la) == 0;
}

/* ------------------------------------------------------------------------------------------
 * get_v_deletions (decombine.py:749-785)
 * ---------------------------------------------------------------------------------------- */
static int get_v_deletions(const gene_t* g, const char* read, int n, int v_match, long temp_end_v,
                           long* end_v, long* deletions_v, uint64_t* counts) {
    long f = temp_end_v;
    int m = g->region_len[v_match];
    const char* region = g->regions[v_match];
    long pos = m - 10;                                   /* :754-756 */
    if (f >= n) {                                        /* :760-762 */
        counts[ORC_v_del_failed_tag_at_end]++;
        return 0;
    }
    f += 1;                                              /* :764 */
    long num_del = 0;
    while (0 <= f && f < n) {                            /* :767 */
        span_t a = py_slice(m, pos, pos + 10);           /* :770 */
        span_t b = py_slice(n, f - 10, f);               /* :771 */
        if (slices_equal(region, a, read, b)) {
            *deletions_v = num_del;                      /* :774-775 */
            *end_v = temp_end_v - num_del;
            return 1;
        }
        pos -= 1; num_del += 1; f -= 1;                  /* :777-779 */
    }
    counts[ORC_v_del_failed]++;                          /* :784 */
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * get_j_deletions (decombine.py:788-817)
 * ---------------------------------------------------------------------------------------- */
static int get_j_deletions(const gene_t* g, const char* read, int n, int j_match, long temp_start_j, long end_of_v,
                           long* start_j, long* deletions_j, uint64_t* counts) {
    long f = temp_start_j;
    int m = g->region_len[j_match];
    const char* region = g->regions[j_match];
    long pos = 0;
    while (0 <= f + 2 && f + 2 < n) {                    /* :795 */
        if (f < end_of_v) {                              /* :798-800 */
            pos += 1; f += 1;
        } else {
            span_t a = py_slice(m, pos, pos + 10);       /* :803 */
            span_t b = py_slice(n, f, f + 10);           /* :804 */
            if (slices_equal(region, a, read, b)) {
                *deletions_j = pos;                      /* :807-808 */
                *start_j = f;
                return 1;
            }
            pos += 1; f += 1;                            /* :810-811 */
        }
    }
    counts[ORC_j_del_failed]++;                          /* :816 */
    return 0;
}

static int hamming_le1(const char* a, const char* b, int n) {
    int d = 0;
    for (int i = 0; i < n; i++) d += (a[i] != b[i]);
    return d <= 1;
}

#define MAX_HITS 2048 /* stored hits per findall(); a read yields at most a handful */

typedef struct { long idx, pos_a, pos_b, seq_pos; } vj_t; /* (match, end_v|start_j, deletions, v_seq_start|j_seq_end) */

/* ------------------------------------------------------------------------------------------
 * vanalysis (decombine.py:273-394)
 * ---------------------------------------------------------------------------------------- */
static int vanalysis(const orc_ctx* c, const char* read, int n, vj_t* out, uint64_t* counts) {
    const gene_t* g = &c->v;
    hit_t hold[MAX_HITS];
    int ok = 0;
    int nh = ac_findall(&g->key, read, n, hold, MAX_HITS);           /* :275 */
    if (nh) {
        if (nh > 1) { counts[ORC_multiple_v_matches]++; goto done; }  /* :278-280 */
        int v_match = hold[0].kw_first;                               /* :282 v_seqs.index(tag) */
        long temp_end_v = (long)hold[0].start + g->jump[v_match] - 1; /* :283-285 */
        long end_v, dels;
        if (get_v_deletions(g, read, n, v_match, temp_end_v, &end_v, &dels, counts)) { /* :288-290 */
            out->idx = v_match; out->pos_a = end_v; out->pos_b = dels; out->seq_pos = hold[0].start;
            ok = 1;
        }
        goto done;
    }
    nh = ac_findall(&g->half1_key, read, n, hold, MAX_HITS);          /* :294 */
    if (nh) {
        if (nh > MAX_HITS) nh = MAX_HITS;
        for (int i = 0; i < nh; i++) {                                /* :297 */
            const char* kw = g->half1[hold[i].kw_first];
            int kwl = hold[i].len;
            int first = hold[i].kw_first;                             /* half1_v_seqs.index(...) :305 */
            long p = hold[i].start;
            for (int k = 0; k < g->n; k++) {                          /* :298-301 */
                if (g->half1_len[k] != kwl || memcmp(g->half1[k], kw, kwl) != 0) continue;
                span_t gs = py_slice(n, p, p + g->seq_len[first]);    /* :302-307 */
                if (g->seq_len[k] != gs.b - gs.a) continue;
                span_t hs = py_slice(n, p, p + g->seq_len[k]);        /* :311-314 */
                if (hs.b - hs.a != g->seq_len[k]) continue;           /* lev.hamming needs equal lengths */
                if (!hamming_le1(g->seqs[k], read + hs.a, g->seq_len[k])) continue; /* :309-316 */
                counts[ORC_verr2]++;                                  /* :318 */
                long temp_end_v = p + g->jump[k] - 1;                 /* :320-322 */
                long end_v, dels;
                if (get_v_deletions(g, read, n, k, temp_end_v, &end_v, &dels, counts)) { /* :323-333 */
                    out->idx = k; out->pos_a = end_v; out->pos_b = dels; out->seq_pos = p;
                    ok = 1; goto done;
                }
            }
        }
        counts[ORC_foundv1notv2]++;                                   /* :334 */
        goto done;
    }
    nh = ac_findall(&g->half2_key, read, n, hold, MAX_HITS);          /* :339 */
    if (nh) {
        if (nh > MAX_HITS) nh = MAX_HITS;
        for (int i = 0; i < nh; i++) {                                /* :341 */
            const char* kw = g->half2[hold[i].kw_first];
            int kwl = hold[i].len;
            int first = hold[i].kw_first;
            long p = hold[i].start;
            for (int k = 0; k < g->n; k++) {                          /* :342-347 */
                if (g->half2_len[k] != kwl || memcmp(g->half2[k], kw, kwl) != 0) continue;
                span_t gs = py_slice(n, p - g->split, p - g->split + g->seq_len[first]); /* :348-357 */
                if (g->seq_len[k] != gs.b - gs.a) continue;
                span_t hs = py_slice(n, p - g->split, p + g->seq_len[k] - g->split);     /* :361-366 */
                if (hs.b - hs.a != g->seq_len[k]) continue;
                if (!hamming_le1(g->seqs[k], read + hs.a, g->seq_len[k])) continue;      /* :359-368 */
                counts[ORC_verr1]++;                                  /* :370 */
                long temp_end_v = p + g->jump[k] - g->split - 1;      /* :372-377 */
                long end_v, dels;
                if (get_v_deletions(g, read, n, k, temp_end_v, &end_v, &dels, counts)) { /* :378-388 */
                    out->idx = k; out->pos_a = end_v; out->pos_b = dels; out->seq_pos = p - g->split;
                    ok = 1; goto done;
                }
            }
        }
        counts[ORC_foundv2notv1]++;                                   /* :389 */
        goto done;
    }
    counts[ORC_no_vtags_found]++;                                     /* :393 */
done:
    return ok;
}

/* ------------------------------------------------------------------------------------------
 * janalysis (decombine.py:397-531)
 * ---------------------------------------------------------------------------------------- */
static int janalysis(const orc_ctx* c, const char* read, int n, long end_of_v, vj_t* out, uint64_t* counts) {
    const gene_t* g = &c->j;
    hit_t hold[MAX_HITS];
    int ok = 0;
    int nh = ac_findall(&g->key, read, n, hold, MAX_HITS);            /* :399 */
    if (nh) {
        if (nh > 1) { counts[ORC_multiple_j_matches]++; goto done; }  /* :402-404 */
        int j_match = hold[0].kw_first;                               /* :406 */
        long temp_start_j = (long)hold[0].start - g->jump[j_match];   /* :407-409 */
        long j_seq_end = (long)hold[0].start + hold[0].len;           /* :411 */
        long start_j, dels;
        if (get_j_deletions(g, read, n, j_match, temp_start_j, end_of_v, &start_j, &dels, counts)) { /* :413-418 */
            out->idx = j_match; out->pos_a = start_j; out->pos_b = dels; out->seq_pos = j_seq_end;
            ok = 1;
        }
        goto done;
    }
    nh = ac_findall(&g->half1_key, read, n, hold, MAX_HITS);          /* :422 */
    if (nh) {
        if (nh > MAX_HITS) nh = MAX_HITS;
        for (int i = 0; i < nh; i++) {                                /* :424 */
            const char* kw = g->half1[hold[i].kw_first];
            int kwl = hold[i].len;
            int first = hold[i].kw_first;
            long p = hold[i].start;
            for (int k = 0; k < g->n; k++) {                          /* :425-428 */
                if (g->half1_len[k] != kwl || memcmp(g->half1[k], kw, kwl) != 0) continue;
                span_t gs = py_slice(n, p, p + g->seq_len[first]);    /* :429-434 */
                if (g->seq_len[k] != gs.b - gs.a) continue;
                span_t hs = py_slice(n, p, p + g->seq_len[k]);        /* :438-441 */
                if (hs.b - hs.a != g->seq_len[k]) continue;
                if (!hamming_le1(g->seqs[k], read + hs.a, g->seq_len[k])) continue; /* :436-443 */
                counts[ORC_jerr2]++;                                  /* :445 */
                long temp_start_j = p - g->jump[k];                   /* :447-449 */
                long j_seq_end = p + kwl + g->split;                  /* :450-454 */
                long start_j, dels;
                if (get_j_deletions(g, read, n, k, temp_start_j, end_of_v, &start_j, &dels, counts)) { /* :455-468 */
                    out->idx = k; out->pos_a = start_j; out->pos_b = dels; out->seq_pos = j_seq_end;
                    ok = 1; goto done;
                }
            }
        }
        counts[ORC_foundj1notj2]++;                                   /* :469 */
        goto done;
    }
    nh = ac_findall(&g->half2_key, read, n, hold, MAX_HITS);          /* :473 */
    if (nh) {
        if (nh > MAX_HITS) nh = MAX_HITS;
        for (int i = 0; i < nh; i++) {                                /* :475 */
            const char* kw = g->half2[hold[i].kw_first];
            int kwl = hold[i].len;
            int first = hold[i].kw_first;
            long p = hold[i].start;
            for (int k = 0; k < g->n; k++) {                          /* :476-481 */
                if (g->half2_len[k] != kwl || memcmp(g->half2[k], kw, kwl) != 0) continue;
                span_t gs = py_slice(n, p - g->split, p - g->split + g->seq_len[first]); /* :482-491 */
                if (g->seq_len[k] != gs.b - gs.a) continue;
                span_t hs = py_slice(n, p - g->split, p + g->seq_len[k] - g->split);     /* :495-500 */
                if (hs.b - hs.a != g->seq_len[k]) continue;
                if (!hamming_le1(g->seqs[k], read + hs.a, g->seq_len[k])) continue;      /* :493-502 */
                counts[ORC_jerr1]++;                                  /* :504 */
                long temp_start_j = p - g->jump[k] - g->split;        /* :506-510 */
                long j_seq_end = p + kwl;                             /* :511 */
                long start_j, dels;
                if (get_j_deletions(g, read, n, k, temp_start_j, end_of_v, &start_j, &dels, counts)) { /* :512-525 */
                    out->idx = k; out->pos_a = start_j; out->pos_b = dels; out->seq_pos = j_seq_end;
                    ok = 1; goto done;
                }
            }
        }
        counts[ORC_foundv2notv1]++;   /* :526 -- the reference bumps the V counter here; preserved */
        goto done;
    }
    counts[ORC_no_j_assigned]++;                                      /* :530 */
done:
    return ok;
}

/* ------------------------------------------------------------------------------------------
 * dcr (decombine.py:534-585)
 * ---------------------------------------------------------------------------------------- */
void orc_dcr(const orc_ctx* c, const char* read, int n, int allowNs, int lenthreshold,
             orc_result* out, uint64_t* counts) {
    memset(out, 0, sizeof(*out));
    vj_t vdat, jdat;
    if (!vanalysis(c, read, n, &vdat, counts)) return;                /* :542-545 */
    long end_of_v = vdat.pos_a + 1;                                   /* :547 */
    if (!janalysis(c, read, n, end_of_v, &jdat, counts)) {            /* :548, :583-585 */
        counts[ORC_VJ_assignment_failed]++;
        return;
    }
    span_t it = py_slice(n, vdat.seq_pos, jdat.seq_pos);              /* :554 */
    int hasN = 0;
    for (int i = it.a; i < it.b; i++) if (read[i] == 'N') { hasN = 1; break; }
    int vlen = c->v.seq_len[vdat.idx], jlen = c->j.seq_len[jdat.idx];
    if (hasN && !allowNs) {
        counts[ORC_dcrfilter_intertagN]++;                            /* :556 */
    } else if ((vdat.seq_pos - jdat.seq_pos) >= lenthreshold) {       /* :557-560 */
        counts[ORC_dcrfilter_toolong_intertag]++;
    } else if (vdat.pos_b > (c->v.jump[vdat.idx] - vlen) || jdat.pos_b > c->j.jump[jdat.idx]) { /* :561-565 */
        counts[ORC_dcrfilter_imposs_deletion]++;
    } else if ((vdat.seq_pos + vlen) > (jdat.seq_pos + jlen)) {       /* :566-569 */
        counts[ORC_dcrfilter_tag_overlap]++;
    } else {                                                          /* :572-581 */
        out->ok = 1;
        out->v = (int32_t)vdat.idx; out->j = (int32_t)jdat.idx;
        out->vdel = (int32_t)vdat.pos_b; out->jdel = (int32_t)jdat.pos_b;
        out->ins_start = (int32_t)(vdat.pos_a + 1); out->ins_end = (int32_t)jdat.pos_a;
        out->v_seq_start = (int32_t)vdat.seq_pos; out->j_seq_end = (int32_t)jdat.seq_pos;
    }
}

/* ------------------------------------------------------------------------------------------
 * revcomp (decombine.py:182-184): Bio.Seq complement table, case preserving, then reversed
 * ---------------------------------------------------------------------------------------- */
static unsigned char g_comp[256];
static pthread_once_t g_comp_once = PTHREAD_ONCE_INIT;

static void comp_init(void) {
    const char* from = "ACGTUMRWSYKVHDBNacgtumrwsykvhdbn";
    const char* to   = "TGCAAKYWSRMBDHVNtgcaakywsrmbdhvn";
    for (int i = 0; i < 256; i++) g_comp[i] = (unsigned char)i;
    for (int i = 0; from[i]; i++) g_comp[(unsigned char)from[i]] = (unsigned char)to[i];
}

void orc_revcomp(const char* src, int n, char* dst) {
    pthread_once(&g_comp_once, comp_init);
    for (int i = 0; i < n; i++) dst[i] = (char)g_comp[(unsigned char)src[n - 1 - i]];
}

/* ------------------------------------------------------------------------------------------
 * orientation handling of the main loop (decombine.py:999-1010) over a batch
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const orc_ctx* c; const char* ascii; const uint64_t* off; const uint32_t* len;
    uint64_t lo, hi; int orientation, allowNs, lenthreshold; orc_result* results;
    uint64_t counts[ORC_NCOUNTERS];
} job_t;

static void* job_run(void* arg) {
    job_t* jb = (job_t*)arg;
    int cap = 1024;
    char* buf = (char*)malloc(cap);
    for (uint64_t r = jb->lo; r < jb->hi; r++) {
        const char* read = jb->ascii + jb->off[r];
        int n = (int)jb->len[r];
        if (n + 1 > cap) { cap = 2 * (n + 1); buf = (char*)realloc(buf, cap); }
        orc_result* out = &jb->results[r];
        if (jb->orientation == 1) {                       /* forward :1002-1004 */
            orc_dcr(jb->c, read, n, jb->allowNs, jb->lenthreshold, out, jb->counts);
            out->frame = 1;
        } else {                                          /* reverse :999-1001 / both :1005-1010 */
            orc_revcomp(read, n, buf);
            orc_dcr(jb->c, buf, n, jb->allowNs, jb->lenthreshold, out, jb->counts);
            out->frame = 0;
            if (jb->orientation == 2 && !out->ok) {
                orc_dcr(jb->c, read, n, jb->allowNs, jb->lenthreshold, out, jb->counts);
                out->frame = 1;
            }
        }
    }
    free(buf);
    return NULL;
}

void orc_decombine(const orc_ctx* c, const char* ascii, const uint64_t* off, const uint32_t* len, uint64_t n_reads,
                   int orientation, int allowNs, int lenthreshold, orc_result* results, uint64_t* counters,
                   int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if ((uint64_t)nthreads > n_reads) nthreads = n_reads ? (int)n_reads : 1;
    job_t* jobs = (job_t*)calloc(nthreads, sizeof(job_t));
    pthread_t* th = (pthread_t*)malloc(nthreads * sizeof(pthread_t));
    pthread_once(&g_comp_once, comp_init);
    for (int t = 0; t < nthreads; t++) {
        jobs[t].c = c; jobs[t].ascii = ascii; jobs[t].off = off; jobs[t].len = len;
        jobs[t].lo = n_reads * (uint64_t)t / nthreads; jobs[t].hi = n_reads * (uint64_t)(t + 1) / nthreads;
        jobs[t].orientation = orientation; jobs[t].allowNs = allowNs; jobs[t].lenthreshold = lenthreshold;
        jobs[t].results = results;
        if (nthreads == 1) job_run(&jobs[t]);
        else pthread_create(&th[t], NULL, job_run, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) {
        if (nthreads > 1) pthread_join(th[t], NULL);
        for (int k = 0; k < ORC_NCOUNTERS; k++) counters[k] += jobs[t].counts[k];
    }
    free(jobs); free(th);
}

int orc_findall(const orc_ctx* c, int which, const char* read, int n, int32_t* kw_first_index, int32_t* start, int cap) {
    const gene_t* g = which < 3 ? &c->v : &c->j;
    const ac_t* ac = (which % 3 == 0) ? &g->key : (which % 3 == 1) ? &g->half1_key : &g->half2_key;
    hit_t* hold = (hit_t*)malloc((cap > 0 ? cap : 1) * sizeof(hit_t));
    int nh = ac_findall(ac, read, n, hold, cap);
    for (int i = 0; i < nh && i < cap; i++) { kw_first_index[i] = hold[i].kw_first; start[i] = hold[i].start; }
    free(hold);
    return nh;
}
