/* _ac.c -- TEST INFRASTRUCTURE: a compiled Aho-Corasick automaton behind the acora stand-in (oracle/standins/acora.py),
 * so that timing the UNMODIFIED reference with the stand-ins (oracle/time_reference.py) charges the tag search what the
 * real acora==2.4 (a Cython automaton walking one transition per byte) would cost, not a Python loop of str.find.
 * Semantics reproduced: findall(s) lists every (overlapping) occurrence of every keyword as (keyword index, start),
 * ordered by END position, the longer keyword first at equal end (acora's report order; decombine.py:275-473 relies on it).
 * Built by oracle/time_reference.py / tests with:  gcc -O2 -shared -fPIC oracle/standins/_ac.c -o oracle/_ref/libac.so  */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int n_states, cap;
    int32_t* next;      /* [n_states][256] goto function, completed into a DFA */
    int32_t* fail;
    int32_t* out_head;  /* per state: first entry of its output list (-1 none) */
    int32_t* out_kw;    /* entries: keyword index */
    int32_t* out_next;  /* entries: next entry */
    int n_out, out_cap;
    int32_t* kw_len;
    int n_kw;
} ac_t;

static int ac_new_state(ac_t* a) {
    if (a->n_states == a->cap) {
        a->cap = a->cap ? 2 * a->cap : 64;
        a->next = (int32_t*)realloc(a->next, sizeof(int32_t) * 256 * (size_t)a->cap);
        a->fail = (int32_t*)realloc(a->fail, sizeof(int32_t) * (size_t)a->cap);
        a->out_head = (int32_t*)realloc(a->out_head, sizeof(int32_t) * (size_t)a->cap);
    }
    for (int c = 0; c < 256; c++) a->next[256 * (size_t)a->n_states + c] = -1;
    a->fail[a->n_states] = 0;
    a->out_head[a->n_states] = -1;
    return a->n_states++;
}
static void ac_add_out(ac_t* a, int state, int kw) {
    if (a->n_out == a->out_cap) {
        a->out_cap = a->out_cap ? 2 * a->out_cap : 64;
        a->out_kw = (int32_t*)realloc(a->out_kw, sizeof(int32_t) * (size_t)a->out_cap);
        a->out_next = (int32_t*)realloc(a->out_next, sizeof(int32_t) * (size_t)a->out_cap);
    }
    /* keep every list ordered longest keyword first */
    int32_t* link = &a->out_head[state];
    while (*link >= 0 && a->kw_len[a->out_kw[*link]] >= a->kw_len[kw]) link = &a->out_next[*link];
    a->out_kw[a->n_out] = kw;
    a->out_next[a->n_out] = *link;
    *link = a->n_out++;
}

ac_t* ac_build(const char* const* kws, const int32_t* lens, int n) {
    ac_t* a = (ac_t*)calloc(1, sizeof(ac_t));
    a->n_kw = n;
    a->kw_len = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n ? n : 1));
    for (int i = 0; i < n; i++) a->kw_len[i] = lens[i];
    ac_new_state(a);
    for (int i = 0; i < n; i++) {
        int s = 0;
        for (int k = 0; k < lens[i]; k++) {
            const unsigned char c = (unsigned char)kws[i][k];
            if (a->next[256 * (size_t)s + c] < 0) { const int t = ac_new_state(a); a->next[256 * (size_t)s + c] = t; }
            s = a->next[256 * (size_t)s + c];
        }
        ac_add_out(a, s, i);
    }
    /* breadth first: failure links, outputs inherited from the failure state, goto completed */
    int32_t* queue = (int32_t*)malloc(sizeof(int32_t) * (size_t)a->n_states);
    int head = 0, tail = 0;
    for (int c = 0; c < 256; c++) {
        const int t = a->next[c];
        if (t < 0) a->next[c] = 0;
        else { a->fail[t] = 0; queue[tail++] = t; }
    }
    while (head < tail) {
        const int s = queue[head++];
        for (int c = 0; c < 256; c++) {
            const int t = a->next[256 * (size_t)s + c];
            if (t < 0) { a->next[256 * (size_t)s + c] = a->next[256 * (size_t)a->fail[s] + c]; continue; }
            a->fail[t] = a->next[256 * (size_t)a->fail[s] + c];
            for (int e = a->out_head[a->fail[t]]; e >= 0; e = a->out_next[e]) ac_add_out(a, t, a->out_kw[e]);
            queue[tail++] = t;
        }
    }
    free(queue);
    return a;
}

/* -> number of occurrences; the first `cap` are written as (keyword index, start) pairs */
int ac_findall(const ac_t* a, const char* s, int n, int32_t* out, int cap) {
    int state = 0, found = 0;
    for (int i = 0; i < n; i++) {
        state = a->next[256 * (size_t)state + (unsigned char)s[i]];
        for (int e = a->out_head[state]; e >= 0; e = a->out_next[e]) {
            if (found < cap) { out[2 * found] = a->out_kw[e]; out[2 * found + 1] = i + 1 - a->kw_len[a->out_kw[e]]; }
            found++;
        }
    }
    return found;
}

void ac_free(ac_t* a) {
    if (!a) return;
    free(a->next); free(a->fail); free(a->out_head); free(a->out_kw); free(a->out_next); free(a->kw_len); free(a);
}
