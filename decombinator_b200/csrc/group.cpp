// group.cpp -- the order-dependent grouping of collapse's read_in_data (reference collapse.py:595-682) over COLUMNS:
// barcode codes, global row indices and the inter-tag sequences of the rows that passed the filters.  Host logic (the
// reference's is a Python loop over rows); the sequence comparisons it needs -- are_seqs_equivalent, collapse.py:355-360 --
// are NOT computed here: every round hands the caller the (sequence, sequence) pairs whose verdict is missing, the caller
// gets them from the GPU (dcb_lev_leq) and feeds them back.
//
// Rules per barcode, rows in input order (collapse.py:595-682, as decombinator_b200/collapse.py::_BarcodeMachine restates
// them): the first row founds the group and is its proto-sequence; a row whose sequence equals the proto-sequence or is
// equivalent to it joins, and the proto-sequence becomes the group's most common sequence (ties: the one whose first copy
// came first), which re-inserts the group at the END of the reference's dict -- recorded as `tick`, the index of the row
// that caused it; the first row that is not equivalent kills the group and blacklists the barcode: that row and every later
// row of the barcode are counted as dropped.
#include "dcb_internal.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

struct SeqCount { uint32_t seq, count, first; };

struct Group {
    uint32_t lo = 0, hi = 0;        // its rows in `order`
    uint32_t pos = 0;               // next row (relative to lo)
    uint32_t proto = 0;             // sequence id
    uint64_t tick = 0;
    bool founded = false, dead = false;
    uint32_t dropped = 0;
    std::vector<SeqCount> counts;   // few distinct sequences per barcode: linear search
};

}  // namespace

struct dcb_group {
    const char* text = nullptr;
    uint64_t n = 0;
    std::vector<uint64_t> off;      // sequence of row i: text[off[i], off[i] + len[i])
    std::vector<uint32_t> len;
    std::vector<uint32_t> seq_id;   // row -> distinct sequence
    std::vector<uint32_t> rep;      // distinct sequence -> a row that holds it
    std::vector<uint64_t> code, idx;
    std::vector<uint32_t> order;    // rows sorted by (barcode code, row index)
    std::vector<Group> groups;
    std::vector<uint32_t> open;     // groups not finished
    std::unordered_map<uint64_t, int8_t> verdict;   // (min id << 32 | max id) -> 0 / 1, -1 while asked for
    std::vector<uint64_t> asked;    // the keys of this round, in the order handed out
};

namespace {

inline uint64_t pair_key(uint32_t a, uint32_t b) { return a < b ? ((uint64_t)a << 32) | b : ((uint64_t)b << 32) | a; }

// Advance one group until its rows are used up (true) or a verdict is missing (false; every verdict the rest of the group
// could need against the current proto-sequence is asked for at once, so that a barcode with many variants does not take
// one round per variant).
bool advance(dcb_group& G, Group& g) {
    const uint32_t n_rows = g.hi - g.lo;
    while (g.pos < n_rows) {
        const uint32_t row = G.order[g.lo + g.pos];
        const uint32_t seq = G.seq_id[row];
        if (g.dead) {
            g.dropped++;
        } else if (!g.founded) {
            g.founded = true; g.proto = seq; g.tick = G.idx[row];
            g.counts.clear();
            g.counts.push_back({seq, 1u, 0u});
        } else {
            bool same = seq == g.proto;
            if (!same) {
                auto it = G.verdict.find(pair_key(g.proto, seq));
                if (it == G.verdict.end() || it->second < 0) {
                    for (uint32_t p = g.pos; p < n_rows; p++) {
                        const uint32_t s2 = G.seq_id[G.order[g.lo + p]];
                        if (s2 == g.proto) continue;
                        const uint64_t key = pair_key(g.proto, s2);
                        if (G.verdict.emplace(key, (int8_t)-1).second) G.asked.push_back(key);
                    }
                    return false;
                }
                same = it->second != 0;
            }
            if (same) {
                // members so far: every row before this one (all of them joined, or the group would be dead)
                SeqCount* mine = nullptr;
                SeqCount* best = nullptr;
                for (SeqCount& c : g.counts) { if (c.seq == seq) mine = &c; if (c.seq == g.proto) best = &c; }
                if (!mine) {
                    g.counts.push_back({seq, 0u, g.pos});
                    mine = &g.counts.back();
                    best = nullptr;
                    for (SeqCount& c : g.counts) if (c.seq == g.proto) best = &c;
                }
                mine->count++;
                if (seq != g.proto && (mine->count > best->count || (mine->count == best->count && mine->first < best->first))) {
                    g.proto = seq; g.tick = G.idx[row];
                }
            } else {
                g.dead = true;
                g.dropped++;
            }
        }
        g.pos++;
    }
    return true;
}

}  // namespace

extern "C" {

/* seqs: the n sequences joined by '\n' (a final newline is optional); code / idx: barcode code and global row index per row. */
int dcb_group_create(const char* seqs, uint64_t seqs_bytes, uint64_t n, const uint64_t* code, const uint64_t* idx, dcb_group** out) {
    if (!out || (n && (!seqs || !code || !idx))) { dcb_set_error("dcb_group_create: null argument"); return DCB_EINVAL; }
    if (n >= 0xFFFFFFFFull) { dcb_set_error("dcb_group_create: at most 2^32-2 rows"); return DCB_EINVAL; }
    dcb_group* G = new dcb_group();
    G->text = seqs; G->n = n;
    G->off.resize(n); G->len.resize(n); G->seq_id.resize(n);
    G->code.assign(code, code + n); G->idx.assign(idx, idx + n);
    uint64_t p = 0;
    for (uint64_t i = 0; i < n; i++) {
        if (p > seqs_bytes) { delete G; dcb_set_error("dcb_group_create: fewer than %llu lines", (unsigned long long)n); return DCB_EINVAL; }
        const char* nl = p < seqs_bytes ? (const char*)std::memchr(seqs + p, '\n', (size_t)(seqs_bytes - p)) : nullptr;
        const uint64_t e = nl ? (uint64_t)(nl - seqs) : seqs_bytes;
        if (!nl && i + 1 < n) { delete G; dcb_set_error("dcb_group_create: fewer than %llu lines", (unsigned long long)n); return DCB_EINVAL; }
        G->off[i] = p; G->len[i] = (uint32_t)(e - p);
        p = e + 1;
    }
    {   // distinct sequences
        std::unordered_map<std::string_view, uint32_t> ids;
        ids.reserve((size_t)n / 2 + 16);
        for (uint64_t i = 0; i < n; i++) {
            auto r = ids.emplace(std::string_view(seqs + G->off[i], G->len[i]), (uint32_t)G->rep.size());
            if (r.second) G->rep.push_back((uint32_t)i);
            G->seq_id[i] = r.first->second;
        }
    }
    G->order.resize(n);
    for (uint64_t i = 0; i < n; i++) G->order[i] = (uint32_t)i;
    std::sort(G->order.begin(), G->order.end(), [&](uint32_t a, uint32_t b) {
        return G->code[a] != G->code[b] ? G->code[a] < G->code[b] : G->idx[a] < G->idx[b];
    });
    for (uint64_t i = 0; i < n;) {
        uint64_t j = i + 1;
        while (j < n && G->code[G->order[j]] == G->code[G->order[i]]) j++;
        Group g;
        g.lo = (uint32_t)i; g.hi = (uint32_t)j;
        G->groups.push_back(std::move(g));
        i = j;
    }
    G->open.resize(G->groups.size());
    for (size_t k = 0; k < G->groups.size(); k++) G->open[k] = (uint32_t)k;
    *out = G;
    return DCB_OK;
}

/* One round: every unfinished barcode advances as far as the verdicts it has allow.  *n_pairs = verdicts to fetch before the
 * next round (0: the grouping is complete). */
int dcb_group_step(dcb_group* G, uint64_t* n_pairs) {
    if (!G || !n_pairs) { dcb_set_error("dcb_group_step: null argument"); return DCB_EINVAL; }
    G->asked.clear();
    std::vector<uint32_t> still;
    for (uint32_t k : G->open) if (!advance(*G, G->groups[k])) still.push_back(k);
    G->open.swap(still);
    *n_pairs = G->asked.size();
    return DCB_OK;
}

/* The pairs of the round as a compact batch for dcb_lev_leq: the distinct sequences involved, back to back in `symbols`
 * (codes 0..7 when they hold at most eight distinct characters, else the characters themselves), their offsets / lengths,
 * and per pair the two sequence numbers.  Sizes first (symbols == NULL), then the arrays. */
int dcb_group_pairs(dcb_group* G, uint64_t* n_seqs, uint64_t* n_symbols, uint8_t* symbols, uint64_t* off, uint32_t* len, uint32_t* a, uint32_t* b,
                    int* coded) {
    if (!G || !n_seqs || !n_symbols) { dcb_set_error("dcb_group_pairs: null argument"); return DCB_EINVAL; }
    std::unordered_map<uint32_t, uint32_t> local;
    std::vector<uint32_t> seqs;
    auto slot = [&](uint32_t id) {
        auto r = local.emplace(id, (uint32_t)seqs.size());
        if (r.second) seqs.push_back(id);
        return r.first->second;
    };
    std::vector<uint32_t> pa(G->asked.size()), pb(G->asked.size());
    for (size_t t = 0; t < G->asked.size(); t++) { pa[t] = slot((uint32_t)(G->asked[t] >> 32)); pb[t] = slot((uint32_t)G->asked[t]); }
    uint64_t total = 0;
    for (uint32_t id : seqs) total += G->len[G->rep[id]];
    *n_seqs = seqs.size(); *n_symbols = total;
    if (!symbols) return DCB_OK;
    if (!off || !len || !a || !b || !coded) { dcb_set_error("dcb_group_pairs: null argument"); return DCB_EINVAL; }
    bool seen[256] = {false};
    uint64_t at = 0;
    for (size_t s = 0; s < seqs.size(); s++) {
        const uint32_t row = G->rep[seqs[s]];
        off[s] = at; len[s] = G->len[row];
        std::memcpy(symbols + at, G->text + G->off[row], G->len[row]);
        for (uint32_t k = 0; k < G->len[row]; k++) seen[symbols[at + k]] = true;
        at += G->len[row];
    }
    int distinct = 0;
    uint8_t table[256];
    for (int c = 0; c < 256; c++) { table[c] = (uint8_t)c; if (seen[c]) { table[c] = (uint8_t)(distinct < 8 ? distinct : 0); distinct++; } }
    *coded = distinct <= 8;
    if (*coded) for (uint64_t k = 0; k < total; k++) symbols[k] = table[symbols[k]];
    std::memcpy(a, pa.data(), pa.size() * 4);
    std::memcpy(b, pb.data(), pb.size() * 4);
    return DCB_OK;
}

/* The verdicts of the round's pairs, in the order dcb_group_pairs gave them. */
int dcb_group_verdicts(dcb_group* G, const uint8_t* same, uint64_t n_pairs) {
    if (!G || (n_pairs && !same) || n_pairs != G->asked.size()) { dcb_set_error("dcb_group_verdicts: %llu verdicts for %zu pairs", (unsigned long long)n_pairs, G ? G->asked.size() : (size_t)0); return DCB_EINVAL; }
    for (size_t t = 0; t < G->asked.size(); t++) G->verdict[G->asked[t]] = same[t] ? 1 : 0;
    G->asked.clear();
    return DCB_OK;
}

/* The surviving groups in the reference's dict order (ascending tick): per group its barcode code, the row whose sequence is
 * the proto-sequence, and its members -- rows[first[g] .. first[g + 1]) in input order.  Sizes first (rows == NULL).
 * dropped: rows counted as multi_tcr_barcode_reads; dead: barcodes blacklisted. */
int dcb_group_result(dcb_group* G, uint64_t* n_groups, uint64_t* n_members, uint64_t* dropped, uint64_t* dead, uint64_t* code, uint64_t* tick,
                     uint32_t* proto_row, uint64_t* first, uint32_t* rows) {
    if (!G || !n_groups || !n_members || !dropped || !dead) { dcb_set_error("dcb_group_result: null argument"); return DCB_EINVAL; }
    if (!G->open.empty()) { dcb_set_error("dcb_group_result: the grouping is not complete"); return DCB_EINVAL; }
    std::vector<uint32_t> alive;
    uint64_t members = 0, dr = 0, dd = 0;
    for (size_t k = 0; k < G->groups.size(); k++) {
        const Group& g = G->groups[k];
        dr += g.dropped; dd += g.dead ? 1 : 0;
        if (!g.dead && g.founded) { alive.push_back((uint32_t)k); members += g.hi - g.lo; }
    }
    *n_groups = alive.size(); *n_members = members; *dropped = dr; *dead = dd;
    if (!rows) return DCB_OK;
    if (!code || !tick || !proto_row || !first) { dcb_set_error("dcb_group_result: null argument"); return DCB_EINVAL; }
    std::sort(alive.begin(), alive.end(), [&](uint32_t x, uint32_t y) { return G->groups[x].tick < G->groups[y].tick; });
    uint64_t at = 0;
    for (size_t s = 0; s < alive.size(); s++) {
        const Group& g = G->groups[alive[s]];
        code[s] = G->code[G->order[g.lo]];
        tick[s] = g.tick;
        proto_row[s] = G->rep[g.proto];
        first[s] = at;
        std::memcpy(rows + at, G->order.data() + g.lo, (size_t)(g.hi - g.lo) * 4);
        at += g.hi - g.lo;
    }
    first[alive.size()] = at;
    return DCB_OK;
}

void dcb_group_free(dcb_group* G) { delete G; }

}  // extern "C"
