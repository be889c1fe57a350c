// core_table.cpp - see core_table.h
#include "core_table.h"

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace scb {

std::string build_core_table(const std::vector<std::string> &cores_in, CoreTable &t) {
    t = CoreTable();
    const int32_t nc = (int32_t)cores_in.size();
    t.cores = cores_in;
    size_t maxlen = 0;
    for (auto &c : t.cores) {
        // pattern_insert stops at NUL or '\n' (reads.cpp:254)
        size_t nl = c.find('\n');
        if (nl != std::string::npos) c.resize(nl);
        if (c.empty()) return "empty core string";
        if (c.size() > 255) return "core longer than 255 bases";
        maxlen = std::max(maxlen, c.size());
    }
    t.max_level = (int32_t)maxlen;

    // ---- trie, level by level; ids come out in the reference's BFS order -------------------
    std::vector<int32_t> child(4, -1);  // child[state*4+c]
    std::vector<uint8_t> level(1, 0);
    std::vector<int32_t> cur(nc, 0);    // node of each core's prefix at the previous level
    std::vector<int32_t> active(nc);
    for (int32_t i = 0; i < nc; i++) active[i] = i;
    int32_t lvl_lo = 0, lvl_hi = 1;     // id range of the previous level
    for (size_t d = 1; d <= maxlen; d++) {
        std::vector<int32_t> slot((size_t)(lvl_hi - lvl_lo) * 4, -1);
        size_t w = 0;
        for (size_t k = 0; k < active.size(); k++) {
            int32_t ci = active[k];
            if (t.cores[ci].size() < d) continue;
            active[w++] = ci;
            slot[(size_t)(cur[ci] - lvl_lo) * 4 + getval((unsigned char)t.cores[ci][d - 1])] = 0;
        }
        active.resize(w);
        int32_t next_id = lvl_hi;
        for (size_t s = 0; s < slot.size(); s++)
            if (slot[s] == 0) {
                slot[s] = next_id++;
                child[(size_t)(lvl_lo + (int32_t)(s / 4)) * 4 + (s % 4)] = slot[s];
            }
        child.resize((size_t)next_id * 4, -1);
        level.resize((size_t)next_id, (uint8_t)d);
        for (int32_t ci : active)
            cur[ci] = slot[(size_t)(cur[ci] - lvl_lo) * 4 + getval((unsigned char)t.cores[ci][d - 1])];
        // cores ending at this level keep cur = their terminal node
        lvl_lo = lvl_hi;
        lvl_hi = next_id;
    }
    const int32_t ns = lvl_hi;
    t.n_states = ns;
    t.state_level = level;

    std::vector<int32_t> output(ns, -1);
    for (int32_t i = 0; i < nc; i++) output[cur[i]] = i;  // later duplicate overwrites (reads.cpp:264)

    // ---- fail links folded into the completed DFA, in id order ------------------------------
    t.next.assign((size_t)ns * 4, 0);
    std::vector<int32_t> fail(ns, 0), nto_node(ns, -1);
    for (int32_t u = 0; u < ns; u++) {
        for (int c = 0; c < 4; c++) {
            int32_t v = child[(size_t)u * 4 + c];
            uint32_t via_fail = (u == 0) ? 0u : t.next[(size_t)fail[u] * 4 + c];
            if (v >= 0) {
                fail[v] = (int32_t)via_fail;
                t.next[(size_t)u * 4 + c] = (uint32_t)v;
            } else {
                t.next[(size_t)u * 4 + c] = via_fail;
            }
        }
        nto_node[u] = output[u] >= 0 ? u : (u == 0 ? -1 : nto_node[fail[u]]);
    }

    // ---- buckets -----------------------------------------------------------------------------
    std::vector<int32_t> node_rank(ns, -1);
    for (int32_t u = 1; u < ns; u++)
        if (output[u] >= 0) {
            node_rank[u] = t.n_buckets++;
            t.rank_node_id.push_back(u);
            t.rank_core.push_back(output[u]);
            t.rank_level.push_back(level[u]);
        }
    t.nto_rank.assign(ns, -1);
    for (int32_t u = 0; u < ns; u++)
        if (nto_node[u] >= 0) t.nto_rank[u] = node_rank[nto_node[u]];
    t.core_to_rank.assign(nc, -1);
    for (int32_t i = 0; i < nc; i++) t.core_to_rank[i] = node_rank[cur[i]];

    // root emission position (reads.cpp:473-495)
    t.root_order_pos = t.n_buckets;
    t.root_counts_unbucketed = true;
    for (int c = 0; c < 4; c++)
        if (child[c] < 0) {  // root->child[c] is the root itself: it is dequeued among the level-1 nodes
            int32_t pos = 0;
            for (int c2 = 0; c2 < c; c2++)
                if (child[c2] >= 0 && output[child[c2]] >= 0) pos++;
            t.root_order_pos = pos;
            t.root_counts_unbucketed = false;
            break;
        }
    return "";
}

std::string load_core_file(const char *path, std::vector<std::string> &cores) {
    cores.clear();
    FILE *f = fopen(path, "rb");
    if (!f) return std::string("cannot open core file ") + path;
    std::vector<unsigned char> buf;
    unsigned char tmp[65536];
    size_t r;
    while ((r = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + r);
    fclose(f);
    bool text = true;
    for (unsigned char c : buf)
        if (!(c == '\n' || c == '\r' || c == '\t' || c == ' ' || (c >= 33 && c < 127))) { text = false; break; }
    if (text) {  // fscanf("%ms") tokens, reads.cpp:389
        size_t i = 0, n = buf.size();
        while (i < n) {
            while (i < n && (buf[i] == ' ' || buf[i] == '\n' || buf[i] == '\r' || buf[i] == '\t')) i++;
            size_t s = i;
            while (i < n && !(buf[i] == ' ' || buf[i] == '\n' || buf[i] == '\r' || buf[i] == '\t')) i++;
            if (i > s) cores.emplace_back((const char *)&buf[s], i - s);
        }
        return "";
    }
    // patterns.bin: {int16 len; int32 cnt; cnt x ceil(len/4) bytes LE, base j at bits 2(len-1-j)} reads.cpp:343-364
    static const char alpha[] = "ACGT";
    size_t pos = 0, size = buf.size();
    while (pos < size) {
        if (pos + 6 > size) return "truncated patterns.bin header";
        int16_t ln; int32_t cnt;
        memcpy(&ln, &buf[pos], 2); memcpy(&cnt, &buf[pos + 2], 4); pos += 6;
        int sz = ln / 4 + (ln % 4 != 0);
        if (ln <= 0 || sz > 8 || cnt < 0) return "bad patterns.bin record";
        if (pos + (size_t)sz * (size_t)cnt > size) return "truncated patterns.bin body";
        for (int32_t i = 0; i < cnt; i++) {
            uint64_t x = 0;
            memcpy(&x, &buf[pos], (size_t)sz); pos += (size_t)sz;
            std::string s((size_t)ln, 'A');
            for (int j = ln - 1, k = 0; j >= 0; j--, k++) s[(size_t)k] = alpha[(x >> (2 * j)) & 3];
            cores.push_back(s);
        }
    }
    return "";
}

}  // namespace scb
