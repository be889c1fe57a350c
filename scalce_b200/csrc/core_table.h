// core_table.h - host-side compiler from a core set to the flat automaton tables the scan
// kernels read. Replaces pattern_insert + prepare_aho_automata (reads.cpp:253-324) with a
// level-synchronous array construction (no pointer-linked nodes).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace scb {

struct CoreTable {
    // states are trie nodes numbered by the reference's BFS id (reads.cpp:296): root 0, then
    // (level asc, prefix lexicographic A<C<G<T). id == state index.
    int32_t n_states = 0;               // including root
    std::vector<uint32_t> next;         // [n_states*4] completed DFA (reads.cpp:298-310)
    std::vector<int32_t> nto_rank;      // [n_states] bucket rank of next_to_output, -1 if none (reads.cpp:292-293,311-313)
    std::vector<uint8_t> state_level;   // [n_states]
    // buckets: output nodes in id order. rank r in [0, n_buckets)
    int32_t n_buckets = 0;
    std::vector<int32_t> rank_node_id;  // BFS id, written as meta "id" (reads.cpp:166)
    std::vector<int32_t> rank_core;     // core index = aho_trie::output ("last duplicate wins", reads.cpp:264)
    std::vector<uint8_t> rank_level;    // core length
    // emission order of the root (no-core) bucket among the buckets: n_buckets (last) unless some
    // base starts no core, in which case aho_output meets the root early (reads.cpp:473-476)
    int32_t root_order_pos = 0;
    bool root_counts_unbucketed = true; // reads.cpp:491-495 only runs when the root was not met early
    int32_t max_level = 0;
    std::vector<std::string> cores;     // patterns[] (reads.h:48)
    std::vector<int32_t> core_to_rank;  // core index -> rank (duplicates share the rank)
};

// 2-bit code of an ASCII base, const.cpp:47-49 / const.h:127 (out-of-table bytes defined as 0).
inline int getval(unsigned char c) {
    c |= 0x20;
    return c == 'c' ? 1 : c == 'g' ? 2 : c == 't' ? 3 : 0;
}

// Builds the table. Returns empty string on success, else an error message.
std::string build_core_table(const std::vector<std::string> &cores, CoreTable &out);

// Loads a core set file: text (-P, reads.cpp:388-394) or patterns.bin (reads.cpp:338-369).
std::string load_core_file(const char *path, std::vector<std::string> &cores);

}  // namespace scb
