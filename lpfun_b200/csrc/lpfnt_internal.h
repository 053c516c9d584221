// Internal declarations shared by the translation units of liblpfnt.so.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/lpfnt.h"

namespace lpfnt {

constexpr int kMaxDim = 12;     // spatial dimensions supported by the fast kernels' line descriptors
constexpr int kMaxRegT = 32;    // longest fibre held in registers by the unrolled kernels (n+1 <= 32)
constexpr int kEvalBlockLines = 256;  // lines per shared-memory block of the eval kernel
constexpr int kEvalBlockElems = 2048;  // padded doubles per block (16 KB)
#ifndef LPFNT_GROUP_LANES
#define LPFNT_GROUP_LANES 8
#endif
constexpr int kGroupLanes = LPFNT_GROUP_LANES;   // consecutive lanes of a fibre that stay together in k_sweep_groups
constexpr int kLineChunk = 128;  // lines per CTA of k_sweep_lines
constexpr int kGroupClass = 4;  // lane groups of k_sweep_groups are ordered by height class of this many rows
constexpr int kWholeSmemLimit = 227 * 1024;   // one CTA per SM
constexpr int kWholeSmemPerCta = 113 * 1024;  // two CTAs per SM (each reserves 1 KB besides its dynamic shared memory)

void set_error(const std::string &msg);

struct SortedItem {  // globally length-sorted work item: x = first entry in ent (h >= 1) or line offset (h = 0); y = lane | t << 8
    int32_t x;
    int32_t y;
};

struct EvalBlock {  // {first line, #lines, padded offset of the block, padded size in doubles}
    int32_t l0, nl, poff, npad;
};

struct FibEnt {  // one line of a fibre: element offset of the line start and its length
    int32_t off;
    int32_t len;
};

// Per-coordinate (h >= 1) grouping of the lines into fibres (lines that differ only in a_h,
// ordered by a_h).  One GPU thread owns one element-fibre = (fibre, lane a_0).
struct DimTable {
    int32_t num_fibres = 0;
    int64_t num_threads = 0;             // sum over fibres of the length of their first line
    int32_t max_lines = 0;               // longest fibre (in lines)
    std::vector<int32_t> fib_ptr;        // num_fibres + 1, CSR into ent
    std::vector<FibEnt> ent;             // N_1 entries
    std::vector<int32_t> thr_start;      // num_fibres + 1, prefix sum of lanes per fibre
    std::vector<int32_t> thr_fib;        // num_threads, thread -> fibre
    std::vector<uint8_t> group_t;        // 8 per group: height of every lane of the run
    std::vector<SortedItem> groups;      // lane runs {first ent entry, lane0 | t << 8 | lanes << 16}, tallest first
    std::vector<uint8_t> group_cta_cap;  // per 16 consecutive runs (one CTA of k_sweep_groups): tallest height
};

// Item tables of the whole-function kernel (k_whole, whole_kernel.cuh), built from the index structure alone.
// Per coordinate one list of 32-bit words: bits 0..15 = element offset of the line (coordinate 0) or first entry of the
// fibre in seo (coordinate h >= 1: element k of the item sits at seo[entry + k] + lane), bits 16..21 = length t,
// bits 22..31 = lane a_0; already in THREAD order: rounds of `threads` items, inside a round chunks of 32 items dealt
// to the warps in snake order over the four SM sub-partitions.  Empty slots are 0 (t = 0).
struct WholeHost {
    bool ok = false;
    int threads = 0;                       // threads per CTA the lists were laid out for (384 or 768)
    int fibres = 1;                        // items per thread and round
    int32_t item_off[3] = {0, 0, 0}, n_items[3] = {0, 0, 0};
    int32_t max_t = 0, max_items = 0;
    std::vector<uint32_t> items;
    std::vector<uint16_t> seo;
};

// Tables of the split whole-function kernel (see HostIndex::build_whole_split).
struct WholeSplitHost {
    bool ok = false;
    int C = 0;                              // planes a_2 < C form part P, the others part Q
    int32_t e0[2] = {0, 0}, ne[2] = {0, 0}; // element range of the parts
    int32_t item_off[2][3] = {{0, 0, 0}, {0, 0, 0}}, n_items[2][3] = {{0, 0, 0}, {0, 0, 0}};
    int32_t cb[kMaxRegT] = {0}, rq[kMaxRegT] = {0};  // columns of Q: cb[a_1] + a_0 for a_0 < rq[a_1]
    int32_t ncolq = 0, max_t = 0;
    std::vector<uint32_t> items;
    std::vector<uint16_t> seo;
};

// Host-side index structure derived from the multi-index array A (colex order).
struct HostIndex {
    int m = 0, n = 0;
    int64_t N = 0, N1 = 0;
    int32_t max_line = 0;
    std::vector<int32_t> line_off;       // N_1 + 1 (cs_T)
    std::vector<int64_t> level_size;     // m + 1: N_0, N_1, ..., N_{m-1}, 1
    std::vector<DimTable> dims;          // index h = 0..m-1 (entry 0 unused)
    std::vector<uint8_t> line_last;      // N_1: a_{m-1} of every line (m >= 2, n <= 255)
    WholeHost whole;
    WholeSplitHost wsplit;
    std::vector<int32_t> line_order;     // N_1: lines ordered by length inside every chunk of kLineChunk
    std::vector<SortedItem> eval_line;   // N_1: {padded coefficient offset, line length}
    std::vector<SortedItem> eval_tline;  // N_1: k_eval_tree descriptors {byte offset in the staged block, steps | 8 a_1 << 8 | depth << 16}
    std::vector<EvalBlock> eval_block;   // blocks of lines staged together by the eval kernel
    int64_t eval_padded = 0;             // size of the padded coefficient layout (doubles)
    std::vector<uint8_t> line_alpha;     // N_1 * (m-1): (a_1..a_{m-1}) of every line, for eval (n <= 255)
    // returns false (and sets the error) if A is not a colex-ordered downward-closed set
    bool build(const int64_t *A, int64_t N, int m, int n);
    // (re)build `whole` for CTAs of `threads` threads (384 or 768) owning `fibres` items each per round; order = 2 bits per coordinate: 0 strict length order,
    // 1 / 2 length classes of 8 / 16 (natural order inside a class), 3 natural; bit 6: chunks dealt to the warps by the
    // pipe-sharing model instead of the snake order
    void build_whole(int threads, int order, int fibres);
    // (re)build `wsplit` (m = 3): C = split plane (<= 0: the one that balances the parts)
    void build_whole_split(int C, int order);
};
constexpr int kWholeDefaultOrder = 0 | (2 << 2) | (2 << 4) | 64 | 128;  // + 128: coordinate 0's chunks dealt to the warps in reverse

}  // namespace lpfnt
