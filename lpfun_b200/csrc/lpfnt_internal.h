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
constexpr int kGroupLanes = 8;   // consecutive lanes of a fibre that stay together in k_sweep_groups
constexpr int kLineChunk = 128;  // lines per CTA of k_sweep_lines
constexpr int kGroupClass = 4;  // lane groups of k_sweep_groups are ordered by height class of this many rows
constexpr int kTileRows = 256;  // lines staged in shared memory by one CTA of the tile kernels
constexpr int kTileItems = 256; // element-fibres per CTA (two per thread, 128 threads)
constexpr int kTileElems = 5120; // packed elements per CTA (40 KB; two data buffers + descriptor ring = 110 KB -> 2 CTAs/SM)

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
    // row tiles (see HostIndex::build): tile -> fibre range; per tile the element-fibres and rows
    // sorted by length.  item = row_rel | lane << 16 | t << 24 ; row0 = row_rel | len << 24
    std::vector<SortedItem> sorted;      // num_threads: element-fibres, longest first
    std::vector<uint8_t> group_t;        // 8 per group: height of every lane of the run
    std::vector<SortedItem> groups;      // lane runs {first ent entry, lane0 | t << 8 | lanes << 16}, tallest first
    std::vector<uint8_t> group_cta_cap;  // per 16 consecutive runs (one CTA of k_sweep_groups): tallest height
    std::vector<int32_t> tile_fib;       // num_tiles + 1
    std::vector<uint32_t> items;         // num_threads (indexed like the threads: thr_start[f0] + j)
    std::vector<uint32_t> rows0;         // N_1 (indexed like ent: fib_ptr[f0] + j)
    std::vector<uint16_t> row_pos;       // N_1: packed shared-memory offset of every row inside its tile
    std::vector<uint8_t> elem_row;       // N: row (relative to the tile) of every packed element
    std::vector<int32_t> tile_elem;      // num_tiles + 1: start of each tile in elem_row (4-aligned)
    std::vector<int32_t> tile_nelem;     // num_tiles: packed elements of each tile
};

// Host-side index structure derived from the multi-index array A (colex order).
struct HostIndex {
    int m = 0, n = 0;
    int64_t N = 0, N1 = 0;
    int32_t max_line = 0;
    std::vector<int32_t> line_off;       // N_1 + 1 (cs_T)
    std::vector<int64_t> level_size;     // m + 1: N_0, N_1, ..., N_{m-1}, 1
    std::vector<DimTable> dims;          // index h = 0..m-1 (entry 0 unused)
    std::vector<SortedItem> sorted0;     // N_1: the lines, longest first (coordinate 0)
    std::vector<int32_t> line_order;     // N_1: lines ordered by length inside every chunk of kLineChunk
    std::vector<SortedItem> eval_line;   // N_1: {padded coefficient offset, line length}
    std::vector<EvalBlock> eval_block;   // blocks of lines staged together by the eval kernel
    int64_t eval_padded = 0;             // size of the padded coefficient layout (doubles)
    std::vector<uint8_t> line_alpha;     // N_1 * (m-1): (a_1..a_{m-1}) of every line, for eval (n <= 255)
    // returns false (and sets the error) if A is not a colex-ordered downward-closed set
    bool build(const int64_t *A, int64_t N, int m, int n);
};

}  // namespace lpfnt
