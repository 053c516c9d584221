// Host-side index structure of the l^p lower set: enumeration, line lengths, trie levels,
// per-coordinate fibre tables.  Pure C++ (no CUDA); exported through the C ABI in include/lpfnt.h.
//
// Reference behaviour restated here (paths relative to the reference repository):
//   lp_set / _lp_set      src/lpfun/core/set.py:12-74
//   lp_tube / _lp_tube    src/lpfun/core/set.py:80-107
//   entropy               src/lpfun/core/set.py:167-182
//   ordinal_embedding     src/lpfun/core/set.py:208-239
// The reference never materialises fibre tables: it either exploits the coordinate symmetry of
// the set (2d/3d kernels), recomputes embeddings inside the loops (md kernels,
// src/lpfun/core/atoms.py:1060-1095) or permutes the whole vector (dx, molecules.py:298-308).
// Here the tables are built ONCE per plan from A alone, without any symmetry assumption.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "lpfnt_internal.h"

namespace lpfnt {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }

namespace {

// Odometer over N^m in colexicographic order with the reference's membership rule.
// The running sum of a_j^p is updated exactly like set.py:31-34/37-39 (subtract the old digit
// power, add the new one), a failing candidate is re-tested with a freshly accumulated sum
// (set.py:46-47) and, if it fails again, digit j is reset and the carry moves on (set.py:52-55).
class LpOdometer {
  public:
    LpOdometer(int m, int n, double p) : m_(m), n_(n), p_(p), digit_(m, 0), digit_pow_(m, 0.0), pow_(n + 1) {
        for (int k = 0; k <= n; ++k) pow_[k] = std::pow(static_cast<double>(k), p);
        limit_ = std::pow(static_cast<double>(n), p);
    }
    const std::vector<int64_t> &digits() const { return digit_; }

    // Moves to the next member; returns false when the enumeration is exhausted.
    bool next() {
        int j = 0;
        for (;;) {
            // increment digit j, carrying over saturated digits
            while (true) {
                if (j >= m_) return false;
                drop(j);
                if (digit_[j] < n_) {
                    digit_[j] += 1;
                    digit_pow_[j] = pow_[digit_[j]];
                    sum_ = sum_ + digit_pow_[j];
                    break;
                }
                digit_[j] = 0;
                digit_pow_[j] = 0.0;
                ++j;
            }
            if (sum_ <= limit_) return true;
            sum_ = fresh_sum();
            if (sum_ <= limit_) return true;
            drop(j);
            digit_[j] = 0;
            digit_pow_[j] = 0.0;
            ++j;
        }
    }

  private:
    void drop(int j) { sum_ = sum_ - digit_pow_[j]; }
    double fresh_sum() const {
        double s = 0.0;
        for (int q = 0; q < m_; ++q) s += std::pow(static_cast<double>(digit_[q]), p_);
        return s;
    }
    int m_, n_;
    double p_, limit_ = 0.0, sum_ = 0.0;
    std::vector<int64_t> digit_;
    std::vector<double> digit_pow_, pow_;
};

int64_t ipow_checked(int64_t base, int e) {
    int64_t r = 1;
    for (int i = 0; i < e; ++i) {
        if (r > (INT64_MAX / base)) return -1;
        r *= base;
    }
    return r;
}

int64_t enumerate(int m, int n, double p, int64_t *A, int64_t cap) {
    if (m == 1) {  // set.py:68-69
        if (A) {
            if (cap < n + 1) return -1;
            for (int k = 0; k <= n; ++k) A[k] = k;
        }
        return n + 1;
    }
    if (std::isinf(p)) {  // set.py:70-73: full tensor grid, coordinate 0 fastest
        int64_t N = ipow_checked(n + 1, m);
        if (N < 0) return -1;
        if (A) {
            if (cap < N) return -1;
            std::vector<int64_t> d(m, 0);
            for (int64_t r = 0; r < N; ++r) {
                for (int j = 0; j < m; ++j) A[r * m + j] = d[j];
                for (int j = 0; j < m; ++j) {
                    if (++d[j] <= n) break;
                    d[j] = 0;
                }
            }
        }
        return N;
    }
    LpOdometer od(m, n, p);
    int64_t rows = 0;
    do {
        if (A) {
            if (rows >= cap) return -1;
            std::memcpy(A + rows * m, od.digits().data(), sizeof(int64_t) * m);
        }
        ++rows;
    } while (od.next());
    return rows;
}

}  // namespace

bool HostIndex::build(const int64_t *A, int64_t N_, int m_, int n_) {
    m = m_;
    n = n_;
    N = N_;
    if (m < 1 || N < 1 || n < 0) {
        set_error("HostIndex: invalid sizes");
        return false;
    }
    if (N >= (int64_t(1) << 31) - 1) {
        set_error("HostIndex: N exceeds the 32-bit offset range of the device tables");
        return false;
    }
    // ---- lines (runs of coordinate 0) ----
    line_off.clear();
    for (int64_t e = 0; e < N; ++e) {
        const int64_t a0 = A[e * m];
        if (a0 == 0) line_off.push_back(static_cast<int32_t>(e));
        else if (line_off.empty() || a0 != e - line_off.back()) {
            set_error("HostIndex: A is not colex ordered (coordinate 0 must count up along each line)");
            return false;
        }
    }
    N1 = static_cast<int64_t>(line_off.size());
    line_off.push_back(static_cast<int32_t>(N));
    max_line = 0;
    for (int64_t l = 0; l < N1; ++l) max_line = std::max(max_line, line_off[l + 1] - line_off[l]);
    if (max_line > n + 1) {
        set_error("HostIndex: a line is longer than n+1");
        return false;
    }

    // ---- trie levels: a level-k node (k = 1..m-1) is a distinct suffix (a_k..a_{m-1});
    //      level 1 = lines.  node[k][l] = level-k ancestor of line l. ----
    auto key = [&](int64_t l, int j) -> int64_t { return A[int64_t(line_off[l]) * m + j]; };  // a_j of line l, j >= 1
    std::vector<std::vector<int32_t>> node(m + 1), first_child(m + 1), nchild(m + 1);
    level_size.assign(m + 1, 1);
    level_size[0] = N;
    if (m >= 2) level_size[1] = N1;
    for (int k = 2; k <= m - 1; ++k) {
        node[k].resize(N1);
        int32_t cur = -1;
        for (int64_t l = 0; l < N1; ++l) {
            bool starts = true;
            for (int j = 1; j < k; ++j)
                if (key(l, j) != 0) { starts = false; break; }
            if (starts) ++cur;
            node[k][l] = cur;
        }
        level_size[k] = cur + 1;
    }
    // first_child[k][v]: index (at level k-1) of the first child of level-k node v, k = 2..m-1
    for (int k = 2; k <= m - 1; ++k) {
        first_child[k].assign(level_size[k], -1);
        nchild[k].assign(level_size[k], 0);
        for (int64_t l = 0; l < N1; ++l) {
            const int32_t v = node[k][l];
            const int32_t child = (k == 2) ? static_cast<int32_t>(l) : node[k - 1][l];
            if (first_child[k][v] < 0) first_child[k][v] = child;
            nchild[k][v] = child - first_child[k][v] + 1;
        }
    }
    // children of the root (level m) are the level-(m-1) nodes 0..level_size[m-1]-1

    auto level_node = [&](int k, int64_t l) -> int32_t {  // level-k ancestor of line l (k = 1..m-1)
        return k == 1 ? static_cast<int32_t>(l) : node[k][l];
    };

    // ---- per-coordinate fibre tables ----
    dims.assign(m, DimTable());
    std::vector<int32_t> nxt(N1);
    for (int h = 1; h < m; ++h) {
        DimTable &D = dims[h];
        std::fill(nxt.begin(), nxt.end(), -1);
        int32_t roots = 0;
        for (int64_t l = 0; l < N1; ++l) {
            const int64_t ah = key(l, h);
            if (ah == 0) { ++roots; continue; }
            // locate the line with a_h - 1: walk down from the level-(h+1) ancestor
            int32_t c;
            if (h + 1 <= m - 1) {
                const int32_t anc = level_node(h + 1, l);
                if (ah - 1 >= nchild[h + 1][anc]) { set_error("HostIndex: set is not downward closed"); return false; }
                c = first_child[h + 1][anc] + static_cast<int32_t>(ah - 1);
            } else {
                c = static_cast<int32_t>(ah - 1);  // child of the root
            }
            for (int j = h - 1; j >= 1; --j) {  // c is a level-(j+1) node; descend to its child a_j
                const int64_t aj = key(l, j);
                if (aj >= nchild[j + 1][c]) { set_error("HostIndex: set is not downward closed"); return false; }
                c = first_child[j + 1][c] + static_cast<int32_t>(aj);
            }
            // c is now a line index
            if (c < 0 || c >= N1 || nxt[c] != -1) { set_error("HostIndex: inconsistent fibre structure"); return false; }
            if (line_off[c + 1] - line_off[c] < line_off[l + 1] - line_off[l]) {
                set_error("HostIndex: line lengths must not increase along a fibre");
                return false;
            }
            nxt[c] = static_cast<int32_t>(l);
        }
        D.num_fibres = roots;
        D.fib_ptr.assign(roots + 1, 0);
        D.thr_start.assign(roots + 1, 0);
        D.ent.resize(N1);
        int32_t f = 0, pos = 0;
        int64_t thr = 0;
        D.max_lines = 0;
        for (int64_t l = 0; l < N1; ++l) {
            if (key(l, h) != 0) continue;
            D.fib_ptr[f] = pos;
            D.thr_start[f] = static_cast<int32_t>(thr);
            thr += line_off[l + 1] - line_off[l];
            int32_t cnt = 0;
            for (int32_t q = static_cast<int32_t>(l); q >= 0; q = nxt[q]) {
                D.ent[pos++] = FibEnt{line_off[q], line_off[q + 1] - line_off[q]};
                ++cnt;
            }
            D.max_lines = std::max(D.max_lines, cnt);
            ++f;
        }
        D.fib_ptr[f] = pos;
        D.thr_start[f] = static_cast<int32_t>(thr);
        if (pos != N1) { set_error("HostIndex: fibre tables do not cover all lines"); return false; }
        D.num_threads = thr;
        D.thr_fib.resize(thr);
        for (int32_t g = 0; g < roots; ++g)
            for (int32_t t = D.thr_start[g]; t < D.thr_start[g + 1]; ++t) D.thr_fib[t] = g;
    }


    // ---- lane groups (k_sweep_groups): every fibre's lanes are cut into runs of <= kGroupLanes
    //      consecutive lanes; the runs are ordered by height (longest element-fibre first), so that
    //      the four runs sharing a warp execute triangles of similar size while every run still reads
    //      and writes >= 64 contiguous bytes per row. ----
    for (int h = 1; h < m; ++h) {
        DimTable &D = dims[h];
        D.groups.clear();
        if (D.max_lines > 255 || max_line > 255) continue;
        for (int32_t f = 0; f < D.num_fibres; ++f) {
            const int32_t p0 = D.fib_ptr[f], nl = D.fib_ptr[f + 1] - p0;
            const int32_t lanes = D.thr_start[f + 1] - D.thr_start[f];
            for (int32_t l0 = 0; l0 < lanes; l0 += kGroupLanes) {
                int32_t t = 0;  // height of the group's first (tallest) lane
                while (t < nl && D.ent[p0 + t].len > l0) ++t;
                D.groups.push_back(SortedItem{p0, l0 | (t << 8) | (std::min(kGroupLanes, lanes - l0) << 16)});
            }
        }
        // by height CLASS (kGroupClass rows), natural order inside a class: the runs sharing a warp are then mostly
        // neighbours in memory (one 256-byte piece per row instead of four 64-byte pieces) -- same trade as in k_whole2
        std::stable_sort(D.groups.begin(), D.groups.end(), [](const SortedItem &a, const SortedItem &b) {
            return (((a.y >> 8) & 0xff) + kGroupClass - 1) / kGroupClass > (((b.y >> 8) & 0xff) + kGroupClass - 1) / kGroupClass;
        });
        constexpr size_t kRunsPerCta = 128 / kGroupLanes;  // one CTA of k_sweep_groups = 128 threads
        D.group_cta_cap.assign((D.groups.size() + kRunsPerCta - 1) / kRunsPerCta, 0);
        for (size_t g = 0; g < D.groups.size(); ++g)
            D.group_cta_cap[g / kRunsPerCta] = std::max<uint8_t>(D.group_cta_cap[g / kRunsPerCta], uint8_t((D.groups[g].y >> 8) & 0xff));
        // height of every lane of every group (8 bytes per group, 0 for absent lanes)
        D.group_t.assign(D.groups.size() * kGroupLanes, 0);
        for (size_t g = 0; g < D.groups.size(); ++g) {
            const int32_t p0 = D.groups[g].x, l0 = D.groups[g].y & 0xff, nl = (D.groups[g].y >> 8) & 0xff, lanes = (D.groups[g].y >> 16) & 0xff;
            for (int32_t sub = 0; sub < lanes; ++sub) {
                int32_t t = 0;
                while (t < nl && D.ent[p0 + t].len > l0 + sub) ++t;
                D.group_t[g * kGroupLanes + sub] = static_cast<uint8_t>(t);
            }
        }
    }
    // lines ordered by length inside every chunk of kLineChunk consecutive lines (k_sweep_lines)
    line_order.resize(N1);
    for (int64_t c0 = 0; c0 < N1; c0 += kLineChunk) {
        const int64_t c1 = std::min<int64_t>(N1, c0 + kLineChunk);
        for (int64_t l = c0; l < c1; ++l) line_order[l] = static_cast<int32_t>(l);
        std::stable_sort(line_order.begin() + c0, line_order.begin() + c1, [&](int32_t a, int32_t b) {
            return (line_off[a + 1] - line_off[a]) > (line_off[b + 1] - line_off[b]);
        });
    }

    // ---- eval program (k_eval_tree): the coefficient vector is re-laid out per call with every line
    //      padded to a multiple of 4 doubles (32-byte aligned, zero filled), and the lines are cut into
    //      blocks of <= kEvalBlockLines lines / <= kEvalBlockElems padded doubles that a CTA stages in
    //      shared memory.  eval_line = {padded offset, t}; eval_block = {first line, #lines, padded
    //      offset of the block, padded size}. ----
    eval_line.clear();
    eval_block.clear();
    eval_padded = 0;
    if (n <= 255) {
        eval_line.resize(N1);
        int64_t pos = 0;
        for (int64_t l = 0; l < N1; ++l) {
            const int32_t t = line_off[l + 1] - line_off[l];
            eval_line[l] = SortedItem{static_cast<int32_t>(pos), t};
            pos += (t + 3) & ~3;
        }
        eval_padded = pos;
        int64_t l0 = 0;
        while (l0 < N1) {
            int64_t l1 = l0;
            int64_t sz = 0;
            while (l1 < N1 && l1 - l0 < kEvalBlockLines) {
                const int64_t add = ((line_off[l1 + 1] - line_off[l1]) + 3) & ~3;
                if (sz + add > kEvalBlockElems) break;
                sz += add;
                ++l1;
            }
            eval_block.push_back(EvalBlock{static_cast<int32_t>(l0), static_cast<int32_t>(l1 - l0), eval_line[l0].x, static_cast<int32_t>(sz)});
            l0 = l1;
        }
    }

    // ---- line descriptors for eval ----
    line_alpha.clear();
    if (n <= 255 && m >= 2) {
        line_alpha.resize(size_t(N1) * (m - 1));
        for (int64_t l = 0; l < N1; ++l)
            for (int j = 1; j < m; ++j) line_alpha[l * (m - 1) + (j - 1)] = static_cast<uint8_t>(key(l, j));
    }
    // k_eval_tree descriptors: x = byte offset of the line inside its staged block, y = ceil(t / 4) | 8 a_1 << 8 | depth << 16 with
    // depth = number of levels the line completes (leading zeros of (a_1, a_2, ..)): the common case -- fold into level 1
    // only -- needs no alpha look-up
    eval_tline.clear();
    if (n < 32 && !eval_line.empty()) {
        eval_tline.resize(N1);
        for (const EvalBlock &b : eval_block)
            for (int64_t l = b.l0; l < int64_t(b.l0) + b.nl; ++l) {
                const int32_t t = line_off[l + 1] - line_off[l];
                int depth = 0;
                while (depth < m - 1 && key(l, 1 + depth) == 0) ++depth;
                const int32_t a1 = m >= 2 ? int32_t(key(l, 1)) : 0;
                eval_tline[l] = SortedItem{(eval_line[l].x - b.poff) * 8, ((t + 3) >> 2) | (a1 * 8) << 8 | depth << 16};
            }
    }
    // ---- last coordinate of every line + the item tables of the whole-function kernel ----
    line_last.clear();
    whole = WholeHost();
    if (n <= 255 && m >= 2) {
        line_last.resize(size_t(N1));
        for (int64_t l = 0; l < N1; ++l) line_last[l] = static_cast<uint8_t>(key(l, m - 1));
        build_whole(768, kWholeDefaultOrder, 1);
    }
    return true;
}

namespace {
struct WholeItem { int t; uint32_t w; };

// Orders the items of one sweep (length classes, natural order inside a class) and appends them to `out` in THREAD order:
// rounds of TG * NF slots, thread i owning slots i, TG + i; inside a round chunks of 32 items dealt to the warps.
// Returns the padded list length.
int32_t layout_whole_items(std::vector<WholeItem> &nat, int order_code, bool model, int TG, int NF, std::vector<uint32_t> &out, bool reverse = false) {
    const int RT = TG * NF;
    const int cls[4] = {1, 8, 16, 1 << 20};
    const int q = cls[order_code & 3];
    std::stable_sort(nat.begin(), nat.end(), [q](const WholeItem &a, const WholeItem &b) { return (a.t + q - 1) / q > (b.t + q - 1) / q; });
    const int32_t nit = int32_t(nat.size()), padded = int32_t((nit + RT - 1) / RT) * RT;
    const int nw = TG / 32, CH = 32 * NF;
    for (int32_t r0 = 0; r0 < padded; r0 += RT) {
        // chunks of CH items (longest first); warp w runs on SM sub-partition w % 4
        const int nch = RT / CH;
        std::vector<double> wgt(nch, 0.0);
        for (int c = 0; c < nch; ++c) {
            int tm = 0;
            for (int i = 0; i < CH; ++i) { const int32_t src = r0 + c * CH + i; if (src < nit) tm = std::max(tm, nat[src].t); }
            const int tt = (tm + 1) & ~1;
            wgt[c] = 0.5 * tt * (tt + 1);
        }
        // a warp's unit of work: one chunk
        std::vector<double> unit(nw, 0.0);
        for (int c = 0; c < nw; ++c) unit[c] = wgt[c];
        std::vector<int> warp_of(nw);  // unit -> warp
        // start: SNAKE order over the four sub-partitions (0 1 2 3 / 3 2 1 0 / ...): equal sums of work
        for (int c = 0; c < nw; ++c) { const int grp = c >> 2; warp_of[c] = grp * 4 + ((grp & 1) ? 3 - (c & 3) : (c & 3)); }
        if (model) {
            // The FP64 pipe of a sub-partition is shared by its warps, and one or two warps alone cannot keep it busy
            // (tools/probe5.cu: equal sums but unequal chunks cost 17-25 % of a sweep): the sweep ends when the
            // sub-partition whose longest unit outlasts its other units by most is done.  Model: processor sharing
            // with rate r(j) of the pipe while j warps are still computing; hill-climb over swaps.
            const double r1 = NF == 2 ? 0.70 : 0.45, r2 = NF == 2 ? 0.88 : 0.75, r3 = NF == 2 ? 0.93 : 0.90, r4 = 0.93;
            auto part_time = [&](int sp) {
                double a[64]; int k = 0;
                for (int c = 0; c < nw; ++c) if ((warp_of[c] & 3) == sp && unit[c] > 0) a[k++] = unit[c];
                std::sort(a, a + k);  // ascending: a[0] finishes first
                double tsum = 0, prev = 0;
                for (int i = 0; i < k; ++i) {
                    const int act = k - i;
                    const double r = act == 1 ? r1 : act == 2 ? r2 : act == 3 ? r3 : r4;
                    tsum += (a[i] - prev) * act / r;
                    prev = a[i];
                }
                return tsum;
            };
            auto cost = [&]() { double mx = 0, sm = 0; for (int sp = 0; sp < 4; ++sp) { const double t = part_time(sp); mx = std::max(mx, t); sm += t; } return mx + 1e-3 * sm; };
            double best = cost();
            for (bool improved = true; improved;) {
                improved = false;
                for (int c1 = 0; c1 < nw; ++c1)
                    for (int c2 = c1 + 1; c2 < nw; ++c2) {
                        if ((warp_of[c1] & 3) == (warp_of[c2] & 3)) continue;
                        std::swap(warp_of[c1], warp_of[c2]);
                        const double t = cost();
                        if (t < best - 1e-9) { best = t; improved = true; }
                        else std::swap(warp_of[c1], warp_of[c2]);
                    }
            }
        }
        // reverse: the warps that carry the longest items of the other sweeps get the shortest ones here (coordinate 0 follows
        // the previous function's LAST sweep without a barrier: the warps that leave it first then start on the long items)
        if (reverse)
            for (int c = 0; c < nw; ++c) warp_of[c] = nw - 1 - warp_of[c];
        std::vector<int> unit_of(nw);
        for (int c = 0; c < nw; ++c) unit_of[warp_of[c]] = c;
        for (int slot = 0; slot < RT; ++slot) {
            // side by side: lane l of the warp owns items l and 32 + l of its chunk (neighbours in the length order)
            const int f = slot / TG, tid = slot % TG;
            const int u = unit_of[tid >> 5];
            const int32_t src = r0 + u * CH + f * 32 + (tid & 31);
            out.push_back(src < nit ? nat[src].w : 0u);
        }
    }
    return padded;
}
}  // namespace


void HostIndex::build_whole(int threads, int order, int fibres) {
    whole = WholeHost();
    WholeHost &W = whole;
    if (m < 2 || m > 3 || N >= 65536 || max_line > kMaxRegT || (threads != 384 && threads != 768) || fibres < 1 || fibres > 2) return;
    for (int h = 1; h < m; ++h)
        if (dims[h].max_lines > kMaxRegT) return;
    // fibres: 1 or 2 (two items of the same length class per thread, folded side by side)
    const int TG = threads, NF = fibres == 1 ? 1 : 2, RT = TG * NF;  // a round = RT items: thread i owns slots i, TG + i, ...
    W.threads = threads;
    W.fibres = NF;
    auto rounds = [&](int64_t items) { return int32_t((items + RT - 1) / RT) * RT; };
    // seo: per coordinate h >= 1 the line offsets in fibre order; every fibre's list starts on a multiple of four entries
    // (8 bytes: the kernel fetches four offsets per load) and the table ends with a full row of padding (the kernel reads
    // whole groups of four / eight entries, also past a short fibre's end)
    std::vector<std::vector<int32_t>> seo_start(m);
    for (int h = 1; h < m; ++h) {
        const DimTable &D = dims[h];
        seo_start[h].resize(D.num_fibres);
        for (int32_t f = 0; f < D.num_fibres; ++f) {
            while (W.seo.size() % 4) W.seo.push_back(0);
            seo_start[h][f] = int32_t(W.seo.size());
            for (int32_t r = D.fib_ptr[f]; r < D.fib_ptr[f + 1]; ++r) W.seo.push_back(uint16_t(D.ent[r].off));
        }
    }
    for (int q = 0; q < kMaxRegT + 8 || W.seo.size() % 8; ++q) W.seo.push_back(0);
    if (W.seo.size() > 65536) return;
    for (int h = 0; h < m; ++h) {
        std::vector<WholeItem> nat;
        if (h == 0) {
            for (int64_t l = 0; l < N1; ++l) {
                const int t = line_off[l + 1] - line_off[l];
                nat.push_back({t, uint32_t(line_off[l]) | (uint32_t(t) << 16)});
            }
        } else {
            const DimTable &D = dims[h];
            for (int32_t f = 0; f < D.num_fibres; ++f) {
                const int32_t p0 = D.fib_ptr[f], nl = D.fib_ptr[f + 1] - p0;
                const int32_t lanes = D.thr_start[f + 1] - D.thr_start[f];
                for (int32_t lane = 0; lane < lanes; ++lane) {
                    int32_t t = 0;
                    while (t < nl && D.ent[p0 + t].len > lane) ++t;
                    nat.push_back({t, uint32_t(seo_start[h][f]) | (uint32_t(t) << 16) | (uint32_t(lane) << 22)});
                }
            }
        }
        const int32_t nit = int32_t(nat.size());
        for (const WholeItem &it : nat) W.max_t = std::max(W.max_t, it.t);
        W.item_off[h] = int32_t(W.items.size());
        W.n_items[h] = layout_whole_items(nat, (order >> (2 * h)) & 3, (order & 64) != 0, TG, NF, W.items, h == 0 && (order & 128) != 0);
        W.max_items = std::max(W.max_items, nit);
    }
    W.ok = true;
}

// Tables of the split whole-function kernel (k_whole_split, m = 3): the function is processed as two units,
//   part P = planes a_2 <  C  (elements [0, e0[1]): contiguous), part Q = planes a_2 >= C,
// so that TWO CTAs fit on an SM.  Coordinates 0 and 1 never couple different planes and run inside each part.  The last
// coordinate: unit P folds the rows k < C of every column (a_0, a_1); unit Q folds the rows k >= C of the columns that
// reach beyond C from the very first term -- the values y[j], j < C, it needs come from unit P's coordinate-1 results,
// exported to a scratch buffer [a_2][column] in global memory (columns of Q numbered a_1-major: cb[a_1] + a_0, a_0 < rq[a_1]).
// Item word: bits 0..15 line offset inside the part (coordinate 0) / first seo entry, 16..21 length, 22..26 lane a_0,
// 27..31 the plane a_2 (coordinate 1) / the row a_1 (coordinate 2).
void HostIndex::build_whole_split(int C, int order) {
    wsplit = WholeSplitHost();
    WholeSplitHost &W = wsplit;
    if (m != 3 || N >= 65536 || max_line > kMaxRegT || n + 1 > kMaxRegT || line_alpha.empty()) return;
    for (int h = 1; h < m; ++h)
        if (dims[h].max_lines > kMaxRegT) return;
    const int n1 = n + 1;
    auto A1 = [&](int64_t l) { return int(line_alpha[l * 2]); };
    auto A2 = [&](int64_t l) { return int(line_alpha[l * 2 + 1]); };
    auto LEN = [&](int64_t l) { return line_off[l + 1] - line_off[l]; };
    // lines of a plane are consecutive (colex) with a_1 ascending from 0; planes ascending in a_2
    std::vector<int64_t> plane0(n1 + 1, N1);  // first line of every plane
    int planes = 0;
    for (int64_t l = 0; l < N1; ++l) {
        if (A1(l) == 0) { if (A2(l) != planes) return; plane0[planes++] = l; }
        else if (l == 0 || A2(l) != A2(l - 1) || A1(l) != A1(l - 1) + 1) return;
    }
    plane0[planes] = N1;
    if (C <= 0) {  // automatic: the split that balances the two parts
        int64_t best = -1;
        for (int c = 1; c < planes; ++c) {
            const int64_t eP = line_off[plane0[c]], big = std::max<int64_t>(eP, N - eP);
            if (best < 0 || big < best) { best = big; C = c; }
        }
    }
    if (C < 1 || C >= planes || C > 16) return;
    W.C = C;
    const int64_t LP = plane0[C];
    W.e0[0] = 0; W.ne[0] = line_off[LP];
    W.e0[1] = line_off[LP]; W.ne[1] = int32_t(N - line_off[LP]);
    // line (a_1, a_2) -> line index (or -1)
    auto line_of = [&](int a1, int a2) -> int64_t {
        if (a2 >= planes) return -1;
        const int64_t l = plane0[a2] + a1;
        return (l < plane0[a2 + 1]) ? l : -1;
    };
    // columns of Q: (a_0, a_1) with t(a_0, a_1) > C, i.e. a_0 < len(line (a_1, C))
    W.ncolq = 0;
    for (int a1 = 0; a1 < kMaxRegT; ++a1) {
        const int64_t l = line_of(a1, C);
        W.cb[a1] = W.ncolq;
        W.rq[a1] = (a1 < n1 && l >= 0) ? LEN(l) : 0;
        W.ncolq += W.rq[a1];
    }
    const int TG = 384;
    for (int p = 0; p < 2; ++p) {
        const int64_t l_lo = p == 0 ? 0 : LP, l_hi = p == 0 ? LP : N1;
        const int a2_lo = p == 0 ? 0 : C, a2_hi = p == 0 ? C : planes;
        for (int h = 0; h < 3; ++h) {
            std::vector<WholeItem> nat;
            if (h == 0) {
                for (int64_t l = l_lo; l < l_hi; ++l)
                    nat.push_back({LEN(l), uint32_t(line_off[l] - W.e0[p]) | (uint32_t(LEN(l)) << 16)});
            } else if (h == 1) {
                for (int a2 = a2_lo; a2 < a2_hi; ++a2) {
                    while (W.seo.size() % 4) W.seo.push_back(0);
                    const uint32_t s0 = uint32_t(W.seo.size());
                    const int64_t f0 = plane0[a2], nl = plane0[a2 + 1] - f0;
                    for (int64_t r = 0; r < nl; ++r) W.seo.push_back(uint16_t(line_off[f0 + r] - W.e0[p]));
                    for (int lane = 0; lane < LEN(f0); ++lane) {
                        int t = 0;
                        while (t < nl && LEN(f0 + t) > lane) ++t;
                        nat.push_back({t, s0 | (uint32_t(t) << 16) | (uint32_t(lane) << 22) | (uint32_t(a2) << 27)});
                    }
                }
            } else {
                for (int a1 = 0; a1 < n1; ++a1) {
                    if (line_of(a1, 0) < 0) break;
                    while (W.seo.size() % 4) W.seo.push_back(0);
                    const uint32_t s0 = uint32_t(W.seo.size());
                    int nl = 0;  // lines (a_1, a_2) of the whole column group
                    while (line_of(a1, nl) >= 0) ++nl;
                    // part Q: C placeholders first, so that entry k of the list belongs to row k in both parts
                    if (p == 1) for (int q = 0; q < C; ++q) W.seo.push_back(0);
                    for (int a2 = a2_lo; a2 < std::min(a2_hi, nl); ++a2) W.seo.push_back(uint16_t(line_off[line_of(a1, a2)] - W.e0[p]));
                    const int lanes = LEN(line_of(a1, 0));
                    for (int lane = 0; lane < lanes; ++lane) {
                        int t = 0;
                        while (t < nl && LEN(line_of(a1, t)) > lane) ++t;
                        if (p == 0) {
                            const int tp = std::min(t, C);
                            nat.push_back({tp, s0 | (uint32_t(tp) << 16) | (uint32_t(lane) << 22) | (uint32_t(a1) << 27)});
                        } else if (t > C) {
                            nat.push_back({t, s0 | (uint32_t(t) << 16) | (uint32_t(lane) << 22) | (uint32_t(a1) << 27)});
                        }
                    }
                }
            }
            for (const WholeItem &it : nat) W.max_t = std::max(W.max_t, it.t);
            W.item_off[p][h] = int32_t(W.items.size());
            W.n_items[p][h] = layout_whole_items(nat, (order >> (2 * h)) & 3, (order & 64) != 0, TG, 1, W.items);
        }
    }
    for (int q = 0; q < kMaxRegT + 8 || W.seo.size() % 8; ++q) W.seo.push_back(0);
    if (W.seo.size() > 65536) return;
    W.ok = true;
}

}  // namespace lpfnt

// ------------------------------------------------------------------------------------------
// C ABI: host-only entry points
// ------------------------------------------------------------------------------------------
using namespace lpfnt;


struct lpfnt_index {
    lpfnt::HostIndex hx;
};

namespace {
// returns pointer/bytes/count of a table, or false
bool table_view(const lpfnt_index *ix, int what, int h, const void **ptr, size_t *bytes, int64_t *count) {
    const HostIndex &x = ix->hx;
    auto set = [&](const void *p, size_t elem, size_t n, int64_t cnt) { *ptr = p; *bytes = elem * n; *count = cnt; return true; };
    if (what == LPFNT_TAB_LINE_OFF) return set(x.line_off.data(), 4, x.line_off.size(), int64_t(x.line_off.size()));
    if (what == LPFNT_TAB_LINE_ALPHA) return set(x.line_alpha.data(), 1, x.line_alpha.size(), int64_t(x.line_alpha.size()));
    if (what == LPFNT_TAB_LEVEL_SIZE) return set(x.level_size.data(), 8, x.level_size.size(), int64_t(x.level_size.size()));
    if (what == LPFNT_TAB_LINE_ORDER) return set(x.line_order.data(), 4, x.line_order.size(), int64_t(x.line_order.size()));
    if (what == LPFNT_TAB_WHOLE_ITEMS) return set(x.whole.items.data(), 4, x.whole.items.size(), int64_t(x.whole.items.size()));
    if (what == LPFNT_TAB_WHOLE_SEO) return set(x.whole.seo.data(), 2, x.whole.seo.size(), int64_t(x.whole.seo.size()));
    if (what == LPFNT_TAB_WSPLIT_ITEMS) return set(x.wsplit.items.data(), 4, x.wsplit.items.size(), int64_t(x.wsplit.items.size()));
    if (what == LPFNT_TAB_WSPLIT_SEO) return set(x.wsplit.seo.data(), 2, x.wsplit.seo.size(), int64_t(x.wsplit.seo.size()));
    if (what == LPFNT_TAB_LINE_LAST) return set(x.line_last.data(), 1, x.line_last.size(), int64_t(x.line_last.size()));
    if (h < 1 || h >= x.m) return false;
    const DimTable &D = x.dims[h];
    if (what == LPFNT_TAB_FIB_PTR) return set(D.fib_ptr.data(), 4, D.fib_ptr.size(), int64_t(D.fib_ptr.size()));
    if (what == LPFNT_TAB_FIB_ENT) return set(D.ent.data(), 8, D.ent.size(), int64_t(2 * D.ent.size()));
    if (what == LPFNT_TAB_THR_START) return set(D.thr_start.data(), 4, D.thr_start.size(), int64_t(D.thr_start.size()));
    if (what == LPFNT_TAB_THR_FIB) return set(D.thr_fib.data(), 4, D.thr_fib.size(), int64_t(D.thr_fib.size()));
    if (what == LPFNT_TAB_GROUPS) return set(D.groups.data(), 8, D.groups.size(), int64_t(2 * D.groups.size()));
    if (what == LPFNT_TAB_GROUP_T) return set(D.group_t.data(), 1, D.group_t.size(), int64_t(D.group_t.size()));
    return false;
}
}  // namespace

extern "C" {

const char *lpfnt_last_error(void) { return g_err.c_str(); }

int64_t lpfnt_lp_set_count(int m, int n, double p) {
    if (m < 1 || n < 0 || !((p > 0.0 && p <= 2.0) || std::isinf(p))) {
        set_error("lp_set: need m >= 1, n >= 0, p in (0,2] or inf");
        return LPFNT_EINVAL;
    }
    int64_t N = enumerate(m, n, p, nullptr, 0);
    if (N < 0) { set_error("lp_set: size overflow"); return LPFNT_EINVAL; }
    return N;
}

int lpfnt_lp_set_fill(int m, int n, double p, int64_t *A, int64_t N) {
    if (!A || m < 1 || n < 0) { set_error("lp_set_fill: bad arguments"); return LPFNT_EINVAL; }
    int64_t got = enumerate(m, n, p, A, N);
    if (got != N) { set_error("lp_set_fill: N does not match the set size"); return LPFNT_EINVAL; }
    return LPFNT_OK;
}

int64_t lpfnt_lp_tube(const int64_t *A, int64_t N, int m, int64_t *T) {
    if (!A || N < 1 || m < 1) { set_error("lp_tube: bad arguments"); return LPFNT_EINVAL; }
    int64_t lines = 0, run = 0;
    for (int64_t e = 0; e < N; ++e) {
        if (A[e * m] == 0 && e > 0) {
            if (T) T[lines] = run;
            ++lines;
            run = 0;
        }
        ++run;
    }
    if (T) T[lines] = run;
    return lines + 1;
}

int lpfnt_entropy(const int64_t *A, int64_t N, int m, int64_t *e_T) {
    if (!A || !e_T || N < 1 || m < 1) { set_error("entropy: bad arguments"); return LPFNT_EINVAL; }
    // level k size = number of members whose first k coordinates vanish
    std::vector<int64_t> cnt(m + 1, 0);
    for (int64_t e = 0; e < N; ++e) {
        int z = 0;
        while (z < m && A[e * m + z] == 0) ++z;
        for (int k = 0; k <= z; ++k) ++cnt[k];
    }
    if (N == 1) { e_T[0] = 1; return 1; }  // set.py:171-172
    for (int k = 0; k <= m; ++k) e_T[k] = cnt[k];
    return m + 1;
}

int lpfnt_ordinal_embedding(const int64_t *As, int64_t Ns, const int64_t *Ab, int64_t Nb, int m, int64_t *phi) {
    if (!As || !Ab || !phi || m < 1) { set_error("ordinal_embedding: bad arguments"); return LPFNT_EINVAL; }
    auto cmp = [m](const int64_t *a, const int64_t *b) {
        for (int j = m - 1; j >= 0; --j) {
            if (a[j] < b[j]) return -1;
            if (a[j] > b[j]) return 1;
        }
        return 0;
    };
    int64_t j = 0;
    for (int64_t i = 0; i < Ns; ++i) {
        while (j < Nb && cmp(Ab + j * m, As + i * m) < 0) ++j;
        if (j >= Nb || cmp(Ab + j * m, As + i * m) != 0) {
            set_error("ordinal_embedding: the small set is not contained in the large one");
            return LPFNT_EINVAL;
        }
        phi[i] = j;
    }
    return LPFNT_OK;
}

int lpfnt_index_create(const int64_t *A, int64_t N, int m, int n, lpfnt_index **index) {
    if (!index) { set_error("index pointer is null"); return LPFNT_EINVAL; }
    *index = nullptr;
    if (!A) { set_error("index_create: A is null"); return LPFNT_EINVAL; }
    lpfnt_index *ix = new lpfnt_index();
    if (!ix->hx.build(A, N, m, n)) { delete ix; return LPFNT_EINVAL; }
    *index = ix;
    return LPFNT_OK;
}
int lpfnt_index_destroy(lpfnt_index *index) { delete index; return LPFNT_OK; }
int lpfnt_index_whole_build(lpfnt_index *index, int threads, int order, int fibres) {
    if (!index) { set_error("index is null"); return LPFNT_EINVAL; }
    if (threads != 384 && threads != 768) { set_error("whole_build: threads must be 384 or 768"); return LPFNT_EINVAL; }
    if (fibres < 1 || fibres > 3) { set_error("whole_build: fibres must be 1, 2 (two per thread, side by side) or 3 (two per thread, one after the other)"); return LPFNT_EINVAL; }
    index->hx.build_whole(threads, order < 0 ? kWholeDefaultOrder : order, fibres);
    return LPFNT_OK;
}
int lpfnt_index_wsplit_build(lpfnt_index *index, int C, int order) {
    if (!index) { set_error("index is null"); return LPFNT_EINVAL; }
    index->hx.build_whole_split(C, order < 0 ? kWholeDefaultOrder : order);
    return LPFNT_OK;
}
int lpfnt_index_wsplit_meta(const lpfnt_index *index, int32_t *meta) {
    if (!index || !meta) { set_error("wsplit_meta: bad arguments"); return LPFNT_EINVAL; }
    const WholeSplitHost &W = index->hx.wsplit;
    int q = 0;
    meta[q++] = W.ok ? 1 : 0; meta[q++] = W.C;
    for (int p = 0; p < 2; ++p) { meta[q++] = W.e0[p]; meta[q++] = W.ne[p]; }
    for (int p = 0; p < 2; ++p) for (int h = 0; h < 3; ++h) { meta[q++] = W.item_off[p][h]; meta[q++] = W.n_items[p][h]; }
    meta[q++] = W.ncolq;
    for (int a = 0; a < kMaxRegT; ++a) meta[q++] = W.cb[a];
    for (int a = 0; a < kMaxRegT; ++a) meta[q++] = W.rq[a];
    return q;  // 83 entries
}
int lpfnt_index_whole_meta(const lpfnt_index *index, int32_t *meta) {
    if (!index || !meta) { set_error("whole_meta: bad arguments"); return LPFNT_EINVAL; }
    const WholeHost &W = index->hx.whole;
    int q = 0;
    meta[q++] = W.ok ? 1 : 0; meta[q++] = W.threads; meta[q++] = W.max_t; meta[q++] = W.max_items;
    for (int h = 0; h < 3; ++h) { meta[q++] = W.item_off[h]; meta[q++] = W.n_items[h]; }
    return q;  // 10 entries
}
int64_t lpfnt_index_size(const lpfnt_index *index, int what, int h) {
    const void *p; size_t b; int64_t c;
    if (!index || !table_view(index, what, h, &p, &b, &c)) { set_error("index_size: no such table"); return LPFNT_EINVAL; }
    return c;
}
int lpfnt_index_copy(const lpfnt_index *index, int what, int h, void *dst) {
    const void *p; size_t b; int64_t c;
    if (!index || !dst || !table_view(index, what, h, &p, &b, &c)) { set_error("index_copy: no such table"); return LPFNT_EINVAL; }
    std::memcpy(dst, p, b);
    return LPFNT_OK;
}

}  // extern "C"
