// Host-side plan of one MSM group: window size, window count, chunk length and the level structure of the two
// multi-level stages.  Pure C++ (no CUDA), shared by the engine (msm.cu) and the logic tests (tests/emul/).
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace cozk {

constexpr int ACC_L = 32;          // entries per thread of the serial accumulate body (level 1: the default, see choose_acc_l)
constexpr int ACC_L_MIN = 16;
constexpr int ACC_L_BIG = 64;      // level-1 chunk once chunks of 64 still give every resident thread 8 of them: half the partial slots for
                                   // the levels above (2^22: accumulate 8.71 -> 8.66 ms, 2^24: 31.66 -> 31.38 ms; 2^20 stays at 32: 2.60 against 2.62 ms)
constexpr int ACC_L_UP = 8;        // partial slots per thread at the serial levels >= 2: 4x the threads of chunks of 32, a quarter of
                                   // the serial chain (2^16 points: 0.68 -> 0.62 ms; 2^20: 3.17 -> 3.15; 2^22: 10.91 -> 10.82)
constexpr int ACC_TILE = 256;      // partial slots per thread BLOCK at small levels >= 2 (block-cooperative segmented scan)
constexpr size_t ACC_SCAN_MAX = 65536;  // levels with more slots than this stay on the serial, work-efficient body
constexpr uint32_t SUM_CHUNK = 1024;  // groups per thread block in the first level of the bucket-reduce sums
constexpr uint32_t GROUP_L = 8;    // buckets per thread in the group step of the bucket reduce
constexpr uint32_t HOST_FINISH_MAX = 4;  // up to this many vectors the final Horner + inversion run on the host
constexpr uint32_t C_MIN = 2, C_MAX = 22;


// Tuning of the level-1 chunk length, carried by the plan: `resident` = threads of the level-1 accumulate kernel that are
// resident on the device at once (SMs x 512; 0 = unknown, keep ACC_L); `force_l` = option "acc_chunk" (a fixed chunk
// length for experiments and tests; 0 = choose).
struct AccTuning {
    size_t resident = 0;
    int force_l = 0;
    int up_l = 0;  // option "acc_chunk_up": partial slots per thread at the serial levels >= 2 (0 = ACC_L_UP)
    int group_l = 0;  // option "group_l": buckets per thread in the group step of the bucket reduce (0 = chosen from the bucket count)
    int reduce_2d = 0;  // option "reduce_2d": 1 = row / column form of the bucket reduce (MsmPlan::reduce_2d)
};

// Level-1 chunk length.  32 pairs per thread is the measured optimum once the kernel fills the device (2^20 points:
// accumulate stage 2.73 / 2.64 / 2.62 / 2.68 / 2.65 ms at 28 / 30 / 32 / 34 / 36 - a chunk of 32 keys is one 128-byte line,
// and the blocks do not run in lock-step waves, so "filling the last wave" buys nothing).  A small MSM that cannot fill
// the device with chunks of 32 gets shorter chunks, down to 16: more threads, a shorter serial chain per thread
// (2^16 points: 0.48 -> 0.40 ms).
inline int choose_acc_l(size_t m, const AccTuning& t) {
    if (t.force_l) return t.force_l;
    const size_t resident = t.resident;
    if (resident == 0 || m == 0) return ACC_L;
    if (m / ACC_L_BIG >= 8 * resident) return ACC_L_BIG;
    if ((m + ACC_L - 1) / ACC_L >= resident) return ACC_L;
    size_t L = (m + resident - 1) / resident;   // the chunk length at which the threads just fill the device
    L = (L + 3) & ~(size_t)3;
    if (L < (size_t)ACC_L_MIN) L = ACC_L_MIN;
    if (L > (size_t)ACC_L) L = ACC_L;
    return (int)L;
}

struct MsmPlan {
    size_t n = 0;        // points per vector
    uint32_t g = 0;      // vectors in the group
    uint32_t bits = 254; // significant scalar bits
    uint32_t c = 0, W = 0, B = 0;
    uint32_t Wb = 0;                 // bucket sets per vector: W, or 1 when a precomputed 2^(c*w)*P table is used
    AccTuning acc;                   // chunk-length tuning the level structure was (and is re-) built with
    size_t m = 0;                    // g*W*n pairs
    size_t total_buckets = 0;        // g*W*B
    uint32_t sort_bits = 0;          // radix-sort key bits
    std::vector<size_t> acc_entries; // entries per accumulate level (level 1 first)
    std::vector<int> acc_tile;       // entries per thread (ACC_L: serial body) or per block (ACC_TILE: segmented scan)
    uint32_t group_l = 1, log_l = 0; // buckets per group, log2
    uint32_t G = 1, NS = 2;          // groups per window, plain sums per window (log2 G + 2)
    uint32_t sum_chunk = 1;          // groups summed per block in the masked level (min(G, SUM_CHUNK))
    uint32_t sum_chunks = 1;         // G / sum_chunk partial sums per (window, id); > 1 needs the second, plain level
    // Row / column form of the bucket reduce: the bucket index b = hi * 2^lo_bits + lo.  With R_hi = sum of row hi and
    // C_lo = sum of column lo,  sum_b b * bucket[b] = sum_lo lo * C_lo + 2^lo_bits * sum_hi hi * R_hi, and each of the two
    // small weighted sums is one masked plain sum per index bit: T_j = sum of the C_lo (R_hi) whose index has bit j set.
    // So NS = (c - 1) + 2 sums per window (id 0 = identity, id 1 = sum of all buckets, id 2 + j = T_j) for 2 additions per
    // bucket, and the depth is two block-cooperative tree sums (over a row / column, then over the row / column sums)
    // instead of a serial group pass plus two.  The finish stage reads it as the group form with l = 1.
    bool reduce_2d = false;
    uint32_t lo_bits = 0, hi_bits = 0;

    // field multiplications this plan performs (for the roofline's "actual" figure): 10 per mixed add, 14 per full add
    double field_mults() const {
        double pairs = (double)m;
        double mul = 10.0 * pairs;
        for (size_t k = 1; k < acc_entries.size(); ++k) mul += 14.0 * 0.5 * (double)acc_entries[k];
        double nb = (double)total_buckets;
        mul += 14.0 * 2.0 * nb;                                   // group step (row / column form: row sums + column sums)
        if (reduce_2d) mul += 14.0 * (double)g * Wb * NS * 0.5 * (double)((1u << lo_bits) + (1u << hi_bits));
        else mul += 14.0 * (double)g * Wb * G * (1.0 + NS / 2.0);      // masked sums + the tree above them
        mul += (double)g * Wb * (9.0 * c + 14.0 * NS);            // bit-position Horner
        return mul;
    }
    // the same figure from the cost model used to choose between plans (cheap, no level vectors needed)
    double model_cost() const {
        return 11.0 * (double)W * (double)n + 45.0 * (double)Wb * (double)B + 400.0 * (double)((c + 2) / 4);
    }
};

inline uint32_t windows_for(uint32_t bits, uint32_t c) { return (bits + 1 + c - 1) / c; }

// A window whose top digit has few bits concentrates all n points of that window in a handful of buckets: long runs
// that every level of the segmented accumulation has to carry (serial latency).  Among plans within 3 % of the best
// cost, take the one whose top window is fullest.
inline uint32_t top_window_bits(uint32_t bits, uint32_t c) { return bits + 1 - (windows_for(bits, c) - 1) * c; }

// cost model in field multiplications: per window n mixed adds (+ the partial-merge overhead) and ~45 per bucket.
// Returns 0 when no window size keeps g * W * 2^(c-1) buckets inside max_buckets (and below 2^31: bucket keys share a
// 32-bit word with the KEY_FILL mark): the caller has to split the group.
inline uint32_t choose_window(size_t n, uint32_t g, uint32_t bits, size_t max_buckets) {
    double cost[C_MAX + 1];
    for (double& x : cost) x = 1e300;
    double best_cost = 1e300;
    for (uint32_t c = C_MIN; c <= C_MAX; ++c) {
        uint32_t W = windows_for(bits, c);
        double B = (double)(1u << (c - 1));
        if ((double)g * W * B > (double)max_buckets || (double)g * W * B >= 2147483647.0) continue;
        cost[c] = (double)W * (11.0 * (double)n + 45.0 * B) + 400.0 * (double)((c + 2) / 4);
        if (cost[c] < best_cost) best_cost = cost[c];
    }
    if (best_cost >= 1e300) return 0;
    uint32_t best = 0, best_top = 0;
    for (uint32_t c = C_MIN; c <= C_MAX; ++c) {
        if (cost[c] >= 1e300 || cost[c] > best_cost * 1.03) continue;
        uint32_t tb = top_window_bits(bits, c);
        if (tb > best_top) {
            best_top = tb;
            best = c;
        }
    }
    return best;
}

// the largest group size g <= want for which some window fits the bucket budget (>= 1: callers pass budgets that hold
// one vector at c = C_MIN)
inline uint32_t max_group_for(size_t n, uint32_t want, uint32_t bits, size_t max_buckets) {
    uint32_t g = want ? want : 1;
    while (g > 1 && choose_window(n, g, bits, max_buckets) == 0) g = (g + 1) / 2;
    return g;
}

// window size for a registered SRS of n points whose windows all share one bucket set (precomputed table)
inline uint32_t choose_table_window(size_t n) {
    double cost[C_MAX + 1];
    double best_cost = 1e300;
    for (uint32_t c = C_MIN; c <= C_MAX; ++c) {
        cost[c] = 11.0 * (double)windows_for(254, c) * (double)n + 45.0 * (double)(1u << (c - 1));
        if (cost[c] < best_cost) best_cost = cost[c];
    }
    uint32_t best = C_MIN, best_top = 0;
    for (uint32_t c = C_MIN; c <= C_MAX; ++c) {
        if (cost[c] > best_cost * 1.03) continue;
        uint32_t tb = top_window_bits(254, c);
        if (tb > best_top) {
            best_top = tb;
            best = c;
        }
    }
    return best;
}

// level structure of the accumulate stage for m (key, val) pairs (m = g*W*n, or fewer in dominant-digit mode)
inline void plan_set_pairs(MsmPlan& p, size_t m) {
    p.m = m;
    p.acc_entries.clear();
    p.acc_tile.clear();
    size_t e = p.m;
    p.acc_entries.push_back(e);
    // Level 1: one thread per l1 pairs (ACC_L; shorter for an MSM too small to fill the device).  Levels >= 2 reduce the partial slots: big levels with the same serial body
    // (one addition per live slot: work-efficient, 16 additions deep), small ones with the block-cooperative segmented
    // scan (log2(ACC_TILE) additions deep, but up to that many additions per slot).  Every thread / block emits two
    // slots; a level that ran as a single thread / block has seen everything and leaves no open run.
    const int l1 = choose_acc_l(e, p.acc);
    p.acc_tile.push_back(l1);
    for (size_t t = (e + l1 - 1) / l1; t > 1;) {
        e = 2 * t;
        int tile = e > ACC_SCAN_MAX ? (p.acc.up_l ? p.acc.up_l : ACC_L_UP) : ACC_TILE;
        p.acc_entries.push_back(e);
        p.acc_tile.push_back(tile);
        t = (e + tile - 1) / tile;
    }
}

// Dominant-digit layout (msm_kernels.cuh, DecomposeArgs): from the per-segment counts of the analysis pass to modes, segment
// offsets and lengths.  A non-zero dominant digit pays off only when nearly every scalar has it (every other scalar costs
// two pairs): 90 %.  Dropping zero digits always pays; it is worth a compacted segment from 25 % on.  Returns the total
// number of pairs, or 0 when no segment is dominant (the caller then keeps the plain layout).
inline size_t dom_layout(const int32_t* cand, const uint32_t* count_cand, const uint32_t* count_zero, size_t segments, size_t n,
                         std::vector<uint32_t>& mode, std::vector<uint64_t>& seg_off, std::vector<uint64_t>& seg_len) {
    mode.assign(segments, 0);
    seg_off.assign(segments, 0);
    seg_len.assign(segments, 0);
    size_t total = 0, special = 0;
    for (size_t sgm = 0; sgm < segments; ++sgm) {
        size_t len = n;
        if (cand[sgm] != 0 && (double)count_cand[sgm] >= 0.9 * (double)n) {
            mode[sgm] = 1;
            len = 2 * (n - count_cand[sgm]) + 1;
        } else if ((double)count_zero[sgm] >= 0.25 * (double)n) {
            mode[sgm] = 2;
            len = n - count_zero[sgm];
        }
        special += mode[sgm] != 0;
        seg_off[sgm] = total;
        seg_len[sgm] = len;
        total += len;
    }
    return special ? total : 0;
}

// table_c != 0: use the SRS's precomputed table (window size fixed at registration, one bucket set per vector)
inline MsmPlan make_plan(size_t n, uint32_t g, uint32_t bits, size_t max_buckets, uint32_t force_c = 0, uint32_t table_c = 0,
                         const AccTuning& acc = AccTuning()) {
    MsmPlan p;
    p.acc = acc;
    p.n = n;
    p.g = g;
    p.bits = bits == 0 || bits > 254 ? 254 : bits;
    p.c = table_c ? table_c : (force_c ? force_c : choose_window(n, g, p.bits, max_buckets));
    if (p.c == 0) p.c = C_MIN;  // no window fits the budget for this many vectors: callers bound g with max_group_for first
    p.W = windows_for(p.bits, p.c);
    p.Wb = table_c ? 1 : p.W;
    p.B = 1u << (p.c - 1);
    p.m = (size_t)g * p.W * n;
    p.total_buckets = (size_t)g * p.Wb * p.B;
    uint32_t sb = 1;
    while (((uint64_t)1 << sb) < (uint64_t)p.total_buckets) ++sb;  // keys are < total_buckets (no sentinel at level 1)
    p.sort_bits = sb;
    plan_set_pairs(p, p.m);
    // Buckets per thread in the group step.  Few buckets: shallow groups (depth is what costs); hundreds of thousands: the
    // stage is throughput bound and the NS masked sums per group dominate, so groups grow.  Measured reduce times (B200):
    // 16 Ki buckets 0.13 / 0.15 / 0.20 ms at l = 2 / 4 / 8; 64 Ki 0.24 / 0.21 / 0.31 ms at 4 / 8 / 16; 512 Ki 1.32 / 0.85 /
    // 0.63 ms at 4 / 8 / 16.
    uint32_t gl = p.total_buckets <= ((size_t)1 << 15) ? 2
                : p.total_buckets <= ((size_t)1 << 17) ? GROUP_L
                : p.total_buckets <= ((size_t)1 << 20) ? 2 * GROUP_L : 4 * GROUP_L;
    if (acc.group_l) gl = (uint32_t)acc.group_l;
    p.group_l = p.B < gl ? p.B : gl;
    p.log_l = 0;
    while ((1u << p.log_l) < p.group_l) ++p.log_l;
    p.G = p.B / p.group_l;
    uint32_t J = 0;
    while ((1u << J) < p.G) ++J;
    p.NS = J + 2;
    p.sum_chunk = p.G < SUM_CHUNK ? p.G : SUM_CHUNK;
    p.sum_chunks = p.G / p.sum_chunk;
    if (acc.reduce_2d) {
        p.reduce_2d = true;
        p.lo_bits = (p.c - 1) / 2;
        p.hi_bits = (p.c - 1) - p.lo_bits;
        p.group_l = 1;
        p.log_l = 0;
        p.G = p.B;
        p.NS = (p.c - 1) + 2;
        p.sum_chunk = 1;
        p.sum_chunks = 1;
    }
    return p;
}

}  // namespace cozk
