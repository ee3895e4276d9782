// Thread bodies of the MSM pipeline.  Every body is a pure function of its thread index (no shared memory, no
// warp intrinsics), so the same source runs as a CUDA thread (msm.cu wraps each body in a __global__ kernel) and,
// for logic tests without a GPU, as a loop iteration in tests/emul/ (host build; never shipped in the library).
//
// Pipeline for a group of g scalar vectors of n points against one base range (what one `batch_msm` call of the
// reference, co-jolt/src/poly/commitment/pst13.rs:319-323, asks for):
//   1 decompose     scalar -> W signed c-bit digits -> (key, val) pairs.  key = (vector*W + window)*B + |digit|-1,
//                   B = 2^(c-1); val = point index | sign << 31.  Zero digits keep a valid key and a skip mark in val.
//   2 sort          pairs by key (sort_kernels.cuh: LSD radix sort, the first pass fused with the decompose step) ->
//                   every bucket is one contiguous run.
//   3 accumulate    load-balanced segmented sum: thread t owns pairs [t*L, (t+1)*L), whatever buckets they belong to.
//                   Runs that lie inside one chunk are written straight to their bucket; the (at most two) runs that
//                   cross a chunk edge are written as partial sums and summed by the same body one level up, until
//                   one thread sees a whole level.  Work per thread is constant for ANY scalar distribution - the
//                   co-jolt party shares are constant vectors (one bucket per window holds all n points).
//   4 bucket reduce sum_b (b+1) * bucket[b] per window: group running sums over l buckets, then J + 2 plain "bit sums"
//                   per window (no serial pass over the window, no doublings).
//   5 finish        one Horner pass over the 254 bit positions, one inversion, affine wire point; on the host for
//                   a few vectors, a GPU thread per vector for large batches.
#pragma once
#include <cstddef>

#include "curve.cuh"

namespace cozk {

constexpr uint32_t KEY_SENTINEL = 0xFFFFFFFFu;  // zero digit / unused partial slot
constexpr uint32_t KEY_FILL = 0x80000000u;      // partial slot that only keeps a run contiguous (identity point)
constexpr uint32_t KEY_MASK = 0x7FFFFFFFu;
constexpr uint32_t VAL_NEG = 0x80000000u;
constexpr uint32_t VAL_SKIP = 0xFFFFFFFFu;      // zero digit: the pair stays in the list (bucket 0 of its window) but adds nothing

constexpr int SCALAR_MONT = 0;   // Fr Montgomery (what msm_field_elements gets)
constexpr int SCALAR_CANON = 1;  // BigInt<4> (what msm_bigint gets)

COZK_HD fq load_fq(const void* p) {
    fq r;
#if defined(__CUDA_ARCH__)
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
#else
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
    for (int i = 0; i < 8; ++i) r.v[i] = q[i];
#endif
    return r;
}
COZK_HD void store_fq(void* p, const fq& a) {
#if defined(__CUDA_ARCH__)
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
#else
    uint32_t* q = reinterpret_cast<uint32_t*>(p);
    for (int i = 0; i < 8; ++i) q[i] = a.v[i];
#endif
}
COZK_HD affine load_affine(const affine* p) {
    affine r;
    r.x = load_fq(&p->x);
    r.y = load_fq(&p->y);
    return r;
}
COZK_HD xyzz load_xyzz(const xyzz* p) {
    xyzz r;
    r.X = load_fq(&p->X);
    r.Y = load_fq(&p->Y);
    r.ZZ = load_fq(&p->ZZ);
    r.ZZZ = load_fq(&p->ZZZ);
    return r;
}
COZK_HD void store_xyzz(xyzz* p, const xyzz& a) {
    store_fq(&p->X, a.X);
    store_fq(&p->Y, a.Y);
    store_fq(&p->ZZ, a.ZZ);
    store_fq(&p->ZZZ, a.ZZZ);
}

// ------------------------------------------------------------------------------------------------ 1 decompose
struct DecomposeArgs {
    const uint8_t* scalars;    // device copy; vector v starts at scalars + v*vector_stride, element i at + i*stride
    const uint8_t* const* vec_ptrs;  // if non-null: vector v starts at vec_ptrs[v] instead (device-resident shares)
    size_t vector_stride;      // bytes between vectors
    size_t stride;             // bytes between elements (32 dense, 64 Rep3 AoS share a)
    int form;                  // SCALAR_MONT / SCALAR_CANON
    size_t n;                  // points per vector
    uint32_t g;                // vectors in this group
    uint32_t c;                // window bits
    uint32_t W;                // windows
    const uint8_t* infinity;   // optional per-base flag (already offset to this base range), or null
    uint32_t* keys;            // [g*W*n]
    uint32_t* vals;            // [g*W*n]
    // Precomputed-table mode (table_stride != 0): the SRS holds 2^(c*w) * P_i at table[w*table_stride + i], so the
    // digit of window w selects a point of table row w and ALL windows share one bucket set per vector.
    uint32_t bucket_windows;   // W normally, 1 in table mode
    size_t table_stride;       // points per table row (the registered SRS length), 0 = no table
    size_t val_offset;         // base_offset of this call inside a table row (0 without table: bases are pre-offset)
    // Dominant-digit mode (dom_mode != null; the call covers the WHOLE registered SRS, which carries the sum of every table
    // row).  Real co-jolt shares are constant or nearly constant vectors: in most windows nearly every
    // scalar has the SAME digit c.  With S = sum_i T_w[i] precomputed,
    //     sum_i d_i T_w[i] = sum_{i: d_i != c} d_i T_w[i]  +  c * (S - sum_{i: d_i != c} T_w[i]),
    // so a window whose dominant digit covers a fraction f of the scalars needs 2 (1 - f) n + 1 pairs instead of n: every
    // other scalar contributes its own pair plus one pair that takes its point OUT of bucket |c|, and one pair puts S in.
    // Segment (v, w) of the pair list starts at seg_off[v*W + w]; its mode is 0 (as above: pair i at offset i), 1 (dominant
    // digit dom_cand[v*W + w] != 0) or 2 (zero digits dropped, nothing to compensate); slots of modes 1 / 2 come from the
    // cursor seg_cursor[v*W + w].
    const uint32_t* dom_mode;
    const int32_t* dom_cand;
    const uint64_t* seg_off;
    uint32_t* seg_cursor;
    const uint64_t* seg_len;   // pairs of the segment (mode 1: its last pair is the total S)
    size_t totals_index;       // index (into the base table) of the total of table row 0; row w: + w (table mode)
    // Ragged group (rag_start != null; fused sort path only): vector v has rag_start[v+1] - rag_start[v] scalars and pairs
    // with bases rag_base[v] .. of the SRS (the levels of a PST13 opening against one concatenated SRS: one launch for
    // vectors of 2^21, 2^20, .. points).  `n` is then unused; val_offset must be 0.
    const uint32_t* rag_start; // [g + 1]
    const uint32_t* rag_base;  // [g]
    size_t total;              // scalars in the group; 0 = g * n
};
COZK_HD size_t decompose_total(const DecomposeArgs& A) { return A.total ? A.total : (size_t)A.g * A.n; }
// index of the first base of vector v inside the SRS row (and inside the infinity flags)
COZK_HD size_t decompose_base(const DecomposeArgs& A, uint32_t v) { return A.rag_base ? (size_t)A.rag_base[v] : A.val_offset; }

// bits [off, off+c) of a 256-bit little-endian integer, c <= 24
COZK_HD uint32_t extract_bits(const fr& s, uint32_t off, uint32_t c) {
    uint32_t limb = off >> 5, sh = off & 31;
    uint64_t lo = s.v[limb];
    uint64_t hi = (limb + 1 < 8) ? s.v[limb + 1] : 0;
    uint64_t x = (lo | (hi << 32)) >> sh;
    return (uint32_t)x & ((1u << c) - 1u);
}

// window w of a canonical scalar as a signed digit: magnitude in [0, B], sign, carry into the next window
COZK_HD uint32_t signed_digit(const fr& s, uint32_t w, uint32_t c, uint32_t& carry, uint32_t& neg) {
    const uint32_t B = 1u << (c - 1);
    uint32_t d = extract_bits(s, w * c, c) + carry;
    neg = 0;
    carry = 0;
    if (d > B) {  // digit in [-B+1, B]; the top window never carries because W*c >= bits + 1
        d = (1u << c) - d;
        neg = VAL_NEG;
        carry = 1;
    }
    return d;
}
COZK_HD fr decompose_load(size_t tid, const DecomposeArgs& A, uint32_t& v, size_t& i) {
    if (A.rag_start) {  // the vector whose range holds tid: largest v with rag_start[v] <= tid
        uint32_t lo = 0, hi = A.g;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (A.rag_start[mid] <= tid) lo = mid;
            else hi = mid;
        }
        v = lo;
        i = tid - A.rag_start[v];
    } else {
        v = (uint32_t)(tid / A.n);
        i = tid - (size_t)v * A.n;
    }
    const uint8_t* vec = A.vec_ptrs ? A.vec_ptrs[v] : A.scalars + (size_t)v * A.vector_stride;
    fr s = load_fq(vec + i * A.stride);
    return A.form == SCALAR_MONT ? fr_from_mont(s) : fr_reduce_canon(s);
}
// one slot counter per segment; on the device the lanes of a warp that ask for the same counter share one atomic
COZK_HD uint32_t seg_take(uint32_t* cursor, uint32_t seg, uint32_t n) {
#if defined(__CUDA_ARCH__)
    const unsigned peers = __match_any_sync(__activemask(), seg);
    const unsigned lane = threadIdx.x & 31u;
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == (unsigned)leader) base = atomicAdd(&cursor[seg], n * (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    return base + n * (uint32_t)__popc(peers & ((1u << lane) - 1u));
#else
    uint32_t base = cursor[seg];
    cursor[seg] += n;
    return base;
#endif
}

// The (key, val) pair of window w of scalar (v, i): digit magnitude d with sign `neg`; `skip` = base at infinity.
// A zero digit keeps a valid key (bucket 0 of its window) and is marked in val instead: the sort then needs no extra key
// bit for a sentinel, which saves a whole radix pass when the bucket index is a multiple of 8 bits.
COZK_HD void make_pair(const DecomposeArgs& A, uint32_t v, size_t i, uint32_t w, uint32_t d, uint32_t neg, bool skip,
                       uint32_t& key, uint32_t& val) {
    const uint32_t B = 1u << (A.c - 1);
    const bool zero = (d == 0) || skip;
    const uint32_t bw = A.table_stride ? 0u : w;
    const uint32_t key0 = (v * A.bucket_windows + bw) * B;
    const uint32_t point = (uint32_t)((size_t)w * A.table_stride + A.val_offset + i);
    key = key0 + (zero ? 0u : d - 1);
    val = zero ? VAL_SKIP : (point | neg);
}

// thread tid = v*n + i
COZK_HD void decompose_body(size_t tid, const DecomposeArgs& A) {
    if (tid >= (size_t)A.g * A.n) return;
    uint32_t v;
    size_t i;
    const fr s = decompose_load(tid, A, v, i);
    bool skip = A.infinity && A.infinity[i];
    const uint32_t B = 1u << (A.c - 1);
    uint32_t carry = 0;
    for (uint32_t w = 0; w < A.W; ++w) {
        uint32_t neg;
        const uint32_t d = signed_digit(s, w, A.c, carry, neg);
        const bool zero = (d == 0) || skip;
        uint32_t key, val;
        make_pair(A, v, i, w, d, neg, skip, key, val);
        const uint32_t bw = A.table_stride ? 0u : w;
        const uint32_t key0 = (v * A.bucket_windows + bw) * B;
        const uint32_t point = (uint32_t)((size_t)w * A.table_stride + A.val_offset + i);
        const uint32_t seg = v * A.W + w;
        const uint32_t mode = A.dom_mode ? A.dom_mode[seg] : 0u;
        if (mode == 0) {
            const size_t o = (A.dom_mode ? (size_t)A.seg_off[seg] : (size_t)seg * A.n) + i;
            A.keys[o] = key;
            A.vals[o] = val;
        } else if (mode == 2) {  // zero digits are the dominant ones: they simply do not appear
            if (zero) continue;
            const size_t o = (size_t)A.seg_off[seg] + seg_take(A.seg_cursor, seg, 1);
            A.keys[o] = key;
            A.vals[o] = val;
        } else {
            const int32_t cand = A.dom_cand[seg];
            const uint32_t cmag = (uint32_t)(cand < 0 ? -cand : cand), cneg = cand < 0 ? VAL_NEG : 0u;
            if (i == 0) {  // the total of the row enters bucket |c| with the sign of c
                const size_t o = (size_t)A.seg_off[seg] + (size_t)A.seg_len[seg] - 1;
                A.keys[o] = key0 + cmag - 1;
                A.vals[o] = (uint32_t)(A.totals_index + (A.table_stride ? w : 0u)) | cneg;
            }
            if (!zero && d == cmag && neg == cneg) continue;  // the dominant digit: covered by the total
            const size_t o = (size_t)A.seg_off[seg] + seg_take(A.seg_cursor, seg, 2);
            A.keys[o] = key;
            A.vals[o] = val;
            A.keys[o + 1] = key0 + cmag - 1;  // and its point leaves bucket |c|: opposite sign of c
            A.vals[o + 1] = point | (cneg ^ VAL_NEG);
        }
    }
}

// Dominant-digit analysis.  Candidate per (vector, window): the digit of the vector's FIRST scalar (for a constant or
// nearly constant share that is the dominant one); then one pass counts, per segment, the scalars with that digit and
// the scalars with a zero digit.  The host turns the counts into modes and segment offsets.
struct DomArgs {
    DecomposeArgs D;       // scalars, form, n, g, c, W as for the decompose pass
    int32_t* cand;         // [g*W] signed digit of element 0
    uint32_t* count_cand;  // [g*W]
    uint32_t* count_zero;  // [g*W]
    size_t count_n;        // scalars per vector the counting pass looks at; 0 = n.  A sample (count_n < n) is taken half
                           // from the head and half from the tail of the vector: polynomials are zero-padded at the end
};
// thread v
COZK_HD void dom_cand_body(size_t v, const DomArgs& A) {
    if (v >= A.D.g) return;
    uint32_t vv;
    size_t i;
    const fr s = decompose_load(v * A.D.n, A.D, vv, i);
    uint32_t carry = 0;
    for (uint32_t w = 0; w < A.D.W; ++w) {
        uint32_t neg;
        const uint32_t d = signed_digit(s, w, A.D.c, carry, neg);
        A.cand[v * A.D.W + w] = neg ? -(int32_t)d : (int32_t)d;
    }
}
// thread tid = v*count_n + i.  SERIAL form (plain increments): the host emulation runs it as it is; on the GPU the
// block-cooperative k_dom_count (msm.cu) produces the same counters with ballots and shared-memory counters.
inline void dom_count_body(size_t tid, const DomArgs& A) {
    const size_t cn = A.count_n ? A.count_n : A.D.n;
    if (tid >= (size_t)A.D.g * cn) return;
    uint32_t v = (uint32_t)(tid / cn);
    size_t i = tid - (size_t)v * cn;
    if (cn < A.D.n && i >= cn / 2) i += A.D.n - cn;
    const fr s = decompose_load((size_t)v * A.D.n + i, A.D, v, i);
    uint32_t carry = 0;
    for (uint32_t w = 0; w < A.D.W; ++w) {
        uint32_t neg;
        const uint32_t d = signed_digit(s, w, A.D.c, carry, neg);
        const uint32_t seg = v * A.D.W + w;
        const int32_t sd = neg ? -(int32_t)d : (int32_t)d;
        if (sd == A.cand[seg]) A.count_cand[seg] += 1;
        if (d == 0) A.count_zero[seg] += 1;
    }
}

// ------------------------------------------------------------------------------------------------ SRS table (setup time)
// table[w*n + i] = 2^(c*w) * P_i for w = 0 .. W-1, affine.  Built once when the SRS is registered (the reference builds
// its SRS once in PST13::setup, outside prove time); thread i walks its point through the windows.
struct TableArgs {
    affine* table;   // [W*n], row 0 already holds the bases
    size_t n;
    uint32_t c, W;
};

COZK_HD void table_body(size_t i, const TableArgs& A) {
    if (i >= A.n) return;
    xyzz p = xyzz_from_affine(load_affine(&A.table[i]));
    for (uint32_t w = 1; w < A.W; ++w) {
        for (uint32_t k = 0; k < A.c; ++k) p = xyzz_dbl(p);
        // back to affine (BN254 G1 has prime order: a finite point never doubles to the identity)
        fq I = fq_inv(fq_mul(p.ZZ, p.ZZZ));
        affine a;
        a.x = fq_mul(p.X, fq_mul(I, p.ZZZ));
        a.y = fq_mul(p.Y, fq_mul(I, p.ZZ));
        store_fq(&A.table[(size_t)w * A.n + i].x, a.x);
        store_fq(&A.table[(size_t)w * A.n + i].y, a.y);
        p = xyzz_from_affine(a);
    }
}

// The same table in two steps, for the GPU (table_body stays the output contract: identical bytes): an inversion is ~390
// multiplications, and table_body spends W - 1 of them per point - more than its (W - 1) * c doublings cost.  Step 1 walks the
// chain in XYZZ and parks the un-normalised multiples in a scratch array; step 2 normalises the W - 1 multiples of a point
// with ONE inversion (Montgomery's trick over ZZ * ZZZ; a finite point of the prime-order group never doubles to the
// identity, and the entries of a point flagged at infinity are never read).  Slab of `slab` points starting at `first`.
constexpr uint32_t TABLE_MAX_ROWS = 128;
struct TableSlabArgs {
    affine* table;   // [W*n], row 0 holds the bases
    xyzz* tmp;       // [(W-1) * slab]: multiple w of slab point j at tmp[(w-1) * slab + j]
    size_t n, first, slab;
    uint32_t c, W;
};
COZK_HD void table_chain_body(size_t j, const TableSlabArgs& A) {
    if (j >= A.slab || A.first + j >= A.n) return;
    xyzz p = xyzz_from_affine(load_affine(&A.table[A.first + j]));
    for (uint32_t w = 1; w < A.W; ++w) {
        for (uint32_t k = 0; k < A.c; ++k) p = xyzz_dbl(p);
        store_xyzz(&A.tmp[(size_t)(w - 1) * A.slab + j], p);
    }
}
COZK_HD void table_norm_body(size_t j, const TableSlabArgs& A) {
    if (j >= A.slab || A.first + j >= A.n) return;
    fq pre[TABLE_MAX_ROWS];  // pre[w] = product of the denominators of rows 1 .. w-1
    fq run = fq_one();
    for (uint32_t w = 1; w < A.W; ++w) {
        const xyzz* q = &A.tmp[(size_t)(w - 1) * A.slab + j];
        pre[w] = run;
        run = fq_mul(run, fq_mul(load_fq(&q->ZZ), load_fq(&q->ZZZ)));
    }
    fq inv = fq_inv(run);
    for (uint32_t w = A.W - 1; w >= 1; --w) {
        const xyzz p = load_xyzz(&A.tmp[(size_t)(w - 1) * A.slab + j]);
        const fq I = fq_mul(inv, pre[w]);  // 1 / (ZZ * ZZZ) of row w
        inv = fq_mul(inv, fq_mul(p.ZZ, p.ZZZ));
        affine* o = &A.table[(size_t)w * A.n + A.first + j];
        store_fq(&o->x, fq_mul(p.X, fq_mul(I, p.ZZZ)));
        store_fq(&o->y, fq_mul(p.Y, fq_mul(I, p.ZZ)));
    }
}

// ------------------------------------------------------------------------------------------------ 3 accumulate
// One body for all levels.  LEVEL1: entries are (key, val) pairs pointing into the base table (mixed additions);
// otherwise entries are (key, xyzz partial sum) produced by the level below (full additions).
struct AccumulateArgs {
    size_t m;                 // number of entries at this level
    const uint32_t* keys;     // [m] sorted (level 1) / partial keys of the level below
    const uint32_t* vals;     // level 1 only
    const affine* bases;      // level 1 only, already offset to the base range
    const xyzz* pts_in;       // levels >= 2
    xyzz* buckets;            // [g*W*B]
    uint32_t* pkeys;          // [2*T] partial keys out, T = ceil(m/L)
    xyzz* ppts;               // [2*T] partial sums out
    uint32_t L;               // entries per thread (chosen per call at level 1: MsmPlan::acc_tile)
};

COZK_HD void emit_bucket(const AccumulateArgs& A, uint32_t key, const xyzz& sum) { store_xyzz(&A.buckets[key], sum); }

// Streamed MSM (one vector fed in point chunks): every chunk after the first accumulates into a scratch bucket set,
// which is then added to the main one - a dense pass, one thread per bucket.  (Adding inside the accumulate kernel
// at the moment a run closes was measured 2x slower: lanes close their runs at different iterations, so the 14
// multiplications of the addition ran once per lane instead of once per warp.)
struct MergeArgs {
    xyzz* buckets;
    const xyzz* scratch;
    size_t total;
};
COZK_HD void merge_body(size_t b, const MergeArgs& A) {
    if (b >= A.total) return;
    xyzz s = load_xyzz(&A.scratch[b]);
    if (xyzz_is_identity(s)) return;
    store_xyzz(&A.buckets[b], xyzz_add(load_xyzz(&A.buckets[b]), s));
}

template <bool LEVEL1>
COZK_HD void accumulate_body(size_t t, const AccumulateArgs& A) {
    size_t start = t * (size_t)A.L;
    if (start >= A.m) return;
    size_t end = start + A.L < A.m ? start + A.L : A.m;
    uint32_t cur = A.keys[start] & KEY_MASK;  // sentinel -> KEY_MASK, never a real key
    bool left_open = start > 0 && cur != KEY_MASK && (A.keys[start - 1] & KEY_MASK) == cur;
    bool first_run = true;
    uint32_t key0 = KEY_SENTINEL, key1 = KEY_SENTINEL;
    xyzz acc = xyzz_identity();
    for (size_t i = start; i < end; ++i) {
        uint32_t raw = A.keys[i];
        uint32_t k = raw & KEY_MASK;
        if (k != cur) {
            // the run `cur` ends inside this chunk: it is closed on the right
            if (cur != KEY_MASK) {
                if (first_run && left_open) {
                    key0 = cur;
                    store_xyzz(&A.ppts[2 * t], acc);
                } else {
                    emit_bucket(A, cur, acc);
                }
            }
            acc = xyzz_identity();
            cur = k;
            first_run = false;
        }
        if (k == KEY_MASK) continue;
        if (LEVEL1) {
            uint32_t val = A.vals[i];
            if (val == VAL_SKIP) continue;
            affine p = load_affine(&A.bases[val & ~VAL_NEG]);
            p.y = fq_cneg(p.y, (val & VAL_NEG) != 0);
            acc = xyzz_madd(acc, p);
        } else {
            if (raw & KEY_FILL) continue;
            acc = xyzz_add(acc, load_xyzz(&A.pts_in[i]));
        }
    }
    bool right_open = end < A.m && cur != KEY_MASK && (A.keys[end] & KEY_MASK) == cur;
    if (cur != KEY_MASK) {
        bool lo = first_run && left_open;
        if (!lo && !right_open) {
            emit_bucket(A, cur, acc);
        } else if (first_run) {
            // the whole chunk is one run, open on at least one side: sum in slot 0, a filler keeps the run contiguous
            key0 = cur;
            store_xyzz(&A.ppts[2 * t], acc);
            key1 = cur | KEY_FILL;
        } else {
            key1 = cur;
            store_xyzz(&A.ppts[2 * t + 1], acc);
        }
    }
    A.pkeys[2 * t] = key0;
    A.pkeys[2 * t + 1] = key1;
}

// ------------------------------------------------------------------------------------------------ 4 bucket reduce
// Window value  V = sum_b (b+1) * bucket[b],  b in [0, B).  Two steps, both shallow (a single GPU thread needs about
// 7 us per group addition, so depth - not work - is what this stage costs):
//   4a group   threads own l consecutive buckets: S_g = sum, W_g = sum_s s * bucket[g*l + s]      (2l additions deep)
//   4b bit sums  with b = g*l + s:  sum_b b * bucket[b] = sum_g W_g + l * sum_g g * S_g  and
//              sum_g g * S_g = sum_j 2^j * T_j,  T_j = sum of the S_g whose index g has bit j set.
//              So per window NS = J + 2 PLAIN sums (J = log2(B/l)): id 0 = sum W_g, id 1 = sum S_g, id 2+j = T_j,
//              each a radix-f tree of additions; no doublings, no serial running sum across the window.
//   V = sums[0] + sums[1] + l * sum_j 2^j * sums[2+j]  is folded into the bit-position Horner of stage 5.
struct GroupArgs {
    const xyzz* buckets;  // [windows * B]
    xyzz* s_out;          // [windows * B / l]
    xyzz* w_out;
    uint32_t l;           // buckets per thread (power of two dividing B)
    size_t threads;       // windows * B / l
};

COZK_HD void group_body(size_t tid, const GroupArgs& A) {
    if (tid >= A.threads) return;
    size_t base = tid * A.l;  // windows are contiguous and B % l == 0, so groups never straddle windows
    xyzz run = xyzz_identity(), acc = xyzz_identity();
    for (uint32_t s = A.l - 1; s >= 1; --s) {
        run = xyzz_add(run, load_xyzz(&A.buckets[base + s]));
        acc = xyzz_add(acc, run);  // ends as sum_s s * bucket[base + s]
    }
    run = xyzz_add(run, load_xyzz(&A.buckets[base]));
    store_xyzz(&A.s_out[tid], run);
    store_xyzz(&A.w_out[tid], acc);
}

struct BitsumArgs {
    const xyzz* s;     // [windows * G]
    const xyzz* w;     // [windows * G]
    xyzz* out;         // [windows * NS * chunks]
    uint32_t G;        // groups per window
    uint32_t NS;       // sums per window = log2(G) + 2
    uint32_t f;        // groups per thread (power of two dividing G)
    uint32_t chunks;   // G / f
    size_t threads;    // windows * NS * chunks
};

// thread tid = (window * NS + id) * chunks + q  sums the f groups of chunk q that belong to sum `id`
COZK_HD void bitsum_body(size_t tid, const BitsumArgs& A) {
    if (tid >= A.threads) return;
    uint32_t q = (uint32_t)(tid % A.chunks);
    size_t rest = tid / A.chunks;
    uint32_t id = (uint32_t)(rest % A.NS);
    size_t win = rest / A.NS;
    const xyzz* src = (id == 0 ? A.w : A.s) + win * A.G;
    xyzz acc = xyzz_identity();
    for (uint32_t e = 0; e < A.f; ++e) {
        uint32_t g = q * A.f + e;
        if (id >= 2 && !((g >> (id - 2)) & 1u)) continue;
        acc = xyzz_add(acc, load_xyzz(&src[g]));
    }
    store_xyzz(&A.out[tid], acc);
}

struct PlainSumArgs {
    const xyzz* in;   // contiguous arrays whose length is a multiple of f
    xyzz* out;
    uint32_t f;
    size_t threads;
};

COZK_HD void plainsum_body(size_t tid, const PlainSumArgs& A) {
    if (tid >= A.threads) return;
    xyzz acc = load_xyzz(&A.in[tid * A.f]);
    for (uint32_t e = 1; e < A.f; ++e) acc = xyzz_add(acc, load_xyzz(&A.in[tid * A.f + e]));
    store_xyzz(&A.out[tid], acc);
}

// Row / column form (MsmPlan::reduce_2d).  Serial output contracts of the two block-cooperative kernels k_rowcol and
// k_masksum (depth_kernels.cu); the GPU adds in tree order, so the XYZZ representation of a sum may differ, the group
// element does not.
struct RowColArgs {
    const xyzz* buckets;  // [windows << (lo_bits + hi_bits)]
    xyzz* rc;             // [windows * (rows + cols)]: per window the row sums R_hi, then the column sums C_lo
    uint32_t lo_bits, hi_bits;
    size_t blocks;        // windows * (rows + cols)
};
// "block" blk = win * (rows + cols) + i:  i < rows: sum of row i;  else: sum of column i - rows
COZK_HD void rowcol_body(size_t blk, const RowColArgs& A) {
    if (blk >= A.blocks) return;
    const size_t rows = (size_t)1 << A.hi_bits, cols = (size_t)1 << A.lo_bits;
    const size_t win = blk / (rows + cols), i = blk % (rows + cols);
    const xyzz* b = A.buckets + win * rows * cols;
    xyzz acc = xyzz_identity();
    if (i < rows) {
        for (size_t e = 0; e < cols; ++e) acc = xyzz_add(acc, load_xyzz(&b[i * cols + e]));
    } else {
        for (size_t e = 0; e < rows; ++e) acc = xyzz_add(acc, load_xyzz(&b[e * cols + (i - rows)]));
    }
    store_xyzz(&A.rc[blk], acc);
}
struct MaskSumArgs {
    const xyzz* rc;   // as written by the row / column pass
    xyzz* out;        // [windows * NS]
    uint32_t lo_bits, hi_bits, NS;  // NS = lo_bits + hi_bits + 2
    size_t blocks;    // windows * NS
};
// "block" blk = win * NS + id:  id 0: identity;  id 1: sum of all buckets (= of the column sums);  id 2 + j: the column
// sums whose index has bit j set (j < lo_bits), the row sums whose index has bit j - lo_bits set (otherwise)
COZK_HD void masksum_body(size_t blk, const MaskSumArgs& A) {
    if (blk >= A.blocks) return;
    const size_t rows = (size_t)1 << A.hi_bits, cols = (size_t)1 << A.lo_bits;
    const size_t win = blk / A.NS;
    const uint32_t id = (uint32_t)(blk % A.NS);
    const xyzz* r = A.rc + win * (rows + cols);
    const xyzz* c = r + rows;
    xyzz acc = xyzz_identity();
    if (id == 1) {
        for (size_t e = 0; e < cols; ++e) acc = xyzz_add(acc, load_xyzz(&c[e]));
    } else if (id >= 2 && id - 2 < A.lo_bits) {
        for (size_t e = 0; e < cols; ++e)
            if ((e >> (id - 2)) & 1u) acc = xyzz_add(acc, load_xyzz(&c[e]));
    } else if (id >= 2) {
        for (size_t e = 0; e < rows; ++e)
            if ((e >> (id - 2 - A.lo_bits)) & 1u) acc = xyzz_add(acc, load_xyzz(&r[e]));
    }
    store_xyzz(&A.out[blk], acc);
}

// ------------------------------------------------------------------------------------------------ 5 finish
// result = sum_w 2^(c*w) * V_w = sum over bit positions: position c*w carries sums[w][0] + sums[w][1], position
// c*w + a + j (a = log2 l, j < J, a + J = c - 1) carries sums[w][2+j].  One Horner pass from the top bit down,
// then one inversion.  The same body runs as a GPU thread per vector (large batches) or on the host (few vectors:
// a CPU core runs this serial chain ~15x faster than a GPU thread).
struct FinishArgs {
    const xyzz* sums;  // [g * W * NS]
    uint32_t g, W, c, NS, log_l;
    uint8_t* out;      // [g] 72-byte wire points
    xyzz* out_xyzz;    // if non-null: the un-normalised sums go here instead (the inversion is 150 us deep on a GPU thread; the
                       // host normalises a whole batch with ONE inversion in a few microseconds)
};

COZK_HD void finish_body(size_t v, const FinishArgs& A) {
    if (v >= A.g) return;
    xyzz acc = xyzz_identity();
    const uint32_t J = A.NS - 2;
    for (uint32_t w = A.W; w-- > 0;) {
        const xyzz* sw = A.sums + ((size_t)v * A.W + w) * A.NS;
        for (uint32_t pos = A.c; pos-- > 0;) {
            if (!xyzz_is_identity(acc)) acc = xyzz_dbl(acc);
            if (pos >= A.log_l && pos - A.log_l < J) acc = xyzz_add(acc, load_xyzz(&sw[2 + pos - A.log_l]));
            if (pos == 0) acc = xyzz_add(acc, xyzz_add(load_xyzz(&sw[0]), load_xyzz(&sw[1])));
        }
    }
    if (A.out_xyzz) store_xyzz(&A.out_xyzz[v], acc);
    else xyzz_to_wire(acc, A.out + 72 * v);
}

}  // namespace cozk
