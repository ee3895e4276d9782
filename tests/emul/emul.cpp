// Host emulation of the MSM pipeline - TEST HARNESS, never part of the shipped library.
//
// Compiles the very thread bodies the CUDA kernels wrap (co-zkvms_b200/csrc/msm_kernels.cuh) with g++ and runs each
// "launch" as a sequential loop over thread indices, with std::sort standing in for the device radix sort.  It lets
// the no-GPU test tier check the pipeline logic (digit recoding, segmented accumulation with open/closed runs,
// the (S, W) reduce tree, Horner, exceptional group-law cases) against the oracle before any GPU time is spent.
// Field arithmetic here is the portable host body of field.cuh; the PTX carry chains are checked separately by
// tools/gen_field_ptx.py's interpreter and, on the GPU, by tests/test_gpu_field.py.
#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../co-zkvms_b200/csrc/affine_kernels.cuh"
#include "../../co-zkvms_b200/csrc/msm_kernels.cuh"
#include "../../co-zkvms_b200/csrc/msm_plan.hpp"
#include "../../co-zkvms_b200/csrc/rep3_kernels.cuh"
#include "../../experiments/radix29/curve29.cuh"

using namespace cozk;

static AccTuning g_acc;  // emul_set_acc_chunk
static uint32_t g_last_pairs = 0, g_last_dominant = 0;  // of the last chunk of the last emul_msm call (stats[4], stats[5])
static int g_dominant = 0;  // 1: run the dominant-digit path of the engine (whole-SRS calls)
static int g_affine_rounds = 0;  // > 0: the batched-affine pre-reduction in front of the accumulate levels (contract body)
static uint32_t g_last_overflow = 0, g_last_reduced = 0;  // of the last chunk: overflow entries of all rounds, entries of the reduced list

extern "C" {

void emul_set_dominant(int on) { g_dominant = on; }
// SRS table of n points (64 B each), rows x n entries out: mode 0 = table_body (the contract), 1 = chain + batch normalisation
// in slabs of `slab` points
void emul_build_table(const uint8_t* bases, size_t n, uint32_t c, uint32_t rows, int mode, size_t slab, uint8_t* out) {
    affine* t = reinterpret_cast<affine*>(out);
    memcpy(t, bases, n * sizeof(affine));
    if (mode == 0) {
        TableArgs A{t, n, c, rows};
        for (size_t i = 0; i < n; ++i) table_body(i, A);
        return;
    }
    std::vector<xyzz> tmp((size_t)(rows - 1) * slab);
    for (size_t first = 0; first < n; first += slab) {
        TableSlabArgs A{t, tmp.data(), n, first, slab, c, rows};
        for (size_t j = 0; j < slab; ++j) table_chain_body(j, A);
        for (size_t j = 0; j < slab; ++j) table_norm_body(j, A);
    }
}
void emul_set_affine_rounds(int rounds) { g_affine_rounds = rounds; }
void emul_affine_stats(uint32_t* out) {
    out[0] = g_last_reduced;
    out[1] = g_last_overflow;
}
// `rounds` halving rounds through the contract body; outputs sized by the caller ((m + 1) / 2 entries for one round)
void emul_affine_round(const uint32_t* keys, const uint32_t* vals, size_t m, const uint8_t* pts, uint32_t* keys_out, uint32_t* vals_out,
                       uint8_t* pts_out, uint32_t* ovf_count, uint32_t* ovf_keys, uint8_t* ovf_pts, uint32_t ovf_cap) {
    AffineRoundArgs A{m, keys, vals, reinterpret_cast<const affine*>(pts), keys_out, vals_out, reinterpret_cast<affine*>(pts_out),
                      ovf_count, ovf_keys, reinterpret_cast<affine*>(ovf_pts), ovf_cap, nullptr, nullptr};
    *ovf_count = 0;
    for (size_t o = 0; o < affine_round_out(m); ++o) affine_round_body(o, A);
}

// The plan's window choice with the caller's stack poisoned first (choose_window once read cost[] entries it had
// never written).  out[0] = c (0 = no window fits), out[1] = max_group_for(n, g, ...).
__attribute__((noinline)) static void poison_stack(double fill) {
    volatile double junk[64];
    for (int i = 0; i < 64; ++i) junk[i] = fill;
}
void emul_choose_window(size_t n, uint32_t g, uint32_t bits, size_t max_buckets, double poison, uint32_t* out) {
    poison_stack(poison);
    out[0] = choose_window(n, g, bits, max_buckets);
    poison_stack(poison);
    out[1] = max_group_for(n, g, bits, max_buckets);
}

// level-1 chunk length of the accumulate stage: resident != 0 lets the plan choose it as the engine does (wave filling),
// force_l != 0 fixes it
void emul_set_acc_chunk(size_t resident, int force_l) {
    g_acc.resident = resident;
    g_acc.force_l = force_l;
}
// 1: the row / column form of the bucket reduce (MsmPlan::reduce_2d)
void emul_set_reduce_2d(int on) { g_acc.reduce_2d = on; }

// bases: n x 64 B; scalars: g vectors, vector v at scalars + v*vector_stride, element i at + i*stride; out: g x 72 B
int emul_msm(const uint8_t* bases, size_t n, const uint8_t* scalars, size_t vector_stride, size_t stride, int form,
             uint32_t g, uint32_t max_bits, uint32_t force_c, const uint8_t* infinity, uint8_t* out, uint32_t* stats,
             uint32_t table_c, size_t srs_n, size_t base_offset, uint32_t stream_chunks) {
    if (n == 0) {
        for (uint32_t v = 0; v < g; ++v) {
            memset(out + 72 * v, 0, 72);
            out[72 * v + 64] = 1;
        }
        return 0;
    }
    // table mode: `bases` holds the whole registered SRS (srs_n points); the call uses [base_offset, base_offset + n)
    MsmPlan P = make_plan(n, g, max_bits, (size_t)1 << 24, force_c, table_c, g_acc);
    std::vector<affine> table;
    const affine* base_ptr = reinterpret_cast<const affine*>(bases);
    if (table_c) {
        uint32_t rows = windows_for(254, table_c);
        table.resize((size_t)rows * srs_n);
        memcpy(table.data(), bases, srs_n * sizeof(affine));
        TableArgs TA{table.data(), srs_n, table_c, rows};
        for (size_t i = 0; i < srs_n; ++i) table_body(i, TA);
        base_ptr = table.data();
    }
    std::vector<xyzz> buckets(P.total_buckets);
    memset(buckets.data(), 0, buckets.size() * sizeof(xyzz));
    // stream_chunks > 1 (single vector only): the engine's streamed mode - the point range goes through decompose, sort
    // and accumulate in chunks that MERGE into one bucket set
    size_t chunks = (stream_chunks > 1 && g == 1) ? stream_chunks : 1;
    size_t cn_max = (((n + chunks - 1) / chunks) + 31) & ~(size_t)31;
    chunks = (n + cn_max - 1) / cn_max;
    std::vector<uint32_t> pk_in, pk_out;
    std::vector<xyzz> pp_in, pp_out;
    for (size_t ci = 0; ci < chunks; ++ci) {
        size_t clo = ci * cn_max, cn = std::min(cn_max, n - clo);
        MsmPlan Pc = chunks == 1 ? P : make_plan(cn, g, max_bits, (size_t)1 << 24, P.c, table_c, g_acc);
        // dominant-digit mode: the call must cover the whole SRS in one piece; the sums of the table rows sit behind the
        // table (the engine computes them at registration), the analysis pass picks the segments that use them
        std::vector<affine> with_totals;
        std::vector<uint32_t> dmode, cursor;
        std::vector<uint64_t> doff, dlen;
        std::vector<int32_t> cand;
        bool dom = false;
        size_t totals_index = 0;
        // (the engine keeps such sums for the whole SRS and for its exact halvings down to 1024 points; here they are
        // computed for whatever prefix the call covers)
        if (g_dominant && chunks == 1 && !infinity && base_offset == 0) {
            const size_t rows = table_c ? windows_for(254, table_c) : 1;
            with_totals.assign(base_ptr, base_ptr + rows * srs_n);
            bool ok = true;
            for (size_t r = 0; r < rows && ok; ++r) {
                xyzz t = xyzz_identity();
                for (size_t i = 0; i < n; ++i) t = xyzz_madd(t, base_ptr[r * srs_n + i]);
                uint8_t wire[72];
                xyzz_to_wire(t, wire);
                ok = wire[64] == 0;
                affine a;
                memcpy(&a, wire, 64);
                with_totals.push_back(a);
            }
            if (ok) {
                const size_t segs = (size_t)g * Pc.W;
                cand.assign(segs, 0);
                std::vector<uint32_t> cc(segs, 0), cz(segs, 0);
                DomArgs DA{{scalars, nullptr, vector_stride, stride, form, cn, g, Pc.c, Pc.W, nullptr, nullptr, nullptr, Pc.Wb,
                            table_c ? srs_n : 0, 0},
                           cand.data(), cc.data(), cz.data()};
                for (size_t v = 0; v < g; ++v) dom_cand_body(v, DA);
                // as the engine: a sample (head and tail of every vector) first, the full count only if it shows something
                size_t m = 0;
                for (int round = 0; round < 2; ++round) {
                    DA.count_n = round == 0 ? std::min<size_t>(cn, 64) : cn;
                    std::fill(cc.begin(), cc.end(), 0u);
                    std::fill(cz.begin(), cz.end(), 0u);
                    for (size_t t = 0; t < (size_t)g * DA.count_n; ++t) dom_count_body(t, DA);
                    m = dom_layout(cand.data(), cc.data(), cz.data(), segs, DA.count_n, dmode, doff, dlen);
                    if (m == 0) break;
                }
                if (m) {
                    dom = true;
                    plan_set_pairs(Pc, m);
                    cursor.assign(segs, 0);
                    totals_index = rows * srs_n;
                }
            }
        }
        g_last_pairs = (uint32_t)Pc.m;
        g_last_dominant = dom ? 1 : 0;
        std::vector<uint32_t> keys(Pc.m), vals(Pc.m);
        DecomposeArgs D{scalars + clo * stride, nullptr, vector_stride, stride, form, cn, g, Pc.c, Pc.W,
                        infinity ? infinity + clo : nullptr, keys.data(), vals.data(),
                        Pc.Wb, table_c ? srs_n : 0, table_c ? base_offset + clo : 0};
        if (dom) {
            D.dom_mode = dmode.data();
            D.dom_cand = cand.data();
            D.seg_off = doff.data();
            D.seg_cursor = cursor.data();
            D.seg_len = dlen.data();
            D.totals_index = totals_index;
        }
        for (size_t t = 0; t < (size_t)g * cn; ++t) decompose_body(t, D);
        std::vector<size_t> order(Pc.m);
        std::iota(order.begin(), order.end(), (size_t)0);
        uint32_t mask = Pc.sort_bits >= 32 ? 0xFFFFFFFFu : ((1u << Pc.sort_bits) - 1u);
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return (keys[a] & mask) < (keys[b] & mask); });
        std::vector<uint32_t> sk(Pc.m), sv(Pc.m);
        for (size_t i = 0; i < Pc.m; ++i) {
            sk[i] = keys[order[i]];
            sv[i] = vals[order[i]];
        }
        const affine* chunk_bases = dom ? with_totals.data() : (table_c ? base_ptr : base_ptr + clo);
        std::vector<xyzz> scratch;
        if (ci > 0) {
            scratch.resize(P.total_buckets);
            memset(scratch.data(), 0, scratch.size() * sizeof(xyzz));
        }
        // batched-affine pre-reduction (the engine's affine_rounds_for rule), then the accumulate levels over the reduced list
        std::vector<std::vector<uint32_t>> rk, rv, ok;
        std::vector<std::vector<affine>> rp, op;
        std::vector<uint32_t> ocount;
        {
            int R = g_affine_rounds < 0 ? -g_affine_rounds : g_affine_rounds;  // negative: exactly that many rounds, whatever the run lengths
            while (g_affine_rounds > 0 && R > 0 && (Pc.m >> R) < 4 * Pc.total_buckets) --R;
            while (R > 0 && (Pc.m >> (R - 1)) < 2) --R;
            const uint32_t *kin = sk.data(), *vin = sv.data();
            const affine* pin = chunk_bases;
            size_t mr = Pc.m;
            rk.resize(R), rv.resize(R), rp.resize(R), ok.resize(R), op.resize(R), ocount.assign(R, 0);
            g_last_overflow = 0;
            for (int r = 0; r < R; ++r) {
                const size_t mo = affine_round_out(mr), cap = std::min(mo, Pc.total_buckets + 1);
                rk[r].assign(mo, 0xDEADBEEFu), rv[r].assign(mo, 0xDEADBEEFu), rp[r].resize(mo), ok[r].resize(cap), op[r].resize(cap);
                AffineRoundArgs A{mr, kin, vin, pin, rk[r].data(), rv[r].data(), rp[r].data(), &ocount[r], ok[r].data(), op[r].data(),
                                  (uint32_t)cap, nullptr, nullptr};
                for (size_t o = 0; o < mo; ++o) affine_round_body(o, A);
                if (ocount[r] > cap) return 3;
                g_last_overflow += ocount[r];
                kin = rk[r].data(), vin = rv[r].data(), pin = rp[r].data();
                mr = mo;
            }
            if (R > 0) {
                plan_set_pairs(Pc, mr);
                sk.assign(kin, kin + mr);
                sv.assign(vin, vin + mr);
                chunk_bases = pin;
            }
            g_last_reduced = (uint32_t)mr;
        }
        for (size_t lvl = 0; lvl < Pc.acc_entries.size(); ++lvl) {
            size_t m = Pc.acc_entries[lvl];
            const int tile = Pc.acc_tile[lvl];
            size_t T = (m + tile - 1) / tile;
            pk_out.assign(2 * T, 0xDEADBEEFu);
            pp_out.assign(2 * T, xyzz_identity());
            AccumulateArgs A{m, lvl == 0 ? sk.data() : pk_in.data(), sv.data(), chunk_bases,
                             pp_in.data(), ci > 0 ? scratch.data() : buckets.data(), pk_out.data(), pp_out.data(),
                             (uint32_t)tile};
            for (size_t t = 0; t < T; ++t) {
                // levels >= 2 run on the GPU as the block-cooperative k_segscan, whose output contract is this body with
                // one "thread" per tile of ACC_TILE slots
                if (lvl == 0) accumulate_body<true>(t, A);
                else accumulate_body<false>(t, A);
            }
            pk_in.swap(pk_out);
            pp_in.swap(pp_out);
        }
        for (size_t r = 0; r < ocount.size(); ++r) {
            OvfAddArgs OA{&ocount[r], ok[r].data(), op[r].data(), ci > 0 ? scratch.data() : buckets.data(), (uint32_t)ok[r].size()};
            for (size_t i = 0; i < ocount[r]; ++i) ovf_add_body(i, OA);
        }
        if (ci > 0) {
            MergeArgs MA{buckets.data(), scratch.data(), P.total_buckets};
            for (size_t b = 0; b < P.total_buckets; ++b) merge_body(b, MA);
        }
    }
    // the top level must not leave any open run
    if (P.acc_entries.size() > 1 || P.m > 0)
        for (uint32_t k : pk_in) if (k != KEY_SENTINEL) return 2;

    size_t windows = (size_t)g * P.Wb;
    std::vector<xyzz> cur, nxt;
    if (P.reduce_2d) {
        const size_t rc_n = windows * (((size_t)1 << P.hi_bits) + ((size_t)1 << P.lo_bits));
        std::vector<xyzz> rc(rc_n);
        RowColArgs RA{buckets.data(), rc.data(), P.lo_bits, P.hi_bits, rc_n};
        for (size_t b = 0; b < rc_n; ++b) rowcol_body(b, RA);
        cur.resize(windows * P.NS);
        MaskSumArgs MA{rc.data(), cur.data(), P.lo_bits, P.hi_bits, P.NS, windows * P.NS};
        for (size_t b = 0; b < MA.blocks; ++b) masksum_body(b, MA);
    } else {
    std::vector<xyzz> gs(windows * P.G), gw(windows * P.G);
    GroupArgs GA{buckets.data(), gs.data(), gw.data(), P.group_l, windows * P.G};
    for (size_t t = 0; t < GA.threads; ++t) group_body(t, GA);
    // on the GPU these two levels are k_treesum (masked, then plain); same output contract as the bodies below
    uint32_t schunks = P.sum_chunks;
    cur.resize(windows * P.NS * schunks);
    BitsumArgs BA{gs.data(), gw.data(), cur.data(), P.G, P.NS, P.sum_chunk, schunks, windows * P.NS * schunks};
    for (size_t t = 0; t < BA.threads; ++t) bitsum_body(t, BA);
    if (schunks > 1) {
        size_t threads = windows * P.NS;
        nxt.assign(threads, xyzz_identity());
        PlainSumArgs SA{cur.data(), nxt.data(), schunks, threads};
        for (size_t t = 0; t < threads; ++t) plainsum_body(t, SA);
        cur.swap(nxt);
    }
    }
    FinishArgs F{cur.data(), g, P.Wb, P.c, P.NS, P.log_l, out};
    for (size_t v = 0; v < g; ++v) finish_body(v, F);
    if (stats) {
        stats[0] = P.c;
        stats[1] = P.W;
        stats[2] = (uint32_t)P.acc_entries.size();
        stats[3] = P.sum_chunks;
        stats[4] = g_last_pairs;
        stats[5] = g_last_dominant;
    }
    return 0;
}

// element-wise checks of the portable field / curve bodies (op: 0 fq_mul 1 fq_add 2 fq_sub 3 fq_sqr 4 fq_inv 5 fr_from_mont)
void emul_field_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
    if (op >= 10) {
        // 9 x 29-bit representation, entered and left through the arkworks wire form: 10 mul 11 add_mod 12 sub_mod 13 sqr 14 inv
        for (size_t i = 0; i < n; ++i) {
            uint32_t wa[8], wb[8] = {0}, wr[8];
            memcpy(wa, a + 32 * i, 32);
            if (b) memcpy(wb, b + 32 * i, 32);
            f29::fe x = f29::from_ark(wa), y = f29::from_ark(wb), r;
            switch (op) {
                case 10: r = f29::mul(x, y); break;
                case 11: r = f29::add_mod(x, y); break;
                case 12: r = f29::sub_mod(x, y); break;
                case 13: r = f29::sqr(x); break;
                default: r = f29::inv(x); break;
            }
            f29::to_ark(r, wr);
            memcpy(out + 32 * i, wr, 32);
        }
        return;
    }
    for (size_t i = 0; i < n; ++i) {
        fq x = load_fq(a + 32 * i), y = b ? load_fq(b + 32 * i) : fq_zero(), r;
        switch (op) {
            case 0: r = fq_mul(x, y); break;
            case 1: r = fq_add(x, y); break;
            case 2: r = fq_sub(x, y); break;
            case 3: r = fq_sqr(x); break;
            case 4: r = fq_inv(x); break;
            default: r = fr_from_mont(x); break;
        }
        store_fq(out + 32 * i, r);
    }
}

// op: 0 xyzz_add 1 xyzz_madd (b finite) 2 xyzz_dbl; inputs/outputs are 72-byte wire points
void emul_g1_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
    if (op == 3) {
        // lazy mixed addition on the 29-bit field: ((a + b) + b) - b, which walks the lazy bounds through three calls
        for (size_t i = 0; i < n; ++i) {
            f29::xyzz29 acc = f29::from_wire29(a + 72 * i);
            f29::xyzz29 q = f29::from_wire29(b + 72 * i);
            f29::affine29 qa{q.X, q.Y}, qn{q.X, f29::sub<2>(f29::zero(), q.Y)};
            acc = f29::madd_lazy(acc, qa);
            acc = f29::madd_lazy(acc, qa);
            acc = f29::madd_lazy(acc, qn);
            f29::to_wire29(f29::normalize29(acc), out + 72 * i);
        }
        return;
    }
    for (size_t i = 0; i < n; ++i) {
        xyzz pa = xyzz_from_wire(a + 72 * i), r;
        if (op == 0) {
            r = xyzz_add(pa, xyzz_from_wire(b + 72 * i));
        } else if (op == 1) {
            affine q;
            q.x = load_fq(b + 72 * i);
            q.y = load_fq(b + 72 * i + 32);
            r = xyzz_madd(pa, q);
        } else {
            r = xyzz_dbl(pa);
        }
        xyzz_to_wire(r, out + 72 * i);
    }
}

// ---------------------------------------------------------------------------------------------- rep3_kernels.cuh bodies
void emul_ingest(uint8_t* data, size_t n_fr, uint32_t* bad) {
    IngestArgs A{data, n_fr, bad};
    for (size_t t = 0; t < n_fr; ++t) ingest_body(t, A);
}
void emul_widen(const uint8_t* src, uint32_t elem_bytes, uint32_t is_signed, size_t n, uint8_t* dst) {
    WidenArgs A{src, elem_bytes, is_signed, n, dst};
    for (size_t t = 0; t < n; ++t) widen_body(t, A);
}
static std::vector<PolyDesc> make_descs(const uint8_t* const* ptrs, const uint64_t* lens, const uint32_t* kinds, uint32_t k) {
    std::vector<PolyDesc> d(k);
    for (uint32_t j = 0; j < k; ++j) d[j] = PolyDesc{ptrs[j], lens[j], kinds[j], 0};
    return d;
}
// coeffs: k x 32 B Montgomery; out: n x 64 B (shared_out) or n x 32 B
void emul_lincomb(const uint8_t* const* ptrs, const uint64_t* lens, const uint32_t* kinds, uint32_t k, const uint8_t* coeffs,
                  uint32_t party, uint32_t shared_out, size_t n, uint8_t* out) {
    std::vector<PolyDesc> d = make_descs(ptrs, lens, kinds, k);
    std::vector<fr> c(2 * (size_t)k);
    for (uint32_t j = 0; j < k; ++j) {
        c[2 * j] = load_fq(coeffs + 32 * j);
        c[2 * j + 1] = fr_mont_from_canon(c[2 * j]);
    }
    LincombArgs A{d.data(), c.data(), k, party, shared_out, n, out};
    for (size_t i = 0; i < n; ++i) lincomb_body(i, A);
}
// the exchange step of a multi-device linear combination: partial joint polynomials (lincomb outputs) -> their sum
void emul_sum_partials(const uint8_t* const* ptrs, const uint64_t* lens, const uint32_t* kinds, uint32_t g, uint32_t party,
                       uint32_t shared_out, size_t n, uint8_t* out) {
    std::vector<PolyDesc> d = make_descs(ptrs, lens, kinds, g);
    SumPartialsArgs A{d.data(), g, party, shared_out, n, out};
    for (size_t i = 0; i < n; ++i) sum_partials_body(i, A);
}
void emul_chi(const uint8_t* const* ptrs, const uint64_t* lens, const uint32_t* kinds, uint32_t k, const uint8_t* chis, size_t n,
              uint32_t T, uint8_t* out) {
    std::vector<PolyDesc> d = make_descs(ptrs, lens, kinds, k);
    std::vector<fr> part((size_t)k * T);
    ChiArgs A{d.data(), k, reinterpret_cast<const fr*>(chis), n, T, part.data()};
    for (size_t t = 0; t < (size_t)k * T; ++t) chi_partial_body(t, A);
    uint32_t Tmid = T < 64 ? T : 64;
    std::vector<fr> mid((size_t)k * Tmid);
    ChiReduceArgs R1{d.data(), k, part.data(), T, mid.data(), Tmid, 0}, R2{d.data(), k, mid.data(), Tmid, reinterpret_cast<fr*>(out), 1, 1};
    for (size_t t = 0; t < (size_t)k * Tmid; ++t) chi_reduce_body(t, R1);
    for (size_t j = 0; j < k; ++j) chi_reduce_body(j, R2);
}
// eq table (both kernels of cozk_eq_evals): point nv x 32 B Montgomery -> out 2^nv x 32 B
void emul_eq(const uint8_t* point, uint32_t nv, uint32_t lo_bits, int msb_first, uint8_t* out) {
    std::vector<fr> lo((size_t)1 << lo_bits), hi((size_t)1 << (nv - lo_bits));
    EqArgs A{reinterpret_cast<const fr*>(point), nv, lo_bits, msb_first, lo.data(), hi.data(), reinterpret_cast<fr*>(out)};
    for (size_t t = 0; t < lo.size() + hi.size(); ++t) eq_small_body(t, A);
    for (size_t b = 0; b < ((size_t)1 << nv); ++b) eq_expand_body(b, A);
}
void emul_pair_sum(const uint8_t* bases, const uint8_t* infinity, size_t half, uint8_t* out, uint8_t* out_inf) {
    PairSumArgs A{reinterpret_cast<const affine*>(bases), infinity, half, reinterpret_cast<affine*>(out), out_inf};
    for (size_t b = 0; b < half; ++b) pair_sum_body(b, A);
}
// out = (sum_i a_i * b_i) / R mod r through the lazy accumulator, `repeat` times over (pushes the top limbs)
void emul_wide_dot(const uint8_t* a, const uint8_t* b, size_t n, uint32_t repeat, uint8_t* out) {
    fr_wide acc = fr_wide_zero();
    for (uint32_t r = 0; r < repeat; ++r)
        for (size_t i = 0; i < n; ++i) fr_wide_mac(acc, load_fq(a + 32 * i), load_fq(b + 32 * i));
    store_fq(out, fr_wide_reduce(acc));
}
}
