// Device-resident Rep3 polynomials (include/cozk_rep3.h): wire ingestion (N2), linear combination and chi dot products
// (N4), pair-sum SRS and the opening of a resident polynomial (N1).  Thread bodies live in rep3_kernels.cuh.
#include "../../include/cozk_rep3.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <thread>
#include <vector>

#include "engine.hpp"
#include "../../include/cozk_pst13.h"
#include "bulk_copy.cuh"
#include "rep3_kernels.cuh"

namespace cozk {

// Four elements per thread, a grid-stride apart, their loads issued before the first conversion.  One element per thread
// (8.4 M threads = 32768 blocks that live for a microsecond each at 2^22 coefficients) is bound by the rate at which the
// SMs can start blocks, not by HBM: 0.152 ms = 0.54 of the measured copy bandwidth.  Output contract = ingest_body.
constexpr int ING_UNROLL = 4;
__global__ void __launch_bounds__(256) k_ingest(IngestArgs A) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    fr x[ING_UNROLL];
#pragma unroll
    for (int u = 0; u < ING_UNROLL; ++u)
        if (t + u * stride < A.n_fr) x[u] = load_fq(A.data + 32 * (t + u * stride));
#pragma unroll
    for (int u = 0; u < ING_UNROLL; ++u) {
        if (t + u * stride >= A.n_fr) continue;
        if (!fr_is_canonical(x[u])) *A.bad = 1;  // every writer stores the same value
        else store_fq(A.data + 32 * (t + u * stride), fr_mont_from_canon(x[u]));
    }
}

// k_ingest with the data streamed through shared memory by the copy engine.  Output contract = ingest_body for every
// element (that body is what the host tier runs).  Both HBM directions of this in-place conversion are bulk copies of
// whole 8 KiB slabs: a block owns tiles b, b + grid, ..; one elected thread keeps two loads in flight ahead of the slab
// being converted and sends every converted slab back with a bulk store; the 256 threads only touch shared memory.
constexpr int ING_THREADS = 256, ING_STAGES = 3;
__global__ void __launch_bounds__(ING_THREADS) k_ingest_bulk(IngestArgs A) {
    __shared__ __align__(128) uint8_t buf[ING_STAGES][ING_THREADS * 32];
    __shared__ __align__(8) uint64_t bar[ING_STAGES];
    const uint32_t t = threadIdx.x;
    const size_t tiles = (A.n_fr + ING_THREADS - 1) / ING_THREADS;
    auto tile_bytes = [&](size_t tile) -> uint32_t {
        const size_t left = A.n_fr - tile * ING_THREADS;
        return (uint32_t)(left < ING_THREADS ? left : ING_THREADS) * 32u;
    };
    if (t == 0) {
        for (int s = 0; s < ING_STAGES; ++s) mbar_init(&bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const size_t first = blockIdx.x, step = gridDim.x;
    if (t == 0) {
        for (int s = 0; s < ING_STAGES - 1; ++s) {
            const size_t tile = first + (size_t)s * step;
            if (tile < tiles) {
                mbar_arrive_expect(&bar[s], tile_bytes(tile));
                bulk_load(buf[s], A.data + tile * ING_THREADS * 32, tile_bytes(tile), &bar[s]);
            }
        }
    }
    uint32_t it = 0;
    for (size_t tile = first; tile < tiles; tile += step, ++it) {
        const int s = it % ING_STAGES;
        if (t == 0) {
            const size_t ahead = tile + (size_t)(ING_STAGES - 1) * step;
            if (ahead < tiles) {
                const int sa = (it + ING_STAGES - 1) % ING_STAGES;  // the stage the slab before this one was stored from
                bulk_wait_read<0>();
                mbar_arrive_expect(&bar[sa], tile_bytes(ahead));
                bulk_load(buf[sa], A.data + ahead * ING_THREADS * 32, tile_bytes(ahead), &bar[sa]);
            }
        }
        mbar_wait(&bar[s], (it / ING_STAGES) & 1u);
        if (tile * ING_THREADS + t < A.n_fr) {
            fr x = load_fq(buf[s] + 32 * t);
            if (!fr_is_canonical(x)) *A.bad = 1;  // every writer stores the same value; the slab goes back unconverted
            else store_fq(buf[s] + 32 * t, fr_mont_from_canon(x));
        }
        fence_async_smem();
        __syncthreads();
        if (t == 0) {
            bulk_store(A.data + tile * ING_THREADS * 32, buf[s], tile_bytes(tile));
            bulk_commit();
        }
    }
    if (t == 0) bulk_wait_all();
}

__global__ void __launch_bounds__(256) k_widen(WidenArgs A) { widen_body((size_t)blockIdx.x * blockDim.x + threadIdx.x, A); }
// descriptors and coefficients are the same for every thread: each block stages them in shared memory once
// (k * (24 + 64) bytes) so that the inner loop reads them with broadcast loads instead of going through L1/L2
__global__ void __launch_bounds__(128, 4) k_lincomb(LincombArgs A, int staged) {
    extern __shared__ uint4 lincomb_smem[];
    if (staged) {
        PolyDesc* sd = reinterpret_cast<PolyDesc*>(lincomb_smem);
        fr* sc = reinterpret_cast<fr*>(reinterpret_cast<uint8_t*>(lincomb_smem) + (((size_t)A.k * sizeof(PolyDesc) + 15) & ~(size_t)15));
        for (uint32_t j = threadIdx.x; j < A.k; j += blockDim.x) sd[j] = A.polys[j];
        for (uint32_t j = threadIdx.x; j < 2 * A.k; j += blockDim.x) store_fq(&sc[j], load_fq(&A.coeffs[j]));
        __syncthreads();
        A.polys = sd;
        A.coeffs = sc;
    }
    lincomb_body((size_t)blockIdx.x * blockDim.x + threadIdx.x, A);
}
__global__ void __launch_bounds__(128) k_chi_partial(ChiArgs A) { chi_partial_body((size_t)blockIdx.x * blockDim.x + threadIdx.x, A); }

// k_chi_partial with the polynomial and chi slabs streamed through shared memory by the copy engine.  Same partial sums
// as chi_partial_body (thread t of polynomial j adds elements t, t + T, ..): a block of 128 threads owns 128 consecutive
// t, so round r needs ONE contiguous slab of the polynomial (128 x 64 or 32 bytes) and one of the chi table (128 x 32):
// an elected thread keeps CHI_STAGES - 1 rounds in flight as bulk copies, the others read their element from shared
// memory - no lane fetches 16 bytes at a 64-byte stride from L1 any more.
constexpr int CHI_THREADS = 128, CHI_STAGES = 3;  // 3 x 12 KiB of slabs: inside the default 48 KiB of dynamic shared memory
struct ChiStage {
    uint8_t poly[CHI_THREADS * 64];
    uint8_t chi[CHI_THREADS * 32];
};
__global__ void __launch_bounds__(CHI_THREADS) k_chi_partial_bulk(ChiArgs A) {
    extern __shared__ __align__(128) uint8_t chi_smem_raw[];
    ChiStage* st = reinterpret_cast<ChiStage*>(chi_smem_raw);
    __shared__ __align__(8) uint64_t bar[CHI_STAGES];
    const uint32_t t = threadIdx.x;
    const uint32_t blocks_per_poly = A.T / CHI_THREADS;  // T is a multiple of the block size (power of two >= 128)
    const uint32_t j = blockIdx.x / blocks_per_poly;
    const size_t tbase = (size_t)(blockIdx.x - j * blocks_per_poly) * CHI_THREADS;
    const PolyDesc d = A.polys[j];
    const size_t lim = d.len < A.n ? d.len : A.n;
    const uint32_t eb = d.kind == POLY_SHARED ? 64u : 32u;
    const size_t rounds = lim > tbase ? (lim - tbase + A.T - 1) / A.T : 0;
    auto issue = [&](size_t r) {  // elected thread: slab of round r into its stage
        const int s = (int)(r % CHI_STAGES);
        const size_t e0 = r * A.T + tbase;
        const uint32_t cnt = (uint32_t)(lim - e0 < CHI_THREADS ? lim - e0 : CHI_THREADS);
        mbar_arrive_expect(&bar[s], cnt * (eb + 32u));
        bulk_load(st[s].poly, d.data + e0 * eb, cnt * eb, &bar[s]);
        bulk_load(st[s].chi, reinterpret_cast<const uint8_t*>(A.chis) + e0 * 32, cnt * 32u, &bar[s]);
    };
    if (t == 0) {
        for (int s = 0; s < CHI_STAGES; ++s) mbar_init(&bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (t == 0)
        for (size_t r = 0; r < (size_t)(CHI_STAGES - 1) && r < rounds; ++r) issue(r);
    fr_wide acc = fr_wide_zero();
    for (size_t r = 0; r < rounds; ++r) {
        const int s = (int)(r % CHI_STAGES);
        // the stage of round r - 1 has been read by everybody (barrier at the end of that round): refill it
        if (t == 0 && r + CHI_STAGES - 1 < rounds) issue(r + CHI_STAGES - 1);
        mbar_wait(&bar[s], (uint32_t)(r / CHI_STAGES) & 1u);
        if (r * A.T + tbase + t < lim) {
            fr v;
            if (d.kind == POLY_SHARED) v = fr_add(load_fq(st[s].poly + 64 * t), load_fq(st[s].poly + 64 * t + 32));
            else {
                v = load_fq(st[s].poly + 32 * t);
                if (d.kind == POLY_CANON) v = fr_mont_from_canon(v);
            }
            fr_wide_mac(acc, v, load_fq(st[s].chi + 32 * t));
        }
        __syncthreads();
    }
    store_fq(&A.partial[(size_t)j * A.T + tbase + t], fr_wide_reduce(acc));
}
__global__ void k_chi_reduce(ChiReduceArgs A) { chi_reduce_body((size_t)blockIdx.x * blockDim.x + threadIdx.x, A); }
__global__ void __launch_bounds__(256) k_sum_partials(SumPartialsArgs A) { sum_partials_body((size_t)blockIdx.x * blockDim.x + threadIdx.x, A); }
__global__ void __launch_bounds__(128) k_eq_small(EqArgs A) { eq_small_body((size_t)blockIdx.x * blockDim.x + threadIdx.x, A); }
__global__ void __launch_bounds__(256) k_eq_expand(EqArgs A) { eq_expand_body((size_t)blockIdx.x * blockDim.x + threadIdx.x, A); }
__global__ void __launch_bounds__(128) k_pair_sum(PairSumArgs A) { pair_sum_body((size_t)blockIdx.x * blockDim.x + threadIdx.x, A); }
// Polynomials come and go by the hundred (138 trace polynomials per proof): allocate them stream-ordered from the device's
// memory pool, whose release threshold cozk_init lifts, so that a freed polynomial's memory is reused instead of being
// handed back to the driver (cudaMalloc / cudaFree of 64 MiB blocks cost more than the ingestion kernel itself).
template <class T>
static inline cudaError_t pool_alloc(Device& D, T** out, size_t bytes) {
    return cudaMallocAsync(reinterpret_cast<void**>(out), bytes ? bytes : 16, D.stream);
}
static inline cudaError_t pool_free(Device& D, void* p) { return cudaFreeAsync(p, D.stream); }

static inline unsigned blocks_for(size_t threads, unsigned block) { return (unsigned)((threads + block - 1) / block); }

static int check_device(cozk_ctx* ctx, int device_index) {
    if (!ctx || device_index < 0 || device_index >= (int)ctx->devs.size()) {
        set_error("null context or device index out of range");
        return COZK_ERR_INVALID_ARG;
    }
    return COZK_OK;
}

static int lookup(cozk_ctx* ctx, cozk_poly h, PolyEntry* out) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    auto it = ctx->polys.find(h);
    if (it == ctx->polys.end()) {
        set_error("unknown polynomial handle");
        return COZK_ERR_BAD_HANDLE;
    }
    *out = it->second;
    return COZK_OK;
}

static cozk_poly publish(cozk_ctx* ctx, const PolyEntry& E0) {
    PolyEntry E = E0;
    if (E.d_data && !E.hold) {
        const cudaStream_t st = ctx->devs[E.dev]->stream;
        const int id = ctx->devs[E.dev]->id;
        E.hold = std::shared_ptr<void>(E.d_data, [st, id](void* p) {
            int cur = -1;
            cudaGetDevice(&cur);
            cudaSetDevice(id);
            cudaFreeAsync(p, st);  // ordered behind everything already queued on the owning device's stream
            if (cur >= 0) cudaSetDevice(cur);
        });
    }
    std::lock_guard<std::mutex> lock(ctx->mu);
    cozk_poly h = ctx->next_handle++;
    ctx->polys[h] = E;
    return h;
}

struct StageTimer {
    Device& D;
    cudaEvent_t a = nullptr, b = nullptr;
    explicit StageTimer(Device& d) : D(d) {
        cudaEventCreate(&a);
        cudaEventCreate(&b);
    }
    ~StageTimer() {
        if (a) cudaEventDestroy(a);
        if (b) cudaEventDestroy(b);
    }
    void start() { cudaEventRecord(a, D.stream); }
    double stop() {
        cudaEventRecord(b, D.stream);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        return ms;
    }
};

// little-endian reader over the wire image
struct Reader {
    const uint8_t* p;
    size_t n, off = 0;
    bool ok = true;
    uint64_t u64() {
        if (!ok || n - off < 8) {
            ok = false;
            return 0;
        }
        uint64_t v;
        memcpy(&v, p + off, 8);
        off += 8;
        return v;
    }
    uint8_t u8() {
        if (!ok || n - off < 1) {
            ok = false;
            return 0;
        }
        return p[off++];
    }
    // skips count * 64 bytes; false when the image is too short (also guards count * 64 against overflow)
    bool skip_shares(uint64_t count) {
        if (!ok || count > (n - off) / 64) {
            ok = false;
            return false;
        }
        off += (size_t)count * 64;
        return true;
    }
};

int srs_pair_sums_into(cozk_ctx* ctx, cozk_srs srs, affine* d_out, uint8_t* d_inf, size_t* half_out) {
    SrsEntry S;
    int lrc = srs_lookup(ctx, srs, &S);
    if (lrc) return lrc;
    if (!S.bases(0)) {
        set_error("pair sums need an SRS that lives on device 0 (not a sliced one)");
        return COZK_ERR_INVALID_ARG;
    }
    if (S.n == 0 || (S.n & 1)) {
        set_error("pair sums need an even, non-zero number of bases");
        return COZK_ERR_INVALID_ARG;
    }
    Device& D = *ctx->devs[0];
    size_t half = S.n / 2;
    std::lock_guard<std::mutex> lock(D.mu);
    COZK_CUDA(cudaSetDevice(D.id));
    PairSumArgs A{S.bases(0), S.inf(0), half, d_out, d_inf};  // row 0 of the table = the bases themselves
    k_pair_sum<<<blocks_for(half, 128), 128, 0, D.stream>>>(A);
    COZK_CUDA(cudaGetLastError());
    COZK_CUDA(cudaStreamSynchronize(D.stream));
    if (half_out) *half_out = half;
    return COZK_OK;
}

}  // namespace cozk

using namespace cozk;

extern "C" {

int cozk_poly_upload(cozk_ctx* ctx, int device_index, const void* coeffs, size_t len, int kind, cozk_poly* out) {
    int rc = check_device(ctx, device_index);
    if (rc) return rc;
    if (!out || (!coeffs && len) || kind < COZK_POLY_SHARED || kind > COZK_POLY_I64) {
        set_error("null pointer or unknown polynomial kind");
        return COZK_ERR_INVALID_ARG;
    }
    Device& D = *ctx->devs[device_index];
    PolyEntry E;
    E.dev = device_index;
    E.user_kind = kind;
    E.total = E.len = len;
    static const unsigned small_bytes[7] = {0, 0, 1, 2, 4, 8, 8};
    E.kind = kind == COZK_POLY_SHARED ? POLY_SHARED : (kind == COZK_POLY_PUBLIC ? POLY_MONT : POLY_CANON);
    E.bits = (kind >= COZK_POLY_U8 && kind <= COZK_POLY_U64) ? 8 * small_bytes[kind] : 0;
    std::lock_guard<std::mutex> lock(D.mu);
    COZK_CUDA(cudaSetDevice(D.id));
    StageTimer T(D);
    uint8_t* d = nullptr;
    COZK_CUDA(pool_alloc(D, &d, std::max<size_t>(len, 1) * E.elem_bytes()));
    E.d_data = d;
    double h2d = 0, conv = 0;
    cudaError_t e = cudaSuccess;
    if (kind <= COZK_POLY_PUBLIC) {
        T.start();
        e = cudaMemcpyAsync(d, coeffs, len * E.elem_bytes(), cudaMemcpyHostToDevice, D.stream);
        h2d = T.stop();
    } else if (len) {
        uint8_t* raw = nullptr;
        size_t rb = len * small_bytes[kind];
        e = pool_alloc(D, &raw, rb);
        if (e == cudaSuccess) {
            T.start();
            e = cudaMemcpyAsync(raw, coeffs, rb, cudaMemcpyHostToDevice, D.stream);
            h2d = T.stop();
        }
        if (e == cudaSuccess) {
            WidenArgs A{raw, small_bytes[kind], kind == COZK_POLY_I64 ? 1u : 0u, len, d};
            T.start();
            k_widen<<<blocks_for(len, 256), 256, 0, D.stream>>>(A);
            e = cudaGetLastError();
            conv = T.stop();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(D.stream);
        if (raw) pool_free(D, raw);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(D.stream);
    if (e != cudaSuccess) {
        pool_free(D, d);
        set_error(std::string("polynomial upload failed: ") + cudaGetErrorString(e));
        return COZK_ERR_CUDA;
    }
    ctx->rep3_stats[0] = h2d;
    ctx->rep3_stats[1] = conv;
    *out = publish(ctx, E);
    return COZK_OK;
}

int cozk_poly_from_device(cozk_ctx* ctx, int device_index, const void* d_coeffs, size_t len, int kind, cozk_poly* out) {
    int rc = check_device(ctx, device_index);
    if (rc) return rc;
    if (!out || (!d_coeffs && len) || (kind != COZK_POLY_SHARED && kind != COZK_POLY_PUBLIC)) {
        set_error("null pointer, or a kind other than SHARED / PUBLIC");
        return COZK_ERR_INVALID_ARG;
    }
    Device& D = *ctx->devs[device_index];
    PolyEntry E;
    E.dev = device_index;
    E.user_kind = kind;
    E.kind = kind == COZK_POLY_SHARED ? POLY_SHARED : POLY_MONT;
    E.total = E.len = len;
    std::lock_guard<std::mutex> lock(D.mu);
    COZK_CUDA(cudaSetDevice(D.id));
    uint8_t* d = nullptr;
    COZK_CUDA(pool_alloc(D, &d, std::max<size_t>(len, 1) * E.elem_bytes()));
    cudaError_t e = cudaMemcpyAsync(d, d_coeffs, len * E.elem_bytes(), cudaMemcpyDeviceToDevice, D.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(D.stream);
    if (e != cudaSuccess) {
        pool_free(D, d);
        set_error(std::string("polynomial copy failed: ") + cudaGetErrorString(e));
        return COZK_ERR_CUDA;
    }
    E.d_data = d;
    *out = publish(ctx, E);
    return COZK_OK;
}

int cozk_poly_from_wire(cozk_ctx* ctx, int device_index, const void* bytes, size_t nbytes, int tagged, cozk_poly* out,
                        size_t* consumed) {
    int rc = check_device(ctx, device_index);
    if (rc) return rc;
    if (!bytes || !out) {
        set_error("null pointer");
        return COZK_ERR_INVALID_ARG;
    }
    Reader R{reinterpret_cast<const uint8_t*>(bytes), nbytes};
    if (tagged) {
        uint8_t tag = R.u8();
        if (R.ok && tag != 1) {
            set_error(tag == 0 ? "Rep3MultilinearPolynomial::Public on the wire: jolt-core's MultilinearPolynomial format is not "
                                 "defined in the reference tree; upload it with cozk_poly_upload"
                               : "unknown Rep3MultilinearPolynomial discriminant");
            return COZK_ERR_WIRE;
        }
    }
    uint64_t num_vars = R.u64();
    uint64_t n_coeffs = R.u64();
    size_t coeff_off = R.off;
    R.skip_shares(n_coeffs);
    uint64_t n_bound = R.u64();
    R.skip_shares(n_bound);
    uint8_t has_scratch = R.u8();
    if (R.ok && has_scratch > 1) {
        set_error("wire image: Option discriminant is not 0/1");
        return COZK_ERR_WIRE;
    }
    if (has_scratch) R.skip_shares(R.u64());
    uint64_t len = R.u64();
    uint64_t lo = R.u64(), hi = R.u64();
    if (!R.ok) {
        set_error("wire image: unexpected end of input");
        return COZK_ERR_WIRE;
    }
    if (lo > hi || hi > n_coeffs || num_vars > 63) {
        set_error("wire image: chunk_range outside the coefficient vector");
        return COZK_ERR_WIRE;
    }
    (void)len;  // the reference's own accessors (coeffs_ref, copy_share_a) go by chunk_range
    Device& D = *ctx->devs[device_index];
    PolyEntry E;
    E.dev = device_index;
    E.user_kind = COZK_POLY_SHARED;
    E.kind = POLY_SHARED;
    E.total = (size_t)n_coeffs;
    E.lo = (size_t)lo;
    E.len = (size_t)(hi - lo);
    std::lock_guard<std::mutex> lock(D.mu);
    COZK_CUDA(cudaSetDevice(D.id));
    StageTimer T(D);
    uint8_t* d = nullptr;
    uint32_t* d_bad = nullptr;
    COZK_CUDA(pool_alloc(D, &d, std::max<size_t>(E.total, 1) * 64 + 16));
    cudaError_t e = pool_alloc(D, &d_bad, 4);
    uint32_t bad = 0;
    double h2d = 0, conv = 0;
    if (e == cudaSuccess) e = cudaMemsetAsync(d_bad, 0, 4, D.stream);
    if (e == cudaSuccess) {
        T.start();
        e = cudaMemcpyAsync(d, R.p + coeff_off, E.total * 64, cudaMemcpyHostToDevice, D.stream);
        h2d = T.stop();
    }
    if (e == cudaSuccess && E.total) {
        IngestArgs A{d, 2 * E.total, d_bad};
        T.start();
        // persistent blocks, a few per SM, each walking its tiles (k_ingest - the plain form - is what the host tier checks)
        const unsigned tiles = blocks_for(A.n_fr, ING_THREADS);
        if (ctx->opt_bulk_copy) k_ingest_bulk<<<std::min<unsigned>(tiles, (unsigned)D.sm_count * 8), ING_THREADS, 0, D.stream>>>(A);
        else k_ingest<<<blocks_for((A.n_fr + ING_UNROLL - 1) / ING_UNROLL, 256), 256, 0, D.stream>>>(A);
        e = cudaGetLastError();
        conv = T.stop();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, D.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(D.stream);
    if (d_bad) pool_free(D, d_bad);
    if (e != cudaSuccess || bad) {
        pool_free(D, d);
        if (e != cudaSuccess) {
            set_error(std::string("wire ingestion failed: ") + cudaGetErrorString(e));
            return COZK_ERR_CUDA;
        }
        set_error("wire image: field element is not below the modulus");
        return COZK_ERR_WIRE;
    }
    E.d_data = d;
    ctx->rep3_stats[0] = h2d;
    ctx->rep3_stats[1] = conv;
    if (consumed) *consumed = R.off;
    *out = publish(ctx, E);
    return COZK_OK;
}

int cozk_poly_release(cozk_ctx* ctx, cozk_poly poly) {
    if (!ctx) return COZK_ERR_INVALID_ARG;
    PolyEntry E;
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        auto it = ctx->polys.find(poly);
        if (it == ctx->polys.end()) {
            set_error("unknown polynomial handle");
            return COZK_ERR_BAD_HANDLE;
        }
        E = it->second;
        ctx->polys.erase(it);
    }
    return COZK_OK;  // E leaves scope outside the context lock; the last copy of the entry frees the device memory
}

int cozk_poly_info(cozk_ctx* ctx, cozk_poly poly, size_t* len, int* kind, int* device_index) {
    if (!ctx) return COZK_ERR_INVALID_ARG;
    PolyEntry E;
    int rc = lookup(ctx, poly, &E);
    if (rc) return rc;
    if (len) *len = E.len;
    if (kind) *kind = E.user_kind;
    if (device_index) *device_index = E.dev;
    return COZK_OK;
}

int cozk_poly_download(cozk_ctx* ctx, cozk_poly poly, void* out) {
    if (!ctx || !out) {
        set_error("null pointer");
        return COZK_ERR_INVALID_ARG;
    }
    PolyEntry E;
    int rc = lookup(ctx, poly, &E);
    if (rc) return rc;
    Device& D = *ctx->devs[E.dev];
    std::lock_guard<std::mutex> lock(D.mu);
    COZK_CUDA(cudaSetDevice(D.id));
    COZK_CUDA(cudaMemcpyAsync(out, E.chunk(), E.len * E.elem_bytes(), cudaMemcpyDeviceToHost, D.stream));
    COZK_CUDA(cudaStreamSynchronize(D.stream));
    return COZK_OK;
}

int cozk_pst13_batch_commit_polys(cozk_ctx* ctx, cozk_srs srs, const cozk_poly* polys, size_t k, int commit_to_public,
                                  void* out_commitments, uint8_t* present) {
    if (!ctx || !polys || !out_commitments || !present || k == 0) {
        set_error("null pointer or k == 0");
        return COZK_ERR_INVALID_ARG;
    }
    std::vector<PolyEntry> E(k);
    for (size_t j = 0; j < k; ++j) {
        int rc = lookup(ctx, polys[j], &E[j]);
        if (rc) return rc;
    }
    size_t n = E[0].len;
    unsigned nv = 0;
    while (((size_t)1 << nv) < n) ++nv;
    if (n == 0 || ((size_t)1 << nv) != n) {
        set_error("polynomial length must be a power of two");
        return COZK_ERR_INVALID_ARG;
    }
    for (size_t j = 0; j < k; ++j) {
        if (E[j].len != n) {
            set_error("batch_commit: polynomials must have equal lengths");  // assert at pst13.rs:307-309
            return COZK_ERR_INVALID_ARG;
        }
    }
    uint8_t* out = reinterpret_cast<uint8_t*>(out_commitments);
    memset(out, 0, k * COZK_COMMITMENT_BYTES);
    // one MSM batch per (device, scalar layout, bit-width hint)
    struct Key {
        int dev;
        uint32_t kind;
        unsigned bits;
        bool operator<(const Key& o) const {
            return dev != o.dev ? dev < o.dev : (kind != o.kind ? kind < o.kind : bits < o.bits);
        }
    };
    std::map<Key, std::vector<size_t>> groups;
    for (size_t j = 0; j < k; ++j) {
        bool shared = E[j].kind == POLY_SHARED;
        present[j] = (shared || commit_to_public) ? 1 : 0;
        if (present[j]) groups[Key{E[j].dev, E[j].kind, E[j].bits}].push_back(j);
    }
    uint64_t nv64 = nv;
    // the groups of one device run one after the other, the devices side by side (one host thread each)
    std::map<int, std::vector<const Key*>> per_dev;
    for (auto& kv : groups) per_dev[kv.first.dev].push_back(&kv.first);
    std::vector<int> rcs(per_dev.size(), COZK_OK);
    std::vector<std::string> errs(per_dev.size());
    auto run_device = [&](size_t slot, const std::vector<const Key*>& keys) {
        for (const Key* key : keys) {
            const std::vector<size_t>& members = groups[*key];
            std::vector<const void*> ptrs;
            for (size_t j : members) ptrs.push_back(E[j].chunk());
            std::vector<uint8_t> pts(ptrs.size() * 72);
            size_t stride = key->kind == POLY_SHARED ? 64 : 32;
            int form = key->kind == POLY_CANON ? COZK_CANON : COZK_MONT;
            int rc = msm_dispatch(ctx, key->dev, srs, 0, n, nullptr, ptrs.data(), ptrs.size(), stride, form, key->bits, pts.data());
            if (rc) {
                rcs[slot] = rc;
                errs[slot] = cozk_last_error();
                return;
            }
            for (size_t i = 0; i < members.size(); ++i) {
                uint8_t* o = out + COZK_COMMITMENT_BYTES * members[i];
                memcpy(o, &nv64, 8);
                memcpy(o + 8, &pts[72 * i], 72);
            }
        }
    };
    if (per_dev.size() == 1) {
        run_device(0, per_dev.begin()->second);
    } else {
        std::vector<std::thread> th;
        size_t slot = 0;
        for (auto& kv : per_dev) th.emplace_back(run_device, slot++, std::cref(kv.second));
        for (auto& t : th) t.join();
    }
    for (size_t i = 0; i < rcs.size(); ++i) {
        if (rcs[i]) {
            set_error(errs[i]);
            return rcs[i];
        }
    }
    return COZK_OK;
}

// Linear combination of polynomials that all live on ctx->devs[dev]; coeffs: one Montgomery value per polynomial.
// shared_out: the result is an array of shares (public terms follow add_public) / a dense public polynomial.
static int lincomb_on_device(cozk_ctx* ctx, int dev, const std::vector<PolyEntry>& E, const std::vector<fr>& coeffs, int party_id,
                             bool shared_out, PolyEntry* out, double* ms_out, double* bytes_out) {
    const size_t k = E.size();
    Device& D = *ctx->devs[dev];
    size_t max_len = 0;
    // coefficient pairs: [2j] = c_j (Montgomery), [2j+1] = c_j * R for terms whose values are canonical integers
    std::vector<fr> hc(2 * k);
    std::vector<PolyDesc> hd(k);
    double bytes = 0;
    for (size_t j = 0; j < k; ++j) {
        hc[2 * j] = coeffs[j];
        hc[2 * j + 1] = fr_mont_from_canon(coeffs[j]);
        hd[j] = PolyDesc{E[j].chunk(), E[j].len, E[j].kind, 0};
        bytes += (double)E[j].len * E[j].elem_bytes();
        max_len = std::max(max_len, E[j].len);
    }
    PolyEntry O;
    O.dev = dev;
    O.kind = shared_out ? POLY_SHARED : POLY_MONT;
    O.user_kind = shared_out ? COZK_POLY_SHARED : COZK_POLY_PUBLIC;
    O.total = O.len = max_len;
    bytes += (double)max_len * O.elem_bytes();
    std::lock_guard<std::mutex> lock(D.mu);
    COZK_CUDA(cudaSetDevice(D.id));
    PolyDesc* d_desc = nullptr;
    fr* d_coef = nullptr;
    uint8_t* d_out = nullptr;
    cudaError_t e = pool_alloc(D, &d_desc, k * sizeof(PolyDesc));
    if (e == cudaSuccess) e = pool_alloc(D, &d_coef, 2 * k * sizeof(fr));
    if (e == cudaSuccess) e = pool_alloc(D, &d_out, std::max<size_t>(max_len, 1) * O.elem_bytes());
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_desc, hd.data(), k * sizeof(PolyDesc), cudaMemcpyHostToDevice, D.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_coef, hc.data(), 2 * k * sizeof(fr), cudaMemcpyHostToDevice, D.stream);
    double ms = 0;
    if (e == cudaSuccess && max_len) {
        LincombArgs A{d_desc, d_coef, (uint32_t)k, (uint32_t)party_id, shared_out ? 1u : 0u, max_len, d_out};
        StageTimer T(D);
        T.start();
        size_t smem = (((size_t)k * sizeof(PolyDesc) + 15) & ~(size_t)15) + 2 * k * sizeof(fr);
        int staged = smem <= 40 * 1024 ? 1 : 0;  // k <= 465; longer lists read the tables from global memory
        k_lincomb<<<blocks_for(max_len, 128), 128, staged ? smem : 0, D.stream>>>(A, staged);
        e = cudaGetLastError();
        ms = T.stop();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(D.stream);
    if (d_desc) pool_free(D, d_desc);
    if (d_coef) pool_free(D, d_coef);
    if (e != cudaSuccess) {
        if (d_out) pool_free(D, d_out);
        set_error(std::string("linear_combination failed: ") + cudaGetErrorString(e));
        return COZK_ERR_CUDA;
    }
    O.d_data = d_out;
    *out = O;
    *ms_out = ms;
    *bytes_out = bytes;
    return COZK_OK;
}

static void free_entry(cozk_ctx* ctx, const PolyEntry& E) {
    if (!E.d_data) return;
    Device& D = *ctx->devs[E.dev];
    std::lock_guard<std::mutex> lock(D.mu);
    cudaSetDevice(D.id);
    pool_free(D, E.d_data);
}

// Device `to` may read stream-ordered pool allocations of device `from` through a peer mapping (NVLink / NVSwitch on the
// 8 x B200 box).  False when the two devices have no peer path: the caller then stages a copy instead.
static bool peer_readable(cozk_ctx* ctx, int from, int to) {
    if (ctx->opt_peer_direct == 0) return false;
    const int from_id = ctx->devs[from]->id, to_id = ctx->devs[to]->id;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, to_id, from_id) != cudaSuccess || !can) return false;
    if (cudaSetDevice(to_id) != cudaSuccess) return false;
    cudaError_t e = cudaDeviceEnablePeerAccess(from_id, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        return false;
    }
    cudaGetLastError();
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, from_id) != cudaSuccess) return false;
    cudaMemAccessDesc desc = {};
    desc.location.type = cudaMemLocationTypeDevice;
    desc.location.id = to_id;
    desc.flags = cudaMemAccessFlagsProtReadWrite;
    if (cudaMemPoolSetAccess(pool, &desc, 1) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return true;
}

int cozk_rep3_linear_combination(cozk_ctx* ctx, const cozk_poly* polys, const void* coeffs, size_t k, int party_id,
                                 cozk_poly* out) {
    if (!ctx || !polys || !coeffs || !out || k == 0 || k > 0xFFFFFFFFu || party_id < 0 || party_id > 2) {
        set_error("null pointer, k == 0 or party id outside 0..2");
        return COZK_ERR_INVALID_ARG;
    }
    std::vector<PolyEntry> E(k);
    size_t max_len = 0, max_shared = 0;
    bool any_shared = false;
    std::map<int, std::vector<size_t>> by_dev;
    for (size_t j = 0; j < k; ++j) {
        int rc = lookup(ctx, polys[j], &E[j]);
        if (rc) return rc;
        by_dev[E[j].dev].push_back(j);
        max_len = std::max(max_len, E[j].len);
        if (E[j].kind == POLY_SHARED) {
            any_shared = true;
            max_shared = std::max(max_shared, E[j].len);
        }
    }
    if (any_shared && max_shared < max_len) {
        // the reference leaves such an index as SharedOrPublic::Public and panics in as_shared() ("Not an arithmetic share")
        set_error("linear_combination: a public polynomial is longer than every shared one");
        return COZK_ERR_INVALID_ARG;
    }
    std::vector<fr> hc(k);
    const uint8_t* cb = reinterpret_cast<const uint8_t*>(coeffs);
    for (size_t j = 0; j < k; ++j) memcpy(hc[j].v, cb + 32 * j, 32);

    if (by_dev.size() == 1) {
        PolyEntry O;
        double ms = 0, bytes = 0;
        int rc = lincomb_on_device(ctx, E[0].dev, E, hc, party_id, any_shared, &O, &ms, &bytes);
        if (rc) return rc;
        ctx->rep3_stats[2] = ms;
        ctx->rep3_stats[4] = bytes;
        ctx->rep3_stats[5] = 0;
        *out = publish(ctx, O);
        return COZK_OK;
    }

    // ---- polynomials on several devices: one partial joint polynomial per device (concurrently), then ONE kernel on
    // device 0 - where the opening runs - adds them, reading the remote partials through NVLink peer mappings.
    const size_t G = by_dev.size();
    std::vector<int> devs;
    for (auto& kv : by_dev) devs.push_back(kv.first);
    std::vector<PolyEntry> part(G);
    std::vector<int> rcs(G, COZK_OK);
    std::vector<std::string> errs(G);
    std::vector<double> mss(G, 0), bys(G, 0);
    {
        std::vector<std::thread> th;
        for (size_t g = 0; g < G; ++g) {
            th.emplace_back([&, g]() {
                const std::vector<size_t>& idx = by_dev[devs[g]];
                std::vector<PolyEntry> Eg;
                std::vector<fr> cg;
                bool shared_g = false;
                for (size_t j : idx) {
                    Eg.push_back(E[j]);
                    cg.push_back(hc[j]);
                    shared_g = shared_g || E[j].kind == POLY_SHARED;
                }
                // party 2 adds public terms to neither share; inside a partial the rule is the same, so it is applied
                // once per term whichever device holds it
                rcs[g] = lincomb_on_device(ctx, devs[g], Eg, cg, party_id, shared_g, &part[g], &mss[g], &bys[g]);
                if (rcs[g]) errs[g] = cozk_last_error();
            });
        }
        for (auto& t : th) t.join();
    }
    auto drop_partials = [&]() {
        for (size_t g = 0; g < G; ++g) free_entry(ctx, part[g]);
    };
    for (size_t g = 0; g < G; ++g) {
        if (rcs[g]) {
            drop_partials();
            set_error(errs[g]);
            return rcs[g];
        }
    }
    const int odev = 0;
    Device& D = *ctx->devs[odev];
    // remote partials: peer-mapped where the devices have a peer path, staged by a peer copy otherwise
    std::vector<PolyDesc> hd(G);
    std::vector<uint8_t*> staged;
    PolyEntry O;
    O.dev = odev;
    O.kind = any_shared ? POLY_SHARED : POLY_MONT;
    O.user_kind = any_shared ? COZK_POLY_SHARED : COZK_POLY_PUBLIC;
    O.total = O.len = max_len;
    double peer_bytes = 0;
    std::vector<bool> direct(G, true);
    for (size_t g = 0; g < G; ++g)
        if (part[g].dev != odev) direct[g] = peer_readable(ctx, part[g].dev, odev);
    std::lock_guard<std::mutex> lock(D.mu);
    cudaError_t e = cudaSetDevice(D.id);
    PolyDesc* d_desc = nullptr;
    uint8_t* d_out = nullptr;
    for (size_t g = 0; g < G && e == cudaSuccess; ++g) {
        const uint8_t* src = part[g].d_data;
        const size_t nbytes = part[g].len * part[g].elem_bytes();
        if (part[g].dev != odev) {
            peer_bytes += (double)nbytes;
            if (!direct[g]) {
                uint8_t* tmp = nullptr;
                e = pool_alloc(D, &tmp, std::max<size_t>(nbytes, 16));
                if (e == cudaSuccess) {
                    staged.push_back(tmp);
                    e = cudaMemcpyPeerAsync(tmp, D.id, src, ctx->devs[part[g].dev]->id, nbytes, D.stream);
                    src = tmp;
                }
            }
        }
        hd[g] = PolyDesc{src, part[g].len, part[g].kind, 0};
    }
    if (e == cudaSuccess) e = pool_alloc(D, &d_desc, G * sizeof(PolyDesc));
    if (e == cudaSuccess) e = pool_alloc(D, &d_out, std::max<size_t>(max_len, 1) * O.elem_bytes());
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_desc, hd.data(), G * sizeof(PolyDesc), cudaMemcpyHostToDevice, D.stream);
    double sum_ms = 0;
    if (e == cudaSuccess && max_len) {
        SumPartialsArgs A{d_desc, (uint32_t)G, (uint32_t)party_id, any_shared ? 1u : 0u, max_len, d_out};
        StageTimer T(D);
        T.start();
        k_sum_partials<<<blocks_for(max_len, 256), 256, 0, D.stream>>>(A);
        e = cudaGetLastError();
        sum_ms = T.stop();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(D.stream);
    if (d_desc) pool_free(D, d_desc);
    for (uint8_t* t : staged) pool_free(D, t);
    if (e != cudaSuccess) {
        if (d_out) pool_free(D, d_out);
        set_error(std::string("linear_combination (partial sums) failed: ") + cudaGetErrorString(e));
    }
    // the partials live on other devices (and on this one): release them without holding this device's lock twice
    for (size_t g = 0; g < G; ++g) {
        if (part[g].dev == odev) pool_free(D, part[g].d_data);
    }
    for (size_t g = 0; g < G; ++g) {
        if (part[g].dev != odev) free_entry(ctx, part[g]);
    }
    if (e != cudaSuccess) return COZK_ERR_CUDA;
    O.d_data = d_out;
    double ms_max = 0, bytes = 0;
    for (size_t g = 0; g < G; ++g) {
        ms_max = std::max(ms_max, mss[g]);
        bytes += bys[g];
    }
    ctx->rep3_stats[2] = ms_max;  // the partial combinations run concurrently: the slowest device
    ctx->rep3_stats[4] = bytes;
    ctx->rep3_stats[5] = sum_ms;  // the exchange step
    ctx->rep3_stats[6] = peer_bytes;
    *out = publish(ctx, O);
    return COZK_OK;
}

// chis: n Fr values, Montgomery, in host memory (h_chis) or already on the polynomials' device (d_chis_in)
static int evaluate_at_chi_core(cozk_ctx* ctx, const cozk_poly* polys, size_t k, const void* h_chis, const fr* d_chis_in,
                                int chis_dev, size_t n, void* out_evals) {
    if (!ctx || !polys || (!h_chis && !d_chis_in) || !out_evals || k == 0 || k > 0xFFFFFFFFu) {
        set_error("null pointer or k == 0");
        return COZK_ERR_INVALID_ARG;
    }
    std::vector<PolyEntry> E(k);
    for (size_t j = 0; j < k; ++j) {
        int rc = lookup(ctx, polys[j], &E[j]);
        if (rc) return rc;
        if (E[j].dev != E[0].dev) {
            set_error("evaluate_at_chi: all polynomials must live on one device");
            return COZK_ERR_INVALID_ARG;
        }
        if (E[j].len != n) {
            set_error("evaluate_at_chi: polynomial and chi lengths differ");  // zip_eq panics
            return COZK_ERR_INVALID_ARG;
        }
    }
    if (d_chis_in && chis_dev != E[0].dev) {
        set_error("evaluate_at_chi: the chi table must live on the polynomials' device");
        return COZK_ERR_INVALID_ARG;
    }
    Device& D = *ctx->devs[E[0].dev];
    std::vector<PolyDesc> hd(k);
    for (size_t j = 0; j < k; ++j) hd[j] = PolyDesc{E[j].chunk(), E[j].len, E[j].kind, 0};
    // threads per polynomial: enough to fill the chip a few times over, never more than one element per thread
    uint32_t T = 32;
    while ((size_t)T * 2 <= n && (size_t)T * k < (size_t)D.sm_count * 2048 * (size_t)ctx->opt_chi_waves && T < 16384) T *= 2;
    std::lock_guard<std::mutex> lock(D.mu);
    COZK_CUDA(cudaSetDevice(D.id));
    PolyDesc* d_desc = nullptr;
    fr *d_chis = nullptr, *d_part = nullptr, *d_mid = nullptr, *d_res = nullptr;
    const uint32_t Tmid = T < 64 ? T : 64;
    cudaError_t e = pool_alloc(D, &d_desc, k * sizeof(PolyDesc));
    if (e == cudaSuccess && !d_chis_in) e = pool_alloc(D, &d_chis, std::max<size_t>(n, 1) * sizeof(fr));
    if (e == cudaSuccess) e = pool_alloc(D, &d_part, k * (size_t)T * sizeof(fr));
    if (e == cudaSuccess) e = pool_alloc(D, &d_mid, k * (size_t)Tmid * sizeof(fr));
    if (e == cudaSuccess) e = pool_alloc(D, &d_res, k * sizeof(fr));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_desc, hd.data(), k * sizeof(PolyDesc), cudaMemcpyHostToDevice, D.stream);
    if (e == cudaSuccess && !d_chis_in) e = cudaMemcpyAsync(d_chis, h_chis, n * sizeof(fr), cudaMemcpyHostToDevice, D.stream);
    double ms = 0;
    if (e == cudaSuccess) {
        ChiArgs A{d_desc, (uint32_t)k, d_chis_in ? d_chis_in : d_chis, n, T, d_part};
        StageTimer St(D);
        St.start();
        if (T >= CHI_THREADS && ctx->opt_bulk_copy) {
            static_assert(sizeof(ChiStage) * CHI_STAGES + 1024 <= 48 * 1024, "chi stages fit the default dynamic shared-memory limit");
            k_chi_partial_bulk<<<(unsigned)(k * (size_t)(T / CHI_THREADS)), CHI_THREADS, sizeof(ChiStage) * CHI_STAGES, D.stream>>>(A);
        } else {
            k_chi_partial<<<blocks_for(k * (size_t)T, 128), 128, 0, D.stream>>>(A);  // tiny tables
        }
        ChiReduceArgs R1{d_desc, (uint32_t)k, d_part, T, d_mid, Tmid, 0}, R2{d_desc, (uint32_t)k, d_mid, Tmid, d_res, 1, 1};
        k_chi_reduce<<<blocks_for(k * (size_t)Tmid, 64), 64, 0, D.stream>>>(R1);
        k_chi_reduce<<<blocks_for(k, 64), 64, 0, D.stream>>>(R2);
        e = cudaGetLastError();
        ms = St.stop();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_evals, d_res, k * sizeof(fr), cudaMemcpyDeviceToHost, D.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(D.stream);
    for (void* p : {(void*)d_desc, (void*)d_chis, (void*)d_part, (void*)d_mid, (void*)d_res})
        if (p) pool_free(D, p);
    if (e != cudaSuccess) {
        set_error(std::string("evaluate_at_chi failed: ") + cudaGetErrorString(e));
        return COZK_ERR_CUDA;
    }
    ctx->rep3_stats[3] = ms;
    return COZK_OK;
}

int cozk_rep3_evaluate_at_chi(cozk_ctx* ctx, const cozk_poly* polys, size_t k, const void* chis, size_t n, void* out_evals) {
    if (!chis) {
        set_error("null pointer or k == 0");
        return COZK_ERR_INVALID_ARG;
    }
    return evaluate_at_chi_core(ctx, polys, k, chis, nullptr, 0, n, out_evals);
}

int cozk_rep3_evaluate_at_chi_poly(cozk_ctx* ctx, const cozk_poly* polys, size_t k, cozk_poly chi, void* out_evals) {
    if (!ctx) {
        set_error("null context");
        return COZK_ERR_INVALID_ARG;
    }
    PolyEntry C;
    int rc = lookup(ctx, chi, &C);
    if (rc) return rc;
    if (C.kind != POLY_MONT) {
        set_error("evaluate_at_chi: the chi table must be a public polynomial of field elements");
        return COZK_ERR_INVALID_ARG;
    }
    return evaluate_at_chi_core(ctx, polys, k, nullptr, reinterpret_cast<const fr*>(C.chunk()), C.dev, C.len, out_evals);
}

int cozk_eq_evals(cozk_ctx* ctx, int device_index, const void* point, size_t nv, int msb_first, cozk_poly* out) {
    int rc = check_device(ctx, device_index);
    if (rc) return rc;
    if (!point || !out || nv == 0 || nv > 30) {
        set_error("eq_evals: null pointer or nv outside 1..30");
        return COZK_ERR_INVALID_ARG;
    }
    Device& D = *ctx->devs[device_index];
    const size_t n = (size_t)1 << nv;
    const uint32_t lo_bits = (uint32_t)(nv / 2);
    fr *d_point = nullptr, *d_lo = nullptr, *d_hi = nullptr, *d_out = nullptr;
    {
        std::lock_guard<std::mutex> lock(D.mu);
        COZK_CUDA(cudaSetDevice(D.id));
        cudaError_t e = pool_alloc(D, &d_point, nv * sizeof(fr));
        if (e == cudaSuccess) e = pool_alloc(D, &d_lo, ((size_t)1 << lo_bits) * sizeof(fr));
        if (e == cudaSuccess) e = pool_alloc(D, &d_hi, ((size_t)1 << (nv - lo_bits)) * sizeof(fr));
        if (e == cudaSuccess) e = pool_alloc(D, &d_out, n * sizeof(fr));
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_point, point, nv * sizeof(fr), cudaMemcpyHostToDevice, D.stream);
        if (e == cudaSuccess) {
            EqArgs A{d_point, (uint32_t)nv, lo_bits, msb_first ? 1 : 0, d_lo, d_hi, d_out};
            k_eq_small<<<blocks_for(((size_t)1 << lo_bits) + ((size_t)1 << (nv - lo_bits)), 128), 128, 0, D.stream>>>(A);
            k_eq_expand<<<blocks_for(n, 256), 256, 0, D.stream>>>(A);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(D.stream);  // `point` is the caller's buffer
        for (void* p : {(void*)d_point, (void*)d_lo, (void*)d_hi})
            if (p) pool_free(D, p);
        if (e != cudaSuccess) {
            if (d_out) pool_free(D, d_out);
            set_error(std::string("eq_evals failed: ") + cudaGetErrorString(e));
            return COZK_ERR_CUDA;
        }
    }
    PolyEntry E;
    E.dev = device_index;
    E.kind = POLY_MONT;
    E.user_kind = COZK_POLY_PUBLIC;
    E.bits = 0;
    E.d_data = reinterpret_cast<uint8_t*>(d_out);
    E.total = n;
    E.lo = 0;
    E.len = n;
    *out = publish(ctx, E);
    return COZK_OK;
}

// distributed_batch_open_poly_worker (co-noir-spartan/co-spartan/src/worker.rs:745-772) without the network send:
//   agg   = aggregate_poly(eta, polys[0..num_comms])      sum_j eta^j * polys[j]          (co-spartan/src/utils.rs:85-107)
//   (pf, r) = distributed_open(ck, agg, point)            nv folds + nv MSMs               (worker.rs:774-809)
//   evals = polys.map(|p| p.evaluate(point))              every polynomial, not only the first num_comms   (:761-764)
// All three on device 0, from device-resident polynomials: the only traffic is the point in and nv + 1 + k values out.
int cozk_spartan_batch_open_worker(cozk_ctx* ctx, cozk_open_key key, const cozk_srs* level_srs, size_t nv, const cozk_poly* polys,
                                   size_t k, size_t num_comms, const void* point, const void* eta, void* out_proofs,
                                   void* out_val, void* out_evals) {
    if (!ctx || !polys || !point || !eta || !out_proofs || !out_val || !out_evals || k == 0 || num_comms == 0 || num_comms > k) {
        set_error("batch_open_worker: null pointer, no polynomial, or num_comms outside 1..k");  // polys[0..num_comms] panics
        return COZK_ERR_INVALID_ARG;
    }
    if (key) {
        OpenKey K;
        int rc = open_key_lookup(ctx, key, &K);
        if (rc) return rc;
        nv = K.nv;
    }
    if (nv == 0 || nv > 30) {
        set_error("bad nv");
        return COZK_ERR_INVALID_ARG;
    }
    // eta^j, Montgomery (host: num_comms - 1 multiplications)
    std::vector<fr> powers(num_comms);
    powers[0] = fr_one();
    fr e1;
    memcpy(e1.v, eta, 32);
    for (size_t j = 1; j < num_comms; ++j) powers[j] = fr_mul(powers[j - 1], e1);
    cozk_poly agg = 0, chi = 0;
    int rc = cozk_rep3_linear_combination(ctx, polys, powers.data(), num_comms, 0, &agg);
    if (!rc) rc = cozk_pst13_open_poly(ctx, level_srs, nv, key, agg, point, out_proofs, out_val);
    if (!rc) rc = cozk_eq_evals(ctx, 0, point, nv, 0, &chi);
    if (!rc) rc = cozk_rep3_evaluate_at_chi_poly(ctx, polys, k, chi, out_evals);
    if (agg) cozk_poly_release(ctx, agg);
    if (chi) cozk_poly_release(ctx, chi);
    return rc;
}

int cozk_srs_pair_sums(cozk_ctx* ctx, cozk_srs srs, cozk_srs* out) {
    if (!ctx || !out) {
        set_error("null pointer");
        return COZK_ERR_INVALID_ARG;
    }
    size_t n = 0;
    int rc = cozk_srs_len(ctx, srs, &n);
    if (rc) return rc;
    if (n == 0 || (n & 1)) {
        set_error("pair sums need an even, non-zero number of bases");
        return COZK_ERR_INVALID_ARG;
    }
    Device& D = *ctx->devs[0];
    size_t half = n / 2;
    affine* d_out = nullptr;
    uint8_t* d_inf = nullptr;
    {
        std::lock_guard<std::mutex> lock(D.mu);
        COZK_CUDA(cudaSetDevice(D.id));
        cudaError_t e = pool_alloc(D, &d_out, half * sizeof(affine));
        if (e == cudaSuccess) e = pool_alloc(D, &d_inf, half);
        if (e != cudaSuccess) {
            if (d_out) pool_free(D, d_out);
            set_error(std::string("pair sums: allocation failed: ") + cudaGetErrorString(e));
            return COZK_ERR_CUDA;
        }
    }
    rc = srs_pair_sums_into(ctx, srs, d_out, d_inf, &half);
    if (!rc) rc = srs_register_from_device(ctx, 0, d_out, d_inf, half, out, 0);  // openings run on device 0 only
    std::lock_guard<std::mutex> lock(D.mu);
    cudaSetDevice(D.id);
    pool_free(D, d_out);
    pool_free(D, d_inf);
    return rc;
}

int cozk_pst13_open_poly(cozk_ctx* ctx, const cozk_srs* level_srs, size_t nv, cozk_open_key key, cozk_poly poly,
                         const void* point, void* out_proofs, void* out_eval) {
    if (!ctx || !point || !out_proofs || !out_eval || (!level_srs && !key)) {
        set_error("null pointer");
        return COZK_ERR_INVALID_ARG;
    }
    OpenKey K;
    int rc = COZK_OK;
    if (key) {
        rc = open_key_lookup(ctx, key, &K);
        if (rc) return rc;
        nv = K.nv;
        level_srs = K.level_srs.data();
    } else {
        if (nv == 0 || nv > 30) {
            set_error("bad nv");
            return COZK_ERR_INVALID_ARG;
        }
        for (size_t i = 0; i < nv; ++i) {
            size_t len = 0;
            rc = cozk_srs_len(ctx, level_srs[i], &len);
            if (rc) return rc;
            if (len != ((size_t)1 << (nv - i))) {
                set_error("Invalid size of polynomial: SRS level length does not match nv");
                return COZK_ERR_KEY_LENGTH;
            }
        }
    }
    PolyEntry E;
    rc = lookup(ctx, poly, &E);
    if (rc) return rc;
    if (E.dev != 0) {
        set_error("open(): the polynomial must live on device 0 of the context");
        return COZK_ERR_INVALID_ARG;
    }
    size_t n = (size_t)1 << nv;
    if (E.len != n) {
        set_error("Invalid size of polynomial");  // assert_eq!(nv, ck.nv), pst13.rs:438
        return COZK_ERR_KEY_LENGTH;
    }
    OpenSource src;
    src.dev = E.chunk();
    src.stride = E.elem_bytes();
    src.canon = E.kind == POLY_CANON ? 1 : 0;
    return pst13_open_device(ctx, level_srs, key ? &K : nullptr, nv, src, point, out_proofs, out_eval);
}

int cozk_rep3_last_stats(cozk_ctx* ctx, double* out8) {
    if (!ctx || !out8) return COZK_ERR_INVALID_ARG;
    for (int i = 0; i < 8; ++i) out8[i] = ctx->rep3_stats[i];
    return COZK_OK;
}

}  // extern "C"
