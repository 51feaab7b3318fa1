// tcgen05 tensor-core convolution for sm_100a: 1x1 and 3x3 dense convs as (implicit) GEMMs
//   D[M,N] = A[M,K] * W[N,K]^T,  M = B*OH*OW pixels, K = kh*kw*Cin, N = Cout
// with fp32 NHWC storage in HBM and tf32 MMA (fp32 accumulate in TMEM).
//
// One persistent, warp-specialised CTA per SM walks the (m-tile, n-tile) list:
//   warp 0      TMA producer   weights (hi/lo planes) every k-block; the A tile too when the conv is a
//                              plain 1x1 (A is then just the [M,K] row-major activation matrix)
//   warp 1      MMA issuer     one elected thread issues tcgen05.mma kind::tf32 (M=128, N=BN, K=8),
//                              accumulators double-buffered in TMEM (2 x BN columns)
//   warps 2-5   A gather       3x3 / strided / concat / nearest-upsample inputs: each thread owns one GEMM row
//                              and cp.async's its 128-byte k-slice (zero-fill = padding) into the 128B-swizzled
//                              K-major layout the UMMA descriptor expects  (only in the GATHER variant)
//   warps 6-9   converter      3xTF32: split the landed fp32 A tile in place into hi = tf32(a) and
//                              lo = tf32(a - hi) planes (the weights were split on the host), so that
//                              hi*hi + lo*hi + hi*lo recovers fp32-class accuracy from tf32 tensor cores
//   warps 10-13 epilogue       tcgen05.ld TMEM -> registers, folded-BN scale/shift, activation, residual, store
// mbarrier pipelines: full_b (TMA tx) / full_a (cp.async) -> conv_done -> [MMA] -> empty (tcgen05.commit);
// tmem_full (commit) -> [epilogue] -> tmem_empty.
//
// Replaces the TF Conv2D + FusedBatchNorm + LeakyRelu/Relu6 (+ ResizeNearestNeighbor/ConcatV2/Add) ops that
// models/yolonet.py:244-260 and models/keras_mobilenet*.py compose; see DESIGN.md for the rooflines.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "gemm_tc.h"
#include "tc_ptx.cuh"

namespace k2y {

namespace {

constexpr int BM = 128;          // rows per tile (UMMA_M)
constexpr int BK = 32;           // fp32 per k-block = 128 bytes = one SWIZZLE_128B row
constexpr int A_TILE_BYTES = BM * BK * 4;  // 16 KB
constexpr int NUM_THREADS = 14 * 32;
constexpr int MAX_STAGES = 8;

using namespace ptx;

struct TcParams {
    // gather-mode A source (also used for the shape arithmetic in both modes)
    const float *src0, *src1;
    int H, W, C0, C1, up0;          // logical input extent (after upsampling src0)
    int OH, OW, kh, kw, stride, pad_t, pad_l;
    // epilogue
    float *dst;
    const float *residual, *scale, *shift;
    int act;
    float alpha;
    int M, N, K;
    int BN, n_tiles, m_tiles, nkb, stages;
    int cluster;                    // 1, or 2: CTA pairs on adjacent m-tiles share every weight tile through TMA multicast
    int k_splits, kb_per_split;     // split-K: work item = (m-tile, n-tile, k-slice); partials go to a scratch buffer
    int three_x;                    // 1 = split scheme (3 MMAs per k-step: lo*hi + hi*lo + hi*hi), 0 = single pass
    int bf16;                       // 1 = bf16x3: operands are bf16 (hi, mid) planes, 64 k per k-block; 0 = tf32, 32 k per k-block
    int a_boxes;                    // 16 KB sub-tiles of raw fp32 A per stage (k per k-block / 32)
    int half_taps;                  // gather: Cin is half a k-block (16 with tf32, 32 with bf16 k-blocks), a k-block holds TWO consecutive taps
    uint32_t tmem_cols;
    float act_slope, act_clamp;     // branch-free activation parameters
    int epi_groups;                 // 1, or 2: the idle gather warps form a second epilogue group (plain 1x1 convs)
    int tma_store;                  // 1: epilogue writes through a TMA store (N % 4 == 0)
    int res_tma;                    // 1: the residual (Add) operand of a TMA-store epilogue is fetched by TMA into the staging buffer
    // fused depthwise producer (gather variant, nkb == 1): A = act(BN(depthwise3x3(src0))) is computed by the gather warps
    int dw;
    const float *dw_w, *dw_scale, *dw_shift;  // [9][C0], [C0], [C0]
    int dw_act;
    float dw_alpha;
    int dbg;                        // bring-up switches (K2Y_TC_DBG): 1 skip stores, 2 skip tmem loads, 4 skip convert math
    long long *trace;               // optional [gridDim.x][16] globaltimer stamps (K2Y_TC_TRACE=1 via k2y_conv2d)
};

__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// wait + (trace mode) add the cycles spent waiting to trace slot `slot` of this CTA; `who` = this thread is the one that reports
#define K2Y_TIMED_WAIT(bar, parity, slot, who)                                            \
    do {                                                                                 \
        if (p.trace) {                                                                   \
            const long long _t0 = clock64();                                             \
            mbar_wait((bar), (parity));                                                  \
            if (who) p.trace[(size_t)blockIdx.x * 64 + (slot)] += clock64() - _t0;       \
        } else {                                                                         \
            mbar_wait((bar), (parity));                                                  \
        }                                                                                \
    } while (0)
#define K2Y_TRACE(slot)                                                                  \
    do {                                                                                 \
        if (p.trace) p.trace[(size_t)blockIdx.x * 64 + (slot)] = gtime();                \
    } while (0)

struct __align__(8) Barriers {
    uint64_t full_b[MAX_STAGES], full_a[MAX_STAGES], conv[MAX_STAGES], empty[MAX_STAGES];
    uint64_t tmem_full[2], tmem_empty[2];
    uint64_t res_full[8];   // one per epilogue warp: its residual chunk (TMA load into the staging buffer) has landed
    uint32_t tmem_slot;
};
// dynamic smem besides the stage ring: alignment slack, barriers, epilogue staging (2 x 4 KB per epilogue warp)
constexpr size_t FIXED_SMEM = 1024 + sizeof(Barriers) + 1024 + 4 * 8192 + 2 * 4096;  // + scale/shift copies of two epilogue groups

// Depthwise 3x3 + BN + activation for 4 horizontally adjacent outputs x 4 channels (fused producer of the 1x1 convs that
// follow the early MobileNet depthwise layers).  The 3 x (3*STRIDE+3) input window is loaded once (clamped addresses, zero
// for padding taps); tap order per output is (ky, kx) ascending like the stand-alone depthwise kernel.
template <int STRIDE>
__device__ __forceinline__ void dw_quad4(const TcParams &p, int b, int oy, int oxb, int c, float4 (&out)[4]) {
    constexpr int NCOL = 3 * STRIDE + 3;
    const int iy0 = oy * STRIDE - p.pad_t, ixb = oxb * STRIDE - p.pad_l;
    float4 acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = iy0 + ky;
        const bool row_ok = iy >= 0 && iy < p.H;
        const float *rowp = p.src0 + (size_t)(b * p.H + min(max(iy, 0), p.H - 1)) * p.W * p.C0 + c;
        float4 v[NCOL];
#pragma unroll
        for (int j = 0; j < NCOL; ++j) {
            const int ix = ixb + j;
            const bool ok = row_ok && ix >= 0 && ix < p.W;
            const float4 t = __ldg(reinterpret_cast<const float4 *>(rowp + (size_t)min(max(ix, 0), p.W - 1) * p.C0));
            v[j] = ok ? t : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const float4 w = __ldg(reinterpret_cast<const float4 *>(p.dw_w + (ky * 3 + kx) * p.C0 + c));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 x = v[STRIDE * j + kx];
                acc[j].x = fmaf(x.x, w.x, acc[j].x);
                acc[j].y = fmaf(x.y, w.y, acc[j].y);
                acc[j].z = fmaf(x.z, w.z, acc[j].z);
                acc[j].w = fmaf(x.w, w.w, acc[j].w);
            }
        }
    }
    const float4 sc = __ldg(reinterpret_cast<const float4 *>(p.dw_scale + c));
    const float4 sh = __ldg(reinterpret_cast<const float4 *>(p.dw_shift + c));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        out[j].x = apply_act(fmaf(acc[j].x, sc.x, sh.x), p.dw_act, p.dw_alpha);
        out[j].y = apply_act(fmaf(acc[j].y, sc.y, sh.y), p.dw_act, p.dw_alpha);
        out[j].z = apply_act(fmaf(acc[j].z, sc.z, sh.z), p.dw_act, p.dw_alpha);
        out[j].w = apply_act(fmaf(acc[j].w, sc.w, sh.w), p.dw_act, p.dw_alpha);
    }
}

template <bool GATHER, bool DWFUSE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
               const __grid_constant__ CUtensorMap map_blo, const __grid_constant__ CUtensorMap map_out,
               const __grid_constant__ CUtensorMap map_res, const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stages] x { A_hi(raw) 16K | A_lo 16K (3x) | B_hi BN*128 | B_lo BN*128 (3x) }, then barriers
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_bytes = (uint32_t)p.BN * 128u;
    // 3xTF32: the converter writes the hi/lo planes of A into TENSOR MEMORY (the MMA reads A from TMEM), so shared
    // memory only holds the raw fp32 A tile and the two weight planes — shared-memory bandwidth was the limiter when
    // all three MMAs of a k-step streamed A from smem.
    const uint32_t a_bytes = (uint32_t)A_TILE_BYTES * (uint32_t)p.a_boxes;
    const uint32_t stage_bytes = a_bytes + b_bytes * (p.three_x ? 2u : 1u);
    const int KBK = p.a_boxes * BK;  // k values per k-block
    uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    Barriers *bars = reinterpret_cast<Barriers *>(smem_gen + (size_t)p.stages * stage_bytes);

    // Role ids run AGAINST the hardware warp ids: the SM's warp arbiter favours high warp ids, and the two single-thread roles
    // everything else waits for (role 0 = TMA producer, role 1 = MMA issuer) must not queue behind busy epilogue / converter
    // warps of their sub-partition.  `warp` below is the role id (0 producer, 1 MMA, 2-5 gather, 6-9 converter, 10-13
    // epilogue), `pq` the tensor-memory lane quarter the hardware lets this warp touch (physical warp id % 4).
    const int phys_warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp = (NUM_THREADS / 32 - 1) - phys_warp;
    const int pq = phys_warp & 3;
    pdl_trigger();
    const bool use_conv = p.three_x || GATHER;
    const int crank = p.cluster == 2 ? (int)cluster_ctarank() : 0;
    const int cluster_id = (int)blockIdx.x / p.cluster, num_clusters = (int)gridDim.x / p.cluster;

    if (threadIdx.x == 0) {
        K2Y_TRACE(0);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(&bars->full_b[s]), 1);
            mbar_init(smem_u32(&bars->full_a[s]), 128);
            mbar_init(smem_u32(&bars->conv[s]), 128);
            mbar_init(smem_u32(&bars->empty[s]), (uint32_t)p.cluster);  // every MMA of the cluster must retire the stage
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(&bars->tmem_full[a]), 1);
            mbar_init(smem_u32(&bars->tmem_empty[a]), 128u * (uint32_t)p.epi_groups);
        }
        for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars->res_full[i]), 1);
        fence_barrier_init();
        if (p.res_tma) prefetch_tmap(&map_res);
        if (!GATHER) prefetch_tmap(&map_a);
        prefetch_tmap(&map_bhi);
        if (p.three_x) prefetch_tmap(&map_blo);
        if (p.tma_store) prefetch_tmap(&map_out);
    }
    if (warp == 1) tmem_alloc(smem_u32(&bars->tmem_slot), p.tmem_cols);
    tc_fence_before();
    if (p.cluster == 2) cluster_sync_all();  // peer barriers are initialised before any multicast can reach them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_slot;
    if (threadIdx.x == 0) K2Y_TRACE(1);
    pdl_wait();  // barriers / TMEM / descriptor prefetch above overlap the previous kernel's tail; activations do not

    const int mp_tiles = (p.m_tiles + p.cluster - 1) / p.cluster;
    const int num_tiles = mp_tiles * p.n_tiles * p.k_splits;   // work items per cluster sequence
    auto stage_a_hi = [&](int s) { return smem_base + (uint32_t)s * stage_bytes; };
    auto stage_b_hi = [&](int s) { return smem_base + (uint32_t)s * stage_bytes + a_bytes; };
    // tensor-memory columns: [0, 2*BN) two accumulators, then per stage 32 columns of A_hi and 32 of A_lo
    auto tmem_a_hi = [&](int s) { return tmem_base + 2u * (uint32_t)p.BN + (uint32_t)s * 64u; };
    auto stage_b_lo = [&](int s) { return stage_b_hi(s) + b_bytes; };

    // ================= epilogue (a lambda: run by warps 10-13, and by the idle gather warps 2-5 as a second group when the
    // conv is a plain 1x1 — the early pointwise layers are epilogue-bound, two groups split a tile's 32-column chunks) ====
    auto epilogue = [&](const int g, const int ew) {
        // ================= epilogue =================
        // TMEM -> registers (lane = GEMM row) -> BN-fold/activation -> 128B-swizzled smem staging [32 rows][32 cols]
        // per warp -> either one TMA store per 32-column chunk (full-line writes, no LSU work) or, for N % 4 != 0 /
        // residual layers, row-contiguous st.global (a warp instruction covers 4 rows x 128 contiguous bytes).
        const int q = pq;                       // TMEM lane quarter this warp may read
        const bool n_vec = (p.N & 3) == 0;
        // two 4 KB staging buffers per warp, 1024-byte aligned (SWIZZLE_128B atom)
        // staging: 32 KB after the barriers.  One epilogue group: 2 x 4 KB per warp (double buffered against its TMA store);
        // two groups: 4 KB per warp each.  Then one 4 KB scale/shift copy per group (double buffered by tile parity).
        const int ng = p.epi_groups;
        const uint32_t epi_base = (smem_base + (uint32_t)p.stages * stage_bytes + (uint32_t)sizeof(Barriers) + 1023u) & ~1023u;
        const uint32_t stg_base = ng == 1 ? epi_base + (uint32_t)ew * 8192u : epi_base + (uint32_t)g * 16384u + (uint32_t)ew * 4096u;
        const uint32_t ss_base = epi_base + 32768u + (uint32_t)g * 4096u;
        const bool tracer = g == 0 && ew == 0 && lane == 0;
        const int chunk = lane & 7, rsub = lane >> 3;
        uint32_t acc_it = 0, stg_it = 0, res_it = 0;
        const uint32_t res_bar = smem_u32(&bars->res_full[g * 4 + ew]);
        for (int t = cluster_id; t < num_tiles; t += num_clusters, ++acc_it) {
            const int ks = t % p.k_splits, tt = t / p.k_splits;
            const int mp = tt / p.n_tiles, nt = tt - mp * p.n_tiles;
            const int mt = mp * p.cluster + crank;
            const bool tile_ok = mt < p.m_tiles;  // odd m-tile count: the pair's second CTA only helps with the weights
            const uint32_t a = acc_it & 1u, aph = (acc_it >> 1) & 1u;
            // With ~227 KB of shared memory carved out there is practically no L1 left, so per-column scale/shift loads
            // were L2 round trips on the epilogue's critical path: stage the tile's BN columns in shared memory (double
            // buffered by tile parity) while the mainloop of this tile is still running.
            const uint32_t ss = ss_base + (acc_it & 1u) * 2048u;
            {
                const int et = ew * 32 + lane;  // 0..127 inside the group
                for (int j = et; j < ((p.BN + 31) & ~31); j += 128) {
                    const int n = nt * p.BN + j;
                    const bool in = j < p.BN && n < p.N;
                    const float scv = in ? __ldg(p.scale + n) : 0.f, shv = in ? __ldg(p.shift + n) : 0.f;
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(ss + (uint32_t)j * 4u), "f"(scv) : "memory");
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(ss + 1024u + (uint32_t)j * 4u), "f"(shv) : "memory");
                }
                if (g == 0) asm volatile("bar.sync 1, 128;" ::: "memory");  // the four warps of this epilogue group only
                else asm volatile("bar.sync 3, 128;" ::: "memory");
            }
            K2Y_TIMED_WAIT(smem_u32(&bars->tmem_full[a]), aph, 56, (g == 0 && ew == 0 && lane == 0));
            if (t == cluster_id && tracer) K2Y_TRACE(8);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * (uint32_t)p.BN;
            // split-K partials of slice ks live at rows [ks * m_tiles * 128, ...) of the scratch tensor
            const int m_base = (ks * p.m_tiles + mt) * BM + q * 32;
            for (int c0 = g * 32; c0 < p.BN; c0 += 32 * ng, ++stg_it) {
                const int ncols = (p.BN - c0) < 32 ? (p.BN - c0) : 32;  // 32 or 16
                const int n0 = nt * p.BN + c0;
                const uint32_t stg = stg_base + (ng == 1 ? (stg_it & 1u) * 4096u : 0u);
                uint32_t r[32];
                const bool tr = p.trace && t == cluster_id && c0 == 32 && tracer;
                long long c_0 = 0, c_1 = 0, c_2 = 0, c_3 = 0, c_4 = 0;
                if (tr) c_0 = clock64();
                if (p.res_tma && lane == 0) {
                    // residual layers: the [32 rows x 32 columns] residual chunk travels by TMA into this chunk's staging buffer
                    // while the accumulator is read and activated; the buffer's previous TMA store must have read it first
                    if (ng == 1) tma_store_wait_read1();
                    else tma_store_wait_read0();
                    if (tile_ok) {
                        mbar_arrive_expect_tx(res_bar, 4096u);
                        tma_load_2d(stg, &map_res, res_bar, n0, m_base);   // rows >= M / columns >= N arrive as zeros
                    }
                }
                tmem_ld16(taddr + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
                if (ncols == 32) tmem_ld16(taddr + (uint32_t)c0 + 16u, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
                else {
#pragma unroll
                    for (int j = 16; j < 32; ++j) r[j] = 0u;
                }
                tmem_ld_wait();
                if (tr) c_1 = clock64();
                if (p.tma_store) {
                    // scale/shift come from the shared-memory copy (zero beyond N, so no bounds checks); the activation is
                    // specialised outside the element loop: leaky = max(v, slope*v), relu = max(v,0), relu6 = min(max(v,0),6)
#define K2Y_EPI_LOOP(ACT_EXPR)                                                                                     \
    _Pragma("unroll") for (int j = 0; j < 32; j += 4) {                                                             \
        const float4 s4 = ld_shared_v4(ss + (uint32_t)(c0 + j) * 4u);                                               \
        const float4 h4 = ld_shared_v4(ss + 1024u + (uint32_t)(c0 + j) * 4u);                                       \
        float v;                                                                                                     \
        v = fmaf(__uint_as_float(r[j]), s4.x, h4.x);     r[j] = __float_as_uint(ACT_EXPR);                           \
        v = fmaf(__uint_as_float(r[j + 1]), s4.y, h4.y); r[j + 1] = __float_as_uint(ACT_EXPR);                       \
        v = fmaf(__uint_as_float(r[j + 2]), s4.z, h4.z); r[j + 2] = __float_as_uint(ACT_EXPR);                       \
        v = fmaf(__uint_as_float(r[j + 3]), s4.w, h4.w); r[j + 3] = __float_as_uint(ACT_EXPR);                       \
    }
                    if (p.act == ACT_LEAKY) {
                        const float slope = p.act_slope;
                        K2Y_EPI_LOOP(fmaxf(v, v * slope))
                    } else if (p.act == ACT_RELU) {
                        K2Y_EPI_LOOP(fmaxf(v, 0.f))
                    } else if (p.act == ACT_RELU6) {
                        K2Y_EPI_LOOP(fminf(fmaxf(v, 0.f), 6.f))
                    } else {
                        K2Y_EPI_LOOP(v)
                    }
#undef K2Y_EPI_LOOP
                    if (tr) c_2 = clock64();
                    if (p.res_tma) {
                        if (tile_ok) {
                            mbar_wait(res_bar, res_it & 1u);
                            ++res_it;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {   // this lane's row of the residual chunk, same swizzle as the store below
                                const float4 rr = ld_shared_v4(stg + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4));
                                r[j * 4] = __float_as_uint(__uint_as_float(r[j * 4]) + rr.x);
                                r[j * 4 + 1] = __float_as_uint(__uint_as_float(r[j * 4 + 1]) + rr.y);
                                r[j * 4 + 2] = __float_as_uint(__uint_as_float(r[j * 4 + 2]) + rr.z);
                                r[j * 4 + 3] = __float_as_uint(__uint_as_float(r[j * 4 + 3]) + rr.w);
                            }
                        }
                    } else if (lane == 0) {
                        // the staging buffer used two chunks ago must have been read by its TMA store
                        if (ng == 1) tma_store_wait_read1();
                        else tma_store_wait_read0();
                    }
                    __syncwarp();
                    if (tr) c_3 = clock64();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        st_shared_v4(stg + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4), r[j * 4], r[j * 4 + 1],
                                     r[j * 4 + 2], r[j * 4 + 3]);
                    fence_proxy_async();
                    __syncwarp();
                    if (tr) c_4 = clock64();
                    if (lane == 0 && tile_ok && !(p.dbg & 1)) {
                        tma_store_2d(&map_out, stg, n0, m_base);  // rows >= M and columns >= N are clipped by the TMA
                        tma_store_commit();
                    }
                    if (tr) {
                        long long *o = p.trace + (size_t)blockIdx.x * 64 + 48;
                        o[0] = c_1 - c_0;
                        o[1] = c_2 - c_1;
                        o[2] = c_3 - c_2;
                        o[3] = c_4 - c_3;
                        o[4] = clock64() - c_4;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        st_shared_v4(stg + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4), r[j * 4], r[j * 4 + 1],
                                     r[j * 4 + 2], r[j * 4 + 3]);
                    __syncwarp();
                    const int n = n0 + chunk * 4;
                    const bool col_ok = (chunk * 4 < ncols) && (n < p.N);
                    const bool full4 = n_vec && (n + 3 < p.N);
                    float sc[4] = {0.f, 0.f, 0.f, 0.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
                    if (col_ok) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (n + j < p.N) {
                                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(sc[j]) : "r"(ss + (uint32_t)(c0 + chunk * 4 + j) * 4u));
                                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(sh[j]) : "r"(ss + 1024u + (uint32_t)(c0 + chunk * 4 + j) * 4u));
                            }
                    }
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int row = it * 4 + rsub;
                        const int m = m_base + row;
                        const float4 v = ld_shared_v4(stg + (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4));
                        if (col_ok && tile_ok && m < p.M && !(p.dbg & 1)) {
                            float o[4];
                            o[0] = act_bf(fmaf(v.x, sc[0], sh[0]), p.act_slope, p.act_clamp);
                            o[1] = act_bf(fmaf(v.y, sc[1], sh[1]), p.act_slope, p.act_clamp);
                            o[2] = act_bf(fmaf(v.z, sc[2], sh[2]), p.act_slope, p.act_clamp);
                            o[3] = act_bf(fmaf(v.w, sc[3], sh[3]), p.act_slope, p.act_clamp);
                            float *out = p.dst + (size_t)m * p.N + n;
                            if (full4) {
                                if (p.residual) {
                                    const float4 rr = __ldg(reinterpret_cast<const float4 *>(p.residual + (size_t)m * p.N + n));
                                    o[0] += rr.x, o[1] += rr.y, o[2] += rr.z, o[3] += rr.w;
                                }
                                *reinterpret_cast<float4 *>(out) = make_float4(o[0], o[1], o[2], o[3]);
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (n + j < p.N) out[j] = o[j] + (p.residual ? __ldg(p.residual + (size_t)m * p.N + n + j) : 0.f);
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&bars->tmem_empty[a]));
            if (t == cluster_id && tracer) K2Y_TRACE(9);
        }
        if (p.tma_store && lane == 0) tma_store_wait_all();  // global writes complete before the CTA retires
        if (tracer) K2Y_TRACE(10);
    };

    if (warp == 0) {
        // ================= TMA producer (whole warp walks the loop, one elected lane issues) =================
        {
            int s = 0;
            uint32_t ph = 0;
            const uint32_t tx = b_bytes * (p.three_x ? 2u : 1u) + (GATHER ? 0u : a_bytes);
            const int bk_elems = p.bf16 ? 64 : 32;  // k-block extent in elements of the weight map (bf16 or fp32)
            for (int t = cluster_id; t < num_tiles; t += num_clusters) {
                const int ks = t % p.k_splits, tt = t / p.k_splits;
                const int mp = tt / p.n_tiles, nt = tt - mp * p.n_tiles;
                const int mt = mp * p.cluster + crank;
                const int kb0 = ks * p.kb_per_split, kb1 = min(p.nkb, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    K2Y_TIMED_WAIT(smem_u32(&bars->empty[s]), ph ^ 1u, 57, (warp == 0 && lane == 0));
                    const uint32_t fb = smem_u32(&bars->full_b[s]);
                    if (elect_one_sync()) {
                    mbar_arrive_expect_tx(fb, tx);
                    if (!GATHER) {
                        for (int bx = 0; bx < p.a_boxes; ++bx)
                            tma_load_2d(stage_a_hi(s) + (uint32_t)bx * A_TILE_BYTES, &map_a, fb, kb * KBK + bx * BK, mt * BM);
                    }
                    if (p.cluster == 2) {
                        // this CTA fetches its half of the weight tile and multicasts it into both CTAs of the pair
                        const int half = p.BN >> 1;
                        const uint32_t off = (uint32_t)(crank * half) * 128u;
                        tma_load_2d_mc(stage_b_hi(s) + off, &map_bhi, fb, kb * bk_elems, nt * p.BN + crank * half, (uint16_t)3);
                        if (p.three_x) tma_load_2d_mc(stage_b_lo(s) + off, &map_blo, fb, kb * bk_elems, nt * p.BN + crank * half, (uint16_t)3);
                    } else {
                        tma_load_2d(stage_b_hi(s), &map_bhi, fb, kb * bk_elems, nt * p.BN);
                        if (p.three_x) tma_load_2d(stage_b_lo(s), &map_blo, fb, kb * bk_elems, nt * p.BN);
                    }
                    if (t == cluster_id && kb - kb0 < 8) K2Y_TRACE(16 + (kb - kb0) * 4);
                    if (t == cluster_id && kb == kb0) K2Y_TRACE(2);
                    if (t == cluster_id && kb == kb1 - 1) K2Y_TRACE(3);
                    }
                    __syncwarp();
                    if (++s == p.stages) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (whole warp walks the loop, one elected lane issues) =================
        {
            const uint32_t idesc = p.bf16 ? make_idesc_bf16(BM, p.BN) : make_idesc_tf32(BM, p.BN);
            int s = 0;
            uint32_t ph = 0, acc_it = 0;
            for (int t = cluster_id; t < num_tiles; t += num_clusters, ++acc_it) {
                const uint32_t a = acc_it & 1u, aph = (acc_it >> 1) & 1u;
                K2Y_TIMED_WAIT(smem_u32(&bars->tmem_empty[a]), aph ^ 1u, 58, (lane == 0));
                tc_fence_after();
                const uint32_t d = tmem_base + a * (uint32_t)p.BN;
                const int ks = t % p.k_splits;
                const int kb0 = ks * p.kb_per_split, kb1 = min(p.nkb, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (!(p.dbg & 8)) K2Y_TIMED_WAIT(smem_u32(&bars->full_b[s]), ph, 59, (lane == 0));
                    if (t == cluster_id && kb == kb0 && lane == 0) K2Y_TRACE(4);
                    if (use_conv && !(p.dbg & 8)) K2Y_TIMED_WAIT(smem_u32(&bars->conv[s]), ph, 60, (lane == 0));
                    if (t == cluster_id && kb == kb0 && lane == 0) K2Y_TRACE(5);
                    tc_fence_after();
                    if (elect_one_sync()) {
                    const uint64_t a_hi = make_desc_sw128(stage_a_hi(s)), b_hi = make_desc_sw128(stage_b_hi(s));
                    const uint64_t b_lo = make_desc_sw128(stage_b_lo(s));
                    const uint32_t ta_hi = tmem_a_hi(s), ta_lo = ta_hi + 32u;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {  // 4 k-steps per k-block: 8 tf32 or 16 bf16 values = 32 bytes of B each
                        if (p.dbg & 16) break;
                        const uint64_t adv = (uint64_t)((kk * 32) >> 4);  // advance inside the 128 B swizzle row
                        const uint32_t acol = (uint32_t)(kk * 8);          // 8 TMEM columns of A per k-step
                        if (p.bf16) {
                            // bf16x3: mid*hi + hi*mid + hi*hi (A planes in tensor memory, 2 bf16 per column)
                            umma_bf16_ts(d, ta_lo + acol, b_hi + adv, idesc, ((kb - kb0) | kk) != 0);
                            umma_bf16_ts(d, ta_hi + acol, b_lo + adv, idesc, 1u);
                            umma_bf16_ts(d, ta_hi + acol, b_hi + adv, idesc, 1u);
                        } else if (p.three_x) {
                            // 3xTF32: small terms first, then the dominant hi*hi; A comes from tensor memory
                            umma_tf32_ts(d, ta_lo + acol, b_hi + adv, idesc, ((kb - kb0) | kk) != 0);
                            umma_tf32_ts(d, ta_hi + acol, b_lo + adv, idesc, 1u);
                            umma_tf32_ts(d, ta_hi + acol, b_hi + adv, idesc, 1u);
                        } else {
                            umma_tf32(d, a_hi + adv, b_hi + adv, idesc, ((kb - kb0) | kk) != 0);
                        }
                    }
                    if (t == cluster_id && kb - kb0 < 8) K2Y_TRACE(16 + (kb - kb0) * 4 + 3);
                    if (p.cluster == 2) umma_commit_mc(smem_u32(&bars->empty[s]), (uint16_t)3);  // frees the slot in BOTH CTAs
                    else umma_commit(smem_u32(&bars->empty[s]));  // frees the smem slot when these MMAs retire
                    }
                    __syncwarp();
                    if (++s == p.stages) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
                if (elect_one_sync()) {
                    umma_commit(smem_u32(&bars->tmem_full[a]));
                    if (t == cluster_id) K2Y_TRACE(6);
                }
                __syncwarp();
            }
            if (lane == 0) K2Y_TRACE(7);
        }
    } else if (warp >= 2 && warp < 6) {
        // ================= A gather (implicit GEMM rows) =================
        // Each k-block is KBK consecutive channels of one filter tap, i.e. one contiguous run per output pixel.  Thread r
        // computes the source address of GEMM row r once per k-block; the addresses are then exchanged with warp shuffles
        // so that 8 (tf32) or 16 (bf16) consecutive lanes copy one row's run: a cp.async warp instruction touches 4 or 2
        // full 128-byte lines instead of 32 different ones (the L1/LSU request rate was the limiter of the 3x3 convs).
        if (!GATHER && p.epi_groups == 2) epilogue(1, warp - 2);
        if (GATHER && DWFUSE) {
            // ---- fused depthwise producer: the A tile is computed, not copied (one k-block per tile, K = C0 <= 64) ----
            // work item = (channel quad, 4 consecutive GEMM rows = 4 horizontally adjacent output pixels; OW % 4 == 0);
            // lanes run over the channel quads first, so a load instruction reads whole pixels' channel runs.
            const int g = (warp - 2) * 32 + lane;
            const int c4n = p.C0 >> 2;
            for (int s2 = 0; s2 < p.stages; ++s2)  // columns >= C0 stay zero for the whole kernel (weights there are zero too)
                for (uint32_t i = (uint32_t)g; i < a_bytes / 16u; i += 128u) st_shared_v4(stage_a_hi(s2) + i * 16u, 0u, 0u, 0u, 0u);
            asm volatile("bar.sync 1, 128;" ::: "memory");
            int s = 0;
            uint32_t ph = 0;
            for (int t = cluster_id; t < num_tiles; t += num_clusters) {
                const int mt = (t / p.n_tiles) * p.cluster + crank;
                const long long tw0 = p.trace ? clock64() : 0;
                mbar_wait(smem_u32(&bars->empty[s]), ph ^ 1u);
                const long long tw1 = p.trace ? clock64() : 0;
                if (mt < p.m_tiles) {
                    const uint32_t stg_a = stage_a_hi(s);
                    for (int it = g; it < c4n * (BM / 4); it += 128) {
                        const int c = (it % c4n) * 4;
                        const int seg = it / c4n;
                        const int m0 = mt * BM + seg * 4;
                        if (m0 >= p.M) continue;  // rows past the end keep older finite data; they are never stored
                        const int oxb = m0 % p.OW;
                        const int q = m0 / p.OW;
                        const int oy = q % p.OH, b = q / p.OH;
                        float4 o[4];
                        if (p.stride == 1) dw_quad4<1>(p, b, oy, oxb, c, o);
                        else dw_quad4<2>(p, b, oy, oxb, c, o);
                        const uint32_t box = (uint32_t)(c >> 5) * A_TILE_BYTES;
                        const uint32_t chunk = (uint32_t)((c & 31) >> 2);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t row = (uint32_t)(seg * 4 + j);
                            st_shared_v4(stg_a + box + row * 128u + ((chunk ^ (row & 7u)) << 4), __float_as_uint(o[j].x),
                                         __float_as_uint(o[j].y), __float_as_uint(o[j].z), __float_as_uint(o[j].w));
                        }
                    }
                }
                mbar_arrive(smem_u32(&bars->full_a[s]));
                if (p.trace && g == 0) {  // producer timeline of this CTA: cycles waiting for a free stage / producing / tiles
                    p.trace[(size_t)blockIdx.x * 64 + 53] += tw1 - tw0;
                    p.trace[(size_t)blockIdx.x * 64 + 54] += clock64() - tw1;
                    p.trace[(size_t)blockIdx.x * 64 + 55] += 1;
                }
                if (++s == p.stages) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        } else if (GATHER) {
            const int r = (warp - 2) * 32 + lane;  // GEMM row inside the tile, 0..127
            const int Cin = p.C0 + p.C1;
            const int wrow0 = (warp - 2) * 32;   // first tile row of this warp
            int s = 0;
            uint32_t ph = 0;
            for (int t = cluster_id; t < num_tiles; t += num_clusters) {
                const int ks = t % p.k_splits, tt = t / p.k_splits;
                const int mt = (tt / p.n_tiles) * p.cluster + crank;
                const int kb0 = ks * p.kb_per_split, kb1 = min(p.nkb, kb0 + p.kb_per_split);
                const int m = mt * BM + r;
                int b = -1, iy0 = 0, ix0 = 0;
                if (m < p.M && mt < p.m_tiles) {
                    const int ox = m % p.OW;
                    const int q = m / p.OW;
                    const int oy = q % p.OH;
                    b = q / p.OH;
                    iy0 = oy * p.stride - p.pad_t;
                    ix0 = ox * p.stride - p.pad_l;
                }
                for (int kb = kb0; kb < kb1; ++kb) {
                    // a k-block = KBK consecutive channels of ONE tap of ONE source (Cin % KBK == 0, C0 % KBK == 0)
                    const int k = kb * KBK;
                    const int tap = k / Cin;
                    const int ci = k - tap * Cin;
                    const int ky = tap / p.kw, kx = tap - ky * p.kw;
                    const int iy = iy0 + ky, ix = ix0 + kx;
                    const bool ok = b >= 0 && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
                    unsigned long long pr = 0ull;  // 0 = zero-fill (padding / rows beyond M)
                    unsigned long long pr2 = 0ull; // half_taps: the second tap of the k-block (upper half of its channels)
                    if (p.half_taps) {
                        // k-block kb = taps 2 kb and 2 kb + 1, C0 channels each; tap kh*kw (K padding) stays zero
                        const int t0 = 2 * kb;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int tp = t0 + h;
                            const int ky2 = tp / p.kw, kx2 = tp - ky2 * p.kw;
                            const int iy2 = iy0 + ky2, ix2 = ix0 + kx2;
                            const bool ok2 = b >= 0 && tp < p.kh * p.kw && iy2 >= 0 && iy2 < p.H && ix2 >= 0 && ix2 < p.W;
                            const unsigned long long a2 =
                                ok2 ? (unsigned long long)(p.src0 + ((size_t)(b * p.H + iy2) * p.W + ix2) * p.C0) : 0ull;
                            if (h == 0) pr = a2;
                            else pr2 = a2;
                        }
                    } else if (ok) {
                        const float *src;
                        if (ci < p.C0) {
                            if (p.up0)
                                src = p.src0 + ((size_t)(b * (p.H >> 1) + (iy >> 1)) * (p.W >> 1) + (ix >> 1)) * p.C0 + ci;
                            else
                                src = p.src0 + ((size_t)(b * p.H + iy) * p.W + ix) * p.C0 + ci;
                        } else {
                            src = p.src1 + ((size_t)(b * p.H + iy) * p.W + ix) * p.C1 + (ci - p.C0);
                        }
                        pr = (unsigned long long)src;
                    }
                    mbar_wait(smem_u32(&bars->empty[s]), ph ^ 1u);
                    const uint32_t stg_a = stage_a_hi(s);
                    if (p.a_boxes == 2) {
                        const int cjj = lane & 15;
                        const uint32_t dcol = (uint32_t)(cjj >> 3) * A_TILE_BYTES;
                        if (p.half_taps) {   // sub-tile 0 = first tap's 32 channels, sub-tile 1 = second tap's
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const int lr = 2 * i + (lane >> 4);
                                const unsigned long long q0 = __shfl_sync(0xffffffffu, pr, lr);
                                const unsigned long long q1 = __shfl_sync(0xffffffffu, pr2, lr);
                                const unsigned long long q = (cjj & 8) ? q1 : q0;
                                const int row = wrow0 + lr;
                                cp_async_16(stg_a + dcol + (uint32_t)row * 128u + (uint32_t)(((cjj & 7) ^ (row & 7)) << 4),
                                            q ? (const void *)(q + (unsigned long long)(cjj & 7) * 16ull) : (const void *)p.src0, q ? 16u : 0u);
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const int lr = 2 * i + (lane >> 4);
                                const unsigned long long q = __shfl_sync(0xffffffffu, pr, lr);
                                const int row = wrow0 + lr;
                                cp_async_16(stg_a + dcol + (uint32_t)row * 128u + (uint32_t)(((cjj & 7) ^ (row & 7)) << 4),
                                            q ? (const void *)(q + (unsigned long long)cjj * 16ull) : (const void *)p.src0, q ? 16u : 0u);
                            }
                        }
                    } else {
                        const int cjj = lane & 7;
                        if (p.half_taps) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int lr = 4 * i + (lane >> 3);
                                const unsigned long long q0 = __shfl_sync(0xffffffffu, pr, lr);
                                const unsigned long long q1 = __shfl_sync(0xffffffffu, pr2, lr);
                                const unsigned long long q = (cjj & 4) ? q1 : q0;   // chunks 0-3: first tap, 4-7: second tap
                                const int row = wrow0 + lr;
                                cp_async_16(stg_a + (uint32_t)row * 128u + (uint32_t)((cjj ^ (row & 7)) << 4),
                                            q ? (const void *)(q + (unsigned long long)(cjj & 3) * 16ull) : (const void *)p.src0, q ? 16u : 0u);
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int lr = 4 * i + (lane >> 3);
                                const unsigned long long q = __shfl_sync(0xffffffffu, pr, lr);
                                const int row = wrow0 + lr;
                                cp_async_16(stg_a + (uint32_t)row * 128u + (uint32_t)((cjj ^ (row & 7)) << 4),
                                            q ? (const void *)(q + (unsigned long long)cjj * 16ull) : (const void *)p.src0, q ? 16u : 0u);
                            }
                        }
                    }
                    cp_async_mbar_arrive_noinc(smem_u32(&bars->full_a[s]));
                    if (++s == p.stages) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else if (warp >= 6 && warp < 10) {
        // ================= converter: fp32 -> (hi, lo) tf32 planes =================
        if (use_conv) {
            const int ct = (warp - 6) * 32 + lane;  // 0..127
            int s = 0;
            uint32_t ph = 0;
            for (int t = cluster_id; t < num_tiles; t += num_clusters) {
                const int ks = t % p.k_splits;
                const int kb0 = ks * p.kb_per_split, kb1 = min(p.nkb, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    K2Y_TIMED_WAIT(smem_u32(GATHER ? &bars->full_a[s] : &bars->full_b[s]), ph, 61, (ct == 0));
                    if (t == cluster_id && kb - kb0 < 8 && ct == 0) K2Y_TRACE(16 + (kb - kb0) * 4 + 1);
                    if (p.bf16) {
                        // thread = GEMM row: 64 fp32 -> bf16 hi plane (32 columns) + bf16 mid plane (32 columns) in TMEM
                        const int row = pq * 32 + lane;
                        uint32_t hi[32], lo[32];
#pragma unroll
                        for (int bx = 0; bx < 2; ++bx) {
                            const uint32_t src = stage_a_hi(s) + (uint32_t)bx * A_TILE_BYTES + (uint32_t)row * 128u;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 v = ld_shared_v4(src + (uint32_t)((j ^ (row & 7)) << 4));
                                const uint32_t h01 = pack_bf16x2(v.x, v.y), h23 = pack_bf16x2(v.z, v.w);
                                const float r0 = v.x - __uint_as_float(h01 << 16), r1 = v.y - __uint_as_float(h01 & 0xffff0000u);
                                const float r2 = v.z - __uint_as_float(h23 << 16), r3 = v.w - __uint_as_float(h23 & 0xffff0000u);
                                hi[bx * 16 + j * 2] = h01;
                                hi[bx * 16 + j * 2 + 1] = h23;
                                lo[bx * 16 + j * 2] = pack_bf16x2(r0, r1);
                                lo[bx * 16 + j * 2 + 1] = pack_bf16x2(r2, r3);
                            }
                        }
                        const uint32_t trow = tmem_a_hi(s) + ((uint32_t)(pq * 32) << 16);
                        tmem_st32(trow, hi);
                        tmem_st32(trow + 32u, lo);
                        tmem_st_wait();
                        tc_fence_before();
                    } else if (p.three_x) {
                        // thread = GEMM row: read its 128-byte (swizzled) k-slice, split, write both planes to TMEM
                        const int row = pq * 32 + lane;
                        const uint32_t src = stage_a_hi(s) + (uint32_t)row * 128u;
                        uint32_t hi[32], lo[32];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 v = ld_shared_v4(src + (uint32_t)((j ^ (row & 7)) << 4));
                            const float h0 = to_tf32_rna(v.x), h1 = to_tf32_rna(v.y), h2 = to_tf32_rna(v.z), h3 = to_tf32_rna(v.w);
                            hi[j * 4] = __float_as_uint(h0);
                            hi[j * 4 + 1] = __float_as_uint(h1);
                            hi[j * 4 + 2] = __float_as_uint(h2);
                            hi[j * 4 + 3] = __float_as_uint(h3);
                            lo[j * 4] = __float_as_uint(to_tf32_rna(v.x - h0));
                            lo[j * 4 + 1] = __float_as_uint(to_tf32_rna(v.y - h1));
                            lo[j * 4 + 2] = __float_as_uint(to_tf32_rna(v.z - h2));
                            lo[j * 4 + 3] = __float_as_uint(to_tf32_rna(v.w - h3));
                        }
                        const uint32_t trow = tmem_a_hi(s) + ((uint32_t)(pq * 32) << 16);
                        tmem_st32(trow, hi);
                        tmem_st32(trow + 32u, lo);
                        tmem_st_wait();
                        tc_fence_before();
                    } else {
                        fence_proxy_async();  // cp.async-written A tile -> visible to the MMA (async proxy)
                    }
                    mbar_arrive(smem_u32(&bars->conv[s]));
                    if (t == cluster_id && kb - kb0 < 8 && ct == 0) K2Y_TRACE(16 + (kb - kb0) * 4 + 2);
                    if (++s == p.stages) {
                        s = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else {
        epilogue(0, warp - 10);
    }

    tc_fence_before();
    if (p.cluster == 2) cluster_sync_all();  // the peer may still multicast into / arrive on this CTA's shared memory
    else __syncthreads();
    if (threadIdx.x == 0) K2Y_TRACE(11);
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
        if (lane == 0) K2Y_TRACE(12);
    }
}

// Split-K reduction + epilogue: out[m,n] = act(scale[n] * sum_s partial[s][m][n] + shift[n]) (+ residual).
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float *__restrict__ part, float *__restrict__ dst,
                                                            const float *__restrict__ residual, const float *__restrict__ scale,
                                                            const float *__restrict__ shift, int M, int N, int splits,
                                                            size_t split_stride, float slope, float clamp) {
    pdl_trigger();
    pdl_wait();
    const int n4 = N >> 2;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)M * n4) return;
    const int n = (int)(idx % n4) * 4;
    const size_t m = idx / n4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(part + (size_t)s * split_stride + m * N + n));
        acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale + n));
    const float4 sh = __ldg(reinterpret_cast<const float4 *>(shift + n));
    float4 o;
    o.x = act_bf(fmaf(acc.x, sc.x, sh.x), slope, clamp);
    o.y = act_bf(fmaf(acc.y, sc.y, sh.y), slope, clamp);
    o.z = act_bf(fmaf(acc.z, sc.z, sh.z), slope, clamp);
    o.w = act_bf(fmaf(acc.w, sc.w, sh.w), slope, clamp);
    if (residual) {
        const float4 r = __ldg(reinterpret_cast<const float4 *>(residual + m * N + n));
        o.x += r.x, o.y += r.y, o.z += r.z, o.w += r.w;
    }
    *reinterpret_cast<float4 *>(dst + m * N + n) = o;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2-D row-major [rows][cols] tensor (fp32 or bf16), box = [box_rows][128 bytes of columns], 128B swizzle, OOB -> 0.
bool make_map_2d(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint64_t pitch_elems = 0,
                 bool bf16 = false) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    const size_t esz = bf16 ? 2 : 4;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {(pitch_elems ? pitch_elems : cols) * esz};
    cuuint32_t box[2] = {(cuuint32_t)(128 / esz), box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

float __int_as_float_host(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}

uint16_t bf16_rne_host(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}
float bf16_to_float_host(uint16_t b) {
    const uint32_t u = (uint32_t)b << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

float tf32_rna_host(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return x;
    u = (u + 0x1000u) & 0xffffe000u;
    float y;
    memcpy(&y, &u, 4);
    return y;
}



// Tile width: UMMA_N is a multiple of 16 and <= 256 (two accumulator stages must fit the 512 TMEM columns).
// Big-M layers take the whole N in one tile (A streams from HBM exactly once); small-M layers (7x10 / 14x20
// grids) split N so that the tile count approaches the SM count.
int g_num_sms = 0;        // SM count / opt-in shared memory of the first device initialised (the planner's view of "a B200")
size_t g_max_smem = 0;
constexpr int ID_LEN = 4096;
constexpr size_t DEFAULT_SCRATCH_BYTES = (size_t)64 << 20;
// Per-device state: function attributes (opt-in shared memory) are per device, and so are the allocations.  A net brings
// its own split-K scratch (ConvArgs::tc_scratch), so nets on different streams never share partials; the per-device
// default scratch below only serves the single-layer test hook (k2y_conv2d).
struct DevState {
    bool init = false;
    float *scratch = nullptr;      // split-K partials [splits][m_tiles*128][N]
    size_t scratch_bytes = 0;
    float *ones = nullptr, *zeros = nullptr;  // identity scale / shift for the partial pass
};
constexpr int MAX_DEVICES = 64;
DevState g_dev[MAX_DEVICES];
DevState *cur_dev() {
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= MAX_DEVICES) return nullptr;
    return &g_dev[d];
}

// Tile width BN and split-K factor from a small cost model (SM cycles, calibrated on the device-side timelines
// in profiles/): a work item costs  kb * max(MMA time, pipeline latency / depth) + epilogue + fixed,  the layer costs
// ceil(items / SMs) item times (+ the reduction pass when K is split).  A is re-read once per n-tile, so ties go to
// the wider tile; BN is a multiple of 32 when N is tiled (TMA-store boxes are 32 columns wide).
// CTA pairs (cluster of 2 on adjacent m-tiles) fetch each weight tile once and multicast it: halves the L2 -> SM
// weight traffic that bounds the K-deep layers.  Not worth the cluster sync for single-tile or 1-2 k-block layers.
int pick_cluster(int M, int nkb) {
    const char *force = getenv("K2Y_TC_CLUSTER");
    if (force) return atoi(force) == 2 ? 2 : 1;
    return ((M + BM - 1) / BM >= 2 && nkb >= 4) ? 2 : 1;
}

// SMs a conv may plan for: the device's, or the caller's budget (even, >= 8) when it asks for less
int budget_sms(int sm_limit) {
    const int all = g_num_sms > 0 ? g_num_sms : 148;
    if (sm_limit <= 0 || sm_limit >= all) return all;
    return (sm_limit < 8 ? 8 : sm_limit) & ~1;
}

void pick_tile(int M, int N, int nkb, bool three_x, bool bf16, bool gather, int cluster, size_t scratch_limit, int sms, int *bn_out,
               int *splits_out) {
    if (scratch_limit == 0) scratch_limit = DEFAULT_SCRATCH_BYTES;
    const int n16 = (N + 15) / 16 * 16;
    const int m_tiles = (M + BM - 1) / BM;
    const char *force = getenv("K2Y_TC_SPLITK");
    int best = 16, best_s = 1;
    double best_cost = 1e30;
    const char *force_bn = getenv("K2Y_TC_BN");
    for (int bn = 16; bn <= 256; bn += 16) {
        if (bn > n16) break;
        if (force_bn && bn != atoi(force_bn)) continue;
        const size_t a_bytes = (size_t)A_TILE_BYTES * (bf16 ? 2 : 1);
        const size_t stage = a_bytes + (size_t)bn * 128 * (three_x ? 2 : 1);
        int stages = (int)(((g_max_smem ? g_max_smem : 232448) - FIXED_SMEM) / stage);
        if (three_x) stages = stages < (512 - 2 * bn) / 64 ? stages : (512 - 2 * bn) / 64;  // A planes live in TMEM: 64 columns / stage
        if (stages > MAX_STAGES) stages = MAX_STAGES;
        if (stages < 3 && bn > 16 && !force_bn && !bf16) continue;  // tf32 modes: not re-measured with 2-stage tiles
        if (stages < 2) continue;
        const int n_tiles = (n16 + bn - 1) / bn;
        if (n_tiles > 1 && (bn % 32) != 0) continue;
        // per k-block (32 k for tf32, 64 k for bf16): tensor time (measured: a tf32 MMA of 128 x bn x 8 takes ~bn cycles, a bf16
        // MMA of 128 x bn x 16 ~bn/2), shared-memory traffic (TMA fill + converter read + MMA operand reads) at 128 B/clk,
        // and the load->convert->mma->free round trip divided by the pipeline depth
        // calibrated on the per-k-block timelines (profiles/): issue-bound MMA floor, converter pass, gather issue, and the
        // load -> convert -> mma -> free round trip (~3000 cycles, ~5000 when A is gathered) divided by the pipeline depth
        const double passes = three_x ? 3.0 : 1.0;
        double mma = passes * 4.0 * (bf16 ? bn / 2.0 : (double)bn);
        if (mma < passes * 4.0 * 30.0) mma = passes * 4.0 * 30.0;
        const double conv = three_x ? (bf16 ? 700.0 : 450.0) : 0.0;
        const double smem_bytes = (double)a_bytes + bn * 128.0 * (three_x ? 2 : 1) + (three_x ? (double)a_bytes : 4.0 * 4096.0) +
                                  passes * 4.0 * bn * 32.0;
        const double smem_t = smem_bytes / 128.0;
        // measured per-k-block times (bf16x3, profiles/r01_tile_model.md): gather 64/3 stages 1310 cycles, 96/3 1370, 128/2 1615,
        // 192/2 1840; plain 96/3 1000, 192/2 1900  ->  round trip ~ L0 + 5.5 * bn, shared by the stages in flight
        const double lat = bf16 ? ((gather ? (stages >= 3 ? 3300.0 : 2600.0) : 2500.0) + 5.5 * bn) / stages
                                : (gather ? 5000.0 : 3000.0) / stages;
        double kbt = mma > lat ? mma : lat;
        if (conv > kbt) kbt = conv;
        if (smem_t > kbt) kbt = smem_t;
        const double waste = (double)(n_tiles * bn) / n16;                   // zero-padded columns still cost MMA time
        for (int sp = 1; sp <= 8; ++sp) {
            if (force && sp != atoi(force) && !(atoi(force) < 1 && sp == 1)) continue;
            if (sp > 1) {
                if ((N & 3) != 0 || N > ID_LEN || nkb / sp < 12) continue;
                if ((size_t)sp * m_tiles * BM * N * sizeof(float) > scratch_limit) continue;
            }
            const int kb = (nkb + sp - 1) / sp;
            const long items = (long)((m_tiles + cluster - 1) / cluster) * n_tiles * sp;  // per cluster
            const int units = sms / cluster;
            const double item = kb * kbt + bn * 35.0 + 2500.0;  // epilogue ~1100 cycles per 32 columns
            double cost = (double)((items + units - 1) / units) * item * (0.9 + 0.1 * waste) + n_tiles * 40.0;
            if (sp > 1) cost += 7000.0 + (double)sp * M * N * 4.0 / 1500.0;  // reduce pass: launch + partial traffic
            if (cost < best_cost) {
                best_cost = cost;
                best = bn;
                best_s = sp;
            }
        }
    }
    *bn_out = best;
    *splits_out = best_s;
}

int tc_init() {
    int dev = 0;
    K2Y_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= MAX_DEVICES) {
        set_error("tc_init: device ordinal %d out of range", dev);
        return K2Y_ERR_INVALID;
    }
    DevState &ds = g_dev[dev];
    if (ds.init) return K2Y_OK;
    cudaDeviceProp prop;
    K2Y_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    K2Y_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)prop.sharedMemPerBlockOptin));
    K2Y_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)prop.sharedMemPerBlockOptin));
    K2Y_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)prop.sharedMemPerBlockOptin));
    if (g_num_sms == 0) {
        g_max_smem = prop.sharedMemPerBlockOptin;
        g_num_sms = prop.multiProcessorCount;
    }
    K2Y_CUDA_CHECK(cudaMalloc(&ds.ones, ID_LEN * sizeof(float)));
    K2Y_CUDA_CHECK(cudaMalloc(&ds.zeros, ID_LEN * sizeof(float)));
    std::vector<float> ones(ID_LEN, 1.f);
    K2Y_CUDA_CHECK(cudaMemcpy(ds.ones, ones.data(), ID_LEN * sizeof(float), cudaMemcpyHostToDevice));
    K2Y_CUDA_CHECK(cudaMemset(ds.zeros, 0, ID_LEN * sizeof(float)));
    ds.init = true;
    return K2Y_OK;
}

}  // namespace

int tc_pack(TcWeights &w, const float *kernel_kn, int K, int N, bool with_permuted) {
    tc_free(w);
    int rc = tc_init();
    if (rc != K2Y_OK) return rc;
    w.K = K;
    w.N = N;
    w.Kpad = (K + BK - 1) / BK * BK;
    w.Npad = (N + 15) / 16 * 16;  // tiles wider than the remainder are zero-filled by TMA
    std::vector<float> hi((size_t)w.Npad * w.Kpad, 0.f), lo((size_t)w.Npad * w.Kpad, 0.f);
    for (int k = 0; k < K; ++k)
        for (int n = 0; n < N; ++n) {
            const float v = kernel_kn[(size_t)k * N + n];
            const float h = tf32_rna_host(v);
            hi[(size_t)n * w.Kpad + k] = h;
            lo[(size_t)n * w.Kpad + k] = tf32_rna_host(v - h);
        }
    K2Y_CUDA_CHECK(cudaMalloc(&w.d_hi, hi.size() * sizeof(float)));
    K2Y_CUDA_CHECK(cudaMalloc(&w.d_lo, lo.size() * sizeof(float)));
    K2Y_CUDA_CHECK(cudaMemcpy(w.d_hi, hi.data(), hi.size() * sizeof(float), cudaMemcpyHostToDevice));
    K2Y_CUDA_CHECK(cudaMemcpy(w.d_lo, lo.data(), lo.size() * sizeof(float), cudaMemcpyHostToDevice));
    // bf16x3 planes
    w.Kpad64 = (K + 63) / 64 * 64;
    std::vector<uint16_t> bh((size_t)w.Npad * w.Kpad64, 0), bm((size_t)w.Npad * w.Kpad64, 0);
    for (int k = 0; k < K; ++k)
        for (int n = 0; n < N; ++n) {
            const float v = kernel_kn[(size_t)k * N + n];
            const uint16_t h = bf16_rne_host(v);
            bh[(size_t)n * w.Kpad64 + k] = h;
            bm[(size_t)n * w.Kpad64 + k] = bf16_rne_host(v - bf16_to_float_host(h));
        }
    K2Y_CUDA_CHECK(cudaMalloc(&w.d_bh, bh.size() * sizeof(uint16_t)));
    K2Y_CUDA_CHECK(cudaMalloc(&w.d_bm, bm.size() * sizeof(uint16_t)));
    K2Y_CUDA_CHECK(cudaMemcpy(w.d_bh, bh.data(), bh.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    K2Y_CUDA_CHECK(cudaMemcpy(w.d_bm, bm.data(), bm.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    if (with_permuted) {
        std::vector<uint16_t> ph((size_t)w.Npad * w.Kpad64, 0), pm((size_t)w.Npad * w.Kpad64, 0);
        for (int kp = 0; kp < w.Kpad64; ++kp) {
            const int base = kp & ~31, kap = kp & 31;
            const int u = kap >> 4, m = (kap >> 2) & 3, e = kap & 3;
            const int k = base + 8 * m + 4 * u + e;   // the channel stored at k position kp
            if (k >= K) continue;
            for (int n = 0; n < N; ++n) {
                ph[(size_t)n * w.Kpad64 + kp] = bh[(size_t)n * w.Kpad64 + k];
                pm[(size_t)n * w.Kpad64 + kp] = bm[(size_t)n * w.Kpad64 + k];
            }
        }
        K2Y_CUDA_CHECK(cudaMalloc(&w.d_bh_p, ph.size() * sizeof(uint16_t)));
        K2Y_CUDA_CHECK(cudaMalloc(&w.d_bm_p, pm.size() * sizeof(uint16_t)));
        K2Y_CUDA_CHECK(cudaMemcpy(w.d_bh_p, ph.data(), ph.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        K2Y_CUDA_CHECK(cudaMemcpy(w.d_bm_p, pm.data(), pm.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    }
    return K2Y_OK;
}

void tc_free(TcWeights &w) {
    cudaFree(w.d_hi);
    cudaFree(w.d_lo);
    cudaFree(w.d_bh);
    cudaFree(w.d_bm);
    cudaFree(w.d_bh_p);
    cudaFree(w.d_bm_p);
    w = TcWeights();
}

// bf16x3 needs whole 64-channel k-blocks per tap in the gather path; layers that do not fit run as 3xTF32
static bool is_plain_1x1(const ConvArgs &a);
// Gather convs over ONE plain source whose channel count is HALF a k-block: the k-block then holds two consecutive taps —
// 16 channels with the 32-wide tf32 k-blocks (tiny_yolo's 16->32 conv), 32 channels with the 64-wide bf16 k-blocks (the
// 32-channel 3x3 convs of Darknet-53 and tiny_yolo), where each tap fills exactly one 16 KB sub-tile of the stage.
static bool is_half_taps(const ConvArgs &a, int channels) {
    return !is_plain_1x1(a) && a.C0 == channels && a.C1 == 0 && a.src1 == nullptr && !a.up0;
}
static int effective_mode(const ConvArgs &a, int math_mode) {
    if (math_mode != K2Y_MATH_TC_BF16X3) return math_mode;
    // K <= 32 fits one 32-wide tf32 k-block: half the A bytes staged and converted per tile of the 64-wide bf16 k-block
    // (conv_pw_1 of yolo_mobilev1-0.75: 56 us as 3xTF32, 64 us as bf16x3); same error class
    if (is_plain_1x1(a)) return (a.C0 + a.C1 <= 32) ? K2Y_MATH_TC_3XTF32 : math_mode;
    const int Cin = a.C0 + a.C1;
    if ((Cin % 64) == 0 && (a.C0 % 64) == 0) return math_mode;
    return is_half_taps(a, 32) ? math_mode : K2Y_MATH_TC_3XTF32;
}

static bool is_plain_1x1(const ConvArgs &a) {
    return a.kh == 1 && a.kw == 1 && a.stride == 1 && a.src1 == nullptr && !a.up0 && a.pad_t == 0 && a.pad_l == 0;
}

bool tc_supported(const ConvArgs &a, const TcWeights &w) {
    if (!w.d_hi || !get_encode()) return false;
    const int Cin = a.C0 + a.C1;
    if (is_plain_1x1(a)) return (Cin % 4) == 0 && (((uintptr_t)a.src0) & 15) == 0;  // TMA: 16-byte row pitch
    // gather path: every k-block is one 128-byte channel run of one tap of one source ...
    if ((Cin % BK) == 0 && (a.C0 % BK) == 0)
        return (((uintptr_t)a.src0) & 15) == 0 && (a.src1 == nullptr || (((uintptr_t)a.src1) & 15) == 0);
    // ... or two 64-byte runs of two consecutive taps (Cin == 16, one plain source; runs as 3xTF32: effective_mode)
    return is_half_taps(a, 16) && (((uintptr_t)a.src0) & 15) == 0;
}

// Depthwise 3x3 (+BN+act) followed by a plain 1x1 conv whose K fits one k-block: the depthwise result is produced straight
// into the A stage of the tensor-core kernel (never written to HBM).  bf16x3 mode only (64 k per k-block).
bool tc_dw_fusable(const DwArgs &dw, const ConvArgs &pw, const TcWeights &w, int math_mode) {
    // opt-in (K2Y_DWPW_FUSION=1): correct, but with only four producer warps per SM the depthwise loads are latency-bound
    // (6.6k / 18.8k cycles per tile on conv_dw_1 / conv_dw_2, see profiles/r01_dw_fusion.md) and the pair runs slower
    // than the two separate launches.
    const char *on = getenv("K2Y_DWPW_FUSION");
    if (!on || on[0] != '1') return false;
    if (math_mode != K2Y_MATH_TC_BF16X3 || !is_plain_1x1(pw) || !tc_supported(pw, w)) return false;
    if (pw.src0 != dw.dst || pw.C0 != dw.C || pw.C1 != 0 || pw.OH != dw.OH || pw.OW != dw.OW || pw.residual) return false;
    if (dw.C > 64 || (dw.C & 3) || (dw.OW & 3) || (pw.N & 3) || (dw.stride != 1 && dw.stride != 2)) return false;
    return (((uintptr_t)dw.src) & 15) == 0 && (((uintptr_t)pw.dst) & 15) == 0;
}

cudaError_t launch_conv_tc(const ConvArgs &a, const TcWeights &w, int math_mode, cudaStream_t st, const DwArgs *dw) {
    DevState *ds = cur_dev();
    if (!ds) return cudaErrorInvalidDevice;
    if (!ds->init && tc_init() != K2Y_OK) return cudaErrorNotReady;  // first launch on this device: attributes, identity vectors
    TcParams p;
    p.dw = 0;
    p.dw_w = p.dw_scale = p.dw_shift = nullptr;
    p.dw_act = ACT_NONE;
    p.dw_alpha = 0.f;
    p.src0 = a.src0;
    p.src1 = a.src1;
    p.H = a.H;
    p.W = a.W;
    p.C0 = a.C0;
    p.C1 = a.C1;
    p.up0 = a.up0;
    p.OH = a.OH;
    p.OW = a.OW;
    p.kh = a.kh;
    p.kw = a.kw;
    p.stride = a.stride;
    p.pad_t = a.pad_t;
    p.pad_l = a.pad_l;
    p.dst = a.dst;
    p.residual = a.residual;
    p.scale = a.scale;
    p.shift = a.shift;
    p.act = a.act;
    p.alpha = a.alpha;
    p.act_slope = a.act == ACT_NONE ? 1.f : (a.act == ACT_LEAKY ? a.alpha : 0.f);
    p.act_clamp = a.act == ACT_RELU6 ? 6.f : __int_as_float_host(0x7f800000);
    p.M = a.B * a.OH * a.OW;
    p.N = a.N;
    p.K = w.K;
    if (dw) {  // fused depthwise producer: the gather warps read the depthwise INPUT
        p.dw = 1;
        p.src0 = dw->src;
        p.H = dw->H;
        p.W = dw->W;
        p.stride = dw->stride;
        p.pad_t = dw->pad_t;
        p.pad_l = dw->pad_l;
        p.dw_w = dw->w;
        p.dw_scale = dw->scale;
        p.dw_shift = dw->shift;
        p.dw_act = dw->act;
        p.dw_alpha = dw->alpha;
    }
    math_mode = effective_mode(a, math_mode);
    p.bf16 = (math_mode == K2Y_MATH_TC_BF16X3) ? 1 : 0;
    p.three_x = (math_mode == K2Y_MATH_TC_TF32) ? 0 : 1;
    p.a_boxes = p.bf16 ? 2 : 1;
    p.half_taps = (!dw && is_half_taps(a, p.bf16 ? 32 : 16)) ? 1 : 0;
    p.nkb = p.bf16 ? w.Kpad64 / 64 : w.Kpad / BK;
    p.cluster = pick_cluster(p.M, p.nkb);
    pick_tile(p.M, a.N, p.nkb, p.three_x != 0, p.bf16 != 0, !is_plain_1x1(a) || dw, p.cluster, a.tc_scratch ? a.tc_scratch_bytes : 0,
              budget_sms(a.sm_limit), &p.BN, &p.k_splits);
    p.n_tiles = (w.Npad + p.BN - 1) / p.BN;
    p.m_tiles = (p.M + BM - 1) / BM;
    p.kb_per_split = (p.nkb + p.k_splits - 1) / p.k_splits;
    p.k_splits = (p.nkb + p.kb_per_split - 1) / p.kb_per_split;  // no empty slices
    const size_t stage_bytes = (size_t)A_TILE_BYTES * p.a_boxes + (size_t)p.BN * 128 * (p.three_x ? 2 : 1);
    const size_t fixed = FIXED_SMEM;
    int stages = (int)((g_max_smem - fixed) / stage_bytes);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (p.three_x && stages > (512 - 2 * p.BN) / 64) stages = (512 - 2 * p.BN) / 64;  // TMEM: 2 accumulators + 64 columns of A per stage
    if (stages < 2) return cudaErrorInvalidConfiguration;
    p.stages = stages;
    uint32_t cols = 32;
    while (cols < 2u * (uint32_t)p.BN + (p.three_x ? 64u * (uint32_t)stages : 0u)) cols <<= 1;
    p.tmem_cols = cols;
    const size_t smem = fixed + (size_t)stages * stage_bytes;

    const bool gather = !is_plain_1x1(a) || dw != nullptr;
    p.epi_groups = (!gather && p.BN > 32 && !getenv("K2Y_TC_ONE_EPI")) ? 2 : 1;
    CUtensorMap map_a, map_bhi, map_blo, map_out, map_res;
    memset(&map_a, 0, sizeof(map_a));
    memset(&map_out, 0, sizeof(map_out));
    memset(&map_res, 0, sizeof(map_res));
    // TMA-store epilogue: 16-byte rows, and 32-column boxes that never reach into a neighbouring n-tile
    p.tma_store = ((a.N & 3) == 0 && (((uintptr_t)a.dst) & 15) == 0 && (a.residual == nullptr || (((uintptr_t)a.residual) & 15) == 0) &&
                   (p.n_tiles == 1 || (p.BN & 31) == 0)) ? 1 : 0;
    p.res_tma = 0;
    if (getenv("K2Y_TC_NO_TMA_STORE") && p.k_splits == 1) p.tma_store = 0;
    const size_t mpad = (size_t)p.m_tiles * BM;
    float *scratch = a.tc_scratch;
    if (p.k_splits > 1) {
        if (!scratch) {  // no caller-owned scratch (k2y_conv2d): the device's default one, created on first use
            if (!ds->scratch) {
                cudaError_t me = cudaMalloc(&ds->scratch, DEFAULT_SCRATCH_BYTES);
                if (me != cudaSuccess) return me;
                ds->scratch_bytes = DEFAULT_SCRATCH_BYTES;
            }
            scratch = ds->scratch;
        }
        // partial pass: raw accumulators -> scratch through the TMA-store epilogue (identity scale, no activation)
        p.tma_store = 1;
        p.dst = scratch;
        p.residual = nullptr;
        p.scale = ds->ones;
        p.shift = ds->zeros;
        p.act = ACT_NONE;
        p.act_slope = 1.f;
        p.act_clamp = __int_as_float_host(0x7f800000);
        if (!make_map_2d(&map_out, scratch, (uint64_t)p.k_splits * mpad, (uint64_t)a.N, 32)) return cudaErrorInvalidValue;
    } else if (p.tma_store) {
        if (!make_map_2d(&map_out, a.dst, (uint64_t)p.M, (uint64_t)a.N, 32)) return cudaErrorInvalidValue;
        if (a.residual) {
            if (!make_map_2d(&map_res, a.residual, (uint64_t)p.M, (uint64_t)a.N, 32)) return cudaErrorInvalidValue;
            p.res_tma = 1;
        }
    }
    if (!gather && !make_map_2d(&map_a, a.src0, (uint64_t)p.M, (uint64_t)(a.C0 + a.C1), BM)) return cudaErrorInvalidValue;
    if (p.bf16) {
        if (!make_map_2d(&map_bhi, w.d_bh, (uint64_t)w.Npad, (uint64_t)w.Kpad64, (uint32_t)(p.BN / p.cluster), 0, true)) return cudaErrorInvalidValue;
        if (!make_map_2d(&map_blo, w.d_bm, (uint64_t)w.Npad, (uint64_t)w.Kpad64, (uint32_t)(p.BN / p.cluster), 0, true)) return cudaErrorInvalidValue;
    } else {
        if (!make_map_2d(&map_bhi, w.d_hi, (uint64_t)w.Npad, (uint64_t)w.Kpad, (uint32_t)(p.BN / p.cluster))) return cudaErrorInvalidValue;
        if (!make_map_2d(&map_blo, w.d_lo, (uint64_t)w.Npad, (uint64_t)w.Kpad, (uint32_t)(p.BN / p.cluster))) return cudaErrorInvalidValue;
    }

    const int tiles = ((p.m_tiles + p.cluster - 1) / p.cluster) * p.n_tiles * p.k_splits;  // work items per cluster
    const int max_clusters = budget_sms(a.sm_limit) / p.cluster;
    const int grid = (tiles < max_clusters ? tiles : max_clusters) * p.cluster;
    p.trace = nullptr;
    p.dbg = getenv("K2Y_TC_DBG") ? atoi(getenv("K2Y_TC_DBG")) : 0;
    const char *tr = getenv("K2Y_TC_TRACE");
    long long *d_trace = nullptr;
    if (tr && tr[0] == '1') {
        cudaMalloc(&d_trace, (size_t)grid * 64 * sizeof(long long));
        cudaMemset(d_trace, 0, (size_t)grid * 64 * sizeof(long long));
        p.trace = d_trace;
        fprintf(stderr, "[tc-trace] M=%d N=%d K=%d BN=%d items=%d (m %d x n %d x k %d) cluster=%d nkb=%d stages=%d grid=%d gather=%d 3x=%d smem=%zu\n",
                p.M, p.N, p.K, p.BN, tiles, p.m_tiles, p.n_tiles, p.k_splits, p.cluster, p.nkb, p.stages, grid, (int)gather, p.three_x + 2 * p.bf16, smem);
    }
    {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(NUM_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)p.cluster;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = (pdl_enabled() && !d_trace) ? 2 : 1;
        cudaError_t le = dw       ? cudaLaunchKernelEx(&cfg, conv_tc_kernel<true, true>, map_a, map_bhi, map_blo, map_out, map_res, p)
                         : gather ? cudaLaunchKernelEx(&cfg, conv_tc_kernel<true, false>, map_a, map_bhi, map_blo, map_out, map_res, p)
                                  : cudaLaunchKernelEx(&cfg, conv_tc_kernel<false, false>, map_a, map_bhi, map_blo, map_out, map_res, p);
        if (le != cudaSuccess) return le;
    }
    if (d_trace) {
        cudaStreamSynchronize(st);
        std::vector<long long> h((size_t)grid * 64);
        cudaMemcpy(h.data(), d_trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        cudaFree(d_trace);
        long long t0 = h[0];
        for (int b = 0; b < grid; ++b)
            if (h[(size_t)b * 64] && h[(size_t)b * 64] < t0) t0 = h[(size_t)b * 64];
        static const char *names[13] = {"start", "setup_done", "tma_first", "tma_tile0_last", "mma_full_b0", "mma_conv0", "mma_tile0_commit",
                                        "mma_all_issued", "epi_tile0_ready", "epi_tile0_done", "epi_all_done", "all_synced", "dealloc"};
        for (int b : {0, grid / 2, grid - 1}) {
            fprintf(stderr, "[tc-trace] cta %d:", b);
            for (int i = 0; i < 13; ++i) fprintf(stderr, " %s=%.2fus", names[i], h[(size_t)b * 64 + i] ? (h[(size_t)b * 64 + i] - t0) * 1e-3 : -1.0);
            fprintf(stderr, "\n");
        }
        // per-k-block pipeline of CTA 0's first tile: tma issue / landed (converter sees it) / converted / mma issued / slot free again
        fprintf(stderr, "[tc-trace] epilogue chunk cycles: tmem_ld+wait=%lld math=%lld wait_read=%lld sts+fence=%lld tma_issue=%lld\n", h[48],
                h[49], h[50], h[51], h[52]);
        fprintf(stderr, "[tc-trace] wait cycles (cta 0): epilogue<-accumulator %lld | tma<-free stage %lld | mma<-free accumulator %lld, <-weights %lld, <-converter %lld | converter<-A tile %lld\n",
                h[56], h[57], h[58], h[59], h[60], h[61]);
        if (p.dw) fprintf(stderr, "[tc-trace] dw producer (cta 0): wait_empty=%lld produce=%lld cycles over %lld tiles\n", h[53], h[54], h[55]);
        fprintf(stderr, "[tc-trace] kb: tma_issue landed converted mma_issued\n");
        for (int kb = 0; kb < 8; ++kb) {
            fprintf(stderr, "[tc-trace]  %d:", kb);
            for (int e = 0; e < 4; ++e) fprintf(stderr, " %.2f", h[16 + kb * 4 + e] ? (h[16 + kb * 4 + e] - t0) * 1e-3 : -1.0);
            fprintf(stderr, "\n");
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (p.k_splits > 1) {
        const float slope = a.act == ACT_NONE ? 1.f : (a.act == ACT_LEAKY ? a.alpha : 0.f);
        const float clamp = a.act == ACT_RELU6 ? 6.f : __int_as_float_host(0x7f800000);
        const size_t total = (size_t)p.M * (a.N / 4);
        launch_k(splitk_reduce_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, (const float *)scratch, a.dst, a.residual,
                 a.scale, a.shift, p.M, a.N, p.k_splits, mpad * a.N, slope, clamp);
        e = cudaGetLastError();
    }
    return e;
}

int tc_launch_count(const ConvArgs &a, const TcWeights &w, int math_mode) {
    int bn = 0, splits = 1;
    const int mode = effective_mode(a, math_mode);
    const bool bf = mode == K2Y_MATH_TC_BF16X3;
    const int M = a.B * a.OH * a.OW, nkb = bf ? w.Kpad64 / 64 : w.Kpad / BK;
    pick_tile(M, a.N, nkb, mode != K2Y_MATH_TC_TF32, bf, !is_plain_1x1(a), pick_cluster(M, nkb), a.tc_scratch ? a.tc_scratch_bytes : 0,
              budget_sms(a.sm_limit), &bn, &splits);
    return splits > 1 ? 2 : 1;
}

// Upper bound of the split-K scratch a conv of this shape can ask for (8 slices of [m_tiles*128][N] partials, capped at the
// planner's default limit); 0 when the shape is never split.  A net sizes its own scratch with the maximum over its layers.
size_t tc_scratch_bound(const ConvArgs &a, const TcWeights &w) {
    const int M = a.B * a.OH * a.OW;
    if ((a.N & 3) != 0 || a.N > ID_LEN || w.Kpad / BK < 24) return 0;
    const size_t need = (size_t)8 * ((M + BM - 1) / BM) * BM * a.N * sizeof(float);
    return need < DEFAULT_SCRATCH_BYTES ? need : DEFAULT_SCRATCH_BYTES;
}

}  // namespace k2y

// Host-only view of the tile planner (no GPU needed: without a device it assumes a B200 — 148 SMs, 227 KB of shared
// memory): which tile width, split-K factor, pipeline depth and CTA-pair mode a dense conv of this GEMM shape gets.
extern "C" int k2y_tc_plan_budget(int M, int N, int K, int ksize, int math_mode, int sm_limit, int *bn, int *k_splits, int *stages,
                                  int *cluster, int *effective_math, int *grid);

extern "C" int k2y_tc_plan(int M, int N, int K, int ksize, int math_mode, int *bn, int *k_splits, int *stages, int *cluster,
                           int *effective_math) {
    int grid = 0;
    return k2y_tc_plan_budget(M, N, K, ksize, math_mode, 0, bn, k_splits, stages, cluster, effective_math, &grid);
}

// The same view under an SM budget (k2y_net_set_sm_limit): also reports the persistent grid size.
extern "C" int k2y_tc_plan_budget(int M, int N, int K, int ksize, int math_mode, int sm_limit, int *bn, int *k_splits, int *stages,
                                  int *cluster, int *effective_math, int *grid) {
    using namespace k2y;
    if (M <= 0 || N <= 0 || K <= 0 || (ksize != 1 && ksize != 3) || !bn || !k_splits || !stages || !cluster || !effective_math ||
        (math_mode != K2Y_MATH_TC_3XTF32 && math_mode != K2Y_MATH_TC_TF32 && math_mode != K2Y_MATH_TC_BF16X3) || sm_limit < 0 || !grid) {
        set_error("k2y_tc_plan: bad arguments");
        return K2Y_ERR_INVALID;
    }
    const bool gather = ksize == 3;
    int mode = math_mode;
    if (mode == K2Y_MATH_TC_BF16X3) {  // effective_mode(): same rules, expressed on (K, ksize)
        const int cin = gather ? K / 9 : K;
        if (!gather) mode = cin <= 32 ? K2Y_MATH_TC_3XTF32 : mode;
        else if ((cin % 64) != 0 && cin != 32) mode = K2Y_MATH_TC_3XTF32;   // 32 channels: two taps per bf16 k-block
    }
    const bool bf = mode == K2Y_MATH_TC_BF16X3, three_x = mode != K2Y_MATH_TC_TF32;
    const int nkb = bf ? (K + 63) / 64 : (K + BK - 1) / BK;
    *cluster = pick_cluster(M, nkb);
    const int sms = budget_sms(sm_limit);
    pick_tile(M, N, nkb, three_x, bf, gather, *cluster, 0, sms, bn, k_splits);
    const size_t stage_bytes = (size_t)A_TILE_BYTES * (bf ? 2 : 1) + (size_t)(*bn) * 128 * (three_x ? 2 : 1);
    int st = (int)(((g_max_smem ? g_max_smem : 232448) - FIXED_SMEM) / stage_bytes);
    if (st > MAX_STAGES) st = MAX_STAGES;
    if (three_x && st > (512 - 2 * *bn) / 64) st = (512 - 2 * *bn) / 64;
    *stages = st;
    const int kb_per_split = (nkb + *k_splits - 1) / *k_splits;
    *k_splits = (nkb + kb_per_split - 1) / kb_per_split;
    *effective_math = mode;
    {
        const int n16 = (N + 15) / 16 * 16, n_tiles = (n16 + *bn - 1) / *bn, m_tiles = (M + BM - 1) / BM;
        const int items = ((m_tiles + *cluster - 1) / *cluster) * n_tiles * *k_splits, max_clusters = sms / *cluster;
        *grid = (items < max_clusters ? items : max_clusters) * *cluster;
    }
    return K2Y_OK;
}

