// Fused depthwise 3x3 (stride 1) + BN + activation -> pointwise 1x1 + BN + activation on sm_100a: one launch per
// MobileNet block (models/keras_mobilenet.py:359-436 `_depthwise_conv_block`: DepthwiseConv2D 3x3 SAME, BN, ReLU,
// Conv2D 1x1, BN, LeakyReLU).  The depthwise result never leaves the SM: it is computed from a TMA-staged NHWC window in
// shared memory straight into TENSOR MEMORY, in the layout the tcgen05 A operand wants, split into (hi, mid) bf16 planes.
//
//   tile      128 output pixels = four 8x4-pixel blocks (2x2, 1x4 or 4x1 blocks, chosen per layer), all N output channels
//   warp 0    TMA producer: 4-D tensor-map loads of [32 channels][tile+halo columns][tile+halo rows] windows of the input
//             (start coordinate -1: the out-of-bounds fill IS the zero padding), 128B-swizzled, 3-deep ring; weight
//             k-blocks (hi/mid bf16 planes, 64 k each) through a 2-deep ring
//   warps 2-9 depthwise: warp = (TMEM lane quarter = one 8x4 block, 32-channel half of the 64-channel k-block).  The
//             16x256b tensor-memory store shape gives lane t the rows t/4 + {0,8,16,24} and column pairs t%4 (+4) of its
//             quarter; GEMM rows are numbered so that these are FOUR VERTICALLY ADJACENT pixels of one column, i.e. a
//             thread slides a 6-row x 3-column register window down its pixel column (4.5 shared-memory loads per output
//             float4 instead of 9), 8 consecutive lanes read 8 consecutive pixels (conflict-free under the 128B swizzle),
//             and the converted bf16 pairs go to TMEM with tcgen05.st — shared memory never holds the depthwise output
//   warp 1    MMA issuer: tcgen05.mma kind::f16, A from tensor memory, B (weights) from shared memory, bf16x3 split
//             (mid*hi + hi*mid + hi*hi), fp32 accumulators in TMEM (two N-passes of 192 for N = 384)
//   warps 10-13 epilogue: tcgen05.ld, folded BN + activation, 128B-swizzled staging, ONE 4-D TMA store per 8x4 block and
//             32-channel chunk (rows/columns beyond the image are clipped by the TMA)
//
// Same arithmetic as the unfused pair (dwconv3x3_kernel + conv_tc_kernel in bf16x3 mode): tap order (ky,kx) ascending with
// fmaf, BN as one fmaf, the same bf16 split.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "gemm_tc.h"
#include "tc_ptx.cuh"

namespace k2y {

namespace {

using namespace ptx;

constexpr int DP_THREADS = 18 * 32;   // warps 0-7 depthwise, 8-15 epilogue, 16 MMA issuer, 17 TMA producer
// (the SM's warp arbiter favours HIGH warp ids: the two single-thread roles everything else waits for get the top ids — with the
// producer as warp 0 its window loads were issued ~1.3 us late behind the busy epilogue / depthwise warps of its sub-partition)
constexpr int W_DW0 = 0, W_EPI0 = 8, W_MMA = 16, W_TMA = 17;
constexpr int WIN_STAGES = 3;
constexpr int B_STAGES = 2;
constexpr int MAX_A_STAGES = 4;

struct DwPwParams {
    int B, H, W, C, N;
    int tw, th;                 // tile = 8*tw x 4*th pixels, tw*th == 4
    int tiles_x, tiles_y, num_tiles;
    int nkb;                    // 64-channel k-blocks
    int n_pass, BN;             // UMMA N per pass (multiple of 16, <= 192); n_pass*BN >= N
    int win_cols, win_rows;
    uint32_t win_bytes;         // bytes one window chunk occupies in shared memory (1024-aligned)
    uint32_t win_tx;            // bytes the TMA actually delivers per chunk
    int acc_stages, a_stages, stg_bufs;
    uint32_t tmem_cols;
    const float *dw_pack;       // [11][cpad]: 9 tap rows, folded-BN scale, shift (zero beyond C)
    int dw_act;
    float dw_alpha;
    const float *scale, *shift; // pointwise folded BN [N]
    int act;
    float act_slope;
    // shared-memory carve (byte offsets from the 1024-aligned base)
    uint32_t off_b, off_stage, off_ss, off_dw, off_bars, cpad;
    long long *trace;           // optional [grid][64] globaltimer stamps (K2Y_TC_TRACE=1)
};

__device__ __forceinline__ long long gtime_ns() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define DP_TRACE(slot)                                                            \
    do {                                                                          \
        if (p.trace) p.trace[(size_t)blockIdx.x * 128 + (slot)] = gtime_ns();      \
    } while (0)

struct __align__(8) DpBarriers {
    uint64_t win_full[WIN_STAGES], win_empty[WIN_STAGES];
    uint64_t b_full[B_STAGES], b_empty[B_STAGES];
    uint64_t a_full[MAX_A_STAGES], a_empty[MAX_A_STAGES];
    uint64_t tmem_full[2], tmem_empty[2];
    uint64_t par_full;
    uint32_t tmem_slot;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(src), "r"(c0),
                 "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}
// 16 lanes x 8 columns of tensor memory: lane t of the warp writes row t/4 (r0, r1) and row t/4 + 8 (r2, r3), columns
// 2*(t%4) + {0,1}, relative to the address
__device__ __forceinline__ void tmem_st_16x256b_x1(uint32_t taddr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
    asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ float dw_act_f(float v, int act, float alpha) {
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
    if (act == ACT_LEAKY) return v >= 0.f ? v : v * alpha;
    return v;
}

// 3x3 depthwise accumulation of one work unit: 4 vertically adjacent output pixels (rows 0..3 of the block) of one channel quad,
// from a 6-row x 3-column window.  `base` points at the thread's top-left window pixel (128-byte lines, 128B-swizzled); WC (window
// columns) is a template constant so that every load is `base + immediate + swz[c & 7]` — the line offset folds into the
// instruction, the swizzle term comes from 8 per-thread constants (the pixel index mod 8 of tap c is (i + c) mod 8).
template <int WC>
__device__ __forceinline__ void dw_accumulate(const uint8_t *base, const uint32_t (&swz)[8], const float4 (&w9)[9], float4 (&acc)[4]) {
#pragma unroll
    for (int wy = 0; wy < 6; ++wy) {
        float4 v[3];
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const int c = wy * WC + dx;
            v[dx] = *reinterpret_cast<const float4 *>(base + c * 128 + swz[c & 7]);
        }
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int j = wy - dy;   // output row fed by this input row through tap row dy
            if (j < 0 || j > 3) continue;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const float4 ww = w9[dy * 3 + dx];
                acc[j].x = fmaf(v[dx].x, ww.x, acc[j].x);
                acc[j].y = fmaf(v[dx].y, ww.y, acc[j].y);
                acc[j].z = fmaf(v[dx].z, ww.z, acc[j].z);
                acc[j].w = fmaf(v[dx].w, ww.w, acc[j].w);
            }
        }
    }
}

__global__ void __launch_bounds__(DP_THREADS, 1)
dwpw_tc_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_bhi,
               const __grid_constant__ CUtensorMap map_blo, const __grid_constant__ CUtensorMap map_out, const DwPwParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    DpBarriers *bars = reinterpret_cast<DpBarriers *>(smem_gen + p.off_bars);
    const float *s_dw = reinterpret_cast<const float *>(smem_gen + p.off_dw);   // [9][cpad] taps, [cpad] scale, [cpad] shift
    const uint32_t b_plane = (uint32_t)p.BN * 128u;                              // one bf16 weight plane of a k-block: BN rows x 128 B

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();

    if (threadIdx.x == 0) {
        DP_TRACE(0);
        for (int s = 0; s < WIN_STAGES; ++s) {
            mbar_init(smem_u32(&bars->win_full[s]), 1);
            mbar_init(smem_u32(&bars->win_empty[s]), 8);    // one arrival per depthwise warp
        }
        for (int s = 0; s < B_STAGES; ++s) {
            mbar_init(smem_u32(&bars->b_full[s]), 1);
            mbar_init(smem_u32(&bars->b_empty[s]), 1);
        }
        for (int s = 0; s < MAX_A_STAGES; ++s) {
            mbar_init(smem_u32(&bars->a_full[s]), 8);
            mbar_init(smem_u32(&bars->a_empty[s]), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(&bars->tmem_full[a]), 1);
            mbar_init(smem_u32(&bars->tmem_empty[a]), 8);  // one arrival per epilogue warp
        }
        mbar_init(smem_u32(&bars->par_full), 1);
        fence_barrier_init();
        // the depthwise taps + folded BN of this layer: constants, not produced by a predecessor -> one bulk copy, before the
        // dependency wait
        const uint32_t par_bytes = 11u * p.cpad * 4u;
        mbar_arrive_expect_tx(smem_u32(&bars->par_full), par_bytes);
        bulk_copy_g2s(smem_base + p.off_dw, p.dw_pack, par_bytes, smem_u32(&bars->par_full));
        prefetch_tmap(&map_in);
        prefetch_tmap(&map_bhi);
        prefetch_tmap(&map_blo);
        prefetch_tmap(&map_out);
    }
    if (warp == W_MMA) tmem_alloc(smem_u32(&bars->tmem_slot), p.tmem_cols);
    {
        float *ss = reinterpret_cast<float *>(smem_gen + p.off_ss);   // [n_pass*BN] scale, then shift
        const int ncol = p.n_pass * p.BN;
        for (int i = threadIdx.x; i < 2 * ncol; i += DP_THREADS) {
            const int n = i < ncol ? i : i - ncol;
            ss[i] = n < p.N ? __ldg((i < ncol ? p.scale : p.shift) + n) : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_slot;
    pdl_wait();  // everything above overlaps the previous kernel's tail; the activations below do not
    if (threadIdx.x == 0) DP_TRACE(1);

    const int acc_cols = p.n_pass * p.BN;
    const uint32_t tmem_a0 = tmem_base + (uint32_t)(p.acc_stages * acc_cols);   // A stages: 64 columns each (hi 32 | mid 32)
    auto win_slot = [&](int s) { return smem_base + (uint32_t)s * p.win_bytes; };
    auto b_slot = [&](int s) { return smem_base + p.off_b + (uint32_t)s * 2u * b_plane; };
    const int tiles_per_img = p.tiles_x * p.tiles_y;

    if (warp == W_TMA) {
        // ================= TMA producer =================
        uint32_t wseq = 0, bseq = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
            const int b = t / tiles_per_img, r = t - b * tiles_per_img;
            const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
            const int x0 = tx * 8 * p.tw, y0 = ty * 4 * p.th;
            for (int kb = 0; kb < p.nkb; ++kb) {
                for (int ch = 0; ch < 2; ++ch) {
                    const int c0 = kb * 64 + ch * 32;
                    if (c0 >= p.C) break;
                    const uint32_t s = wseq % WIN_STAGES, ph = (wseq / WIN_STAGES) & 1u;
                    mbar_wait(smem_u32(&bars->win_empty[s]), ph ^ 1u);
                    if (elect_one_sync()) {
                        const uint32_t fb = smem_u32(&bars->win_full[s]);
                        mbar_arrive_expect_tx(fb, p.win_tx);
                        tma_load_4d(win_slot((int)s), &map_in, fb, c0, x0 - 1, y0 - 1, b);
                        if (t == (int)blockIdx.x && wseq < 6) DP_TRACE(2 + wseq);
                        if (p.trace && kb == 0 && ch == 0 && (t - (int)blockIdx.x) / (int)gridDim.x < 10) DP_TRACE(64 + 6 * ((t - (int)blockIdx.x) / (int)gridDim.x) + 5);
                    }
                    __syncwarp();
                    ++wseq;
                }
                for (int np = 0; np < p.n_pass; ++np) {
                    const uint32_t s = bseq % B_STAGES, ph = (bseq / B_STAGES) & 1u;
                    mbar_wait(smem_u32(&bars->b_empty[s]), ph ^ 1u);
                    if (elect_one_sync()) {
                        const uint32_t fb = smem_u32(&bars->b_full[s]);
                        mbar_arrive_expect_tx(fb, 2u * b_plane);
                        tma_load_2d(b_slot((int)s), &map_bhi, fb, kb * 64, np * p.BN);
                        tma_load_2d(b_slot((int)s) + b_plane, &map_blo, fb, kb * 64, np * p.BN);
                    }
                    __syncwarp();
                    ++bseq;
                }
            }
        }
    } else if (warp == W_MMA) {
        // ================= MMA issuer =================
        const uint32_t idesc = make_idesc_bf16(128, p.BN);
        uint32_t kc = 0, bseq = 0, it = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
            const uint32_t a = it % (uint32_t)p.acc_stages, aph = (it / (uint32_t)p.acc_stages) & 1u;
            mbar_wait(smem_u32(&bars->tmem_empty[a]), aph ^ 1u);
            tc_fence_after();
            for (int kb = 0; kb < p.nkb; ++kb, ++kc) {
                const uint32_t as = kc % (uint32_t)p.a_stages, asph = (kc / (uint32_t)p.a_stages) & 1u;
                mbar_wait(smem_u32(&bars->a_full[as]), asph);
                tc_fence_after();
                if (lane == 0 && t == (int)blockIdx.x && kb < 8) DP_TRACE(32 + kb);
                const int kvalid = min(64, p.C - kb * 64);
                const int nks = (kvalid + 15) >> 4;                     // 16-channel k-steps that hold data
                const uint32_t ta_hi = tmem_a0 + as * 64u, ta_lo = ta_hi + 32u;
                for (int np = 0; np < p.n_pass; ++np, ++bseq) {
                    const uint32_t bs = bseq % B_STAGES, bph = (bseq / B_STAGES) & 1u;
                    mbar_wait(smem_u32(&bars->b_full[bs]), bph);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint32_t d = tmem_base + a * (uint32_t)acc_cols + (uint32_t)(np * p.BN);
                        const uint64_t b_hi = make_desc_sw128(b_slot((int)bs)), b_lo = make_desc_sw128(b_slot((int)bs) + b_plane);
                        for (int kk = 0; kk < nks; ++kk) {
                            const uint64_t adv = (uint64_t)((kk * 32) >> 4);
                            const uint32_t acol = (uint32_t)(kk * 8);
                            umma_bf16_ts(d, ta_lo + acol, b_hi + adv, idesc, (kb | kk) != 0);
                            umma_bf16_ts(d, ta_hi + acol, b_lo + adv, idesc, 1u);
                            umma_bf16_ts(d, ta_hi + acol, b_hi + adv, idesc, 1u);
                        }
                        umma_commit(smem_u32(&bars->b_empty[bs]));
                        if (np == p.n_pass - 1) {
                            if (t == (int)blockIdx.x && kb < 8) DP_TRACE(40 + kb);
                            umma_commit(smem_u32(&bars->a_empty[as]));
                            if (kb == p.nkb - 1) {
                                umma_commit(smem_u32(&bars->tmem_full[a]));
                                if (p.trace && it < 10) DP_TRACE(64 + 6 * it + 2);
                            }
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp < W_EPI0) {
        // ================= depthwise: shared-memory window -> registers -> (hi, mid) bf16 planes in tensor memory =================
        // Work unit = (8x4 block q, 16-channel half h of the k-block): thread (i = lane/4, m = lane%4) computes the 4 vertically
        // adjacent pixels (x = i, y = 0..3) of ONE channel quad.  Under the weight k-permutation (gemm_tc.h d_bh_p) the tensor
        // memory columns 8u + 2m (+1) of a 32-channel chunk hold channels 8m + 4u .. +3, i.e. thread m of half u reads quad
        // 2m + u: the 8 lanes of a quarter warp touch all eight 16-byte chunks of their pixels' 128-byte lines (no bank conflict).
        const int q = warp & 3;            // TMEM lane quarter this warp may touch == the 8x4 block of the tile it computes
        const int hsel = (warp - W_DW0) >> 2;  // this warp takes halves hsel and hsel + 2 of every k-block (= unit hsel of chunk 0 / 1)
        const int bx = q % p.tw, by = q / p.tw;
        const int i = lane >> 2, m = lane & 3;
        const int quad = 2 * m + hsel;     // channel quad inside a 32-channel chunk
        const int WC = p.win_cols;   // 10, 18 or 34 (8*tw + 2)
        // the thread's top-left window pixel p0 = (4*by)*WC + 8*bx + i has p0 mod 8 == i (4*WC is a multiple of 8), so tap c = wy*WC + dx
        // sits in line p0 + c whose swizzle phase is (i + c) mod 8
        const uint32_t p0_bytes = (uint32_t)((by * 4) * WC + bx * 8 + i) * 128u;
        uint32_t swz[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) swz[k] = (((uint32_t)quad) ^ ((uint32_t)(i + k) & 7u)) << 4;
        const bool tracer = warp == W_DW0 + 2 && lane == 0;
        mbar_wait(smem_u32(&bars->par_full), 0u);
        float4 w9[9];
        auto load_taps = [&](int cbase) {
            const float *wp = s_dw + cbase + quad * 4;
#pragma unroll
            for (int k = 0; k < 9; ++k) w9[k] = *reinterpret_cast<const float4 *>(wp + k * p.cpad);
        };
        const bool hoist = p.nkb == 1 && p.C <= 32;   // one chunk per tile: the taps of this thread never change
        if (hoist) load_taps(0);
        uint32_t wseq = 0, kc = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
            for (int kb = 0; kb < p.nkb; ++kb, ++kc) {
                const int nvalid = min(2, (p.C - kb * 64 + 31) >> 5);    // window chunks of this k-block
                const int nhalf = min(4, (p.C - kb * 64 + 15) >> 4);     // 16-channel halves that hold data
                const uint32_t as = kc % (uint32_t)p.a_stages, asph = (kc / (uint32_t)p.a_stages) & 1u;
                mbar_wait(smem_u32(&bars->a_empty[as]), asph ^ 1u);       // the MMAs that read this A stage have retired
                tc_fence_after();
                bool stored = false;
                for (int ch = 0; ch < nvalid; ++ch) {
                    const uint32_t seq = wseq + (uint32_t)ch;
                    const uint32_t s = seq % WIN_STAGES, ph = (seq / WIN_STAGES) & 1u;
                    mbar_wait(smem_u32(&bars->win_full[s]), ph);
                    if (tracer && t == (int)blockIdx.x && kb < 8 && ch == 0) DP_TRACE(8 + kb);
                    if (p.trace && tracer && kb == 0 && ch == 0 && (t - (int)blockIdx.x) / (int)gridDim.x < 10) DP_TRACE(64 + 6 * ((t - (int)blockIdx.x) / (int)gridDim.x));
                    const long long d_0 = (p.trace && tracer) ? clock64() : 0;
                    if (2 * ch + hsel < nhalf) {
                        const uint8_t *wptr = smem_gen + (size_t)s * p.win_bytes;   // plain loads: the compiler may batch them
                        const int cbase = kb * 64 + ch * 32;
                        if (!hoist) load_taps(cbase);
                        float4 acc[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                        const uint8_t *tbase = wptr + p0_bytes;
                        if (WC == 18) dw_accumulate<18>(tbase, swz, w9, acc);
                        else if (WC == 10) dw_accumulate<10>(tbase, swz, w9, acc);
                        else dw_accumulate<34>(tbase, swz, w9, acc);
                        const float *bnp = s_dw + 9 * p.cpad + cbase + quad * 4;
                        const float4 sc = *reinterpret_cast<const float4 *>(bnp);
                        const float4 sh = *reinterpret_cast<const float4 *>(bnp + p.cpad);
                        const uint32_t ta = tmem_a0 + as * 64u + (uint32_t)(ch * 16 + hsel * 8);
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            // pixel rows 2hf, 2hf+1 of the block = GEMM rows i + 8*(2hf), i + 8*(2hf+1) of the quarter: lane half hf
                            uint32_t hi[4], lo[4];
#pragma unroll
                            for (int jj = 0; jj < 2; ++jj) {
                                const int j = 2 * hf + jj;
                                const float o0 = dw_act_f(fmaf(acc[j].x, sc.x, sh.x), p.dw_act, p.dw_alpha);
                                const float o1 = dw_act_f(fmaf(acc[j].y, sc.y, sh.y), p.dw_act, p.dw_alpha);
                                const float o2 = dw_act_f(fmaf(acc[j].z, sc.z, sh.z), p.dw_act, p.dw_alpha);
                                const float o3 = dw_act_f(fmaf(acc[j].w, sc.w, sh.w), p.dw_act, p.dw_alpha);
                                const uint32_t h01 = pack_bf16x2(o0, o1), h23 = pack_bf16x2(o2, o3);
                                const float r0 = o0 - __uint_as_float(h01 << 16), r1 = o1 - __uint_as_float(h01 & 0xffff0000u);
                                const float r2 = o2 - __uint_as_float(h23 << 16), r3 = o3 - __uint_as_float(h23 & 0xffff0000u);
                                hi[jj * 2] = h01;
                                hi[jj * 2 + 1] = h23;
                                lo[jj * 2] = pack_bf16x2(r0, r1);
                                lo[jj * 2 + 1] = pack_bf16x2(r2, r3);
                            }
                            const uint32_t lane_off = (uint32_t)(q * 32 + hf * 16) << 16;
                            tmem_st_16x256b_x1(ta + lane_off, hi[0], hi[1], hi[2], hi[3]);
                            tmem_st_16x256b_x1(ta + 32u + lane_off, lo[0], lo[1], lo[2], lo[3]);
                        }
                        stored = true;
                    }
                    // the window slot is free as soon as every lane has consumed its values; ONE arrival per warp (256 per-thread
                    // arrivals on one mbarrier serialise in the shared-memory atomic unit)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bars->win_empty[s]));
                    if (p.trace && tracer && t == (int)blockIdx.x + 2 * (int)gridDim.x && kb == 0 && ch == 0)
                        p.trace[(size_t)blockIdx.x * 128 + 61] = clock64() - d_0;
                    if (tracer && t == (int)blockIdx.x && kb < 8 && ch == nvalid - 1) DP_TRACE(16 + kb);
                }
                if (stored) tmem_st_wait();
                if (tracer && t == (int)blockIdx.x && kb < 8) DP_TRACE(24 + kb);
                if (p.trace && tracer && kb == p.nkb - 1 && (t - (int)blockIdx.x) / (int)gridDim.x < 10) DP_TRACE(64 + 6 * ((t - (int)blockIdx.x) / (int)gridDim.x) + 1);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars->a_full[as]));
                wseq += (uint32_t)nvalid;
            }
        }
    } else {
        // ================= epilogue: two warps per TMEM quarter, alternating 32-column chunks =================
        const int ew = warp - W_EPI0;
        const int q = warp & 3;
        const int esel = ew >> 2;
        const int bx = q % p.tw, by = q / p.tw;
        const uint32_t stg_base = smem_base + p.off_stage + (uint32_t)ew * 4096u * (uint32_t)p.stg_bufs;
        const float *ssp = reinterpret_cast<const float *>(smem_gen + p.off_ss);   // [acc_cols] scale, [acc_cols] shift
        const int n_lim = (p.N + 15) & ~15;
        uint32_t it = 0, stg_it = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
            const int b = t / tiles_per_img, r = t - b * tiles_per_img;
            const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
            const int px0 = tx * 8 * p.tw + bx * 8, py0 = ty * 4 * p.th + by * 4;
            const bool inside = px0 < p.W && py0 < p.H;
            const uint32_t a = it % (uint32_t)p.acc_stages, aph = (it / (uint32_t)p.acc_stages) & 1u;
            mbar_wait(smem_u32(&bars->tmem_full[a]), aph);
            tc_fence_after();
            if (ew == 0 && lane == 0 && t == (int)blockIdx.x) DP_TRACE(48);
            if (p.trace && ew == 0 && lane == 0 && it < 10) DP_TRACE(64 + 6 * it + 3);
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * (uint32_t)acc_cols;
            for (int c0 = esel * 32; c0 < acc_cols && c0 < n_lim; c0 += 64, ++stg_it) {
                const int ncols = (acc_cols - c0) < 32 ? (acc_cols - c0) : 32;   // 32 or 16
                const uint32_t stg = stg_base + (p.stg_bufs == 2 ? (stg_it & 1u) * 4096u : 0u);
                uint32_t rr[32];
                const bool tr = p.trace && ew == 0 && lane == 0 && it == 2 && c0 == esel * 32;   // a steady-state tile
                long long c_0 = 0, c_1 = 0, c_2 = 0, c_3 = 0, c_4 = 0;
                if (tr) c_0 = clock64();
                tmem_ld16(taddr + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[16]>(&rr[0]));
                if (ncols == 32) tmem_ld16(taddr + (uint32_t)c0 + 16u, *reinterpret_cast<uint32_t(*)[16]>(&rr[16]));
                else {
#pragma unroll
                    for (int j = 16; j < 32; ++j) rr[j] = 0u;
                }
                tmem_ld_wait();
                if (tr) c_1 = clock64();
                // folded BN + activation, specialised outside the element loop
#define K2Y_DP_EPI(ACT_EXPR)                                                                                        \
    _Pragma("unroll") for (int j = 0; j < 32; j += 4) {                                                             \
        const float4 s4 = *reinterpret_cast<const float4 *>(ssp + c0 + j);                                          \
        const float4 h4 = *reinterpret_cast<const float4 *>(ssp + acc_cols + c0 + j);                               \
        float v;                                                                                                     \
        v = fmaf(__uint_as_float(rr[j]), s4.x, h4.x);     rr[j] = __float_as_uint(ACT_EXPR);                         \
        v = fmaf(__uint_as_float(rr[j + 1]), s4.y, h4.y); rr[j + 1] = __float_as_uint(ACT_EXPR);                     \
        v = fmaf(__uint_as_float(rr[j + 2]), s4.z, h4.z); rr[j + 2] = __float_as_uint(ACT_EXPR);                     \
        v = fmaf(__uint_as_float(rr[j + 3]), s4.w, h4.w); rr[j + 3] = __float_as_uint(ACT_EXPR);                     \
    }
                if (p.act == ACT_LEAKY) {
                    const float slope = p.act_slope;
                    K2Y_DP_EPI(fmaxf(v, v * slope))
                } else if (p.act == ACT_RELU) {
                    K2Y_DP_EPI(fmaxf(v, 0.f))
                } else if (p.act == ACT_RELU6) {
                    K2Y_DP_EPI(fminf(fmaxf(v, 0.f), 6.f))
                } else {
                    K2Y_DP_EPI(v)
                }
#undef K2Y_DP_EPI
                if (tr) c_2 = clock64();
                if (lane == 0) {   // the staging buffer about to be overwritten has been read by its store
                    if (p.stg_bufs == 2) tma_store_wait_read1();
                    else tma_store_wait_read0();
                }
                __syncwarp();
                if (tr) c_3 = clock64();
                // lane = GEMM row i + 8j of the quarter = pixel (x = i, y = j) of the 8x4 block = row of the [y][x][32 ch] store box
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    st_shared_v4(stg + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4), rr[j * 4], rr[j * 4 + 1], rr[j * 4 + 2],
                                 rr[j * 4 + 3]);
                fence_proxy_async();
                __syncwarp();
                if (tr) c_4 = clock64();
                if (lane == 0 && inside) {
                    tma_store_4d(&map_out, stg, c0, px0, py0, b);
                    tma_store_commit();
                }
                if (tr) {
                    long long *o = p.trace + (size_t)blockIdx.x * 128 + 56;
                    o[0] = c_1 - c_0;
                    o[1] = c_2 - c_1;
                    o[2] = c_3 - c_2;
                    o[3] = c_4 - c_3;
                    o[4] = clock64() - c_4;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars->tmem_empty[a]));
            if (ew == 0 && lane == 0 && t == (int)blockIdx.x) DP_TRACE(49);
            if (p.trace && ew == 0 && lane == 0 && it < 10) DP_TRACE(64 + 6 * it + 4);
        }
        if (lane == 0) tma_store_wait_all();
        if (ew == 0 && lane == 0) DP_TRACE(50);
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) DP_TRACE(51);
    if (warp == W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

typedef CUresult (*EncodeTiledFn4)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn4 get_encode4() {
    static EncodeTiledFn4 fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn4)ptr;
    }
    return fn;
}

// NHWC fp32 tensor [B][H][W][C] as a 4-D map (C, W, H, B), box = [box_c][box_w][box_h][1], 128B swizzle (box_c * 4 <= 128).
bool make_map_nhwc(CUtensorMap *map, const void *base, int B, int H, int W, int C, int box_c, int box_w, int box_h) {
    EncodeTiledFn4 enc = get_encode4();
    if (!enc) return false;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
bool make_map_weights(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn4 enc = get_encode4();
    if (!enc) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct DwPwPlan {
    DwPwParams p;
    size_t smem;
    int grid;
};

int g_dp_sms[64] = {0};
size_t g_dp_smem[64] = {0};

// Tile shape: the arrangement of the four 8x4 blocks that wastes the fewest pixels on this map (ties: the squarest).
void pick_blocks(int H, int W, int *tw, int *th) {
    const int opts[3][2] = {{2, 2}, {1, 4}, {4, 1}};
    long best = -1;
    for (auto &o : opts) {
        const long tiles = (long)((W + 8 * o[0] - 1) / (8 * o[0])) * ((H + 4 * o[1] - 1) / (4 * o[1]));
        if (best < 0 || tiles < best) {
            best = tiles;
            *tw = o[0];
            *th = o[1];
        }
    }
}

bool plan_dwpw(const DwArgs &dw, const ConvArgs &pw, int dev, DwPwPlan *out) {
    DwPwParams &p = out->p;
    memset(&p, 0, sizeof(p));
    p.B = dw.B;
    p.H = dw.H;
    p.W = dw.W;
    p.C = dw.C;
    p.N = pw.N;
    pick_blocks(p.H, p.W, &p.tw, &p.th);
    p.tiles_x = (p.W + 8 * p.tw - 1) / (8 * p.tw);
    p.tiles_y = (p.H + 4 * p.th - 1) / (4 * p.th);
    p.num_tiles = p.B * p.tiles_x * p.tiles_y;
    p.nkb = (p.C + 63) / 64;
    const int n16 = (p.N + 15) / 16 * 16;
    p.n_pass = n16 > 192 ? 2 : 1;
    p.BN = p.n_pass == 1 ? n16 : ((n16 / 2 + 15) / 16 * 16);
    if (p.BN > 192 || p.n_pass * p.BN > 384) return false;
    p.win_cols = 8 * p.tw + 2;
    p.win_rows = 4 * p.th + 2;
    p.win_tx = (uint32_t)(p.win_cols * p.win_rows) * 128u;
    p.win_bytes = (p.win_tx + 1023u) & ~1023u;
    const int acc = p.n_pass * p.BN;
    p.acc_stages = (2 * acc + 2 * 64 <= 512) ? 2 : 1;
    p.a_stages = (512 - p.acc_stages * acc) / 64;
    if (p.a_stages > MAX_A_STAGES) p.a_stages = MAX_A_STAGES;
    if (p.a_stages < 2) return false;
    uint32_t cols = 32;
    while (cols < (uint32_t)(p.acc_stages * acc + p.a_stages * 64)) cols <<= 1;
    p.tmem_cols = cols;
    p.cpad = (uint32_t)(p.nkb * 64);
    uint32_t off = WIN_STAGES * p.win_bytes;
    p.off_b = off;
    off += B_STAGES * 2u * (uint32_t)p.BN * 128u;
    off = (off + 1023u) & ~1023u;
    p.off_stage = off;
    // staging: 8 epilogue warps, double-buffered against their TMA stores when shared memory allows
    const uint32_t rest = ((uint32_t)(2 * acc) * 4u + 255u) / 256u * 256u + (11u * p.cpad * 4u + 255u) / 256u * 256u + (uint32_t)sizeof(DpBarriers) + 1024u;
    p.stg_bufs = (size_t)off + 8u * 8192u + rest <= g_dp_smem[dev] ? 2 : 1;
    off += 8u * 4096u * (uint32_t)p.stg_bufs;
    p.off_ss = off;
    off += ((uint32_t)(2 * acc) * 4u + 255u) & ~255u;
    p.off_dw = off;
    off += (11u * p.cpad * 4u + 255u) & ~255u;
    p.off_bars = off;
    off += (uint32_t)sizeof(DpBarriers);
    out->smem = (size_t)off + 1024;  // + alignment slack
    if (out->smem > g_dp_smem[dev]) return false;
    int sms = g_dp_sms[dev];
    if (pw.sm_limit > 0 && pw.sm_limit < sms) sms = pw.sm_limit < 8 ? 8 : pw.sm_limit;   // the net's SM budget (ConvArgs::sm_limit)
    out->grid = p.num_tiles < sms ? p.num_tiles : sms;
    return true;
}

int dp_init(int *dev_out) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return K2Y_ERR_CUDA;
    *dev_out = dev;
    if (g_dp_sms[dev]) return K2Y_OK;
    cudaDeviceProp prop;
    K2Y_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    K2Y_CUDA_CHECK(cudaFuncSetAttribute(dwpw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin));
    g_dp_smem[dev] = prop.sharedMemPerBlockOptin;
    g_dp_sms[dev] = prop.multiProcessorCount;
    return K2Y_OK;
}

}  // namespace

// The MobileNet block `dw` (depthwise 3x3, stride 1, SAME) followed by the plain 1x1 conv `pw` reading exactly its output.
bool dwpw_supported(const DwArgs &dw, const ConvArgs &pw, const TcWeights &w, int math_mode) {
    if (math_mode != K2Y_MATH_TC_BF16X3 || !w.d_bh_p || !w.d_bm_p || !get_encode4()) return false;
    if (dw.stride != 1 || dw.pad_t != 1 || dw.pad_l != 1 || dw.OH != dw.H || dw.OW != dw.W) return false;
    if (pw.kh != 1 || pw.kw != 1 || pw.stride != 1 || pw.src1 || pw.up0 || pw.residual || pw.pad_t || pw.pad_l) return false;
    if (pw.src0 != dw.dst || pw.C0 != dw.C || pw.C1 != 0 || pw.OH != dw.OH || pw.OW != dw.OW || pw.B != dw.B) return false;
    if ((dw.C & 3) || (pw.N & 3) || dw.C > 512 || pw.N > 384 || !dw.pack || dw.cpad != (dw.C + 63) / 64 * 64) return false;
    if ((((uintptr_t)dw.src) & 15) || (((uintptr_t)pw.dst) & 15)) return false;
    int dev = 0;
    if (dp_init(&dev) != K2Y_OK) return false;
    DwPwPlan plan;
    return plan_dwpw(dw, pw, dev, &plan);
}

cudaError_t launch_dwpw_tc(const DwArgs &dw, const ConvArgs &pw, const TcWeights &w, cudaStream_t st) {
    int dev = 0;
    if (dp_init(&dev) != K2Y_OK) return cudaErrorNotReady;
    DwPwPlan plan;
    if (!plan_dwpw(dw, pw, dev, &plan)) return cudaErrorInvalidConfiguration;
    DwPwParams &p = plan.p;
    p.dw_pack = dw.pack;
    p.dw_act = dw.act;
    p.dw_alpha = dw.alpha;
    p.scale = pw.scale;
    p.shift = pw.shift;
    p.act = pw.act;
    p.act_slope = pw.alpha;
    CUtensorMap map_in, map_bhi, map_blo, map_out;
    if (!make_map_nhwc(&map_in, dw.src, p.B, p.H, p.W, p.C, 32, p.win_cols, p.win_rows)) return cudaErrorInvalidValue;
    if (!make_map_nhwc(&map_out, pw.dst, p.B, p.H, p.W, p.N, 32, 8, 4)) return cudaErrorInvalidValue;
    if (!make_map_weights(&map_bhi, w.d_bh_p, (uint64_t)w.Npad, (uint64_t)w.Kpad64, (uint32_t)p.BN)) return cudaErrorInvalidValue;
    if (!make_map_weights(&map_blo, w.d_bm_p, (uint64_t)w.Npad, (uint64_t)w.Kpad64, (uint32_t)p.BN)) return cudaErrorInvalidValue;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)plan.grid);
    cfg.blockDim = dim3(DP_THREADS);
    cfg.dynamicSmemBytes = plan.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    p.trace = nullptr;
    const char *tr = getenv("K2Y_TC_TRACE");
    long long *d_trace = nullptr;
    if (tr && tr[0] == '1') {
        cudaMalloc(&d_trace, (size_t)plan.grid * 128 * sizeof(long long));
        cudaMemset(d_trace, 0, (size_t)plan.grid * 128 * sizeof(long long));
        p.trace = d_trace;
        cfg.numAttrs = 0;
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, dwpw_tc_kernel, map_in, map_bhi, map_blo, map_out, p);
    if (e != cudaSuccess) return e;
    if (d_trace) {
        cudaStreamSynchronize(st);
        long long h[128];
        fprintf(stderr, "[dwpw-trace] %dx%d C=%d N=%d tile %dx%d tiles=%d grid=%d nkb=%d BN=%d n_pass=%d acc_stages=%d a_stages=%d smem=%zu\n", p.H, p.W,
                p.C, p.N, 8 * p.tw, 4 * p.th, p.num_tiles, plan.grid, p.nkb, p.BN, p.n_pass, p.acc_stages, p.a_stages, plan.smem);
        for (int cta : {0, plan.grid - 1}) {
            cudaMemcpy(h, d_trace + (size_t)cta * 128, sizeof(h), cudaMemcpyDeviceToHost);
            auto us = [&](int i) { return h[i] ? (h[i] - h[0]) * 1e-3 : -1.0; };
            fprintf(stderr, "[dwpw-trace] cta %d: setup %.2f | win issue", cta, us(1));
            for (int i = 0; i < 6; ++i) fprintf(stderr, " %.2f", us(2 + i));
            fprintf(stderr, "\n[dwpw-trace]   kb: win_landed dw_computed tmem_stored | mma_got_A mma_issued\n");
            for (int kb = 0; kb < p.nkb && kb < 8; ++kb)
                fprintf(stderr, "[dwpw-trace]   %d: %.2f %.2f %.2f | %.2f %.2f\n", kb, us(8 + kb), us(16 + kb), us(24 + kb), us(32 + kb), us(40 + kb));
            fprintf(stderr, "[dwpw-trace]   epilogue: acc ready %.2f, tile done %.2f, stores drained %.2f, cta end %.2f\n", us(48), us(49), us(50), us(51));
            fprintf(stderr, "[dwpw-trace]   epilogue chunk cycles (3rd tile): tmem_ld+wait %lld, bn+act %lld, wait_read %lld, sts+fence %lld, tma issue %lld\n",
                    h[56], h[57], h[58], h[59], h[60]);
            fprintf(stderr, "[dwpw-trace]   depthwise unit cycles (3rd tile, window landed -> stored): %lld\n", h[61]);
            fprintf(stderr, "[dwpw-trace]   tile: win_issued dw_start dw_done mma_issued epi_start epi_done\n");
            for (int ti = 0; ti < 10; ++ti)
                fprintf(stderr, "[dwpw-trace]   %d: %.2f %.2f %.2f %.2f %.2f %.2f\n", ti, us(64 + 6 * ti + 5), us(64 + 6 * ti), us(64 + 6 * ti + 1), us(64 + 6 * ti + 2),
                        us(64 + 6 * ti + 3), us(64 + 6 * ti + 4));
        }
        cudaFree(d_trace);
    }
    return cudaGetLastError();
}

}  // namespace k2y
