// tcgen05 / TMA implicit GEMM, second generation: halo-resident activations.
//
//   y[nb][to][v][co] (+)= bias[co] + sum_tap sum_ci x[nb][ti(to,tap)][v][ci] * w[co][tap][ci]
//
// One CTA tile = tt consecutive output timesteps x all V joints (tt*V <= 128 rows = the UMMA M atom) x BN channels.
// For every 32-channel K chunk the activation rows of ALL temporal taps of the tile are brought in ONCE: a TMA box of
// tt+taps-1 timesteps (stride 1; two boxes of even / odd timesteps for the stride-2 forward gather; one box per parity
// class for the transposed stride-2 gather).  Tap `d` is then the same shared-memory tile read d*V rows further down:
// the UMMA descriptor start address is simply advanced by whole 128-byte rows.  tcgen05 applies the 128B swizzle to
// absolute shared-memory address bits, so a start address that is not a multiple of the 8-row swizzle repeat is fine
// (probed on B200 with tools/umma_rowoff_test.cu; the descriptor's base-offset field must stay 0).  Compared with
// re-loading a shifted box per tap this divides the L2->SMEM activation traffic (and, in 3xTF32 mode, the hi/lo
// operand-split work) of the 9x1 temporal conv by ~3.5.
//
// Warp roles (7 warps, +8 converter warps in the parity modes), persistent over tiles, 1 CTA / SM:
//   warp 0      activation producer: TMA boxes of one K chunk -> A ring (NA stages of payload only)
//   warp 1      MMA issuer: per (K chunk, tap): 4 x tcgen05.mma.kind::tf32 (M128, N=BN, K8); strict mode (MODE 1): these are the
//               hi*hi products (the MMA truncates the raw fp32 rows to TF32 itself), followed by 2 x 2 kind::f16 MMAs (K16) for the
//               cross terms lo*hi + hi*lo from bf16 [hi16 | lo16] rows, into their own TMEM columns where the tile is narrow enough
//   warps 2..5  epilogue: tcgen05.ld -> (+bias) -> 1x1 convs: swizzled shared-memory chunk -> TMA bulk store / reduce-add;
//               9-tap convs: per-warp staging -> full-line st.global (+old when accumulating); optional BatchNorm column sums;
//               in 3xTF32 mode also the fp32 promotion of accumulator segments
//   warp 6      weight producer: one (tap, K chunk) BN x 32 box -> B ring (NB stages); small 1x1 weights stay resident
//   warps 7..14 (strict mode) build the [bf16(x) | bf16(x - trunc_tf32(x))] cross row of every activation row once per K chunk into a
//               2-deep ring; the payload itself is left alone, so the hi*hi MMAs do not wait for them (the weights' rna_tf32 hi
//               tensor and cross rows are precomputed into the caller's workspace and TMA-loaded)
// Warps 0, 1 and 6 run their loops on all 32 lanes (warp-uniform control flow, mbarrier waits included) and let ONE elected lane
// issue the TMA / MMA / commit instructions: ptxas then keeps the descriptors, coordinates and TMEM addresses in uniform
// registers.  Under `if (lane == 0)` every UTCHMMA / UTMALDG is wrapped in an ELECT + R2UR.BROADCAST + branch waterfall that
// costs ~100 cycles per instruction and bounded every tile narrower than 256 columns (profiles/r2f, r2y).
// The achieved HBM bandwidth of these kernels is (payload bytes in flight per SM) / (~3 us loaded latency) (measured,
// profiles/r1k): the A ring therefore holds only TMA payload and is as deep as shared memory allows; the lo residuals live
// in their own two-slot ring between the split warps and the MMA issuer.
// MODE 2 (AGCN_PREC_BF16X3, the second fp32-parity mode): x = h + m + O(2^-17 x) with h = bf16(x), m = bf16(x - h); the product is
// h_a h_b + h_a m_b + m_a h_b on tcgen05 kind::f16 (BF16 operands, fp32 accumulate): three MMAs at TWICE the TF32 rate on operands of
// HALF the bytes, i.e. half the tensor time and half the shared-memory operand traffic of 3xTF32, at ~1e-5 instead of ~2e-7 unit-level
// error (both inside the 1e-4 contract).  A K chunk is then 64 channels = two 32-channel fp32 TMA boxes per box slot; the converter
// warps (one thread per activation row) rewrite them IN PLACE as one 128-byte bf16 row of h (over the first box) and of m (over the
// second), same 128B swizzle, so no extra ring exists.  Weights: bf16 h | m tensors precomputed into the caller's workspace.
// TMEM: 2 accumulator buffers of 128 (or 256) fp32 columns (epilogue of tile i overlaps the main loop of tile i+1; 3xTF32 segments
// ping-pong) + 128 columns of fp32 master sums for the 3xTF32 segment promotion (tmem_promote16).
#include "tc_common.cuh"
#include <stdlib.h>

namespace agcn {
namespace tc2 {
using namespace agcn::tc;

constexpr int kMaxTaps = 9;
constexpr int kMaxA = 8, kMaxB = 8, kMaxLo = 4;
constexpr int kSplitWarps = 8;
constexpr int kThreads2 = 7 * 32;
constexpr int kThreads2Split = (7 + kSplitWarps) * 32;
constexpr uint32_t kBarBytes = 512;
constexpr int kTmemCols2 = 512;
constexpr uint32_t kStagePitch = 144;                        // bytes per staged output row: 32 floats + 16 bytes (conflict-free 16-byte accesses)
constexpr uint32_t kStageBytes = 4u * 32u * kStagePitch;     // four epilogue warps
constexpr uint32_t kSmemBudget = 222u * 1024u;
constexpr uint32_t kTmaStageBytes = 2u * 128u * 128u;         // TMA-store epilogue: two chunk buffers of 128 rows x 32 floats
constexpr int kStatCols = 256;                               // fused BatchNorm statistics: widest output the per-warp column sums cover
constexpr uint32_t kStatWarpFloats = 2u * kStatCols + kStatCols / 2u;      // per epilogue warp: shifted sum | shifted sum of squares (fp32) | pivots (bf16)
constexpr uint32_t kStatBytes = 4u * kStatWarpFloats * 4u;

struct Tc2Args {
    float* y; const float* bias;
    PostOp post;                  // eval-mode fused tail: y = act(scale * (acc + bias) + shift + res)
    float* stat_part;             // != NULL: per-warp statistics of the written output, [gridDim.x * 4][4][cout]: sum(o - p) | sum((o - p)^2) |
                                  // p | rows, p = the first value of the column this warp saw (BatchNorm statistics, agcn_bn_finalize)
    int nb, t_out, v, cin, cout, stride, transposed, accumulate;
    int tt, bn, n_tiles_n, kchunks, tiles_t, nparity;
    long long total_tiles;
    int seg_iters;                // 3xTF32: (tap, K chunk) iterations per accumulator segment
    int acc_stride;               // TMEM columns between the two accumulator buffers (128, or 256 for tiles wider than 128)
    int dbg;                      // bring-up only (env AGCN_CONV_DEBUG): 1 skip operand split, 2 one MMA per tap, 4 skip global stores, 8 no split warps / handshake
    int na, nbst, nlo;            // ring depths (nlo: 3xTF32 lo-residual ring, 1 or 2 slots)
    int nblk;                     // activation boxes per stage
    uint32_t blk_rows_bytes;      // bytes TMA writes per box
    uint32_t blk_bytes;           // box slot (1024-aligned, >= the rows the last tap's MMA touches)
    uint32_t a_stage_bytes;       // nblk * blk_bytes  (3xTF32: the lo residuals go to a separate two-slot ring)
    uint32_t b_stage_bytes;       // bn * 128          (3xTF32: the lo copy follows)
    int tmul;                     // box time coordinate = tmul * jt * tt + blk_t0[par][box]
    int ntap[2];
    signed char tap_id[2][kMaxTaps], tap_blk[2][kMaxTaps], tap_off[2][kMaxTaps];
    int blk_t0[2][2];
    int par_val[2];
    int dual;                     // 3xTF32, bn <= 64, multi-segment: the hi*hi products and the two cross terms accumulate in SEPARATE TMEM columns
                                  // (main at +0, cross at +bn); only the main chain carries full-magnitude truncation error, so segments are 3x longer
    int w_resident;               // 1: 1x1 conv whose whole weight tile (all K chunks, hi [+ lo]) fits the weight ring: loaded ONCE per CTA, kept for every tile
    int tma_store;                // 1: epilogue stages 32-column chunks in shared memory and writes them with TMA (bulk tensor stores / reduce-adds)
};

// MODE: 0 single-pass TF32, 1 3xTF32 (fp32 parity), 2 BF16x3 (fp32 parity, bf16 triple products)
// EPI: 0 plain epilogue, 1 the eval-mode fused tail (PostOp), 2 fused BatchNorm statistics -- a template parameter so that every
// variant carries only its own epilogue registers: the 480-thread parity-mode kernels sit at the 128-register cap (with the eval
// tail folded in at run time the epilogue spilled and the output-heavy 1x1 convolutions lost 25 %, profiles/r3q; with the
// statistics code shared, the input-gradient kernels lost 10-15 %, profiles/r4g).
template <int MODE, int EPI>
__global__ void __launch_bounds__(MODE != 0 ? kThreads2Split : kThreads2, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ CUtensorMap map_blo, const __grid_constant__ CUtensorMap map_y0,
                const __grid_constant__ CUtensorMap map_y1, Tc2Args a) {
    constexpr bool POST = EPI == 1;        // eval-mode fused tail
    constexpr bool STATS = EPI == 2;       // fused BatchNorm statistics (a.stat_part != NULL)
    constexpr bool SPLIT = MODE == 1;      // 3xTF32: hi in place, lo into its own ring
    constexpr bool BF = MODE == 2;         // BF16x3: h | m in place over the two fp32 boxes of a 64-channel chunk
    constexpr bool CONV = MODE != 0;       // converter warps, [hi ; lo] weight slots, segment promotion
    constexpr int kChunkCh = BF ? 64 : kKChunk;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_slot = a.a_stage_bytes;
    const uint32_t b_slot = a.b_stage_bytes * (CONV ? 2u : 1u);
    const uint32_t lo_ring = smem_base + (uint32_t)a.na * a_slot;                       // SPLIT: two slots of a_stage_bytes
    const uint32_t b_ring = lo_ring + (SPLIT ? (uint32_t)a.nlo * a_slot : 0u);
    const uint32_t bar_base = b_ring + (uint32_t)a.nbst * b_slot;
    auto a_full = [&](int s) { return bar_base + 8u * s; };
    auto a_empty = [&](int s) { return bar_base + 8u * (kMaxA + s); };
    auto a_lo = [&](int s) { return bar_base + 8u * (2 * kMaxA + s); };
    auto b_full = [&](int s) { return bar_base + 8u * (3 * kMaxA + s); };
    auto b_empty = [&](int s) { return bar_base + 8u * (3 * kMaxA + kMaxB + s); };
    auto lo_empty = [&](int s) { return bar_base + 8u * (3 * kMaxA + 2 * kMaxB + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (3 * kMaxA + 2 * kMaxB + kMaxLo + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (3 * kMaxA + 2 * kMaxB + kMaxLo + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (3 * kMaxA + 2 * kMaxB + kMaxLo + 4);
    const uint32_t stage_base = bar_base + kBarBytes + (uint32_t)((threadIdx.x >> 5) & 3) * (32u * kStagePitch);   // epilogue staging tile of this warp's lane quarter

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        if (a.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y0) : "memory");
        for (int s = 0; s < kMaxA; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); mbar_init(a_lo(s), kSplitWarps); }
        for (int s = 0; s < kMaxB; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
        for (int s = 0; s < kMaxLo; ++s) mbar_init(lo_empty(s), 1);
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(kTmemCols2) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    // fused BatchNorm statistics: per epilogue warp, the sum and the sum of squares of (o - pivot) for every output column it has written.
    // pivot = the first value of the column the warp saw (row 0 of its slice in its first tile of that column block) rounded to bf16:
    // any value within a fraction of a percent of the channel mean removes the cancellation of E[o^2] - mean^2, and two bytes per
    // column keep the shared-memory cost of the shifted sums at 2 KB (the operand rings of the 9-tap kernels have none to spare).
    // The row counts of the partials are recomputed from the tile sequence at the end.
    const uint32_t tbuf_base = bar_base + 1024u;                   // TMA-store epilogue: two 128-row x 128-byte chunk buffers (1024-aligned)
    const uint32_t epi_end = a.tma_store ? tbuf_base + kTmaStageBytes : bar_base + kBarBytes + kStageBytes;
    float* stat_sm = reinterpret_cast<float*>(smem_raw + (epi_end - smem_u32(smem_raw))) + ((threadIdx.x >> 5) & 3) * kStatWarpFloats;

    const uint32_t a_tx = (uint32_t)a.nblk * a.blk_rows_bytes;
    const uint32_t b_tx = (uint32_t)a.bn * 128u * (CONV ? 2u : 1u);

    if (warp == 0) {
        // ===================================================== activation producer (warp-uniform loop, elected lane issues: the TMA
        // operands then stay in uniform registers, like the MMA operands)
        {
            const bool leader = elect_one_sync() != 0;
            int sa = 0; uint32_t pa = 0;
            for (long long tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
                long long r = tile / a.n_tiles_n;
                const int jt = (int)(r % a.tiles_t); r /= a.tiles_t;
                const int par = (int)(r % a.nparity);
                const int n = (int)(r / a.nparity);
                const int tb = a.tmul * jt * a.tt;
                for (int kc = 0; kc < a.kchunks; ++kc) {
                    mbar_wait(a_empty(sa), pa ^ 1u);
                    const uint32_t dst = smem_base + (uint32_t)sa * a_slot;
                    if (leader) {
                        if (BF) {
                            // 64-channel chunk = two 32-channel fp32 boxes per slot (the second only when it holds real channels)
                            const int nh = (a.cin - kc * 64 > 32) ? 2 : 1;
                            mbar_expect_tx(a_full(sa), a_tx * (uint32_t)nh);
                            for (int b = 0; b < a.nblk; ++b)
                                for (int h = 0; h < nh; ++h)
                                    tma_load_4d(dst + (uint32_t)(b * 2 + h) * a.blk_bytes, &map_a, a_full(sa), kc * 64 + h * 32, 0, tb + a.blk_t0[par][b], n);
                        } else {
                            mbar_expect_tx(a_full(sa), a_tx);
                            for (int b = 0; b < a.nblk; ++b)
                                tma_load_4d(dst + (uint32_t)b * a.blk_bytes, &map_a, a_full(sa), kc * kKChunk, 0, tb + a.blk_t0[par][b], n);
                        }
                    }
                    __syncwarp();
                    if (++sa == a.na) { sa = 0; pa ^= 1u; }
                }
            }
        }
    } else if (warp == 6) {
        // ===================================================== weight producer (same pattern)
        {
            const bool leader = elect_one_sync() != 0;
            int sb = 0; uint32_t pb = 0;
            for (long long tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
                if (a.w_resident && tile != blockIdx.x) break;            // resident weights: one load per CTA
                const int nt = (int)(tile % a.n_tiles_n);
                const int par = (int)((tile / a.n_tiles_n / a.tiles_t) % a.nparity);
                for (int kc = 0; kc < a.kchunks; ++kc) {
                    for (int i = 0; i < a.ntap[par]; ++i) {
                        mbar_wait(b_empty(sb), pb ^ 1u);
                        if (leader) {
                            mbar_expect_tx(b_full(sb), b_tx);
                            tma_load_3d(b_ring + (uint32_t)sb * b_slot, &map_b, b_full(sb), kc * kChunkCh, a.tap_id[par][i], nt * a.bn);
                            // second half of the weight slot: BF16x3 m pieces / strict mode [hi16 | lo16] cross rows -- 64 bf16 per K chunk either way
                            if (CONV) tma_load_3d(b_ring + (uint32_t)sb * b_slot + a.b_stage_bytes, &map_blo, b_full(sb), kc * 64, a.tap_id[par][i], nt * a.bn);
                        }
                        __syncwarp();
                        if (++sb == a.nbst) { sb = 0; pb ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        // All 32 lanes run the (warp-uniform) control flow and the mbarrier waits; ONE elected lane issues tcgen05.mma / commit.
        // Under `if (lane == 0)` ptxas cannot prove the descriptor / TMEM operands warp-uniform and wraps every UTCHMMA in an
        // ELECT + 5 x R2UR.BROADCAST + branch waterfall (~100 cycles per instruction, which bounded every tile narrower than 256
        // columns, profiles/r2f); in uniform control flow the operands live in uniform registers.
        {
            const bool leader = elect_one_sync() != 0;
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            constexpr uint32_t kFmt = BF ? 1u : 2u;          // operand format: 1 = BF16 (kind::f16), 2 = TF32
            const uint32_t idesc = (1u << 4) | (kFmt << 7) | (kFmt << 10) | ((uint32_t)(a.bn >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t idesc_wide = (1u << 4) | (kFmt << 7) | (kFmt << 10) | ((uint32_t)((2 * a.bn) >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t idesc_bf = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.bn >> 3) << 17) | ((128u >> 4) << 24);     // BF16 x BF16, N = bn
            int sa = 0; uint32_t pa = 0;
            int sb = 0; uint32_t pb = 0;
            int acc = 0; uint32_t acc_phase = 0;
            int sl = 0;
            const uint32_t row_bytes_v = (uint32_t)a.v * 128u;
            const uint32_t cross_off = a.dual ? (uint32_t)a.bn : 0u;      // TMEM column offset of the cross-term accumulator
            const uint32_t main_first = a.dual ? 1u : 0u;                 // dual: the hi*hi chain starts its own accumulator each segment
            for (long long tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
                const int par = (int)((tile / a.n_tiles_n / a.tiles_t) % a.nparity);
                const int ntap = a.ntap[par];
                const int iters = ntap * a.kchunks;
                const bool wait_b = !(a.w_resident && tile != blockIdx.x);
                int it = 0;
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t d_tmem = tmem_u + (uint32_t)(acc * a.acc_stride);
                uint32_t first = 1;
                for (int kc = 0; kc < a.kchunks; ++kc) {
                    mbar_wait(a_full(sa), pa);
                    // BF16x3 converts in place: wait for the converter first.  The strict mode's converter leaves the fp32 row alone
                    // (kind::tf32 truncates it itself), so its hi*hi MMAs are issued at once and only the cross terms wait.
                    if (BF && !(a.dbg & 8)) mbar_wait(a_lo(sa), pa);
                    const int ksteps = BF ? ((a.cin - kc * 64 >= 64) ? 4 : (a.cin - kc * 64) / 16) : 4;      // BF16x3: UMMA K = 16 channels
                    const uint32_t abase = smem_base + (uint32_t)sa * a_slot;
                    const uint32_t lobase = lo_ring + (uint32_t)sl * a_slot;
                    for (int i = 0; i < ntap; ++i) {
                        if (wait_b) mbar_wait(b_full(sb), pb);                     // resident weights arrive once
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t aoff = (uint32_t)a.tap_blk[par][i] * a.blk_bytes * (BF ? 2u : 1u) + (uint32_t)a.tap_off[par][i] * row_bytes_v;
                        const uint32_t aaddr = abase + aoff;
                        const uint32_t baddr = b_ring + (uint32_t)sb * b_slot;
                        const uint64_t da = make_smem_desc(aaddr), db = make_smem_desc(baddr);
                        // second operand piece: 3xTF32 lo residuals (own ring) / BF16x3 m piece (the box after the h piece)
                        const uint64_t dalo = make_smem_desc(BF ? aaddr + a.blk_bytes : lobase + aoff), dblo = make_smem_desc(baddr + a.b_stage_bytes);
                        if (leader) {
                        if (BF) {
                            // h.Wh (+ h.Wm) + m.Wh on kind::f16; dual: one 2*bn-wide MMA against [W_h ; W_m] yields main | cross columns
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (k < ksteps) {
                                    const uint64_t ko = (uint64_t)(k * 2);           // 16 bf16 = 32 bytes per K step
                                    const uint32_t fresh = (k == 0) ? first : 0u;
                                    if (a.dual) {
                                        umma_bf16(d_tmem, da + ko, db + ko, idesc_wide, fresh ^ 1u);
                                        umma_bf16(d_tmem + cross_off, dalo + ko, db + ko, idesc, 1u);
                                    } else {
                                        umma_bf16(d_tmem, dalo + ko, db + ko, idesc, fresh ^ 1u);
                                        umma_bf16(d_tmem, da + ko, dblo + ko, idesc, 1u);
                                        umma_bf16(d_tmem, da + ko, db + ko, idesc, 1u);
                                    }
                                }
                            }
                        } else if (SPLIT) {
                            // strict fp32 mode: hi*hi on kind::tf32 (four K steps of 8 channels) + the two cross terms lo*hi + hi*lo on
                            // kind::f16 from the [hi16 | lo16] rows (two K steps of 16 channels; lo half 64 bytes = 4 descriptor units
                            // along K).  a.dual: the cross terms accumulate in their own TMEM columns.
#pragma unroll
                            for (int k = 0; k < kKChunk / 8; ++k) {
                                if ((a.dbg & 2) && k) break;
                                umma_tf32(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, ((k == 0) ? first : 0u) ^ 1u);
                            }
                            if (i == 0 && !(a.dbg & 8)) {                         // the [hi16 | lo16] rows of this K chunk
                                mbar_wait(a_lo(sa), pa);
                                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            }
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const uint64_t ko = (uint64_t)(j * 2);
                                const uint32_t fresh_c = (j == 0) ? (first & main_first) : 0u;       // dual: the cross columns start fresh too
                                umma_bf16(d_tmem + cross_off, dalo + 4u + ko, dblo + ko, idesc_bf, fresh_c ^ 1u);
                                umma_bf16(d_tmem + cross_off, dalo + ko, dblo + 4u + ko, idesc_bf, 1u);
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < kKChunk / 8; ++k) {
                                if ((a.dbg & 2) && k) break;
                                umma_tf32(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, ((k == 0) ? first : 0u) ^ 1u);
                            }
                        }
                        if (!a.w_resident) umma_commit(b_empty(sb));
                        }
                        __syncwarp();
                        first = 0;
                        if (++sb == a.nbst) { sb = 0; pb ^= 1u; }
                        ++it;
                        if (CONV && (it % a.seg_iters) == 0 && it < iters) {
                            // promote the partial accumulator to the epilogue's fp32 registers, continue in the other buffer
                            if (leader) umma_commit(tfull_bar(acc));
                            __syncwarp();
                            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                            mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            d_tmem = tmem_u + (uint32_t)(acc * a.acc_stride);
                            first = 1;
                        }
                    }
                    if (leader) { umma_commit(a_empty(sa)); if (SPLIT) umma_commit(lo_empty(sl)); }
                    __syncwarp();
                    if (SPLIT) { if (++sl == a.nlo) sl = 0; }
                    if (++sa == a.na) { sa = 0; pa ^= 1u; }
                }
                if (leader) umma_commit(tfull_bar(acc));
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else if (warp < 6) {
        // ===================================================== epilogue warps (TMEM lane quarter = warp % 4)
        const int q = warp & 3;
        const int row_local = q * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        if constexpr (STATS) {
            for (int i = lane; i < (int)kStatWarpFloats; i += 32) stat_sm[i] = 0.f;
            __syncwarp();
        }
        uint32_t stat_seen = 0;                            // bit nt: this warp has already fixed its pivots of column block nt
        // (valid rows of a tile in this warp's slice: recomputed where used -- the epilogue has no registers to spare, see the POST
        // template parameter)
        uint16_t* stat_piv = reinterpret_cast<uint16_t*>(stat_sm + 2 * kStatCols);
        auto piv_load = [&](int col) {                       // four bf16 pivots -> fp32
            const uint2 u = *reinterpret_cast<const uint2*>(stat_piv + col);
            return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
        };
        auto piv_fix = [&](float4 v, int col, bool writer) { // round four first values to bf16, keep them (lanes < 8 write), return them as fp32
            const uint32_t p0 = bf16x2_rn(v.y, v.x), p1 = bf16x2_rn(v.w, v.z);
            if (writer) *reinterpret_cast<uint2*>(stat_piv + col) = make_uint2(p0, p1);
            return make_float4(__uint_as_float(p0 << 16), __uint_as_float(p0 & 0xffff0000u), __uint_as_float(p1 << 16), __uint_as_float(p1 & 0xffff0000u));
        };
        auto stat_rvalid = [&](int jt) {
            int rv = (a.tt < a.t_out - jt * a.tt ? a.tt : a.t_out - jt * a.tt) * a.v - q * 32;      // valid rows are a prefix (forward gather)
            return rv < 0 ? 0 : (rv > 32 ? 32 : rv);
        };
        uint32_t ck = 0;                                   // chunks stored so far by this CTA (TMA-store epilogue: buffer = ck & 1)
        const bool issuer = (warp == 2 && lane == 0);      // the thread that issues and tracks the bulk stores
        for (long long tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
            long long r = tile;
            const int nt = (int)(r % a.n_tiles_n); r /= a.n_tiles_n;
            const int jt = (int)(r % a.tiles_t); r /= a.tiles_t;
            const int par = (int)(r % a.nparity);
            const int n = (int)(r / a.nparity);
            const int tl = row_local / a.v, vv = row_local - tl * a.v;
            const int j = jt * a.tt + tl;
            const int to = a.transposed ? a.stride * j + a.par_val[par] : j;
            const bool row_ok = (tl < a.tt) && (to < a.t_out);
            // element offset of this lane's output row (first column of the tile), -1 for rows outside the tensor
            const long long my_off = row_ok ? (((long long)n * a.t_out + to) * a.v + vv) * a.cout + (long long)nt * a.bn : -1;
            const int iters = a.ntap[par] * a.kchunks;
            const int nseg = CONV ? (iters + a.seg_iters - 1) / a.seg_iters : 1;
            const uint32_t lane_base = (uint32_t)(q * 32) << 16;
            const uint32_t master = tmem_base + 256u + lane_base;          // 3xTF32: fp32 master sums (columns 256..383)
            for (int sg = 0; sg < nseg; ++sg) {
                mbar_wait(tfull_bar(acc), acc_phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t taddr = tmem_base + (uint32_t)(acc * a.acc_stride) + lane_base;
                const bool last = sg == nseg - 1;
                // 32 columns at a time: TMEM (lane = row, 32 consecutive columns per thread) -> per-warp staging tile in shared
                // memory (144-byte row pitch) -> read back with 8 lanes per row, so that every st.global instruction writes four
                // full 128-byte row segments.  Writing the TMEM fragment directly (one 16-byte piece of 32 different rows per
                // instruction) held the 1x1 convolutions at a third of the HBM rate (measured, profiles/r1m).
                // NOT unrolled: eight copies of this body (three accumulator flavours, two store paths, statistics) made the 3xTF32
                // kernel 229 KB of SASS (14 300 instructions; 3 950 rolled) for no measurable gain (profiles/r2l)
#pragma unroll 1
                for (int cc = 0; cc < 8; ++cc) {
                    const int c0 = cc * 32;
                    if (c0 < a.bn) {
                        const bool wide = c0 + 16 < a.bn;                  // bn is a multiple of 16: the last chunk may be 16 wide
                        float vals[32];
                        if (CONV && a.dual) {
                            tmem_promote16_dual(taddr + (uint32_t)c0, taddr + (uint32_t)(a.bn + c0), master + (uint32_t)c0, sg == 0, !last, vals);
                            if (wide) tmem_promote16_dual(taddr + (uint32_t)c0 + 16u, taddr + (uint32_t)(a.bn + c0) + 16u, master + (uint32_t)c0 + 16u, sg == 0, !last, vals + 16);
                        } else if (CONV) {
                            tmem_promote16(taddr + (uint32_t)c0, master + (uint32_t)c0, sg == 0, !last, vals);
                            if (wide) tmem_promote16(taddr + (uint32_t)c0 + 16u, master + (uint32_t)c0 + 16u, sg == 0, !last, vals + 16);
                        } else {
                            tmem_ld16(taddr + (uint32_t)c0, vals);
                            if (wide) tmem_ld16(taddr + (uint32_t)c0 + 16u, vals + 16);
                        }
                        if (last && a.tma_store) {
                            // ---- TMA-store epilogue.  The 128 epilogue threads write their rows of the 32-column chunk (+ bias) into a
                            // 128B-swizzled shared-memory tile, one elected thread hands it to the TMA unit as ONE bulk tensor store (or
                            // reduce-add when accumulating) of the [tt][v][32] box -- rows outside the tensor are clipped by the hardware,
                            // and the warps are free to fetch the next chunk while the copy engine drains the tile to HBM.
                            const uint32_t buf = tbuf_base + (ck & 1u) * (128u * 128u);
                            if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");     // the store that last read `buf` is done
                            asm volatile("bar.sync 1, 128;" ::: "memory");
                            {
                                const uint32_t srow = buf + (uint32_t)row_local * 128u;
                                const uint32_t sw = (uint32_t)(row_local & 7);
#pragma unroll
                                for (int g = 0; g < 8; ++g) {
                                    float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
                                    if (a.bias) bq = __ldg(reinterpret_cast<const float4*>(a.bias + nt * a.bn + c0 + g * 4));
                                    float4 o = make_float4(vals[g * 4] + bq.x, vals[g * 4 + 1] + bq.y, vals[g * 4 + 2] + bq.z, vals[g * 4 + 3] + bq.w);
                                    if constexpr (POST) {
                                    if (a.post.scale) {
                                        const float4 sc = __ldg(reinterpret_cast<const float4*>(a.post.scale + nt * a.bn + c0 + g * 4));
                                        const float4 sh = __ldg(reinterpret_cast<const float4*>(a.post.shift + nt * a.bn + c0 + g * 4));
                                        o = make_float4(fmaf(o.x, sc.x, sh.x), fmaf(o.y, sc.y, sh.y), fmaf(o.z, sc.z, sh.z), fmaf(o.w, sc.w, sh.w));
                                    }
                                    if (a.post.res && my_off >= 0) {          // this thread's own row: eight 16-byte pieces = one 128-byte line
                                        const float4 rq = __ldg(reinterpret_cast<const float4*>(a.post.res + my_off + c0 + g * 4));
                                        o.x += rq.x; o.y += rq.y; o.z += rq.z; o.w += rq.w;
                                    }
                                    if (a.post.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                                    }
                                    sts128(srow + (((uint32_t)g ^ sw) << 4), o);
                                }
                            }
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            asm volatile("bar.sync 1, 128;" ::: "memory");
                            if (issuer && !(a.dbg & 4)) {
                                const CUtensorMap* my = (a.transposed && a.par_val[par] == 1) ? &map_y1 : &map_y0;
                                const int cc0 = nt * a.bn + c0, ct = jt * a.tt;
                                if (a.accumulate)
                                    asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                                 ::"l"(my), "r"(buf), "r"(cc0), "r"(0), "r"(ct), "r"(n) : "memory");
                                else
                                    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                                 ::"l"(my), "r"(buf), "r"(cc0), "r"(0), "r"(ct), "r"(n) : "memory");
                                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                            } else if (issuer) {
                                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                            }
                            ++ck;
                            if constexpr (STATS) {
                                // column sums from the staged tile: lane -> 16-byte column piece (lane & 7), rows (lane >> 3) + 4 i of this warp,
                                // shifted by the warp's pivot of the column
                                const int cq = lane & 7;
                                const int col = nt * a.bn + c0 + cq * 4;
                                const int rvalid = stat_rvalid(jt);
                                if (rvalid > 0) {
                                    float4 pv;
                                    if (!((stat_seen >> nt) & 1u)) {
                                        const int r = q * 32;
                                        pv = piv_fix(lds128(buf + (uint32_t)r * 128u + (((uint32_t)cq ^ (uint32_t)(r & 7)) << 4)), col, lane < 8);
                                    } else {
                                        pv = piv_load(col);
                                    }
                                    float4 ssum = make_float4(0.f, 0.f, 0.f, 0.f), ssq = ssum;
#pragma unroll
                                    for (int i = 0; i < 8; ++i) {
                                        const int r = q * 32 + i * 4 + (lane >> 3);
                                        if (i * 4 + (lane >> 3) < rvalid) {
                                            float4 o = lds128(buf + (uint32_t)r * 128u + (((uint32_t)cq ^ (uint32_t)(r & 7)) << 4));
                                            o.x -= pv.x; o.y -= pv.y; o.z -= pv.z; o.w -= pv.w;
                                            ssum.x += o.x; ssum.y += o.y; ssum.z += o.z; ssum.w += o.w;
                                            ssq.x = fmaf(o.x, o.x, ssq.x); ssq.y = fmaf(o.y, o.y, ssq.y); ssq.z = fmaf(o.z, o.z, ssq.z); ssq.w = fmaf(o.w, o.w, ssq.w);
                                        }
                                    }
#pragma unroll
                                    for (int d = 8; d <= 16; d <<= 1) {
                                        ssum.x += __shfl_xor_sync(0xffffffffu, ssum.x, d); ssum.y += __shfl_xor_sync(0xffffffffu, ssum.y, d);
                                        ssum.z += __shfl_xor_sync(0xffffffffu, ssum.z, d); ssum.w += __shfl_xor_sync(0xffffffffu, ssum.w, d);
                                        ssq.x += __shfl_xor_sync(0xffffffffu, ssq.x, d); ssq.y += __shfl_xor_sync(0xffffffffu, ssq.y, d);
                                        ssq.z += __shfl_xor_sync(0xffffffffu, ssq.z, d); ssq.w += __shfl_xor_sync(0xffffffffu, ssq.w, d);
                                    }
                                    if (lane < 8) {
                                        float4* s0 = reinterpret_cast<float4*>(stat_sm + col);
                                        float4* s1 = reinterpret_cast<float4*>(stat_sm + kStatCols + col);
                                        float4 u0 = *s0, u1 = *s1;
                                        u0.x += ssum.x; u0.y += ssum.y; u0.z += ssum.z; u0.w += ssum.w;
                                        u1.x += ssq.x; u1.y += ssq.y; u1.z += ssq.z; u1.w += ssq.w;
                                        *s0 = u0; *s1 = u1;
                                    }
                                }
                            }
                        } else if (last && !(a.dbg & 4)) {
                            const uint32_t srow = stage_base + (uint32_t)lane * kStagePitch;
#pragma unroll
                            for (int g = 0; g < 8; ++g)
                                if (g < 4 || wide) sts128(srow + (uint32_t)g * 16u, make_float4(vals[g * 4], vals[g * 4 + 1], vals[g * 4 + 2], vals[g * 4 + 3]));
                            __syncwarp();
                            const int cq = lane & 7;                       // 16-byte column piece handled by this lane
                            const bool col_ok = cq < 4 || wide;
                            float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (a.bias && col_ok) bq = __ldg(reinterpret_cast<const float4*>(a.bias + nt * a.bn + c0 + cq * 4));
                            float4 psc = make_float4(1.f, 1.f, 1.f, 1.f), psh = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (POST && a.post.scale && col_ok) {
                                psc = __ldg(reinterpret_cast<const float4*>(a.post.scale + nt * a.bn + c0 + cq * 4));
                                psh = __ldg(reinterpret_cast<const float4*>(a.post.shift + nt * a.bn + c0 + cq * 4));
                            }
                            float4 ssum = make_float4(0.f, 0.f, 0.f, 0.f), ssq = ssum;      // column sums over this lane's 8 rows (of o - pivot)
                            float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
                            const int scol = nt * a.bn + c0 + cq * 4;
                            if (STATS && stat_rvalid(jt) > 0 && col_ok) {
                                if (!((stat_seen >> nt) & 1u)) {               // the warp's row 0 of this tile, as it is about to be written
                                    pv = lds128(stage_base + (uint32_t)cq * 16u);
                                    pv.x += bq.x; pv.y += bq.y; pv.z += bq.z; pv.w += bq.w;
                                    pv = piv_fix(pv, scol, lane < 8);
                                } else {
                                    pv = piv_load(scol);
                                }
                            }
                            // four row offsets (and, when accumulating, four old values) are fetched before the first dependent
                            // add / store, so the global-load latency is paid twice per chunk, not once per row
#pragma unroll
                            for (int half = 0; half < 2; ++half) {
                                long long offs[4];
                                float4 oldv[4];
#pragma unroll
                                for (int r4 = 0; r4 < 4; ++r4) {
                                    offs[r4] = __shfl_sync(0xffffffffu, my_off, (half * 4 + r4) * 4 + (lane >> 3));
                                    oldv[r4] = make_float4(0.f, 0.f, 0.f, 0.f);
                                    if (a.accumulate && offs[r4] >= 0 && col_ok)
                                        oldv[r4] = *reinterpret_cast<const float4*>(a.y + offs[r4] + c0 + cq * 4);
                                    else if (POST && a.post.res && offs[r4] >= 0 && col_ok)   // eval tail: residual added AFTER the affine
                                        oldv[r4] = __ldg(reinterpret_cast<const float4*>(a.post.res + offs[r4] + c0 + cq * 4));
                                }
#pragma unroll
                                for (int r4 = 0; r4 < 4; ++r4) {
                                    const int src_lane = (half * 4 + r4) * 4 + (lane >> 3);
                                    if (offs[r4] >= 0 && col_ok) {
                                        float4 o = lds128(stage_base + (uint32_t)src_lane * kStagePitch + (uint32_t)cq * 16u);
                                        if constexpr (POST) {
                                            o.x = fmaf(o.x + bq.x, psc.x, psh.x) + oldv[r4].x; o.y = fmaf(o.y + bq.y, psc.y, psh.y) + oldv[r4].y;
                                            o.z = fmaf(o.z + bq.z, psc.z, psh.z) + oldv[r4].z; o.w = fmaf(o.w + bq.w, psc.w, psh.w) + oldv[r4].w;
                                            if (a.post.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                                        } else {
                                            o.x += bq.x + oldv[r4].x; o.y += bq.y + oldv[r4].y; o.z += bq.z + oldv[r4].z; o.w += bq.w + oldv[r4].w;
                                        }
                                        *reinterpret_cast<float4*>(a.y + offs[r4] + c0 + cq * 4) = o;
                                        if constexpr (STATS) {
                                            o.x -= pv.x; o.y -= pv.y; o.z -= pv.z; o.w -= pv.w;
                                            ssum.x += o.x; ssum.y += o.y; ssum.z += o.z; ssum.w += o.w;
                                            ssq.x = fmaf(o.x, o.x, ssq.x); ssq.y = fmaf(o.y, o.y, ssq.y); ssq.z = fmaf(o.z, o.z, ssq.z); ssq.w = fmaf(o.w, o.w, ssq.w);
                                        }
                                    }
                                }
                            }
                            if constexpr (STATS) {
                                // lanes l, l+8, l+16, l+24 hold the same four columns (different rows): fixed-order butterfly, then
                                // lanes 0..7 add the warp's 32-row sums into the warp's shared-memory accumulators
#pragma unroll
                                for (int d = 8; d <= 16; d <<= 1) {
                                    ssum.x += __shfl_xor_sync(0xffffffffu, ssum.x, d); ssum.y += __shfl_xor_sync(0xffffffffu, ssum.y, d);
                                    ssum.z += __shfl_xor_sync(0xffffffffu, ssum.z, d); ssum.w += __shfl_xor_sync(0xffffffffu, ssum.w, d);
                                    ssq.x += __shfl_xor_sync(0xffffffffu, ssq.x, d); ssq.y += __shfl_xor_sync(0xffffffffu, ssq.y, d);
                                    ssq.z += __shfl_xor_sync(0xffffffffu, ssq.z, d); ssq.w += __shfl_xor_sync(0xffffffffu, ssq.w, d);
                                }
                                if (lane < 8 && col_ok && stat_rvalid(jt) > 0) {
                                    float4* s0 = reinterpret_cast<float4*>(stat_sm + scol);
                                    float4* s1 = reinterpret_cast<float4*>(stat_sm + kStatCols + scol);
                                    float4 u0 = *s0, u1 = *s1;
                                    u0.x += ssum.x; u0.y += ssum.y; u0.z += ssum.z; u0.w += ssum.w;
                                    u1.x += ssq.x; u1.y += ssq.y; u1.z += ssq.z; u1.w += ssq.w;
                                    *s0 = u0; *s1 = u1;
                                }
                            }
                            __syncwarp();
                        }
                    }
                }
                if (CONV && !last) tmem_st_wait();
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
            if (STATS && stat_rvalid(jt) > 0) { stat_seen |= 1u << nt; __syncwarp(); }      // pivots of this column block are fixed (and visible to the warp)
        }
        if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // shared memory must outlive the bulk stores reading it
        if (a.stat_part != nullptr) {
            __syncwarp();
            // rows this warp accumulated per column block: lane = column block, recomputed from the CTA's tile sequence
            float cnt = 0.f;
            for (long long tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
                const int nt = (int)(tile % a.n_tiles_n);
                const int jt = (int)((tile / a.n_tiles_n) % a.tiles_t);
                if (nt == lane) cnt += (float)stat_rvalid(jt);
            }
            float* dst = a.stat_part + ((long long)blockIdx.x * 4 + q) * 4 * a.cout;
            for (int i0 = 0; i0 < a.cout; i0 += 32) {
                const int i = i0 + lane;
                const float ci = __shfl_sync(0xffffffffu, cnt, (i < a.cout ? i : 0) / a.bn);
                if (i < a.cout) {
                    dst[i] = stat_sm[i]; dst[a.cout + i] = stat_sm[kStatCols + i];
                    dst[2 * a.cout + i] = __uint_as_float((uint32_t)stat_piv[i] << 16); dst[3 * a.cout + i] = ci;
                }
            }
        }
    } else if (BF) {
        // ===================================================== BF16x3 converter: one thread per activation row, in place
        const int tids = threadIdx.x - 7 * 32;
        const int rows = (int)(a.blk_rows_bytes >> 7);
        int sa = 0; uint32_t pa = 0;
        for (long long tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
            for (int kc = 0; kc < a.kchunks; ++kc) {
                mbar_wait(a_full(sa), pa);
                const uint32_t src = smem_base + (uint32_t)sa * a_slot;
                const bool has1 = a.cin - kc * 64 > 32;
                // one thread per row: the two-lanes-per-row variant measured 3-25 % slower on every shape (profiles/r3e)
                for (int idx = tids; idx < a.nblk * rows; idx += kSplitWarps * 32) {
                    const int b = idx >= rows ? 1 : 0;
                    const int r = idx - b * rows;
                    const uint32_t p0 = src + (uint32_t)(b * 2) * a.blk_bytes + (uint32_t)r * 128u;
                    bf16_split_row(p0, p0 + a.blk_bytes, (uint32_t)(r & 7), has1);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(a_lo(sa));
                if (++sa == a.na) { sa = 0; pa ^= 1u; }
            }
        }
    } else if (SPLIT) {
        // ===================================================== operand split of the activation stage (kSplitWarps warps)
        const int tids = threadIdx.x - 7 * 32;
        int sa = 0; uint32_t pa = 0;
        int sl = 0; uint32_t pl = 0;
        for (long long tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
            if (a.dbg & 8) break;                       // timing probe: no split warps at all (results wrong)
            for (int kc = 0; kc < a.kchunks; ++kc) {
                mbar_wait(a_full(sa), pa);
                mbar_wait(lo_empty(sl), pl ^ 1u);
                const uint32_t src = smem_base + (uint32_t)sa * a_slot;
                const uint32_t dst = lo_ring + (uint32_t)sl * a_slot;
                // one thread per QUARTER row (8 channels): [hi16 | lo16] cross row into the lo ring slot
                const int rows = (int)(a.blk_rows_bytes >> 7);
                for (int idx = tids; idx < ((a.dbg & 1) ? 0 : a.nblk * rows * 4); idx += kSplitWarps * 32) {
                    const int rr = idx >> 2;
                    const int b = rr >= rows ? 1 : 0;
                    const int r = rr - b * rows;
                    const uint32_t off = (uint32_t)b * a.blk_bytes + (uint32_t)r * 128u;
                    tf32_cross_quarter(src + off, dst + off, (uint32_t)(r & 7), (uint32_t)(idx & 3));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(a_lo(sa));
                if (++sa == a.na) { sa = 0; pa ^= 1u; }
                if (++sl == a.nlo) { sl = 0; pl ^= 1u; }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols2) : "memory");
    }
}

}  // namespace tc2
}  // namespace agcn

using namespace agcn;

// bf16 h | m pieces of the weights for the BF16x3 mode: w_split[0 .. n) = bf16(w), w_split[n .. 2n) = bf16(w - h) (16-bit elements)
static __global__ void split_weights_bf16_kernel(const float* w, uint16_t* w_split, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float v = w[i];
        const uint32_t hb = (__float_as_uint(v) + 0x8000u) & 0xFFFF0000u;
        w_split[i] = (uint16_t)(hb >> 16);
        w_split[n + i] = (uint16_t)((__float_as_uint(v - __uint_as_float(hb)) + 0x8000u) >> 16);
    }
}

// Weights of the strict fp32 mode: w_hi [cout][taps][cin] fp32 = rna_tf32(w) (kind::tf32 operand of hi*hi) and cross
// [cout][taps][nchunk][64] bf16, every 32-channel K chunk stored as one 128-byte row [bf16(hi) | bf16(w - hi)] (channels past cin: 0).
static __global__ void split_weights_cross_kernel(const float* w, float* w_hi, uint16_t* cross, long long rows, int cin, int nchunk) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // (row, chunk, channel in chunk)
    if (i >= rows * nchunk * 32) return;
    const int c = (int)(i & 31);
    const long long rc = i >> 5;
    const int kc = (int)(rc % nchunk);
    const long long row = rc / nchunk;
    const int ch = kc * 32 + c;
    uint16_t h = 0, l = 0;
    if (ch < cin) {
        const float v = w[row * cin + ch];
        const float hi = agcn::tc::tf32_rna(v);
        w_hi[row * cin + ch] = hi;
        h = (uint16_t)((__float_as_uint(hi) + 0x8000u) >> 16);
        l = (uint16_t)((__float_as_uint(v - hi) + 0x8000u) >> 16);
    }
    cross[rc * 64 + c] = h;
    cross[rc * 64 + 32 + c] = l;
}

// Returns AGCN_ERR_UNSUPPORTED for shapes outside this path (the caller then tries the next kernel).
// split: 0 single-pass TF32, 1 3xTF32 (fp32 parity mode), 2 BF16x3 (fp32 parity mode on bf16 triple products).
// stat_part != NULL (forward gather, no accumulate, cout <= 256): the epilogue also leaves [*stat_nparts][2][cout] column sums.
int agcn_conv_fwd_tc2(const float* x, const float* w, const float* bias, float* y,
                      int nb, int t_in, int t_out, int v, int cin, int cout,
                      int taps, int stride, int pad, int transposed, int accumulate, int split, float* w_split, void* stream,
                      float* stat_part, int* stat_nparts, const agcn::PostOp* post) {
    using namespace agcn::tc;
    using namespace agcn::tc2;
    static const bool disabled = probe_env("AGCN_TC_V1") != nullptr;
    static const bool no_1x1 = probe_env("AGCN_TC2_NO1X1") != nullptr;
    static const bool no_skip_parity = probe_env("AGCN_TC2_NO_SKIP_PARITY") != nullptr;
    static const int dbg = probe_env("AGCN_CONV_DEBUG") ? atoi(probe_env("AGCN_CONV_DEBUG")) : 0;
    if (disabled || (no_1x1 && taps == 1)) return AGCN_ERR_UNSUPPORTED;
    if (cin % 4 || cout % 16 || v > 128 || stride > 2 || taps > kMaxTaps) return AGCN_ERR_UNSUPPORTED;
    if (transposed && stride > 1 && taps < stride && no_skip_parity) return AGCN_ERR_UNSUPPORTED;
    if (!aligned16(x) || !aligned16(w) || !aligned16(y) || (bias && !aligned16(bias))) return AGCN_ERR_UNSUPPORTED;
    if (split && (w_split == nullptr || !aligned16(w_split) || ((long long)cout * taps * cin) % 4)) return AGCN_ERR_UNSUPPORTED;
    const bool bf = split == 2;
    if (bf && (cin % 16 || ((long long)cout * taps * cin) % 8)) return AGCN_ERR_UNSUPPORTED;      // UMMA K = 16 bf16; m tensor 16-byte aligned
    const int chunk_ch = bf ? 64 : kKChunk;
    if (stat_part != nullptr && (transposed || accumulate || cout > kStatCols || stat_nparts == nullptr)) return AGCN_ERR_UNSUPPORTED;
    // Output-channel tile.  A 3xTF32 tile whose K reduction needs more than one accumulator segment keeps fp32 master sums
    // in TMEM columns 256..383, so it is at most 128 wide; single-segment tiles (1x1 convs with cin <= 256) and the TF32
    // mode use up to 256 columns per accumulator buffer, which avoids re-loading (and re-splitting) the activations per tile.
    const int kiters = taps * ((cin + chunk_ch - 1) / chunk_ch);
    // (tap, K chunk) iterations one accumulator segment may hold without promotion.  3xTF32: 8 (the strict mode keeps the chained
    // truncating accumulations of the hi*hi products <= 96 per segment).  BF16x3 never promotes: its K reduction is at most 144
    // chained adds into the main columns (9 taps x 256 channels / 16), whose truncation (<= 144 x 2^-24, one-sided) stays below the
    // mode's own 2^-17 operand error -- and without master sums in TMEM every tile up to 128 wide takes the wide [W_h ; W_m] MMA.
    const int seg_cap = bf ? (1 << 28) : 8;
    const bool one_segment = !split || kiters <= seg_cap;
    const int bn_cap = bf ? ((taps == 1 && kiters <= 4) ? 256 : 128) : (one_segment ? 256 : 128);
    int bn = 0;
    for (int cand = bn_cap; cand >= 16; cand -= 16)
        if (cout % cand == 0) { bn = cand; break; }
    if (bn == 0 || (bn < 64 && cout > 128)) return AGCN_ERR_UNSUPPORTED;
    if (split && taps == 1 && bn > 128) {
        // Ring depth.  The [hi ; lo] weight slot of a tile wider than 128 columns is 48..64 KB: next to the epilogue buffers only ONE
        // activation stage then fits and load -> convert -> MMA serialise (measured: 768 -> 256 at 0.48 ms against 0.22 ms in the
        // TF32 mode).  Unless the whole weight tile can stay resident, take the widest tile that leaves two weight slots and two
        // activation stages (plus the 3xTF32 lo ring); the activation tile is then re-read (from L2) and re-converted once per extra
        // tile, which is why tiles that already get two stages keep their width (128 -> 192 at 192 columns beats 2 x 96, r3h).
        const uint64_t a_st = (uint64_t)(bf ? 2 : 1) * 16384u;
        const int kch = (cin + chunk_ch - 1) / chunk_ch;
        auto fits = [&](int cand) {
            const uint64_t b = (uint64_t)cand * 256u;
            const bool resident = cout == cand && kch <= kMaxB && (uint64_t)kch * b + (split == 1 ? 5u : 3u) * a_st <= 176u * 1024u;
            return resident || 2u * b + (split == 1 ? 4u : 2u) * a_st <= 176u * 1024u;
        };
        if (!fits(bn))
            for (int cand = 128; cand >= 64; cand -= 16)
                if (cout % cand == 0 && fits(cand)) { bn = cand; break; }
    }
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(AGCN_ERR_CUDA, "agcn_conv_fwd_tc2: cuTensorMapEncodeTiled is not available from the driver");

    Tc2Args a;
    a.y = y; a.bias = bias; a.stat_part = stat_part;
    a.post = post ? *post : PostOp{nullptr, nullptr, nullptr, 0};
    if (post && (accumulate || transposed || stat_part)) return AGCN_ERR_UNSUPPORTED;
    a.nb = nb; a.t_out = t_out; a.v = v; a.cin = cin; a.cout = cout; a.stride = stride; a.transposed = transposed; a.accumulate = accumulate;
    a.dbg = dbg;
    static const bool no_dual = probe_env("AGCN_TC2_NO_DUAL") != nullptr;
    // main | cross accumulator pairs: multi-segment tiles up to 64 wide (2*bn <= 128 next to the master sums), single-segment
    // tiles (1x1 convs with cin <= 256: no master sums) up to 128 wide (2*bn <= 256 = one of the two accumulator buffers)
    a.dual = (split && !no_dual && ((kiters > seg_cap && bn <= 64) || (kiters <= seg_cap && bn <= 128))) ? 1 : 0;
    // strict mode: an iteration chains 4 hi*hi MMAs (+ 4 cross MMAs when they share the accumulator): 8 iterations = 64 adds per
    // segment without the dual columns, 12 iterations = 48 main adds with them
    a.seg_iters = (split && kiters <= seg_cap) ? seg_cap : (a.dual ? 3 * kSegment : 2 * kSegment);
    a.acc_stride = (bn > 128 || (a.dual && 2 * bn > 128)) ? 256 : 128;
    a.tt = 128 / v;
    a.bn = bn;
    a.n_tiles_n = cout / bn;
    a.kchunks = (cin + chunk_ch - 1) / chunk_ch;
    a.nparity = transposed ? stride : 1;
    const int t_per_class = transposed ? (t_out + stride - 1) / stride : t_out;
    a.tiles_t = (t_per_class + a.tt - 1) / a.tt;
    a.total_tiles = (long long)nb * a.nparity * a.tiles_t * a.n_tiles_n;

    // ---- tap tables: which box and which timestep offset inside it every (parity, tap) reads
    const int es = transposed ? 1 : stride;          // element stride of the activation box along T
    a.tmul = transposed ? 1 : stride;
    int nt_max = 0;                                  // timesteps per box
    a.nblk = 1;
    for (int p = 0; p < 2; ++p) { a.ntap[p] = 0; a.blk_t0[p][0] = a.blk_t0[p][1] = 0; a.par_val[p] = p; }
    bool skipped_parity = false;
    if (!transposed) {
        a.nblk = (stride == 2 && taps > 1) ? 2 : 1;
        for (int tap = 0; tap < taps; ++tap) {
            const int b = a.nblk == 2 ? (tap & 1) : 0;
            const int off = a.nblk == 2 ? tap / 2 : (stride == 1 ? tap : 0);
            const int i = a.ntap[0]++;
            a.tap_id[0][i] = (signed char)tap; a.tap_blk[0][i] = (signed char)b; a.tap_off[0][i] = (signed char)off;
            if (a.tt + off > nt_max) nt_max = a.tt + off;
        }
        if (stride == 2 && taps == 1) { /* 1x1 strided: single box, offset 0 */ }
        a.blk_t0[0][0] = -pad;
        a.blk_t0[0][1] = 1 - pad;
        if (stride != 1 && stride != 2 && taps > 1) return AGCN_ERR_UNSUPPORTED;
    } else {
        int ncls = 0;
        for (int pv = 0; pv < stride; ++pv) {
            int qmin = 1 << 30, qmax = -(1 << 30);
            for (int tap = 0; tap < taps; ++tap) {
                const int num = pv + pad - tap;
                if (num % stride) continue;
                const int q = num / stride;
                if (q < qmin) qmin = q;
                if (q > qmax) qmax = q;
            }
            if (qmin > qmax) {                                     // no tap reaches this parity (strided 1x1 residual conv): its
                skipped_parity = true;                             // output rows receive no contribution
                continue;
            }
            const int par = ncls++;
            a.par_val[par] = pv;
            for (int tap = 0; tap < taps; ++tap) {
                const int num = pv + pad - tap;
                if (num % stride) continue;
                const int i = a.ntap[par]++;
                a.tap_id[par][i] = (signed char)tap; a.tap_blk[par][i] = 0; a.tap_off[par][i] = (signed char)(num / stride - qmin);
            }
            a.blk_t0[par][0] = qmin;
            if (a.tt + qmax - qmin > nt_max) nt_max = a.tt + qmax - qmin;
        }
        if (ncls == 0) return AGCN_ERR_UNSUPPORTED;
        if (ncls != a.nparity) {
            if (no_skip_parity) return AGCN_ERR_UNSUPPORTED;
            a.nparity = ncls;
            a.total_tiles = (long long)nb * a.nparity * a.tiles_t * a.n_tiles_n;
        }
    }
    if (nt_max * es > 256) return AGCN_ERR_UNSUPPORTED;
    a.blk_rows_bytes = (uint32_t)nt_max * v * 128u;
    uint32_t need = ((uint32_t)(nt_max - a.tt) * v + 128u) * 128u;       // rows the last tap's M=128 operand touches
    if (need < a.blk_rows_bytes) need = a.blk_rows_bytes;
    a.blk_bytes = (need + 1023u) & ~1023u;
    a.a_stage_bytes = (uint32_t)a.nblk * a.blk_bytes * (bf ? 2u : 1u);      // BF16x3: two 32-channel fp32 boxes per slot -> h | m in place
    a.b_stage_bytes = (uint32_t)bn * 128u;
    const uint32_t a_slot = a.a_stage_bytes, b_slot = a.b_stage_bytes * (split ? 2u : 1u);
    const uint32_t stat_bytes = stat_part != nullptr ? kStatBytes : 0u;
    // TMA-store epilogue for the 1x1 convolutions (their time is the output stream: skipping the stores halves it, profiles/r2b);
    // the 9-tap kernels are bound elsewhere and keep their shared memory for the operand rings
    static const bool no_tma_store = probe_env("AGCN_TC2_NO_TMA_STORE") != nullptr;
    static const bool tma_store_all = probe_env("AGCN_TC2_TMA_STORE_ALL") != nullptr;      // experiment: also the 9-tap kernels
    a.tma_store = (!no_tma_store && (taps == 1 || tma_store_all) && bn % 32 == 0 && aligned16(y)) ? 1 : 0;
    const uint32_t epi_bytes = a.tma_store ? 1024u + kTmaStageBytes : kBarBytes + kStageBytes;
    const uint32_t budget = kSmemBudget - 1024u - epi_bytes - stat_bytes;
    // Ring depths.  Weights: 3 slots (2 when tight).  3xTF32 lo residuals: 2 slots, 1 when two would leave a single
    // activation stage.  Everything else goes to the activation ring: payload bytes in flight set the achieved bandwidth.
    static const int nlo_env = probe_env("AGCN_TC2_NLO") ? atoi(probe_env("AGCN_TC2_NLO")) : 0;
    // lo-residual ring: 2 slots (3 or 4 measured no faster for the 1x1 kernels and slower where they shorten the payload ring, profiles/r2g)
    int nlo_want = split == 1 ? 2 : 0;
    if (split == 1 && nlo_env >= 1 && nlo_env <= kMaxLo) nlo_want = nlo_env;
    a.na = 0; a.nbst = 0; a.nlo = nlo_want;
    auto fit = [&](int nlo, int min_b) -> int {
        const uint64_t fixed = (uint64_t)nlo * a_slot + (uint64_t)min_b * b_slot;
        if (fixed + a_slot > budget) return 0;
        int na = (int)((budget - fixed) / a_slot);
        return na > kMaxA ? kMaxA : na;
    };
    int best_na = 0, best_lo = a.nlo, best_b = 3;
    for (int nlo = a.nlo; nlo >= (split == 1 ? 1 : 0) && best_na < 2; --nlo)
        for (int min_b = 3; min_b >= 2 && best_na < 2; --min_b) {
            const int na = fit(nlo, min_b);
            if (na > best_na) { best_na = na; best_lo = nlo; best_b = min_b; }
        }
    if (best_na == 0) return AGCN_ERR_UNSUPPORTED;
    a.na = best_na; a.nlo = best_lo;
    {
        int nbst = (int)((budget - (uint64_t)a.nlo * a_slot - (uint64_t)a.na * a_slot) / b_slot);
        if (nbst > kMaxB) nbst = kMaxB;
        if (nbst < best_b) nbst = best_b;
        a.nbst = nbst;
    }
    // Resident weights: a 1x1 conv with a single output-channel tile re-reads the SAME weight boxes for every 125-row tile (for
    // 64 -> 192 channels in 3xTF32 that is 98 KB of weights per 32 KB of activations through the SM's ingress).  When all K chunks
    // fit the weight ring next to at least three activation stages they are loaded once per CTA and never released.
    static const bool no_resident = probe_env("AGCN_TC2_NO_RESIDENT_W") != nullptr;
    a.w_resident = 0;
    if (!no_resident && taps == 1 && a.n_tiles_n == 1 && a.nparity == 1 && a.kchunks <= kMaxB) {
        const uint64_t wbytes = (uint64_t)a.kchunks * b_slot;
        const uint64_t fixed_lo = (uint64_t)a.nlo * a_slot;
        if (wbytes + fixed_lo + 3ull * a_slot <= budget) {
            a.w_resident = 1;
            a.nbst = a.kchunks;
            int na = (int)((budget - wbytes - fixed_lo) / a_slot);
            a.na = na > kMaxA ? kMaxA : na;
        }
    }
    const size_t smem = (size_t)(a.na + a.nlo) * a_slot + (size_t)a.nbst * b_slot + 1024 + epi_bytes + stat_bytes;

    CUtensorMap map_a, map_b, map_blo;
    {
        cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)v, (cuuint64_t)t_in, (cuuint64_t)nb};
        cuuint64_t strides[3] = {(cuuint64_t)cin * 4, (cuuint64_t)v * cin * 4, (cuuint64_t)t_in * v * cin * 4};
        cuuint32_t box[4] = {(cuuint32_t)kKChunk, (cuuint32_t)v, (cuuint32_t)(nt_max * es), 1};
        cuuint32_t estr[4] = {1, 1, (cuuint32_t)es, 1};
        CUresult r = enc(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(AGCN_ERR_CUDA, "agcn_conv_fwd_tc2: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        // weight boxes: bn rows x one 128-byte K chunk (32 fp32 or 64 bf16 channels; channels past cin are zero-filled)
        const size_t esz = bf ? 2 : 4;
        auto encode_w = [&](CUtensorMap* m, const void* ptr) -> CUresult {
            cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)taps, (cuuint64_t)cout};
            cuuint64_t strides[2] = {(cuuint64_t)cin * esz, (cuuint64_t)taps * cin * esz};
            cuuint32_t box[3] = {(cuuint32_t)chunk_ch, 1, (cuuint32_t)bn};
            cuuint32_t estr[3] = {1, 1, 1};
            return enc(m, bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        };
        const long long nw = (long long)cout * taps * cin;
        CUresult r = encode_w(&map_b, split ? static_cast<const void*>(w_split) : static_cast<const void*>(w));
        if (r != CUDA_SUCCESS) return fail(AGCN_ERR_CUDA, "agcn_conv_fwd_tc2: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
        map_blo = map_b;
        if (bf) {
            r = encode_w(&map_blo, static_cast<const void*>(reinterpret_cast<uint16_t*>(w_split) + nw));
            if (r != CUDA_SUCCESS) return fail(AGCN_ERR_CUDA, "agcn_conv_fwd_tc2: cuTensorMapEncodeTiled(B lo) failed with %d", (int)r);
            split_weights_bf16_kernel<<<ceil_div(nw, 256), 256, 0, st>>>(w, reinterpret_cast<uint16_t*>(w_split), nw);
            int rc = check_launch("agcn_conv_fwd_tc2(split weights)");
            if (rc) return rc;
        } else if (split) {
            // strict mode: cross rows [hi16 | lo16] of every 32-channel K chunk behind the fp32 hi tensor
            const int nchunk = (cin + kKChunk - 1) / kKChunk;
            uint16_t* cross = reinterpret_cast<uint16_t*>(w_split + nw);
            cuuint64_t dims[3] = {(cuuint64_t)nchunk * 64, (cuuint64_t)taps, (cuuint64_t)cout};
            cuuint64_t strides[2] = {(cuuint64_t)nchunk * 128, (cuuint64_t)taps * nchunk * 128};
            cuuint32_t box[3] = {64u, 1u, (cuuint32_t)bn};
            cuuint32_t estr[3] = {1, 1, 1};
            r = enc(&map_blo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, cross, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(AGCN_ERR_CUDA, "agcn_conv_fwd_tc2: cuTensorMapEncodeTiled(B cross) failed with %d", (int)r);
            const long long n = (long long)cout * taps * nchunk * 32;
            split_weights_cross_kernel<<<ceil_div(n, 256), 256, 0, st>>>(w, w_split, cross, (long long)cout * taps, cin, nchunk);
            int rc = check_launch("agcn_conv_fwd_tc2(split weights)");
            if (rc) return rc;
        }
    }
    CUtensorMap map_y0 = map_a, map_y1 = map_a;
    if (a.tma_store) {
        // output tile box: 32 channels x V joints x tt timesteps of one sample.  Transposed stride-s gather: one map per output-time
        // parity (base shifted by `parity` timesteps, time stride s) so that a tile's rows are dense in the map's coordinates.
        const int tstep = transposed ? stride : 1;
        for (int cls = 0; cls < a.nparity; ++cls) {
            const int pv = transposed ? a.par_val[cls] : 0;
            const long long nt_cls = transposed ? (t_out - pv + stride - 1) / stride : t_out;
            cuuint64_t dims[4] = {(cuuint64_t)cout, (cuuint64_t)v, (cuuint64_t)nt_cls, (cuuint64_t)nb};
            cuuint64_t strides[3] = {(cuuint64_t)cout * 4, (cuuint64_t)tstep * v * cout * 4, (cuuint64_t)t_out * v * cout * 4};
            cuuint32_t box[4] = {32u, (cuuint32_t)v, (cuuint32_t)a.tt, 1u};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            CUresult r = enc(pv == 1 ? &map_y1 : &map_y0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, y + (long long)pv * v * cout, dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(AGCN_ERR_CUDA, "agcn_conv_fwd_tc2: cuTensorMapEncodeTiled(Y) failed with %d", (int)r);
        }
    }
    {   // per call: the attribute is per device / context, a process-wide flag would skip the second GPU
        auto opt_in = [](auto kern) { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget); };
        const int epi = post ? 1 : (stat_part != nullptr ? 2 : 0);
        cudaError_t e = cudaSuccess;
        if (epi == 0) { e = opt_in(conv_tc2_kernel<0, 0>); if (e == cudaSuccess) e = opt_in(conv_tc2_kernel<1, 0>); if (e == cudaSuccess) e = opt_in(conv_tc2_kernel<2, 0>); }
        if (epi == 1) { e = opt_in(conv_tc2_kernel<0, 1>); if (e == cudaSuccess) e = opt_in(conv_tc2_kernel<1, 1>); if (e == cudaSuccess) e = opt_in(conv_tc2_kernel<2, 1>); }
        if (epi == 2) { e = opt_in(conv_tc2_kernel<0, 2>); if (e == cudaSuccess) e = opt_in(conv_tc2_kernel<1, 2>); if (e == cudaSuccess) e = opt_in(conv_tc2_kernel<2, 2>); }
        if (e != cudaSuccess) return fail(AGCN_ERR_CUDA, "agcn_conv_fwd_tc2: %s", cudaGetErrorString(e));
    }
    if (skipped_parity && !accumulate) {
        // the kernel only visits the parity classes that have taps; the other output timesteps are exact zeros
        cudaError_t e = cudaMemsetAsync(y, 0, (size_t)nb * t_out * v * cout * sizeof(float), st);
        if (e != cudaSuccess) return fail(AGCN_ERR_CUDA, "agcn_conv_fwd_tc2: cudaMemsetAsync: %s", cudaGetErrorString(e));
    }
    const long long grid = a.total_tiles < kNumSMs ? a.total_tiles : kNumSMs;
    // EPI: 0 plain, 1 eval tail (PostOp), 2 fused BatchNorm statistics -- template parameters, so that every variant carries only its
    // own epilogue code (the 480-thread parity-mode kernels sit at their 128-register cap; see the note at the template)
    auto launch = [&](auto k0, auto k1, auto k2) {
        if (split == 2) k2<<<(unsigned)grid, kThreads2Split, smem, st>>>(map_a, map_b, map_blo, map_y0, map_y1, a);
        else if (split) k1<<<(unsigned)grid, kThreads2Split, smem, st>>>(map_a, map_b, map_blo, map_y0, map_y1, a);
        else k0<<<(unsigned)grid, kThreads2, smem, st>>>(map_a, map_b, map_blo, map_y0, map_y1, a);
    };
    if (post) launch(conv_tc2_kernel<0, 1>, conv_tc2_kernel<1, 1>, conv_tc2_kernel<2, 1>);
    else if (stat_part != nullptr) launch(conv_tc2_kernel<0, 2>, conv_tc2_kernel<1, 2>, conv_tc2_kernel<2, 2>);
    else launch(conv_tc2_kernel<0, 0>, conv_tc2_kernel<1, 0>, conv_tc2_kernel<2, 0>);
    if (stat_nparts != nullptr) *stat_nparts = stat_part != nullptr ? (int)grid * 4 : 0;
    return check_launch("agcn_conv_fwd_tc2");
}
