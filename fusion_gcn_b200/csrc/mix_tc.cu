// Per-sample mixing over the joint axis on the tensor cores (tcgen05 / TMA), sm_100a only.
//
//   AGG_FWD : out[n][t][v][k*W + c]  = sum_u         in[n][t][u][c]       * G[n][k][u][v]       (agcn.py:110, X . G_k)
//   AGG_BWD : out[n][t][u][c]      (+)= sum_k sum_v  in[n][t][v][k*W + c] * G[n][k][u][v]       (its input gradient)
//
//   SCORE_BWD: in = e = [theta_0 phi_0 theta_1 phi_1 theta_2 phi_2] (6 groups of W channels), M = dS:
//             out[n][t][u][theta_k c] = sum_v M[n][k][u][v] * in[n][t][v][phi_k c]                 (d theta, agcn.py:104-106 backward)
//             out[n][t][v][phi_k c]   = sum_u M[n][k][u][v] * in[n][t][u][theta_k c]               (d phi)
//
// GEMM view (per timestep): the contraction index is the JOINT, the channels are the M dimension:
//   D[(t, cb, c)][col] = sum_joint A[(t, cb, c)][joint] * B[col][joint],   cb = 32-channel block, c = channel in block
// so an M = 128 tile is four (timestep, channel-block) pairs.  Activations are channels-contiguous, i.e. M-contiguous
// ("MN-major", 128B swizzle with 32-byte atoms, the same operand form as the weight-gradient kernel): one TMA box of
// 32 channels x 32 joint rows per pair (rows >= V are out of bounds of the tensor map and arrive as zeros), LBO = one box.
// B is the per-sample matrix, padded to 32 x 32 per subset by pad_mats_kernel and TMA-loaded once per sample:
//   AGG_FWD: B[N = (k, v)][K = u] = G[k][u][v]  (N contiguous: Gp [k][u][v32]),  N = 96, K = 32  (4 UMMA K steps)
//   AGG_BWD: B[N = u][K = (k, v)] = G[k][u][v]  (N contiguous: GpT[k][v][u32]),  N = 32, K = 3 x 32 (12 UMMA K steps)
//   SCORE_BWD: a tile is one 32-channel block over four consecutive timesteps (so all 128 rows share the subset k);
//            B[N = 64][K = 32] = [dS_k^T | dS_k] (two 32-column atoms): every channel row gets both mixes, and the epilogue
//            lane keeps the 32 columns that belong to it (a theta channel produces d phi and vice versa) and writes them
//            to the partner channel (+-W).  The tensor core does twice the needed MACs, which is free here (HBM bound).
// Strict fp32 mode (SPLIT): hi*hi on kind::tf32 straight from the fp32 payload (the MMA truncates it); the cross terms hi*lo + lo*hi on kind::f16 from bf16 blocks
// the converter warps build next to the operands (tf32_cross_quarter_mn, tc_common.cuh).  The K reduction is at most 96 long, so
// no accumulator promotion is needed.
// The epilogue thread of TMEM lane (pair, c) owns one channel of one timestep: for every accumulator column (an output
// joint) the 32 lanes of a warp write 32 consecutive channels, a full 128-byte line.
// Warp roles: 0 activation producer (TMA), 1 MMA issuer, 2..5 epilogue, 6 matrix producer, 7..14 operand converters (strict mode).
// Persistent: CTA c walks a contiguous range of tiles (tiles of one sample are consecutive, so B changes at most a few times).
#include "tc_common.cuh"
#include <stdlib.h>

namespace agcn {
namespace mtc {
using namespace agcn::tc;

constexpr int kMaxA = 8;
constexpr int kSplitWarps = 8;
constexpr int kThreadsM = 7 * 32;
constexpr int kThreadsMSplit = (7 + kSplitWarps) * 32;
constexpr uint32_t kBarBytes = 512;
constexpr uint32_t kBoxBytes = 4096;      // 32 rows x 128 bytes
constexpr uint32_t kMatBytes = 3u * kBoxBytes;     // AGG modes; SCORE_BWD holds 6 boxes per sample (MArgs::mat_bytes)
constexpr uint32_t kCrossBlk = 8192;      // strict mode: one bf16 block = 64 K rows ([hi16 ; lo16] of 32 joints) x 64 channels

struct MArgs {
    float* out;
    int nb, t, v, width, bwd, accumulate;
    int ncb;                 // 32-channel blocks per timestep
    int tiles_per_sample;    // ceil(t * ncb / 4)
    long long total_tiles;
    int kb;                  // K blocks per tile: 1 (fwd) or 3 (bwd)
    int ncols;               // accumulator columns: 96 (fwd) or 32 (bwd)
    int na, nlo;
    uint32_t a_tile;         // kb * 4 * kBoxBytes
    int ldout;
    int tma_out;             // 1: outputs leave through per-warp shared-memory tiles and TMA bulk stores / reduce-adds
    int score;               // SCORE_BWD
    int tgroups;             // SCORE_BWD: ceil(t / 4) timestep groups per channel block
    uint32_t mat_bytes;      // padded matrices of one sample: 3 boxes (AGG) or 6 boxes (SCORE_BWD)
    uint32_t mat_cross;      // strict mode: bytes of their bf16 [lo ; hi] blocks (8 KB each: two boxes per block)
    float* colsum_part;      // SCORE_BWD, optional: [gridDim.x * 4][ldout] per-warp column sums of `out` (the bias gradient of the theta / phi
                             // convolutions, agcn.py:104-105, which would otherwise take its own pass over the 1.5 x Cout-wide tensor)
};

constexpr int kColsumMax = 384;       // widest SCORE_BWD output with fused column sums: 6 x 64 channels

// out[c] = sum over the per-warp partials: one warp per column, lane l adds partials l, l + 32, ... and a fixed shuffle tree
// combines the lanes (deterministic)
__global__ void colsum_reduce_kernel(const float* part, int nparts, int ld, float* out) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= ld) return;
    double s = 0.0;
#pragma unroll 4
    for (int p = lane; p < nparts; p += 32) s += (double)part[(long long)p * ld + c];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) out[c] = (float)s;
}

// which (timestep, 32-channel block) the j-th quarter of tile `ts` of a sample covers
__device__ __forceinline__ void tile_pair(const MArgs& p, int ts, int j, int& tt, int& cb) {
    if (p.score) { const int tg = ts / p.ncb; cb = ts - tg * p.ncb; tt = tg * 4 + j; }
    else { const int pair = ts * 4 + j; tt = pair / p.ncb; cb = pair - tt * p.ncb; }
}

// SCORE_BWD: mats [nb][3][v][v] -> gp [nb][3][2][32][32]: box 0 = transpose ([v][u]), box 1 = as is ([u][v])
__global__ void pad_mats_score_kernel(const float* mats, float* gp, int nb, int v) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nb * 6 * 1024) return;
    const int c = idx & 31, r = (idx >> 5) & 31, box = (idx >> 10) & 1, nk = idx >> 11;
    float val = 0.f;
    if (r < v && c < v) val = box == 0 ? mats[((long long)nk * v + c) * v + r] : mats[((long long)nk * v + r) * v + c];
    gp[idx] = val;
}

// mats [nb][3][v][v] -> gp [nb][3][32][32] (zero padded): fwd keeps [k][u][v], bwd stores the transpose [k][v][u]
__global__ void pad_mats_kernel(const float* mats, float* gp, int nb, int v, int transpose) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nb * 3 * 1024) return;
    const int c = idx & 31, r = (idx >> 5) & 31, nk = idx >> 10;
    float val = 0.f;
    if (r < v && c < v) val = transpose ? mats[((long long)nk * v + c) * v + r] : mats[((long long)nk * v + r) * v + c];
    gp[idx] = val;
}

// MODE: 0 AGG_FWD, 1 AGG_BWD, 2 SCORE_BWD -- a template parameter so that each variant carries only its own epilogue
template <bool SPLIT, int MODE>
__global__ void __launch_bounds__(SPLIT ? kThreadsMSplit : kThreadsM, 1)
mix_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
              const __grid_constant__ CUtensorMap map_o, MArgs p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr bool kScore = MODE == 2, kBwd = MODE == 1;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t lo_ring = smem_base + (uint32_t)p.na * p.a_tile;                       // SPLIT: nlo slots of a_tile
    const uint32_t b_ring = lo_ring + (SPLIT ? (uint32_t)p.nlo * p.a_tile : 0u);          // 2 slots of mat_bytes (+ 2 lo slots when SPLIT)
    const uint32_t b_lo = b_ring + 2u * p.mat_bytes;
    const uint32_t bar_base = b_lo + (SPLIT ? 2u * p.mat_cross : 0u);
    auto a_full = [&](int s) { return bar_base + 8u * s; };
    auto a_empty = [&](int s) { return bar_base + 8u * (kMaxA + s); };
    auto a_lo = [&](int s) { return bar_base + 8u * (2 * kMaxA + s); };
    auto lo_empty = [&](int s) { return bar_base + 8u * (3 * kMaxA + s); };
    auto b_full = [&](int s) { return bar_base + 8u * (3 * kMaxA + 2 + s); };
    auto b_empty = [&](int s) { return bar_base + 8u * (3 * kMaxA + 4 + s); };
    auto b_lo_bar = [&](int s) { return bar_base + 8u * (3 * kMaxA + 6 + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (3 * kMaxA + 8 + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (3 * kMaxA + 10 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (3 * kMaxA + 12);
    constexpr int kCols = 256;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int s = 0; s < kMaxA; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); mbar_init(a_lo(s), kSplitWarps); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(lo_empty(s), 1); mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); mbar_init(b_lo_bar(s), kSplitWarps);
            mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(kCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    // contiguous tile range of this CTA
    const long long per = (p.total_tiles + gridDim.x - 1) / gridDim.x;
    const long long tile_begin = (long long)blockIdx.x * per;
    long long tile_end = tile_begin + per;
    if (tile_end > p.total_tiles) tile_end = p.total_tiles;
    const uint32_t a_tx = p.a_tile;

    if (warp == 0) {
        // ===================================================== activation producer (warp-uniform loop, elected lane issues)
        {
            const bool leader = elect_one_sync() != 0;
            int sa = 0; uint32_t pa = 0;
            for (long long tile = tile_begin; tile < tile_end; ++tile) {
                const int n = (int)(tile / p.tiles_per_sample);
                const int ts = (int)(tile - (long long)n * p.tiles_per_sample);
                mbar_wait(a_empty(sa), pa ^ 1u);
                const uint32_t dst = smem_base + (uint32_t)sa * p.a_tile;
                if (leader) {
                    mbar_expect_tx(a_full(sa), a_tx);
                    for (int k = 0; k < p.kb; ++k)
                        for (int j = 0; j < 4; ++j) {
                            int tt, cb;                                             // tt >= t for the pairs past the end: zero-filled box
                            tile_pair(p, ts, j, tt, cb);
                            tma_load_4d(dst + (uint32_t)(k * 4 + j) * kBoxBytes, &map_a, a_full(sa), k * p.width + cb * 32, 0, tt, n);
                        }
                }
                __syncwarp();
                if (++sa == p.na) { sa = 0; pa ^= 1u; }
            }
        }
    } else if (warp == 6) {
        // ===================================================== matrix producer: one padded (3 x 32 x 32) block per sample
        if (lane == 0) {
            int m = -1;
            int cur = -1;
            for (long long tile = tile_begin; tile < tile_end; ++tile) {
                const int n = (int)(tile / p.tiles_per_sample);
                if (n == cur) continue;
                cur = n;
                ++m;
                const int sb = m & 1;
                mbar_wait(b_empty(sb), ((uint32_t)(m >> 1) & 1u) ^ 1u);
                mbar_expect_tx(b_full(sb), p.mat_bytes);
                tma_load_4d(b_ring + (uint32_t)sb * p.mat_bytes, &map_b, b_full(sb), 0, 0, 0, n);
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        // warp-uniform control flow, one elected lane issues (uniform-register MMA operands, see conv_tc2.cu)
        {
            const bool leader = elect_one_sync() != 0;
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(p.ncols >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t idesc_bf = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                      ((uint32_t)(p.ncols >> 3) << 17) | ((128u >> 4) << 24);
            int sa = 0; uint32_t pa = 0;
            int m = -1, sb = 0;                   // m: index of the current sample in this CTA's sequence; its matrix sits in slot m & 1
            int acc = 0; uint32_t acc_phase = 0;
            int sl = 0;
            int cur = -1;
            for (long long tile = tile_begin; tile < tile_end; ++tile) {
                const int n = (int)(tile / p.tiles_per_sample);
                if (n != cur) {
                    if (m >= 0 && leader) umma_commit(b_empty(sb));   // every MMA that reads the previous matrix has been issued
                    __syncwarp();
                    cur = n;
                    ++m;
                    sb = m & 1;
                    mbar_wait(b_full(sb), (uint32_t)(m >> 1) & 1u);
                    if (SPLIT) mbar_wait(b_lo_bar(sb), (uint32_t)(m >> 1) & 1u);
                }
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                mbar_wait(a_full(sa), pa);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_u + (uint32_t)(acc * 128);
                const uint32_t abase = smem_base + (uint32_t)sa * p.a_tile;
                const uint32_t alo = lo_ring + (uint32_t)sl * p.a_tile;
                const uint32_t bbase = b_ring + (uint32_t)sb * p.mat_bytes;
                const uint32_t blo = b_lo + (uint32_t)sb * p.mat_cross;
                uint32_t score_off = 0, score_cross = 0;      // SCORE_BWD: the [dS_k^T | dS_k] box pair (and its cross block) of this tile's subset
                if (kScore) {
                    const int ts = (int)(tile - (long long)n * p.tiles_per_sample);
                    const int cb = ts % p.ncb;
                    const uint32_t subset = (uint32_t)((cb * 32) / (2 * p.width));
                    score_off = subset * 2u * kBoxBytes;
                    score_cross = subset * kCrossBlk;
                }
                if (leader) {
                    // hi*hi on kind::tf32 straight from the TMA payload (the MMA truncates the fp32 words itself) ...
                    for (int k = 0; k < p.kb; ++k) {
                        // fwd: B atoms = the three subsets (LBO = one box), K rows inside each box;  bwd: one atom, K block k = box k
                        const uint32_t bo = kScore ? score_off : (kBwd ? (uint32_t)k * kBoxBytes : 0u);
                        const uint64_t da = make_smem_desc_mn(abase + (uint32_t)k * 4u * kBoxBytes, kBoxBytes);
                        const uint64_t db = make_smem_desc_mn(bbase + bo, kBoxBytes);
#pragma unroll
                        for (int kg = 0; kg < 4; ++kg) {
                            const uint64_t ko = (uint64_t)(kg * 64);          // 8 rows = 1024 bytes, in 16-byte units
                            const uint32_t acc_flag = (k == 0 && kg == 0) ? 0u : 1u;
                            umma_tf32(d_tmem, da + ko, db + ko, idesc, acc_flag);
                        }
                    }
                    if (SPLIT) {
                        // ... then hi*lo + lo*hi as ONE chain of four K = 16 bf16 MMAs per K block over the K-stacked blocks the converter
                        // warps build: A' = [hi16 ; lo16] (two 64-channel blocks, 8 KB apart) and B' = [lo16 ; hi16]
                        mbar_wait(a_lo(sa), pa);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        for (int k = 0; k < p.kb; ++k) {
                            const uint64_t dac = make_smem_desc_mn16(alo + (uint32_t)k * 2u * kCrossBlk, kCrossBlk);
                            // bwd: subset k = the 32-column half (k & 1) of block k >> 1 (start address + 64 bytes inside the swizzled rows)
                            const uint64_t dbc = make_smem_desc_mn16(blo + (kScore ? score_cross : (kBwd ? (uint32_t)(k >> 1) * kCrossBlk + (uint32_t)(k & 1) * 64u : 0u)), kCrossBlk);
#pragma unroll
                            for (int kg = 0; kg < 4; ++kg)
                                umma_bf16(d_tmem, dac + (uint64_t)(kg * 128), dbc + (uint64_t)(kg * 128), idesc_bf, 1u);     // 16 rows = 2048 bytes
                        }
                    }
                    umma_commit(a_empty(sa));
                    if (SPLIT) umma_commit(lo_empty(sl));
                    umma_commit(tfull_bar(acc));
                }
                __syncwarp();
                if (SPLIT) { if (++sl == p.nlo) sl = 0; }
                if (++sa == p.na) { sa = 0; pa ^= 1u; }
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else if (warp < 6) {
        // ===================================================== epilogue: warp quarter q = pair q of the tile, lane = channel
        const int q = warp & 3;
        int acc = 0; uint32_t acc_phase = 0;
        // fused column sums (SCORE_BWD): every lane owns one output channel per tile; its sum over the V joints goes to the warp's
        // shared-memory accumulator of that channel (behind the TMA staging tiles)
        float* csum = nullptr;
        if constexpr (kScore) {
            if (p.colsum_part != nullptr) {
                csum = reinterpret_cast<float*>(smem_raw + (bar_base + kBarBytes + 4u * kBoxBytes - smem_u32(smem_raw))) + q * kColsumMax;
                for (int i = lane; i < kColsumMax; i += 32) csum[i] = 0.f;
                __syncwarp();
            }
        }
        for (long long tile = tile_begin; tile < tile_end; ++tile) {
            const int n = (int)(tile / p.tiles_per_sample);
            const int ts = (int)(tile - (long long)n * p.tiles_per_sample);
            int tt, cb;
            tile_pair(p, ts, q, tt, cb);
            const bool ok = tt < p.t;
            float* obase = p.out + ((long long)n * p.t + (ok ? tt : 0)) * p.v * p.ldout + cb * 32 + lane;
            mbar_wait(tfull_bar(acc), acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + (uint32_t)(acc * 128) + ((uint32_t)(q * 32) << 16);
            if (p.tma_out) {
                // ---- TMA-store epilogue: the warp transposes its 32 channels x V joints block(s) through a private shared-memory
                // tile ([joint][channel], 128-byte rows) and lane 0 hands each [V][32] box to the copy engine (reduce-add when
                // accumulating): no per-thread global stores, no read-modify-write loads, rows past V never leave the SM.
                constexpr int nbox = (kScore || kBwd) ? 1 : 3;
                const uint32_t sbuf = bar_base + kBarBytes + (uint32_t)q * (uint32_t)nbox * kBoxBytes;
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the previous tile's stores have read the tile
                __syncwarp();
                int chan0 = cb * 32;
                uint32_t col = (uint32_t)lane;
                bool is_phi = false;
                if (kScore) {
                    is_phi = (((cb * 32 + lane) / p.width) & 1) != 0;
                    if (p.width == 16) col = (uint32_t)(lane ^ 16);                      // theta | phi share the block: swap halves
                    else chan0 += ((((cb * 32) / p.width) & 1) != 0) ? -p.width : p.width;   // whole block is theta (-> +W) or phi (-> -W)
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (k < nbox) {
                        uint32_t r0[16], r1[16];
                        if constexpr (kScore) {
                            // phi channels own columns 0..31 (dS^T mix = d theta), theta channels columns 32..63 (d phi)
                            uint32_t a0[16], a1[16];
                            tmem_ld16_nowait(taddr, a0);
                            tmem_ld16_nowait(taddr + 16u, a1);
                            tmem_ld16_nowait(taddr + 32u, r0);
                            tmem_ld16_nowait(taddr + 48u, r1);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) { if (is_phi) { r0[i] = a0[i]; r1[i] = a1[i]; } }
                            if (csum != nullptr && ok) {
                                float sv[4] = {0.f, 0.f, 0.f, 0.f};         // four chains: the adds sit on the epilogue's critical path
#pragma unroll
                                for (int i = 0; i < 16; ++i) {
                                    if (i < p.v) sv[i & 1] += __uint_as_float(r0[i]);
                                    if (16 + i < p.v) sv[2 + (i & 1)] += __uint_as_float(r1[i]);
                                }
                                csum[chan0 + (int)col] += (sv[0] + sv[1]) + (sv[2] + sv[3]);      // (a lane's output channel is distinct within the warp)
                            }
                        } else {
                            tmem_ld16_nowait(taddr + (uint32_t)(k * 32), r0);
                            tmem_ld16_nowait(taddr + (uint32_t)(k * 32 + 16), r1);
                            tmem_ld_wait();
                        }
                        const uint32_t base = sbuf + (uint32_t)k * kBoxBytes + col * 4u;
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (i < p.v) asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + (uint32_t)i * 128u), "r"(r0[i]) : "memory");
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (16 + i < p.v) asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + (uint32_t)(16 + i) * 128u), "r"(r1[i]) : "memory");
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    if (ok) {
                        for (int k = 0; k < nbox; ++k) {
                            const int c0 = chan0 + k * p.width;
                            const uint32_t src = sbuf + (uint32_t)k * kBoxBytes;
                            if (p.accumulate)
                                asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                             ::"l"(&map_o), "r"(src), "r"(c0), "r"(0), "r"(tt), "r"(n) : "memory");
                            else
                                asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                             ::"l"(&map_o), "r"(src), "r"(c0), "r"(0), "r"(tt), "r"(n) : "memory");
                        }
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            } else if constexpr (kScore) {
                // this lane's input channel is a phi channel (is_phi) -> it owns d theta = columns 0..31 (dS^T mix), written to the
                // partner theta channel (ch - W); a theta channel owns d phi = columns 32..63, written to ch + W
                const int ch = cb * 32 + lane;
                const bool is_phi = ((ch / p.width) & 1) != 0;
                float* ob = obase + (is_phi ? -p.width : p.width);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t ra[16], rb[16];
                    tmem_ld16_nowait(taddr + (uint32_t)(h * 16), ra);
                    tmem_ld16_nowait(taddr + 32u + (uint32_t)(h * 16), rb);
                    tmem_ld_wait();
                    if (ok) {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (h * 16 + i < p.v) ob[(long long)(h * 16 + i) * p.ldout] = __uint_as_float(is_phi ? ra[i] : rb[i]);
                    }
                }
            } else if constexpr (!kBwd) {
                // column j = k*32 + v -> out[.., v, k*W + cb*32 + c]
                // two 16-column loads (one subset's 32 joint columns) per TMEM round trip; batching all six cost registers the
                // 480-thread 3xTF32 variant does not have (spills, 0.30 -> 0.45 ms, profiles/r2c)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    uint32_t r0[16], r1[16];
                    tmem_ld16_nowait(taddr + (uint32_t)(k * 32), r0);
                    tmem_ld16_nowait(taddr + (uint32_t)(k * 32 + 16), r1);
                    tmem_ld_wait();
                    if (ok) {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (i < p.v) obase[(long long)i * p.ldout + k * p.width] = __uint_as_float(r0[i]);
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (16 + i < p.v) obase[(long long)(16 + i) * p.ldout + k * p.width] = __uint_as_float(r1[i]);
                    }
                }
            } else {
                // column j = u -> out[.., u, cb*32 + c]
                if constexpr (!SPLIT) {
                    // both TMEM loads and all the old values (accumulate) are in flight before the first dependent add / store
                    // (TF32: 0.305 -> 0.237 ms; the 480-thread 3xTF32 variant has no registers for it and keeps the 16-column loop)
                    uint32_t rv[2][16];
                    tmem_ld16_nowait(taddr, rv[0]);
                    tmem_ld16_nowait(taddr + 16u, rv[1]);
                    float oldv[2][16];
#pragma unroll
                    for (int cg = 0; cg < 2; ++cg)
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            oldv[cg][i] = (ok && p.accumulate && cg * 16 + i < p.v) ? obase[(long long)(cg * 16 + i) * p.ldout] : 0.f;
                    tmem_ld_wait();
                    if (ok) {
#pragma unroll
                        for (int cg = 0; cg < 2; ++cg)
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (cg * 16 + i < p.v) obase[(long long)(cg * 16 + i) * p.ldout] = __uint_as_float(rv[cg][i]) + oldv[cg][i];
                    }
                } else {
#pragma unroll
                    for (int cg = 0; cg < 2; ++cg) {
                        float vals[16];
                        tmem_ld16(taddr + (uint32_t)(cg * 16), vals);
                        if (ok) {
                            float oldv[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                oldv[i] = (p.accumulate && cg * 16 + i < p.v) ? obase[(long long)(cg * 16 + i) * p.ldout] : 0.f;
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (cg * 16 + i < p.v) obase[(long long)(cg * 16 + i) * p.ldout] = vals[i] + oldv[i];
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        if (p.tma_out && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if constexpr (kScore) {
            if (csum != nullptr) {
                __syncwarp();
                float* dst = p.colsum_part + ((long long)blockIdx.x * 4 + q) * p.ldout;
                for (int i = lane; i < p.ldout; i += 32) dst[i] = csum[i];
            }
        }
    } else if (SPLIT && warp >= 7) {
        // ===================================================== operand split: activations per tile, matrices per sample
        const int tids = threadIdx.x - 7 * 32;
        int sa = 0; uint32_t pa = 0;
        int sl = 0; uint32_t pl = 0;
        int m = -1;
        int cur = -1;
        for (long long tile = tile_begin; tile < tile_end; ++tile) {
            const int n = (int)(tile / p.tiles_per_sample);
            if (n != cur) {
                cur = n;
                ++m;
                const int sb = m & 1;
                mbar_wait(b_full(sb), (uint32_t)(m >> 1) & 1u);
                // the lo slot of this matrix slot is free once the matrix slot itself was released (b_empty), which the matrix
                // producer already waited for before refilling it
                // matrices: two boxes side by side per 64-wide cross block, K-stacked [lo16 ; hi16]
                const uint32_t msrc = b_ring + (uint32_t)sb * p.mat_bytes, mdst = b_lo + (uint32_t)sb * p.mat_cross;
                const int nboxes = (int)(p.mat_bytes / kBoxBytes);
                for (int idx = tids; idx < nboxes * 128; idx += kSplitWarps * 32) {
                    const uint32_t bx = (uint32_t)idx >> 7, r = ((uint32_t)idx >> 2) & 31u, qd = (uint32_t)idx & 3u;
                    tf32_cross_quarter_mn(msrc + bx * kBoxBytes, mdst + (bx >> 1) * kCrossBlk, r, qd, (bx & 1u) * 4u + qd, 32u + r, r);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(b_lo_bar(sb));
            }
            mbar_wait(a_full(sa), pa);
            mbar_wait(lo_empty(sl), pl ^ 1u);
            {
                // activations: boxes (k, j) of the tile -> block (k, j >> 1), half j & 1, K-stacked [hi16 ; lo16]
                const uint32_t asrc = smem_base + (uint32_t)sa * p.a_tile, adst = lo_ring + (uint32_t)sl * p.a_tile;
                for (int idx = tids; idx < p.kb * 512; idx += kSplitWarps * 32) {
                    const uint32_t bx = (uint32_t)idx >> 7, r = ((uint32_t)idx >> 2) & 31u, qd = (uint32_t)idx & 3u;
                    tf32_cross_quarter_mn(asrc + bx * kBoxBytes, adst + (bx >> 1) * kCrossBlk, r, qd, (bx & 1u) * 4u + qd, r, 32u + r);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(a_lo(sa));
            if (++sa == p.na) { sa = 0; pa ^= 1u; }
            if (++sl == p.nlo) { sl = 0; pl ^= 1u; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kCols) : "memory");
    }
}

}  // namespace mtc
}  // namespace agcn

using namespace agcn;

size_t agcn_joint_mix_tc_workspace_bytes(int nb) { return (size_t)nb * 6 * 1024 * sizeof(float); }   // SCORE_BWD pads 6 boxes per sample
size_t agcn_joint_mix_tc_colsum_floats(int ldout) { return (size_t)agcn::kNumSMs * 4 * ldout; }          // per-warp column-sum partials

// Returns AGCN_ERR_UNSUPPORTED for shapes / modes outside this path (the caller then runs the FFMA kernel).
// gp: nb*3*1024 floats of scratch for the padded matrices.
// colsum (SCORE_BWD only, may be NULL): receives sum over all rows of out[., c]; colsum_part: agcn_joint_mix_tc_colsum_floats(ldout) floats
int agcn_joint_mix_tc(const float* in, const float* mats, float* out, float* gp,
                      int nb, int t, int v, int ldin, int ldout, int width, int mode, int accumulate, int split, void* stream,
                      float* colsum, float* colsum_part) {
    using namespace agcn::tc;
    using namespace agcn::mtc;
    static const bool disabled = probe_env("AGCN_MIX_SIMT") != nullptr;
    if (disabled) return AGCN_ERR_UNSUPPORTED;
    static const bool score_simt = probe_env("AGCN_MIX_SCORE_SIMT") != nullptr;
    const bool score = mode == AGCN_MIX_SCORE_BWD;
    if (mode != AGCN_MIX_AGG_FWD && mode != AGCN_MIX_AGG_BWD && !(score && !score_simt)) return AGCN_ERR_UNSUPPORTED;
    if (v > 32 || width % (score ? 16 : 32) || !aligned16(in) || !aligned16(out) || gp == nullptr || !aligned16(gp)) return AGCN_ERR_UNSUPPORTED;
    if (mode != AGCN_MIX_AGG_BWD && accumulate) return AGCN_ERR_UNSUPPORTED;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(AGCN_ERR_CUDA, "agcn_joint_mix_tc: cuTensorMapEncodeTiled is not available from the driver");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MArgs p;
    p.out = out; p.nb = nb; p.t = t; p.v = v; p.width = width; p.bwd = mode == AGCN_MIX_AGG_BWD ? 1 : 0; p.accumulate = accumulate;
    p.score = score ? 1 : 0;
    p.ncb = (score ? 6 * width : width) / 32;
    p.tgroups = (t + 3) / 4;
    p.tiles_per_sample = score ? p.ncb * p.tgroups : (t * p.ncb + 3) / 4;
    p.total_tiles = (long long)nb * p.tiles_per_sample;
    p.kb = p.bwd ? 3 : 1;
    p.ncols = score ? 64 : (p.bwd ? 32 : 96);
    p.a_tile = (uint32_t)p.kb * 4u * kBoxBytes;
    p.ldout = ldout;
    p.mat_bytes = score ? 2u * kMatBytes : kMatBytes;
    p.mat_cross = (score ? 3u : 2u) * kCrossBlk;
    static const bool no_tma_out = probe_env("AGCN_MIX_NO_TMA_STORE") != nullptr;
    p.tma_out = (!no_tma_out && ldout % 4 == 0 && (!score || width == 16 || width % 32 == 0)) ? 1 : 0;
    p.colsum_part = nullptr;
    if (colsum != nullptr) {          // fused column sums ride in the TMA-store epilogue of the score backward
        if (!score || !p.tma_out || ldout > kColsumMax || colsum_part == nullptr) return AGCN_ERR_UNSUPPORTED;
        p.colsum_part = colsum_part;
    }
    const uint32_t out_stage = (p.tma_out ? 4u * ((score || p.bwd) ? 1u : 3u) * kBoxBytes : 0u) + (p.colsum_part ? 4u * kColsumMax * 4u : 0u);
    const uint32_t fixed = 2u * p.mat_bytes + (split ? 2u * p.mat_cross : 0u) + kBarBytes + out_stage + 1024u;
    const uint32_t budget = 220u * 1024u - fixed;
    p.nlo = split ? 2 : 0;
    int na = (int)((budget - (uint32_t)p.nlo * p.a_tile) / p.a_tile);
    if (split && na < 2) { p.nlo = 1; na = (int)((budget - p.a_tile) / p.a_tile); }
    if (na > kMaxA) na = kMaxA;
    if (na < 1) return AGCN_ERR_UNSUPPORTED;
    p.na = na;
    const size_t smem = (size_t)(p.na + p.nlo) * p.a_tile + fixed;

    if (score) pad_mats_score_kernel<<<ceil_div((long long)nb * 6 * 1024, 256), 256, 0, st>>>(mats, gp, nb, v);
    else pad_mats_kernel<<<ceil_div((long long)nb * 3 * 1024, 256), 256, 0, st>>>(mats, gp, nb, v, p.bwd);
    int rc = check_launch("agcn_joint_mix_tc(pad)");
    if (rc) return rc;

    CUtensorMap map_a, map_b;
    {
        // activations: dims (c, v, t, n); box 32 channels x 32 joint rows (rows >= v are out of bounds -> zeros) x 1 x 1
        cuuint64_t dims[4] = {(cuuint64_t)ldin, (cuuint64_t)v, (cuuint64_t)t, (cuuint64_t)nb};
        cuuint64_t strides[3] = {(cuuint64_t)ldin * 4, (cuuint64_t)v * ldin * 4, (cuuint64_t)t * v * ldin * 4};
        cuuint32_t box[4] = {32u, 32u, 1u, 1u};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(in), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(AGCN_ERR_CUDA, "agcn_joint_mix_tc: cuTensorMapEncodeTiled(in) failed with %d", (int)r);
    }
    {
        // padded matrices: dims (32, 32, 3, n); one box = the three 32 x 32 blocks of a sample
        const cuuint32_t nbox = score ? 6u : 3u;
        cuuint64_t dims[4] = {32u, 32u, nbox, (cuuint64_t)nb};
        cuuint64_t strides[3] = {128u, 4096u, 4096u * nbox};
        cuuint32_t box[4] = {32u, 32u, nbox, 1u};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, gp, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(AGCN_ERR_CUDA, "agcn_joint_mix_tc: cuTensorMapEncodeTiled(mats) failed with %d", (int)r);
    }
    CUtensorMap map_o = map_a;
    if (p.tma_out) {
        // output: dims (c, v, t, n); one box = 32 channels x V joints of one timestep (dense 128-byte rows in shared memory)
        cuuint64_t dims[4] = {(cuuint64_t)ldout, (cuuint64_t)v, (cuuint64_t)t, (cuuint64_t)nb};
        cuuint64_t strides[3] = {(cuuint64_t)ldout * 4, (cuuint64_t)v * ldout * 4, (cuuint64_t)t * v * ldout * 4};
        cuuint32_t box[4] = {32u, (cuuint32_t)v, 1u, 1u};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&map_o, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(AGCN_ERR_CUDA, "agcn_joint_mix_tc: cuTensorMapEncodeTiled(out) failed with %d", (int)r);
    }
    auto launch = [&](auto kern, int threads) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)(p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs), threads, smem, st>>>(map_a, map_b, map_o, p);
        return cudaSuccess;
    };
    cudaError_t e;
    if (split) e = score ? launch(mix_tc_kernel<true, 2>, kThreadsMSplit) : p.bwd ? launch(mix_tc_kernel<true, 1>, kThreadsMSplit) : launch(mix_tc_kernel<true, 0>, kThreadsMSplit);
    else e = score ? launch(mix_tc_kernel<false, 2>, kThreadsM) : p.bwd ? launch(mix_tc_kernel<false, 1>, kThreadsM) : launch(mix_tc_kernel<false, 0>, kThreadsM);
    if (e != cudaSuccess) return fail(AGCN_ERR_CUDA, "agcn_joint_mix_tc: %s", cudaGetErrorString(e));
    int rc2 = check_launch("agcn_joint_mix_tc");
    if (rc2 || colsum == nullptr) return rc2;
    const int nparts = (int)(p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs) * 4;
    colsum_reduce_kernel<<<ceil_div(ldout, 8), 256, 0, st>>>(colsum_part, nparts, ldout, colsum);
    return check_launch("agcn_joint_mix_tc(column sums)");
}
