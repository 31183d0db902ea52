// Joint-gram reduction on the tensor cores (tcgen05 / TMA), sm_100a only.
//
//   out[n][chunk][g][u][v] = sum_{t in chunk} sum_{c<width} a[n][t][u][offa + ga*stridea + c] * b[n][t][v][offb + g*strideb + c]
//   (ga = g, or 0 when stridea == 0)
//
// used for the attention score theta^T phi (agcn.py:104-106) and for dG = X^T dZ (gradient of agcn.py:110).  The SIMT
// kernel in joint.cu is issue-bound (FMA pipe 36-43 %, ncu profiles/r1j); the arithmetic is tiny for the tensor cores, so
// here the kernel is a pure HBM stream:
//   per timestep and 32-channel K chunk one TMA box per operand lands the rows (v, ga) / (v, g) K-major (128-byte rows, 128B
//   swizzle); one UMMA M=128 x N=80 covers ALL (ga, g) pairs of the timestep at once -- D[(u,ga)][(v,g)] -- and the epilogue
//   keeps the ga == g blocks.  Rows beyond V*groups of the M=128 / N=80 operand windows are whatever lies behind the box in
//   shared memory; they only reach accumulator rows / columns that are never read.
// Strict fp32 mode (SPLIT): 8 converter warps write bf16 [hi16 | lo16] rows of the operands into a two-slot ring (the fp32 payload
// stays as it is); hi*hi runs on kind::tf32 and the cross terms lo*hi + hi*lo on kind::f16, as in conv_tc2.cu; segment promotion into TMEM master sums.
// Warp roles: 0 TMA producer, 1 MMA issuer, 2..5 epilogue, 6..13 operand split (3xTF32 only).  Grid = (nb * nchunk) CTAs.
#include "tc_common.cuh"
#include <stdlib.h>

namespace agcn {
namespace gtc {
using namespace agcn::tc;

constexpr int kRing = 8;                 // barrier slots of the stage ring
constexpr int kSplitWarps = 8;
constexpr int kThreadsG = 6 * 32;
constexpr int kThreadsGSplit = (6 + kSplitWarps) * 32;
constexpr uint32_t kBarBytes = 512;
constexpr int kSegStages = 4;            // strict mode: stages (timestep x K-chunk group) per accumulator segment (<= 64 chained MMAs)

struct GArgs {
    float* out;
    int nb, t, v, groups, ga, width, nchunk;
    int swap;               // 1 (ga == 1, the dG gram): the wide operand b (rows (v, g)) is the M operand and a (V rows) the N = 32 operand --
                            // D[(v, g)][u]; with a on M the MMA is N = 80 wide for 25 useful rows of M, 2.5x the tensor time
    int shared_tile;        // 1: theta and phi of a group share one 128-byte row (width 16): B = A tile + 64 bytes
    int kw;                 // floats per K chunk row (16 or 32)
    int nkc;                // K chunks (boxes) per operand
    int kpg;                // K chunks per stage
    int nkg;                // stages per timestep
    uint32_t a_tile, b_tile;        // bytes reserved per K chunk tile (1024-aligned)
    uint32_t a_rows_bytes, b_rows_bytes;   // bytes TMA writes per tile
    uint32_t stage_bytes;
    int stages;
    int offa, offb;
};

template <bool SPLIT>
__global__ void __launch_bounds__(SPLIT ? kThreadsGSplit : kThreadsG, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, GArgs p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t lo_ring = smem_base + (uint32_t)p.stages * p.stage_bytes;
    const uint32_t bar_base = lo_ring + (SPLIT ? 2u * p.stage_bytes : 0u) + 16384u;      // 16 KB guard: M = 128 operand windows overrun the last tile
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kRing + s); };
    auto lo_bar = [&](int s) { return bar_base + 8u * (2 * kRing + s); };
    auto lo_empty = [&](int s) { return bar_base + 8u * (3 * kRing + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (3 * kRing + 2 + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (3 * kRing + 4 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (3 * kRing + 6);
    constexpr int kCols = SPLIT ? 512 : 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); mbar_init(lo_bar(s), kSplitWarps); }
        for (int s = 0; s < 2; ++s) { mbar_init(lo_empty(s), 1); mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4); }
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

    const int n = blockIdx.x / p.nchunk, chunk = blockIdx.x % p.nchunk;
    const int t_per = (p.t + p.nchunk - 1) / p.nchunk;
    const int t0 = chunk * t_per;
    int t1 = t0 + t_per; if (t1 > p.t) t1 = p.t;
    const int nstage = (t1 > t0 ? t1 - t0 : 0) * p.nkg;                 // stages this CTA streams
    const uint32_t stage_tx = (uint32_t)p.kpg * (p.a_rows_bytes + (p.shared_tile ? 0u : p.b_rows_bytes));

    if (warp == 0) {
        {   // warp-uniform loop, elected lane issues the TMA loads (uniform-register operands)
            const bool leader = elect_one_sync() != 0;
            int stage = 0; uint32_t phase = 0;
            for (int s = 0; s < nstage; ++s) {
                const int tt = t0 + s / p.nkg, kg = s % p.nkg;
                mbar_wait(empty_bar(stage), phase ^ 1u);
                // the last K-chunk group of a timestep may hold fewer chunks
                const int nk = (p.nkc - kg * p.kpg) < p.kpg ? (p.nkc - kg * p.kpg) : p.kpg;
                const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                if (leader) {
                    mbar_expect_tx(full_bar(stage), (uint32_t)nk * (p.a_rows_bytes + (p.shared_tile ? 0u : p.b_rows_bytes)));
                    for (int k = 0; k < nk; ++k) {
                        const int c0 = (kg * p.kpg + k) * 32;
                        tma_load_5d(sa + (uint32_t)k * p.a_tile, &map_a, full_bar(stage), p.offa + c0, 0, 0, tt, n);
                        if (!p.shared_tile)
                            tma_load_5d(sa + (uint32_t)p.kpg * p.a_tile + (uint32_t)k * p.b_tile, &map_b, full_bar(stage), p.offb + c0, 0, 0, tt, n);
                    }
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
            (void)stage_tx;
        }
    } else if (warp == 1) {
        // warp-uniform control flow, one elected lane issues (uniform-register MMA operands, see conv_tc2.cu)
        {
            const bool leader = elect_one_sync() != 0;
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t ncol = p.swap ? 32u : 80u;
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((ncol >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t idesc_bf = (1u << 4) | (1u << 7) | (1u << 10) | ((ncol >> 3) << 17) | ((128u >> 4) << 24);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            int sl = 0;
            uint32_t d_tmem = tmem_u;
            uint32_t first = 1;
            const int ksteps = p.kw / 8;
            const uint32_t b_off = p.shared_tile ? 64u : (uint32_t)p.kpg * p.a_tile;      // B tile relative to the stage base
            for (int s = 0; s < nstage; ++s) {
                const int kg = s % p.nkg;
                const int nk = (p.nkc - kg * p.kpg) < p.kpg ? (p.nkc - kg * p.kpg) : p.kpg;
                mbar_wait(full_bar(stage), phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
                const uint32_t slo = lo_ring + (uint32_t)sl * p.stage_bytes;
                if (leader) {
                    // hi*hi on kind::tf32 straight from the TMA payload (the MMA truncates the fp32 words itself) ...
                    for (int k = 0; k < nk; ++k) {
                        const uint32_t ao = (uint32_t)k * p.a_tile, bo = b_off + (uint32_t)k * (p.shared_tile ? p.a_tile : p.b_tile);
                        const uint64_t da = make_smem_desc(sa + (p.swap ? bo : ao)), db = make_smem_desc(sa + (p.swap ? ao : bo));
                        for (int ks = 0; ks < ksteps; ++ks) {
                            const uint32_t fresh = (k == 0 && ks == 0) ? first : 0u;
                            umma_tf32(d_tmem, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), idesc, fresh ^ 1u);
                        }
                    }
                    if (SPLIT) {
                        // ... then, once the converter warps have built the bf16 [hi16 | lo16] rows of this stage, the two cross terms
                        // lo*hi + hi*lo on kind::f16 (K = 16 channels per MMA)
                        mbar_wait(lo_bar(stage), phase);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        for (int k = 0; k < nk; ++k) {
                            const uint32_t ao = (uint32_t)k * p.a_tile, bo = b_off + (uint32_t)k * (p.shared_tile ? p.a_tile : p.b_tile);
                            const uint64_t dalo = make_smem_desc(slo + (p.swap ? bo : ao)), dblo = make_smem_desc(slo + (p.swap ? ao : bo));
                            if (p.shared_tile) {
                                // one row = [theta hi16 | phi hi16 | theta lo16 | phi lo16], 32 bytes (2 descriptor units) each
                                umma_bf16(d_tmem, dalo + 4u, dalo + 2u, idesc_bf, 1u);
                                umma_bf16(d_tmem, dalo, dalo + 6u, idesc_bf, 1u);
                            } else {
#pragma unroll
                                for (int j = 0; j < 2; ++j) {
                                    const uint64_t ko = (uint64_t)(j * 2);
                                    umma_bf16(d_tmem, dalo + 4u + ko, dblo + ko, idesc_bf, 1u);
                                    umma_bf16(d_tmem, dalo + ko, dblo + 4u + ko, idesc_bf, 1u);
                                }
                            }
                        }
                    }
                    umma_commit(empty_bar(stage));
                    if (SPLIT) umma_commit(lo_empty(sl));
                }
                __syncwarp();
                first = 0;
                if (SPLIT) sl ^= 1;
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                if (SPLIT && ((s + 1) % kSegStages) == 0 && s + 1 < nstage) {
                    if (leader) umma_commit(tfull_bar(acc));
                    __syncwarp();
                    if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
                    mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    d_tmem = tmem_u + (uint32_t)(acc * 128);
                    first = 1;
                }
            }
            if (leader) umma_commit(tfull_bar(acc));
            __syncwarp();
        }
    } else if (warp < 6) {
        // epilogue: lane quarter q holds accumulator rows r = q*32 + lane = u*GA + ga; columns j = v*groups + g
        // (swap: rows r = v*groups + g, columns j = u)
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int u = r / p.ga, ga = r - u * p.ga;
        const int sv = r / p.groups, sg_ = r - sv * p.groups;          // swap: this row's (v, g)
        const int nseg = SPLIT ? (nstage + kSegStages - 1) / kSegStages : 1;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const uint32_t master = tmem_base + 256u + lane_base;
        float* o = p.out + ((long long)n * p.nchunk + chunk) * p.groups * p.v * p.v;
        int acc = 0; uint32_t acc_phase = 0;
        for (int sg = 0; sg < (nseg > 0 ? nseg : 1); ++sg) {
            mbar_wait(tfull_bar(acc), acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + (uint32_t)(acc * 128) + lane_base;
            const bool last = sg >= nseg - 1;
#pragma unroll
            for (int cg = 0; cg < 5; ++cg) {
                const int c = cg * 16;
                if (p.swap && cg >= 2) break;
                float vals[16];
                if (SPLIT) tmem_promote16(taddr + (uint32_t)c, master + (uint32_t)c, sg == 0, !last, vals);
                else tmem_ld16(taddr + (uint32_t)c, vals);
                if (last && p.swap) {
                    if (sv < p.v) {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (c + i < p.v) o[((long long)sg_ * p.v + (c + i)) * p.v + sv] = nstage > 0 ? vals[i] : 0.f;
                    }
                } else if (last && u < p.v) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int j = c + i;
                        const int vv = j / p.groups, g = j - vv * p.groups;
                        if (vv < p.v && (p.ga == 1 || g == ga))
                            o[((long long)g * p.v + u) * p.v + vv] = nstage > 0 ? vals[i] : 0.f;
                    }
                }
            }
            if (SPLIT && !last) tmem_st_wait();
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    } else if (SPLIT) {
        const int tidx = threadIdx.x - 6 * 32;
        int stage = 0; uint32_t phase = 0;
        int sl = 0; uint32_t pl = 0;
        for (int s = 0; s < nstage; ++s) {
            const int kg = s % p.nkg;
            const int nk = (p.nkc - kg * p.kpg) < p.kpg ? (p.nkc - kg * p.kpg) : p.kpg;
            mbar_wait(full_bar(stage), phase);
            mbar_wait(lo_empty(sl), pl ^ 1u);
            const uint32_t sa = smem_base + (uint32_t)stage * p.stage_bytes;
            const uint32_t slo = lo_ring + (uint32_t)sl * p.stage_bytes;
            // one thread per QUARTER row (8 channels): [hi16 | lo16] cross row into the lo slot (tc_common.cuh)
            const int rows_a = (int)(p.a_rows_bytes >> 7), rows_b = p.shared_tile ? 0 : (int)(p.b_rows_bytes >> 7);
            const int per_k = (rows_a + rows_b) * 4;
            for (int idx = tidx; idx < nk * per_k; idx += kSplitWarps * 32) {
                const int k = idx / per_k, rem = idx - k * per_k;
                int rr = rem >> 2;
                uint32_t off = (uint32_t)k * p.a_tile;
                if (rr >= rows_a) { rr -= rows_a; off = (uint32_t)p.kpg * p.a_tile + (uint32_t)k * p.b_tile; }
                off += (uint32_t)rr * 128u;
                tf32_cross_quarter(sa + off, slo + off, (uint32_t)(rr & 7), (uint32_t)(rem & 3));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(lo_bar(stage));
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            sl ^= 1; if (sl == 0) pl ^= 1u;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kCols) : "memory");
    }
}

}  // namespace gtc
}  // namespace agcn

using namespace agcn;

// Returns AGCN_ERR_UNSUPPORTED for shapes outside this path (the caller then runs the SIMT kernel).
int agcn_joint_gram_tc(const float* a, const float* b, float* out,
                       int nb, int t, int v, int lda, int ldb, int groups,
                       int offa, int stridea, int offb, int strideb, int width, int nchunk, int split, void* stream) {
    using namespace agcn::tc;
    using namespace agcn::gtc;
    static const bool disabled = probe_env("AGCN_GRAM_SIMT") != nullptr;
    if (disabled) return AGCN_ERR_UNSUPPORTED;
    const int ga = stridea == 0 ? 1 : groups;
    if (groups != 3 || v * groups > 80 || v * ga > 128) return AGCN_ERR_UNSUPPORTED;
    if (!(width == 16 || width % 32 == 0)) return AGCN_ERR_UNSUPPORTED;
    if ((ga > 1 && offa + width > stridea) || offb + width > strideb) return AGCN_ERR_UNSUPPORTED;   // windows must sit inside one group stride
    if (lda % 4 || ldb % 4 || offa % 4 || offb % 4 || stridea % 4 || strideb % 4 || !aligned16(a) || !aligned16(b)) return AGCN_ERR_UNSUPPORTED;
    GArgs p;
    p.shared_tile = 0;
    if (width == 16) {
        // theta_g | phi_g side by side in one 128-byte row (the score of the 64-channel units); anything else with 16-wide
        // groups stays on the SIMT kernel
        if (!(a == b && lda == ldb && ga == groups && stridea == 32 && strideb == 32 && offb == offa + 16)) return AGCN_ERR_UNSUPPORTED;
        p.shared_tile = 1;
    }
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(AGCN_ERR_CUDA, "agcn_joint_gram_tc: cuTensorMapEncodeTiled is not available from the driver");
    p.out = out; p.nb = nb; p.t = t; p.v = v; p.groups = groups; p.ga = ga; p.width = width; p.nchunk = nchunk;
    p.offa = offa; p.offb = offb;
    p.swap = (ga == 1 && v <= 32) ? 1 : 0;
    p.kw = width == 16 ? 16 : 32;
    p.nkc = width == 16 ? 1 : width / 32;
    p.kpg = p.nkc < 2 ? p.nkc : 2;
    p.nkg = (p.nkc + p.kpg - 1) / p.kpg;
    p.a_rows_bytes = (uint32_t)(v * ga) * 128u;
    p.b_rows_bytes = (uint32_t)(v * groups) * 128u;
    p.a_tile = (p.a_rows_bytes + 1023u) & ~1023u;
    p.b_tile = (p.b_rows_bytes + 1023u) & ~1023u;
    p.stage_bytes = (uint32_t)p.kpg * (p.a_tile + (p.shared_tile ? 0u : p.b_tile));
    const uint32_t budget = 200u * 1024u - (split ? 2u * p.stage_bytes : 0u);
    int stages = (int)(budget / p.stage_bytes);
    if (stages > kRing) stages = kRing;
    if (stages < 2) return AGCN_ERR_UNSUPPORTED;
    p.stages = stages;
    const size_t smem = (size_t)(stages + (split ? 2 : 0)) * p.stage_bytes + 16384 + kBarBytes + 1024;

    CUtensorMap map_a, map_b;
    auto encode = [&](CUtensorMap* m, const float* ptr, int ld, int ngrp, int gstride) -> CUresult {
        // dims (c, group, v, t, n).  With several groups the channel axis is one group stride long (c = window offset inside
        // the group's stride), so that the strides grow monotonically; a single-group operand spans the whole row.
        cuuint64_t dims[5] = {(cuuint64_t)(ngrp > 1 ? gstride : ld), (cuuint64_t)ngrp, (cuuint64_t)v, (cuuint64_t)t, (cuuint64_t)nb};
        cuuint64_t strides[4] = {(cuuint64_t)(ngrp > 1 ? gstride : ld) * 4, (cuuint64_t)ld * 4, (cuuint64_t)v * ld * 4, (cuuint64_t)t * v * ld * 4};
        cuuint32_t box[5] = {32u, (cuuint32_t)ngrp, (cuuint32_t)v, 1u, 1u};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    CUresult r = encode(&map_a, a, lda, ga, stridea);
    if (r != CUDA_SUCCESS) return fail(AGCN_ERR_CUDA, "agcn_joint_gram_tc: cuTensorMapEncodeTiled(a) failed with %d", (int)r);
    r = encode(&map_b, b, ldb, groups, strideb);
    if (r != CUDA_SUCCESS) return fail(AGCN_ERR_CUDA, "agcn_joint_gram_tc: cuTensorMapEncodeTiled(b) failed with %d", (int)r);
    {   // per call: the attribute is per device / context, a process-wide flag would skip the second GPU
        cudaError_t e = cudaFuncSetAttribute(gram_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gram_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        if (e != cudaSuccess) return fail(AGCN_ERR_CUDA, "agcn_joint_gram_tc: %s", cudaGetErrorString(e));
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (split) gram_tc_kernel<true><<<nb * nchunk, kThreadsGSplit, smem, st>>>(map_a, map_b, p);
    else gram_tc_kernel<false><<<nb * nchunk, kThreadsG, smem, st>>>(map_a, map_b, p);
    return check_launch("agcn_joint_gram_tc");
}
