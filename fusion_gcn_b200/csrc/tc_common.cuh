// tcgen05 / TMA / mbarrier PTX wrappers shared by the tensor-core kernels (sm_100a only).
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace agcn {
namespace tc {

constexpr int kKChunk = 32;       // fp32 elements per 128-byte swizzle row
constexpr int kTmemCols = 256;
constexpr int kSegment = 4;       // 3xTF32: promote the TMEM accumulator to fp32 registers every 4 (tap, k-chunk) steps (K = 128)

// round-to-nearest (ties away) to the 10-bit TF32 mantissa, i.e. what cvt.rna.tf32.f32 returns -- done with two full-rate
// integer instructions (add half an ulp, clear the low 13 bits) because the conversion pipe runs at a quarter of that rate and
// was a measurable share of the operand-split warps' time (profiles/r1y)
__device__ __forceinline__ float tf32_rna(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// 3xTF32 operand split of `bytes` bytes at src: hi = rna_tf32(x) overwrites src in place, lo = x - hi (exact) goes to dst
// at the same offsets, so any swizzle is preserved.  Called by the 128 transform threads.
__device__ __forceinline__ void transform_split(uint32_t src, uint32_t dst, uint32_t bytes, int tid128) {
    for (uint32_t off = (uint32_t)tid128 * 16u; off < bytes; off += 128u * 16u) {
        const float4 v = lds128(src + off);
        const float4 hi = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
        sts128(src + off, hi);
        sts128(dst + off, make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}


// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// 1 in exactly one (elected) lane of a fully converged warp
__device__ __forceinline__ uint32_t elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// raw (no wait) 16-column TMEM load / store; the caller batches them and issues one wait
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                   "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                   "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                   "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 3xTF32 accumulator promotion inside TMEM.  The tensor core's fp32 accumulate truncates, so a long K reduction is cut into
// short segments; each finished segment P (16 columns at `part`) is added in fp32 registers (round to nearest) to the master
// sum S (16 columns at `master`) which also lives in TMEM:  first segment S = P, later S += P.  Returns the new S in v.
__device__ __forceinline__ void tmem_promote16(uint32_t part, uint32_t master, bool first, bool write_back, float* v) {
    uint32_t p[16], s[16];
    tmem_ld16_nowait(part, p);
    if (!first) tmem_ld16_nowait(master, s);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = first ? __uint_as_float(p[i]) : __uint_as_float(s[i]) + __uint_as_float(p[i]);
    if (write_back) tmem_st16(master, v);
}

// Same with the segment's products split over two accumulators: `part` holds the hi*hi chain, `cross` the lo*hi + hi*lo terms.
__device__ __forceinline__ void tmem_promote16_dual(uint32_t part, uint32_t cross, uint32_t master, bool first, bool write_back, float* v) {
    {   // two round trips keep at most 32 loaded registers alive (the 480-thread kernels have no room for 48)
        uint32_t p[16], c[16];
        tmem_ld16_nowait(part, p);
        tmem_ld16_nowait(cross, c);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(p[i]) + __uint_as_float(c[i]);
    }
    if (!first) {
        uint32_t s[16];
        tmem_ld16_nowait(master, s);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += __uint_as_float(s[i]);
    }
    if (write_back) tmem_st16(master, v);
}

// Operand split of `bytes` bytes (multiple of 16) at src for the 3xTF32 path: hi = rna_tf32(x) overwrites src in place,
// lo = x - hi (exact) goes to dst at the same offsets (any swizzle is preserved).  `nthreads` threads cooperate; every
// thread keeps four 16-byte loads in flight before the dependent converts and stores.
// (Tried twice and dropped: leaving src untouched -- the tensor core truncates the raw operand itself -- and writing only
// lo = x - trunc(x).  Without rounding lo the error is one-sided (2^-20) and stage tolerances fail (profiles/r2a); with
// lo = rna_tf32(x - trunc(x)) parity holds but no kernel gets faster, several get 5-15 % slower (profiles/r2i).)
__device__ __forceinline__ void transform_split4(uint32_t src, uint32_t dst, uint32_t bytes, int tid, int nthreads) {
    const uint32_t step = (uint32_t)nthreads * 16u;
    for (uint32_t off = (uint32_t)tid * 16u; off < bytes; off += 4u * step) {
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t o = off + (uint32_t)i * step;
            v[i] = (o < bytes) ? lds128(src + o) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t o = off + (uint32_t)i * step;
            if (o < bytes) {
                const float4 hi = make_float4(tf32_rna(v[i].x), tf32_rna(v[i].y), tf32_rna(v[i].z), tf32_rna(v[i].w));
                sts128(src + o, hi);
                sts128(dst + o, make_float4(v[i].x - hi.x, v[i].y - hi.y, v[i].z - hi.z, v[i].w - hi.w));
            }
        }
    }
}

// ---------------------------------------------------------------- BF16x3 (AGCN_PREC_BF16X3) helpers
// round-to-nearest (ties away) bf16 pieces of eight fp32 values: h = bf16(x), m = bf16(x - h), packed in channel order
__device__ __forceinline__ void bf16_split8(const float4& a, const float4& b, uint4& h, uint4& m) {
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t hb[8], mb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        hb[i] = (__float_as_uint(x[i]) + 0x8000u) & 0xFFFF0000u;
        mb[i] = __float_as_uint(x[i] - __uint_as_float(hb[i])) + 0x8000u;          // exact difference; only the upper half is kept below
    }
    h = make_uint4(__byte_perm(hb[0], hb[1], 0x7632), __byte_perm(hb[2], hb[3], 0x7632), __byte_perm(hb[4], hb[5], 0x7632), __byte_perm(hb[6], hb[7], 0x7632));
    m = make_uint4(__byte_perm(mb[0], mb[1], 0x7632), __byte_perm(mb[2], mb[3], 0x7632), __byte_perm(mb[4], mb[5], 0x7632), __byte_perm(mb[6], mb[7], 0x7632));
}
__device__ __forceinline__ void sts128u(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// One activation row of a 64-channel K chunk, in place: p0 / p1 = the row in the first / second 32-channel fp32 box (128 bytes each,
// 16-byte chunk j stored at j ^ sw), rewritten as 64 bf16 of h at p0 and 64 bf16 of m at p1 (chunk q = channels 8q..8q+7 at q ^ sw).
// Every fp32 chunk is read before the bf16 chunk that overwrites its bytes is stored.
__device__ __forceinline__ void bf16_split_row(uint32_t p0, uint32_t p1, uint32_t sw, bool has1) {
    float4 f[8];
    uint4 hq[4], mq[4], h2[4], m2[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = lds128(p0 + (((uint32_t)j ^ sw) << 4));
#pragma unroll
    for (int q = 0; q < 4; ++q) bf16_split8(f[2 * q], f[2 * q + 1], hq[q], mq[q]);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = has1 ? lds128(p1 + (((uint32_t)j ^ sw) << 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 4; ++q) bf16_split8(f[2 * q], f[2 * q + 1], h2[q], m2[q]);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        sts128u(p0 + (((uint32_t)q ^ sw) << 4), hq[q]);
        sts128u(p0 + (((uint32_t)(q + 4) ^ sw) << 4), h2[q]);
        sts128u(p1 + (((uint32_t)q ^ sw) << 4), mq[q]);
        sts128u(p1 + (((uint32_t)(q + 4) ^ sw) << 4), m2[q]);
    }
}
// One 128-byte activation row (32 fp32 channels, 16-byte chunk j stored at j ^ sw) rewritten IN PLACE as [h of the 32 channels |
// m of the 32 channels] in bf16 (chunk q of h at q ^ sw, chunk q of m at (4 + q) ^ sw): the K-interleaved BF16x3 operand row.
__device__ __forceinline__ void bf16_interleave_row(uint32_t p, uint32_t sw) {
    float4 f[8];
    uint4 hq[4], mq[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = lds128(p + (((uint32_t)j ^ sw) << 4));
#pragma unroll
    for (int q = 0; q < 4; ++q) bf16_split8(f[2 * q], f[2 * q + 1], hq[q], mq[q]);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        sts128u(p + (((uint32_t)q ^ sw) << 4), hq[q]);
        sts128u(p + (((uint32_t)(q + 4) ^ sw) << 4), mq[q]);
    }
}
// Strict fp32 mode ("TF32 + BF16 cross terms").  The kind::tf32 MMA truncates its fp32 operands to their upper 19 bits, so the raw
// row x IS the hi*hi operand (hi = trunc_tf32(x), nothing is written back) and the converter only builds the kind::f16 operand of the
// two cross products lo_a*hi_b + hi_a*lo_b: the row [bf16(x) of the 32 channels | bf16(lo) of the 32 channels] with lo = x - hi
// (exact in fp32), same swizzle (chunk q of the hi16 half at q ^ sw, of the lo16 half at (4 + q) ^ sw).  |lo| < 2^-10 |x| and bf16
// keeps 8 bits (round to nearest), so each cross term carries a relative error of 2^-19 at most; the dropped lo*lo term is 2^-20 at
// most and has zero mean against the weights' round-to-nearest split.  2 instead of 3 TF32-equivalent MMAs, 2/3 of the operand
// bytes of 3xTF32, and per 8 channels 2 loads + 2 stores + ~30 ALU instructions (the first version rounded hi to nearest and wrote
// it back: 2 more stores and 2.5x the ALU work, which made the CONVERTER the limiter of the 1x1 convolutions, grams and mixes).
// Work item = a QUARTER of a row (8 channels: fp32 chunks 2*qd, 2*qd + 1 -> one 16-byte chunk of hi16 and one of lo16).
__device__ __forceinline__ uint32_t bf16x2_rn(float upper, float lower) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(upper), "f"(lower));
    return d;
}
__device__ __forceinline__ void cross_pack8(const float4& f0, const float4& f1, uint4& h, uint4& l) {
    const float x[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
    float lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) lo[i] = x[i] - __uint_as_float(__float_as_uint(x[i]) & 0xffffe000u);
    h = make_uint4(bf16x2_rn(x[1], x[0]), bf16x2_rn(x[3], x[2]), bf16x2_rn(x[5], x[4]), bf16x2_rn(x[7], x[6]));
    l = make_uint4(bf16x2_rn(lo[1], lo[0]), bf16x2_rn(lo[3], lo[2]), bf16x2_rn(lo[5], lo[4]), bf16x2_rn(lo[7], lo[6]));
}
__device__ __forceinline__ void tf32_cross_quarter(uint32_t p, uint32_t pc, uint32_t sw, uint32_t qd) {
    const float4 f0 = lds128(p + (((2u * qd) ^ sw) << 4)), f1 = lds128(p + (((2u * qd + 1u) ^ sw) << 4));
    uint4 h, l;
    cross_pack8(f0, f1, h, l);
    sts128u(pc + ((qd ^ sw) << 4), h);
    sts128u(pc + (((qd + 4u) ^ sw) << 4), l);
}
// The same split for MN-major operands (channels contiguous, the joint / row index is K).  Source: quarter `qd` (8 channels) of row
// `r` of a 32-channel fp32 box that TMA landed in the 32-byte-atom swizzle (Swizzle<2,5,2>: 32-byte chunk index ^= row & 3).
// Destination: a plain-128B-swizzled bf16 block of 64 channels per 128-byte row (8-row swizzle groups): bf16(x) goes to the
// logical 16-byte chunk `c` of row `row_hi`, bf16(lo) to the same chunk of row `row_lo` (the two pieces are stacked along K, so that
// ONE chain of K = 16 MMAs against the other operand's [lo ; hi] block yields hi*lo + lo*hi).
__device__ __forceinline__ void tf32_cross_quarter_mn(uint32_t box, uint32_t blk, uint32_t r, uint32_t qd, uint32_t c, uint32_t row_hi, uint32_t row_lo) {
    const uint32_t p = box + r * 128u + ((qd ^ (r & 3u)) << 5);
    const float4 f0 = lds128(p), f1 = lds128(p + 16u);
    uint4 h, l;
    cross_pack8(f0, f1, h, l);
    sts128u(blk + row_hi * 128u + ((c ^ (row_hi & 7u)) << 4), h);
    sts128u(blk + row_lo * 128u + ((c ^ (row_lo & 7u)) << 4), l);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// MN-major (channels contiguous) 16-bit operand, plain 128B swizzle: 64 bf16 along M/N per 128-byte row, rows along K in groups of
// 8 (1024 bytes); `lbo_bytes` is the distance between successive 64-wide M/N blocks
__device__ __forceinline__ uint64_t make_smem_desc_mn16(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
    return d;
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (rows at 128 B pitch, 8-row atoms at 1024 B pitch)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);        // start address, 16-byte units
    d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset between 8-row atoms
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}


// MN-major (channels contiguous) operand, 128B swizzle with 32-byte atoms: 32 fp32 along M/N per 128-byte row, rows along K in
// groups of 4 (512 bytes); `lbo_bytes` is the distance between successive 32-wide M/N atoms
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;   // leading byte offset: between 32-channel (MN) atoms
    d |= (uint64_t)(512 >> 4) << 32;                     // stride byte offset: between 4-row (K) groups of the 32B-atom swizzle
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                              // SWIZZLE_128B_BASE32B: the only layout for MN-major tf32 operands
    return d;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}



// weights: w_split[0 .. n) = rna_tf32(w), w_split[n .. 2n) = w - hi

static __global__ void split_weights_kernel(const float* w, float* w_split, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float v = w[i];
        const float hi = tf32_rna(v);
        w_split[i] = hi;
        w_split[n + i] = v - hi;
    }
}


}  // namespace tc
}  // namespace agcn
