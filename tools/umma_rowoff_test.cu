// Hardware probe (run on the GPU box): do tcgen05.mma shared-memory descriptors accept a start address that is shifted
// by a whole number of 128-byte rows which is NOT a multiple of the swizzle repeat (8 rows for K-major SWIZZLE_128B,
// 4 rows for the MN-major 128B/32B-atom layout)?  The halo-resident temporal-conv kernels rely on it: one TMA box holds
// tt+8 timesteps and tap `d` of the 9x1 conv reads the same tile starting d*V rows further down.
//   build: nvcc -std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a tools/umma_rowoff_test.cu -o tools/umma_rowoff_test -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW1:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D1;\n\tbra W1;\n\tD1:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// mode 0: K-major SW128.  A: 256 rows x 32 fp32 (one box), B: 64 rows x 32.  D[m][n] = sum_k A[roff+m][k] B[n][k], K = 32.
// mode 1: MN-major SW128/32B atom.  A = dy: 4 sub-boxes (32 ch x RP rows), B = x: 2 sub-boxes (32 ch x RP rows, rows shifted by roff)
//         D[co][ci] = sum_{r<64} dy[r][co] x[r+roff][ci]
constexpr int RP = 128;   // rows per sub-box in mode 1
__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                int mode, int roff, int bo_mode, float* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sa = base, sb = base + 64 * 1024, bar = base + 128 * 1024, slot = bar + 64;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 8, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "n"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    if (threadIdx.x == 0) {
        if (mode == 0) {
            mbar_expect_tx(bar, 256 * 128 + 64 * 128);
            tma_load_2d(sa, &map_a, bar, 0, 0);
            tma_load_2d(sb, &map_b, bar, 0, 0);
        } else {
            mbar_expect_tx(bar, 6 * RP * 128);
            for (int i = 0; i < 4; ++i) tma_load_2d(sa + i * RP * 128, &map_a, bar, 32 * i, 0);
            for (int j = 0; j < 2; ++j) tma_load_2d(sb + j * RP * 128, &map_b, bar, 32 * j, 0);
        }
        mbar_wait(bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        auto bo = [&](uint32_t addr) -> uint64_t {
            if (bo_mode == 0) return 0;
            if (bo_mode == 1) return (uint64_t)((addr >> 7) & 7u);
            return (uint64_t)((addr >> 7) & 3u);
        };
        if (mode == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t a0 = sa + (uint32_t)roff * 128u;
            uint64_t da = (uint64_t)((a0 & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | (bo(a0) << 49) | ((uint64_t)2 << 61);
            uint64_t db = (uint64_t)((sb & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
            for (int k = 0; k < 4; ++k) umma_tf32(tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, k ? 1u : 0u);
        } else {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t b0 = sb + (uint32_t)roff * 128u;
            const uint32_t lbo = RP * 128;
            uint64_t da = (uint64_t)((sa & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
            uint64_t db = (uint64_t)((b0 & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | (bo(b0) << 49) | ((uint64_t)1 << 61);
            for (int kg = 0; kg < 8; ++kg) umma_tf32(tmem, da + (uint64_t)(kg * 64), db + (uint64_t)(kg * 64), idesc, kg ? 1u : 0u);
        }
        umma_commit(bar + 8);
    }
    mbar_wait(bar + 8, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < 64; c += 16) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
        for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 64 + c + i] = v[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= ~0x1FFFu; memcpy(&x, &u, 4); return x; }

int main() {
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const int RA = 256, CA = 128, RB = 256, CB = 64;
    std::vector<float> ha(RA * CA), hb(RB * CB);
    srand(1);
    for (auto& v : ha) v = trunc_tf32((float)rand() / RAND_MAX - 0.5f);
    for (auto& v : hb) v = trunc_tf32((float)rand() / RAND_MAX - 0.5f);
    float *da, *db, *dout;
    CK(cudaMalloc(&da, ha.size() * 4)); CK(cudaMalloc(&db, hb.size() * 4)); CK(cudaMalloc(&dout, 128 * 64 * 4));
    CK(cudaMemcpy(da, ha.data(), ha.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024));
    std::vector<float> ho(128 * 64);
    for (int mode = 0; mode < 2; ++mode) {
        // mode 0: A viewed as [256 rows][32 ch] = first 32 channels of each row (row pitch CA), B as [64 rows][32] of hb (pitch CB)
        CUtensorMap ma, mb;
        cuuint32_t estr[2] = {1, 1};
        if (mode == 0) {
            cuuint64_t dims[2] = {32, (cuuint64_t)RA}; cuuint64_t str[1] = {(cuuint64_t)CA * 4}; cuuint32_t box[2] = {32, 256};
            CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, da, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            cuuint64_t dimsb[2] = {32, 64}; cuuint64_t strb[1] = {(cuuint64_t)CB * 4}; cuuint32_t boxb[2] = {32, 64};
            CUresult r2 = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, db, dimsb, strb, boxb, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r || r2) { printf("encode failed %d %d\n", (int)r, (int)r2); return 1; }
        } else {
            cuuint64_t dims[2] = {(cuuint64_t)CA, (cuuint64_t)RA}; cuuint64_t str[1] = {(cuuint64_t)CA * 4}; cuuint32_t box[2] = {32, RP};
            CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, da, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            cuuint64_t dimsb[2] = {(cuuint64_t)CB, (cuuint64_t)RB}; cuuint64_t strb[1] = {(cuuint64_t)CB * 4};
            CUresult r2 = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, db, dimsb, strb, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r || r2) { printf("encode failed %d %d\n", (int)r, (int)r2); return 1; }
        }
        const int roffs[] = {0, 8, 4, 1, 3, 25, 50, 22, 40, 63};
        for (int roff : roffs) {
            for (int bo = 0; bo < 3; ++bo) {
                CK(cudaMemset(dout, 0, 128 * 64 * 4));
                probe<<<1, 128, 140 * 1024>>>(ma, mb, mode, roff, bo, dout);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d roff %d bo %d: %s\n", mode, roff, bo, cudaGetErrorString(e)); return 1; }
                CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
                double worst = 0, scale = 0;
                for (int m = 0; m < 128; ++m)
                    for (int n = 0; n < 64; ++n) {
                        double ref = 0;
                        if (mode == 0) { for (int k = 0; k < 32; ++k) ref += (double)ha[(size_t)(roff + m) * CA + k] * hb[(size_t)n * CB + k]; }
                        else { for (int r = 0; r < 64; ++r) ref += (double)ha[(size_t)r * CA + m] * hb[(size_t)(r + roff) * CB + n]; }
                        worst = fmax(worst, fabs(ref - ho[m * 64 + n])); scale = fmax(scale, fabs(ref));
                    }
                printf("mode %d (%s) row offset %2d base_offset mode %d: max err %.3e (scale %.3e) %s\n", mode, mode ? "MN-major SW128/32B" : "K-major SW128",
                       roff, bo, worst, scale, worst < 1e-4 * scale ? "OK" : "MISMATCH");
            }
        }
    }
    return 0;
}
