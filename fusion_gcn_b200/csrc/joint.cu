// V x V attention kernels: the score / dG joint-gram reduction, the column softmax, and the per-sample
// mixing over the joint axis with A_k + B_k + C_k resident in shared memory.
// Reference arithmetic: torch_src/models/mmargcn/agcn.py:98-110 (matmul, softmax(-2), matmul) and its backward.
#include "common.cuh"

namespace agcn {

constexpr int kMaxV = 32;

// ------------------------------------------------------------------------------------------ joint gram
// out[n][chunk][g][u][v] = sum_{t in chunk, c<width} a[n][t][u][offa+g*sa+c] * b[n][t][v][offb+g*sb+c]
// CTA = (n, chunk).  Thread tile: 5x5 block of (u,v) for one group, over one of 4 interleaved channel slices.
struct GramArgs {
    const float* a; const float* b; float* out;
    int nb, t, v, lda, ldb, groups, offa, sa, offb, sb, width, nchunk;
};

constexpr int kGramCW = 64;     // channels staged per pass
constexpr int kGramThreads = 320;

__global__ void __launch_bounds__(kGramThreads) joint_gram_kernel(GramArgs p) {
    extern __shared__ __align__(16) float smem[];
    const int V = p.v;
    const int cw = p.width < kGramCW ? p.width : kGramCW;
    const int ld = cw + 4;
    // A tile: groups_a x V x ld  (groups_a = 1 when sa == 0), B tile: groups x V x ld
    const int ga = (p.sa == 0) ? 1 : p.groups;
    float* As = smem;
    float* Bs = smem + (size_t)ga * V * ld;
    const int n = blockIdx.x / p.nchunk, chunk = blockIdx.x % p.nchunk;
    const int t_per = (p.t + p.nchunk - 1) / p.nchunk;
    const int t0 = chunk * t_per;
    int t1 = t0 + t_per; if (t1 > p.t) t1 = p.t;

    const int nblk = (V + 4) / 5;                 // 5-wide blocks per axis
    const int tiles = p.groups * nblk * nblk;
    const int tid = threadIdx.x;
    const int cs = tid & 3;                       // channel slice
    // each thread owns up to 2 tiles (tiles <= 3*7*7 = 147 <= 2 * 80)
    int tile_id[2]; int tg[2], tu[2], tv[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        int id = (tid >> 2) + s * (kGramThreads / 4);
        tile_id[s] = id < tiles ? id : -1;
        int g = id / (nblk * nblk); int r = id % (nblk * nblk);
        tg[s] = g; tu[s] = (r / nblk) * 5; tv[s] = (r % nblk) * 5;
    }
    float acc[2][5][5];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int j = 0; j < 5; ++j) acc[s][i][j] = 0.f;

    for (int t = t0; t < t1; ++t) {
        const float* arow = p.a + ((long long)n * p.t + t) * V * p.lda;
        const float* brow = p.b + ((long long)n * p.t + t) * V * p.ldb;
        for (int c0 = 0; c0 < p.width; c0 += cw) {
            const int cn = (p.width - c0) < cw ? (p.width - c0) : cw;
            __syncthreads();
            // stage A
            for (int idx = tid; idx < ga * V * cw; idx += kGramThreads) {
                int c = idx % cw; int r = idx / cw; int u = r % V; int g = r / V;
                float val = 0.f;
                if (c < cn) val = __ldg(arow + (long long)u * p.lda + p.offa + g * p.sa + c0 + c);
                As[(g * V + u) * ld + c] = val;
            }
            for (int idx = tid; idx < p.groups * V * cw; idx += kGramThreads) {
                int c = idx % cw; int r = idx / cw; int u = r % V; int g = r / V;
                float val = 0.f;
                if (c < cn) val = __ldg(brow + (long long)u * p.ldb + p.offb + g * p.sb + c0 + c);
                Bs[(g * V + u) * ld + c] = val;
            }
            __syncthreads();
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (tile_id[s] < 0) continue;
                const float* Ab = As + (size_t)((p.sa == 0 ? 0 : tg[s]) * V) * ld;
                const float* Bb = Bs + (size_t)(tg[s] * V) * ld;
                for (int c = cs; c < cn; c += 4) {
                    float av[5], bv[5];
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        int u = tu[s] + i; av[i] = (u < V) ? Ab[u * ld + c] : 0.f;
                        int w = tv[s] + i; bv[i] = (w < V) ? Bb[w * ld + c] : 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 5; ++i)
#pragma unroll
                        for (int j = 0; j < 5; ++j) acc[s][i][j] = fmaf(av[i], bv[j], acc[s][i][j]);
                }
            }
        }
    }
    // reduce the 4 channel slices (adjacent lanes) and store
    float* o = p.out + ((long long)n * p.nchunk + chunk) * p.groups * V * V;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                float x = acc[s][i][j];
                x += __shfl_xor_sync(0xffffffffu, x, 1);
                x += __shfl_xor_sync(0xffffffffu, x, 2);
                if (cs == 0 && tile_id[s] >= 0) {
                    int u = tu[s] + i, w = tv[s] + j;
                    if (u < V && w < V) o[((long long)tg[s] * V + u) * V + w] = x;
                }
            }
    }
}

// ------------------------------------------------------------------------------------------ attention fwd / bwd
__global__ void attention_fwd_kernel(const float* s_part, const float* adj_a, const float* adj_b, float* p, float* g,
                                     int nb, int nchunk, int groups, int V, float scale) {
    // one thread per (n, k, v) column
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)nb * groups * V;
    if (idx >= total) return;
    const int v = (int)(idx % V); const int k = (int)((idx / V) % groups); const int n = (int)(idx / ((long long)V * groups));
    float s[kMaxV];
    float mx = -INFINITY;
    for (int u = 0; u < V; ++u) {
        float acc = 0.f;
        for (int c = 0; c < nchunk; ++c)
            acc += s_part[((((long long)n * nchunk + c) * groups + k) * V + u) * V + v];
        acc *= scale;
        s[u] = acc;
        mx = fmaxf(mx, acc);
    }
    float den = 0.f;
    for (int u = 0; u < V; ++u) { s[u] = expf(s[u] - mx); den += s[u]; }
    const float inv = 1.f / den;
    for (int u = 0; u < V; ++u) {
        const long long o = (((long long)n * groups + k) * V + u) * V + v;
        const float pv = s[u] * inv;
        p[o] = pv;
        const int ao = (k * V + u) * V + v;
        g[o] = pv + adj_a[ao] + adj_b[ao];
    }
}

__global__ void attention_bwd_kernel(const float* dg_part, const float* p, float* dg_sum, float* ds,
                                     int nb, int nchunk, int groups, int V, float scale) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)nb * groups * V;
    if (idx >= total) return;
    const int v = (int)(idx % V); const int k = (int)((idx / V) % groups); const int n = (int)(idx / ((long long)V * groups));
    float dg[kMaxV], pv[kMaxV];
    float dot = 0.f;
    for (int u = 0; u < V; ++u) {
        float acc = 0.f;
        for (int c = 0; c < nchunk; ++c)
            acc += dg_part[((((long long)n * nchunk + c) * groups + k) * V + u) * V + v];
        const long long o = (((long long)n * groups + k) * V + u) * V + v;
        dg[u] = acc; pv[u] = p[o];
        dg_sum[o] = acc;
        dot = fmaf(pv[u], acc, dot);
    }
    for (int u = 0; u < V; ++u) {
        const long long o = (((long long)n * groups + k) * V + u) * V + v;
        ds[o] = scale * pv[u] * (dg[u] - dot);
    }
}

__global__ void sum_over_samples_kernel(const float* dg_sum, float* dadj_b, int nb, int per_sample) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per_sample) return;
    float s = 0.f;
    for (int n = 0; n < nb; ++n) s += dg_sum[(long long)n * per_sample + i];
    dadj_b[i] = s;
}

// ------------------------------------------------------------------------------------------ joint mix
// Work item = (t in tile, output group, block of 5 output joints, CV consecutive channels).
// For each term of the output group:  acc[j][c] += in_s[i][ingroup*w + c] * M_s[mat][i][j]   (i over joints)
struct MixArgs {
    const float* in; const float* mats; float* out;
    int nb, t, v, ldin, ldout, width, mode, accumulate, tt;
};

template <int CV>
__global__ void __launch_bounds__(256) joint_mix_kernel(MixArgs p) {
    extern __shared__ __align__(16) float smem[];
    const int V = p.v, W = p.width;
    const int nblk = (V + 4) / 5;
    const int mld = nblk * 8;                       // padded row of a staged matrix (5 of every 8 used)
    // staged matrices: index [slot][i][jb*8 + jj]; slot meaning depends on mode
    const int nslots = (p.mode == AGCN_MIX_SCORE_BWD) ? 6 : 3;
    float* Ms = smem;
    float* in_s = smem + (size_t)nslots * V * mld;  // [tt][V][ldin]
    const int tid = threadIdx.x;
    const int tiles_t = (p.t + p.tt - 1) / p.tt;
    const int n = blockIdx.x / tiles_t;
    const int t0 = (blockIdx.x % tiles_t) * p.tt;
    const int tn = (p.t - t0) < p.tt ? (p.t - t0) : p.tt;

    // stage matrices.  slot s holds Mat[i][j]:
    //  AGG_FWD : slot k = G_k[i=u][j=v]
    //  AGG_BWD : slot k = G_k^T  -> [i=v][j=u] = G_k[u][v]
    //  SCORE_BWD: slot 2k   (output dtheta_k, input phi_k):   [i=v][j=u] = dS_k[u][v]
    //             slot 2k+1 (output dphi_k,   input theta_k): [i=u][j=v] = dS_k[u][v]
    const float* mat_n = p.mats + (long long)n * 3 * V * V;
    for (int idx = tid; idx < nslots * V * V; idx += 256) {
        const int j = idx % V; const int i = (idx / V) % V; const int s = idx / (V * V);
        float val;
        if (p.mode == AGCN_MIX_AGG_FWD) val = mat_n[(s * V + i) * V + j];
        else if (p.mode == AGCN_MIX_AGG_BWD) val = mat_n[(s * V + j) * V + i];
        else {
            const int k = s >> 1;
            val = (s & 1) ? mat_n[(k * V + i) * V + j] : mat_n[(k * V + j) * V + i];
        }
        Ms[((size_t)s * V + i) * mld + (j / 5) * 8 + (j % 5)] = val;
    }
    // stage input rows (contiguous in memory: tn * V * ldin floats)
    {
        const float* src = p.in + ((long long)n * p.t + t0) * V * p.ldin;
        const int total = tn * V * p.ldin;
        if (CV == 4) {
            const float4* s4 = reinterpret_cast<const float4*>(src);
            float4* d4 = reinterpret_cast<float4*>(in_s);
            for (int idx = tid; idx < total / 4; idx += 256) d4[idx] = __ldg(s4 + idx);
        } else {
            for (int idx = tid; idx < total; idx += 256) in_s[idx] = __ldg(src + idx);
        }
    }
    __syncthreads();

    const int ogroups = (p.mode == AGCN_MIX_AGG_FWD) ? 3 : (p.mode == AGCN_MIX_AGG_BWD ? 1 : 6);
    const int cq = W / CV;                          // channel vectors per group
    const int items = tn * ogroups * nblk * cq;
    for (int item = tid; item < items; item += 256) {
        const int c = (item % cq) * CV;
        int r = item / cq;
        const int jb = r % nblk; r /= nblk;
        const int og = r % ogroups;
        const int tl = r / ogroups;
        float acc[5][CV];
#pragma unroll
        for (int j = 0; j < 5; ++j)
#pragma unroll
            for (int e = 0; e < CV; ++e) acc[j][e] = 0.f;
        const int nterms = (p.mode == AGCN_MIX_AGG_BWD) ? 3 : 1;
        for (int term = 0; term < nterms; ++term) {
            int ig, slot;
            if (p.mode == AGCN_MIX_AGG_FWD) { ig = 0; slot = og; }
            else if (p.mode == AGCN_MIX_AGG_BWD) { ig = term; slot = term; }
            else { ig = og ^ 1; slot = og; }
            const float* xin = in_s + (size_t)tl * V * p.ldin + ig * W + c;
            const float* m = Ms + (size_t)slot * V * mld + jb * 8;
            for (int i = 0; i < V; ++i) {
                float xv[CV];
                if (CV == 4) {
                    float4 q = *reinterpret_cast<const float4*>(xin + (size_t)i * p.ldin);
                    xv[0] = q.x; xv[1] = q.y; xv[2] = q.z; xv[3] = q.w;
                } else {
#pragma unroll
                    for (int e = 0; e < CV; ++e) xv[e] = xin[(size_t)i * p.ldin + e];
                }
                const float4 m0 = *reinterpret_cast<const float4*>(m + (size_t)i * mld);
                const float m4 = m[(size_t)i * mld + 4];
                const float mv[5] = {m0.x, m0.y, m0.z, m0.w, m4};
#pragma unroll
                for (int j = 0; j < 5; ++j)
#pragma unroll
                    for (int e = 0; e < CV; ++e) acc[j][e] = fmaf(xv[e], mv[j], acc[j][e]);
            }
        }
        const int t = t0 + tl;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const int jj = jb * 5 + j;
            if (jj >= V) continue;
            float* o = p.out + (((long long)n * p.t + t) * V + jj) * p.ldout + og * W + c;
            if (CV == 4) {
                float4 q = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
                if (p.accumulate) {
                    float4 old = *reinterpret_cast<const float4*>(o);
                    q.x += old.x; q.y += old.y; q.z += old.z; q.w += old.w;
                }
                *reinterpret_cast<float4*>(o) = q;
            } else {
#pragma unroll
                for (int e = 0; e < CV; ++e) o[e] = p.accumulate ? o[e] + acc[j][e] : acc[j][e];
            }
        }
    }
}

}  // namespace agcn

using namespace agcn;

extern "C" AGCN_API int agcn_joint_gram(const float* a, const float* b, float* out,
                               int nb, int t, int v, int lda, int ldb, int groups,
                               int offa, int stridea, int offb, int strideb, int width, int nchunk, void* stream) {
    AGCN_REQUIRE(a && b && out, AGCN_ERR_NULL, "agcn_joint_gram: null pointer");
    AGCN_REQUIRE(nb > 0 && t > 0 && v > 0 && groups > 0 && width > 0 && nchunk > 0 && nchunk <= t,
                 AGCN_ERR_BAD_SHAPE, "agcn_joint_gram: bad shape nb=%d t=%d v=%d groups=%d width=%d nchunk=%d", nb, t, v, groups, width, nchunk);
    AGCN_REQUIRE(v <= kMaxV && groups <= 3, AGCN_ERR_UNSUPPORTED, "agcn_joint_gram: V=%d > %d or groups=%d > 3", v, kMaxV, groups);
    AGCN_REQUIRE(offa >= 0 && offb >= 0 && offa + (groups - 1) * stridea + width <= lda && offb + (groups - 1) * strideb + width <= ldb,
                 AGCN_ERR_BAD_SHAPE, "agcn_joint_gram: channel window outside the row");
    GramArgs p{a, b, out, nb, t, v, lda, ldb, groups, offa, stridea, offb, strideb, width, nchunk};
    const int cw = width < kGramCW ? width : kGramCW;
    const int ga = stridea == 0 ? 1 : groups;
    size_t smem = (size_t)(ga + groups) * v * (cw + 4) * sizeof(float);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    joint_gram_kernel<<<nb * nchunk, kGramThreads, smem, s>>>(p);
    return check_launch("agcn_joint_gram");
}

extern "C" AGCN_API int agcn_attention_fwd(const float* s_part, const float* adj_a, const float* adj_b, float* p, float* g,
                                  int nb, int nchunk, int groups, int v, float scale, void* stream) {
    AGCN_REQUIRE(s_part && adj_a && adj_b && p && g, AGCN_ERR_NULL, "agcn_attention_fwd: null pointer");
    AGCN_REQUIRE(nb > 0 && nchunk > 0 && groups > 0 && v > 0, AGCN_ERR_BAD_SHAPE, "agcn_attention_fwd: bad shape");
    AGCN_REQUIRE(v <= kMaxV, AGCN_ERR_UNSUPPORTED, "agcn_attention_fwd: V=%d > %d", v, kMaxV);
    long long total = (long long)nb * groups * v;
    attention_fwd_kernel<<<ceil_div(total, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(s_part, adj_a, adj_b, p, g, nb, nchunk, groups, v, scale);
    return check_launch("agcn_attention_fwd");
}

extern "C" AGCN_API int agcn_attention_bwd(const float* dg_part, const float* p, float* dg_sum, float* ds, float* dadj_b,
                                  int nb, int nchunk, int groups, int v, float scale, void* stream) {
    AGCN_REQUIRE(dg_part && p && dg_sum && ds && dadj_b, AGCN_ERR_NULL, "agcn_attention_bwd: null pointer");
    AGCN_REQUIRE(nb > 0 && nchunk > 0 && groups > 0 && v > 0, AGCN_ERR_BAD_SHAPE, "agcn_attention_bwd: bad shape");
    AGCN_REQUIRE(v <= kMaxV, AGCN_ERR_UNSUPPORTED, "agcn_attention_bwd: V=%d > %d", v, kMaxV);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    long long total = (long long)nb * groups * v;
    attention_bwd_kernel<<<ceil_div(total, 128), 128, 0, s>>>(dg_part, p, dg_sum, ds, nb, nchunk, groups, v, scale);
    int rc = check_launch("agcn_attention_bwd");
    if (rc) return rc;
    const int per = groups * v * v;
    sum_over_samples_kernel<<<ceil_div(per, 128), 128, 0, s>>>(dg_sum, dadj_b, nb, per);
    return check_launch("agcn_attention_bwd(dadj_b)");
}

extern "C" AGCN_API int agcn_joint_mix(const float* in, const float* mats, float* out,
                              int nb, int t, int v, int ldin, int ldout, int width, int mode, int accumulate, void* stream) {
    AGCN_REQUIRE(in && mats && out, AGCN_ERR_NULL, "agcn_joint_mix: null pointer");
    AGCN_REQUIRE(nb > 0 && t > 0 && v > 0 && width > 0, AGCN_ERR_BAD_SHAPE, "agcn_joint_mix: bad shape");
    AGCN_REQUIRE(v <= kMaxV, AGCN_ERR_UNSUPPORTED, "agcn_joint_mix: V=%d > %d", v, kMaxV);
    int need_in, need_out;
    if (mode == AGCN_MIX_AGG_FWD) { need_in = width; need_out = 3 * width; }
    else if (mode == AGCN_MIX_AGG_BWD) { need_in = 3 * width; need_out = width; }
    else if (mode == AGCN_MIX_SCORE_BWD) { need_in = 6 * width; need_out = 6 * width; }
    else return fail(AGCN_ERR_UNSUPPORTED, "agcn_joint_mix: unknown mode %d", mode);
    AGCN_REQUIRE(ldin == need_in && ldout == need_out, AGCN_ERR_BAD_SHAPE,
                 "agcn_joint_mix: mode %d expects ldin=%d ldout=%d, got %d %d", mode, need_in, need_out, ldin, ldout);
    const int nblk = (v + 4) / 5, mld = nblk * 8;
    const int nslots = mode == AGCN_MIX_SCORE_BWD ? 6 : 3;
    // choose tt so that the staged rows stay below ~64 KB and there are enough work items per CTA
    int tt = 1;
    const size_t row_bytes = (size_t)v * ldin * sizeof(float);
    while (tt < 8 && (size_t)(tt * 2) * row_bytes <= 48 * 1024 && tt * 2 <= t) tt *= 2;
    size_t smem = (size_t)nslots * v * mld * sizeof(float) + (size_t)tt * row_bytes;
    AGCN_REQUIRE(smem <= 200 * 1024, AGCN_ERR_UNSUPPORTED, "agcn_joint_mix: row too wide for shared memory (%zu bytes)", smem);
    MixArgs p{in, mats, out, nb, t, v, ldin, ldout, width, mode, accumulate, tt};
    const int tiles_t = (t + tt - 1) / tt;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool vec = (width % 4 == 0) && aligned16(in) && aligned16(out);
    cudaError_t e;
    if (vec) {
        e = cudaFuncSetAttribute(joint_mix_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return fail(AGCN_ERR_CUDA, "agcn_joint_mix: %s", cudaGetErrorString(e));
        joint_mix_kernel<4><<<nb * tiles_t, 256, smem, s>>>(p);
    } else {
        e = cudaFuncSetAttribute(joint_mix_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return fail(AGCN_ERR_CUDA, "agcn_joint_mix: %s", cudaGetErrorString(e));
        joint_mix_kernel<1><<<nb * tiles_t, 256, smem, s>>>(p);
    }
    return check_launch("agcn_joint_mix");
}
