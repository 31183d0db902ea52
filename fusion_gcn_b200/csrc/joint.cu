// V x V attention kernels: the score / dG joint-gram reduction, the column softmax, and the per-sample
// mixing over the joint axis with A_k + B_k + C_k resident in shared memory.
// Reference arithmetic: torch_src/models/mmargcn/agcn.py:98-110 (matmul, softmax(-2), matmul) and its backward.
#include "common.cuh"
#include <stdlib.h>

namespace agcn {

constexpr int kMaxV = 32;

// ------------------------------------------------------------------------------------------ joint gram
// out[n][chunk][g][u][v] = sum_{t in chunk, c<width} a[n][t][u][offa+g*sa+c] * b[n][t][v][offb+g*sb+c]
// CTA = (n, chunk), 12 warps.  The (g, u, v) outputs are cut into 5x5 register tiles; tile ids run along the lanes of
// `ntw` "tile warps", and the remaining warp index is a slice of the reduction axis (t, c), so every lane of a warp
// reads the same channel quad: the five distinct joint rows a warp touches per operand are shared-memory broadcasts
// (row pitch ld = cw+4 floats keeps them in distinct bank groups), giving 100 FMAs per 10 LDS.128.
// Operands stay channels-contiguous in shared memory exactly as they lie in HBM: staging is straight 16-byte
// cp.async (double buffered); the k-slices are summed through shared memory in a fixed order (deterministic).
struct GramArgs {
    const float* a; const float* b; float* out;
    int nb, t, v, lda, ldb, groups, offa, sa, offb, sb, width, nchunk, tt, vec;
};

constexpr int kGramCW = 64;     // channels staged per pass
constexpr int kGramWarps = 12;
constexpr int kGramThreads = kGramWarps * 32;

__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(kGramThreads, 2) joint_gram_kernel(GramArgs p) {
    extern __shared__ __align__(16) float smem[];
    const int V = p.v;
    const int cwp = ((p.width < kGramCW ? p.width : kGramCW) + 3) & ~3;      // staged channels, padded to a quad
    const int ld = cwp + 4;
    const int cq = cwp >> 2;
    const int ga = (p.sa == 0) ? 1 : p.groups;
    const int rows_a = ga * V, rows_b = p.groups * V;
    const int stage_floats = p.tt * (rows_a + rows_b) * ld;
    const int n = blockIdx.x / p.nchunk, chunk = blockIdx.x % p.nchunk;
    const int t_per = (p.t + p.nchunk - 1) / p.nchunk;
    const int t0 = chunk * t_per;
    int t1 = t0 + t_per; if (t1 > p.t) t1 = p.t;

    const int nblk = (V + 4) / 5;
    const int tiles = p.groups * nblk * nblk;
    const int ntw = (tiles + 31) >> 5;
    const int nks = kGramWarps / ntw;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tw = warp % ntw, ks = warp / ntw;
    const int tile = tw * 32 + lane;
    const bool active = (ks < nks) && (tile < tiles);
    int g = 0, u0 = 0, v0 = 0;
    if (tile < tiles) { g = tile / (nblk * nblk); const int r = tile % (nblk * nblk); u0 = (r / nblk) * 5; v0 = (r % nblk) * 5; }
    int offu[5], offv[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int u = (u0 + i < V) ? u0 + i : V - 1;          // clamped rows compute values that are never stored
        const int w = (v0 + i < V) ? v0 + i : V - 1;
        offu[i] = ((p.sa == 0 ? 0 : g) * V + u) * ld;
        offv[i] = (rows_a * p.tt) * ld + (g * V + w) * ld;
    }
    float acc[5][5];
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) acc[i][j] = 0.f;

    // stage list: (timestep block, channel pass)
    const int npass = (p.width + kGramCW - 1) / kGramCW;
    const int ntb = (t1 - t0 + p.tt - 1) / p.tt;
    const int nstage = ntb * npass;

    // Per-timestep copy table (built once): element e of operand A / B -> (source offset relative to the timestep's first
    // row and the pass's first channel, destination offset inside the stage).  Staging a timestep is then one table read,
    // two adds and a 16-byte cp.async per element; no index arithmetic in the stage loop.
    int2* tab_a = reinterpret_cast<int2*>(smem + 2 * stage_floats);
    const int ne_a = p.vec ? rows_a * cq : rows_a * cwp;
    const int ne_b = p.vec ? rows_b * cq : rows_b * cwp;
    int2* tab_b = tab_a + ne_a;
    {
        const int per_row = p.vec ? cq : cwp, cstep = p.vec ? 4 : 1;
        for (int e = tid; e < ne_a; e += kGramThreads) {
            const int c = (e % per_row) * cstep, row = e / per_row, gg = row / V, u = row - gg * V;
            tab_a[e] = make_int2(u * p.lda + p.offa + gg * p.sa + c, (row * ld + c) | (c << 20));
        }
        for (int e = tid; e < ne_b; e += kGramThreads) {
            const int c = (e % per_row) * cstep, row = e / per_row, gg = row / V, u = row - gg * V;
            tab_b[e] = make_int2(u * p.ldb + p.offb + gg * p.sb + c, (row * ld + c) | (c << 20));
        }
    }
    __syncthreads();

    auto issue = [&](int s, float* dst) {
        const int tb = s / npass, ps = s - tb * npass;
        const int ts = t0 + tb * p.tt;
        const int ttn = (t1 - ts) < p.tt ? (t1 - ts) : p.tt;
        const int c0 = ps * kGramCW;
        const int cn = (p.width - c0) < cwp ? (p.width - c0) : cwp;      // valid channels in this pass
        for (int tl = 0; tl < ttn; ++tl) {
            const float* srca = p.a + ((long long)n * p.t + ts + tl) * V * p.lda + c0;
            const float* srcb = p.b + ((long long)n * p.t + ts + tl) * V * p.ldb + c0;
            float* As = dst + tl * rows_a * ld;
            float* Bs = dst + p.tt * rows_a * ld + tl * rows_b * ld;
            if (p.vec) {
                for (int e = tid; e < ne_a; e += kGramThreads) {
                    const int2 q = tab_a[e];
                    float* d = As + (q.y & 0xFFFFF);
                    if ((q.y >> 20) < cn) cp_async16(d, srca + q.x);
                    else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                for (int e = tid; e < ne_b; e += kGramThreads) {
                    const int2 q = tab_b[e];
                    float* d = Bs + (q.y & 0xFFFFF);
                    if ((q.y >> 20) < cn) cp_async16(d, srcb + q.x);
                    else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
                for (int e = tid; e < ne_a; e += kGramThreads) {
                    const int2 q = tab_a[e];
                    As[q.y & 0xFFFFF] = ((q.y >> 20) < cn) ? __ldg(srca + q.x) : 0.f;
                }
                for (int e = tid; e < ne_b; e += kGramThreads) {
                    const int2 q = tab_b[e];
                    Bs[q.y & 0xFFFFF] = ((q.y >> 20) < cn) ? __ldg(srcb + q.x) : 0.f;
                }
            }
        }
        cp_async_commit();
    };

    if (nstage > 0) issue(0, smem);
    for (int s = 0; s < nstage; ++s) {
        float* cur = smem + (s & 1) * stage_floats;
        if (s + 1 < nstage) { issue(s + 1, smem + ((s + 1) & 1) * stage_floats); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        if (active) {
            const int tb = s / npass;
            const int ts = t0 + tb * p.tt;
            const int ttn = (t1 - ts) < p.tt ? (t1 - ts) : p.tt;
            for (int tl = 0; tl < ttn; ++tl) {
                const float* ab = cur + tl * rows_a * ld;
                const float* bb = cur + tl * rows_b * ld;
                for (int c4 = ks; c4 < cq; c4 += nks) {
                    float4 av[5];
#pragma unroll
                    for (int i = 0; i < 5; ++i) av[i] = *reinterpret_cast<const float4*>(ab + offu[i] + c4 * 4);
#pragma unroll
                    for (int jj = 0; jj < 5; ++jj) {
                        const float4 bv = *reinterpret_cast<const float4*>(bb + offv[jj] + c4 * 4);
#pragma unroll
                        for (int i = 0; i < 5; ++i) {
                            float x = acc[i][jj];
                            x = fmaf(av[i].x, bv.x, x); x = fmaf(av[i].y, bv.y, x);
                            x = fmaf(av[i].z, bv.z, x); x = fmaf(av[i].w, bv.w, x);
                            acc[i][jj] = x;
                        }
                    }
                }
            }
        }
        __syncthreads();            // everyone is done with `cur` before the next issue overwrites it
    }
    // sum the k-slices in a fixed order through shared memory: red[ks][tile][25]
    float* red = smem;
    if (active) {
        float* r = red + ((size_t)ks * tiles + tile) * 25;
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int j = 0; j < 5; ++j) r[i * 5 + j] = acc[i][j];
    }
    __syncthreads();
    float* o = p.out + ((long long)n * p.nchunk + chunk) * p.groups * V * V;
    for (int idx = tid; idx < tiles * 25; idx += kGramThreads) {
        float sum = 0.f;
        for (int k = 0; k < nks; ++k) sum += red[(size_t)k * tiles * 25 + idx];
        const int tl = idx / 25, e = idx - tl * 25;
        const int gg = tl / (nblk * nblk), r = tl % (nblk * nblk);
        const int u = (r / nblk) * 5 + e / 5, w = (r % nblk) * 5 + e % 5;
        if (u < V && w < V) o[((long long)gg * V + u) * V + w] = sum;
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

// CV = 1: scalar channels; CV = 4: one 16-byte channel quad per item; CV = 8: two quads W/2 channels apart per item
// (the lanes of a warp then read contiguous quads -> conflict-free LDS.128, and each matrix row feeds 40 FMAs).
template <int CV>
__global__ void __launch_bounds__(256) joint_mix_kernel(MixArgs p) {
    extern __shared__ __align__(16) float smem[];
    constexpr int NQ = CV == 8 ? 2 : 1;             // quads per item
    constexpr int CE = CV == 1 ? 1 : 4;             // channels per quad
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
        if (CV >= 4) {
            const float4* s4 = reinterpret_cast<const float4*>(src);
            float4* d4 = reinterpret_cast<float4*>(in_s);
            for (int idx = tid; idx < total / 4; idx += 256) d4[idx] = __ldg(s4 + idx);
        } else {
            for (int idx = tid; idx < total; idx += 256) in_s[idx] = __ldg(src + idx);
        }
    }
    __syncthreads();

    const int ogroups = (p.mode == AGCN_MIX_AGG_FWD) ? 3 : (p.mode == AGCN_MIX_AGG_BWD ? 1 : 6);
    const int cq = W / (CE * NQ);                   // items along the channel axis of one group
    const int qstride = NQ == 2 ? W / 2 : 0;        // channel distance between the two quads of an item
    const int items = tn * ogroups * nblk * cq;
    for (int item = tid; item < items; item += 256) {
        const int c = (item % cq) * CE;
        int r = item / cq;
        const int jb = r % nblk; r /= nblk;
        const int og = r % ogroups;
        const int tl = r / ogroups;
        float acc[NQ][5][CE];
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int j = 0; j < 5; ++j)
#pragma unroll
                for (int e = 0; e < CE; ++e) acc[q][j][e] = 0.f;
        const int nterms = (p.mode == AGCN_MIX_AGG_BWD) ? 3 : 1;
        for (int term = 0; term < nterms; ++term) {
            int ig, slot;
            if (p.mode == AGCN_MIX_AGG_FWD) { ig = 0; slot = og; }
            else if (p.mode == AGCN_MIX_AGG_BWD) { ig = term; slot = term; }
            else { ig = og ^ 1; slot = og; }
            const float* xin = in_s + (size_t)tl * V * p.ldin + ig * W + c;
            const float* m = Ms + (size_t)slot * V * mld + jb * 8;
#pragma unroll 5
            for (int i = 0; i < V; ++i) {
                float xv[NQ][CE];
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    if (CE == 4) {
                        const float4 v4 = *reinterpret_cast<const float4*>(xin + (size_t)i * p.ldin + q * qstride);
                        xv[q][0] = v4.x; xv[q][1 % CE] = v4.y; xv[q][2 % CE] = v4.z; xv[q][3 % CE] = v4.w;
                    } else {
                        xv[q][0] = xin[(size_t)i * p.ldin];
                    }
                }
                const float4 m0 = *reinterpret_cast<const float4*>(m + (size_t)i * mld);
                const float m4 = m[(size_t)i * mld + 4];
                const float mv[5] = {m0.x, m0.y, m0.z, m0.w, m4};
#pragma unroll
                for (int q = 0; q < NQ; ++q)
#pragma unroll
                    for (int j = 0; j < 5; ++j)
#pragma unroll
                        for (int e = 0; e < CE; ++e) acc[q][j][e] = fmaf(xv[q][e], mv[j], acc[q][j][e]);
            }
        }
        const int t = t0 + tl;
        // when accumulating, all old values are loaded before the first add / store (one exposed load latency per item)
        float oldv[NQ][5][CE];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const int jj = jb * 5 + j;
            const float* o = p.out + (((long long)n * p.t + t) * V + (jj < V ? jj : 0)) * p.ldout + og * W + c;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                if (p.accumulate && jj < V) {
                    if (CE == 4) {
                        const float4 v4 = *reinterpret_cast<const float4*>(o + q * qstride);
                        oldv[q][j][0] = v4.x; oldv[q][j][1 % CE] = v4.y; oldv[q][j][2 % CE] = v4.z; oldv[q][j][3 % CE] = v4.w;
                    } else {
                        oldv[q][j][0] = o[0];
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < CE; ++e) oldv[q][j][e] = 0.f;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const int jj = jb * 5 + j;
            if (jj >= V) continue;
            float* o = p.out + (((long long)n * p.t + t) * V + jj) * p.ldout + og * W + c;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                if (CE == 4) {
                    *reinterpret_cast<float4*>(o + q * qstride) =
                        make_float4(acc[q][j][0] + oldv[q][j][0], acc[q][j][1 % CE] + oldv[q][j][1 % CE],
                                    acc[q][j][2 % CE] + oldv[q][j][2 % CE], acc[q][j][3 % CE] + oldv[q][j][3 % CE]);
                } else {
                    o[0] = acc[q][j][0] + oldv[q][j][0];
                }
            }
        }
    }
}

}  // namespace agcn

using namespace agcn;

// implemented in joint_big.cu: graphs with more than kMaxV nodes (the 1-D adaptive graph convolution)
int agcn_joint_gram_big(const float* a, const float* b, float* out, int nb, int t, int v, int lda, int ldb, int groups,
                        int offa, int stridea, int offb, int strideb, int width, void* stream);
int agcn_attention_fwd_big(const float* s_part, const float* adj_a, const float* adj_b, float* p, float* g,
                           int nb, int groups, int v, float scale, void* stream);
int agcn_attention_bwd_big(const float* dg_part, const float* p, float* dg_sum, float* ds, float* dadj_b,
                           int nb, int groups, int v, float scale, void* stream);
int agcn_joint_mix_big(const float* in, const float* mats, float* out, int nb, int t, int v, int ldin, int ldout, int width,
                       int mode, int accumulate, void* stream);

// implemented in gram_tc.cu; AGCN_ERR_UNSUPPORTED when the shape is outside the tensor-core path
int agcn_joint_gram_tc(const float* a, const float* b, float* out,
                       int nb, int t, int v, int lda, int ldb, int groups,
                       int offa, int stridea, int offb, int strideb, int width, int nchunk, int split, void* stream);

extern "C" AGCN_API int agcn_joint_gram(const float* a, const float* b, float* out,
                               int nb, int t, int v, int lda, int ldb, int groups,
                               int offa, int stridea, int offb, int strideb, int width, int nchunk, int precision, void* stream) {
    AGCN_REQUIRE(a && b && out, AGCN_ERR_NULL, "agcn_joint_gram: null pointer");
    AGCN_REQUIRE(nb > 0 && t > 0 && v > 0 && groups > 0 && width > 0 && nchunk > 0 && nchunk <= t,
                 AGCN_ERR_BAD_SHAPE, "agcn_joint_gram: bad shape nb=%d t=%d v=%d groups=%d width=%d nchunk=%d", nb, t, v, groups, width, nchunk);
    AGCN_REQUIRE(groups <= 3, AGCN_ERR_UNSUPPORTED, "agcn_joint_gram: groups=%d > 3", groups);
    AGCN_REQUIRE(offa >= 0 && offb >= 0 && offa + (groups - 1) * stridea + width <= lda && offb + (groups - 1) * strideb + width <= ldb,
                 AGCN_ERR_BAD_SHAPE, "agcn_joint_gram: channel window outside the row");
    if (v > kMaxV) {
        AGCN_REQUIRE(nchunk == 1, AGCN_ERR_UNSUPPORTED, "agcn_joint_gram: V=%d > %d takes one chunk (nchunk=%d)", v, kMaxV, nchunk);
        return agcn_joint_gram_big(a, b, out, nb, t, v, lda, ldb, groups, offa, stridea, offb, strideb, width, stream);
    }
    AGCN_REQUIRE(precision >= AGCN_PREC_FP32 && precision <= AGCN_PREC_BF16X3, AGCN_ERR_UNSUPPORTED,
                 "agcn_joint_gram: unknown precision %d", precision);
    if (precision != AGCN_PREC_FP32_FFMA) {
        const int rc = agcn_joint_gram_tc(a, b, out, nb, t, v, lda, ldb, groups, offa, stridea, offb, strideb, width, nchunk,
                                          precision == AGCN_PREC_FP32 || precision == AGCN_PREC_BF16X3 /* the V x V stages run the strict split in both parity modes */, stream);
        if (rc != AGCN_ERR_UNSUPPORTED) return rc;
    }
    GramArgs p{a, b, out, nb, t, v, lda, ldb, groups, offa, stridea, offb, strideb, width, nchunk, 1, 0};
    const int cwp = ((width < kGramCW ? width : kGramCW) + 3) & ~3;
    const int ld = cwp + 4;
    const int ga = stridea == 0 ? 1 : groups;
    const size_t step_bytes = (size_t)(ga + groups) * v * ld * sizeof(float);     // one staged timestep
    const int t_per = (t + nchunk - 1) / nchunk;
    int tt = (int)((28 * 1024) / step_bytes);
    if (tt < 1) tt = 1;
    if (tt > t_per) tt = t_per;
    p.tt = tt;
    p.vec = (width % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 && offa % 4 == 0 && offb % 4 == 0 && stridea % 4 == 0 && strideb % 4 == 0 &&
             aligned16(a) && aligned16(b)) ? 1 : 0;
    const int nblk = (v + 4) / 5, tiles = groups * nblk * nblk;
    const size_t tab_bytes = (size_t)(ga + groups) * v * (p.vec ? cwp / 4 : cwp) * sizeof(int2);   // per-timestep copy table
    size_t smem = 2 * (size_t)tt * step_bytes + tab_bytes;
    const int ntw = (tiles + 31) / 32;
    const size_t red_bytes = (size_t)(kGramWarps / ntw) * tiles * 25 * sizeof(float);   // k-slice partials
    if (smem < red_bytes) smem = red_bytes;
    AGCN_REQUIRE(smem <= 200 * 1024, AGCN_ERR_UNSUPPORTED, "agcn_joint_gram: staged rows need %zu bytes of shared memory", smem);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaFuncSetAttribute(joint_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return fail(AGCN_ERR_CUDA, "agcn_joint_gram: %s", cudaGetErrorString(e));
    joint_gram_kernel<<<nb * nchunk, kGramThreads, smem, s>>>(p);
    return check_launch("agcn_joint_gram");
}

extern "C" AGCN_API int agcn_attention_fwd(const float* s_part, const float* adj_a, const float* adj_b, float* p, float* g,
                                  int nb, int nchunk, int groups, int v, float scale, void* stream) {
    AGCN_REQUIRE(s_part && adj_a && adj_b && p && g, AGCN_ERR_NULL, "agcn_attention_fwd: null pointer");
    AGCN_REQUIRE(nb > 0 && nchunk > 0 && groups > 0 && v > 0, AGCN_ERR_BAD_SHAPE, "agcn_attention_fwd: bad shape");
    if (v > kMaxV) {
        AGCN_REQUIRE(nchunk == 1, AGCN_ERR_UNSUPPORTED, "agcn_attention_fwd: V=%d > %d takes one chunk", v, kMaxV);
        return agcn_attention_fwd_big(s_part, adj_a, adj_b, p, g, nb, groups, v, scale, stream);
    }
    long long total = (long long)nb * groups * v;
    attention_fwd_kernel<<<ceil_div(total, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(s_part, adj_a, adj_b, p, g, nb, nchunk, groups, v, scale);
    return check_launch("agcn_attention_fwd");
}

extern "C" AGCN_API int agcn_attention_bwd(const float* dg_part, const float* p, float* dg_sum, float* ds, float* dadj_b,
                                  int nb, int nchunk, int groups, int v, float scale, void* stream) {
    AGCN_REQUIRE(dg_part && p && dg_sum && ds && dadj_b, AGCN_ERR_NULL, "agcn_attention_bwd: null pointer");
    AGCN_REQUIRE(nb > 0 && nchunk > 0 && groups > 0 && v > 0, AGCN_ERR_BAD_SHAPE, "agcn_attention_bwd: bad shape");
    if (v > kMaxV) {
        AGCN_REQUIRE(nchunk == 1, AGCN_ERR_UNSUPPORTED, "agcn_attention_bwd: V=%d > %d takes one chunk", v, kMaxV);
        return agcn_attention_bwd_big(dg_part, p, dg_sum, ds, dadj_b, nb, groups, v, scale, stream);
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    long long total = (long long)nb * groups * v;
    attention_bwd_kernel<<<ceil_div(total, 128), 128, 0, s>>>(dg_part, p, dg_sum, ds, nb, nchunk, groups, v, scale);
    int rc = check_launch("agcn_attention_bwd");
    if (rc) return rc;
    const int per = groups * v * v;
    sum_over_samples_kernel<<<ceil_div(per, 128), 128, 0, s>>>(dg_sum, dadj_b, nb, per);
    return check_launch("agcn_attention_bwd(dadj_b)");
}

// implemented in mix_tc.cu; AGCN_ERR_UNSUPPORTED when the shape / mode is outside the tensor-core path
size_t agcn_joint_mix_tc_workspace_bytes(int nb);
size_t agcn_joint_mix_tc_colsum_floats(int ldout);
int agcn_joint_mix_tc(const float* in, const float* mats, float* out, float* gp,
                      int nb, int t, int v, int ldin, int ldout, int width, int mode, int accumulate, int split, void* stream,
                      float* colsum, float* colsum_part);

extern "C" AGCN_API size_t agcn_joint_mix_workspace_bytes(int nb) { return nb > 0 ? agcn_joint_mix_tc_workspace_bytes(nb) : 0; }

extern "C" AGCN_API int agcn_joint_mix(const float* in, const float* mats, float* out,
                              int nb, int t, int v, int ldin, int ldout, int width, int mode, int accumulate,
                              int precision, void* workspace, size_t workspace_bytes, void* stream) {
    AGCN_REQUIRE(in && mats && out, AGCN_ERR_NULL, "agcn_joint_mix: null pointer");
    AGCN_REQUIRE(nb > 0 && t > 0 && v > 0 && width > 0, AGCN_ERR_BAD_SHAPE, "agcn_joint_mix: bad shape");
    int need_in, need_out;
    if (mode == AGCN_MIX_AGG_FWD) { need_in = width; need_out = 3 * width; }
    else if (mode == AGCN_MIX_AGG_BWD) { need_in = 3 * width; need_out = width; }
    else if (mode == AGCN_MIX_SCORE_BWD) { need_in = 6 * width; need_out = 6 * width; }
    else return fail(AGCN_ERR_UNSUPPORTED, "agcn_joint_mix: unknown mode %d", mode);
    AGCN_REQUIRE(ldin == need_in && ldout == need_out, AGCN_ERR_BAD_SHAPE,
                 "agcn_joint_mix: mode %d expects ldin=%d ldout=%d, got %d %d", mode, need_in, need_out, ldin, ldout);
    AGCN_REQUIRE(precision >= AGCN_PREC_FP32 && precision <= AGCN_PREC_BF16X3, AGCN_ERR_UNSUPPORTED,
                 "agcn_joint_mix: unknown precision %d", precision);
    if (v > kMaxV) return agcn_joint_mix_big(in, mats, out, nb, t, v, ldin, ldout, width, mode, accumulate, stream);
    if (precision != AGCN_PREC_FP32_FFMA && workspace != nullptr && workspace_bytes >= agcn_joint_mix_tc_workspace_bytes(nb)) {
        const int rc = agcn_joint_mix_tc(in, mats, out, static_cast<float*>(workspace), nb, t, v, ldin, ldout, width, mode, accumulate,
                                         precision == AGCN_PREC_FP32 || precision == AGCN_PREC_BF16X3 /* the V x V stages run the strict split in both parity modes */, stream,
                                         nullptr, nullptr);
        if (rc != AGCN_ERR_UNSUPPORTED) return rc;
    }
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
    static const bool mix8 = probe_env("AGCN_MIX8") != nullptr;      // 2-quad items: measured slower on B200 (114 registers), kept for experiments
    if (vec && width % 8 == 0 && mix8) {
        e = cudaFuncSetAttribute(joint_mix_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return fail(AGCN_ERR_CUDA, "agcn_joint_mix: %s", cudaGetErrorString(e));
        joint_mix_kernel<8><<<nb * tiles_t, 256, smem, s>>>(p);
    } else if (vec) {
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

extern "C" AGCN_API size_t agcn_joint_mix_score_bwd_colsum_workspace_bytes(int nb, int width) {
    if (nb <= 0 || width <= 0) return 0;
    return agcn_joint_mix_tc_workspace_bytes(nb) + agcn_joint_mix_tc_colsum_floats(6 * width) * sizeof(float);
}

// AGCN_MIX_SCORE_BWD of agcn_joint_mix that also leaves colsum[6 * width] = sum over all (nb, t, v) rows of `out`: the bias gradient of
// the theta / phi convolutions (agcn.py:104-105) out of the epilogue that writes d theta / d phi, instead of a pass over that tensor.
// Tensor-core path only: AGCN_ERR_UNSUPPORTED for shapes it does not take (run agcn_joint_mix and sum the columns separately then).
extern "C" AGCN_API int agcn_joint_mix_score_bwd_colsum(const float* in, const float* mats, float* out, float* colsum,
                                                        int nb, int t, int v, int width, int precision,
                                                        void* workspace, size_t workspace_bytes, void* stream) {
    AGCN_REQUIRE(in && mats && out && colsum && workspace, AGCN_ERR_NULL, "agcn_joint_mix_score_bwd_colsum: null pointer");
    AGCN_REQUIRE(nb > 0 && t > 0 && v > 0 && width > 0, AGCN_ERR_BAD_SHAPE, "agcn_joint_mix_score_bwd_colsum: bad shape");
    AGCN_REQUIRE(precision >= AGCN_PREC_FP32 && precision <= AGCN_PREC_BF16X3, AGCN_ERR_UNSUPPORTED,
                 "agcn_joint_mix_score_bwd_colsum: unknown precision %d", precision);
    AGCN_REQUIRE(workspace_bytes >= agcn_joint_mix_score_bwd_colsum_workspace_bytes(nb, width), AGCN_ERR_WORKSPACE,
                 "agcn_joint_mix_score_bwd_colsum: workspace too small");
    AGCN_REQUIRE(aligned16(workspace), AGCN_ERR_MISALIGNED, "agcn_joint_mix_score_bwd_colsum: workspace not 16-byte aligned");
    if (v > kMaxV || precision == AGCN_PREC_FP32_FFMA) return AGCN_ERR_UNSUPPORTED;       // (quiet: the caller falls back)
    float* gp = static_cast<float*>(workspace);
    float* part = gp + agcn_joint_mix_tc_workspace_bytes(nb) / sizeof(float);
    return agcn_joint_mix_tc(in, mats, out, gp, nb, t, v, 6 * width, 6 * width, width, AGCN_MIX_SCORE_BWD, 0,
                             precision == AGCN_PREC_FP32 || precision == AGCN_PREC_BF16X3, stream, colsum, part);
}
