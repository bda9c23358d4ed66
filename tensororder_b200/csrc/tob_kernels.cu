// Hand-written sm_100a kernels of the contraction executor.
//
// Every join arrives in canonical form (tob_internal.h):
//     C[pdep(mi, mask_m) | pdep(ni, mask_n)] = sum_{kk < 2^k} A[mi << k | kk] * B[ni << k | kk]
// i.e. a "TN" GEMM with both operands K-contiguous and an interleaving scatter on the output.
//
//  * k_gemm_dmma   — compute-bound joins.  FP64 has no tcgen05 kind on Blackwell, so the tensor-core
//                    path is DMMA (mma.sync.m8n8k4.f64; SASS DMMA.8x8x4).  CTA tile TM x TN x 16,
//                    multi-stage cp.async (LDGSTS) pipeline of 128-byte rows, padded K-major shared
//                    memory (conflict-free fragment loads), split-K across CTAs for short grids.
//  * k_generic_*   — bandwidth/latency-bound joins: one thread, one warp or one CTA per output element,
//                    vectorised K-contiguous loads, warp-shuffle reductions, deterministic split-K.
//  * k_permute     — stand-alone index (address-bit) permutation through shared memory.
//  * k_begin_slice / k_accum — slice bookkeeping kept on the device so a slice replays as a CUDA graph.
#include <cstdio>
#include <cstdlib>

#include "tob_kernels.cuh"

namespace tob {

// ------------------------------------------------------------------------------------------------
// bit scatter / gather by runs
// ------------------------------------------------------------------------------------------------
BitRuns make_runs(uint64_t mask) {
    BitRuns r;
    r.n = 0;
    for (int i = 0; i < 24; i++) r.dst[i] = r.src[i] = r.len[i] = 0;
    int src = 0;
    int p = 0;
    while (p < 64) {
        if (!((mask >> p) & 1)) { p++; continue; }
        int q = p;
        while (q < 64 && ((mask >> q) & 1)) q++;
        r.dst[r.n] = (uint8_t)p;
        r.src[r.n] = (uint8_t)src;
        r.len[r.n] = (uint8_t)(q - p);
        r.n++;
        src += q - p;
        p = q;
    }
    return r;
}

__device__ __forceinline__ unsigned long long pext_runs(unsigned long long x, const BitRuns& r) {
    unsigned long long o = 0;
    for (int i = 0; i < r.n; i++) o |= ((x >> r.dst[i]) & ((1ull << r.len[i]) - 1ull)) << r.src[i];
    return o;
}
__device__ __forceinline__ unsigned long long pdep_runs(unsigned long long x, const BitRuns& r) {
    unsigned long long o = 0;
    for (int i = 0; i < r.n; i++) o |= ((x >> r.src[i]) & ((1ull << r.len[i]) - 1ull)) << r.dst[i];
    return o;
}

__device__ __forceinline__ const double* operand_base(const double* base, const long long* leaf_off, int leaf) {
    return (leaf >= 0) ? base + leaf_off[leaf] : base;
}

// x mod p for an integer-valued 0 <= x < 2^53 (exact: the quotient estimate is off by at most one)
__device__ __forceinline__ double mod_reduce(double x, double p, double inv_p) {
    const double q = floor(x * inv_p);
    double r = fma(-q, p, x);
    if (r < 0.0) r += p;
    if (r >= p) r -= p;
    return r;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------
// generic kernels
// ------------------------------------------------------------------------------------------------
// one thread per output element; K <= 64
__global__ void __launch_bounds__(256) k_generic_t1(KParams p) {
    const double* A = operand_base(p.a, p.leaf_off, p.a_leaf);
    const double* B = operand_base(p.b, p.leaf_off, p.b_leaf);
    const unsigned long long total = 1ull << (p.m + p.n);
    const int k = p.k;
    for (unsigned long long c = blockIdx.x * 256ull + threadIdx.x; c < total; c += (unsigned long long)gridDim.x * 256ull) {
        const unsigned long long mi = pext_runs(c, p.runs_m);
        const unsigned long long ni = pext_runs(c, p.runs_n);
        const double* ar = A + (mi << k);
        const double* br = B + (ni << k);
        double s;
        if (k == 0) {
            s = ar[0] * br[0];
        } else {
            const double2* a2 = reinterpret_cast<const double2*>(ar);
            const double2* b2 = reinterpret_cast<const double2*>(br);
            const int K2 = 1 << (k - 1);
            s = 0.0;
            for (int i = 0; i < K2; i++) {
                const double2 x = a2[i], y = b2[i];
                s = fma(x.x, y.x, s);
                s = fma(x.y, y.y, s);
            }
        }
        if (p.modp > 0.0) s = mod_reduce(s, p.modp, p.inv_modp);  // K <= 64 products < 2^46: exact so far
        p.c[c] = s;
    }
}

// Streaming form of k_generic_t1 for large outputs: one thread produces FOUR consecutive outputs (one
// 32-byte store), so the two pext's are paid once per four outputs and the A / B rows of neighbouring
// outputs are fetched with 16-byte loads.  KL = log2 K (compile time, 0..3) or -1 for a run-time K <= 64;
// LOW = mask_m & 3 says which of the two lowest C address bits belong to the M side.
template <int KL, int LOW>
__global__ void __launch_bounds__(256) k_generic_t1x4(KParams p) {
    const double* A = operand_base(p.a, p.leaf_off, p.a_leaf);
    const double* B = operand_base(p.b, p.leaf_off, p.b_leaf);
    const unsigned long long quads = 1ull << (p.m + p.n - 2);
    const int k = (KL >= 0) ? KL : p.k;
    for (unsigned long long q = blockIdx.x * 256ull + threadIdx.x; q < quads; q += (unsigned long long)gridDim.x * 256ull) {
        const unsigned long long c0 = q << 2;
        const unsigned long long mi0 = pext_runs(c0, p.runs_m);
        const unsigned long long ni0 = pext_runs(c0, p.runs_n);
        double out[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int dm = (LOW == 3) ? j : (LOW == 1) ? (j & 1) : (LOW == 2) ? (j >> 1) : 0;
            const int dn = (LOW == 0) ? j : (LOW == 1) ? (j >> 1) : (LOW == 2) ? (j & 1) : 0;
            const double* ar = A + ((mi0 + dm) << k);
            const double* br = B + ((ni0 + dn) << k);
            double s;
            if (KL == 0) {
                s = ar[0] * br[0];
            } else if (KL > 0) {
                const double2* a2 = reinterpret_cast<const double2*>(ar);
                const double2* b2 = reinterpret_cast<const double2*>(br);
                s = 0.0;
#pragma unroll
                for (int i = 0; i < (1 << (KL > 0 ? KL - 1 : 0)); i++) {
                    const double2 x = a2[i], y = b2[i];
                    s = fma(x.x, y.x, s);
                    s = fma(x.y, y.y, s);
                }
            } else {
                if (k == 0) {
                    s = ar[0] * br[0];
                } else {
                    const double2* a2 = reinterpret_cast<const double2*>(ar);
                    const double2* b2 = reinterpret_cast<const double2*>(br);
                    const int K2 = 1 << (k > 0 ? k - 1 : 0);
                    s = 0.0;
                    for (int i = 0; i < K2; i++) {
                        const double2 x = a2[i], y = b2[i];
                        s = fma(x.x, y.x, s);
                        s = fma(x.y, y.y, s);
                    }
                }
            }
            out[j] = (p.modp > 0.0) ? mod_reduce(s, p.modp, p.inv_modp) : s;
        }
        double2* dst = reinterpret_cast<double2*>(p.c + c0);
        dst[0] = make_double2(out[0], out[1]);
        dst[1] = make_double2(out[2], out[3]);
    }
}

template <int KL>
static void launch_t1x4(const KParams& p, unsigned grid, cudaStream_t stream) {
    switch ((int)(p.mask_m & 3ull)) {
        case 0: k_generic_t1x4<KL, 0><<<grid, 256, 0, stream>>>(p); break;
        case 1: k_generic_t1x4<KL, 1><<<grid, 256, 0, stream>>>(p); break;
        case 2: k_generic_t1x4<KL, 2><<<grid, 256, 0, stream>>>(p); break;
        default: k_generic_t1x4<KL, 3><<<grid, 256, 0, stream>>>(p); break;
    }
}

// one warp per output element; 128 <= K
__global__ void __launch_bounds__(256) k_generic_t32(KParams p) {
    const double* A = operand_base(p.a, p.leaf_off, p.a_leaf);
    const double* B = operand_base(p.b, p.leaf_off, p.b_leaf);
    const unsigned long long total = 1ull << (p.m + p.n);
    const int k = p.k;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long K2 = 1ull << (k - 1);
    for (unsigned long long c = blockIdx.x * 8ull + warp; c < total; c += (unsigned long long)gridDim.x * 8ull) {
        const unsigned long long mi = pext_runs(c, p.runs_m);
        const unsigned long long ni = pext_runs(c, p.runs_n);
        const double2* a2 = reinterpret_cast<const double2*>(A + (mi << k));
        const double2* b2 = reinterpret_cast<const double2*>(B + (ni << k));
        double s0 = 0.0, s1 = 0.0;
        for (unsigned long long i = lane; i < K2; i += 32) {
            const double2 x = a2[i], y = b2[i];
            s0 = fma(x.x, y.x, s0);
            s1 = fma(x.y, y.y, s1);
        }
        double s;
        if (p.modp > 0.0) {  // <= 32 products per accumulator (k <= 11), then 32 residues per warp sum
            s = warp_sum(mod_reduce(s0, p.modp, p.inv_modp) + mod_reduce(s1, p.modp, p.inv_modp));
            s = mod_reduce(s, p.modp, p.inv_modp);
        } else {
            s = warp_sum(s0 + s1);
        }
        if (lane == 0) p.c[c] = s;
    }
}

// one CTA per (output element, K chunk); K >= 128
__global__ void __launch_bounds__(256) k_generic_t256(KParams p) {
    __shared__ double red[8];
    const double* A = operand_base(p.a, p.leaf_off, p.a_leaf);
    const double* B = operand_base(p.b, p.leaf_off, p.b_leaf);
    const int k = p.k, ks = p.ksplit_log2;
    const unsigned long long outs = 1ull << (p.m + p.n);
    const unsigned long long work = outs << ks;
    const unsigned long long chunk2 = 1ull << (k - ks - 1);  // double2 elements per chunk
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (unsigned long long w = blockIdx.x; w < work; w += gridDim.x) {
        const unsigned long long c = w & (outs - 1);
        const unsigned long long split = w >> (p.m + p.n);
        const unsigned long long mi = pext_runs(c, p.runs_m);
        const unsigned long long ni = pext_runs(c, p.runs_n);
        const double2* a2 = reinterpret_cast<const double2*>(A + (mi << k)) + split * chunk2;
        const double2* b2 = reinterpret_cast<const double2*>(B + (ni << k)) + split * chunk2;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        unsigned long long i = threadIdx.x;
        for (; i + 768 < chunk2; i += 1024) {
            const double2 x0 = a2[i], y0 = b2[i];
            const double2 x1 = a2[i + 256], y1 = b2[i + 256];
            const double2 x2 = a2[i + 512], y2 = b2[i + 512];
            const double2 x3 = a2[i + 768], y3 = b2[i + 768];
            s0 = fma(x0.x, y0.x, s0); s0 = fma(x0.y, y0.y, s0);
            s1 = fma(x1.x, y1.x, s1); s1 = fma(x1.y, y1.y, s1);
            s2 = fma(x2.x, y2.x, s2); s2 = fma(x2.y, y2.y, s2);
            s3 = fma(x3.x, y3.x, s3); s3 = fma(x3.y, y3.y, s3);
            if (p.modp > 0.0 && (((i >> 10) & 31) == 31)) {  // every 32 iterations: 64 products per accumulator
                s0 = mod_reduce(s0, p.modp, p.inv_modp); s1 = mod_reduce(s1, p.modp, p.inv_modp);
                s2 = mod_reduce(s2, p.modp, p.inv_modp); s3 = mod_reduce(s3, p.modp, p.inv_modp);
            }
        }
        for (; i < chunk2; i += 256) {
            const double2 x = a2[i], y = b2[i];
            s0 = fma(x.x, y.x, s0);
            s0 = fma(x.y, y.y, s0);
        }
        if (p.modp > 0.0) {
            s0 = mod_reduce(s0, p.modp, p.inv_modp); s1 = mod_reduce(s1, p.modp, p.inv_modp);
            s2 = mod_reduce(s2, p.modp, p.inv_modp); s3 = mod_reduce(s3, p.modp, p.inv_modp);
        }
        double s = warp_sum((s0 + s1) + (s2 + s3));  // 128 residues: far below 2^53
        __syncthreads();  // red[] reuse across iterations of the work loop
        if (lane == 0) red[warp] = s;
        __syncthreads();
        if (warp == 0) {
            s = (lane < 8) ? red[lane] : 0.0;
            s = warp_sum(s);
            if (p.modp > 0.0) s = mod_reduce(s, p.modp, p.inv_modp);
            if (lane == 0) {
                if (ks > 0) p.ws[split * outs + c] = s;
                else p.c[c] = s;
            }
        }
    }
}

// C[o] = sum_j ws[j * outs + o], j ascending (deterministic)
__global__ void __launch_bounds__(256) k_reduce_splits(const double* __restrict__ ws, double* __restrict__ c,
                                                      unsigned long long outs, int nsplit, double modp, double inv_modp) {
    for (unsigned long long o = blockIdx.x * 256ull + threadIdx.x; o < outs; o += (unsigned long long)gridDim.x * 256ull) {
        double s = ws[o];
        for (int j = 1; j < nsplit; j++) s += ws[(unsigned long long)j * outs + o];
        c[o] = (modp > 0.0) ? mod_reduce(s, modp, inv_modp) : s;  // <= 1024 residues: exact
    }
}

// few outputs, many splits (the final dot products): one warp per output, lane-strided partial sums, then
// a fixed shuffle tree — deterministic, and 32 independent add chains instead of one
__global__ void __launch_bounds__(256) k_reduce_splits_warp(const double* __restrict__ ws, double* __restrict__ c,
                                                           unsigned long long outs, int nsplit, double modp, double inv_modp) {
    const int lane = threadIdx.x & 31;
    const unsigned long long o = blockIdx.x * 8ull + (threadIdx.x >> 5);
    if (o >= outs) return;
    double s = 0.0;
    for (int j = lane; j < nsplit; j += 32) s += ws[(unsigned long long)j * outs + o];  // <= 32 residues per lane
    s = warp_sum(s);                                                                    // <= 1024 residues: exact
    if (lane == 0) c[o] = (modp > 0.0) ? mod_reduce(s, modp, inv_modp) : s;
}

// ------------------------------------------------------------------------------------------------
// DMMA GEMM
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
// 16-byte cp.async whose source contributes only src_bytes (0 or 16); the rest of the chunk is zero-filled
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, int src_bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// K step per pipeline stage (doubles) and the padded shared-memory row stride: stride == 4 (mod 16)
// makes the 8x4 DMMA fragment loads (LDS.64) conflict-free for both TK = 16 and TK = 32.
// EXACT: the residue-arithmetic instantiation (entry type bigint); the float64 instantiation carries none
// of its code (the extra live values cost registers the two-CTAs-per-SM budget does not have)
template <int TM_LOG2, int TN_LOG2, int WM, int WN, int TK, int STAGES, int MINB, bool EXACT = false>
__global__ void __launch_bounds__(WM * WN * 32, MINB) k_gemm_dmma(KParams p) {
    constexpr int TM = 1 << TM_LOG2, TN = 1 << TN_LOG2;
    constexpr int NT = WM * WN * 32;             // threads per CTA
    constexpr int WTM = TM / WM, WTN = TN / WN;  // warp tile
    constexpr int MB = WTM / 8, NB = WTN / 8;    // 8x8 DMMA blocks per warp tile
    constexpr int LDS = TK + 4;                  // row stride == 4 (mod 16): conflict-free LDS.64 fragments
    constexpr int CHUNKS = TK / 2;               // 16-byte chunks per tile row
    constexpr int RPP = NT / CHUNKS;             // rows per loader pass
    constexpr int K4 = TK / 4;
    static_assert(TM % RPP == 0 && TN % RPP == 0, "loader passes");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* As = reinterpret_cast<double*>(smem_raw);
    double* Bs = As + STAGES * TM * LDS;
    unsigned long long* cM = reinterpret_cast<unsigned long long*>(Bs + STAGES * TN * LDS);
    unsigned long long* cN = cM + TM;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp % WM, wn = warp / WM;
    const int k = p.k, ks = p.ksplit_log2;

    for (int i = tid; i < TM; i += NT) cM[i] = pdep_runs((unsigned long long)i, p.runs_m);
    for (int i = tid; i < TN; i += NT) cN[i] = pdep_runs((unsigned long long)i, p.runs_n);

    // ---- work decode: 1-D grid over (split, tile), tiles rasterised in groups of 16 M-tiles ----
    const unsigned long long tilesM = 1ull << (p.m - TM_LOG2), tilesN = 1ull << (p.n - TN_LOG2);
    const unsigned long long tiles = tilesM * tilesN;
    const unsigned long long id = blockIdx.x;
    const unsigned long long split = id / tiles, tid_in = id % tiles;
    const unsigned long long group = tilesM < 16 ? tilesM : 16;
    const unsigned long long per_group = group * tilesN;
    const unsigned long long gidx = tid_in / per_group, r = tid_in % per_group;
    const unsigned long long tile_m = gidx * group + (r % group), tile_n = r / group;

    const unsigned long long Ksplit = (1ull << k) >> ks;
    const int KT = (int)((Ksplit + TK - 1) / TK);
    const bool partial = Ksplit < (unsigned long long)TK;  // K = 2, 4, 8 (< one K step): zero-fill the tail chunks
    const int k4_end = partial ? (int)((Ksplit + 3) / 4) : K4;
    const double* A = operand_base(p.a, p.leaf_off, p.a_leaf) + ((tile_m << TM_LOG2) << k) + split * Ksplit;
    const double* B = operand_base(p.b, p.leaf_off, p.b_leaf) + ((tile_n << TN_LOG2) << k) + split * Ksplit;

    const int chunk = tid % CHUNKS, row0 = tid / CHUNKS;
    const double* a_src = A + ((unsigned long long)row0 << k) + chunk * 2;
    const double* b_src = B + ((unsigned long long)row0 << k) + chunk * 2;
    const unsigned long long pass_stride = (unsigned long long)RPP << k;
    const int dst_off = row0 * LDS + chunk * 2;
    auto load_stage = [&](int s, int kt) {
        double* as = As + s * TM * LDS + dst_off;
        double* bs = Bs + s * TN * LDS + dst_off;
        const double* ag = a_src + kt * TK;
        const double* bg = b_src + kt * TK;
        if (partial) {
            const int nbytes = ((unsigned long long)(chunk * 2) < Ksplit) ? 16 : 0;
            const int back = nbytes ? 0 : chunk * 2;  // keep the (unused) source address inside the row
#pragma unroll
            for (int i = 0; i < TM / RPP; i++) cp_async16_zfill(as + i * RPP * LDS, ag + i * pass_stride - back, nbytes);
#pragma unroll
            for (int i = 0; i < TN / RPP; i++) cp_async16_zfill(bs + i * RPP * LDS, bg + i * pass_stride - back, nbytes);
            return;
        }
#pragma unroll
        for (int i = 0; i < TM / RPP; i++) cp_async16(as + i * RPP * LDS, ag + i * pass_stride);
#pragma unroll
        for (int i = 0; i < TN / RPP; i++) cp_async16(bs + i * RPP * LDS, bg + i * pass_stride);
    };

    double acc[MB][NB][2];
#pragma unroll
    for (int i = 0; i < MB; i++)
#pragma unroll
        for (int j = 0; j < NB; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < KT) load_stage(s, s);
        cp_async_commit();
    }
    const int frag_off_a = (wm * WTM + g) * LDS + t;
    const int frag_off_b = (wn * WTN + g) * LDS + t;
    double af[2][MB], bf[2][NB];  // register double buffer: fragments of step k4+1 load while k4 computes
    for (int kt = 0; kt < KT; kt++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nk = kt + STAGES - 1;
        if (nk < KT) load_stage(nk % STAGES, nk);
        cp_async_commit();
        const double* as = As + (kt % STAGES) * TM * LDS + frag_off_a;
        const double* bs = Bs + (kt % STAGES) * TN * LDS + frag_off_b;
#pragma unroll
        for (int i = 0; i < MB; i++) af[0][i] = as[i * 8 * LDS];
#pragma unroll
        for (int j = 0; j < NB; j++) bf[0][j] = bs[j * 8 * LDS];
#pragma unroll
        for (int k4 = 0; k4 < K4; k4++) {
            if (k4 >= k4_end) break;  // K = 2, 4, 8: the zero-filled tail of the K step adds nothing
            const int cur = k4 & 1, nxt = cur ^ 1;
            if (k4 + 1 < K4) {
#pragma unroll
                for (int i = 0; i < MB; i++) af[nxt][i] = as[i * 8 * LDS + (k4 + 1) * 4];
#pragma unroll
                for (int j = 0; j < NB; j++) bf[nxt][j] = bs[j * 8 * LDS + (k4 + 1) * 4];
            }
#pragma unroll
            for (int i = 0; i < MB; i++)
#pragma unroll
                for (int j = 0; j < NB; j++) dmma884(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
        }
        // exact mode: at most 128 products (< 2^46 each) per accumulator between reductions
        if (EXACT && ((kt + 1) % (128 / TK) == 0 || kt + 1 == KT)) {
#pragma unroll
            for (int i = 0; i < MB; i++)
#pragma unroll
                for (int j = 0; j < NB; j++) {
                    acc[i][j][0] = mod_reduce(acc[i][j][0], p.modp, p.inv_modp);
                    acc[i][j][1] = mod_reduce(acc[i][j][1], p.modp, p.inv_modp);
                }
        }
    }
    cp_async_wait<0>();

    // ---- epilogue: scatter the tile into C (or the split-K workspace) ----
    double* Cout = (ks > 0) ? p.ws + (split << (p.m + p.n)) : p.c;
    const unsigned long long cbase = pdep_runs(tile_m << TM_LOG2, p.runs_m) | pdep_runs(tile_n << TN_LOG2, p.runs_n);
    const bool vec = (p.mask_n & 1ull) != 0;  // ni bit 0 is C address bit 0: the two fragment columns are adjacent
#pragma unroll
    for (int i = 0; i < MB; i++) {
        const unsigned long long rbase = cbase | cM[wm * WTM + i * 8 + g];
#pragma unroll
        for (int j = 0; j < NB; j++) {
            const int col = wn * WTN + j * 8 + 2 * t;
            if (vec) {
                *reinterpret_cast<double2*>(Cout + (rbase | cN[col])) = make_double2(acc[i][j][0], acc[i][j][1]);
            } else {
                Cout[rbase | cN[col]] = acc[i][j][0];
                Cout[rbase | cN[col + 1]] = acc[i][j][1];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Persistent variant for SHORT K (K <= 2 * TK per split): a CTA walks tiles  blockIdx.x, blockIdx.x + gridDim.x, ...
// and its cp.async ring runs across tile boundaries, so the operands of the next tile(s) are in flight while
// the current tile computes and scatters its 64 KB of output.  With one K step per tile the one-tile-per-CTA
// kernel is a serial load -> DMMA -> store per CTA (two CTAs per SM are the only overlap): store-bound joins
// (k <= 5) reach 0.65 of HBM there.  256 threads at two CTAs per SM leave 128 registers, so the extra tile
// state costs no spills (the warp-specialised kernel at 320 threads has no such room: DESIGN.md §4).
// ------------------------------------------------------------------------------------------------
template <int TM_LOG2, int TN_LOG2, int WM, int WN, int TK, int STAGES, int MINB>
__global__ void __launch_bounds__(WM * WN * 32, MINB) k_gemm_dmma_p(KParams p) {
    constexpr int TM = 1 << TM_LOG2, TN = 1 << TN_LOG2;
    constexpr int NT = WM * WN * 32;
    constexpr int WTM = TM / WM, WTN = TN / WN;
    constexpr int MB = WTM / 8, NB = WTN / 8;
    constexpr int LDS = TK + 4;
    constexpr int CHUNKS = TK / 2;
    constexpr int RPP = NT / CHUNKS;
    constexpr int K4 = TK / 4;
    static_assert(TM % RPP == 0 && TN % RPP == 0, "loader passes");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* As = reinterpret_cast<double*>(smem_raw);
    double* Bs = As + STAGES * TM * LDS;
    unsigned long long* cM = reinterpret_cast<unsigned long long*>(Bs + STAGES * TN * LDS);
    unsigned long long* cN = cM + TM;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp % WM, wn = warp / WM;
    const int k = p.k, ks = p.ksplit_log2;

    for (int i = tid; i < TM; i += NT) cM[i] = pdep_runs((unsigned long long)i, p.runs_m);
    for (int i = tid; i < TN; i += NT) cN[i] = pdep_runs((unsigned long long)i, p.runs_n);

    const unsigned long long tilesM = 1ull << (p.m - TM_LOG2), tilesN = 1ull << (p.n - TN_LOG2);
    const unsigned long long tiles = tilesM * tilesN;
    const unsigned long long total = tiles << ks;
    const unsigned long long group = (tilesM >> p.raster_group_log2) ? (1ull << p.raster_group_log2) : tilesM;
    const unsigned long long per_group = group * tilesN;
    const unsigned long long Ksplit = (1ull << k) >> ks;
    const int KT = (int)((Ksplit + TK - 1) / TK);
    const bool partial = Ksplit < (unsigned long long)TK;
    const int k4_end = partial ? (int)((Ksplit + 3) / 4) : K4;
    const unsigned long long nmine = total > blockIdx.x ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long nsteps = (long long)nmine * KT;
    const double* Abase = operand_base(p.a, p.leaf_off, p.a_leaf);
    const double* Bbase = operand_base(p.b, p.leaf_off, p.b_leaf);

    // same rasterisation as k_gemm_dmma: id -> (split, tile_m, tile_n); every count is a power of two
    const int tiles_log2 = (p.m - TM_LOG2) + (p.n - TN_LOG2);
    const int group_log2 = (p.m - TM_LOG2) < p.raster_group_log2 ? (p.m - TM_LOG2) : p.raster_group_log2;
    const int pg_log2 = group_log2 + (p.n - TN_LOG2);
    auto decode = [&](unsigned long long id, unsigned long long& split, unsigned long long& tile_m, unsigned long long& tile_n) {
        split = id >> tiles_log2;
        const unsigned long long tid_in = id & (tiles - 1);
        const unsigned long long gidx = tid_in >> pg_log2, r = tid_in & (per_group - 1);
        tile_m = (gidx << group_log2) + (r & (group - 1));
        tile_n = r >> group_log2;
    };

    // ---- loader: runs STAGES-1 steps ahead of the math, across tile boundaries ----
    const int chunk = tid % CHUNKS, row0 = tid / CHUNKS;
    const unsigned long long pass_stride = (unsigned long long)RPP << k;
    const int dst_off = row0 * LDS + chunk * 2;
    unsigned long long l_id = blockIdx.x;
    int l_kt = 0;
    const double *a_src = nullptr, *b_src = nullptr;
    auto set_load_tile = [&](unsigned long long id) {
        unsigned long long split, tile_m, tile_n;
        decode(id, split, tile_m, tile_n);
        a_src = Abase + ((tile_m << TM_LOG2) << k) + split * Ksplit + ((unsigned long long)row0 << k) + chunk * 2;
        b_src = Bbase + ((tile_n << TN_LOG2) << k) + split * Ksplit + ((unsigned long long)row0 << k) + chunk * 2;
    };
    auto load_step = [&](int s) {
        double* as = As + s * TM * LDS + dst_off;
        double* bs = Bs + s * TN * LDS + dst_off;
        const double* ag = a_src + l_kt * TK;
        const double* bg = b_src + l_kt * TK;
        if (partial) {
            const int nbytes = ((unsigned long long)(chunk * 2) < Ksplit) ? 16 : 0;
            const int back = nbytes ? 0 : chunk * 2;
#pragma unroll
            for (int i = 0; i < TM / RPP; i++) cp_async16_zfill(as + i * RPP * LDS, ag + i * pass_stride - back, nbytes);
#pragma unroll
            for (int i = 0; i < TN / RPP; i++) cp_async16_zfill(bs + i * RPP * LDS, bg + i * pass_stride - back, nbytes);
        } else {
#pragma unroll
            for (int i = 0; i < TM / RPP; i++) cp_async16(as + i * RPP * LDS, ag + i * pass_stride);
#pragma unroll
            for (int i = 0; i < TN / RPP; i++) cp_async16(bs + i * RPP * LDS, bg + i * pass_stride);
        }
        if (++l_kt == KT) {
            l_kt = 0;
            l_id += gridDim.x;
            if (l_id < total) set_load_tile(l_id);
        }
    };

    double acc[MB][NB][2];
#pragma unroll
    for (int i = 0; i < MB; i++)
#pragma unroll
        for (int j = 0; j < NB; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    if (nmine > 0) set_load_tile(l_id);
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nsteps) load_step(s);
        cp_async_commit();
    }
    const int frag_off_a = (wm * WTM + g) * LDS + t;
    const int frag_off_b = (wn * WTN + g) * LDS + t;
    const bool vec = (p.mask_n & 1ull) != 0;
    unsigned long long c_id = blockIdx.x;
    int c_kt = 0;
    double af[2][MB], bf[2][NB];
    for (long long s = 0; s < nsteps; s++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const long long nk = s + STAGES - 1;
        if (nk < nsteps) load_step((int)(nk % STAGES));
        cp_async_commit();
        const double* as = As + (int)(s % STAGES) * TM * LDS + frag_off_a;
        const double* bs = Bs + (int)(s % STAGES) * TN * LDS + frag_off_b;
#pragma unroll
        for (int i = 0; i < MB; i++) af[0][i] = as[i * 8 * LDS];
#pragma unroll
        for (int j = 0; j < NB; j++) bf[0][j] = bs[j * 8 * LDS];
#pragma unroll
        for (int k4 = 0; k4 < K4; k4++) {
            if (k4 >= k4_end) break;
            const int cur = k4 & 1, nxt = cur ^ 1;
            if (k4 + 1 < K4) {
#pragma unroll
                for (int i = 0; i < MB; i++) af[nxt][i] = as[i * 8 * LDS + (k4 + 1) * 4];
#pragma unroll
                for (int j = 0; j < NB; j++) bf[nxt][j] = bs[j * 8 * LDS + (k4 + 1) * 4];
            }
#pragma unroll
            for (int i = 0; i < MB; i++)
#pragma unroll
                for (int j = 0; j < NB; j++) dmma884(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
        }
        if (++c_kt == KT) {
            // ---- this tile is complete: scatter it (the loads of the next tiles are already in flight) ----
            unsigned long long split, tile_m, tile_n;
            decode(c_id, split, tile_m, tile_n);
            double* Cout = (ks > 0) ? p.ws + (split << (p.m + p.n)) : p.c;
            const unsigned long long cbase = pdep_runs(tile_m << TM_LOG2, p.runs_m) | pdep_runs(tile_n << TN_LOG2, p.runs_n);
#pragma unroll
            for (int i = 0; i < MB; i++) {
                const unsigned long long rbase = cbase | cM[wm * WTM + i * 8 + g];
#pragma unroll
                for (int j = 0; j < NB; j++) {
                    const int col = wn * WTN + j * 8 + 2 * t;
                    if (vec) {
                        *reinterpret_cast<double2*>(Cout + (rbase | cN[col])) = make_double2(acc[i][j][0], acc[i][j][1]);
                    } else {
                        Cout[rbase | cN[col]] = acc[i][j][0];
                        Cout[rbase | cN[col + 1]] = acc[i][j][1];
                    }
                    acc[i][j][0] = acc[i][j][1] = 0.0;
                }
            }
            c_kt = 0;
            c_id += gridDim.x;
        }
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// K <= 16 (ONE K step per tile — every join the persistent kernel is dispatched for): the same persistent tile walk,
// but a warp finishes its 32x32 block RP rows of 8 at a time and stores each group of rows while the DMMAs of the
// next group issue.  In k_gemm_dmma_p a tile is 64 DMMAs per warp followed by a burst of 16 stores per thread, the
// CTAs of an SM drift into the same phase, and the DMMA pipe (55 % busy at k = 4, ncu) and the store path take turns
// instead of overlapping; here both are fed all the time.  The B fragments of the whole K step stay in registers
// (NB x 4 doubles), the accumulators shrink from 64 to 16 * RP registers.
// ------------------------------------------------------------------------------------------------
template <int TM_LOG2, int TN_LOG2, int WM, int WN, int STAGES, int MINB, int RP>
__global__ void __launch_bounds__(WM * WN * 32, MINB) k_gemm_dmma_p1(KParams p) {
    constexpr int TM = 1 << TM_LOG2, TN = 1 << TN_LOG2;
    constexpr int NT = WM * WN * 32;
    constexpr int WTM = TM / WM, WTN = TN / WN;
    constexpr int MB = WTM / 8, NB = WTN / 8;
    constexpr int TK = 16, LDS = TK + 4, CHUNKS = TK / 2, RPP = NT / CHUNKS, K4 = TK / 4;
    static_assert(TM % RPP == 0 && TN % RPP == 0 && MB % RP == 0, "loader passes / row groups");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* As = reinterpret_cast<double*>(smem_raw);
    double* Bs = As + STAGES * TM * LDS;
    unsigned long long* cM = reinterpret_cast<unsigned long long*>(Bs + STAGES * TN * LDS);
    unsigned long long* cN = cM + TM;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp % WM, wn = warp / WM;
    const int k = p.k, ks = p.ksplit_log2;

    for (int i = tid; i < TM; i += NT) cM[i] = pdep_runs((unsigned long long)i, p.runs_m);
    for (int i = tid; i < TN; i += NT) cN[i] = pdep_runs((unsigned long long)i, p.runs_n);

    const int tiles_log2 = (p.m - TM_LOG2) + (p.n - TN_LOG2);
    const unsigned long long tiles = 1ull << tiles_log2;
    const unsigned long long total = tiles << ks;
    const int group_log2 = (p.m - TM_LOG2) < p.raster_group_log2 ? (p.m - TM_LOG2) : p.raster_group_log2;
    const int pg_log2 = group_log2 + (p.n - TN_LOG2);
    const unsigned long long Ksplit = (1ull << k) >> ks;   // <= TK
    const int k4_end = (int)((Ksplit + 3) / 4);
    const long long nmine = total > blockIdx.x ? (long long)((total - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    const double* Abase = operand_base(p.a, p.leaf_off, p.a_leaf);
    const double* Bbase = operand_base(p.b, p.leaf_off, p.b_leaf);
    auto decode = [&](unsigned long long id, unsigned long long& split, unsigned long long& tile_m, unsigned long long& tile_n) {
        split = id >> tiles_log2;
        const unsigned long long tid_in = id & (tiles - 1);
        const unsigned long long gidx = tid_in >> pg_log2, r = tid_in & ((1ull << pg_log2) - 1ull);
        tile_m = (gidx << group_log2) + (r & ((1ull << group_log2) - 1ull));
        tile_n = r >> group_log2;
    };

    // ---- loader: one stage = the operands of one tile; runs STAGES-1 tiles ahead ----
    const int chunk = tid % CHUNKS, row0 = tid / CHUNKS;
    const unsigned long long pass_stride = (unsigned long long)RPP << k;
    const int dst_off = row0 * LDS + chunk * 2;
    const int nbytes = ((unsigned long long)(chunk * 2) < Ksplit) ? 16 : 0;  // K = 2, 4, 8: zero-fill the tail chunks
    const int back = nbytes ? 0 : chunk * 2;                                 // keep the (unused) source address inside the row
    auto load_tile = [&](int s, unsigned long long id) {
        unsigned long long split, tile_m, tile_n;
        decode(id, split, tile_m, tile_n);
        const double* ag = Abase + ((tile_m << TM_LOG2) << k) + split * Ksplit + ((unsigned long long)row0 << k) + chunk * 2 - back;
        const double* bg = Bbase + ((tile_n << TN_LOG2) << k) + split * Ksplit + ((unsigned long long)row0 << k) + chunk * 2 - back;
        double* as = As + s * TM * LDS + dst_off;
        double* bs = Bs + s * TN * LDS + dst_off;
#pragma unroll
        for (int i = 0; i < TM / RPP; i++) cp_async16_zfill(as + i * RPP * LDS, ag + i * pass_stride, nbytes);
#pragma unroll
        for (int i = 0; i < TN / RPP; i++) cp_async16_zfill(bs + i * RPP * LDS, bg + i * pass_stride, nbytes);
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nmine) load_tile(s, blockIdx.x + (unsigned long long)s * gridDim.x);
        cp_async_commit();
    }
    const int frag_off_a = (wm * WTM + g) * LDS + t;
    const int frag_off_b = (wn * WTN + g) * LDS + t;
    const bool vec = (p.mask_n & 1ull) != 0;
    auto store_rows = [&](const double (&v)[RP][NB][2], double* out, const unsigned long long (&rb)[RP]) {
#pragma unroll
        for (int r = 0; r < RP; r++)
#pragma unroll
            for (int j = 0; j < NB; j++) {
                const int col = wn * WTN + j * 8 + 2 * t;
                if (vec) {
                    *reinterpret_cast<double2*>(out + (rb[r] | cN[col])) = make_double2(v[r][j][0], v[r][j][1]);
                } else {
                    out[rb[r] | cN[col]] = v[r][j][0];
                    out[rb[r] | cN[col + 1]] = v[r][j][1];
                }
            }
    };
    for (long long s = 0; s < nmine; s++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const long long nk = s + STAGES - 1;
        if (nk < nmine) load_tile((int)(nk % STAGES), blockIdx.x + (unsigned long long)nk * gridDim.x);
        cp_async_commit();
        const double* as = As + (int)(s % STAGES) * TM * LDS + frag_off_a;
        const double* bs = Bs + (int)(s % STAGES) * TN * LDS + frag_off_b;
        unsigned long long split, tile_m, tile_n;
        decode(blockIdx.x + (unsigned long long)s * gridDim.x, split, tile_m, tile_n);
        double* Cout = (ks > 0) ? p.ws + (split << (p.m + p.n)) : p.c;
        const unsigned long long cbase = pdep_runs(tile_m << TM_LOG2, p.runs_m) | pdep_runs(tile_n << TN_LOG2, p.runs_n);
        double bf[NB][K4];
#pragma unroll
        for (int j = 0; j < NB; j++)
#pragma unroll
            for (int k4 = 0; k4 < K4; k4++) bf[j][k4] = (k4 < k4_end) ? bs[j * 8 * LDS + k4 * 4] : 0.0;
#pragma unroll
        for (int i0 = 0; i0 < MB; i0 += RP) {
            double acc[RP][NB][2];
            double af[RP][K4];
#pragma unroll
            for (int r = 0; r < RP; r++) {
#pragma unroll
                for (int k4 = 0; k4 < K4; k4++) af[r][k4] = (k4 < k4_end) ? as[(i0 + r) * 8 * LDS + k4 * 4] : 0.0;
#pragma unroll
                for (int j = 0; j < NB; j++) acc[r][j][0] = acc[r][j][1] = 0.0;
            }
#pragma unroll
            for (int k4 = 0; k4 < K4; k4++) {
                if (k4 >= k4_end) break;  // K = 2, 4, 8: the zero-filled tail of the K step adds nothing
#pragma unroll
                for (int r = 0; r < RP; r++)
#pragma unroll
                    for (int j = 0; j < NB; j++) dmma884(acc[r][j][0], acc[r][j][1], af[r][k4], bf[j][k4]);
            }
            unsigned long long rb[RP];
#pragma unroll
            for (int r = 0; r < RP; r++) rb[r] = cbase | cM[wm * WTM + (i0 + r) * 8 + g];
            store_rows(acc, Cout, rb);
        }
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// Warp-specialised DMMA GEMM: one producer warp streams whole 128-byte operand rows into the padded
// K-major tiles with bulk async copies (cp.async.bulk, SASS UBLKCP) that signal a per-stage "full"
// mbarrier by transaction bytes; eight consumer warps wait on "full", run the DMMA steps, and arrive on
// the stage's "empty" mbarrier.  No CTA-wide barrier in the main loop (the barrier stall was the largest
// non-math stall of k_gemm_dmma, profiles/r01c_gemm_ncu_summary.json) and no per-thread address math or
// LDGSTS issue in the consumer warps.  Same tiling (128x64, two CTAs per SM), fragments and epilogue.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(a), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

// SWZ: dense 128-byte rows with the 16-byte chunk index XOR-ed by ((row & 3) << 1) instead of the padded
// 160-byte rows: the 8x4 fragment loads stay conflict-free and a stage shrinks from 30 KB to 24 KB, so FOUR
// stages fit twice per SM (ncu showed consumers waiting on `full` 24 % of the time with three).
template <int TM_LOG2, int TN_LOG2, int WM, int WN, int STAGES, int MINB, bool BULK, bool SWZ, int NPROD = 1, bool EXACT = false>
__global__ void __launch_bounds__(WM * WN * 32 + 32 * NPROD, MINB) k_gemm_dmma_ws(KParams p) {
    constexpr int TM = 1 << TM_LOG2, TN = 1 << TN_LOG2;
    constexpr int NC = WM * WN * 32;  // consumer threads
    constexpr int TK = 16, LDS = SWZ ? TK : TK + 4, K4 = TK / 4;
    static_assert(!(BULK && SWZ), "bulk row copies cannot swizzle");
    constexpr int WTM = TM / WM, WTN = TN / WN;
    constexpr int MB = WTM / 8, NB = WTN / 8;
    constexpr unsigned STAGE_BYTES = (TM + TN) * TK * 8;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* As = reinterpret_cast<double*>(smem_raw);
    double* Bs = As + STAGES * TM * LDS;
    unsigned long long* cM = reinterpret_cast<unsigned long long*>(Bs + STAGES * TN * LDS);
    unsigned long long* cN = cM + TM;
    unsigned long long* full = cN + TN;
    unsigned long long* empty = full + STAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = p.k, ks = p.ksplit_log2;

    // ---- work decode (same rasterisation as k_gemm_dmma) ----
    const unsigned long long tilesM = 1ull << (p.m - TM_LOG2), tilesN = 1ull << (p.n - TN_LOG2);
    const unsigned long long tiles = tilesM * tilesN;
    const unsigned long long id = blockIdx.x;
    const unsigned long long split = id / tiles, tid_in = id % tiles;
    const unsigned long long group = tilesM < 16 ? tilesM : 16;
    const unsigned long long per_group = group * tilesN;
    const unsigned long long gidx = tid_in / per_group, r = tid_in % per_group;
    const unsigned long long tile_m = gidx * group + (r % group), tile_n = r / group;
    const unsigned long long Ksplit = (1ull << k) >> ks;
    const int KT = (int)(Ksplit / TK);

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], BULK ? 1 : 32 * NPROD);  // bulk: the producer's expect_tx arrive + STAGE_BYTES of transactions
            mbar_init(&empty[s], WM * WN);      // one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int i = tid; i < TM; i += NC + 32 * NPROD) cM[i] = pdep_runs((unsigned long long)i, p.runs_m);
    for (int i = tid; i < TN; i += NC + 32 * NPROD) cN[i] = pdep_runs((unsigned long long)i, p.runs_n);
    __syncthreads();

    if (warp >= WM * WN) {
        const int pw = warp - WM * WN;  // producer warp index: takes every NPROD-th group of 4 rows
        // ===================== producer warp =====================
        const double* A = operand_base(p.a, p.leaf_off, p.a_leaf) + ((tile_m << TM_LOG2) << k) + split * Ksplit;
        const double* B = operand_base(p.b, p.leaf_off, p.b_leaf) + ((tile_n << TN_LOG2) << k) + split * Ksplit;
        // this lane's first source element of A / B (row = (lane >> 3) + 4 * pw, 16-byte chunk = lane & 7)
        const double* a_lane = A + ((unsigned long long)((lane >> 3) + 4 * pw) << k) + (lane & 7) * 2;
        const double* b_lane = B + ((unsigned long long)((lane >> 3) + 4 * pw) << k) + (lane & 7) * 2;
        for (int kt = 0; kt < KT; kt++) {
            const int s = kt % STAGES;
            if (kt >= STAGES) mbar_wait(&empty[s], ((kt / STAGES) - 1) & 1);  // consumers released this slot
            double* as = As + s * TM * LDS;
            double* bs = Bs + s * TN * LDS;
            if (BULK) {
                if (lane == 0) mbar_expect_tx(&full[s], STAGE_BYTES);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < TM / 32; i++) {
                    const int row = lane + 32 * i;
                    bulk_copy_g2s(as + row * LDS, A + ((unsigned long long)row << k) + kt * TK, TK * 8, &full[s]);
                }
#pragma unroll
                for (int i = 0; i < TN / 32; i++) {
                    const int row = lane + 32 * i;
                    bulk_copy_g2s(bs + row * LDS, B + ((unsigned long long)row << k) + kt * TK, TK * 8, &full[s]);
                }
            } else {
                // LDGSTS from the producer warp only; the stage's "full" barrier (count 32) completes when
                // every lane's copies have landed (cp.async.mbarrier.arrive.noinc)
                const int chunk = lane & 7, r0 = lane >> 3;  // r0 = row & 3 for every row this lane copies
                const int dchunk = SWZ ? (chunk ^ (r0 << 1)) : chunk;
                const unsigned long long step = (4ull * NPROD) << k;  // global stride between this warp's row groups
                const double* ag = a_lane + kt * TK;
                const double* bg = b_lane + kt * TK;
                double* ad = as + (r0 + 4 * pw) * LDS + dchunk * 2;
                double* bd = bs + (r0 + 4 * pw) * LDS + dchunk * 2;
#pragma unroll
                for (int i = 0; i < TM / 4 / NPROD; i++) cp_async16(ad + i * (4 * NPROD * LDS), ag + i * step);
#pragma unroll
                for (int i = 0; i < TN / 4 / NPROD; i++) cp_async16(bd + i * (4 * NPROD * LDS), bg + i * step);
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(&full[s]))
                             : "memory");
            }
        }
        return;
    }

    // ===================== consumer warps =====================
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp % WM, wn = warp / WM;
    double acc[MB][NB][2];
#pragma unroll
    for (int i = 0; i < MB; i++)
#pragma unroll
        for (int j = 0; j < NB; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int frag_off_a = (wm * WTM + g) * LDS;
    const int frag_off_b = (wn * WTN + g) * LDS;
    int koff[K4];  // column of fragment element (k4, t) inside a tile row
#pragma unroll
    for (int k4 = 0; k4 < K4; k4++)
        koff[k4] = SWZ ? ((((2 * k4 + (t >> 1)) ^ ((g & 3) << 1)) << 1) + (t & 1)) : (k4 * 4 + t);
    double af[2][MB], bf[2][NB];
    for (int kt = 0; kt < KT; kt++) {
        const int s = kt % STAGES;
        mbar_wait(&full[s], (kt / STAGES) & 1);
        const double* as = As + s * TM * LDS + frag_off_a;
        const double* bs = Bs + s * TN * LDS + frag_off_b;
#pragma unroll
        for (int i = 0; i < MB; i++) af[0][i] = as[i * 8 * LDS + koff[0]];
#pragma unroll
        for (int j = 0; j < NB; j++) bf[0][j] = bs[j * 8 * LDS + koff[0]];
#pragma unroll
        for (int k4 = 0; k4 < K4; k4++) {
            const int cur = k4 & 1, nxt = cur ^ 1;
            if (k4 + 1 < K4) {
#pragma unroll
                for (int i = 0; i < MB; i++) af[nxt][i] = as[i * 8 * LDS + koff[(k4 + 1) % K4]];
#pragma unroll
                for (int j = 0; j < NB; j++) bf[nxt][j] = bs[j * 8 * LDS + koff[(k4 + 1) % K4]];
            }
#pragma unroll
            for (int i = 0; i < MB; i++)
#pragma unroll
                for (int j = 0; j < NB; j++) dmma884(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
        }
        __syncwarp();                       // every lane's fragment loads of this stage have been consumed
        if (lane == 0) mbar_arrive(&empty[s]);
        // exact mode: at most 128 products (< 2^46 each) per accumulator between reductions
        if (EXACT && ((kt + 1) % (128 / TK) == 0 || kt + 1 == KT)) {
#pragma unroll
            for (int i = 0; i < MB; i++)
#pragma unroll
                for (int j = 0; j < NB; j++) {
                    acc[i][j][0] = mod_reduce(acc[i][j][0], p.modp, p.inv_modp);
                    acc[i][j][1] = mod_reduce(acc[i][j][1], p.modp, p.inv_modp);
                }
        }
    }

    // ---- epilogue (identical to k_gemm_dmma) ----
    double* Cout = (ks > 0) ? p.ws + (split << (p.m + p.n)) : p.c;
    const unsigned long long cbase = pdep_runs(tile_m << TM_LOG2, p.runs_m) | pdep_runs(tile_n << TN_LOG2, p.runs_n);
    const bool vec = (p.mask_n & 1ull) != 0;
#pragma unroll
    for (int i = 0; i < MB; i++) {
        const unsigned long long rbase = cbase | cM[wm * WTM + i * 8 + g];
#pragma unroll
        for (int j = 0; j < NB; j++) {
            const int col = wn * WTN + j * 8 + 2 * t;
            if (vec) {
                *reinterpret_cast<double2*>(Cout + (rbase | cN[col])) = make_double2(acc[i][j][0], acc[i][j][1]);
            } else {
                Cout[rbase | cN[col]] = acc[i][j][0];
                Cout[rbase | cN[col + 1]] = acc[i][j][1];
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Persistent + warp-specialised + row-streamed: the short-K joins (K = 16 per tile, or K = 32 as two ring stages) without a
// CTA barrier in the steady state.  k_gemm_dmma_p1 brings every warp of the CTA to a __syncthreads per tile (ncu: barrier
// stall 1.2 per issued instruction, DMMA pipe 77 % busy at K = 32); here producer warps walk the CTA's tiles
// (blockIdx.x + j * gridDim.x) and fill a 4-stage ring of swizzled K steps, each consumer warp waits for the stages of its
// tile on the `full` mbarriers, computes its 32x32 block RP row groups at a time, stores each group while the next
// computes, and hands the stages back on `empty`.  Same tiling, swizzle and fragment layout as k_gemm_dmma_ws.
// ------------------------------------------------------------------------------------------------
template <int STAGES, int NPROD, int RP, int KS>
__global__ void __launch_bounds__(256 + 32 * NPROD, 2) k_gemm_dmma_wp(KParams p) {
    constexpr int TM_LOG2 = 7, TN_LOG2 = 6, WM = 4, WN = 2;
    constexpr int TM = 1 << TM_LOG2, TN = 1 << TN_LOG2;
    constexpr int NC = WM * WN * 32;
    constexpr int TK = 16, LDS = TK, K4 = TK / 4;
    constexpr int WTM = TM / WM, WTN = TN / WN;
    constexpr int MB = WTM / 8, NB = WTN / 8;
    static_assert(MB % RP == 0 && STAGES % KS == 0, "row groups / stages per tile");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* As = reinterpret_cast<double*>(smem_raw);
    double* Bs = As + STAGES * TM * LDS;
    unsigned long long* cM = reinterpret_cast<unsigned long long*>(Bs + STAGES * TN * LDS);
    unsigned long long* cN = cM + TM;
    unsigned long long* full = cN + TN;
    unsigned long long* empty = full + STAGES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = p.k, ks = p.ksplit_log2;
    const int tiles_log2 = (p.m - TM_LOG2) + (p.n - TN_LOG2);
    const unsigned long long tiles = 1ull << tiles_log2;
    const unsigned long long total = tiles << ks;
    const int group_log2 = (p.m - TM_LOG2) < p.raster_group_log2 ? (p.m - TM_LOG2) : p.raster_group_log2;
    const int pg_log2 = group_log2 + (p.n - TN_LOG2);
    const unsigned long long Ksplit = (1ull << k) >> ks;   // == TK * KS
    const long long nmine = total > blockIdx.x ? (long long)((total - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    auto decode = [&](unsigned long long id, unsigned long long& split, unsigned long long& tile_m, unsigned long long& tile_n) {
        split = id >> tiles_log2;
        const unsigned long long tid_in = id & (tiles - 1);
        const unsigned long long gidx = tid_in >> pg_log2, r = tid_in & ((1ull << pg_log2) - 1ull);
        tile_m = (gidx << group_log2) + (r & ((1ull << group_log2) - 1ull));
        tile_n = r >> group_log2;
    };

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 32 * NPROD);
            mbar_init(&empty[s], WM * WN);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int i = tid; i < TM; i += NC + 32 * NPROD) cM[i] = pdep_runs((unsigned long long)i, p.runs_m);
    for (int i = tid; i < TN; i += NC + 32 * NPROD) cN[i] = pdep_runs((unsigned long long)i, p.runs_n);
    __syncthreads();

    if (warp >= WM * WN) {
        // ===================== producer warps: one ring stage per K step, tile after tile =====================
        const int pw = warp - WM * WN;
        const double* Abase = operand_base(p.a, p.leaf_off, p.a_leaf);
        const double* Bbase = operand_base(p.b, p.leaf_off, p.b_leaf);
        const int chunk = lane & 7, r0 = lane >> 3;
        const int dchunk = chunk ^ (r0 << 1);
        const unsigned long long step = (4ull * NPROD) << k;
        const unsigned long long lane_off = ((unsigned long long)(r0 + 4 * pw) << k) + chunk * 2;
        const int dst_off = (r0 + 4 * pw) * LDS + dchunk * 2;
        long long it = 0;
        for (long long j = 0; j < nmine; j++) {
            unsigned long long split, tile_m, tile_n;
            decode(blockIdx.x + (unsigned long long)j * gridDim.x, split, tile_m, tile_n);
            const double* a_lane = Abase + ((tile_m << TM_LOG2) << k) + split * Ksplit + lane_off;
            const double* b_lane = Bbase + ((tile_n << TN_LOG2) << k) + split * Ksplit + lane_off;
#pragma unroll
            for (int h = 0; h < KS; h++, it++) {
                const int s = (int)(it % STAGES);
                if (it >= STAGES) mbar_wait(&empty[s], (unsigned)(((it / STAGES) - 1) & 1));
                const double* ag = a_lane + h * TK;
                const double* bg = b_lane + h * TK;
                double* ad = As + s * TM * LDS + dst_off;
                double* bd = Bs + s * TN * LDS + dst_off;
#pragma unroll
                for (int i = 0; i < TM / 4 / NPROD; i++) cp_async16(ad + i * (4 * NPROD * LDS), ag + i * step);
#pragma unroll
                for (int i = 0; i < TN / 4 / NPROD; i++) cp_async16(bd + i * (4 * NPROD * LDS), bg + i * step);
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(&full[s]))
                             : "memory");
            }
        }
        return;
    }

    // ===================== consumer warps =====================
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp % WM, wn = warp / WM;
    const int frag_off_a = (wm * WTM + g) * LDS;
    const int frag_off_b = (wn * WTN + g) * LDS;
    int koff[K4];
#pragma unroll
    for (int k4 = 0; k4 < K4; k4++) koff[k4] = (((2 * k4 + (t >> 1)) ^ ((g & 3) << 1)) << 1) + (t & 1);
    const bool vec = (p.mask_n & 1ull) != 0;
    long long it = 0;
    for (long long j = 0; j < nmine; j++, it += KS) {
        unsigned long long split, tile_m, tile_n;
        decode(blockIdx.x + (unsigned long long)j * gridDim.x, split, tile_m, tile_n);
        double* Cout = (ks > 0) ? p.ws + (split << (p.m + p.n)) : p.c;
        const unsigned long long cbase = pdep_runs(tile_m << TM_LOG2, p.runs_m) | pdep_runs(tile_n << TN_LOG2, p.runs_n);
#pragma unroll
        for (int h = 0; h < KS; h++) mbar_wait(&full[(int)((it + h) % STAGES)], (unsigned)(((it + h) / STAGES) & 1));
        const int s0 = (int)(it % STAGES);  // KS divides STAGES: the tile's stages are s0 .. s0 + KS - 1
        double bf[NB][K4];
        if (KS == 1) {
#pragma unroll
            for (int jj = 0; jj < NB; jj++)
#pragma unroll
                for (int k4 = 0; k4 < K4; k4++) bf[jj][k4] = Bs[s0 * TN * LDS + frag_off_b + jj * 8 * LDS + koff[k4]];
        }
#pragma unroll
        for (int i0 = 0; i0 < MB; i0 += RP) {
            double acc[RP][NB][2];
#pragma unroll
            for (int r = 0; r < RP; r++)
#pragma unroll
                for (int jj = 0; jj < NB; jj++) acc[r][jj][0] = acc[r][jj][1] = 0.0;
#pragma unroll
            for (int h = 0; h < KS; h++) {
                const double* as = As + (s0 + h) * TM * LDS + frag_off_a;
                double af[RP][K4];
#pragma unroll
                for (int r = 0; r < RP; r++)
#pragma unroll
                    for (int k4 = 0; k4 < K4; k4++) af[r][k4] = as[(i0 + r) * 8 * LDS + koff[k4]];
                if (KS > 1) {
#pragma unroll
                    for (int jj = 0; jj < NB; jj++)
#pragma unroll
                        for (int k4 = 0; k4 < K4; k4++) bf[jj][k4] = Bs[(s0 + h) * TN * LDS + frag_off_b + jj * 8 * LDS + koff[k4]];
                }
#pragma unroll
                for (int k4 = 0; k4 < K4; k4++)
#pragma unroll
                    for (int r = 0; r < RP; r++)
#pragma unroll
                        for (int jj = 0; jj < NB; jj++) dmma884(acc[r][jj][0], acc[r][jj][1], af[r][k4], bf[jj][k4]);
            }
            if (i0 + RP >= MB) {
                // the last row group's fragments are in registers (its DMMAs were issued behind the loads): the stages can be refilled
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int h = 0; h < KS; h++) mbar_arrive(&empty[s0 + h]);
                }
            }
#pragma unroll
            for (int r = 0; r < RP; r++) {
                const unsigned long long rbase = cbase | cM[wm * WTM + (i0 + r) * 8 + g];
#pragma unroll
                for (int jj = 0; jj < NB; jj++) {
                    const int col = wn * WTN + jj * 8 + 2 * t;
                    if (vec) {
                        *reinterpret_cast<double2*>(Cout + (rbase | cN[col])) = make_double2(acc[r][jj][0], acc[r][jj][1]);
                    } else {
                        Cout[rbase | cN[col]] = acc[r][jj][0];
                        Cout[rbase | cN[col + 1]] = acc[r][jj][1];
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Stream-K form of the warp-specialised kernel, for joins whose tile count quantises badly on the CTA slots
// (e.g. m=11,n=10: 256 tiles of 128x64 on 296 slots — 40 SMs run one CTA while 108 run two, 0.79 of peak).
// The K steps of ALL tiles form one sequence  g = tile * KT + kt  that is cut into gridDim.x equal ranges, one per
// CTA (grid = the CTA slots), so every SM carries the same number of K steps.  A range is a list of segments
// (tile, kt0, kt1): only its first segment can start inside a tile and only its last can stop short of a tile's end.
//  * a segment that reaches its tile's end makes the CTA that tile's OWNER: it adds the partial tiles of the CTAs that
//    computed the earlier K steps (ascending CTA order: deterministic) and scatters the result into C;
//  * a segment that stops short ("open") is written as a partial tile, in fragment order (coalesced 16-byte stores),
//    into the CTA's 64 KB slot of the workspace and published with a release flag.  A CTA runs its open segment FIRST,
//    so partials exist long before their owners need them, and the only wait of the kernel is on a CTA that itself
//    waits for nobody at that point.
// CTA ranks are handed out by an atomic counter in start order (not blockIdx), so a waited-for CTA is always resident:
// no co-residency assumption, safe next to other kernels.  Flags are reset by their one consumer, the counter wraps
// to zero with the last CTA: the 4 KB flag block only has to be zero once (plan upload).
// The segment list lives in shared memory: the main loop carries no more live state than k_gemm_dmma_ws (the
// persistent-tile variant measured in round 1 lost 5 % to spills at the 96-register budget).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* ptr) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* ptr, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(ptr), "r"(v) : "memory");
}


template <int TM_LOG2, int TN_LOG2, int WM, int WN, int STAGES, int MINB, int NPROD>
__global__ void __launch_bounds__(WM * WN * 32 + 32 * NPROD, MINB) k_gemm_dmma_sk(KParams p) {
    constexpr int TM = 1 << TM_LOG2, TN = 1 << TN_LOG2;
    constexpr int NC = WM * WN * 32;  // consumer threads
    constexpr int TK = 16, LDS = TK, K4 = TK / 4;
    constexpr int WTM = TM / WM, WTN = TN / WN;
    constexpr int MB = WTM / 8, NB = WTN / 8;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* As = reinterpret_cast<double*>(smem_raw);
    double* Bs = As + STAGES * TM * LDS;
    unsigned long long* cM = reinterpret_cast<unsigned long long*>(Bs + STAGES * TN * LDS);
    unsigned long long* cN = cM + TM;
    unsigned long long* full = cN + TN;
    unsigned long long* empty = full + STAGES;
    // the segment list sits BEHIND the ring in the dynamic block: static __shared__ variables would shift the ring off its
    // 128-byte alignment, and the swizzled rows then straddle bank lines (measured: 4x the LDGSTS shared-memory wavefronts,
    // 2x the L2 sectors, the main loop 15-40 % slower — profiles/r02h_streamk_misaligned_ncu.md)
    int* seg_tile = reinterpret_cast<int*>(empty + STAGES);
    int* seg_kt0 = seg_tile + kSkMaxSegs;
    int* seg_kt1 = seg_kt0 + kSkMaxSegs;
    int& s_nseg = seg_kt1[kSkMaxSegs];
    int& s_rank = seg_kt1[kSkMaxSegs + 1];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = p.k;
    const int KT = 1 << (k - 4);
    const int tiles_log2 = (p.m - TM_LOG2) + (p.n - TN_LOG2);
    const int group_log2 = (p.m - TM_LOG2) < 4 ? (p.m - TM_LOG2) : 4;
    const int pg_log2 = group_log2 + (p.n - TN_LOG2);
    unsigned* flags = p.sk_flags;  // [0]: start-order counter, [1 + rank]: "partial of CTA rank is ready"

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full[s], 32 * NPROD);
            mbar_init(&empty[s], WM * WN);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        // ---- this CTA's rank (start order) and its range of the global K-step sequence ----
        const unsigned P = gridDim.x;
        const unsigned rank = atomicInc(flags, P - 1);  // wraps to 0 after the last CTA of the launch
        const unsigned long long G = (unsigned long long)KT << tiles_log2;
        const unsigned long long q = G / P, r = G % P;
        const unsigned long long g0 = rank * q + (rank < r ? rank : r), g1 = g0 + q + (rank < r ? 1 : 0);
        const int t_first = (int)(g0 >> (k - 4)), t_last = (int)((g1 - 1) >> (k - 4));
        const int nseg = t_last - t_first + 1;
        const int end1 = (int)(g1 & (unsigned long long)(KT - 1));  // != 0: the last segment is open
        const bool rotate = end1 != 0 && nseg > 1;                  // run the open segment first
        for (int x = 0; x < nseg; x++) {
            const int s = rotate ? (x == 0 ? nseg - 1 : x - 1) : x;
            seg_tile[x] = t_first + s;
            seg_kt0[x] = s == 0 ? (int)(g0 & (unsigned long long)(KT - 1)) : 0;
            seg_kt1[x] = (s == nseg - 1 && end1 != 0) ? end1 : KT;
        }
        s_nseg = nseg;
        s_rank = (int)rank;
    }
    for (int i = tid; i < TM; i += NC + 32 * NPROD) cM[i] = pdep_runs((unsigned long long)i, p.runs_m);
    for (int i = tid; i < TN; i += NC + 32 * NPROD) cN[i] = pdep_runs((unsigned long long)i, p.runs_n);
    __syncthreads();
    const int nseg = s_nseg;

    if (warp >= WM * WN) {
        // ===================== producer warps =====================
        const int pw = warp - WM * WN;
        const double* Abase = operand_base(p.a, p.leaf_off, p.a_leaf);
        const double* Bbase = operand_base(p.b, p.leaf_off, p.b_leaf);
        const int chunk = lane & 7, r0 = lane >> 3;
        const int dchunk = chunk ^ (r0 << 1);
        const unsigned long long step = (4ull * NPROD) << k;
        const unsigned long long lane_off = ((unsigned long long)(r0 + 4 * pw) << k) + chunk * 2;
        const int dst_off = (r0 + 4 * pw) * LDS + dchunk * 2;
        int it = 0;
        for (int x = 0; x < nseg; x++) {
            const unsigned tile = (unsigned)seg_tile[x];
            const unsigned gidx = tile >> pg_log2, rr = tile & ((1u << pg_log2) - 1u);
            const unsigned long long tile_m = ((unsigned long long)gidx << group_log2) + (rr & ((1u << group_log2) - 1u));
            const unsigned long long tile_n = rr >> group_log2;
            const double* a_lane = Abase + ((tile_m << TM_LOG2) << k) + lane_off;
            const double* b_lane = Bbase + ((tile_n << TN_LOG2) << k) + lane_off;
            const int kt1 = seg_kt1[x];
            for (int kt = seg_kt0[x]; kt < kt1; kt++, it++) {
                const int s = it % STAGES;
                if (it >= STAGES) mbar_wait(&empty[s], ((it / STAGES) - 1) & 1);
                const double* ag = a_lane + kt * TK;
                const double* bg = b_lane + kt * TK;
                double* ad = As + s * TM * LDS + dst_off;
                double* bd = Bs + s * TN * LDS + dst_off;
#pragma unroll
                for (int i = 0; i < TM / 4 / NPROD; i++) cp_async16(ad + i * (4 * NPROD * LDS), ag + i * step);
#pragma unroll
                for (int i = 0; i < TN / 4 / NPROD; i++) cp_async16(bd + i * (4 * NPROD * LDS), bg + i * step);
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(&full[s]))
                             : "memory");
            }
        }
        return;
    }

    // ===================== consumer warps =====================
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp % WM, wn = warp / WM;
    double acc[MB][NB][2];
    const int frag_off_a = (wm * WTM + g) * LDS;
    const int frag_off_b = (wn * WTN + g) * LDS;
    int koff[K4];
#pragma unroll
    for (int k4 = 0; k4 < K4; k4++) koff[k4] = (((2 * k4 + (t >> 1)) ^ ((g & 3) << 1)) << 1) + (t & 1);
    double af[2][MB], bf[2][NB];
    int it = 0;
    for (int x = 0; x < nseg; x++) {
#pragma unroll
        for (int i = 0; i < MB; i++)
#pragma unroll
            for (int j = 0; j < NB; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
        const int nsteps = seg_kt1[x] - seg_kt0[x];
        for (int e = it + nsteps; it < e; it++) {
            const int s = it % STAGES;
            mbar_wait(&full[s], (it / STAGES) & 1);
            const double* as = As + s * TM * LDS + frag_off_a;
            const double* bs = Bs + s * TN * LDS + frag_off_b;
#pragma unroll
            for (int i = 0; i < MB; i++) af[0][i] = as[i * 8 * LDS + koff[0]];
#pragma unroll
            for (int j = 0; j < NB; j++) bf[0][j] = bs[j * 8 * LDS + koff[0]];
#pragma unroll
            for (int k4 = 0; k4 < K4; k4++) {
                const int cur = k4 & 1, nxt = cur ^ 1;
                if (k4 + 1 < K4) {
#pragma unroll
                    for (int i = 0; i < MB; i++) af[nxt][i] = as[i * 8 * LDS + koff[(k4 + 1) % K4]];
#pragma unroll
                    for (int j = 0; j < NB; j++) bf[nxt][j] = bs[j * 8 * LDS + koff[(k4 + 1) % K4]];
                }
#pragma unroll
                for (int i = 0; i < MB; i++)
#pragma unroll
                    for (int j = 0; j < NB; j++) dmma884(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        // ---- the segment is done: a partial tile for its owner, or this CTA owns the tile ----
        const int rank = s_rank;
        double2* slots = reinterpret_cast<double2*>(p.ws);  // slot of CTA `rank`: [MB * NB][NC] double2, fragment order
        if (seg_kt1[x] < KT) {
            double2* mine = slots + (size_t)rank * (MB * NB * NC) + tid;
#pragma unroll
            for (int i = 0; i < MB; i++)
#pragma unroll
                for (int j = 0; j < NB; j++) mine[(i * NB + j) * NC] = make_double2(acc[i][j][0], acc[i][j][1]);
            __threadfence();
            named_bar_sync(1, NC);
            if (tid == 0) st_release_u32(flags + 1 + rank, 1u);
            continue;
        }
        const unsigned tile = (unsigned)seg_tile[x];
        if (seg_kt0[x] > 0) {
            // contributors: the CTAs whose ranges cover K steps [tile * KT, tile * KT + kt0) — ranks first .. rank - 1
            const unsigned P = gridDim.x;
            const unsigned long long G = (unsigned long long)KT << tiles_log2;
            const unsigned long long q = G / P, r = G % P;
            const unsigned long long s0 = (unsigned long long)tile << (k - 4);
            const int first = (int)(s0 < r * (q + 1) ? s0 / (q + 1) : r + (s0 - r * (q + 1)) / q);
            for (int j = first; j < rank; j++) {
                if (lane == 0) {
                    // the contributor is resident and publishes before it waits for anyone, so this ends within its open
                    // segment's run time (microseconds); seconds of spinning mean a broken invariant: fail the launch
                    // (the host sees a CUDA error) instead of hanging the device
                    unsigned spins = 0;
                    while (ld_acquire_u32(flags + 1 + j) == 0u) {
                        __nanosleep(128);
                        if (++spins > (1u << 25)) __trap();
                    }
                }
                __syncwarp();
                const double2* theirs = slots + (size_t)j * (MB * NB * NC) + tid;
#pragma unroll
                for (int i = 0; i < MB; i++)
#pragma unroll
                    for (int jj = 0; jj < NB; jj++) {
                        const double2 v = __ldcg(theirs + (i * NB + jj) * NC);
                        acc[i][jj][0] += v.x;
                        acc[i][jj][1] += v.y;
                    }
                named_bar_sync(1, NC);  // every consumer warp has seen the flag and read the partial
                if (tid == 0) flags[1 + j] = 0u;
            }
        }
        const unsigned gidx = tile >> pg_log2, rr = tile & ((1u << pg_log2) - 1u);
        const unsigned long long tile_m = ((unsigned long long)gidx << group_log2) + (rr & ((1u << group_log2) - 1u));
        const unsigned long long tile_n = rr >> group_log2;
        const unsigned long long cbase = pdep_runs(tile_m << TM_LOG2, p.runs_m) | pdep_runs(tile_n << TN_LOG2, p.runs_n);
        const bool vec = (p.mask_n & 1ull) != 0;
#pragma unroll
        for (int i = 0; i < MB; i++) {
            const unsigned long long rbase = cbase | cM[wm * WTM + i * 8 + g];
#pragma unroll
            for (int j = 0; j < NB; j++) {
                const int col = wn * WTN + j * 8 + 2 * t;
                if (vec) {
                    *reinterpret_cast<double2*>(p.c + (rbase | cN[col])) = make_double2(acc[i][j][0], acc[i][j][1]);
                } else {
                    p.c[rbase | cN[col]] = acc[i][j][0];
                    p.c[rbase | cN[col + 1]] = acc[i][j][1];
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// A TMA-fed variant of this kernel (2-D tensor maps, cp.async.bulk.tensor.2d / UTMALDG.2D from one elected thread, 128-byte
// swizzle, one producer warp) was built and measured in round 2 and is NOT part of the build: it ran 0.4-2.3 % slower than
// the two-producer LDGSTS feed on every join measured (profiles/r02b_kernel_lab_tma_staged.md; ncu side by side:
// profiles/r02c_gemm_11_11_12{,_tma}_ncu_summary.json — same DRAM traffic, 2.5x the shared-memory bank conflicts), and it
// returned slightly wrong counts when two slice lanes ran it concurrently (profiles/r02d_tma_race_*.log; never with the
// LDGSTS feed).  The code is in the history: commit "TMA-fed DMMA GEMM variant" (k_gemm_dmma_tma).
// ------------------------------------------------------------------------------------------------
template <int TM_LOG2, int TN_LOG2, int STAGES, bool SWZ = false>
constexpr size_t gemm_ws_smem_bytes() {
    return (size_t)STAGES * ((1 << TM_LOG2) + (1 << TN_LOG2)) * (SWZ ? 16 : 20) * 8 + ((1 << TM_LOG2) + (1 << TN_LOG2)) * 8 + 2 * STAGES * 8 + 16;
}

constexpr size_t gemm_sk_smem_bytes() { return gemm_ws_smem_bytes<7, 6, 4, true>() + (3 * kSkMaxSegs + 4) * sizeof(int); }

template <int TM_LOG2, int TN_LOG2, int TK, int STAGES>
constexpr size_t gemm_smem_bytes() {
    return (size_t)STAGES * ((1 << TM_LOG2) + (1 << TN_LOG2)) * (TK + 4) * 8 + ((1 << TM_LOG2) + (1 << TN_LOG2)) * 8;
}

// The shipped instantiations (the 128x128, one-producer, bulk-copy and four-producer variants measured in round 1
// are documented in DESIGN.md §4 and were removed from the build): 128x64 tiles at two CTAs per SM.
#define GEMM_76 k_gemm_dmma<7, 6, 4, 2, 16, 3, 2>
#define GEMM_66 k_gemm_dmma<6, 6, 2, 4, 16, 4, 1>
#define GEMM_76_P k_gemm_dmma_p<7, 6, 4, 2, 16, 3, 2>
// K = 16: one K step per tile, two row groups of 8 stored while the next two compute.  Measured next to it and not kept
// (profiles/r02h_kernel_lab_store_*.md): one row group at a time (0.83 of HBM at k = 4: too few independent DMMA chains),
// three CTAs per SM at 80 registers (0.80), stores deferred behind the next group's DMMAs (0.83), 64x64 tiles at four
// CTAs per SM on k_gemm_dmma_p (0.83, and 0.73 instead of 0.72 of the FP64 peak at k = 5).
#define GEMM_76_P1 k_gemm_dmma_p1<7, 6, 4, 2, 3, 2, 2>
// K = 32 as two ring stages per tile, two producer warps.  Measured next to it (profiles/r02i_kernel_lab_k32.md): one producer
// warp (0.81 instead of 0.85 at m=15,n=14,k=5), k_gemm_dmma_p1 with both K steps in one stage and a CTA barrier per tile (0.79),
// and the same kernel for K = 16 (0.80-0.82 where k_gemm_dmma_p1 reaches 0.89: kept there).
#define GEMM_76_WP2 k_gemm_dmma_wp<4, 2, 2, 2>
#define GEMM_76_WL k_gemm_dmma_ws<7, 6, 4, 2, 3, 2, false, false>
#define GEMM_76_WZ2 k_gemm_dmma_ws<7, 6, 4, 2, 4, 2, false, true, 2>
#define GEMM_76_SK k_gemm_dmma_sk<7, 6, 4, 2, 4, 2, 2>
// residue-arithmetic instantiations (entry type bigint)
#define GEMM_76_X k_gemm_dmma<7, 6, 4, 2, 16, 3, 2, true>
#define GEMM_66_X k_gemm_dmma<6, 6, 2, 4, 16, 4, 1, true>
#define GEMM_76_WZ2_X k_gemm_dmma_ws<7, 6, 4, 2, 4, 2, false, true, 2, true>

template <int NT, int FWD>
__global__ void k_microtree(const MicroOpDev* __restrict__ ops, const int32_t* __restrict__ cta_start, const double* leaves,
                            double* arena, const double* arena0, const long long* leaf_off, int smem_ops, double modp);

cudaError_t configure_kernels() {
    cudaError_t e;
    e = cudaFuncSetAttribute(GEMM_76, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<7, 6, 16, 3>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(GEMM_66, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<6, 6, 16, 4>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(GEMM_76_P, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<7, 6, 16, 3>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(GEMM_76_P1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<7, 6, 16, 3>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(GEMM_76_WP2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_ws_smem_bytes<7, 6, 4, true>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(GEMM_76_WL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_ws_smem_bytes<7, 6, 3>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(GEMM_76_WZ2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_ws_smem_bytes<7, 6, 4, true>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(GEMM_76_SK, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_sk_smem_bytes());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(GEMM_76_X, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<7, 6, 16, 3>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(GEMM_66_X, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<6, 6, 16, 4>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(GEMM_76_WZ2_X, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_ws_smem_bytes<7, 6, 4, true>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_microtree<256, kMicroFwdMax>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kMicroDescBytes + (int)((kMicroLeafCache + 2 * kMicroFwdMax) * sizeof(double)));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_microtree<1024, kMicroFwdMaxBig>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kMicroDescBytes + (int)((kMicroLeafCache + 2 * kMicroFwdMaxBig) * sizeof(double)));
    return e;
}

// ------------------------------------------------------------------------------------------------
// slice bookkeeping
// ------------------------------------------------------------------------------------------------
__global__ void k_begin_slice(DevState* st, SliceTables t) {
    __shared__ unsigned long long sid;
    if (threadIdx.x == 0) sid = st->next_slice;
    __syncthreads();
    for (int l = threadIdx.x; l < t.n_leaves; l += blockDim.x) {
        long long off = 0;
        for (int j = t.term_start[l]; j < t.term_start[l + 1]; j++)
            off |= (long long)((sid >> t.id_bit[j]) & 1ull) << t.addr_bit[j];
        t.leaf_off[l] = off;
    }
    __syncthreads();
    if (threadIdx.x == 0) st->next_slice = sid + st->stride;
}

// the rank-0 result of the lane's current slice goes to its slot of the result buffer; the slices of a
// run may execute on several lanes concurrently, the SUM is taken afterwards in slice order
__global__ void k_accum(DevState* st, const double* root, const long long* leaf_off, int root_leaf, double* results) {
    const double* r = operand_base(root, leaf_off, root_leaf);
    const unsigned long long slot = st->slot;
    results[slot] = r[0];
    st->slot = slot + st->slot_stride;
}

// acc = (previous acc | initial) + results[0] + results[1] + ...   sequentially, exactly like the
// reference's `result += tensor_result[()]` loop (base_api.py:25-27)
__global__ void k_final_sum(double* acc, const double* results, int count, double initial, int use_previous, double modp) {
    double s = use_previous ? acc[0] : initial;
    for (int i = 0; i < count; i++) s += results[i];  // exact mode: <= 4096 residues + one residue: exact
    acc[0] = (modp > 0.0) ? mod_reduce(s, modp, 1.0 / modp) : s;
}

cudaError_t launch_begin_slice(DevState* st, SliceTables t, cudaStream_t stream) {
    k_begin_slice<<<1, 256, 0, stream>>>(st, t);
    return cudaGetLastError();
}

cudaError_t launch_accum(DevState* st, const double* root, const long long* leaf_off, int root_leaf, double* results,
                         cudaStream_t stream) {
    k_accum<<<1, 1, 0, stream>>>(st, root, leaf_off, root_leaf, results);
    return cudaGetLastError();
}

cudaError_t launch_final_sum(double* acc, const double* results, int count, double initial, int use_previous, double modp,
                             cudaStream_t stream) {
    k_final_sum<<<1, 1, 0, stream>>>(acc, results, count, initial, use_previous, modp);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// micro subtrees: CTA b runs joins [cta_start[b], cta_start[b+1]) one after the other; every join is
// tiny (<= 2^12 outputs, <= 2^15 multiply-adds), its operands were produced by this CTA or are leaves,
// so a __syncthreads between joins is the only synchronisation.  Replaces hundreds of launch-bound
// kernel launches per slice by one.
// ------------------------------------------------------------------------------------------------
// One join of a micro stage, for the thread's NU outputs  c0, c0 + NT, ...  (NU = 1 when the join has at most
// NT outputs, so small joins carry no dead work).  AG / BG: the operand comes from global memory (L2, .cg
// loads) rather than shared memory.  The loads of all NU outputs are issued before the first FMA.
template <int NT, int LOG_NT, int NU, bool AG, bool BG>
__device__ __forceinline__ void micro_join(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
                                           double* __restrict__ mine, bool fwd_out, unsigned mask, int tot, int k,
                                           unsigned mi_lo, unsigned ni_lo, const uint2* __restrict__ hi_tab, double modp,
                                           double inv) {
    const unsigned outs = 1u << tot;
    for (unsigned c0 = threadIdx.x; c0 < outs; c0 += NT * NU) {
        const double* ar[NU];
        const double* br[NU];
#pragma unroll
        for (int u = 0; u < NU; u++) {
            const unsigned c = c0 + u * NT;  // < outs: outs is a multiple of NT * NU whenever NU > 1
            // the bits above log2(NT) were decoded once per join into hi_tab (<= 16 entries)
            const uint2 hi = NU > 1 ? hi_tab[c >> LOG_NT] : make_uint2(0u, 0u);
            ar[u] = A + ((size_t)(mi_lo | hi.x) << k);
            br[u] = B + ((size_t)(ni_lo | hi.y) << k);
        }
        double s[NU];
        if (k == 0) {
            double x[NU], y[NU];
#pragma unroll
            for (int u = 0; u < NU; u++) x[u] = AG ? __ldcg(ar[u]) : ar[u][0];
#pragma unroll
            for (int u = 0; u < NU; u++) y[u] = BG ? __ldcg(br[u]) : br[u][0];
#pragma unroll
            for (int u = 0; u < NU; u++) s[u] = x[u] * y[u];
        } else {
            const int K2 = 1 << (k - 1);  // k <= 4 for mini joins: at most 8 double2 steps
            double s0[NU], s1[NU];
#pragma unroll
            for (int u = 0; u < NU; u++) s0[u] = s1[u] = 0.0;
            for (int j = 0; j < K2; j++) {
                double2 x[NU], y[NU];
#pragma unroll
                for (int u = 0; u < NU; u++)
                    x[u] = AG ? __ldcg(reinterpret_cast<const double2*>(ar[u]) + j) : reinterpret_cast<const double2*>(ar[u])[j];
#pragma unroll
                for (int u = 0; u < NU; u++)
                    y[u] = BG ? __ldcg(reinterpret_cast<const double2*>(br[u]) + j) : reinterpret_cast<const double2*>(br[u])[j];
#pragma unroll
                for (int u = 0; u < NU; u++) {
                    s0[u] = fma(x[u].x, y[u].x, s0[u]);
                    s1[u] = fma(x[u].y, y[u].y, s1[u]);
                }
                if (modp > 0.0 && (j & 63) == 63) {  // exact mode: reduce before 65 products pile up
#pragma unroll
                    for (int u = 0; u < NU; u++) {
                        s0[u] = mod_reduce(s0[u], modp, inv);
                        s1[u] = mod_reduce(s1[u], modp, inv);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < NU; u++) s[u] = s0[u] + s1[u];
        }
#pragma unroll
        for (int u = 0; u < NU; u++) {
            const unsigned c = c0 + u * NT;
            const double v = modp > 0.0 ? mod_reduce(s[u], modp, inv) : s[u];
            C[c] = v;
            if (fwd_out) mine[c] = v;
        }
    }
}

// Shared memory: [join descriptors | leaf cache | two forward buffers].  The serial chain of joins a CTA
// walks is latency-bound: with the descriptors and the (tiny) leaf operands staged up front and each
// result handed to the next join through the forward buffer, the critical path of a join is shared-memory
// latency instead of two L2 round trips (store result, load it back).  The address bits of an output that
// come from the thread index are decoded once per join, only the few bits above log2(NT) per output.
template <int NT, int FWD>
__global__ void __launch_bounds__(NT) k_microtree(const MicroOpDev* __restrict__ ops, const int32_t* __restrict__ cta_start,
                                                  const double* leaves, double* arena, const double* arena0,
                                                  const long long* leaf_off, int smem_ops, double modp) {
    constexpr int LOG_NT = NT == 1024 ? 10 : 8;
    static_assert(NT == (1 << LOG_NT), "NT");
    __shared__ uint2 hi_tab[64];  // 2^(14 - 8) entries at most
    extern __shared__ __align__(16) unsigned char micro_smem[];
    MicroOpDev* sops = reinterpret_cast<MicroOpDev*>(micro_smem);
    double* cache = reinterpret_cast<double*>(micro_smem + (size_t)smem_ops * sizeof(MicroOpDev));
    double* fwd = cache + kMicroLeafCache;  // [2][FWD]
    const int first = cta_start[blockIdx.x], last = cta_start[blockIdx.x + 1];
    const bool staged = (last - first) <= smem_ops;
    const double inv = modp > 0.0 ? 1.0 / modp : 0.0;
    if (staged) {
        const int4* src = reinterpret_cast<const int4*>(ops + first);
        int4* dst = reinterpret_cast<int4*>(sops);
        const int n16 = (last - first) * (int)(sizeof(MicroOpDev) / 16);
        for (int i = threadIdx.x; i < n16; i += NT) dst[i] = src[i];
        __syncthreads();
        // leaf operands: one thread per (join, operand), a handful of doubles each
        for (int i = threadIdx.x; i < 2 * (last - first); i += NT) {
            const MicroOpDev& op = sops[i >> 1];
            const bool is_b = i & 1;
            if ((is_b ? op.b_src : op.a_src) != 2) continue;
            const int leaf = is_b ? op.b_leaf : op.a_leaf;
            const double* g = leaves + (is_b ? op.b_off : op.a_off) + (leaf >= 0 ? leaf_off[leaf] : 0);
            double* d = cache + (is_b ? op.b_soff : op.a_soff);
            const int n = 1 << ((is_b ? op.n : op.m) + op.k);
            for (int e = 0; e < n; e++) d[e] = __ldcg(g + e);
        }
        __syncthreads();
    }
    for (int i = first; i < last; i++) {
        const MicroOpDev op = staged ? sops[i - first] : ops[i];
        const int a_src = staged ? op.a_src : 0, b_src = staged ? op.b_src : 0;
        const double* prev = fwd + ((i - first + 1) & 1) * FWD;  // written by join i-1
        double* mine = fwd + ((i - first) & 1) * FWD;
        const int tot = op.m + op.n, k = op.k;
        if (tot > LOG_NT) {  // decode the (at most 4) output-index bits above log2(NT) once: hi_tab[c >> LOG_NT]
            if (threadIdx.x < (1u << (tot - LOG_NT))) {
                const unsigned maskv = op.mask_m;
                int im = __popc(maskv & ((1u << LOG_NT) - 1u)), in = LOG_NT - im;  // bits the low part already used
                unsigned mi = 0, ni = 0;
                for (int b = LOG_NT; b < tot; b++) {
                    const unsigned bit = (threadIdx.x >> (b - LOG_NT)) & 1u;
                    if ((maskv >> b) & 1u) { mi |= bit << im; im++; }
                    else { ni |= bit << in; in++; }
                }
                hi_tab[threadIdx.x] = make_uint2(mi, ni);
            }
            __syncthreads();
        }
        if (threadIdx.x < (1u << tot)) {  // warps without an output go straight to the barrier
            const double* A = a_src == 1 ? prev : a_src == 2 ? cache + op.a_soff
                              : (op.a_space == 0 ? leaves : (op.a_space == 2 ? arena0 : arena)) + op.a_off +
                                    (op.a_leaf >= 0 ? leaf_off[op.a_leaf] : 0);
            const double* B = b_src == 1 ? prev : b_src == 2 ? cache + op.b_soff
                              : (op.b_space == 0 ? leaves : (op.b_space == 2 ? arena0 : arena)) + op.b_off +
                                    (op.b_leaf >= 0 ? leaf_off[op.b_leaf] : 0);
            double* C = arena + op.c_off;
            const unsigned mask = op.mask_m;
            const bool fwd_out = staged && op.fwd_out;
            // address bits supplied by the thread index: decoded once per join
            unsigned mi_lo = 0, ni_lo = 0;
            int im_lo = 0, in_lo = 0;
            const int lo_bits = tot < LOG_NT ? tot : LOG_NT;
            for (int b = 0; b < lo_bits; b++) {
                const unsigned bit = (threadIdx.x >> b) & 1u;
                if ((mask >> b) & 1u) { mi_lo |= bit << im_lo; im_lo++; }
                else { ni_lo |= bit << in_lo; in_lo++; }
            }
            const int nu = tot <= LOG_NT ? 1 : (tot == LOG_NT + 1 ? 2 : 4);
            const int sel = (a_src == 0 ? 2 : 0) | (b_src == 0 ? 1 : 0);
#define TOB_MICRO_CALL(NU, AG, BG) \
    micro_join<NT, LOG_NT, NU, AG, BG>(A, B, C, mine, fwd_out, mask, tot, k, mi_lo, ni_lo, hi_tab, modp, inv)
#define TOB_MICRO_SEL(NU)                                    \
    switch (sel) {                                           \
        case 0: TOB_MICRO_CALL(NU, false, false); break;     \
        case 1: TOB_MICRO_CALL(NU, false, true); break;      \
        case 2: TOB_MICRO_CALL(NU, true, false); break;      \
        default: TOB_MICRO_CALL(NU, true, true); break;      \
    }
            if (nu == 1) { TOB_MICRO_SEL(1) }
            else if (nu == 2) { TOB_MICRO_SEL(2) }
            else { TOB_MICRO_SEL(4) }
#undef TOB_MICRO_SEL
#undef TOB_MICRO_CALL
        }
        __syncthreads();
    }
}

cudaError_t launch_microtree(const MicroOpDev* ops, const int32_t* cta_start, int n_ctas, int smem_ops, int threads, const double* leaves,
                             double* arena, const double* arena0, const long long* leaf_off, double modp,
                             cudaStream_t stream) {
    if (threads > 256) {  // stages with results above 2^12 doubles: four times the threads and forward buffers
        const size_t smem = (size_t)smem_ops * sizeof(MicroOpDev) + (size_t)(kMicroLeafCache + 2 * kMicroFwdMaxBig) * sizeof(double);
        k_microtree<1024, kMicroFwdMaxBig><<<n_ctas, 1024, smem, stream>>>(ops, cta_start, leaves, arena, arena0, leaf_off, smem_ops, modp);
    } else {
        const size_t smem = (size_t)smem_ops * sizeof(MicroOpDev) + (size_t)(kMicroLeafCache + 2 * kMicroFwdMax) * sizeof(double);
        k_microtree<256, kMicroFwdMax><<<n_ctas, 256, smem, stream>>>(ops, cta_start, leaves, arena, arena0, leaf_off, smem_ops, modp);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// contract launcher
// ------------------------------------------------------------------------------------------------
static unsigned grid_for(unsigned long long items, unsigned long long per_block, unsigned long long cap) {
    unsigned long long b = (items + per_block - 1) / per_block;
    if (b < 1) b = 1;
    if (b > cap) b = cap;
    return (unsigned)b;
}

cudaError_t launch_contract(const Op& op, const KParams& p, cudaStream_t stream, int* launches) {
    const unsigned long long outs = 1ull << (op.m + op.n);
    const unsigned long long cap = 1ull << 30;
    if (op.kind == OP_GEMM) {
        const unsigned long long tiles = 1ull << ((op.m - op.tm_log2) + (op.n - op.tn_log2));
        const unsigned long long blocks = tiles << op.ksplit_log2;
        if (blocks > 0x7fffffffull) return cudaErrorInvalidConfiguration;
        if (p.modp > 0.0) {
            // exact mode: residue-arithmetic instantiations of the default kernels
            if (op.tm_log2 == 7 && op.tn_log2 == 6 && (op.k - op.ksplit_log2) >= 8)
                GEMM_76_WZ2_X<<<(unsigned)blocks, 320, gemm_ws_smem_bytes<7, 6, 4, true>(), stream>>>(p);
            else if (op.tm_log2 == 7 && op.tn_log2 == 6)
                GEMM_76_X<<<(unsigned)blocks, 256, gemm_smem_bytes<7, 6, 16, 3>(), stream>>>(p);
            else if (op.tm_log2 == 6 && op.tn_log2 == 6)
                GEMM_66_X<<<(unsigned)blocks, 256, gemm_smem_bytes<6, 6, 16, 4>(), stream>>>(p);
            else
                return cudaErrorInvalidConfiguration;
        } else if (op.tm_log2 == 7 && op.tn_log2 == 6) {
            const Tuning& T = tuning();
            const int kk = op.k - op.ksplit_log2;  // log2 of the K range one CTA walks
            if (kk == 5 && T.store_tile == 2 && blocks >= 2048) {
                // K = 32, >= 2048 tiles (tensor time 1.5x the store time): persistent, warp-specialised, rows stored while the
                // next rows compute: 0.72 -> 0.85 of the FP64 peak at m=15,n=14; at 1024 tiles the 3-stage ring below is as fast
                KParams pp = p;
                pp.raster_group_log2 = T.store_group_log2;
                GEMM_76_WP2<<<2 * num_sms(), 320, gemm_ws_smem_bytes<7, 6, 4, true>(), stream>>>(pp);
            } else if (kk <= T.persist_max_k && blocks > 2ull * (unsigned long long)num_sms())
                // short K, more tiles than CTA slots (the store-bound joins): persistent CTAs prefetch the next tiles
                // (a shared-memory-staged epilogue writing 512-byte runs was measured in round 2 and is 25-30 % SLOWER than
                // the direct 16-byte scatter: profiles/r02b_kernel_lab_tma_staged.md — its barriers serialise the tile)
            {
                KParams pp = p;
                pp.raster_group_log2 = T.store_group_log2;  // raster order of the persistent tile walk
                if (T.store_tile >= 1 && kk == 4)
                    // K = 16 (the DMMA time is 3/4 of the store time): rows leave while the next rows compute, 0.78 -> 0.89 of HBM
                    GEMM_76_P1<<<2 * num_sms(), 256, gemm_smem_bytes<7, 6, 16, 3>(), stream>>>(pp);
                else
                    GEMM_76_P<<<2 * num_sms(), 256, gemm_smem_bytes<7, 6, 16, 3>(), stream>>>(pp);
            }
            else if (op.streamk > 0 && kk >= 8) {
                // stream-K: the K steps of all tiles in equal ranges over the CTA slots (badly quantised tile counts)
                if (!p.sk_flags || !p.ws || op.streamk > kSkFlagBytes / 4 - 1) return cudaErrorInvalidConfiguration;
                GEMM_76_SK<<<(unsigned)op.streamk, 320, gemm_sk_smem_bytes(), stream>>>(p);
            } else if (kk >= 8)
                // warp-specialised pipeline, TWO producer warps issuing LDGSTS into a swizzled 4-stage ring, mbarrier
                // full/empty stages: 34.7-35.2 TFLOP/s on the dominant joins
                GEMM_76_WZ2<<<(unsigned)blocks, 320, gemm_ws_smem_bytes<7, 6, 4, true>(), stream>>>(p);
            else if (kk >= T.ws_min_k)  // K = 32, 64, 128: the shorter 3-stage ring fills faster (K = 32 with >= 2048 tiles went to k_gemm_dmma_p1 above)
                GEMM_76_WL<<<(unsigned)blocks, 288, gemm_ws_smem_bytes<7, 6, 3>(), stream>>>(p);
            else
                GEMM_76<<<(unsigned)blocks, 256, gemm_smem_bytes<7, 6, 16, 3>(), stream>>>(p);
        }
        else if (op.tm_log2 == 6 && op.tn_log2 == 6)
            GEMM_66<<<(unsigned)blocks, 256, gemm_smem_bytes<6, 6, 16, 4>(), stream>>>(p);
        else
            return cudaErrorInvalidConfiguration;
        (*launches)++;
    } else if (op.threads_per_out == 1 && op.m + op.n >= 12) {
        // large outputs: 4 outputs per thread, 32-byte stores
        const unsigned grid = grid_for(outs >> 2, 256, cap);
        switch (op.k) {
            case 0: launch_t1x4<0>(p, grid, stream); break;
            case 1: launch_t1x4<1>(p, grid, stream); break;
            case 2: launch_t1x4<2>(p, grid, stream); break;
            case 3: launch_t1x4<3>(p, grid, stream); break;
            default: launch_t1x4<-1>(p, grid, stream); break;
        }
        (*launches)++;
    } else if (op.threads_per_out == 1) {
        k_generic_t1<<<grid_for(outs, 256, cap), 256, 0, stream>>>(p);
        (*launches)++;
    } else if (op.threads_per_out == 32) {
        k_generic_t32<<<grid_for(outs, 8, cap), 256, 0, stream>>>(p);
        (*launches)++;
    } else {
        k_generic_t256<<<grid_for(outs << op.ksplit_log2, 1, cap), 256, 0, stream>>>(p);
        (*launches)++;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (op.ksplit_log2 > 0) {
        if (outs <= 4096 && op.ksplit_log2 >= 5)
            k_reduce_splits_warp<<<grid_for(outs, 8, cap), 256, 0, stream>>>(p.ws, p.c, outs, 1 << op.ksplit_log2, p.modp, p.inv_modp);
        else
            k_reduce_splits<<<grid_for(outs, 256, (unsigned long long)num_sms() * 16), 256, 0, stream>>>(p.ws, p.c, outs, 1 << op.ksplit_log2, p.modp, p.inv_modp);
        (*launches)++;
        e = cudaGetLastError();
    }
    return e;
}

// ------------------------------------------------------------------------------------------------
// index permutation:  out[o] = in[sigma(o)],  address bit p of o comes from address bit src_bit[p]
// ------------------------------------------------------------------------------------------------
// EPT = tile elements per thread (tile = 256 * EPT elements); OFF = 32-bit element offsets when rank <= 32
// (halves the registers held across the tile loop so 3-4 CTAs fit per SM and loads of one CTA overlap
// the stores of another)
template <int EPT, typename OFF>
__global__ void __launch_bounds__(256, (EPT >= 16 ? (sizeof(OFF) == 8 ? 2 : 3) : 4)) k_permute(PermuteParams p) {
    extern __shared__ double tile[];
    const int tid = threadIdx.x;
    OFF in_off[EPT], out_off[EPT];
    unsigned short tile_idx[EPT];
#pragma unroll
    for (int e = 0; e < EPT; e++) {
        const unsigned ei = tid + 256 * e;  // tile index, input order
        unsigned long long io = 0, oo = 0;
        unsigned ti = 0;
        for (int j = 0; j < p.tbits; j++) {
            const unsigned long long bi = (ei >> j) & 1u;
            io |= bi << p.in_pos[j];
            ti |= (unsigned)bi << p.in_to_tile[j];
            oo |= bi << p.out_pos[j];  // the same index read in output order
        }
        in_off[e] = (OFF)io;
        tile_idx[e] = (unsigned short)(ti ^ (((ti >> 5) ^ (ti >> 10)) & 31u));  // XOR swizzle: strided scatters hit distinct banks
        out_off[e] = (OFF)oo;
    }
    const unsigned long long ntiles = 1ull << p.nrest;
    for (unsigned long long tb = blockIdx.x; tb < ntiles; tb += gridDim.x) {
        unsigned long long base_in = 0, base_out = 0;
        for (int j = 0; j < p.nrest; j++) {
            const unsigned long long b = (tb >> j) & 1ull;
            base_in |= b << p.rest_in[j];
            base_out |= b << p.rest_out[j];
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < EPT; e++) tile[tile_idx[e]] = p.in[base_in | in_off[e]];
        __syncthreads();
#pragma unroll
        for (int e = 0; e < EPT; e++) {
            const unsigned idx = tid + 256 * e;
            p.out[base_out | out_off[e]] = tile[idx ^ (((idx >> 5) ^ (idx >> 10)) & 31u)];
        }
    }
}

// tensors smaller than one tile
__global__ void k_permute_small(const double* in, double* out, int rank, PermuteParams p) {
    const unsigned long long total = 1ull << rank;
    for (unsigned long long o = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; o < total;
         o += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long i = 0;
        for (int j = 0; j < rank; j++) i |= ((o >> p.rest_out[j]) & 1ull) << p.rest_in[j];
        out[o] = in[i];
    }
}

cudaError_t launch_permute(const double* in, double* out, int rank, const int32_t* src_bit, cudaStream_t stream) {
    PermuteParams p;
    p.in = in;
    p.out = out;
    p.rank = rank;
    if (rank < 8) {
        p.tbits = 0;
        p.nrest = rank;
        for (int j = 0; j < rank; j++) { p.rest_out[j] = (uint8_t)j; p.rest_in[j] = (uint8_t)src_bit[j]; }
        k_permute_small<<<1, 256, 0, stream>>>(in, out, rank, p);
        return cudaGetLastError();
    }
    // tile = low TB output bits  U  output bits fed by the low TB input bits, padded to >= 8 bits
    const int tb_knob = tuning().permute_low_bits;  // 0: 6 low bits of both sides from rank 12 on, else 4
    const int TB = tb_knob > 0 ? (rank >= 2 * tb_knob ? tb_knob : 4) : (rank >= 12 ? 6 : 4);
    bool in_tile[64] = {false};
    int tbits = 0;
    for (int q = 0; q < rank; q++)
        if (q < TB || src_bit[q] < TB) { in_tile[q] = true; tbits++; }
    // pad to 4096-element tiles (amortises the per-tile barriers and base computation): alternately the
    // next output bit and the output bit fed by the next input bit, so both run lengths grow
    const int want = rank < 12 ? (rank < 8 ? rank : 8) : 12;
    int inv[64];
    for (int q = 0; q < rank; q++) inv[src_bit[q]] = q;  // input bit -> output bit
    for (int step = 0, oi = TB, ii = TB; tbits < want && (oi < rank || ii < rank); step++) {
        if ((step & 1) == 0 && oi < rank) {
            if (!in_tile[oi]) { in_tile[oi] = true; tbits++; }
            oi++;
        } else if (ii < rank) {
            if (!in_tile[inv[ii]]) { in_tile[inv[ii]] = true; tbits++; }
            ii++;
        } else if (oi < rank) {
            if (!in_tile[oi]) { in_tile[oi] = true; tbits++; }
            oi++;
        }
    }
    for (int q = 0; q < rank && tbits < 8; q++)
        if (!in_tile[q]) { in_tile[q] = true; tbits++; }
    // output order of the tile bits
    int outs[16], n = 0;
    for (int q = 0; q < rank; q++)
        if (in_tile[q]) outs[n++] = q;
    // input order: sort tile bits by their source position
    int order[16];
    for (int j = 0; j < n; j++) order[j] = j;
    for (int a = 0; a < n; a++)
        for (int b = a + 1; b < n; b++)
            if (src_bit[outs[order[b]]] < src_bit[outs[order[a]]]) { int tmp = order[a]; order[a] = order[b]; order[b] = tmp; }
    p.tbits = n;
    for (int j = 0; j < n; j++) {
        p.out_pos[j] = (uint8_t)outs[j];
        p.in_pos[j] = (uint8_t)src_bit[outs[order[j]]];
        p.in_to_tile[j] = (uint8_t)order[j];
    }
    // out_off[] in the kernel is computed from the same thread index interpreted in OUTPUT order
    p.nrest = 0;
    for (int q = 0; q < rank; q++)
        if (!in_tile[q]) { p.rest_out[p.nrest] = (uint8_t)q; p.rest_in[p.nrest] = (uint8_t)src_bit[q]; p.nrest++; }
    const unsigned long long ntiles = 1ull << p.nrest;
    const int bl_env = tuning().permute_ctas_per_sm > 0 ? tuning().permute_ctas_per_sm : 8;
    const unsigned blocks = (unsigned)(ntiles < (unsigned long long)num_sms() * bl_env ? ntiles : (unsigned long long)num_sms() * bl_env);
    const size_t smem = ((size_t)1 << n) * 8;
    if (rank <= 32) {
        switch (n - 8) {
            case 0: k_permute<1, unsigned><<<blocks, 256, smem, stream>>>(p); break;
            case 1: k_permute<2, unsigned><<<blocks, 256, smem, stream>>>(p); break;
            case 2: k_permute<4, unsigned><<<blocks, 256, smem, stream>>>(p); break;
            case 3: k_permute<8, unsigned><<<blocks, 256, smem, stream>>>(p); break;
            case 4: k_permute<16, unsigned><<<blocks, 256, smem, stream>>>(p); break;
            default: return cudaErrorInvalidConfiguration;
        }
    } else {
        switch (n - 8) {
            case 0: k_permute<1, unsigned long long><<<blocks, 256, smem, stream>>>(p); break;
            case 1: k_permute<2, unsigned long long><<<blocks, 256, smem, stream>>>(p); break;
            case 2: k_permute<4, unsigned long long><<<blocks, 256, smem, stream>>>(p); break;
            case 3: k_permute<8, unsigned long long><<<blocks, 256, smem, stream>>>(p); break;
            case 4: k_permute<16, unsigned long long><<<blocks, 256, smem, stream>>>(p); break;
            default: return cudaErrorInvalidConfiguration;
        }
    }
    return cudaGetLastError();
}

}  // namespace tob
