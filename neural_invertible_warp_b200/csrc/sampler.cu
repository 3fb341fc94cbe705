// Depth sampling.
//   a4  Graph.sample_depth                      reference model/nerf.py:334-344
//   a5  Graph.sample_depth_from_pdf + cat+sort  reference model/nerf.py:346-365, :313-315
// Both are HBM-bound streaming kernels.  Arithmetic uses explicit round-to-nearest intrinsics so
// that no multiply-add is contracted: the depths and, for a5, the bin indices are bit-exact with
// the reference's fp32 CPU evaluation.
#include "common.cuh"
#include <math_constants.h>

namespace {

// Philox4x32-10 (Salmon et al. 2011): four 32-bit words per (counter, key) -- the generator behind the uniforms of the
// device-RNG path (niw_sample_stratified_rng): counter = (group index, call number), key = seed.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}

// ((u+k)/N)*scale + dmin ; inverse: 1/(d+1e-8).  One thread per 4 consecutive samples (128-bit accesses), grid-stride.
// POW2: N is a power of two, so (u+k)/N == (u+k)*(1/N) bit for bit and the sample index is a mask.
// rng != nullptr: the uniforms are drawn here (24-bit, [0, 1), as torch.rand) from Philox4x32-10 keyed by `seed`, counter
// (group, call number rng[0]); the last block to finish advances rng[0] (ticket in rng[1]), so every launch -- and every
// replay of a captured graph -- draws afresh without a torch RNG op (whose graph-safe state costs two fill launches per replay).
template <bool POW2>
__global__ void __launch_bounds__(256)
stratified_kernel(const float* __restrict__ u, int64_t total, int N, float scale, float dmin, int inverse,
                  float* __restrict__ depth, const float* __restrict__ range_dev, unsigned long long* __restrict__ rng,
                  unsigned long long seed) {
    const unsigned long long call = rng ? rng[0] : 0ull;
    if (range_dev) { dmin = range_dev[0]; scale = __fsub_rn(range_dev[1], range_dev[0]); }   // (max - min) as the reference evaluates it
    const float fN = (float)N, rN = 1.0f / (float)N;
    const int64_t ngroups = (total + 3) >> 2, gstride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gi < ngroups; gi += gstride) {
        const int64_t i0 = gi * 4;
        float uu[4] = {0.5f, 0.5f, 0.5f, 0.5f};
        const bool vec = (i0 + 3 < total);
        if (rng) {
            const uint4 x = philox4x32_10(make_uint4((uint32_t)gi, (uint32_t)((uint64_t)gi >> 32), (uint32_t)call, (uint32_t)(call >> 32)),
                                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
            uu[0] = (float)(x.x >> 8) * 5.9604644775390625e-08f; uu[1] = (float)(x.y >> 8) * 5.9604644775390625e-08f;
            uu[2] = (float)(x.z >> 8) * 5.9604644775390625e-08f; uu[3] = (float)(x.w >> 8) * 5.9604644775390625e-08f;
        } else if (u) {
            if (vec) {
                const float4 v = __ldcs(reinterpret_cast<const float4*>(u + i0));
                uu[0] = v.x; uu[1] = v.y; uu[2] = v.z; uu[3] = v.w;
            } else {
                for (int j = 0; j < 4 && i0 + j < total; ++j) uu[j] = u[i0 + j];
            }
        }
        int k = POW2 ? (int)(i0 & (int64_t)(N - 1)) : (int)(i0 % N);
        float out[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float s = __fadd_rn(uu[j], (float)k);
            const float q = POW2 ? __fmul_rn(s, rN) : __fdiv_rn(s, fN);
            float d = __fadd_rn(__fmul_rn(q, scale), dmin);
            if (inverse) d = __frcp_rn(__fadd_rn(d, 1e-8f));        // correctly rounded, == 1/(d+1e-8)
            out[j] = d;
            if (++k == N) k = 0;
        }
        if (vec) {
            __stcs(reinterpret_cast<float4*>(depth + i0), make_float4(out[0], out[1], out[2], out[3]));
        } else {
            for (int j = 0; j < 4 && i0 + j < total; ++j) depth[i0 + j] = out[j];
        }
    }
    if (rng) {      // every block has read rng[0] before it gets here: the last one to arrive moves on to the next call number
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(rng + 1, 1ull) == gridDim.x - 1) { rng[1] = 0ull; __threadfence(); rng[0] = call + 1; }
    }
}

// ---- inverse-CDF sampling + merge, generic path: one warp per ray, any N / Nf ------------------------------
// Shared memory per warp: cdf[N+1] | sort buffer[npow2].
__device__ __noinline__ void pdf_cdf_sequential(const float* __restrict__ p, int N, float* cdf, int stride) {
    // torch.cumsum on CPU: sequential fp64 accumulation, each output rounded to fp32
    double acc = 0.0;
    cdf[0] = 0.f;
    for (int i = 1; i <= N; ++i) { acc += (double)p[i - 1]; cdf[i * stride] = (float)acc; }
}

// bitonic sort of npow2 keys in shared memory by one warp (values only; equals torch.sort(...).values)
__device__ __noinline__ void warp_bitonic_sort(float* buf, int npow2, int lane) {
    for (int k = 2; k <= npow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < npow2; i += 32) {
                int ixj = i ^ j;
                if (ixj > i) {
                    float a = buf[i], b = buf[ixj];
                    bool up = ((i & k) == 0);
                    if ((a > b) == up) { buf[i] = b; buf[ixj] = a; }
                }
            }
            __syncwarp();
        }
    }
}

template <int WARPS>
__global__ void pdf_merge_kernel(const float* __restrict__ pdf, const float* __restrict__ depth_coarse,
                                 const float* __restrict__ unif, const float* __restrict__ bins, int64_t R, int N,
                                 int Nf, int npow2, float* __restrict__ fine, int64_t* __restrict__ idx_out,
                                 float* __restrict__ merged) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per_warp = (N + 1) + npow2;
    float* cdf = smem + (size_t)warp * per_warp;
    float* buf = cdf + (N + 1);
    for (int64_t r = (int64_t)blockIdx.x * WARPS + warp; r < R; r += (int64_t)gridDim.x * WARPS) {
        if (lane == 0) pdf_cdf_sequential(pdf + r * N, N, cdf, 1);
        __syncwarp();
        for (int j = lane; j < Nf; j += 32) {
            float uj = unif[j];
            // searchsorted(right=True): first index with cdf[index] > u, in [0, N+1]
            int lo = 0, hi = N + 1;
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (cdf[mid] <= uj) lo = mid + 1; else hi = mid;
            }
            int il = lo - 1 < 0 ? 0 : lo - 1;
            int ih = lo > N ? N : lo;
            float cl = cdf[il], ch = cdf[ih], dl = bins[il], dh = bins[ih];
            float t = __fdiv_rn(__fsub_rn(uj, cl), __fadd_rn(__fsub_rn(ch, cl), 1e-8f));
            float d = __fadd_rn(dl, __fmul_rn(t, __fsub_rn(dh, dl)));
            if (fine) fine[r * Nf + j] = d;
            if (idx_out) idx_out[r * Nf + j] = (int64_t)lo;
            buf[N + j] = d;
        }
        if (merged) {
            const float* dc = depth_coarse + r * N;
            for (int i = lane; i < N; i += 32) buf[i] = dc[i];
            for (int i = N + Nf + lane; i < npow2; i += 32) buf[i] = CUDART_INF_F;
            __syncwarp();
            warp_bitonic_sort(buf, npow2, lane);
            float* m = merged + r * (N + Nf);
            for (int i = lane; i < N + Nf; i += 32) m[i] = buf[i];
        }
        __syncwarp();
    }
}

// ---- fast path: N = 32*CN, Nf = 32*CF (CN, CF even), one warp per ray, lane-contiguous register chunks ------
//  1. CDF: per-lane fp64 prefix of CN consecutive pdf values + warp scan of the lane totals.  The reference result is
//     the SEQUENTIAL fp64 sum rounded to fp32; the scan associates differently, so every value is certified:
//     (a) if the prefix so far is a sum of multiples of q = (smallest ulp of its non-zero terms) that stays below
//         q 2^53, every fp64 addition in ANY order is exact and the two orders agree bit for bit (the usual case);
//     (b) else, if fp32(v(1-e)) == fp32(v(1+e)) for e = (2N+8) 2^-53 (a bound on the distance of either order from
//         the exact sum of non-negative terms), both orders round to the same fp32;
//     (c) else (a lossy prefix that lands next to an fp32 rounding boundary) the warp recomputes the ray sequentially.
//  2. searchsorted(cdf, u, right=True) for all Nf sorted queries at once: idx_j = #{i : cdf[i] <= u_j}.  When the
//     queries are the exact grid u_j = (2j+1)/(2Nf) (Nf a power of two), cdf[i] <= u_j  <=>  j >= ceil((2Nf cdf[i] - 1)/2)
//     with every operation exact in fp32, so each CDF entry drops one count into a histogram over j and a warp prefix
//     sum yields the indices (no per-query search).  Any other table: per-query binary search.
//  3. merge: coarse and fine depths are both ascending (checked; else bitonic sort), so the sorted union is a
//     two-way merge: every lane finds its split of the two lists on its merge-path diagonal and emits CN+CF outputs.
template <int N> struct VecIO {
    // CN consecutive floats per lane, as 128-bit accesses when possible
    template <int C> static __device__ __forceinline__ void load(const float* __restrict__ p, float (&v)[C]) {
        if constexpr (C % 4 == 0) {
#pragma unroll
            for (int q = 0; q < C / 4; ++q) {
                const float4 t = __ldcs(reinterpret_cast<const float4*>(p) + q);
                v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int q = 0; q < C / 2; ++q) {
                const float2 t = __ldcs(reinterpret_cast<const float2*>(p) + q);
                v[2 * q] = t.x; v[2 * q + 1] = t.y;
            }
        }
    }
    template <int C> static __device__ __forceinline__ void store(float* __restrict__ p, const float (&v)[C]) {
        if constexpr (C % 4 == 0) {
#pragma unroll
            for (int q = 0; q < C / 4; ++q)
                __stcs(reinterpret_cast<float4*>(p) + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
        } else {
#pragma unroll
            for (int q = 0; q < C / 2; ++q) __stcs(reinterpret_cast<float2*>(p) + q, make_float2(v[2 * q], v[2 * q + 1]));
        }
    }
};

__device__ __forceinline__ double shfl_up_f64(double v, int d) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_up_sync(0xffffffffu, lo, d); hi = __shfl_up_sync(0xffffffffu, hi, d);
    return __hiloint2double(hi, lo);
}

template <int CN, int CF, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
pdf_merge_fast_kernel(const float* __restrict__ pdf, const float* __restrict__ depth_coarse,
                      const float* __restrict__ unif, const float* __restrict__ bins, int64_t R,
                      float* __restrict__ fine, int64_t* __restrict__ idx_out, float* __restrict__ merged) {
    constexpr int N = 32 * CN, NF = 32 * CF, CM = CN + CF;
    constexpr int NPOW2 = (N + NF) <= 128 ? 128 : ((N + NF) <= 256 ? 256 : 512);
    // per-warp regions; la | lf | hist are contiguous and double as the bitonic buffer of the fallback
    // cb[i] = (cdf[i], bins[i]) for i <= N, cb[N+1] = cb[N]: the two ends of bin idx are cb[idx-1], cb[idx], no clamps
    constexpr int CDF_W = 2 * (N + 2), LA_W = N + 1, LF_W = NF + 1, PER_WARP = CDF_W + LA_W + LF_W + NF + ((LA_W + LF_W) & 1);
    static_assert(CN % 2 == 0 && CF % 2 == 0, "lane chunks are moved as 64/128-bit vectors");
    static_assert((NF & (NF - 1)) == 0, "the histogram search needs a power-of-two query count");
    static_assert(LA_W + LF_W + NF >= NPOW2, "fallback sort buffer");
    static_assert(PER_WARP % 2 == 0, "8-byte alignment of the (cdf, bin) pairs");
    __shared__ __align__(16) float s_warp[WARPS * PER_WARP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* cb = reinterpret_cast<float2*>(s_warp + warp * PER_WARP);
    float* la = s_warp + warp * PER_WARP + CDF_W;          // coarse depths + inf sentinel
    float* lf = la + LA_W;            // fine depths + inf sentinel
    int* hist = reinterpret_cast<int*>(lf + LF_W);
    for (int i = lane; i <= N + 1; i += 32) cb[i] = make_float2(0.f, bins[i < N ? i : N]);
    float uq[CF];                     // this lane's queries (the same for every ray)
    bool grid_ok = true;
#pragma unroll
    for (int k = 0; k < CF; ++k) {
        uq[k] = unif[lane * CF + k];
        grid_ok = grid_ok && (uq[k] == (float)(2 * (lane * CF + k) + 1) * (1.0f / (float)(2 * NF)));
    }
    const bool exact_grid = __all_sync(0xffffffffu, grid_ok);
    __syncwarp();
    const double eps = (double)(2 * N + 8) * 1.1102230246251565e-16;      // (2N+8) 2^-53

    for (int64_t r = (int64_t)blockIdx.x * WARPS + warp; r < R; r += (int64_t)gridDim.x * WARPS) {
        // ---- 1. CDF ----
        float p[CN], ca[CN];
        VecIO<N>::template load<CN>(pdf + r * N + lane * CN, p);
        if (merged) VecIO<N>::template load<CN>(depth_coarse + r * N + lane * CN, ca);
        double s[CN];
        s[0] = (double)p[0];
#pragma unroll
        for (int k = 1; k < CN; ++k) s[k] = s[k - 1] + (double)p[k];
        double incl = s[CN - 1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = shfl_up_f64(incl, o);
            if (lane >= o) incl += t;
        }
        double excl = shfl_up_f64(incl, 1);
        if (lane == 0) excl = 0.0;
        // smallest exponent field among the non-zero terms of each prefix (255: none yet)
        uint32_t qe[CN];
        bool ok = true;
#pragma unroll
        for (int k = 0; k < CN; ++k) {
            const uint32_t e = __float_as_uint(p[k]) >> 23;
            const uint32_t ek = p[k] > 0.f ? (e ? e : 1u) : 255u;
            qe[k] = k ? (qe[k - 1] < ek ? qe[k - 1] : ek) : ek;
            ok = ok && (p[k] >= 0.f);
        }
        uint32_t qincl = qe[CN - 1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, qincl, o);
            if (lane >= o && t < qincl) qincl = t;
        }
        uint32_t qbefore = __shfl_up_sync(0xffffffffu, qincl, 1);
        if (lane == 0) qbefore = 255u;
        float c[CN];
        double v[CN];
        bool exact = ok;
#pragma unroll
        for (int k = 0; k < CN; ++k) {
            v[k] = excl + s[k];
            c[k] = (float)v[k];
            const uint32_t q = qe[k] < qbefore ? qe[k] : qbefore;
            // q 2^53 with q = 2^(e - 150): the double with exponent field e - 97 + 1023
            exact = exact && (v[k] < __hiloint2double((int)((q + 926u) << 20), 0));
        }
        bool safe = __all_sync(0xffffffffu, exact);
        if (!safe) {
            bool near = !ok;
#pragma unroll
            for (int k = 0; k < CN; ++k) near = near || ((float)(v[k] - v[k] * eps) != (float)(v[k] + v[k] * eps));
            safe = __all_sync(0xffffffffu, !near);
        }
        if (safe) {
#pragma unroll
            for (int k = 0; k < CN; ++k) cb[lane * CN + k + 1].x = c[k];
            if (lane == 31) cb[N + 1].x = c[CN - 1];
            __syncwarp();
        } else {                                                  // a lossy prefix next to an fp32 rounding boundary
            if (lane == 0) pdf_cdf_sequential(pdf + r * N, N, &cb[0].x, 2);
            __syncwarp();
#pragma unroll
            for (int k = 0; k < CN; ++k) c[k] = cb[lane * CN + k + 1].x;
            if (lane == 31) cb[N + 1].x = c[CN - 1];
            __syncwarp();
        }
        // ---- 2. idx_j = #{i in [0, N] : cdf[i] <= u_j} ----
        int lo[CF];
        if (exact_grid) {
#pragma unroll
            for (int k = 0; k < CF; ++k) hist[lane * CF + k] = 0;
            __syncwarp();
#pragma unroll
            for (int k = 0; k < CN; ++k) {
                // first query index j with u_j >= c[k]:  ceil((2 NF c - 1) / 2), exact in fp32
                const float x = __fmul_rn(c[k], (float)(2 * NF));
                const float jf = ceilf(__fmul_rn(__fadd_rn(x, -1.0f), 0.5f));
                if (jf < (float)NF) atomicAdd(&hist[jf > 0.f ? (int)jf : 0], 1);
            }
            __syncwarp();
            int h[CF];
#pragma unroll
            for (int k = 0; k < CF; ++k) h[k] = hist[lane * CF + k];
#pragma unroll
            for (int k = 1; k < CF; ++k) h[k] += h[k - 1];
            int tot = h[CF - 1];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, tot, o);
                if (lane >= o) tot += t;
            }
            const int before = tot - h[CF - 1] + 1;                // + 1: cdf[0] = 0 <= every query
#pragma unroll
            for (int k = 0; k < CF; ++k) lo[k] = before + h[k];
        } else {
#pragma unroll
            for (int k = 0; k < CF; ++k) {
                int a = 0, b = N + 1;
                while (a < b) {
                    const int mid = (a + b) >> 1;
                    if (cb[mid].x <= uq[k]) a = mid + 1; else b = mid;
                }
                lo[k] = a < 1 ? 1 : a;                             // cdf[0] = 0 <= u for the reference's grids
            }
        }
        // ---- inverse CDF: linear interpolation inside the bin (model/nerf.py:357-364) ----
        float fv[CF];
#pragma unroll
        for (int k = 0; k < CF; ++k) {
            const float2 l = cb[lo[k] - 1], h = cb[lo[k]];         // idx in [1, N+1]; entry N+1 repeats entry N
            const float t = __fdiv_rn(__fsub_rn(uq[k], l.x), __fadd_rn(__fsub_rn(h.x, l.x), 1e-8f));
            fv[k] = __fadd_rn(l.y, __fmul_rn(t, __fsub_rn(h.y, l.y)));
        }
        if (fine) VecIO<N>::template store<CF>(fine + r * NF + lane * CF, fv);
        if (idx_out) {
            longlong2* o = reinterpret_cast<longlong2*>(idx_out + r * NF + lane * CF);
#pragma unroll
            for (int k = 0; k < CF; k += 2) __stcs(o + k / 2, make_longlong2((long long)lo[k], (long long)lo[k + 1]));
        }
        // ---- 3. merged = sort(coarse ++ fine) ----
        if (merged) {
            // both lists ascending?  (comparisons are false on NaN -> fallback)
            bool asc = true;
#pragma unroll
            for (int k = 1; k < CN; ++k) asc = asc && (ca[k - 1] <= ca[k]);
#pragma unroll
            for (int k = 1; k < CF; ++k) asc = asc && (fv[k - 1] <= fv[k]);
            const float na = __shfl_down_sync(0xffffffffu, ca[0], 1), nf = __shfl_down_sync(0xffffffffu, fv[0], 1);
            if (lane < 31) asc = asc && (ca[CN - 1] <= na) && (fv[CF - 1] <= nf);
            asc = asc && (ca[CN - 1] < CUDART_INF_F) && (fv[CF - 1] < CUDART_INF_F);   // the merge uses +inf sentinels
            asc = __all_sync(0xffffffffu, asc);
            __syncwarp();                                          // everyone is done with hist / cdf reads above
#pragma unroll
            for (int k = 0; k < CN; ++k) la[lane * CN + k] = ca[k];
            float* fdst = asc ? lf : la + N;                      // fallback: one contiguous buffer
#pragma unroll
            for (int k = 0; k < CF; ++k) fdst[lane * CF + k] = fv[k];
            float* m = merged + r * (N + NF);
            if (asc) {
                if (lane == 0) { la[N] = CUDART_INF_F; lf[NF] = CUDART_INF_F; }
                __syncwarp();
                const int d0 = lane * CM;
                // merge path: number of coarse elements among the first d0 outputs = first x in [a, a + n] with
                // la[x] > lf[d0 - 1 - x]; fixed trip count (n <= min(N, NF)), no divergence
                int a = d0 - NF > 0 ? d0 - NF : 0, n = (d0 < N ? d0 : N) - a;
                constexpr int TRIPS = (N < NF ? N : NF) >= 128 ? 8 : ((N < NF ? N : NF) >= 64 ? 7 : 6);
#pragma unroll
                for (int it = 0; it < TRIPS; ++it) {
                    const int half = n >> 1, mid = a + half;
                    const bool go = n > 0 && la[mid] <= lf[d0 - 1 - mid];      // n == 0: mid is still a valid index pair
                    a = go ? mid + 1 : a;
                    n = go ? n - half - 1 : half;
                }
                int ai = a, fi = d0 - a;
                float av = la[ai], bv = lf[fi], out[CM];
#pragma unroll
                for (int k = 0; k < CM; ++k) {
                    const bool take_a = av <= bv;
                    out[k] = take_a ? av : bv;
                    if (take_a) { ++ai; av = la[ai]; } else { ++fi; bv = lf[fi]; }
                }
                VecIO<N>::template store<CM>(m + d0, out);
            } else {
                float* buf = la;
                for (int i = N + NF + lane; i < NPOW2; i += 32) buf[i] = CUDART_INF_F;
                __syncwarp();
                warp_bitonic_sort(buf, NPOW2, lane);
                for (int i = lane; i < N + NF; i += 32) m[i] = buf[i];
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// Random pixel subset: the first k entries of a random permutation of [0, n) -- what the reference draws with
// torch.randperm(H*W)[:rand_rays//B] (model/nerf.py:268), which sorts H*W random keys (5 radix-sort passes,
// ~70 us at 480x640) to keep 64 of them.  Here entry i is pi(i) for a keyed bijection pi of [0, 2^b) (4-round
// Feistel network, b = bit length of n-1 rounded up to even) with cycle walking back into [0, n): a prefix of a
// permutation by construction, O(k) work, no global state besides a call counter (so CUDA-graph replays draw new
// subsets).  Different random stream than torch's generator: used where the draw itself is the only requirement.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t x) {      // lowbias32
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__global__ void sample_pixels_kernel(int64_t n, int k, uint64_t seed, unsigned long long* __restrict__ counter,
                                     int64_t* __restrict__ out) {
    const unsigned long long call = *counter;
    int bits = 2;
    while ((1ull << bits) < (unsigned long long)n) bits += 2;
    const int hb = bits / 2;
    const uint32_t hmask = (1u << hb) - 1u;
    uint32_t key[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) key[r] = mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + 0x9E3779B9u * (uint32_t)(4 * call + r + 1)) ^ (uint32_t)(call >> 32));
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) {
        uint64_t x = (uint64_t)i;
        do {                                             // cycle walking: expected < 4 trips (2^b < 4n)
            uint32_t L = (uint32_t)(x >> hb) & hmask, R = (uint32_t)x & hmask;
#pragma unroll
            for (int r = 0; r < 4; ++r) { const uint32_t t = L ^ (mix32(R ^ key[r]) & hmask); L = R; R = t; }
            x = ((uint64_t)L << hb) | R;
        } while (x >= (uint64_t)n);
        out[i] = (int64_t)x;
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) *counter = call + 1;      // single block (k <= a few thousand)
}

}  // namespace

static int launch_stratified(const float* u, int64_t n_rays, int N, float scale, float dmin, const float* range_dev,
                             int inverse, float* depth, cudaStream_t st, unsigned long long* rng = nullptr,
                             unsigned long long seed = 0) {
    int64_t total = n_rays * N;
    int64_t blocks = niw_blocks((total + 3) / 4, 256);
    const int64_t cap = (int64_t)niw_num_sms() * 16;              // grid-stride: 8 resident CTAs / SM x 2 rounds
    if (blocks > cap) blocks = cap;
    niw::note_launch();
    if ((N & (N - 1)) == 0)
        stratified_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(u, total, N, scale, dmin, inverse, depth, range_dev, rng, seed);
    else
        stratified_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(u, total, N, scale, dmin, inverse, depth, range_dev, rng, seed);
    NIW_LAUNCH_CHECK();
    return 0;
}

extern "C" int niw_sample_stratified(const float* u, int64_t n_rays, int N, float scale, float dmin, int inverse,
                                     float* depth, void* stream) {
    NIW_CHECK_ARG(depth && n_rays > 0 && N > 0);
    return launch_stratified(u, n_rays, N, scale, dmin, nullptr, inverse, depth, niw_stream(stream));
}

extern "C" int niw_sample_stratified_dev(const float* u, int64_t n_rays, int N, const float* range_dev, int inverse,
                                         float* depth, void* stream) {
    NIW_CHECK_ARG(depth && range_dev && n_rays > 0 && N > 0);
    return launch_stratified(u, n_rays, N, 0.f, 0.f, range_dev, inverse, depth, niw_stream(stream));
}

// stratified depths with the uniforms drawn inside the kernel; range_dev (device [min, max]) may be NULL: then scale / dmin.
// rng: two zero-initialised 64-bit device words owned by the caller (call number, block ticket)
extern "C" int niw_sample_stratified_rng(int64_t n_rays, int N, float scale, float dmin, const float* range_dev, int inverse,
                                         unsigned long long seed, unsigned long long* rng, float* depth, void* stream) {
    NIW_CHECK_ARG(depth && rng && n_rays > 0 && N > 0);
    return launch_stratified(nullptr, n_rays, N, scale, dmin, range_dev, inverse, depth, niw_stream(stream), rng, seed);
}

template <int CN, int CF>
static int launch_pdf_fast(const float* pdf, const float* depth_coarse, const float* unif, const float* bins, int64_t R,
                           float* fine, int64_t* idx, float* merged, cudaStream_t st) {
    constexpr int WARPS = 8;
    static const unsigned full = niw_resident_grid(pdf_merge_fast_kernel<CN, CF, WARPS>, WARPS * 32, 0, INT32_MAX);
    int64_t blocks = (R + WARPS - 1) / WARPS;
    if (blocks > full) blocks = full;
    niw::note_launch();
    pdf_merge_fast_kernel<CN, CF, WARPS><<<(unsigned)blocks, WARPS * 32, 0, st>>>(pdf, depth_coarse, unif, bins, R, fine, idx, merged);
    NIW_LAUNCH_CHECK();
    return 0;
}

extern "C" int niw_sample_pdf_merge(const float* pdf, const float* depth_coarse, const float* unif, const float* bins,
                                    int64_t R, int N, int Nf, float* fine, int64_t* idx, float* merged,
                                    void* stream) {
    NIW_CHECK_ARG(pdf && unif && bins && R > 0 && N > 0 && Nf > 0 && (!merged || depth_coarse));
    cudaStream_t st = niw_stream(stream);
    // register-resident warp-per-ray kernel for the shapes the reference configurations use
    // (options/nerf_inn_dtu.yaml: 64 + 128; 128 + 128 and 64 + 64 for scaled variants); anything else: generic kernel
    const bool al = ((uintptr_t)pdf % 16 == 0) && (!depth_coarse || (uintptr_t)depth_coarse % 16 == 0) &&
                    (!fine || (uintptr_t)fine % 16 == 0) && (!idx || (uintptr_t)idx % 16 == 0) &&
                    (!merged || (uintptr_t)merged % 16 == 0);
    if (al && N == 64 && Nf == 128) return launch_pdf_fast<2, 4>(pdf, depth_coarse, unif, bins, R, fine, idx, merged, st);
    if (al && N == 128 && Nf == 128) return launch_pdf_fast<4, 4>(pdf, depth_coarse, unif, bins, R, fine, idx, merged, st);
    if (al && N == 64 && Nf == 64) return launch_pdf_fast<2, 2>(pdf, depth_coarse, unif, bins, R, fine, idx, merged, st);
    if (al && N == 128 && Nf == 64) return launch_pdf_fast<4, 2>(pdf, depth_coarse, unif, bins, R, fine, idx, merged, st);
    int npow2 = 1;
    while (npow2 < N + Nf) npow2 <<= 1;
    if (npow2 > 4096) return NIW_E_UNSUPP;
    constexpr int WARPS = 4;
    size_t smem = sizeof(float) * WARPS * ((size_t)(N + 1) + npow2);
    if (smem > 200 * 1024) return NIW_E_UNSUPP;
    if (smem > 48 * 1024)
        NIW_CUDA(cudaFuncSetAttribute(pdf_merge_kernel<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (R + WARPS - 1) / WARPS;
    int64_t cap = (int64_t)niw_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    niw::note_launch(), pdf_merge_kernel<WARPS><<<(unsigned)blocks, WARPS * 32, smem, st>>>(
        pdf, depth_coarse, unif, bins, R, N, Nf, npow2, fine, idx, merged);
    NIW_LAUNCH_CHECK();
    return 0;
}

extern "C" int niw_sample_pixels(int64_t n, int k, uint64_t seed, uint64_t* counter, int64_t* out, void* stream) {
    NIW_CHECK_ARG(counter && out && n > 0 && k > 0 && k <= n && n <= (int64_t(1) << 40));
    // one block: the counter is read by every thread before the last one advances it
    niw::note_launch(), sample_pixels_kernel<<<1, 1024, 0, niw_stream(stream)>>>(n, k, seed, reinterpret_cast<unsigned long long*>(counter), out);
    NIW_LAUNCH_CHECK();
    return 0;
}
