// knn_morton.cu -- spatially ordered sweep for group_knn on large clouds.
//
// Why: the streaming top-k of knn.cu is dominated, on unordered clouds, by its "rare" path --
// with 32 lanes x Q queries looking at unrelated neighbourhoods, some lane finds a candidate
// in most 4-point steps, so the warp keeps leaving the FP32-bound hot loop.  Here both clouds
// are first sorted along a Morton (Z-order) curve, a CTA's queries are therefore spatial
// neighbours, and the sweep over the (sorted) points STARTS at the CTA's own position on the
// curve and works outwards.  After the first tiles every query's k-th distance is nearly final,
// candidates become rare AND coincide across lanes, and the rest of the sweep stays in the
// hot loop.  The result is exactly the same as the unordered sweep: selection and final order
// use the lexicographic key (distance, ORIGINAL index), distances are evaluated in the same
// rounding order, and the permutation is undone when rows are written.
//
// Pipeline (all on the caller's stream, scratch in the caller's workspace):
//   bbox (atomic min/max) -> 30-bit Morton keys tagged with the batch index -> cub radix sort
//   -> gather sorted coordinates + original indices -> knn_morton_kernel.
#include <cub/device/device_radix_sort.cuh>

#include "pp_common.cuh"

namespace pp {
namespace {

constexpr int KM_THREADS = 32;
constexpr int KM_TILE = 64;
constexpr int KM_SUB = 1;                   // sub-tiles per tile, each with its own bounding box
constexpr int KM_SUBLEN = KM_TILE / KM_SUB;  // 64 points = two warps of the box kernel

__device__ __forceinline__ int float_to_ordered(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) {
    return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff);
}

// bbox[0..2] = min (ordered ints), bbox[3..5] = max
__global__ void __launch_bounds__(256)
km_bbox_kernel(const float *__restrict__ xyz, long long n, int *__restrict__ bbox) {
    float lo[3] = {PP_INF, PP_INF, PP_INF}, hi[3] = {-PP_INF, -PP_INF, -PP_INF};
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float v = __ldg(xyz + e * 3 + c);
            if (v == v && fabsf(v) != PP_INF) {  // ignore NaN / inf
                lo[c] = fminf(lo[c], v);
                hi[c] = fmaxf(hi[c], v);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(FULL_MASK, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(FULL_MASK, hi[c], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            atomicMin(bbox + c, float_to_ordered(lo[c]));
            atomicMax(bbox + 3 + c, float_to_ordered(hi[c]));
        }
    }
}

__device__ __forceinline__ unsigned spread10(unsigned v) {  // 10 bits -> every third bit
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__global__ void __launch_bounds__(256)
km_keys_kernel(const float *__restrict__ xyz, int per_cloud, long long n, const int *__restrict__ bbox,
               unsigned long long *__restrict__ keys, unsigned *__restrict__ vals) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    unsigned code = 0;
    unsigned q[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float lo = ordered_to_float(bbox[c]), hi = ordered_to_float(bbox[3 + c]);
        const float ext = hi - lo;
        const float v = __ldg(xyz + e * 3 + c);
        float t = ext > 0.f ? (v - lo) / ext * 1023.f : 0.f;
        t = (t == t) ? fminf(fmaxf(t, 0.f), 1023.f) : 0.f;
        q[c] = (unsigned)t;
    }
    code = spread10(q[0]) | (spread10(q[1]) << 1) | (spread10(q[2]) << 2);
    keys[e] = ((unsigned long long)(e / per_cloud) << 32) | code;
    vals[e] = (unsigned)e;
}

__global__ void __launch_bounds__(256)
km_gather_kernel(const float *__restrict__ xyz, const unsigned *__restrict__ order, int per_cloud,
                 long long n, float *__restrict__ sorted_xyz, int *__restrict__ sorted_idx) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const unsigned src = order[e];
    sorted_xyz[e * 3 + 0] = __ldg(xyz + (size_t)src * 3 + 0);
    sorted_xyz[e * 3 + 1] = __ldg(xyz + (size_t)src * 3 + 1);
    sorted_xyz[e * 3 + 2] = __ldg(xyz + (size_t)src * 3 + 2);
    sorted_idx[e] = (int)(src % (unsigned)per_cloud);
}

// Axis-aligned bounding box of every KM_TILE consecutive sorted points (the sweep's unit of
// work).  Non-finite coordinates are left out: such points only ever produce inf/NaN distances,
// which no list accepts.  box = {lo.x, lo.y, lo.z, hi.x | hi.y, hi.z, -, -}; empty: lo = +inf, hi = -inf.
__global__ void __launch_bounds__(KM_TILE)
km_tilebox_kernel(const float *__restrict__ sorted_xyz, int N, int ntiles, float4 *__restrict__ boxes) {
    __shared__ float red[KM_TILE / 32][6];
    const int b = blockIdx.y, t = blockIdx.x, u = threadIdx.x;
    const int j = t * KM_TILE + u;
    float lo[3] = {PP_INF, PP_INF, PP_INF}, hi[3] = {-PP_INF, -PP_INF, -PP_INF};
    if (j < N) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float v = sorted_xyz[((size_t)b * N + j) * 3 + c];
            if (v == v && fabsf(v) != PP_INF) lo[c] = hi[c] = v;
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(FULL_MASK, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(FULL_MASK, hi[c], o));
        }
    }
    if ((u & 31) == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            red[u >> 5][c] = lo[c];
            red[u >> 5][3 + c] = hi[c];
        }
    }
    __syncthreads();
    // thread 0: the whole tile; threads 1..KM_SUB: sub-tile u-1 (KM_SUBLEN/32 warps each)
    if (u <= KM_SUB) {
        const int w0 = u == 0 ? 0 : (u - 1) * (KM_SUBLEN / 32);
        const int w1 = u == 0 ? KM_TILE / 32 : w0 + KM_SUBLEN / 32;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            lo[c] = PP_INF;
            hi[c] = -PP_INF;
            for (int w = w0; w < w1; w++) {
                lo[c] = fminf(lo[c], red[w][c]);
                hi[c] = fmaxf(hi[c], red[w][3 + c]);
            }
        }
        // per cloud: [ntiles] tile boxes, then [ntiles][KM_SUB] sub-tile boxes
        const size_t slot = u == 0 ? (size_t)t : (size_t)ntiles + (size_t)t * KM_SUB + (u - 1);
        float4 *o = boxes + ((size_t)b * ntiles * (1 + KM_SUB) + slot) * 2;
        o[0] = make_float4(lo[0], lo[1], lo[2], hi[0]);
        o[1] = make_float4(hi[1], hi[2], 0.f, 0.f);
    }
}

// Squared distance between two boxes (0 when they overlap).  NaN/inf propagate, and every
// comparison against the result is written so that "not provably far" means "visit".
__device__ __forceinline__ float km_box_gap2(const float (&alo)[3], const float (&ahi)[3], const float4 b0, const float4 b1) {
    const float blo[3] = {b0.x, b0.y, b0.z}, bhi[3] = {b0.w, b1.x, b1.y};
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float g = fmaxf(0.f, fmaxf(blo[c] - ahi[c], alo[c] - bhi[c]));
        acc = fmaf(g, g, acc);
    }
    return acc;
}
// A tile may be skipped only when even its nearest possible point is farther than every
// query's current k-th distance.  Rounded distances can undershoot the real ones by a few ulp
// (and by absolute denormal steps near zero): hence the margin and the floor.
__device__ __forceinline__ bool km_can_skip(float gap2, float taumax) {
    return gap2 * 0.9999f > taumax && gap2 > 1e-30f;
}

// Insertion of (d0,j0) into the ascending list: it enters in front of the first slot that is
// larger; from there on every slot takes its predecessor (pure shift).  LEX = false compares
// distances only (5 instructions per slot) and reports whether an exactly equal distance was
// met; LEX = true uses the full (distance, original index) key (8 per slot).
template <int K, bool LEX>
__device__ __forceinline__ bool km_insert(float (&ld)[K], int (&li)[K], float d, int j) {
    const float d0 = d;
    const int j0 = j;
    bool tie = false;
#pragma unroll
    for (int s = 0; s < K; s++) {
        const float td = ld[s];
        const int ti = li[s];
        bool sw;
        if (LEX) {
            sw = d0 < td || (d0 == td && j0 < ti);
        } else {
            sw = d0 < td;
            tie |= d0 == td;
        }
        ld[s] = sw ? d : td;
        li[s] = sw ? j : ti;
        d = sw ? td : d;
        j = sw ? ti : j;
    }
    return tie;
}

template <int K, int Q, int KM_CB, bool PRECHECK, bool ESTIMATE>
__global__ void __launch_bounds__(KM_THREADS)
knn_morton_kernel(const float *__restrict__ sq, const int *__restrict__ sqi, const unsigned long long *__restrict__ qkeys,
                  const float *__restrict__ sp, const int *__restrict__ spi, const unsigned long long *__restrict__ pkeys,
                  const float4 *__restrict__ tileboxes, int prune, int M, int N, int k, float *__restrict__ dist,
                  int *__restrict__ idx, unsigned long long *__restrict__ visited) {
    __shared__ __align__(16) float sX[KM_TILE];
    __shared__ __align__(16) float sY[KM_TILE];
    __shared__ __align__(16) float sZ[KM_TILE];
    __shared__ int sI[KM_TILE];
    __shared__ float sBD[Q][KM_CB][KM_THREADS];
    __shared__ int sBI[Q][KM_CB][KM_THREADS];
    __shared__ int s_start;
    __shared__ int s_wtau[2][KM_THREADS / 32];   // per-warp max k-th distance (float bits), double buffered
    __shared__ float s_qbox[KM_THREADS / 32][6];  // per-warp query bounding boxes

    const int b = blockIdx.y;
    const int tid = threadIdx.x;
    const float *qp = sq + (size_t)b * M * 3;
    const float *pp_ = sp + (size_t)b * N * 3;
    const int *pi = spi + (size_t)b * N;
    const int qbase = blockIdx.x * (KM_THREADS * Q);
    const int ntiles = ceil_div(N, KM_TILE);

    // where on the points' curve do this CTA's queries sit?  lower_bound of the CTA's middle
    // query key among the sorted point keys of this cloud (warp 0, 32-ary search)
    if (tid < 32) {
        const int mid = min(M - 1, qbase + (KM_THREADS * Q) / 2);
        const unsigned long long want = qkeys[(size_t)b * M + mid];
        const unsigned long long *pk = pkeys + (size_t)b * N;
        int lo = 0, hi = N;  // answer in [lo, hi]
        while (hi - lo > 0) {
            const int span = hi - lo;
            const int step = (span + 31) / 32;
            const int probe = lo + tid * step;
            const bool below = probe < hi && pk[probe] < want;
            const unsigned m = __ballot_sync(FULL_MASK, below);
            const int nb = __popc(m);  // probes 0..nb-1 are below (keys sorted)
            if (nb == 0) {
                hi = lo;
            } else {
                const int nlo = lo + (nb - 1) * step + 1;
                const int nhi = min(hi, lo + nb * step);
                lo = nlo;
                hi = nhi;
            }
        }
        if (tid == 0) s_start = min(ntiles - 1, lo / KM_TILE);
    }

    float nqx[Q], nqy[Q], nqz[Q], tau[Q], tau0[Q];
    int cnt[Q];
    float ld[Q][K];
    int li[Q][K];
    bool active[Q];
    float wlo[3] = {PP_INF, PP_INF, PP_INF}, whi[3] = {-PP_INF, -PP_INF, -PP_INF};  // this warp's query box
#pragma unroll
    for (int q = 0; q < Q; q++) {
        tau0[q] = PP_INF;
        // a warp's 32*Q queries are consecutive on the curve: its bounding box stays small
        const int i = qbase + (tid >> 5) * (32 * Q) + q * 32 + (tid & 31);
        float x = PP_INF, y = PP_INF, z = PP_INF;
        if (i < M) {
            x = __ldg(qp + (size_t)i * 3);
            y = __ldg(qp + (size_t)i * 3 + 1);
            z = __ldg(qp + (size_t)i * 3 + 2);
        }
        active[q] = i < M;
        {
            const float c3[3] = {x, y, z};
#pragma unroll
            for (int c = 0; c < 3; c++)
                if (i < M && c3[c] == c3[c] && fabsf(c3[c]) != PP_INF) {
                    wlo[c] = fminf(wlo[c], c3[c]);
                    whi[c] = fmaxf(whi[c], c3[c]);
                }
        }
        nqx[q] = -x; nqy[q] = -y; nqz[q] = -z;
        tau[q] = PP_INF;
        cnt[q] = 0;
#pragma unroll
        for (int s = 0; s < K; s++) {
            ld[q][s] = s < k ? PP_INF : -PP_INF;  // slots >= k never accept anything
            li[q][s] = s < k ? 0x7fffffff : -1;
        }
    }

    unsigned n_iter = 0, n_cand = 0, n_stale = 0;
    auto drain = [&]() {
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const int most = __reduce_max_sync(FULL_MASK, cnt[q]);
            n_iter += most; n_cand += cnt[q];
            for (int e = 0; e < most; e++) {
                float d = PP_INF;
                int j = 0x7fffffff;
                if (e < cnt[q]) {
                    d = sBD[q][e][tid];
                    j = sBI[q][e][tid];
                }
                // buffered against an older threshold: by now it may be beaten already (k-th
                // entry of the list); if that holds for every lane the insertion is a no-op
                float kth = ld[q][K - 1];
                if (k < K) {
#pragma unroll
                    for (int s = 0; s < K - 1; s++) kth = (s == k - 1) ? ld[q][s] : kth;
                }
                if (!__any_sync(FULL_MASK, d <= kth)) { n_stale++; continue; }
                // An exactly equal distance already in some lane's list?  Only then does the
                // original index decide and the (more expensive) full-key insertion run.
                if (PRECHECK) {
                    bool tie = false;
#pragma unroll
                    for (int s = 0; s < K; s++) tie |= d == ld[q][s];
                    if (__any_sync(FULL_MASK, tie && d < PP_INF))
                        km_insert<K, true>(ld[q], li[q], d, j);
                    else
                        km_insert<K, false>(ld[q], li[q], d, j);
                } else {
                    km_insert<K, true>(ld[q], li[q], d, j);
                }
            }
            cnt[q] = 0;
            float t = ld[q][0];
#pragma unroll
            for (int s = 1; s < K; s++) t = (s < k) ? ld[q][s] : t;
            tau[q] = fminf(t, tau0[q]);
        }
    };

#pragma unroll
    for (int c = 0; c < 3; c++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            wlo[c] = fminf(wlo[c], __shfl_xor_sync(FULL_MASK, wlo[c], o));
            whi[c] = fmaxf(whi[c], __shfl_xor_sync(FULL_MASK, whi[c], o));
        }
    }
    if ((tid & 31) == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            s_qbox[tid >> 5][c] = wlo[c];
            s_qbox[tid >> 5][3 + c] = whi[c];
        }
    }
    __syncthreads();
    const int t0 = s_start;
    float clo[3], chi[3];  // the CTA's query box
#pragma unroll
    for (int c = 0; c < 3; c++) {
        clo[c] = s_qbox[0][c];
        chi[c] = s_qbox[0][3 + c];
#pragma unroll
        for (int w = 1; w < KM_THREADS / 32; w++) {
            clo[c] = fminf(clo[c], s_qbox[w][c]);
            chi[c] = fmaxf(chi[c], s_qbox[w][3 + c]);
        }
    }
    const float4 *boxes = tileboxes + (size_t)b * ntiles * (1 + KM_SUB) * 2;
    const float4 *subboxes = boxes + (size_t)ntiles * 2;
    float wtaumax = PP_INF;  // this warp's largest k-th distance, refreshed by block_taumax()
    int tau_par = 0;
    // largest k-th distance over the CTA's live queries (block-uniform result; one barrier)
    auto block_taumax = [&]() -> float {
        float m = 0.f;
#pragma unroll
        for (int q = 0; q < Q; q++) m = fmaxf(m, active[q] ? tau[q] : 0.f);
        const int wm = __reduce_max_sync(FULL_MASK, __float_as_int(m));  // non-negative floats order as ints
        wtaumax = __int_as_float(wm);
        if ((tid & 31) == 0) s_wtau[tau_par][tid >> 5] = wm;
        __syncthreads();
        int bm = s_wtau[tau_par][0];
#pragma unroll
        for (int w = 1; w < KM_THREADS / 32; w++) bm = max(bm, s_wtau[tau_par][w]);
        tau_par ^= 1;
        return __int_as_float(bm);
    };
    unsigned long long n_visited = 0;

    // ---- threshold seed: a cheap UPPER BOUND tau0 on every query's k-th distance, taken from the
    // home tile before the real sweep.  The tile is cut into G = K/2 groups; per group the
    // two smallest distances are tracked with three FMNMX per pair; the largest "second
    // smallest" over the groups has 2G >= k distinct points at or below it.  The sweep then
    // starts with the filter d <= tau0 instead of d <= inf, which removes most of the warm-up
    // candidates (the expensive part of a streaming top-k on a few thousand points).
    if (ESTIMATE) {
        const int tile0 = t0 * KM_TILE;
        for (int u = tid; u < KM_TILE; u += KM_THREADS) {
            const int j = tile0 + u;
            float x = PP_INF, y = PP_INF, z = PP_INF;
            if (j < N) {
                x = __ldg(pp_ + (size_t)j * 3);
                y = __ldg(pp_ + (size_t)j * 3 + 1);
                z = __ldg(pp_ + (size_t)j * 3 + 2);
            }
            sX[u] = x; sY[u] = y; sZ[u] = z;
        }
        __syncthreads();
        // K/2 interleaved groups (point u -> group u mod G): every group is a uniform sample of the
        // tile, so each group's second-smallest distance is already close to the k-th distance
        // (consecutive groups would be dominated by the group farthest from the query)
        constexpr int G = K / 2;
        float m1[Q][G], m2[Q][G];
#pragma unroll
        for (int q = 0; q < Q; q++)
#pragma unroll
            for (int g = 0; g < G; g++) m1[q][g] = m2[q][g] = PP_INF;
#pragma unroll 1
        for (int jj0 = 0; jj0 < KM_TILE; jj0 += G) {
#pragma unroll
            for (int v = 0; v < G / 4; v++) {
                const int jj = jj0 + 4 * v;
                const float4 X = *reinterpret_cast<const float4 *>(sX + jj);
                const float4 Y = *reinterpret_cast<const float4 *>(sY + jj);
                const float4 Z = *reinterpret_cast<const float4 *>(sZ + jj);
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    const float2 a = sqdist2_xyz(make_float2(X.x, X.y), make_float2(Y.x, Y.y), make_float2(Z.x, Z.y),
                                                 nqx[q], nqy[q], nqz[q]);
                    const float2 c = sqdist2_xyz(make_float2(X.z, X.w), make_float2(Y.z, Y.w), make_float2(Z.z, Z.w),
                                                 nqx[q], nqy[q], nqz[q]);
                    const float dd[4] = {a.x, a.y, c.x, c.y};
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        m2[q][4 * v + r] = fminf(m2[q][4 * v + r], fmaxf(m1[q][4 * v + r], dd[r]));  // second smallest
                        m1[q][4 * v + r] = fminf(m1[q][4 * v + r], dd[r]);                            // smallest
                    }
                }
            }
        }
        float est[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) {
            est[q] = 0.f;
#pragma unroll
            for (int g = 0; g < G; g++) est[q] = fmaxf(est[q], m2[q][g]);
        }
#pragma unroll
        for (int q = 0; q < Q; q++) {
            tau0[q] = est[q];  // +inf when the home tile is too short: the seed is then simply unused
            tau[q] = est[q];
        }
    }
    // outward sweep: t0, t0+1, t0-1, t0+2, ... (wrapping), nearest tiles first
    auto tile_of = [&](int s) -> int {
        int t = (s & 1) ? t0 + (s + 1) / 2 : t0 - s / 2;
        t %= ntiles;
        return t < 0 ? t + ntiles : t;
    };
    // Exact pruning.  Pass 0 visits the tiles whose box touches the queries' box (that is where the
    // neighbours are, wherever the curve put them), pass 1 the rest -- by then the thresholds
    // are tight and nearly all of them are provably too far.  Evaluation order never changes the
    // result (selection is by the total order (distance, original index)).
    for (int pass = 0; pass < (prune ? 2 : 1); pass++) {
    for (int s0 = 0; s0 < ntiles; s0 += 32) {
      // 32 sweep steps at a time: lane l tests the box of step s0+l against the CTA's query box
      // and the largest k-th distance so far.  Thresholds only shrink, so a tile that is provably
      // too far now stays too far; survivors are re-tested with the fresh threshold right before
      // they are loaded (their gap is fetched from the lane that computed it).
      unsigned todo = 0xffffffffu;
      float gap = 0.f;
      if (prune) {
          const float taumax = block_taumax();
          const int s = s0 + (tid & 31);
          bool need = false;
          if (s < ntiles) {
              const int t = tile_of(s);
              gap = km_box_gap2(clo, chi, __ldg(boxes + t * 2), __ldg(boxes + t * 2 + 1));
              const bool touching = !(gap > 0.f);  // NaN counts as touching: visit early, never skip
              need = pass == 0 ? touching : (!touching && !km_can_skip(gap, taumax));
          }
          todo = __ballot_sync(FULL_MASK, need);
      } else if (ntiles - s0 < 32) {
          todo = (1u << (ntiles - s0)) - 1u;
      }
      while (todo != 0u) {
        const int bit = __ffs(todo) - 1;
        const int s = s0 + bit;
        todo &= todo - 1u;
        const int t = tile_of(s);
        const int tile0 = t * KM_TILE;
        if (prune && pass == 1) {
            const float taumax = block_taumax();  // also the barrier that frees the tile buffers
            if (km_can_skip(__shfl_sync(FULL_MASK, gap, bit), taumax)) continue;
        } else {
            __syncthreads();
        }
        for (int u = tid; u < KM_TILE; u += KM_THREADS) {
            const int j = tile0 + u;
            float x = PP_INF, y = PP_INF, z = PP_INF;  // padding: d = inf
            int oi = 0x7fffffff;
            if (j < N) {
                x = __ldg(pp_ + (size_t)j * 3);
                y = __ldg(pp_ + (size_t)j * 3 + 1);
                z = __ldg(pp_ + (size_t)j * 3 + 2);
                oi = __ldg(pi + j);
            }
            sX[u] = x; sY[u] = y; sZ[u] = z; sI[u] = oi;
        }
        __syncthreads();
        // the tile is here because SOME warp may need it; each warp now tests its own (smaller)
        // query box against the tile's sub-boxes and sweeps only the sub-tiles it cannot rule out
        unsigned sub = (1u << KM_SUB) - 1u;
        if (prune && KM_SUB > 1) {
            bool need = false;
            if ((tid & 31) < KM_SUB) {
                const float4 *sb = subboxes + ((size_t)t * KM_SUB + (tid & 31)) * 2;
                need = !km_can_skip(km_box_gap2(wlo, whi, __ldg(sb), __ldg(sb + 1)), wtaumax);
            }
            sub = __ballot_sync(FULL_MASK, need);
        }
        if ((tid & 31) == 0) n_visited += __popc(sub);
#pragma unroll 1
        for (; sub != 0u; sub &= sub - 1u) {
        const int j0 = (__ffs(sub) - 1) * KM_SUBLEN;
#pragma unroll 1
        for (int jj = j0; jj < j0 + KM_SUBLEN; jj += 4) {
            const float4 X = *reinterpret_cast<const float4 *>(sX + jj);
            const float4 Y = *reinterpret_cast<const float4 *>(sY + jj);
            const float4 Z = *reinterpret_cast<const float4 *>(sZ + jj);
            float2 d01[Q], d23[Q];
            bool cand = false;
#pragma unroll
            for (int q = 0; q < Q; q++) {
                d01[q] = sqdist2_xyz(make_float2(X.x, X.y), make_float2(Y.x, Y.y), make_float2(Z.x, Z.y),
                                     nqx[q], nqy[q], nqz[q]);
                d23[q] = sqdist2_xyz(make_float2(X.z, X.w), make_float2(Y.z, Y.w), make_float2(Z.z, Z.w),
                                     nqx[q], nqy[q], nqz[q]);
                // '<=': an equal distance with a lower original index still has to get in
                cand |= fminf(fmin3(d01[q].x, d01[q].y, d23[q].x), d23[q].y) <= tau[q];
            }
            if (__any_sync(FULL_MASK, cand)) {
                bool full = false;
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    const float dd[4] = {d01[q].x, d01[q].y, d23[q].x, d23[q].y};
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        if (dd[r] <= tau[q] && dd[r] < PP_INF) {
                            sBD[q][cnt[q]][tid] = dd[r];
                            sBI[q][cnt[q]][tid] = sI[jj + r];
                            cnt[q]++;
                        }
                    }
                    full |= cnt[q] > KM_CB - 4;
                }
                if (__any_sync(FULL_MASK, full)) drain();
            }
        }
        }
      }
    }
    }
    drain();
    if (visited != nullptr) {
        if ((tid & 31) == 0) { atomicAdd(visited, n_visited); atomicAdd(visited + 1, (unsigned long long)n_iter); atomicAdd(visited + 3, (unsigned long long)n_stale); }
        atomicAdd(visited + 2, (unsigned long long)n_cand);
    }

    const int *qi = sqi + (size_t)b * M;
#pragma unroll
    for (int q = 0; q < Q; q++) {
        const int i = qbase + (tid >> 5) * (32 * Q) + q * 32 + (tid & 31);
        if (i < M) {
            const int orig = __ldg(qi + i);
            float *od = dist + ((size_t)b * M + orig) * k;
            int *oi = idx + ((size_t)b * M + orig) * k;
#pragma unroll
            for (int s = 0; s < K; s++) {
                if (s < k) {
                    od[s] = ld[q][s];
                    oi[s] = li[q][s] == 0x7fffffff ? -1 : li[q][s];
                }
            }
        }
    }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct KmLayout {
    size_t bbox, keys_in, keys_out, vals_in, vals_out, sorted_xyz, sorted_idx, boxes, cub_temp, cub_bytes, total;
};

// One set of buffers sized for max(nq, np) elements is laid out twice (points, then queries).
KmLayout km_layout(size_t n) {
    KmLayout L;
    size_t off = 0;
    L.bbox = off; off = align_up(off + 6 * sizeof(int) + 64, 256);  // + a 64-bit visit counter at byte 32
    L.keys_in = off; off = align_up(off + n * 8, 256);
    L.keys_out = off; off = align_up(off + n * 8, 256);
    L.vals_in = off; off = align_up(off + n * 4, 256);
    L.vals_out = off; off = align_up(off + n * 4, 256);
    L.sorted_xyz = off; off = align_up(off + n * 12, 256);
    L.sorted_idx = off; off = align_up(off + n * 4, 256);
    // one 32-byte box per tile and per sub-tile; clouds on this path hold >= 4096 points, so there are at most
    // n/KM_TILE + n/4096 tiles
    L.boxes = off; off = align_up(off + (n / KM_TILE + n / 4096 + 2) * (1 + KM_SUB) * 32, 256);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (const unsigned *)nullptr, (unsigned *)nullptr, (long long)n, 0, 64);
    L.cub_bytes = tb;
    L.cub_temp = off; off = align_up(off + tb, 256);
    L.total = off;
    return L;
}

int km_sort_cloud(const float *xyz, int B, int per_cloud, unsigned char *ws, const KmLayout &L, const int *bbox,
                  bool with_boxes, cudaStream_t st) {
    const long long n = (long long)B * per_cloud;
    unsigned long long *keys_in = (unsigned long long *)(ws + L.keys_in), *keys_out = (unsigned long long *)(ws + L.keys_out);
    unsigned *vals_in = (unsigned *)(ws + L.vals_in), *vals_out = (unsigned *)(ws + L.vals_out);
    const unsigned blocks = (unsigned)ceil_div_ll(n, 256);
    km_keys_kernel<<<blocks, 256, 0, st>>>(xyz, per_cloud, n, bbox, keys_in, vals_in);
    PP_LAUNCH_CHECK();
    int batch_bits = 1;
    while ((1 << batch_bits) < B) batch_bits++;
    size_t tb = L.cub_bytes;
    PP_CUDA(cub::DeviceRadixSort::SortPairs(ws + L.cub_temp, tb, keys_in, keys_out, vals_in, vals_out, n, 0,
                                            32 + batch_bits, st));
    km_gather_kernel<<<blocks, 256, 0, st>>>(xyz, vals_out, per_cloud, n, (float *)(ws + L.sorted_xyz),
                                             (int *)(ws + L.sorted_idx));
    PP_LAUNCH_CHECK();
    if (with_boxes) {
        const int ntiles = ceil_div(per_cloud, KM_TILE);
        km_tilebox_kernel<<<dim3(ntiles, B), KM_TILE, 0, st>>>((const float *)(ws + L.sorted_xyz), per_cloud, ntiles,
                                                             (float4 *)(ws + L.boxes));
        PP_LAUNCH_CHECK();
    }
    return PP_OK;
}

}  // namespace

double g_knn_tiles_visited = 0, g_knn_tiles_total = 0;

size_t knn_morton_workspace_bytes(int B, int M, int N) {
    const size_t n = (size_t)B * (size_t)(M > N ? M : N);
    return 2 * km_layout(n).total + 256;
}

// Returns PP_OK after launching everything, or a negative/positive error.
int knn_morton_launch(const float *query, const float *points, int B, int M, int N, int k, float *dist, int *idx,
                      void *workspace, size_t workspace_bytes, cudaStream_t st) {
    const size_t n = (size_t)B * (size_t)(M > N ? M : N);
    const KmLayout L = km_layout(n);
    unsigned char *ws = (unsigned char *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    if (workspace == nullptr || (size_t)(ws - (unsigned char *)workspace) + 2 * L.total > workspace_bytes) {
        set_error("knn: workspace %zu < %zu bytes", workspace_bytes, 2 * L.total + 256);
        return PP_ENOSPC;
    }
    unsigned char *wsP = ws, *wsQ = ws + L.total;
    int *bbox = (int *)(wsP + L.bbox);
    // bbox over both clouds: min <- big positive ints (0x7f7f7f7f), max <- big negative (0x80808080)
    PP_CUDA(cudaMemsetAsync(bbox, 0x7f, 3 * sizeof(int), st));
    PP_CUDA(cudaMemsetAsync(bbox + 3, 0x80, 3 * sizeof(int), st));
    const bool self = (query == points && M == N);
    km_bbox_kernel<<<NUM_SMS_B200 * 2, 256, 0, st>>>(points, (long long)B * N, bbox);
    PP_LAUNCH_CHECK();
    if (!self) {
        km_bbox_kernel<<<NUM_SMS_B200 * 2, 256, 0, st>>>(query, (long long)B * M, bbox);
        PP_LAUNCH_CHECK();
    }
    int rc = km_sort_cloud(points, B, N, wsP, L, bbox, true, st);
    if (rc != PP_OK) return rc;
    if (!self) {
        rc = km_sort_cloud(query, B, M, wsQ, L, bbox, false, st);
        if (rc != PP_OK) return rc;
    } else {
        wsQ = wsP;
    }
    const float *sp = (const float *)(wsP + L.sorted_xyz), *sq = (const float *)(wsQ + L.sorted_xyz);
    const int *spi = (const int *)(wsP + L.sorted_idx), *sqi = (const int *)(wsQ + L.sorted_idx);
    const unsigned long long *pk = (const unsigned long long *)(wsP + L.keys_out), *qk = (const unsigned long long *)(wsQ + L.keys_out);
    const float4 *boxes = (const float4 *)(wsP + L.boxes);
    const int prune = get_option("knn_prune", 1);
    unsigned long long *visited = nullptr;
    if (get_option("knn_stats", 0)) {
        visited = (unsigned long long *)(wsP + L.bbox + 32);
        PP_CUDA(cudaMemsetAsync(visited, 0, 4 * sizeof(unsigned long long), st));
    }
    {
    KernelTimer timer("knn", st);
    // Two tunings of the same kernel (measured on B200, k=16): small clouds sweep few points per
    // candidate, so selection dominates -> deeper candidate buffers and the cheaper
    // distances-only insertion with a tie pre-check (1.39 vs 1.51 ms at B=32 N=8192); large
    // clouds live in the hot loop, where the leaner full-key-only variant wins (18.6 vs 20.6 ms
    // at B=4 N=131072).
    const bool small = get_option("knn_small_tuning", N <= 32768 ? 1 : 0) != 0;
#define KM_LAUNCH(KK, QQ)                                                                                   \
    do {                                                                                                    \
        dim3 grid(ceil_div(M, KM_THREADS * QQ), B);                                                         \
        if (small && est)                                                                                   \
            knn_morton_kernel<KK, QQ, 16, true, true><<<grid, KM_THREADS, 0, st>>>(sq, sqi, qk, sp, spi, pk, boxes, prune, M, N, k, dist, idx, visited); \
        else if (small)                                                                                     \
            knn_morton_kernel<KK, QQ, 16, true, false><<<grid, KM_THREADS, 0, st>>>(sq, sqi, qk, sp, spi, pk, boxes, prune, M, N, k, dist, idx, visited); \
        else if (est)                                                                                       \
            knn_morton_kernel<KK, QQ, 8, false, true><<<grid, KM_THREADS, 0, st>>>(sq, sqi, qk, sp, spi, pk, boxes, prune, M, N, k, dist, idx, visited); \
        else                                                                                                \
            knn_morton_kernel<KK, QQ, 8, false, false><<<grid, KM_THREADS, 0, st>>>(sq, sqi, qk, sp, spi, pk, boxes, prune, M, N, k, dist, idx, visited); \
    } while (0)
    const bool est = get_option("knn_estimate", 1) != 0;
    if (k <= 8) KM_LAUNCH(8, 2);
    else if (k <= 16) { if (get_option("knn_q1", 1)) KM_LAUNCH(16, 1); else KM_LAUNCH(16, 2); }
    else KM_LAUNCH(32, 1);
#undef KM_LAUNCH
    PP_LAUNCH_CHECK();
    }
    if (visited != nullptr) {  // diagnostics only: synchronises the stream
        unsigned long long v4[4] = {0, 0, 0, 0};
        PP_CUDA(cudaMemcpyAsync(v4, visited, sizeof(v4), cudaMemcpyDeviceToHost, st));
        PP_CUDA(cudaStreamSynchronize(st));
        const unsigned long long v = v4[0];
        fprintf(stderr, "[knn stats] warp drain iterations %llu (per warp-query %.1f), lane candidates %llu (per query %.1f), stale-skipped %llu\n", v4[1], (double)v4[1] / ((double)B * M / 32), v4[2], (double)v4[2] / ((double)B * M), v4[3]);
        g_knn_tiles_visited = (double)v;
        const int qper = KM_THREADS * ((k <= 16 && !(k > 8 && get_option("knn_q1", 1))) ? 2 : 1);
        g_knn_tiles_total = (double)B * ceil_div(M, qper) * (KM_THREADS / 32) * ceil_div(N, KM_TILE) * KM_SUB;
    }
    return PP_OK;
}

}  // namespace pp

extern "C" int pp_knn_stats(double *tiles_visited, double *tiles_total) {
    if (tiles_visited) *tiles_visited = pp::g_knn_tiles_visited;
    if (tiles_total) *tiles_total = pp::g_knn_tiles_total;
    return PP_OK;
}
