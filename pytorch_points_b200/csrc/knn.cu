// knn.cu -- group_knn core (pairwise squared distance + top-k) for sm_100a.
//
// The reference snapshot ships no KNN kernel: its callers use
// pytorch3d.ops.knn_points.  Contract implemented here (include/pp_b200.h):
// squared distance in the Chamfer rounding order, k neighbours per query sorted
// ascending by (distance, index).
//
// Design: streaming fused distance + selection, the B x M x N distance matrix is
// never materialised (config 4 would need 275 GB).
//   * thread owns Q=2 queries in registers; reference points stream through a
//     shared-memory SoA tile, four per LDS.128 broadcast, distances evaluated in
//     packed FADD2/FMUL2/FFMA2 pairs exactly as in chamfer.cu;
//   * hot path per (query, 4 references): two FMNMX + one compare against the
//     query's current k-th distance tau -- nothing else;
//   * rare path: candidates (d < tau) are appended to a small per-query buffer in
//     shared memory; when any lane of the warp runs out of buffer space the whole
//     warp merges its buffers into the per-query sorted lists (lane-parallel
//     insertion, so the divergent work is shared by all 32 lanes) and refreshes tau.
//   * references are visited in ascending index order by a single thread per query,
//     so "strict < with stable insertion" yields the (distance, index) order.
#include "pp_common.cuh"

namespace pp {
namespace {

constexpr int KNN_THREADS = 128;
constexpr int KNN_Q = 2;
constexpr int KNN_LISTS = KNN_THREADS * KNN_Q;
constexpr int KNN_TILE = 256;
constexpr int KNN_CB = 12;  // candidate buffer entries per query

struct KnnSmem {
    float *x, *y, *z;  // [KNN_TILE]
    float *ld;         // [k][KNN_LISTS] sorted list distances
    int *li;           // [k][KNN_LISTS] sorted list indices
    float *bd;         // [KNN_CB][KNN_LISTS] candidate buffer
    int *bi;
};

__device__ __forceinline__ void knn_flush(const KnnSmem &s, int list, int k, int &cnt, float &tau) {
    for (int e = 0; e < cnt; e++) {
        const float d = s.bd[e * KNN_LISTS + list];
        const int j = s.bi[e * KNN_LISTS + list];
        if (d < s.ld[(k - 1) * KNN_LISTS + list]) {
            int pos = k - 1;
            while (pos > 0) {
                const float pd = s.ld[(pos - 1) * KNN_LISTS + list];
                if (!(d < pd)) break;  // stable: equal distance stays behind the earlier index
                s.ld[pos * KNN_LISTS + list] = pd;
                s.li[pos * KNN_LISTS + list] = s.li[(pos - 1) * KNN_LISTS + list];
                pos--;
            }
            s.ld[pos * KNN_LISTS + list] = d;
            s.li[pos * KNN_LISTS + list] = j;
        }
    }
    cnt = 0;
    tau = s.ld[(k - 1) * KNN_LISTS + list];
}

__global__ void __launch_bounds__(KNN_THREADS)
knn_kernel(const float *__restrict__ query, const float *__restrict__ points, int M, int N, int k,
           float *__restrict__ dist, int *__restrict__ idx) {
    extern __shared__ __align__(16) unsigned char knn_smem_raw[];
    KnnSmem s;
    s.x = reinterpret_cast<float *>(knn_smem_raw);
    s.y = s.x + KNN_TILE;
    s.z = s.y + KNN_TILE;
    s.ld = s.z + KNN_TILE;
    s.li = reinterpret_cast<int *>(s.ld + (size_t)k * KNN_LISTS);
    s.bd = reinterpret_cast<float *>(s.li + (size_t)k * KNN_LISTS);
    s.bi = reinterpret_cast<int *>(s.bd + KNN_CB * KNN_LISTS);

    const int b = blockIdx.y;
    const int tid = threadIdx.x;
    const float *qp = query + (size_t)b * M * 3;
    const float *pp_ = points + (size_t)b * N * 3;
    const int qbase = blockIdx.x * KNN_LISTS;

    // query q of this thread: qbase + q*KNN_THREADS + tid  (list index q*KNN_THREADS + tid)
    float nqx[KNN_Q], nqy[KNN_Q], nqz[KNN_Q], tau[KNN_Q];
    int cnt[KNN_Q];
#pragma unroll
    for (int q = 0; q < KNN_Q; q++) {
        const int i = qbase + q * KNN_THREADS + tid;
        float x = PP_INF, y = PP_INF, z = PP_INF;
        if (i < M) {
            x = __ldg(qp + (size_t)i * 3);
            y = __ldg(qp + (size_t)i * 3 + 1);
            z = __ldg(qp + (size_t)i * 3 + 2);
        }
        nqx[q] = -x; nqy[q] = -y; nqz[q] = -z;
        tau[q] = PP_INF;
        cnt[q] = 0;
        const int list = q * KNN_THREADS + tid;
        for (int e = 0; e < k; e++) {
            s.ld[e * KNN_LISTS + list] = PP_INF;
            s.li[e * KNN_LISTS + list] = -1;
        }
    }

    for (int tile0 = 0; tile0 < N; tile0 += KNN_TILE) {
        __syncthreads();
        for (int t = tid; t < KNN_TILE; t += KNN_THREADS) {
            const int j = tile0 + t;
            float x = PP_INF, y = PP_INF, z = PP_INF;  // padding: d = inf, never < tau
            if (j < N) {
                x = __ldg(pp_ + (size_t)j * 3);
                y = __ldg(pp_ + (size_t)j * 3 + 1);
                z = __ldg(pp_ + (size_t)j * 3 + 2);
            }
            s.x[t] = x; s.y[t] = y; s.z[t] = z;
        }
        __syncthreads();
#pragma unroll 1
        for (int jj = 0; jj < KNN_TILE; jj += 4) {
            const float4 X = *reinterpret_cast<const float4 *>(s.x + jj);
            const float4 Y = *reinterpret_cast<const float4 *>(s.y + jj);
            const float4 Z = *reinterpret_cast<const float4 *>(s.z + jj);
            bool full = false;
#pragma unroll
            for (int q = 0; q < KNN_Q; q++) {
                const float2 d01 = sqdist2_xyz(make_float2(X.x, X.y), make_float2(Y.x, Y.y),
                                               make_float2(Z.x, Z.y), nqx[q], nqy[q], nqz[q]);
                const float2 d23 = sqdist2_xyz(make_float2(X.z, X.w), make_float2(Y.z, Y.w),
                                               make_float2(Z.z, Z.w), nqx[q], nqy[q], nqz[q]);
                const float mn = fminf(fmin3(d01.x, d01.y, d23.x), d23.y);
                if (mn < tau[q]) {
                    const int list = q * KNN_THREADS + tid;
                    const float dd[4] = {d01.x, d01.y, d23.x, d23.y};
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        if (dd[r] < tau[q]) {
                            s.bd[cnt[q] * KNN_LISTS + list] = dd[r];
                            s.bi[cnt[q] * KNN_LISTS + list] = tile0 + jj + r;
                            cnt[q]++;
                        }
                    }
                    full |= cnt[q] > KNN_CB - 4;
                }
            }
            if (__any_sync(FULL_MASK, full)) {
#pragma unroll
                for (int q = 0; q < KNN_Q; q++) knn_flush(s, q * KNN_THREADS + tid, k, cnt[q], tau[q]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < KNN_Q; q++) knn_flush(s, q * KNN_THREADS + tid, k, cnt[q], tau[q]);
    __syncthreads();
    // coalesced write-out: consecutive threads write consecutive (query, slot) elements
    const int nq = min(KNN_LISTS, M - qbase);
    float *od = dist + ((size_t)b * M + qbase) * k;
    int *oi = idx + ((size_t)b * M + qbase) * k;
    for (int t = tid; t < nq * k; t += KNN_THREADS) {
        const int ql = t / k, e = t % k;  // ql = local query (qbase + ql)
        // local query ql lives in list (ql / KNN_THREADS)*KNN_THREADS + ql % KNN_THREADS == ql
        od[t] = s.ld[e * KNN_LISTS + ql];
        oi[t] = s.li[e * KNN_LISTS + ql];
    }
}

// ---------------------------------------------------------------------------
// Register-list kernel (k <= 32): the k best (distance, index) pairs of each query live in
// REGISTERS, sorted ascending.  Candidates found by the hot loop are parked in a small
// per-thread shared-memory buffer; when any lane of the warp is about to run out of buffer the
// whole warp drains its buffers in lock-step: every drain pass pushes one buffered candidate
// per lane (or +inf for lanes that have none) through a branch-free compare-exchange chain
// over the K slots, so the selection work is shared by all 32 lanes instead of diverging.
// ---------------------------------------------------------------------------
constexpr int KR_THREADS = 128;
constexpr int KR_TILE = 256;
constexpr int KR_CB = 8;  // buffered candidates per query

template <int K>
__device__ __forceinline__ void knn_insert(float (&ld)[K], int (&li)[K], float d, int j) {
    // (d, j) enters at the first slot whose distance is strictly larger; everything after it
    // shifts down one slot; the former last entry falls out.  Strict '<' keeps the earlier
    // (lower) index in front on equal distances, because candidates arrive in ascending index.
    // The list is sorted, so (d0 < ld[s]) is false..false,true..true: from the first true on,
    // every slot takes its predecessor (a pure shift, which keeps equal distances in order).
    const float d0 = d;
#pragma unroll
    for (int s = 0; s < K; s++) {
        const float td = ld[s];
        const int ti = li[s];
        const bool sw = d0 < td;
        ld[s] = sw ? d : td;
        li[s] = sw ? j : ti;
        d = sw ? td : d;
        j = sw ? ti : j;
    }
}

template <int K, int Q>
__global__ void __launch_bounds__(KR_THREADS)
knn_reg_kernel(const float *__restrict__ query, const float *__restrict__ points, int M, int N, int k,
               float *__restrict__ dist, int *__restrict__ idx) {
    __shared__ __align__(16) float sX[KR_TILE];
    __shared__ __align__(16) float sY[KR_TILE];
    __shared__ __align__(16) float sZ[KR_TILE];
    __shared__ float sBD[Q][KR_CB][KR_THREADS];
    __shared__ int sBI[Q][KR_CB][KR_THREADS];

    const int b = blockIdx.y;
    const int tid = threadIdx.x;
    const float *qp = query + (size_t)b * M * 3;
    const float *pp_ = points + (size_t)b * N * 3;
    const int qbase = blockIdx.x * (KR_THREADS * Q);

    float nqx[Q], nqy[Q], nqz[Q], tau[Q];
    int cnt[Q];
    float ld[Q][K];
    int li[Q][K];
#pragma unroll
    for (int q = 0; q < Q; q++) {
        const int i = qbase + q * KR_THREADS + tid;
        float x = PP_INF, y = PP_INF, z = PP_INF;
        if (i < M) {
            x = __ldg(qp + (size_t)i * 3);
            y = __ldg(qp + (size_t)i * 3 + 1);
            z = __ldg(qp + (size_t)i * 3 + 2);
        }
        nqx[q] = -x; nqy[q] = -y; nqz[q] = -z;
        tau[q] = PP_INF;
        cnt[q] = 0;
#pragma unroll
        for (int s = 0; s < K; s++) {
            // slots >= k are never reported: keep them at -inf so they never accept anything
            ld[q][s] = s < k ? PP_INF : -PP_INF;
            li[q][s] = -1;
        }
    }

    auto drain = [&]() {
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const int most = __reduce_max_sync(FULL_MASK, cnt[q]);
            for (int e = 0; e < most; e++) {
                float d = PP_INF;
                int j = -1;
                if (e < cnt[q]) {
                    d = sBD[q][e][tid];
                    j = sBI[q][e][tid];
                }
                knn_insert<K>(ld[q], li[q], d, j);
            }
            cnt[q] = 0;
            // current k-th best; slots >= k hold -inf so index k-1 is found with a static select
            float t = ld[q][0];
#pragma unroll
            for (int s = 1; s < K; s++) t = (s < k) ? ld[q][s] : t;
            tau[q] = t;
        }
    };

    for (int tile0 = 0; tile0 < N; tile0 += KR_TILE) {
        __syncthreads();
        for (int t = tid; t < KR_TILE; t += KR_THREADS) {
            const int j = tile0 + t;
            float x = PP_INF, y = PP_INF, z = PP_INF;  // padding: d = inf, never < tau
            if (j < N) {
                x = __ldg(pp_ + (size_t)j * 3);
                y = __ldg(pp_ + (size_t)j * 3 + 1);
                z = __ldg(pp_ + (size_t)j * 3 + 2);
            }
            sX[t] = x; sY[t] = y; sZ[t] = z;
        }
        __syncthreads();
#pragma unroll 1
        for (int jj = 0; jj < KR_TILE; jj += 4) {
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
                cand |= fminf(fmin3(d01[q].x, d01[q].y, d23[q].x), d23[q].y) < tau[q];
            }
            // hot path ends here: one vote, one (warp-uniform) branch per 4 points x Q queries
            if (__any_sync(FULL_MASK, cand)) {
                bool full = false;
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    const float dd[4] = {d01[q].x, d01[q].y, d23[q].x, d23[q].y};
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        if (dd[r] < tau[q]) {
                            sBD[q][cnt[q]][tid] = dd[r];
                            sBI[q][cnt[q]][tid] = tile0 + jj + r;
                            cnt[q]++;
                        }
                    }
                    full |= cnt[q] > KR_CB - 4;
                }
                if (__any_sync(FULL_MASK, full)) drain();
            }
        }
    }
    drain();

#pragma unroll
    for (int q = 0; q < Q; q++) {
        const int i = qbase + q * KR_THREADS + tid;
        if (i < M) {
            float *od = dist + ((size_t)b * M + i) * k;
            int *oi = idx + ((size_t)b * M + i) * k;
#pragma unroll
            for (int s = 0; s < K; s++) {
                if (s < k) {
                    od[s] = ld[q][s];
                    oi[s] = li[q][s];
                }
            }
        }
    }
}

// Generic point dimension (c != 3): simple thread-per-query kernel, candidates inserted
// directly.  Correctness path only.
__global__ void __launch_bounds__(128)
knn_generic_kernel(const float *__restrict__ query, const float *__restrict__ points, int M, int N,
                   int c, int k, float *__restrict__ dist, int *__restrict__ idx) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const float *q = query + ((size_t)b * M + i) * c;
    const float *p = points + (size_t)b * N * c;
    float *od = dist + ((size_t)b * M + i) * k;
    int *oi = idx + ((size_t)b * M + i) * k;
    for (int e = 0; e < k; e++) {
        od[e] = PP_INF;
        oi[e] = -1;
    }
    for (int j = 0; j < N; j++) {
        float d = 0.f;
        for (int cc = 0; cc < c; cc++) {
            const float t = __fsub_rn(__ldg(p + (size_t)j * c + cc), q[cc]);
            d = __fmaf_rn(t, t, d);
        }
        if (d < od[k - 1]) {
            int pos = k - 1;
            while (pos > 0 && d < od[pos - 1]) {
                od[pos] = od[pos - 1];
                oi[pos] = oi[pos - 1];
                pos--;
            }
            od[pos] = d;
            oi[pos] = j;
        }
    }
}

size_t knn_smem_bytes(int k) {
    return sizeof(float) * 3 * KNN_TILE + (size_t)k * KNN_LISTS * 8 + (size_t)KNN_CB * KNN_LISTS * 8;
}

}  // namespace
}  // namespace pp

using namespace pp;

// Spatially ordered sweep (knn_morton.cu) pays off once the clouds are large enough for the
// sort (a handful of small launches) to vanish next to the distance work.
static bool knn_use_morton(int B, int M, int N, int c, int k) {
    if (c != 3 || k > 32) return false;
    const int opt = get_option("knn_morton", -1);
    if (opt >= 0) return opt != 0;
    // measured crossover on B200 (k = 16): with the one-launch preparation (clouds <= 16384 points)
    // the ordered sweep wins from N = 2048 on, whatever the batch (0.196 vs 0.200 ms at B = 2,
    // 0.236 vs 0.265 ms at B = 32); at N = 1024 the streaming kernel is still ahead
    return N >= 2048;
}

// Tensor-core path (knn_tc.cu): the dense contraction on tcgen05.mma, exact resolution on the FP32 pipe.
static bool knn_use_tc(int B, int M, int N, int c, int k) {
    if (c != 3 || !knn_tc_supported(B, M, N, k)) return false;
    const int opt = get_option("knn_tc", -1);
    if (opt >= 0) return opt != 0;
    // measured on B200 (tools/knn_tc_check.py, profiles/r02_knn_tc.txt):
    //  * lists wider than 16 (K = k + 1 = 17, 20, 21 are what the snapshot's callers ask for): ahead everywhere,
    //    0.54 vs 0.76 ms at K = 17, B = 32, N = 8192; 1.27 vs 1.82 ms at B = 4, N = 131072 -- the ordered sweep's
    //    per-candidate insertion switches to full-warp lists there, the exact pass only ranks a few more keys;
    //  * k <= 16: ahead on small jobs (up to ~65536 queries per call, clouds up to 16384 points: 0.125 vs 0.156 ms
    //    at B = 2, N = 2048; 0.19 vs 0.22 at B = 4, N = 8192), on par or slightly behind on large ones
    //    (0.50 vs 0.47 ms at B = 32, N = 8192), where the ordered sweep stays.
    if (get_option("knn_morton", -1) >= 0 || N < 2048) return false;
    if (k > 16) return true;
    return N <= 16384 && (long long)B * M <= 65536;
}

extern "C" size_t pp_knn_workspace_bytes(int B, int M, int N, int c, int k) {
    if (B <= 0 || M <= 0 || N <= 0) return 0;
    if (c != 3 || k > 32) return 0;
    // what the path pp_knn would choose right now needs (the choice depends on the options in force: query
    // the size after setting them; pp_knn refuses a workspace that is too small, it never overruns it)
    if (knn_use_tc(B, M, N, c, k)) return knn_tc_workspace_bytes(B, M, N);
    return knn_morton_workspace_bytes(B, M, N);
}

extern "C" int pp_knn(const float *query, const float *points, int B, int M, int N, int c, int k,
                      float *dist, int32_t *idx, void *workspace, size_t workspace_bytes, int device,
                      void *stream) {
    PP_REQUIRE(B >= 0 && M >= 0 && N >= 0 && c >= 1, "knn: bad sizes");
    PP_REQUIRE(k >= 1 && k <= PP_KNN_MAX_K, "knn: k=%d outside [1,%d]", k, PP_KNN_MAX_K);
    if (B == 0 || M == 0) return PP_OK;
    PP_REQUIRE(k <= N, "knn: k=%d exceeds the number of points N=%d", k, N);
    PP_REQUIRE(query && points && dist && idx, "knn: null pointer");
    PP_REQUIRE(B <= 65535, "knn: B=%d too large", B);
    DeviceGuard guard(device);
    PP_CUDA(guard.err);
    cudaStream_t st = (cudaStream_t)stream;
    if (c != 3 || get_option("knn_generic", 0)) {
        dim3 grid(ceil_div(M, 128), B);
        knn_generic_kernel<<<grid, 128, 0, st>>>(query, points, M, N, c, k, dist, idx);
        PP_LAUNCH_CHECK();
        return PP_OK;
    }
    if (knn_use_tc(B, M, N, c, k) && !get_option("knn_smem_lists", 0))
        return knn_tc_launch(query, points, B, M, N, k, dist, idx, workspace, workspace_bytes, st);
    if (knn_use_morton(B, M, N, c, k) && !get_option("knn_smem_lists", 0))
        return knn_morton_launch(query, points, B, M, N, k, dist, idx, workspace, workspace_bytes, st);
    if (k <= 32 && !get_option("knn_smem_lists", 0)) {
        KernelTimer timer("knn", st);
        if (k <= 8) {
            dim3 grid(ceil_div(M, KR_THREADS * 2), B);
            knn_reg_kernel<8, 2><<<grid, KR_THREADS, 0, st>>>(query, points, M, N, k, dist, idx);
        } else if (k <= 16) {
            dim3 grid(ceil_div(M, KR_THREADS * 2), B);
            knn_reg_kernel<16, 2><<<grid, KR_THREADS, 0, st>>>(query, points, M, N, k, dist, idx);
        } else {
            dim3 grid(ceil_div(M, KR_THREADS), B);
            knn_reg_kernel<32, 1><<<grid, KR_THREADS, 0, st>>>(query, points, M, N, k, dist, idx);
        }
        PP_LAUNCH_CHECK();
        return PP_OK;
    }
    const size_t smem = knn_smem_bytes(k);
    PP_CUDA(cudaFuncSetAttribute(knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(M, KNN_LISTS), B);
    KernelTimer timer("knn", st);
    knn_kernel<<<grid, KNN_THREADS, smem, st>>>(query, points, M, N, k, dist, idx);
    PP_LAUNCH_CHECK();
    return PP_OK;
}
