// knn_morton.cu -- spatially ordered, exactly pruned sweep for group_knn on large clouds.
//
// Why: a brute-force top-k evaluates B*M*N distances (8 flop each) and, on unordered clouds,
// spends even more in its selection path.  Here both clouds are first sorted along a Morton
// (Z-order) curve.  A warp's 32 queries are then spatial neighbours and every 64 consecutive
// points form a compact tile with a small bounding box, so a tile whose box is farther from the
// warp's query box than the largest current k-th distance cannot contribute and is never
// loaded.  On uniform clouds 1.6 % (N = 131072) to 18 % (N = 8192) of the tiles survive.
// The result is bit-identical to the unordered brute-force sweep: distances are evaluated in the
// same rounding order, selection and final order use the total order (distance, ORIGINAL
// index), a tile is only skipped when that is provably safe in floating point, and the
// permutation is undone when rows are written.
//
// Pipeline (all on the caller's stream, scratch in the caller's workspace):
//   bbox (atomic min/max) -> 30-bit Morton keys tagged with the batch index -> cub radix sort
//   -> gather sorted coordinates + original indices -> per-tile bounding boxes -> knn_sweep_kernel.
#include <cub/block/block_radix_sort.cuh>
#include <cub/device/device_radix_sort.cuh>

#include "pp_common.cuh"

namespace pp {
namespace {

constexpr int KM_TILE = 64;

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
    if (u == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            for (int w = 1; w < KM_TILE / 32; w++) {
                lo[c] = fminf(lo[c], red[w][c]);
                hi[c] = fmaxf(hi[c], red[w][3 + c]);
            }
        }
        float4 *o = boxes + ((size_t)b * ntiles + t) * 2;
        o[0] = make_float4(lo[0], lo[1], lo[2], hi[0]);
        o[1] = make_float4(hi[1], hi[2], 0.f, 0.f);
    }
}

// ---------------------------------------------------------------------------------------------
// Whole preparation in ONE launch for clouds of up to 16384 points: a 1024-thread CTA per cloud
// computes the batch element's bounding box, the Morton keys, sorts them in shared memory
// (cub::BlockRadixSort), gathers the sorted coordinates / original indices and reduces the tile
// boxes.  Replaces the 12 dependent launches of the general path (memsets, bbox, keys, cub device
// sort, gather, boxes: ~0.14 ms at B*N = 262144, mostly launch latency) by ~30 us.
// The order of equal keys and the exact box used for the keys only influence speed, never results.
// ---------------------------------------------------------------------------------------------
constexpr int KP_THREADS = 1024;
constexpr int KP_RADIX = 5;      // digits of 5 bits: 19 key bits in 4 passes
constexpr int KP_AXIS_BITS = 6;

template <int ITEMS>
__global__ void __launch_bounds__(KP_THREADS)
km_prepare_small_kernel(const float *__restrict__ points, const float *__restrict__ query, int N, int M, int self,
                        unsigned char *__restrict__ wsP, unsigned char *__restrict__ wsQ, size_t off_keys,
                        size_t off_xyz, size_t off_idx, size_t off_boxes) {
    // 6 bits per axis are plenty for <= 16384 points (262144 cells); bit 18 flags padding
    using Sort = cub::BlockRadixSort<unsigned, KP_THREADS, ITEMS, int, KP_RADIX>;
    extern __shared__ __align__(16) unsigned char kp_smem[];
    typename Sort::TempStorage &temp = *reinterpret_cast<typename Sort::TempStorage *>(kp_smem);
    __shared__ float red[KP_THREADS / 32][6];
    __shared__ float box[6];

    const int b = blockIdx.x;
    const bool is_query = blockIdx.y == 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *mine = (is_query ? query + (size_t)b * M * 3 : points + (size_t)b * N * 3);
    const int n_mine = is_query ? M : N;

    // 1. bounding box of this batch element over both clouds (finite coordinates only)
    float lo[3] = {PP_INF, PP_INF, PP_INF}, hi[3] = {-PP_INF, -PP_INF, -PP_INF};
    for (int pass = 0; pass < (self ? 1 : 2); pass++) {
        const float *src = pass == 0 ? points + (size_t)b * N * 3 : query + (size_t)b * M * 3;
        const int cnt3 = (pass == 0 ? N : M) * 3;
        // thread t reads elements t, t + 1024, ...: 1024 = 1 (mod 3), so the component advances by one
        // each step and three unrolled steps cover x, y, z with compile-time indices
        const int cnt = pass == 0 ? N : M;
        const int c0 = tid % 3;
#pragma unroll 4
        for (int e = tid; e < cnt3; e += 3 * KP_THREADS) {
            float v[3];
#pragma unroll
            for (int u = 0; u < 3; u++) v[u] = (e + u * KP_THREADS < cnt3) ? __ldg(src + e + u * KP_THREADS) : __int_as_float(0x7fc00000);
#pragma unroll
            for (int u = 0; u < 3; u++) {
                if (v[u] == v[u] && fabsf(v[u]) != PP_INF) {  // NaN/inf (and the out-of-range filler) are skipped
#pragma unroll
                    for (int cc = 0; cc < 3; cc++)
                        if (cc == (c0 + u) % 3) {
                            lo[cc] = fminf(lo[cc], v[u]);
                            hi[cc] = fmaxf(hi[cc], v[u]);
                        }
                }
            }
        }
        (void)cnt;
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(FULL_MASK, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(FULL_MASK, hi[c], o));
        }
        if (lane == 0) {
            red[warp][c] = lo[c];
            red[warp][3 + c] = hi[c];
        }
    }
    __syncthreads();
    if (tid < 6) {
        float v = red[0][tid];
        for (int w = 1; w < KP_THREADS / 32; w++) v = tid < 3 ? fminf(v, red[w][tid]) : fmaxf(v, red[w][tid]);
        box[tid] = v;
    }
    __syncthreads();

    // 2. keys (blocked arrangement: thread t owns items t*ITEMS ..)
    unsigned keys[ITEMS];
    int vals[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const int e = tid * ITEMS + i;
        keys[i] = 1u << (3 * KP_AXIS_BITS);  // padding sorts behind every real key
        vals[i] = -1;
        if (e < n_mine) {
            unsigned q[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float ext = box[3 + c] - box[c];
                const float v = __ldg(mine + (size_t)e * 3 + c);
                float t = ext > 0.f ? (v - box[c]) / ext * 1023.f : 0.f;
                t = (t == t) ? fminf(fmaxf(t, 0.f), 1023.f) : 0.f;
                q[c] = (unsigned)t;
            }
            keys[i] = (spread10(q[0]) | (spread10(q[1]) << 1) | (spread10(q[2]) << 2)) >> (3 * (10 - KP_AXIS_BITS));
            vals[i] = e;
        }
    }
    // 3. sort; striped output (item i of thread t = position i*1024 + t) keeps the stores coalesced
    Sort(temp).SortBlockedToStriped(keys, vals, 0, 3 * KP_AXIS_BITS + 1);

    // 4. sorted coordinates, original indices, keys
    unsigned char *ws = is_query ? wsQ : wsP;
    unsigned long long *keys_out = (unsigned long long *)(ws + off_keys) + (size_t)b * n_mine;
    float *sxyz = (float *)(ws + off_xyz) + (size_t)b * n_mine * 3;
    int *sidx = (int *)(ws + off_idx) + (size_t)b * n_mine;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const int pos = i * KP_THREADS + tid;
        if (pos < n_mine) {
            const int src = vals[i];
            keys_out[pos] = ((unsigned long long)b << 32) | keys[i];
            sidx[pos] = src;
#pragma unroll
            for (int c = 0; c < 3; c++) sxyz[(size_t)pos * 3 + c] = __ldg(mine + (size_t)src * 3 + c);
        }
    }
    if (is_query) return;
    // 5. tile boxes from the rows this CTA has just written (visible after the barrier)
    __syncthreads();
    const int ntiles = ceil_div(N, KM_TILE);
    for (int t = warp; t < ntiles; t += KP_THREADS / 32) {
        float tlo[3] = {PP_INF, PP_INF, PP_INF}, thi[3] = {-PP_INF, -PP_INF, -PP_INF};
#pragma unroll
        for (int h = 0; h < KM_TILE / 32; h++) {
            const int pos = t * KM_TILE + h * 32 + lane;
            if (pos < N) {
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const float v = sxyz[(size_t)pos * 3 + c];
                    if (v == v && fabsf(v) != PP_INF) {
                        tlo[c] = fminf(tlo[c], v);
                        thi[c] = fmaxf(thi[c], v);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                tlo[c] = fminf(tlo[c], __shfl_xor_sync(FULL_MASK, tlo[c], o));
                thi[c] = fmaxf(thi[c], __shfl_xor_sync(FULL_MASK, thi[c], o));
            }
        }
        if (lane == 0) {
            float4 *o = (float4 *)(ws + off_boxes) + ((size_t)b * ntiles + t) * 2;
            o[0] = make_float4(tlo[0], tlo[1], tlo[2], thi[0]);
            o[1] = make_float4(thi[1], thi[2], 0.f, 0.f);
        }
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

// ---------------------------------------------------------------------------------------------
// The sweep.  One WARP per CTA owns 32 consecutive queries of the sorted query cloud (one per
// lane) and walks over the 64-point tiles of the sorted point cloud:
//   * pruning -- 32 tiles at a time, lane l tests "box of tile l vs the warp's query box vs the
//     largest k-th distance in the warp"; only tiles that survive are loaded.  Pass 0 takes
//     the tiles touching the query box (where the neighbours are, wherever the curve put them),
//     pass 1 whatever is still not provably too far.  Exact: a skipped tile cannot hold a point
//     any list would accept, and evaluation order never changes the result because selection is
//     by the total order (distance, original index);
//   * hot loop -- 4 points per LDS.128 triple, packed FADD2/FMUL2/FFMA2 distances in the
//     reference rounding order, FMNMX3 + one vote per 4 points; lanes that see a distance <=
//     their current k-th push (distance, original index) into a small per-lane buffer;
//   * selection -- the k-best lists live in SHARED memory as 64-bit keys
//     (distance bits << 32 | original index; unsigned order == the (distance, index) order) and
//     insertion is warp-cooperative: slot s of a list sits in lane s of a half-warp, a candidate
//     is compared against all slots at once, one ballot finds its position and one shuffle
//     shifts the tail.  Cost is proportional to the TOTAL number of candidates of the warp; a
//     per-lane register list pays max-over-lanes x K per drain, 5x more on these sweeps.
// ---------------------------------------------------------------------------------------------
constexpr int KM_CB = 8;  // per-lane candidate buffer depth

constexpr unsigned long long KM_EMPTY = 0x7f8000007fffffffull;  // (+inf, no index)

template <int K>
__global__ void __launch_bounds__(32)
knn_sweep_kernel(const float *__restrict__ sq, const int *__restrict__ sqi, const unsigned long long *__restrict__ qkeys,
                 const float *__restrict__ sp, const int *__restrict__ spi, const unsigned long long *__restrict__ pkeys,
                 const float4 *__restrict__ tileboxes, int prune, int seed, int M, int N, int k,
                 float *__restrict__ dist, int *__restrict__ idx, unsigned long long *__restrict__ visited) {
    constexpr int W = K <= 16 ? 16 : 32;  // lanes per list
    constexpr int LP = K + 1;             // padded row length (keys)
    // one array for the tile (x | y | z planes): the hot loop addresses all three from one base
    __shared__ __align__(16) float sP[3 * KM_TILE];
    float *const sX = sP, *const sY = sP + KM_TILE, *const sZ = sP + 2 * KM_TILE;
    __shared__ int sI[KM_TILE];
    __shared__ unsigned long long sBK[KM_CB][32];  // per-lane candidate keys
    __shared__ unsigned long long sL[32][LP];

    const int b = blockIdx.y;
    const int lane = threadIdx.x;
    const float *qp = sq + (size_t)b * M * 3;
    const float *pp_ = sp + (size_t)b * N * 3;
    const int *pi = spi + (size_t)b * N;
    const int qbase = blockIdx.x * 32;
    const int ntiles = ceil_div(N, KM_TILE);

    // where on the points' curve do these queries sit?  lower_bound of the middle query's key
    // among the sorted point keys of this cloud (32-ary search)
    int t0;
    {
        const int mid = min(M - 1, qbase + 16);
        const unsigned long long want = qkeys[(size_t)b * M + mid];
        const unsigned long long *pk = pkeys + (size_t)b * N;
        int lo = 0, hi = N;  // answer in [lo, hi]
        while (hi - lo > 0) {
            const int span = hi - lo;
            const int step = (span + 31) / 32;
            const int probe = lo + lane * step;
            const bool below = probe < hi && pk[probe] < want;
            const int nb = __popc(__ballot_sync(FULL_MASK, below));  // probes 0..nb-1 are below (keys sorted)
            if (nb == 0) {
                hi = lo;
            } else {
                const int nlo = lo + (nb - 1) * step + 1;
                hi = min(hi, lo + nb * step);
                lo = nlo;
            }
        }
        t0 = min(ntiles - 1, lo / KM_TILE);
    }

    // this lane's query, the warp's query box, the empty lists
    const int qi_ = qbase + lane;
    const bool active = qi_ < M;
    float qx = PP_INF, qy = PP_INF, qz = PP_INF;
    if (active) {
        qx = __ldg(qp + (size_t)qi_ * 3);
        qy = __ldg(qp + (size_t)qi_ * 3 + 1);
        qz = __ldg(qp + (size_t)qi_ * 3 + 2);
    }
    const float nqx = -qx, nqy = -qy, nqz = -qz;
    float clo[3], chi[3];
    {
        const float c3[3] = {qx, qy, qz};
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const bool fin = active && c3[c] == c3[c] && fabsf(c3[c]) != PP_INF;
            float lo = fin ? c3[c] : PP_INF, hi = fin ? c3[c] : -PP_INF;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo = fminf(lo, __shfl_xor_sync(FULL_MASK, lo, o));
                hi = fmaxf(hi, __shfl_xor_sync(FULL_MASK, hi, o));
            }
            clo[c] = lo;
            chi[c] = hi;
        }
    }
    for (int e = lane; e < 32 * LP; e += 32) (&sL[0][0])[e] = KM_EMPTY;
#pragma unroll
    for (int e = 0; e < KM_CB; e++) sBK[e][lane] = ~0ull;
    __syncwarp();

    float tau = PP_INF, tau0 = PP_INF;  // accept d <= tau; tau0 = seed (upper bound of the k-th distance)
    int cnt = 0;          // candidates waiting in this lane's buffer
    int fill = 0;         // entries appended to this lane's list while it is not yet full (unsorted)
    bool sorted = false;  // list complete and in order: from here on candidates go through the buffer
    unsigned long long n_visited = 0;

    // Drain the per-lane buffers into the shared lists.  K <= 16: the two half-warps serve two
    // source lanes at once.
    auto drain = [&](bool final) {
        __syncwarp();
        const int half = W == 16 ? (lane >> 4) : 0;
        const int sl = lane & (W - 1);
        // Lists that just became complete were filled by plain appends (the first k accepted
        // candidates need no ordering to be kept): order them now, one bitonic network per list
        // across its lanes -- about 100 instructions instead of k insertions.
        unsigned unsorted = __ballot_sync(FULL_MASK, !sorted && (fill >= k || (final && fill > 0)));
        if (!sorted && (fill >= k || final)) sorted = true;
        while (unsorted != 0u) {
            const int a = __ffs(unsorted) - 1;
            unsorted &= unsorted - 1u;
            int bsrc = -1;
            if (W == 16 && unsorted != 0u) {
                bsrc = __ffs(unsorted) - 1;
                unsorted &= unsorted - 1u;
            }
            const int src = half == 0 ? a : (bsrc < 0 ? a : bsrc);
            const bool mine = sl < k && (half == 0 || bsrc >= 0);  // an idle half touches no list
            unsigned long long my = mine ? sL[src][sl] : ~0ull;    // ~0 sorts behind every real key
#pragma unroll
            for (int size = 2; size <= W; size <<= 1) {
#pragma unroll
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    const unsigned olo = __shfl_xor_sync(FULL_MASK, (unsigned)my, stride);
                    const unsigned ohi = __shfl_xor_sync(FULL_MASK, (unsigned)(my >> 32), stride);
                    const unsigned long long other = ((unsigned long long)ohi << 32) | olo;
                    const bool take_min = ((sl & size) == 0) == ((sl & stride) == 0);
                    my = (other < my) == take_min ? other : my;
                }
            }
            if (mine) sL[src][sl] = my;
        }
        __syncwarp();
        unsigned have = __ballot_sync(FULL_MASK, cnt > 0);
        while (have != 0u) {
            const int a = __ffs(have) - 1;
            have &= have - 1u;
            int bsrc = -1;
            if (W == 16 && have != 0u) {
                bsrc = __ffs(have) - 1;
                have &= have - 1u;
            }
            const int na = __shfl_sync(FULL_MASK, cnt, a);
            const int nb = __shfl_sync(FULL_MASK, cnt, bsrc < 0 ? 0 : bsrc);
            const int src = half == 0 ? a : (bsrc < 0 ? a : bsrc);
            const int n = half == 0 ? na : (bsrc < 0 ? 0 : nb);
            const int steps = max(na, bsrc < 0 ? 0 : nb);
            const bool slot = sl < k && (half == 0 || bsrc >= 0);  // an idle half touches no list
            unsigned long long my = slot ? sL[src][sl] : 0ull;     // 0 never compares greater: inert lanes
            const unsigned long long *cand = &sBK[0][src];
            for (int e = 0; e < steps; e++) {
                // slots past a lane's count hold ~0 (greater than every entry: nothing to insert)
                const unsigned long long key = cand[e * 32];
                const bool lt = key < my;  // true exactly for the slots from the insertion point on
                if (!__any_sync(FULL_MASK, lt)) continue;  // warp-uniform: both candidates already beaten
                const unsigned plo = __shfl_up_sync(FULL_MASK, (unsigned)my, 1, W);
                const unsigned phi = __shfl_up_sync(FULL_MASK, (unsigned)(my >> 32), 1, W);
                // every slot behind the insertion point takes its predecessor, the insertion point
                // itself (predecessor <= key, or no predecessor) takes the candidate: max(key, prev)
                const unsigned long long prev = sl == 0 ? 0ull : (((unsigned long long)phi << 32) | plo);
                if (lt) my = key > prev ? key : prev;
            }
            if (slot && n > 0) sL[src][sl] = my;
        }
        __syncwarp();
        for (int e = 0; e < cnt; e++) sBK[e][lane] = ~0ull;  // consumed: back to "nothing here"
        cnt = 0;
        if (sorted) tau = fminf(tau0, __uint_as_float((unsigned)(sL[lane][k - 1] >> 32)));
    };

    auto load_tile = [&](int t, bool with_index) {
#pragma unroll
        for (int u = lane; u < KM_TILE; u += 32) {
            const int j = t * KM_TILE + u;
            float x = PP_INF, y = PP_INF, z = PP_INF;  // padding: d = inf
            int oi = 0x7fffffff;
            if (j < N) {
                x = __ldg(pp_ + (size_t)j * 3);
                y = __ldg(pp_ + (size_t)j * 3 + 1);
                z = __ldg(pp_ + (size_t)j * 3 + 2);
                if (with_index) oi = __ldg(pi + j);
            }
            sX[u] = x; sY[u] = y; sZ[u] = z; sI[u] = oi;
        }
    };

    // ---- threshold seed: a cheap UPPER BOUND tau0 on every query's k-th distance, taken from the
    // home tile before the real sweep.  The tile is cut into G = K/2 interleaved groups (point u
    // -> group u mod G: every group is a uniform sample of the tile); per group the two smallest
    // distances are tracked with three FMNMX per pair; the largest "second smallest" over the
    // groups has 2G >= k distinct points at or below it.  The sweep then starts with the filter
    // d <= tau0 instead of d <= inf, which removes most of the warm-up candidates.
    if (seed) {
        load_tile(t0, false);
        __syncwarp();
        constexpr int G = K / 2;
        float m1[G], m2[G];
#pragma unroll
        for (int g = 0; g < G; g++) m1[g] = m2[g] = PP_INF;
#pragma unroll 1
        for (int jj0 = 0; jj0 < KM_TILE; jj0 += G) {
#pragma unroll
            for (int v = 0; v < G / 4; v++) {
                const int jj = jj0 + 4 * v;
                const float4 X = *reinterpret_cast<const float4 *>(sX + jj);
                const float4 Y = *reinterpret_cast<const float4 *>(sY + jj);
                const float4 Z = *reinterpret_cast<const float4 *>(sZ + jj);
                const float2 a2 = sqdist2_xyz(make_float2(X.x, X.y), make_float2(Y.x, Y.y), make_float2(Z.x, Z.y), nqx, nqy, nqz);
                const float2 c2 = sqdist2_xyz(make_float2(X.z, X.w), make_float2(Y.z, Y.w), make_float2(Z.z, Z.w), nqx, nqy, nqz);
                const float dd[4] = {a2.x, a2.y, c2.x, c2.y};
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    m2[4 * v + r] = fminf(m2[4 * v + r], fmaxf(m1[4 * v + r], dd[r]));  // second smallest
                    m1[4 * v + r] = fminf(m1[4 * v + r], dd[r]);                        // smallest
                }
            }
        }
        float est = 0.f;
#pragma unroll
        for (int g = 0; g < G; g++) est = fmaxf(est, m2[g]);
        tau0 = est;  // +inf when the home tile is too short: the seed is then simply unused
        tau = est;
        __syncwarp();
    }

    const float4 *boxes = tileboxes + (size_t)b * ntiles * 2;
    // largest k-th distance over the warp's live queries (non-negative floats order as ints)
    auto warp_taumax = [&]() -> float {
        return __int_as_float(__reduce_max_sync(FULL_MASK, __float_as_int(active ? tau : 0.f)));
    };
    // outward along the curve: t0, t0+1, t0-1, t0+2, ... (wrapping)
    auto tile_of = [&](int s) -> int {
        int t = (s & 1) ? t0 + (s + 1) / 2 : t0 - s / 2;  // in (-ntiles, 2 * ntiles): no division needed
        if (t >= ntiles) t -= ntiles;
        return t < 0 ? t + ntiles : t;
    };

    for (int pass = 0; pass < (prune ? 2 : 1); pass++) {
        for (int s0 = 0; s0 < ntiles; s0 += 32) {
            // Thresholds only shrink, so a tile that is provably too far now stays too far;
            // survivors are re-tested with the fresh threshold right before they are loaded
            // (their gap is fetched from the lane that computed it).
            unsigned todo = 0xffffffffu;
            float gap = 0.f;
            if (prune) {
                const float taumax = warp_taumax();
                const int s = s0 + lane;
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
                todo &= todo - 1u;
                const int t = tile_of(s0 + bit);
                if (prune) {
                    // Second, sharper test with the thresholds as they are NOW: every lane measures
                    // the gap between ITS query and the tile's box against ITS k-th distance; the
                    // tile is loaded only if some lane cannot rule it out.  (The lane-parallel test
                    // above used the warp's whole query box and its largest threshold.)
                    const float4 b0 = __ldg(boxes + t * 2), b1 = __ldg(boxes + t * 2 + 1);
                    const float plo[3] = {qx, qy, qz};
                    const bool need = active && !km_can_skip(km_box_gap2(plo, plo, b0, b1), tau);
                    if (!__any_sync(FULL_MASK, need)) continue;
                }
                n_visited++;
                __syncwarp();
                load_tile(t, true);
                __syncwarp();
#pragma unroll 1
                for (int jj = 0; jj < KM_TILE; jj += 8) {
                    // 8 points per trip: two LDS.128 triples, four packed distance pairs, ONE vote
                    float2 d[4];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const float4 X = *reinterpret_cast<const float4 *>(sX + jj + 4 * h);
                        const float4 Y = *reinterpret_cast<const float4 *>(sY + jj + 4 * h);
                        const float4 Z = *reinterpret_cast<const float4 *>(sZ + jj + 4 * h);
                        d[2 * h] = sqdist2_xyz(make_float2(X.x, X.y), make_float2(Y.x, Y.y), make_float2(Z.x, Z.y), nqx, nqy, nqz);
                        d[2 * h + 1] = sqdist2_xyz(make_float2(X.z, X.w), make_float2(Y.z, Y.w), make_float2(Z.z, Z.w), nqx, nqy, nqz);
                    }
                    // '<=': an equal distance with a lower original index still has to get in
                    const float lo8 = fmin3(fmin3(d[0].x, d[0].y, d[1].x), fmin3(d[1].y, d[2].x, d[2].y), fminf(d[3].x, d[3].y));
                    if (__any_sync(FULL_MASK, lo8 <= tau)) {
                        const float dd[8] = {d[0].x, d[0].y, d[1].x, d[1].y, d[2].x, d[2].y, d[3].x, d[3].y};
#pragma unroll
                        for (int h = 0; h < 2; h++) {
#pragma unroll
                            for (int r = 0; r < 4; r++) {
                                if (dd[4 * h + r] <= tau && dd[4 * h + r] < PP_INF) {
                                    if (fill < k) {  // list not full yet: append, order comes later
                                        sL[lane][fill] = ((unsigned long long)__float_as_uint(dd[4 * h + r]) << 32) |
                                                         (unsigned)sI[jj + 4 * h + r];
                                        fill++;
                                    } else {
                                        sBK[cnt][lane] = ((unsigned long long)__float_as_uint(dd[4 * h + r]) << 32) |
                                                         (unsigned)sI[jj + 4 * h + r];
                                        cnt++;
                                    }
                                }
                            }
                            // a lane can add 4 per half-trip: drain while 4 more still fit everywhere,
                            // and as soon as a list is complete (its k-th distance becomes the filter)
                            if (__any_sync(FULL_MASK, cnt > KM_CB - 4 || (fill >= k && !sorted))) drain(false);
                        }
                    }
                }
            }
        }
    }
    drain(true);
    if (visited != nullptr && lane == 0) atomicAdd(visited, n_visited);

    if (active) {
        const int orig = __ldg(sqi + (size_t)b * M + qi_);
        float *od = dist + ((size_t)b * M + orig) * k;
        int *oi = idx + ((size_t)b * M + orig) * k;
        for (int s = 0; s < k; s++) {
            const unsigned long long key = sL[lane][s];
            const int j = (int)(unsigned)key;
            od[s] = __uint_as_float((unsigned)(key >> 32));
            oi[s] = j == 0x7fffffff ? -1 : j;
        }
    }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct KmLayout {
    size_t bbox, keys_in, keys_out, vals_in, vals_out, sorted_xyz, sorted_idx, boxes, cub_temp, cub_bytes, total;
};

// One set of buffers sized for max(nq, np) elements is laid out twice (points, then queries).
KmLayout km_layout(size_t n, size_t clouds) {
    KmLayout L;
    size_t off = 0;
    L.bbox = off; off = align_up(off + 6 * sizeof(int) + 64, 256);  // + a 64-bit visit counter at byte 32
    L.keys_in = off; off = align_up(off + n * 8, 256);
    L.keys_out = off; off = align_up(off + n * 8, 256);
    L.vals_in = off; off = align_up(off + n * 4, 256);
    L.vals_out = off; off = align_up(off + n * 4, 256);
    L.sorted_xyz = off; off = align_up(off + n * 12, 256);
    L.sorted_idx = off; off = align_up(off + n * 4, 256);
    // one 32-byte box per tile: every cloud has ceil(points / KM_TILE) <= points / KM_TILE + 1 tiles
    L.boxes = off; off = align_up(off + (n / KM_TILE + clouds + 2) * 32, 256);
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
    return 2 * km_layout(n, (size_t)B).total + 256;
}

// Sorts both clouds along the Morton curve (sorted coordinates, original indices, keys, 64-point tile
// boxes of the points) into the caller's workspace; shared by the ordered sweep below and by the
// tensor-core path (knn_tc.cu).
int knn_morton_prepare(const float *query, const float *points, int B, int M, int N, void *workspace,
                       size_t workspace_bytes, cudaStream_t st, KmSorted *out) {
    const size_t n = (size_t)B * (size_t)(M > N ? M : N);
    const KmLayout L = km_layout(n, (size_t)B);
    unsigned char *ws = (unsigned char *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    if (workspace == nullptr || (size_t)(ws - (unsigned char *)workspace) + 2 * L.total > workspace_bytes) {
        set_error("knn: workspace %zu < %zu bytes", workspace_bytes, 2 * L.total + 256);
        return PP_ENOSPC;
    }
    unsigned char *wsP = ws, *wsQ = ws + L.total;
    const bool self = (query == points && M == N);
    const int big = M > N ? M : N;
    if (big <= 16 * KP_THREADS && get_option("knn_fused_prep", 1)) {
        // one launch: a CTA per cloud does bbox, keys, sort, gather and tile boxes
        if (self) wsQ = wsP;
        dim3 grid(B, self ? 1 : 2);
#define KP_LAUNCH(ITEMS)                                                                                       \
    do {                                                                                                       \
        auto kern = km_prepare_small_kernel<ITEMS>;                                                            \
        const size_t smem = sizeof(cub::BlockRadixSort<unsigned, KP_THREADS, ITEMS, int, KP_RADIX>::TempStorage);        \
        PP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
        kern<<<grid, KP_THREADS, smem, st>>>(points, query, N, M, self ? 1 : 0, wsP, wsQ, L.keys_out, L.sorted_xyz, \
                                             L.sorted_idx, L.boxes);                                           \
    } while (0)
        if (big <= 4 * KP_THREADS) KP_LAUNCH(4);
        else if (big <= 8 * KP_THREADS) KP_LAUNCH(8);
        else KP_LAUNCH(16);
#undef KP_LAUNCH
        PP_LAUNCH_CHECK();
    } else {
        int *bbox = (int *)(wsP + L.bbox);
        // bbox over both clouds: min <- big positive ints (0x7f7f7f7f), max <- big negative (0x80808080)
        PP_CUDA(cudaMemsetAsync(bbox, 0x7f, 3 * sizeof(int), st));
        PP_CUDA(cudaMemsetAsync(bbox + 3, 0x80, 3 * sizeof(int), st));
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
    }
    out->sp = (const float *)(wsP + L.sorted_xyz); out->sq = (const float *)(wsQ + L.sorted_xyz);
    out->spi = (const int *)(wsP + L.sorted_idx); out->sqi = (const int *)(wsQ + L.sorted_idx);
    out->pk = (const unsigned long long *)(wsP + L.keys_out); out->qk = (const unsigned long long *)(wsQ + L.keys_out);
    out->boxes = (const float4 *)(wsP + L.boxes);
    out->counter = (unsigned long long *)(wsP + L.bbox + 32);
    return PP_OK;
}

// Returns PP_OK after launching everything, or a negative/positive error.
int knn_morton_launch(const float *query, const float *points, int B, int M, int N, int k, float *dist, int *idx,
                      void *workspace, size_t workspace_bytes, cudaStream_t st) {
    KmSorted S;
    const int rc = knn_morton_prepare(query, points, B, M, N, workspace, workspace_bytes, st, &S);
    if (rc != PP_OK) return rc;
    const float *sp = S.sp, *sq = S.sq;
    const int *spi = S.spi, *sqi = S.sqi;
    const unsigned long long *pk = S.pk, *qk = S.qk;
    const float4 *boxes = S.boxes;
    const int prune = get_option("knn_prune", 1);
    unsigned long long *visited = nullptr;
    if (get_option("knn_stats", 0)) {
        visited = S.counter;
        PP_CUDA(cudaMemsetAsync(visited, 0, sizeof(unsigned long long), st));
    }
    {
        KernelTimer timer("knn", st);
        const int seed = get_option("knn_estimate", 1);
        dim3 grid(ceil_div(M, 32), B);
        if (k <= 8)
            knn_sweep_kernel<8><<<grid, 32, 0, st>>>(sq, sqi, qk, sp, spi, pk, boxes, prune, seed, M, N, k, dist, idx, visited);
        else if (k <= 16)
            knn_sweep_kernel<16><<<grid, 32, 0, st>>>(sq, sqi, qk, sp, spi, pk, boxes, prune, seed, M, N, k, dist, idx, visited);
        else
            knn_sweep_kernel<32><<<grid, 32, 0, st>>>(sq, sqi, qk, sp, spi, pk, boxes, prune, seed, M, N, k, dist, idx, visited);
        PP_LAUNCH_CHECK();
    }
    if (visited != nullptr) {  // diagnostics only: synchronises the stream
        unsigned long long v = 0;
        PP_CUDA(cudaMemcpyAsync(&v, visited, sizeof(v), cudaMemcpyDeviceToHost, st));
        PP_CUDA(cudaStreamSynchronize(st));
        g_knn_tiles_visited = (double)v;
        g_knn_tiles_total = (double)B * ceil_div(M, 32) * ceil_div(N, KM_TILE);
    }
    return PP_OK;
}

}  // namespace pp

extern "C" int pp_knn_stats(double *tiles_visited, double *tiles_total) {
    if (tiles_visited) *tiles_visited = pp::g_knn_tiles_visited;
    if (tiles_total) *tiles_total = pp::g_knn_tiles_total;
    return PP_OK;
}
