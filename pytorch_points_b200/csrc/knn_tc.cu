// knn_tc.cu -- group_knn (c == 3, k <= 32) with the dense pairwise-distance contraction on the
// 5th-generation tensor cores: tcgen05.mma decides WHERE a query's neighbours can be, the FP32 pipe
// evaluates only those places with the exact distance chain.  Results are bit-identical to the
// brute-force definition (pp_knn's contract: squared L2 in the Chamfer rounding order, ascending by
// (distance, original index)) -- the same outputs as the ordered sweep of knn_morton.cu.
//
// Serves the snapshot's pytorch3d.ops.knn_points call sites (network/layers.py:52,
// network/geo_operations.py:112,139, network/model_loss.py:120,147,378) and README.md:12's group_knn.
//
// Pipeline (one stream, scratch in the caller's workspace):
//   1. Morton preparation (knn_morton.cu): both clouds sorted along the Z-curve, so that 128
//      consecutive queries / references are spatial neighbours.
//   2. kt_prep_kernel     sorted clouds -> tensor-core operands (3xTF32 split of the centred
//                         coordinates, K = 16: the layout of chamfer_sweep.cu), |q|^2, R^2, a float4
//                         copy of the sorted references (xyz + original index), the bounding box of
//                         every 128-reference block.
//   3. kt_seed_kernel     CTA = 128 queries.  Every query's exact k-th smallest distance inside a
//                         window of 256 references around the tile's place on the curve: an upper
//                         bound TAU0 of its true k-th distance.  Also the box and the largest TAU0 of
//                         each of the tile's four 32-query groups.
//   4. kt_rowpass_kernel  CTA = 128 queries (TMEM lane = query) x the reference blocks that one of the
//                         four groups cannot rule out (box farther than its largest TAU0: exact test).  The
//                         accumulator holds e = |r|^2 - 2 q.r for 128 x 128 pairs; a thread reduces
//                         every 32 of its values to their minimum and flags the 32-reference granule
//                         when that minimum is within the approximation's error bound of TAU0 - |q|^2.
//                         Output: 4 flag bits per (query, visited block).  No selection, no atomics.
//   5. kt_select_kernel   a warp takes 32 queries one after the other: lane = reference of a flagged
//                         granule (exact chain), the references with d <= TAU0 (at least k exist: those
//                         of the window) are compacted into the warp's buffer as keys (distance bits << 32
//                         | original index), and every candidate counts the keys below its own: that
//                         rank is its output slot.
// Why it is exact: the approximation error of e + |q|^2 against the exact chain is below
// EPS = 128 u R^2; a reference with d <= TAU0 therefore has e <= TAU0 - |q|^2 + EPS and
// its granule is flagged (the kernel adds 2.5 EPS); a block is skipped only when its box is provably
// farther than TAU0 from every query of the tile; and the k nearest neighbours all have d <= TAU0.
// Non-finite points are never neighbours (as in the ordered sweep): they are kept out of the operands,
// boxes and R^2, and a non-finite query gets (inf, -1) rows.
#include <atomic>
#include <vector>

#include "pp_common.cuh"
#include "tc_common.cuh"

namespace pp {
namespace {

constexpr int KT_GR = 32;                  // granule (references)
constexpr int KT_STAGES = 8;               // reference blocks in flight per CTA
constexpr int KT_WIN = 256;                // references in the seed window
constexpr int KT_MAX_BLOCKS = 2048;        // reference blocks per cloud (N <= 262144)
constexpr float KT_SLACK_PER_R2 = 320.f * 5.9604644775390625e-8f;  // 2.5 * EPS, EPS = 128 u R^2
constexpr int KT_EPI = 4;
constexpr int KT_THREADS = (KT_EPI + 2) * 32;

struct KtArgs {
    const float *aform;        // (B, qtiles, 8 KB) query operand tiles
    const float *bform;        // (B, rblk, 8 KB) reference operand tiles
    const float *qnorm;        // (B, qtiles * 128) |q|^2 of the centred queries
    const float *tau0;         // (B, qtiles * 128) seed thresholds (-1: no query)
    const float4 *blockbox;    // (B, rblk, 2) boxes of the reference blocks
    const float *tilerec;      // (B, qtiles, 4 warps, 8): box lo/hi and largest TAU0 of each warp's 32 queries
    const unsigned *r2bits;    // (B) bits of R^2
    unsigned short *vis;       // (B, qtiles, rblk) visited blocks of a tile, ascending
    int *viscnt;               // (B, qtiles)
    unsigned *flags;           // (B, qtiles, words, 128): 4 bits per visited block, 8 blocks per word
    int M, N, qtiles, rblk, words;
};

struct KtLayout {
    size_t ctrl, aform, bform, qnorm, ref4, blockbox, tau0, tilerec, vis, viscnt, flags, total;
    int qtiles, rblk, words;
};

KtLayout kt_layout(int B, int M, int N) {
    KtLayout L;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    L.qtiles = ceil_div(M, CS_RB);
    L.rblk = ceil_div(N, CS_RB);
    L.words = ceil_div(L.rblk, 8);
    const size_t qrows = (size_t)B * L.qtiles * CS_RB, rrows = (size_t)B * L.rblk * CS_RB;
    size_t o = 0;
    L.ctrl = o;     o += up(4 * (size_t)B);
    L.aform = o;    o += up(64 * qrows);
    L.bform = o;    o += up(64 * rrows);
    L.qnorm = o;    o += up(4 * qrows);
    L.ref4 = o;     o += up(16 * rrows);
    L.blockbox = o; o += up(32 * (size_t)B * L.rblk);
    L.tau0 = o;     o += up(4 * qrows);
    L.tilerec = o;  o += up(128 * (size_t)B * L.qtiles);
    L.vis = o;      o += up(2 * (size_t)B * L.qtiles * L.rblk);
    L.viscnt = o;   o += up(4 * (size_t)B * L.qtiles);
    L.flags = o;    o += up(4 * (size_t)B * L.qtiles * L.words * CS_RB);
    L.total = o;
    return L;
}

__device__ __forceinline__ bool kt_finite3(float x, float y, float z) {  // false for NaN and +-inf
    return fabsf(x) < PP_INF && fabsf(y) < PP_INF && fabsf(z) < PP_INF;
}

// ---- operands -----------------------------------------------------------------------------------
// grid (max(qtiles, rblk), B, 2): blockIdx.z = 0 the references (B tiles, float4 copy, block boxes),
// 1 the queries (A tiles, norms).  One CTA per 128 rows.
__global__ void __launch_bounds__(CS_RB)
kt_prep_kernel(const float *__restrict__ sq, const float *__restrict__ sp, const int *__restrict__ spi, int M, int N,
               int qtiles, int rblk, float *__restrict__ aform, float *__restrict__ bform, float *__restrict__ qnorm,
               float4 *__restrict__ ref4, float4 *__restrict__ blockbox, unsigned *__restrict__ r2bits) {
    const int b = blockIdx.y, role = blockIdx.z;
    if ((int)blockIdx.x >= (role ? qtiles : rblk)) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __shared__ float s_c[3];
    __shared__ float s_box[CS_RB / 32][6];
    if (threadIdx.x < 32) {
        // centre = mean of 32 + 32 points spread evenly over the two sorted clouds (the curve visits the
        // whole cloud): the same instruction sequence in every CTA of this batch element, hence the same bits
        float sx = 0.f, sy = 0.f, sz = 0.f, cnt = 0.f;
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const int n = s ? M : N;
            const float *p = (s ? sq : sp) + ((size_t)b * n + (size_t)(((long long)lane * n) >> 5)) * 3;
            const float x = p[0], y = p[1], z = p[2];
            if (kt_finite3(x, y, z)) { sx += x; sy += y; sz += z; cnt += 1.f; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sx += __shfl_xor_sync(FULL_MASK, sx, o);
            sy += __shfl_xor_sync(FULL_MASK, sy, o);
            sz += __shfl_xor_sync(FULL_MASK, sz, o);
            cnt += __shfl_xor_sync(FULL_MASK, cnt, o);
        }
        const float inv = cnt > 0.f ? 1.f / cnt : 0.f;
        if (lane == 0) { s_c[0] = sx * inv; s_c[1] = sy * inv; s_c[2] = sz * inv; }
    }
    __syncthreads();
    const float cx = s_c[0], cy = s_c[1], cz = s_c[2];
    const int n = role ? M : N;
    const int i = blockIdx.x * CS_RB + threadIdx.x;
    float px = PP_INF, py = PP_INF, pz = PP_INF;
    if (i < n) {
        const float *p = (role ? sq : sp) + ((size_t)b * n + i) * 3;
        px = p[0]; py = p[1]; pz = p[2];
    }
    const bool ok = i < n && kt_finite3(px, py, pz);
    const float x = ok ? __fsub_rn(px, cx) : 0.f, y = ok ? __fsub_rn(py, cy) : 0.f, z = ok ? __fsub_rn(pz, cz) : 0.f;
    const float nn = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
    constexpr int CH = TC_TILE_BYTES / 16;  // floats between two 16-byte chunks of one operand row
    const size_t tile = ((size_t)b * (role ? qtiles : rblk) + blockIdx.x) * (TC_TILE_BYTES / 4) + (size_t)threadIdx.x * 4;
    if (role) {
        // A row (query role, q = -2 p); a missing / non-finite query is a row of zeros
        const float qx = -2.f * x, qy = -2.f * y, qz = -2.f * z;
        const float qhx = to_tf32(qx), qhy = to_tf32(qy), qhz = to_tf32(qz);
        const float qlx = to_tf32(qx - qhx), qly = to_tf32(qy - qhy), qlz = to_tf32(qz - qhz);
        const float one = ok ? 1.f : 0.f;
        float *at = aform + tile;
        *reinterpret_cast<float4 *>(at) = make_float4(qhx, qhy, qhz, qhx);
        *reinterpret_cast<float4 *>(at + CH) = make_float4(qhy, qhz, qlx, qly);
        *reinterpret_cast<float4 *>(at + 2 * CH) = make_float4(qlz, one, one, one);
        *reinterpret_cast<float4 *>(at + 3 * CH) = make_float4(0.f, 0.f, 0.f, 0.f);
        qnorm[(size_t)b * qtiles * CS_RB + i] = nn;
    } else {
        const float rhx = to_tf32(x), rhy = to_tf32(y), rhz = to_tf32(z);
        const float rlx = to_tf32(x - rhx), rly = to_tf32(y - rhy), rlz = to_tf32(z - rhz);
        // padding / non-finite references carry a huge (finite, TF32-exact) norm: never flagged by a finite threshold
        const float nv = ok ? nn : 1.0e30f;
        const float n0 = to_tf32(nv), n1 = ok ? to_tf32(nv - n0) : 0.f, n2 = ok ? to_tf32(nv - n0 - n1) : 0.f;
        float *bt = bform + tile;
        *reinterpret_cast<float4 *>(bt) = make_float4(rhx, rhy, rhz, rlx);
        *reinterpret_cast<float4 *>(bt + CH) = make_float4(rly, rlz, rhx, rhy);
        *reinterpret_cast<float4 *>(bt + 2 * CH) = make_float4(rhz, n0, n1, n2);
        *reinterpret_cast<float4 *>(bt + 3 * CH) = make_float4(0.f, 0.f, 0.f, 0.f);
        // the exact pass reads this copy: coordinates as they are (a non-finite point only ever gives
        // inf / NaN distances, which nothing accepts), +inf for the padding
        const int oi = i < n ? __ldg(spi + (size_t)b * n + i) : 0x7fffffff;
        ref4[(size_t)b * rblk * CS_RB + i] = make_float4(px, py, pz, __int_as_float(oi));
        float lo[3] = {ok ? px : PP_INF, ok ? py : PP_INF, ok ? pz : PP_INF};
        float hi[3] = {ok ? px : -PP_INF, ok ? py : -PP_INF, ok ? pz : -PP_INF};
#pragma unroll
        for (int c = 0; c < 3; c++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo[c] = fminf(lo[c], __shfl_xor_sync(FULL_MASK, lo[c], o));
                hi[c] = fmaxf(hi[c], __shfl_xor_sync(FULL_MASK, hi[c], o));
            }
            if (lane == 0) { s_box[w][c] = lo[c]; s_box[w][3 + c] = hi[c]; }
        }
    }
    const unsigned rb = __reduce_max_sync(FULL_MASK, __float_as_uint(ok ? nn : 0.f));  // nn >= 0: bits order like values
    if (lane == 0 && rb != 0u) atomicMax(r2bits + b, rb);
    if (!role) {
        __syncthreads();
        if (threadIdx.x == 0) {
            float lo[3], hi[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                lo[c] = s_box[0][c]; hi[c] = s_box[0][3 + c];
#pragma unroll
                for (int v = 1; v < CS_RB / 32; v++) { lo[c] = fminf(lo[c], s_box[v][c]); hi[c] = fmaxf(hi[c], s_box[v][3 + c]); }
            }
            float4 *o = blockbox + ((size_t)b * rblk + blockIdx.x) * 2;
            o[0] = make_float4(lo[0], lo[1], lo[2], hi[0]);
            o[1] = make_float4(hi[1], hi[2], 0.f, 0.f);
        }
    }
}

// ---- seed -----------------------------------------------------------------------------------------
// grid (qtiles, B), 128 threads: thread = query of the sorted query cloud.
template <int K>
__global__ void __launch_bounds__(CS_RB)
kt_seed_kernel(const float *__restrict__ sq, const unsigned long long *__restrict__ qk,
               const unsigned long long *__restrict__ pk, const float4 *__restrict__ ref4, int M, int N, int qtiles,
               int rblk, int k, int self, float *__restrict__ tau0, float *__restrict__ tilerec) {
    __shared__ __align__(16) float sX[KT_WIN], sY[KT_WIN], sZ[KT_WIN];
    __shared__ int s_home;
    const int b = blockIdx.y, tile = blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i_q = tile * CS_RB + (int)threadIdx.x;
    // where on the references' curve does this tile sit?
    if (w == 0) {
        int home = min(N - 1, tile * CS_RB + CS_RB / 2);
        if (!self) {  // lower bound of the middle query's key among the sorted reference keys (32-ary search)
            const unsigned long long want = qk[(size_t)b * M + min(M - 1, tile * CS_RB + CS_RB / 2)];
            const unsigned long long *keys = pk + (size_t)b * N;
            int lo = 0, hi = N;
            while (hi - lo > 0) {
                const int span = hi - lo, step = (span + 31) / 32;
                const int probe = lo + lane * step;
                const bool below = probe < hi && keys[probe] < want;
                const int nb = __popc(__ballot_sync(FULL_MASK, below));
                if (nb == 0) {
                    hi = lo;
                } else {
                    const int nlo = lo + (nb - 1) * step + 1;
                    hi = min(hi, lo + nb * step);
                    lo = nlo;
                }
            }
            home = min(N - 1, lo);
        }
        if (lane == 0) s_home = home;
    }
    __syncthreads();
    const int rows = rblk * CS_RB;
    const int wlen = min(KT_WIN, rows);
    const int w0 = max(0, min(s_home - wlen / 2, rows - wlen)) & ~3;  // (rows and wlen are multiples of 128)
    for (int u = threadIdx.x; u < wlen; u += CS_RB) {
        const float4 r = __ldg(ref4 + (size_t)b * rows + w0 + u);
        sX[u] = r.x; sY[u] = r.y; sZ[u] = r.z;
    }
    float qx = PP_INF, qy = PP_INF, qz = PP_INF;
    if (i_q < M) {
        const float *q = sq + ((size_t)b * M + i_q) * 3;
        qx = q[0]; qy = q[1]; qz = q[2];
    }
    const bool ok = i_q < M && kt_finite3(qx, qy, qz);
    const float nqx = -qx, nqy = -qy, nqz = -qz;
    __syncthreads();
    // the K smallest window distances, ascending, in registers (values only)
    float l[K];
#pragma unroll
    for (int s = 0; s < K; s++) l[s] = PP_INF;
    // The window is walked in groups of 32 references, outward from the group where this warp's own queries
    // sit on the curve (their nearest neighbours come first, so the lists settle early): wlen is a multiple
    // of 128, group order g0, g0+1, g0-1, g0+2, ... wrapped into the window.
    const int ngr = wlen >> 5;
    const int g0 = min(ngr - 1, max(0, (min(s_home, rows - 1) - w0 - CS_RB / 2 + w * 32 + 16) >> 5));
#pragma unroll 1
    for (int jj = 0; jj < wlen; jj += 4) {
        const int step = jj >> 5;
        int g = (step & 1) ? g0 + ((step + 1) >> 1) : g0 - (step >> 1);
        g = g >= ngr ? g - ngr : (g < 0 ? g + ngr : g);
        const int j = (g << 5) + (jj & 31);
        const float4 X = *reinterpret_cast<const float4 *>(sX + j);
        const float4 Y = *reinterpret_cast<const float4 *>(sY + j);
        const float4 Z = *reinterpret_cast<const float4 *>(sZ + j);
        const float2 a2 = sqdist2_xyz(make_float2(X.x, X.y), make_float2(Y.x, Y.y), make_float2(Z.x, Z.y), nqx, nqy, nqz);
        const float2 c2 = sqdist2_xyz(make_float2(X.z, X.w), make_float2(Y.z, Y.w), make_float2(Z.z, Z.w), nqx, nqy, nqz);
        const float dd[4] = {a2.x, a2.y, c2.x, c2.y};
        if (fminf(fminf(dd[0], dd[1]), fminf(dd[2], dd[3])) < l[K - 1]) {
#pragma unroll
            for (int r = 0; r < 4; r++) {
                float v = dd[r];
                if (v < l[K - 1]) {  // (NaN never enters)
#pragma unroll
                    for (int s = 0; s < K; s++) {
                        const float t = fmaxf(l[s], v);
                        l[s] = fminf(l[s], v);
                        v = t;
                    }
                }
            }
        }
    }
    float t0 = PP_INF;
#pragma unroll
    for (int s = 0; s < K; s++)
        if (s == k - 1) t0 = l[s];
    if (!ok) t0 = -1.f;
    tau0[(size_t)b * qtiles * CS_RB + i_q] = t0;
    // box and largest threshold of every warp's 32 queries (Morton neighbours: much tighter than the tile's box)
    float red[7] = {ok ? qx : PP_INF, ok ? qy : PP_INF, ok ? qz : PP_INF, ok ? qx : -PP_INF, ok ? qy : -PP_INF,
                    ok ? qz : -PP_INF, t0};
#pragma unroll
    for (int c = 0; c < 7; c++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float other = __shfl_xor_sync(FULL_MASK, red[c], o);
            red[c] = c < 3 ? fminf(red[c], other) : fmaxf(red[c], other);
        }
    }
    if (lane < 7) {
        float v = red[0];
#pragma unroll
        for (int c = 1; c < 7; c++)
            if (lane == c) v = red[c];
        tilerec[(((size_t)b * qtiles + tile) * (CS_RB / 32) + w) * 8 + lane] = v;
    }
}

// ---- the flagging pass on the tensor cores ------------------------------------------------------------
// Warp roles as in cs_rowpass_tc_kernel: 0-3 epilogue (thread = query = TMEM lane), 4 copy issuer + TMEM
// owner, 5 MMA issuer.  Before the roles start, all six warps test the reference blocks' boxes against the
// tile's query box (exact, see km_can_skip in knn_morton.cu) and compact the survivors into a list.
__global__ void __launch_bounds__(KT_THREADS, 2)
kt_rowpass_kernel(const KtArgs a) {
    extern __shared__ __align__(128) unsigned char kt_dyn_smem[];  // A tile | KT_STAGES B tiles
    unsigned char *sA = kt_dyn_smem;
    unsigned char (*sB)[TC_TILE_BYTES] = reinterpret_cast<unsigned char (*)[TC_TILE_BYTES]>(kt_dyn_smem + TC_TILE_BYTES);
    __shared__ __align__(8) unsigned long long sBar[1 + 2 * KT_STAGES + 4];
    __shared__ unsigned sTmem;
    __shared__ unsigned sSurvive[KT_MAX_BLOCKS / 32];
    __shared__ unsigned short sList[KT_MAX_BLOCKS];
    __shared__ int sCnt;

    const int b = blockIdx.y, tile = blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned bar_a = smem_u32(sBar), bar_bf = smem_u32(sBar + 1), bar_be = smem_u32(sBar + 1 + KT_STAGES);
    const unsigned bar_tf = smem_u32(sBar + 1 + 2 * KT_STAGES), bar_te = smem_u32(sBar + 3 + 2 * KT_STAGES);

    if (threadIdx.x == 0) {
        mbar_init(bar_a, 1);
#pragma unroll
        for (int i = 0; i < KT_STAGES; i++) { mbar_init(bar_bf + 8 * i, 1); mbar_init(bar_be + 8 * i, 1); }
#pragma unroll
        for (int i = 0; i < 2; i++) { mbar_init(bar_tf + 8 * i, 1); mbar_init(bar_te + 8 * i, KT_EPI); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (w == KT_EPI) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&sTmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- which reference blocks can hold a neighbour of one of this tile's queries?
    {
        // a block is visited when one of the four 32-query groups of the tile cannot rule it out
        const float *rec = a.tilerec + ((size_t)b * a.qtiles + tile) * (CS_RB / 32) * 8;
        float qlo[CS_RB / 32][3], qhi[CS_RB / 32][3], taumax[CS_RB / 32];
#pragma unroll
        for (int v = 0; v < CS_RB / 32; v++) {
#pragma unroll
            for (int c = 0; c < 3; c++) { qlo[v][c] = __ldg(rec + v * 8 + c); qhi[v][c] = __ldg(rec + v * 8 + 3 + c); }
            taumax[v] = __ldg(rec + v * 8 + 6);
        }
        const float4 *boxes = a.blockbox + (size_t)b * a.rblk * 2;
        const int ngroups = ceil_div(a.rblk, 32);
        for (int g = w; g < ngroups; g += KT_THREADS / 32) {
            const int blk = g * 32 + lane;
            bool need = false;
            if (blk < a.rblk) {
                const float4 b0 = __ldg(boxes + blk * 2), b1 = __ldg(boxes + blk * 2 + 1);
                const float blo[3] = {b0.x, b0.y, b0.z}, bhi[3] = {b0.w, b1.x, b1.y};
#pragma unroll
                for (int v = 0; v < CS_RB / 32; v++) {
                    float gap2 = 0.f;
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const float gsep = fmaxf(0.f, fmaxf(blo[c] - qhi[v][c], qlo[v][c] - bhi[c]));
                        gap2 = fmaf(gsep, gsep, gap2);
                    }
                    // skipped only when provably too far (margin: the rounded chain can undershoot the real
                    // distance by a few ulp); NaN compares false -> visited
                    need = need || !(gap2 * 0.9999f > taumax[v] && gap2 > 1e-30f);
                }
            }
            const unsigned m = __ballot_sync(FULL_MASK, need);
            if (lane == 0) sSurvive[g] = m;
        }
        __syncthreads();
        if (w == 0) {
            int total = 0;
            for (int g0 = 0; g0 < ngroups; g0 += 32) {
                unsigned m = g0 + lane < ngroups ? sSurvive[g0 + lane] : 0u;
                const int c = __popc(m);
                int incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(FULL_MASK, incl, o);
                    if (lane >= o) incl += t;
                }
                int p = total + incl - c;
                while (m) {
                    const int bit = __ffs(m) - 1;
                    m &= m - 1u;
                    sList[p++] = (unsigned short)((g0 + lane) * 32 + bit);
                }
                total += __shfl_sync(FULL_MASK, incl, 31);
            }
            if (lane == 0) sCnt = total;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = sTmem;
    const int nblk = sCnt;
    {   // the list, for the exact pass
        unsigned short *vis = a.vis + ((size_t)b * a.qtiles + tile) * a.rblk;
        for (int i = threadIdx.x; i < nblk; i += KT_THREADS) vis[i] = sList[i];
        if (threadIdx.x == 0) a.viscnt[(size_t)b * a.qtiles + tile] = nblk;
    }

    if (nblk > 0) {
        if (w == KT_EPI) {
            if (lane == 0) {  // ---- copy issuer
                mbar_expect_tx(bar_a, TC_TILE_BYTES);
                bulk_g2s(smem_u32(sA), a.aform + ((size_t)b * a.qtiles + tile) * (TC_TILE_BYTES / 4), TC_TILE_BYTES, bar_a);
                const float *src = a.bform + (size_t)b * a.rblk * (TC_TILE_BYTES / 4);
                for (int i = 0; i < nblk; i++) {
                    const int st = i % KT_STAGES;
                    if (i >= KT_STAGES) mbar_wait(bar_be + 8 * st, (unsigned)(i / KT_STAGES - 1) & 1u);
                    mbar_expect_tx(bar_bf + 8 * st, TC_TILE_BYTES);
                    bulk_g2s(smem_u32(sB[st]), src + (size_t)sList[i] * (TC_TILE_BYTES / 4), TC_TILE_BYTES, bar_bf + 8 * st);
                }
            }
        } else if (w == KT_EPI + 1) {
            if (lane == 0) {  // ---- MMA issuer (unrolled over the ring: stage, accumulator and parities are constants)
                static_assert(KT_STAGES % 4 == 0, "the slot decides accumulator and its wait parity");
                constexpr unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
                const unsigned long long adesc = tc_smem_desc(smem_u32(sA));
                const unsigned long long bdesc0 = tc_smem_desc(smem_u32(sB[0]));
                mbar_wait(bar_a, 0);
                unsigned ring_parity = 0u;
                for (int i0 = 0; i0 < nblk; i0 += KT_STAGES, ring_parity ^= 1u) {
#pragma unroll
                    for (int u = 0; u < KT_STAGES; u++) {
                        if (i0 + u >= nblk) break;
                        const int acc = u & 1;
                        mbar_wait(bar_bf + 8 * u, ring_parity);
                        if (i0 + u >= 2) mbar_wait(bar_te + 8 * acc, (unsigned)((u >> 1) + 1) & 1u);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const unsigned long long bdesc = bdesc0 + (unsigned long long)(u * (TC_TILE_BYTES >> 4));
                        const unsigned d = tmem + (unsigned)acc * 128u;
                        tc_mma_tf32(d, adesc, bdesc, idesc, 0u);
                        tc_mma_tf32(d, adesc + (2 * 2048 >> 4), bdesc + (2 * 2048 >> 4), idesc, 1u);
                        tc_commit(bar_be + 8 * u);
                        tc_commit(bar_tf + 8 * acc);
                    }
                }
            }
        } else {
            // ---- epilogue warps: thread = query = TMEM lane; four granules of 32 references per block
            const size_t qrow = (size_t)b * a.qtiles * CS_RB + (size_t)tile * CS_RB + threadIdx.x;
            const float r2 = __uint_as_float(__ldg(a.r2bits + b));
            const float t0 = __ldg(a.tau0 + qrow);
            // flag a granule unless its minimum is provably above  TAU0 - |q|^2 + 2.5 EPS
            float thr = __fadd_rn(__fsub_rn(t0, __ldg(a.qnorm + qrow)), r2 * KT_SLACK_PER_R2);
            if (!(r2 < 1.0e30f)) thr = PP_INF;  // the expansion may overflow: flag everything
            if (t0 < 0.f) thr = -PP_INF;        // no query in this lane
            unsigned *fl = a.flags + ((size_t)b * a.qtiles + tile) * a.words * CS_RB + threadIdx.x;
            unsigned word = 0u;
            for (int i = 0; i < nblk; i++) {
                const int acc = i & 1;
                mbar_wait(bar_tf + 8 * acc, (unsigned)(i / 2) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned tbase = tmem + ((unsigned)(w * 32) << 16) + (unsigned)acc * 128u;
                unsigned raw[CS_RB / KT_GR][32];
#pragma unroll
                for (int gi = 0; gi < CS_RB / KT_GR; gi++) tc_ld32_issue(tbase + gi * KT_GR, raw[gi]);
#pragma unroll
                for (int gi = 0; gi < CS_RB / KT_GR; gi++) tc_ld_wait(raw[gi]);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_te + 8 * acc);
                unsigned nib = 0u;
#pragma unroll
                for (int gi = 0; gi < CS_RB / KT_GR; gi++) {
                    float v[32];
#pragma unroll
                    for (int q = 0; q < 32; q++) v[q] = __uint_as_float(raw[gi][q]);
                    float m0 = fmin3(v[0], v[1], v[2]), m1 = fmin3(v[3], v[4], v[5]);
                    float m2 = fmin3(v[6], v[7], v[8]), m3 = fmin3(v[9], v[10], v[11]);
                    m0 = fmin3(m0, v[12], v[13]); m1 = fmin3(m1, v[14], v[15]);
                    m2 = fmin3(m2, v[16], v[17]); m3 = fmin3(m3, v[18], v[19]);
                    m0 = fmin3(m0, v[20], v[21]); m1 = fmin3(m1, v[22], v[23]);
                    m2 = fmin3(m2, v[24], v[25]); m3 = fmin3(m3, v[26], v[27]);
                    m0 = fmin3(m0, v[28], v[29]); m1 = fmin3(m1, v[30], v[31]);
                    const float gm = fminf(fmin3(m0, m1, m2), m3);
                    if (!(gm > thr)) nib |= 1u << gi;
                }
                word |= nib << (4 * (i & 7));
                if ((i & 7) == 7 || i == nblk - 1) {
                    fl[(size_t)(i >> 3) * CS_RB] = word;
                    word = 0u;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == KT_EPI) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

// ---- exact pass ------------------------------------------------------------------------------------------
// grid (qtiles, B), 128 threads.  A warp takes 32 queries of the tile, one after the other with all 32 lanes.
// Everything that steers the loop is warp-uniform (the query's flag words are read by all lanes at once):
//   * up to two flagged granules per trip: lane = reference (independent coalesced 512-byte reads in flight),
//     the references with d <= TAU0 are compacted (ballot + prefix count) into the warp's candidate buffer as
//     keys (distance bits << 32 | original index);
//   * finally every candidate counts the keys below its own -- that rank is its output slot.
// A buffer about to overflow (degenerate data: hundreds of equal distances) is cut back to its k smallest
// keys and the threshold drops to the k-th of them.
// rank += (a < b) on 64-bit keys: two compares and one predicated add
__device__ __forceinline__ void kt_count_less(int &rank, unsigned long long a, unsigned long long b) {
    asm("{\n\t"
        ".reg .pred p;\n\t"
        "setp.lt.u64 p, %1, %2;\n\t"
        "@p add.s32 %0, %0, 1;\n\t"
        "}"
        : "+r"(rank)
        : "l"(a), "l"(b));
}

constexpr int KT_SEL_CAP = 256;   // candidate keys per warp (two buffers)
// flagged granules per trip of the exact pass: a query's ~5 granules are spread over ~4 flag words, so wider trips
// mostly carry idle slots (measured: 4 slots 0.230 ms, 2 slots 0.213, 1 slot 0.230 at B = 32, N = 8192, k = 16)
constexpr int KT_SEL_SLOTS = 2;

__global__ void __launch_bounds__(CS_RB)
kt_select_kernel(const float *__restrict__ sq, const int *__restrict__ sqi, const float4 *__restrict__ ref4,
                 const float *__restrict__ tau0, const unsigned short *__restrict__ vis, const int *__restrict__ viscnt,
                 const unsigned *__restrict__ flags, int M, int qtiles, int rblk, int words, int k,
                 float *__restrict__ dist, int *__restrict__ idx, unsigned long long *__restrict__ stats) {
    extern __shared__ __align__(16) unsigned char kt_sel_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned long long *buf = reinterpret_cast<unsigned long long *>(kt_sel_smem) + w * 2 * KT_SEL_CAP;
    unsigned long long *buf2 = buf + KT_SEL_CAP;
    unsigned short *sVis = reinterpret_cast<unsigned short *>(kt_sel_smem + (size_t)(CS_RB / 32) * 2 * KT_SEL_CAP * 8);
    const int b = blockIdx.y, tile = blockIdx.x;
    const int nvis = __ldg(viscnt + (size_t)b * qtiles + tile);
    {
        const unsigned short *v = vis + ((size_t)b * qtiles + tile) * rblk;
        for (int i = threadIdx.x; i < nvis; i += CS_RB) sVis[i] = v[i];
    }
    __syncthreads();
    const float4 *refs = ref4 + (size_t)b * rblk * CS_RB + lane;  // this lane's reference of granule 0
    const int nwords = (nvis + 7) >> 3;
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned n_gran = 0, n_cand = 0, n_cut = 0;  // statistics (option knn_stats)
    const int q_first = tile * CS_RB + w * 32;
    const int q_count = min(32, M - q_first);  // queries of this warp (<= 0: none)
    const unsigned *fl_warp = flags + ((size_t)b * qtiles + tile) * words * CS_RB + w * 32;
    const float *tau_warp = tau0 + (size_t)b * qtiles * CS_RB + (size_t)tile * CS_RB + w * 32;

    // what a query needs first is requested one query ahead: threshold, coordinates, first flag word, output row
    float nx_tau = -1.f, nx_x = 0.f, nx_y = 0.f, nx_z = 0.f;
    unsigned nx_word = 0u;
    int nx_orig = 0;
    auto request = [&](int qi) {
        if (qi < q_count) {
            const float *q = sq + ((size_t)b * M + q_first + qi) * 3;
            nx_tau = __ldg(tau_warp + qi);
            nx_x = __ldg(q); nx_y = __ldg(q + 1); nx_z = __ldg(q + 2);
            nx_word = nwords > 0 ? __ldg(fl_warp + qi) : 0u;
            nx_orig = __ldg(sqi + (size_t)b * M + q_first + qi);
        }
    };
    request(0);
    for (int qi = 0; qi < q_count; qi++) {
        float tau = nx_tau;
        const float qx = nx_x, qy = nx_y, qz = nx_z;
        const int orig = nx_orig;
        unsigned word = nx_word;
        request(qi + 1);
        int cnt = 0;
        if (!(tau < 0.f)) {
            const unsigned *fw = fl_warp + qi;
            for (int wd = 0; wd < nwords; wd++) {
                fw += CS_RB;
                const unsigned next = wd + 1 < nwords ? __ldg(fw) : 0u;
                const unsigned short *v8 = sVis + wd * 8;
                while (word) {  // warp-uniform
                    // up to KT_SEL_SLOTS flagged granules per trip
                    unsigned off[KT_SEL_SLOTS];
                    bool have[KT_SEL_SLOTS];
#pragma unroll
                    for (int u = 0; u < KT_SEL_SLOTS; u++) {
                        have[u] = word != 0u;
                        const int bit = have[u] ? __ffs(word) - 1 : 0;
                        word &= word - 1u;  // (0 stays 0)
                        off[u] = (unsigned)v8[bit >> 2] * CS_RB + (bit & 3) * KT_GR;
                    }
                    float4 r[KT_SEL_SLOTS];
#pragma unroll
                    for (int u = 0; u < KT_SEL_SLOTS; u++) r[u] = __ldg(refs + off[u]);
                    #pragma unroll
                    for (int u = 0; u < KT_SEL_SLOTS; u++) n_gran += have[u] ? 1u : 0u;
                    if (cnt >= KT_SEL_CAP - 136) {  // room for four granules and the ranking loop's pad (rare)
                        n_cut++;
                        __syncwarp();  // the keys other lanes wrote in the previous trips are read below
                        // every key counts the keys below it; those ranked below k move to their rank
                        for (int i = lane; i < cnt; i += 32) {
                            const unsigned long long mine = buf[i];
                            int rank = 0;
                            for (int j = 0; j < cnt; j++) rank += buf[j] < mine ? 1 : 0;
                            if (rank < k) buf2[rank] = mine;
                        }
                        __syncwarp();
                        unsigned long long *t = buf; buf = buf2; buf2 = t;
                        cnt = min(cnt, k);
                        if (cnt == k) tau = __uint_as_float((unsigned)(buf[k - 1] >> 32));
                    }
#pragma unroll
                    for (int u = 0; u < KT_SEL_SLOTS; u++) {
                        const float d = sqdist_xyz(r[u].x, r[u].y, r[u].z, qx, qy, qz);
                        const bool pass = have[u] && d <= tau && d < PP_INF;
                        const unsigned m = __ballot_sync(FULL_MASK, pass);
                        if (pass) buf[cnt + __popc(m & lt_mask)] = ((unsigned long long)__float_as_uint(d) << 32) | __float_as_uint(r[u].w);
                        cnt += __popc(m);
                    }
                }
                word = next;
            }
        }
        if (lane < 8) buf[cnt + lane] = ~0ull;  // pads the last group of the ranking loop (cnt <= KT_SEL_CAP - 8)
        __syncwarp();
        n_cand += cnt;
        float *od = dist + ((size_t)b * M + orig) * k;
        int *oi = idx + ((size_t)b * M + orig) * k;
        // rank = number of keys below mine (keys are distinct: every reference appears once); two keys per LDS.128
        for (int base = 0; base < cnt; base += 32) {
            const int i = base + lane;
            const unsigned long long mine = i < cnt ? buf[i] : 0ull;
            int rank = 0;
            for (int j = 0; j < cnt; j += 8) {  // eight keys per trip, four LDS.128
#pragma unroll
                for (int u = 0; u < 8; u += 2) {
                    const ulonglong2 kk = *reinterpret_cast<const ulonglong2 *>(buf + j + u);
                    kt_count_less(rank, kk.x, mine);
                    kt_count_less(rank, kk.y, mine);
                }
            }
            if (i < cnt && rank < k) {
                od[rank] = __uint_as_float((unsigned)(mine >> 32));
                oi[rank] = (int)(unsigned)mine;
            }
        }
        for (int s = cnt + lane; s < k; s += 32) {  // fewer than k acceptable references (non-finite data)
            od[s] = PP_INF;
            oi[s] = -1;
        }
        __syncwarp();  // the buffer is reused by the next query
    }
    if (stats != nullptr && lane == 0) {
        atomicAdd(stats + 1, (unsigned long long)n_gran);
        atomicAdd(stats + 2, (unsigned long long)n_cand);
        atomicAdd(stats + 3, (unsigned long long)n_cut);
    }
}

template <typename Kern>
int kt_opt_in(Kern kern, size_t smem, std::atomic<bool> *flags) {
    int dev = 0;
    PP_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !flags[dev].load(std::memory_order_acquire)) {
        PP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) flags[dev].store(true, std::memory_order_release);
    }
    return PP_OK;
}

}  // namespace

bool knn_tc_supported(int B, int M, int N, int k) {
    return k >= 1 && k <= 32 && N >= 1 && M >= 1 && B >= 1 && ceil_div(N, CS_RB) <= KT_MAX_BLOCKS &&
           (long long)B * ceil_div(M, CS_RB) * CS_RB < (1ll << 31) && (long long)B * ceil_div(N, CS_RB) * CS_RB < (1ll << 31);
}

size_t knn_tc_workspace_bytes(int B, int M, int N) {
    return knn_morton_workspace_bytes(B, M, N) + 256 + kt_layout(B, M, N).total;
}

int knn_tc_launch(const float *query, const float *points, int B, int M, int N, int k, float *dist, int *idx,
                  void *workspace, size_t workspace_bytes, cudaStream_t st) {
    const size_t morton_bytes = knn_morton_workspace_bytes(B, M, N);
    const KtLayout L = kt_layout(B, M, N);
    if (workspace == nullptr || workspace_bytes < morton_bytes + 256 + L.total) {
        set_error("knn: workspace %zu < %zu bytes", workspace_bytes, morton_bytes + 256 + L.total);
        return PP_ENOSPC;
    }
    KmSorted S;
    {
        KernelTimer timer("knn_sort", st);
        const int rc = knn_morton_prepare(query, points, B, M, N, workspace, morton_bytes, st, &S);
        if (rc != PP_OK) return rc;
    }
    const bool self = (query == points && M == N);
    unsigned char *ws = (unsigned char *)(((uintptr_t)workspace + morton_bytes + 255) & ~(uintptr_t)255);
    float *aform = (float *)(ws + L.aform), *bform = (float *)(ws + L.bform), *qnorm = (float *)(ws + L.qnorm);
    float4 *ref4 = (float4 *)(ws + L.ref4), *blockbox = (float4 *)(ws + L.blockbox);
    float *tau0 = (float *)(ws + L.tau0), *tilerec = (float *)(ws + L.tilerec);
    unsigned *r2bits = (unsigned *)(ws + L.ctrl);
    unsigned short *vis = (unsigned short *)(ws + L.vis);
    int *viscnt = (int *)(ws + L.viscnt);
    unsigned *flags = (unsigned *)(ws + L.flags);
    {
        KernelTimer timer("knn_prep", st);
        PP_CUDA(cudaMemsetAsync(r2bits, 0, 4 * (size_t)B, st));
        kt_prep_kernel<<<dim3(max(L.qtiles, L.rblk), B, 2), CS_RB, 0, st>>>(S.sq, S.sp, S.spi, M, N, L.qtiles, L.rblk, aform,
                                                                          bform, qnorm, ref4, blockbox, r2bits);
        PP_LAUNCH_CHECK();
    }
    {
        KernelTimer timer("knn_seed", st);
        const dim3 grid(L.qtiles, B);
        // the K smallest window distances live in registers: the narrowest list that holds k (the insertion
        // chain is 2 K FMNMX and it is what the kernel spends its time on)
#define KT_SEED(KK) kt_seed_kernel<KK><<<grid, CS_RB, 0, st>>>(S.sq, S.qk, S.pk, ref4, M, N, L.qtiles, L.rblk, k, self ? 1 : 0, tau0, tilerec)
        if (k <= 4) KT_SEED(4);
        else if (k <= 8) KT_SEED(8);
        else if (k <= 12) KT_SEED(12);
        else if (k <= 16) KT_SEED(16);
        else if (k <= 20) KT_SEED(20);
        else if (k <= 24) KT_SEED(24);
        else if (k <= 28) KT_SEED(28);
        else KT_SEED(32);
#undef KT_SEED
        PP_LAUNCH_CHECK();
    }
    {
        KernelTimer timer("knn", st);
        KtArgs A;
        A.aform = aform; A.bform = bform; A.qnorm = qnorm; A.tau0 = tau0; A.blockbox = blockbox; A.tilerec = tilerec;
        A.r2bits = r2bits; A.vis = vis; A.viscnt = viscnt; A.flags = flags;
        A.M = M; A.N = N; A.qtiles = L.qtiles; A.rblk = L.rblk; A.words = L.words;
        constexpr size_t smem = (size_t)(1 + KT_STAGES) * TC_TILE_BYTES;
        static std::atomic<bool> opted_in[64];
        const int rc = kt_opt_in(kt_rowpass_kernel, smem, opted_in);
        if (rc != PP_OK) return rc;
        kt_rowpass_kernel<<<dim3(L.qtiles, B), KT_THREADS, smem, st>>>(A);
        PP_LAUNCH_CHECK();
    }
    {
        KernelTimer timer("knn_select", st);
        const dim3 grid(L.qtiles, B);
        const size_t vis_bytes = ((size_t)L.rblk * 2 + 15) & ~(size_t)15;
        unsigned long long *stats = nullptr;
        if (get_option("knn_stats", 0)) {
            stats = S.counter;
            PP_CUDA(cudaMemsetAsync(stats, 0, 4 * sizeof(unsigned long long), st));
        }
        const size_t smem = (size_t)(CS_RB / 32) * 2 * KT_SEL_CAP * 8 + vis_bytes;
        kt_select_kernel<<<grid, CS_RB, smem, st>>>(S.sq, S.sqi, ref4, tau0, vis, viscnt, flags, M, L.qtiles, L.rblk, L.words, k,
                                                    dist, idx, stats);
        PP_LAUNCH_CHECK();
    }
    if (get_option("knn_stats", 0)) {  // diagnostics only: synchronises the stream
        std::vector<int> h((size_t)B * L.qtiles);
        PP_CUDA(cudaMemcpyAsync(h.data(), viscnt, h.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
        PP_CUDA(cudaStreamSynchronize(st));
        double v = 0;
        for (int c : h) v += c;
        unsigned long long hs[4];
        PP_CUDA(cudaMemcpy(hs, S.counter, sizeof(hs), cudaMemcpyDeviceToHost));
        const double nq = (double)B * M;
        if (get_option("knn_stats", 0) >= 2)
            fprintf(stderr, "knn_tc stats: per query %.1f flagged granules, %.1f candidates, %.3f cuts\n", hs[1] / nq, hs[2] / nq,
                hs[3] / nq);
        g_knn_tiles_visited = v;  // here in units of 128 x 128 (query tile, reference block) pairs
        g_knn_tiles_total = (double)B * L.qtiles * L.rblk;
    }
    return PP_OK;
}

}  // namespace pp
