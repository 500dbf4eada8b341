// chamfer.cu -- Chamfer / nndistance forward + backward for sm_100a.
//
// Replaces the reference's NmDistanceKernel x2 / NmDistanceGradKernel x2
// (_ext/nmdistance_cuda.cu:8-49,169-185).  Design:
//
//  * ONE pass over the B*N*M unique point pairs feeds both directions: d(i,j) is
//    bit-identical either way round ((a-b) = -(b-a) exactly, squares drop the sign), so the
//    reference's second launch is redundant arithmetic.
//  * Each thread keeps Q query points of cloud 1 in registers and streams cloud 2 through a
//    shared-memory SoA tile, four reference points per LDS.128 broadcast.  Distances are
//    evaluated two at a time with the packed FADD2/FMUL2/FFMA2 pipe in the reference's exact
//    rounding order; running minima use FMNMX3.
//  * Hot loop tracks minima VALUES only.  Indices are recovered lazily:
//      - row side (dist1/idx1): per query remember the 32-reference granule in which the
//        running min last strictly improved (= lowest granule holding the final min);
//      - column side (dist2/idx2): per reference, a warp-wide REDUX.MIN of the per-lane column
//        minimum is compared with a shared-memory filter of the best value seen so far; only
//        when it may improve does one lane push (value, query-group) with a 64-bit atomicMin.
//    A tiny finalize kernel re-evaluates the <=32 (row) / <=Q (column) candidates of the
//    recorded granule and picks the first exact match -> lowest index on ties, as the reference.
//  * Keys are (float bits << 32 | granule): distances are >= +0 so their bit patterns order
//    like unsigned integers, and atomicMin over the packed key resolves ties to the lower
//    granule for free.
#include "pp_common.cuh"

namespace pp {
namespace {

constexpr int CH_TILE = 128;  // reference points per shared-memory tile
constexpr int CH_GR = 32;     // row-side index granule (references)
constexpr unsigned long long KEY_INIT = 0xffffffffffffffffull;

template <int Q, int THREADS>
__global__ void __launch_bounds__(THREADS)
chamfer_fwd_kernel(const float *__restrict__ xyz1, const float *__restrict__ xyz2, int N, int M,
                   unsigned long long *__restrict__ key1, unsigned long long *__restrict__ key2,
                   int refs_per_block) {
    __shared__ __align__(16) float sX[CH_TILE];
    __shared__ __align__(16) float sY[CH_TILE];
    __shared__ __align__(16) float sZ[CH_TILE];
    __shared__ __align__(16) unsigned sW[CH_TILE];  // filter: best column value seen (bits)

    const int b = blockIdx.z;
    const int group = blockIdx.y * THREADS + threadIdx.x;  // query group of this thread
    const int q0 = group * Q;                               // first query of this thread
    const int ref_begin = blockIdx.x * refs_per_block;
    const int ref_end = min(M, ref_begin + refs_per_block);
    const int lane = threadIdx.x & 31;

    const float *p1 = xyz1 + (size_t)b * N * 3;
    const float *p2 = xyz2 + (size_t)b * M * 3;
    unsigned long long *k1 = key1 + (size_t)b * N;
    unsigned long long *k2 = key2 + (size_t)b * M;

    // Negated query coordinates; queries past N become +inf (distance inf, never a minimum).
    float nqx[Q], nqy[Q], nqz[Q], best[Q], prev[Q];
    int granule[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) {
        const int i = q0 + q;
        float x = PP_INF, y = PP_INF, z = PP_INF;
        if (i < N) {
            x = __ldg(p1 + (size_t)i * 3 + 0);
            y = __ldg(p1 + (size_t)i * 3 + 1);
            z = __ldg(p1 + (size_t)i * 3 + 2);
        }
        nqx[q] = -x;
        nqy[q] = -y;
        nqz[q] = -z;
        best[q] = PP_INF;
        prev[q] = PP_INF;
        granule[q] = ref_begin / CH_GR;
    }

    for (int tile0 = ref_begin; tile0 < ref_end; tile0 += CH_TILE) {
        __syncthreads();  // previous tile fully consumed
        for (int t = threadIdx.x; t < CH_TILE; t += THREADS) {
            const int j = tile0 + t;
            float x = PP_INF, y = PP_INF, z = PP_INF;
            unsigned w = 0u;  // padded reference: filter can never pass
            if (j < ref_end) {
                x = __ldg(p2 + (size_t)j * 3 + 0);
                y = __ldg(p2 + (size_t)j * 3 + 1);
                z = __ldg(p2 + (size_t)j * 3 + 2);
                // upper 32 bits of the global key = best value any block has published so far
                w = (unsigned)(__ldcg(k2 + j) >> 32);
            }
            sX[t] = x;
            sY[t] = y;
            sZ[t] = z;
            sW[t] = w;
        }
        __syncthreads();

#pragma unroll 1
        for (int jj = 0; jj < CH_TILE; jj += 4) {
            const float4 X = *reinterpret_cast<const float4 *>(sX + jj);
            const float4 Y = *reinterpret_cast<const float4 *>(sY + jj);
            const float4 Z = *reinterpret_cast<const float4 *>(sZ + jj);
            const float2 x01 = make_float2(X.x, X.y), x23 = make_float2(X.z, X.w);
            const float2 y01 = make_float2(Y.x, Y.y), y23 = make_float2(Y.z, Y.w);
            const float2 z01 = make_float2(Z.x, Z.y), z23 = make_float2(Z.z, Z.w);
            float c0 = PP_INF, c1 = PP_INF, c2 = PP_INF, c3 = PP_INF;
#pragma unroll
            for (int q = 0; q < Q; q += 2) {
                const float2 a01 = sqdist2_xyz(x01, y01, z01, nqx[q], nqy[q], nqz[q]);
                const float2 a23 = sqdist2_xyz(x23, y23, z23, nqx[q], nqy[q], nqz[q]);
                const float2 b01 = sqdist2_xyz(x01, y01, z01, nqx[q + 1], nqy[q + 1], nqz[q + 1]);
                const float2 b23 = sqdist2_xyz(x23, y23, z23, nqx[q + 1], nqy[q + 1], nqz[q + 1]);
                best[q] = fmin3(fmin3(best[q], a01.x, a01.y), a23.x, a23.y);
                best[q + 1] = fmin3(fmin3(best[q + 1], b01.x, b01.y), b23.x, b23.y);
                c0 = fmin3(c0, a01.x, b01.x);
                c1 = fmin3(c1, a01.y, b01.y);
                c2 = fmin3(c2, a23.x, b23.x);
                c3 = fmin3(c3, a23.y, b23.y);
            }
            // ---- column side: warp minimum per reference, filtered publish ----
            const uint4 W = *reinterpret_cast<const uint4 *>(sW + jj);
            const unsigned m0 = __reduce_min_sync(FULL_MASK, __float_as_uint(c0));
            const unsigned m1 = __reduce_min_sync(FULL_MASK, __float_as_uint(c1));
            const unsigned m2 = __reduce_min_sync(FULL_MASK, __float_as_uint(c2));
            const unsigned m3 = __reduce_min_sync(FULL_MASK, __float_as_uint(c3));
            if ((m0 <= W.x) | (m1 <= W.y) | (m2 <= W.z) | (m3 <= W.w)) {
                const unsigned mm[4] = {m0, m1, m2, m3};
                const unsigned ww[4] = {W.x, W.y, W.z, W.w};
                const float cc[4] = {c0, c1, c2, c3};
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    if (mm[r] <= ww[r]) {
                        // lowest lane holding the minimum == lowest query group in this warp
                        const unsigned hit = __ballot_sync(FULL_MASK, __float_as_uint(cc[r]) == mm[r]);
                        if (lane == __ffs(hit) - 1) {
                            atomicMin(k2 + tile0 + jj + r,
                                      ((unsigned long long)mm[r] << 32) | (unsigned)group);
                            atomicMin(sW + jj + r, mm[r]);
                        }
                    }
                }
            }
            // ---- row side: remember the granule of the last strict improvement ----
            if ((jj & (CH_GR - 1)) == CH_GR - 4) {
                const int g = (tile0 + jj) / CH_GR;
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    if (best[q] < prev[q]) {
                        prev[q] = best[q];
                        granule[q] = g;
                    }
                }
            }
        }
    }

#pragma unroll
    for (int q = 0; q < Q; q++) {
        const int i = q0 + q;
        if (i < N)
            atomicMin(k1 + i, ((unsigned long long)__float_as_uint(best[q]) << 32) |
                                  (unsigned)granule[q]);
    }
}

// Resolve (value, granule) keys into (dist, idx): re-evaluate the candidates of the granule in
// ascending index order and take the first whose distance equals the minimum bit for bit.
template <int Q>
__global__ void __launch_bounds__(256)
chamfer_finalize_kernel(const float *__restrict__ xyz1, const float *__restrict__ xyz2, int B, int N,
                        int M, const unsigned long long *__restrict__ key1,
                        const unsigned long long *__restrict__ key2, float *__restrict__ dist1,
                        float *__restrict__ dist2, int *__restrict__ idx1, int *__restrict__ idx2,
                        float *__restrict__ sums) {
    const long long total1 = (long long)B * N, total = total1 + (long long)B * M;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float s1 = 0.f, s2 = 0.f;
    if (t < total1) {
        const int b = (int)(t / N), i = (int)(t % N);
        const unsigned long long key = key1[t];
        const unsigned want = (unsigned)(key >> 32);
        const int g = (int)(unsigned)key;
        const float *q = xyz1 + ((size_t)b * N + i) * 3;
        const float qx = q[0], qy = q[1], qz = q[2];
        const float *r = xyz2 + (size_t)b * M * 3;
        const int j0 = g * CH_GR, j1 = min(M, j0 + CH_GR);
        int found = j0;
        for (int j = j0; j < j1; j++) {
            const float d = sqdist_xyz(__ldg(r + (size_t)j * 3), __ldg(r + (size_t)j * 3 + 1),
                                       __ldg(r + (size_t)j * 3 + 2), qx, qy, qz);
            if (__float_as_uint(d) == want) {
                found = j;
                break;
            }
        }
        dist1[t] = __uint_as_float(want);
        idx1[t] = found;
        s1 = __uint_as_float(want);
    } else if (t < total) {
        const long long u = t - total1;
        const int b = (int)(u / M), j = (int)(u % M);
        const unsigned long long key = key2[u];
        const unsigned want = (unsigned)(key >> 32);
        const int g = (int)(unsigned)key;
        const float *r = xyz2 + ((size_t)b * M + j) * 3;
        const float rx = r[0], ry = r[1], rz = r[2];
        const float *q = xyz1 + (size_t)b * N * 3;
        const int i0 = g * Q, i1 = min(N, i0 + Q);
        int found = i0;
        for (int i = i0; i < i1; i++) {
            // same operand roles as the hot loop: (cloud-2 point) - (cloud-1 point)
            const float d = sqdist_xyz(rx, ry, rz, __ldg(q + (size_t)i * 3), __ldg(q + (size_t)i * 3 + 1),
                                       __ldg(q + (size_t)i * 3 + 2));
            if (__float_as_uint(d) == want) {
                found = i;
                break;
            }
        }
        dist2[u] = __uint_as_float(want);
        idx2[u] = found;
        s2 = __uint_as_float(want);
    }
    if (sums != nullptr) {
        // fused loss partial sums: warp shuffle -> shared -> one atomicAdd pair per block
        __shared__ float sh1[8], sh2[8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(FULL_MASK, s1, o);
            s2 += __shfl_xor_sync(FULL_MASK, s2, o);
        }
        const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
        if (l == 0) {
            sh1[w] = s1;
            sh2[w] = s2;
        }
        __syncthreads();
        if (w == 0) {
            s1 = l < 8 ? sh1[l] : 0.f;
            s2 = l < 8 ? sh2[l] : 0.f;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                s1 += __shfl_xor_sync(FULL_MASK, s1, o);
                s2 += __shfl_xor_sync(FULL_MASK, s2, o);
            }
            if (l == 0) {
                if (s1 != 0.f) atomicAdd(sums + 0, s1);
                if (s2 != 0.f) atomicAdd(sums + 1, s2);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Generic one-direction kernel: arbitrary point dimension c and the labeled
// variant.  Literal restatement of the reference's 512-chunk structure
// (_ext/nmdistance_cuda.cu:20-45 and :63-114) so that even its quirks (k==0 ||
// d<best inside the label test, per-chunk 1e10 initial value) carry over.
// Secondary path: simple thread-per-query tiling, not the roofline kernel.
// ---------------------------------------------------------------------------
constexpr int GEN_CHUNK = 512;
constexpr int GEN_MAXC = 8;  // chunk staged in shared memory for c <= 8

template <bool LABELED>
__global__ void __launch_bounds__(256)
nmdist_generic_kernel(int n, int c, const float *__restrict__ q, const float *__restrict__ ql, int m,
                      const float *__restrict__ r, const float *__restrict__ rl,
                      float *__restrict__ result, int *__restrict__ result_i) {
    extern __shared__ float buf[];  // GEN_CHUNK * (c + LABELED)
    const int b = blockIdx.y;
    q += (size_t)b * n * c;
    r += (size_t)b * m * c;
    if (LABELED) {
        ql += (size_t)b * n;
        rl += (size_t)b * m;
    }
    result += (size_t)b * n;
    result_i += (size_t)b * n;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = j < n;
    const float l1 = (LABELED && active) ? ql[j] : 0.f;
    float res = 0.f;
    int res_i = 0;
    for (int k2 = 0; k2 < m; k2 += GEN_CHUNK) {
        const int end_k = min(m, k2 + GEN_CHUNK) - k2;
        __syncthreads();
        for (int t = threadIdx.x; t < end_k * c; t += blockDim.x) buf[t] = r[(size_t)k2 * c + t];
        if (LABELED)
            for (int t = threadIdx.x; t < end_k; t += blockDim.x) buf[GEN_CHUNK * c + t] = rl[k2 + t];
        __syncthreads();
        if (active) {
            int best_i = LABELED ? -1 : 0;
            float best = LABELED ? 1e10f : 0.f;
            for (int k = 0; k < end_k; k++) {
                if (LABELED && !(l1 == buf[GEN_CHUNK * c + k])) continue;
                float d = 0.f;
                for (int cc = 0; cc < c; cc++) {
                    const float tmp = __fsub_rn(buf[k * c + cc], q[(size_t)j * c + cc]);
                    d = __fmaf_rn(tmp, tmp, d);
                }
                if (k == 0 || d < best) {
                    best = d;
                    best_i = k + k2;
                }
            }
            if (k2 == 0 || res > best) {
                res = best;
                res_i = best_i;
            }
        }
    }
    if (active) {
        if (LABELED && res_i < 0) res = 0.f;
        result[j] = res;
        result_i[j] = res_i;
    }
}

// ---------------------------------------------------------------------------
// Backward (_ext/nmdistance_cuda.cu:169-185): g = 2*gd; v = g*(a - b[idx]);
// grad_a[j] += v; grad_b[idx[j]] -= v.  Fused over both sides in two phases so
// that no zero-fill pass is needed and half of the reference's atomics vanish:
//   phase 0: every point STORES its own term (the reference's first add onto 0);
//   phase 1: every point scatters -v onto its nearest neighbour with RED.ADD.
// ---------------------------------------------------------------------------
template <int PHASE>
__global__ void __launch_bounds__(256)
chamfer_bwd_kernel(const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                   const float *__restrict__ gd1, const float *__restrict__ gd2,
                   const int *__restrict__ idx1, const int *__restrict__ idx2, int B, int N, int M,
                   int c, float *__restrict__ g1, float *__restrict__ g2) {
    const long long total1 = (long long)B * N, total = total1 + (long long)B * M;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const float *a, *bb, *gd;
    const int *idx;
    float *ga, *gb;
    long long u;
    int na, nb;
    if (t < total1) {
        u = t; a = xyz1; bb = xyz2; gd = gd1; idx = idx1; ga = g1; gb = g2; na = N; nb = M;
    } else {
        u = t - total1; a = xyz2; bb = xyz1; gd = gd2; idx = idx2; ga = g2; gb = g1; na = M; nb = N;
    }
    const int b = (int)(u / na);
    const int j2 = idx[u];
    const float g = __fmul_rn(gd[u], 2.f);
    const float *pa = a + (size_t)u * c;
    float *pga = ga + (size_t)u * c;
    if (j2 < 0) {  // labeled variant: no neighbour, no gradient (:175)
        if (PHASE == 0)
            for (int cc = 0; cc < c; cc++) pga[cc] = 0.f;
        return;
    }
    const float *pb = bb + ((size_t)b * nb + j2) * c;
    float *pgb = gb + ((size_t)b * nb + j2) * c;
    for (int cc = 0; cc < c; cc++) {
        const float v = __fmul_rn(g, __fsub_rn(pa[cc], pb[cc]));
        if (PHASE == 0)
            pga[cc] = v;
        else
            atomicAdd(pgb + cc, -v);
    }
}

}  // namespace
}  // namespace pp

using namespace pp;

extern "C" size_t pp_chamfer_fwd_workspace_bytes(int B, int N, int M) {
    if (B <= 0 || N < 0 || M < 0) return 0;
    return sizeof(unsigned long long) * ((size_t)B * N + (size_t)B * M);
}

template <int Q, int THREADS>
static int launch_chamfer_fwd(const float *xyz1, const float *xyz2, int B, int N, int M,
                              unsigned long long *key1, unsigned long long *key2, float *dist1,
                              float *dist2, int *idx1, int *idx2, float *sums, cudaStream_t st) {
    const int qtiles = ceil_div(N, Q * THREADS);
    // enough blocks for ~8 per SM, each covering a whole number of tiles
    const long long want_blocks = (long long)NUM_SMS_B200 * get_option("chamfer_blocks_per_sm", 8);
    int splits = (int)ceil_div_ll(want_blocks, (long long)B * qtiles);
    const int max_splits = ceil_div(M, CH_TILE);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    const int refs_per_block = ceil_div(ceil_div(M, splits), CH_TILE) * CH_TILE;
    splits = ceil_div(M, refs_per_block);
    PP_REQUIRE(qtiles <= 65535 && B <= 65535, "chamfer: grid too large (N=%d B=%d)", N, B);
    dim3 grid(splits, qtiles, B);
    chamfer_fwd_kernel<Q, THREADS><<<grid, THREADS, 0, st>>>(xyz1, xyz2, N, M, key1, key2,
                                                           refs_per_block);
    PP_LAUNCH_CHECK();
    const long long total = (long long)B * N + (long long)B * M;
    chamfer_finalize_kernel<Q><<<(unsigned)ceil_div_ll(total, 256), 256, 0, st>>>(
        xyz1, xyz2, B, N, M, key1, key2, dist1, dist2, idx1, idx2, sums);
    PP_LAUNCH_CHECK();
    return PP_OK;
}

static int launch_generic(bool labeled, const float *xyz1, const float *xyz2, const float *label1,
                          const float *label2, int B, int N, int M, int c, float *dist1, float *dist2,
                          int *idx1, int *idx2, cudaStream_t st) {
    const size_t smem = sizeof(float) * GEN_CHUNK * (c + (labeled ? 1 : 0));
    PP_REQUIRE(smem <= 48 * 1024, "chamfer: point dimension c=%d too large for the generic path", c);
    PP_REQUIRE(B <= 65535, "chamfer: B=%d too large", B);
    if (N > 0) {
        dim3 g1(ceil_div(N, 256), B);
        if (labeled)
            nmdist_generic_kernel<true><<<g1, 256, smem, st>>>(N, c, xyz1, label1, M, xyz2, label2, dist1, idx1);
        else
            nmdist_generic_kernel<false><<<g1, 256, smem, st>>>(N, c, xyz1, nullptr, M, xyz2, nullptr, dist1, idx1);
        PP_LAUNCH_CHECK();
    }
    if (M > 0) {
        dim3 g2(ceil_div(M, 256), B);
        if (labeled)
            nmdist_generic_kernel<true><<<g2, 256, smem, st>>>(M, c, xyz2, label2, N, xyz1, label1, dist2, idx2);
        else
            nmdist_generic_kernel<false><<<g2, 256, smem, st>>>(M, c, xyz2, nullptr, N, xyz1, nullptr, dist2, idx2);
        PP_LAUNCH_CHECK();
    }
    return PP_OK;
}

extern "C" int pp_chamfer_fwd(const float *xyz1, const float *xyz2, int B, int N, int M, int c,
                              float *dist1, float *dist2, int32_t *idx1, int32_t *idx2, float *sums,
                              void *workspace, size_t workspace_bytes, int device, void *stream) {
    PP_REQUIRE(B >= 0 && N >= 0 && M >= 0 && c >= 1, "chamfer_fwd: bad sizes B=%d N=%d M=%d c=%d", B, N, M, c);
    PP_REQUIRE((long long)B * N * c < (1ll << 31) && (long long)B * M * c < (1ll << 31),
               "chamfer_fwd: B*N*c must fit int32 indexing like the reference");
    if (B == 0 || (N == 0 && M == 0)) return PP_OK;
    PP_REQUIRE(xyz1 && xyz2 && dist1 && dist2 && idx1 && idx2, "chamfer_fwd: null pointer");
    DeviceGuard guard(device);
    PP_CUDA(guard.err);
    cudaStream_t st = (cudaStream_t)stream;
    if (sums) PP_CUDA(cudaMemsetAsync(sums, 0, 2 * sizeof(float), st));
    if (N == 0 || M == 0) {
        // the reference's loops never run and leave the Python-side zero fill in place
        if (N) { PP_CUDA(cudaMemsetAsync(dist1, 0, sizeof(float) * (size_t)B * N, st)); PP_CUDA(cudaMemsetAsync(idx1, 0, sizeof(int) * (size_t)B * N, st)); }
        if (M) { PP_CUDA(cudaMemsetAsync(dist2, 0, sizeof(float) * (size_t)B * M, st)); PP_CUDA(cudaMemsetAsync(idx2, 0, sizeof(int) * (size_t)B * M, st)); }
        return PP_OK;
    }
    if (c != 3 || get_option("chamfer_generic", 0)) {
        PP_REQUIRE(sums == nullptr, "chamfer_fwd: fused sums are only available for c == 3");
        return launch_generic(false, xyz1, xyz2, nullptr, nullptr, B, N, M, c, dist1, dist2, idx1, idx2, st);
    }
    const size_t need = pp_chamfer_fwd_workspace_bytes(B, N, M);
    PP_REQUIRE(workspace != nullptr && ((uintptr_t)workspace & 7) == 0, "chamfer_fwd: workspace null or misaligned");
    if (workspace_bytes < need) {
        set_error("chamfer_fwd: workspace %zu < %zu bytes", workspace_bytes, need);
        return PP_ENOSPC;
    }
    unsigned long long *key1 = (unsigned long long *)workspace;
    unsigned long long *key2 = key1 + (size_t)B * N;
    PP_CUDA(cudaMemsetAsync(workspace, 0xff, need, st));
    // query-tile shape: keep padding waste low for small clouds, more reuse for large ones
    const int variant = get_option("chamfer_variant", 0);
    int pick = variant;
    if (pick == 0) pick = (N <= 4096) ? 1 : 2;
    switch (pick) {
        case 1: return launch_chamfer_fwd<8, 64>(xyz1, xyz2, B, N, M, key1, key2, dist1, dist2, idx1, idx2, sums, st);
        case 2: return launch_chamfer_fwd<8, 128>(xyz1, xyz2, B, N, M, key1, key2, dist1, dist2, idx1, idx2, sums, st);
        case 3: return launch_chamfer_fwd<4, 128>(xyz1, xyz2, B, N, M, key1, key2, dist1, dist2, idx1, idx2, sums, st);
        case 4: return launch_chamfer_fwd<16, 64>(xyz1, xyz2, B, N, M, key1, key2, dist1, dist2, idx1, idx2, sums, st);
        case 5: return launch_chamfer_fwd<16, 128>(xyz1, xyz2, B, N, M, key1, key2, dist1, dist2, idx1, idx2, sums, st);
        default: break;
    }
    set_error("chamfer_fwd: unknown variant %d", variant);
    return PP_EINVAL;
}

extern "C" int pp_chamfer_labeled_fwd(const float *xyz1, const float *xyz2, const float *label1,
                                      const float *label2, int B, int N, int M, int c, float *dist1,
                                      float *dist2, int32_t *idx1, int32_t *idx2, int device,
                                      void *stream) {
    PP_REQUIRE(B >= 0 && N >= 0 && M >= 0 && c >= 1, "chamfer_labeled_fwd: bad sizes");
    if (B == 0 || (N == 0 && M == 0)) return PP_OK;
    PP_REQUIRE(xyz1 && xyz2 && label1 && label2 && dist1 && dist2 && idx1 && idx2, "chamfer_labeled_fwd: null pointer");
    DeviceGuard guard(device);
    PP_CUDA(guard.err);
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0 || M == 0) {
        if (N) { PP_CUDA(cudaMemsetAsync(dist1, 0, sizeof(float) * (size_t)B * N, st)); PP_CUDA(cudaMemsetAsync(idx1, 0, sizeof(int) * (size_t)B * N, st)); }
        if (M) { PP_CUDA(cudaMemsetAsync(dist2, 0, sizeof(float) * (size_t)B * M, st)); PP_CUDA(cudaMemsetAsync(idx2, 0, sizeof(int) * (size_t)B * M, st)); }
        return PP_OK;
    }
    return launch_generic(true, xyz1, xyz2, label1, label2, B, N, M, c, dist1, dist2, idx1, idx2, st);
}

extern "C" int pp_chamfer_bwd(const float *xyz1, const float *xyz2, const float *graddist1,
                              const float *graddist2, const int32_t *idx1, const int32_t *idx2, int B,
                              int N, int M, int c, float *gradxyz1, float *gradxyz2, int device,
                              void *stream) {
    PP_REQUIRE(B >= 0 && N >= 0 && M >= 0 && c >= 1, "chamfer_bwd: bad sizes");
    if (B == 0 || (N == 0 && M == 0)) return PP_OK;
    PP_REQUIRE(xyz1 && xyz2 && graddist1 && graddist2 && idx1 && idx2 && gradxyz1 && gradxyz2, "chamfer_bwd: null pointer");
    DeviceGuard guard(device);
    PP_CUDA(guard.err);
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0 || M == 0) {  // nothing to match against: gradients are zero
        if (N) PP_CUDA(cudaMemsetAsync(gradxyz1, 0, sizeof(float) * (size_t)B * N * c, st));
        if (M) PP_CUDA(cudaMemsetAsync(gradxyz2, 0, sizeof(float) * (size_t)B * M * c, st));
        return PP_OK;
    }
    const long long total = (long long)B * N + (long long)B * M;
    const unsigned blocks = (unsigned)ceil_div_ll(total, 256);
    chamfer_bwd_kernel<0><<<blocks, 256, 0, st>>>(xyz1, xyz2, graddist1, graddist2, idx1, idx2, B, N, M, c, gradxyz1, gradxyz2);
    PP_LAUNCH_CHECK();
    chamfer_bwd_kernel<1><<<blocks, 256, 0, st>>>(xyz1, xyz2, graddist1, graddist2, idx1, idx2, B, N, M, c, gradxyz1, gradxyz2);
    PP_LAUNCH_CHECK();
    return PP_OK;
}
