// chamfer.cu -- Chamfer / nndistance forward + backward for sm_100a.
//
// Replaces the reference's NmDistanceKernel x2 / NmDistanceGradKernel x2
// (_ext/nmdistance_cuda.cu:8-49,169-185).  Design:
//
//  * ONE pass over the B*N*M unique point pairs feeds both directions: d(i,j) is
//    bit-identical either way round ((a-b) = -(b-a) exactly, squares drop the sign), so the
//    reference's second launch is redundant arithmetic.
//  * A CTA keeps a block of RB cloud-2 points ("references") resident in shared memory (SoA)
//    together with their running column state, and sweeps tiles of cloud-1 points ("queries")
//    past them: each thread holds Q queries in registers and reads four references per
//    LDS.128 broadcast.  Distances are evaluated two at a time on the packed
//    FADD2/FMUL2/FFMA2 pipe in the reference's exact rounding order; minima use FMNMX3.
//  * The hot loop tracks minima VALUES only.  Indices are recovered lazily:
//      - row side (dist1/idx1): per query, the 32-reference granule in which the running
//        min last strictly improved (= lowest granule holding the final min) -> one
//        64-bit RED.MIN of (value, granule) per query and reference block;
//      - column side (dist2/idx2): per reference, a warp-wide REDUX.MIN of the per-lane
//        column minimum is compared with a shared-memory filter (best value seen so far
//        by this CTA, seeded from what earlier CTAs published); only when it may improve
//        does one lane push (value, query-group) into the CTA's shared key array.
//    A small finalize kernel re-evaluates the <=32 (row) / <=Q (column) candidates of the
//    recorded granule and picks the first exact match -> lowest index on ties, like the
//    reference's strict '<' scan.
//  * Keys are (float bits << 32 | granule): distances are >= +0 so their bit patterns order
//    like unsigned integers, and a min over the packed key resolves ties to the lower
//    granule for free.
//  * Query splits ride the slowest grid dimension, so CTAs of split s start after the CTAs
//    of split s-1 have published their keys: the column filter starts tight.
#include "pp_common.cuh"

namespace pp {
namespace {

constexpr int CH_GR = 32;  // row-side index granule (references)
constexpr unsigned long long KEY_INIT = 0xffffffffffffffffull;

// The leader lane publishes (value, group) to the global key with one fire-and-forget
// RED.MIN.64 and tightens the CTA's shared filter.  Predicated PTX: only one lane acts, nobody
// branches.  The filter update is a plain store on purpose: every value ever stored is an
// observed warp minimum, hence >= the final minimum, so a lost or reordered update can only let
// a few extra candidates through -- it can never hide one.
__device__ __forceinline__ void publish_column_min(int lane, int leader, unsigned long long *gkey,
                                                   unsigned long long key, unsigned *filt,
                                                   unsigned value) {
    const unsigned faddr = (unsigned)__cvta_generic_to_shared(filt);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.eq.s32 p, %0, %1;\n\t"
        "@p red.global.min.u64 [%2], %3;\n\t"
        "@p st.shared.u32 [%4], %5;\n\t"
        "}"
        :
        : "r"(lane), "r"(leader), "l"(gkey), "l"(key), "r"(faddr), "r"(value)
        : "memory");
}

// Election-free form: EVERY lane whose column minimum equals the warp minimum publishes its own
// (value, group) key.  Exact for the same reason the election was: RED.MIN on the packed key
// keeps the lowest group among equal values.  Saves the ballot / find-first / lane compare of
// the elected form; ties (more than one publishing lane) are rare.
__device__ __forceinline__ void publish_column_min_all(unsigned mine, unsigned mn, unsigned long long *gkey,
                                                       unsigned long long key, unsigned *filt) {
    const unsigned faddr = (unsigned)__cvta_generic_to_shared(filt);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.eq.u32 p, %0, %1;\n\t"
        "@p red.global.min.u64 [%2], %3;\n\t"
        "@p st.shared.u32 [%4], %1;\n\t"
        "}"
        :
        : "r"(mine), "r"(mn), "l"(gkey), "l"(key), "r"(faddr)
        : "memory");
}

// LABELED (LabeledNmdistanceFunction, _ext/nmdistance_cuda.cu:56-115): only pairs with equal fp32
// labels are candidates -- every other distance is replaced by +inf right after it is computed,
// so both minima see same-label partners only; a point without a partner keeps +inf and the
// finalize kernel turns that into the reference's (dist 0, idx -1).
// STAGEQ: a thread's Q query points are 12*Q contiguous bytes, so loading them straight from
// global memory makes every LDG of a warp touch 24 different 128-byte lines (24 such loads per
// tile: ~1200 cycles of the SM's single L1 wavefront queue, which the other warps' LDS.128 of the
// hot loop wait behind).  Staged, the CTA reads its tile as whole lines into shared memory (one
// pad word per 24 so that the per-thread readback at stride 25 words is conflict free).
template <int Q, int THREADS, int RB, int MINB, bool ROTATE, bool LABELED, bool STAGEQ = false, bool ELECT = true>
__global__ void __launch_bounds__(THREADS, MINB)
chamfer_fwd_kernel(const float *__restrict__ xyz1, const float *__restrict__ xyz2, int N, int M,
                   unsigned long long *__restrict__ key1, unsigned long long *__restrict__ key2,
                   int queries_per_split, const float *__restrict__ label1,
                   const float *__restrict__ label2, float *__restrict__ zero1,
                   float *__restrict__ zero2) {
    __shared__ __align__(16) float sX[RB];
    __shared__ __align__(16) float sY[RB];
    __shared__ __align__(16) float sZ[RB];
    __shared__ __align__(16) unsigned sW[RB];  // filter: best column value seen (bits)
    __shared__ __align__(16) float sLab[LABELED ? RB : 4];
    __shared__ float sQ[STAGEQ ? Q * THREADS * 3 + THREADS : 1];

    pdl_launch_dependents();  // the finalize kernel may take the SM slots this grid's tail frees
    constexpr int TQ = Q * THREADS;
    const int b = blockIdx.y;
    const int ref_begin = blockIdx.x * RB;
    const int q_begin = blockIdx.z * queries_per_split;
    const int q_end = min(N, q_begin + queries_per_split);
    const int lane = threadIdx.x & 31;

    const float *p1 = xyz1 + (size_t)b * N * 3;
    const float *p2 = xyz2 + (size_t)b * M * 3;
    unsigned long long *k1 = key1 + (size_t)b * N;
    unsigned long long *k2 = key2 + (size_t)b * M;

    for (int t = threadIdx.x; t < RB; t += THREADS) {
        const int j = ref_begin + t;
        float x = PP_INF, y = PP_INF, z = PP_INF;  // padding: distance inf, never a minimum
        unsigned w = 0u;                           // padding: the filter can never pass
        if (j < M) {
            x = __ldg(p2 + (size_t)j * 3 + 0);
            y = __ldg(p2 + (size_t)j * 3 + 1);
            z = __ldg(p2 + (size_t)j * 3 + 2);
            // upper half of the global key = best value any finished CTA has published
            w = (unsigned)(__ldcg(k2 + j) >> 32);
        }
        sX[t] = x;
        sY[t] = y;
        sZ[t] = z;
        sW[t] = w;
        // padding carries a NaN label: equal to nothing
        if (LABELED) sLab[t] = j < M ? __ldg(label2 + (size_t)b * M + j) : __int_as_float(0x7fc00000);
    }
    // Fused backward (pp_chamfer_fwd_bwd_uniform): the finalize kernel accumulates the gradients
    // with RED.ADD, so they start from zero -- each (reference block, split 0) CTA clears its
    // slice of gradxyz2, each (reference block 0, split) CTA its slice of gradxyz1.  The finalize
    // kernel starts after this whole grid has completed (griddepcontrol.wait).
    if (zero2 != nullptr && blockIdx.z == 0) {
        const int lim = (min(M, ref_begin + RB) - ref_begin) * 3;
        float *z = zero2 + ((size_t)b * M + ref_begin) * 3;
        for (int t = threadIdx.x; t < lim; t += THREADS) z[t] = 0.f;
    }
    if (zero1 != nullptr && blockIdx.x == 0) {
        const int lim = (q_end - q_begin) * 3;
        float *z = zero1 + ((size_t)b * N + q_begin) * 3;
        for (int t = threadIdx.x; t < lim; t += THREADS) z[t] = 0.f;
    }
    __syncthreads();

    for (int qt = q_begin; qt < q_end; qt += TQ) {
        if (STAGEQ) {
            if (qt != q_begin) __syncthreads();  // the previous tile has been read by everybody
            const int nfl = (min(q_end, qt + TQ) - qt) * 3;
            const float *src = p1 + (size_t)qt * 3;
#pragma unroll  // 3*Q = 24 independent loads in flight: one memory round trip per tile
            for (int f = threadIdx.x; f < TQ * 3; f += THREADS)
                sQ[f + f / (3 * Q)] = f < nfl ? __ldg(src + f) : PP_INF;  // padding: +inf like below
            __syncthreads();
        }
        // thread owns Q consecutive queries; a warp whose queries are all padding sits out
        const int q0 = qt + threadIdx.x * Q;
        // (the shuffle tells the compiler the test is warp-uniform, so the sweep below keeps
        //  its loop state in uniform registers)
        if (__shfl_sync(FULL_MASK, qt + (int)(threadIdx.x & ~31u) * Q, 0) >= q_end) continue;
        const int group = q0 / Q;

        float nqx[Q], nqy[Q], nqz[Q], best[Q], prev[Q];
        float ql[LABELED ? Q : 1];
        int granule[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const int i = q0 + q;
            float x = PP_INF, y = PP_INF, z = PP_INF;
            if (LABELED) ql[q] = __int_as_float(0x7fc00000);
            if (STAGEQ) {
                const float *mine = sQ + threadIdx.x * (3 * Q + 1) + 3 * q;
                x = mine[0]; y = mine[1]; z = mine[2];
                if (LABELED && i < q_end) ql[q] = __ldg(label1 + (size_t)b * N + i);
            } else if (i < q_end) {
                x = __ldg(p1 + (size_t)i * 3 + 0);
                y = __ldg(p1 + (size_t)i * 3 + 1);
                z = __ldg(p1 + (size_t)i * 3 + 2);
                if (LABELED) ql[q] = __ldg(label1 + (size_t)b * N + i);
            }
            nqx[q] = -x;  // negated: rn(r + (-q)) == rn(r - q)
            nqy[q] = -y;
            nqz[q] = -z;
            best[q] = PP_INF;
            prev[q] = PP_INF;
            granule[q] = ref_begin / CH_GR;
        }

        // Warp w starts its sweep over the resident points at offset w*RB/WARPS and wraps: at any
        // moment the warps of a CTA work on DIFFERENT resident points, so each point meets the
        // CTA's warps one after the other and the filter a warp sees already contains what the
        // previous warps found (fewer publishes).  The row side stays exact: when the sweep wraps
        // to index 0 the row minima found so far are published and tracking restarts -- the
        // packed (value, granule) RED.MIN resolves ties towards the lower granule.
        constexpr int WARPS = THREADS / 32;
        // (offsets are whole granules: the row side attributes an improvement to the granule
        //  whose last step it is checked at, so a sweep must start on a granule boundary)
        static_assert(RB % CH_GR == 0, "reference blocks are whole granules");
        const int rot = ROTATE ? __shfl_sync(FULL_MASK, ((int)(threadIdx.x >> 5) * (RB / WARPS)) & ~(CH_GR - 1), 0) : 0;
#pragma unroll 1
        for (int step = 0; step < RB; step += 4) {
            int jj = step + rot;
            if (ROTATE && jj >= RB) {
                jj -= RB;
                if (jj == 0) {  // wrap point (warp-uniform): publish and restart the row tracking
#pragma unroll
                    for (int q = 0; q < Q; q++) {
                        const int i = q0 + q;
                        if (i < q_end)
                            atomicMin(k1 + i, ((unsigned long long)__float_as_uint(best[q]) << 32) |
                                                  (unsigned)granule[q]);
                        best[q] = PP_INF;
                        prev[q] = PP_INF;
                        granule[q] = ref_begin / CH_GR;
                    }
                }
            }
            const float4 X = *reinterpret_cast<const float4 *>(sX + jj);
            const float4 Y = *reinterpret_cast<const float4 *>(sY + jj);
            const float4 Z = *reinterpret_cast<const float4 *>(sZ + jj);
            const float2 x01 = make_float2(X.x, X.y), x23 = make_float2(X.z, X.w);
            const float2 y01 = make_float2(Y.x, Y.y), y23 = make_float2(Y.z, Y.w);
            const float2 z01 = make_float2(Z.x, Z.y), z23 = make_float2(Z.z, Z.w);
            float c0 = PP_INF, c1 = PP_INF, c2 = PP_INF, c3 = PP_INF;
            float4 LB = make_float4(0.f, 0.f, 0.f, 0.f);
            if (LABELED) LB = *reinterpret_cast<const float4 *>(sLab + jj);
#pragma unroll
            for (int q = 0; q < Q; q += 2) {
                float2 a01 = sqdist2_xyz(x01, y01, z01, nqx[q], nqy[q], nqz[q]);
                float2 a23 = sqdist2_xyz(x23, y23, z23, nqx[q], nqy[q], nqz[q]);
                float2 b01 = sqdist2_xyz(x01, y01, z01, nqx[q + 1], nqy[q + 1], nqz[q + 1]);
                float2 b23 = sqdist2_xyz(x23, y23, z23, nqx[q + 1], nqy[q + 1], nqz[q + 1]);
                if (LABELED) {
                    a01.x = LB.x == ql[q] ? a01.x : PP_INF;
                    a01.y = LB.y == ql[q] ? a01.y : PP_INF;
                    a23.x = LB.z == ql[q] ? a23.x : PP_INF;
                    a23.y = LB.w == ql[q] ? a23.y : PP_INF;
                    b01.x = LB.x == ql[q + 1] ? b01.x : PP_INF;
                    b01.y = LB.y == ql[q + 1] ? b01.y : PP_INF;
                    b23.x = LB.z == ql[q + 1] ? b23.x : PP_INF;
                    b23.y = LB.w == ql[q + 1] ? b23.y : PP_INF;
                }
                best[q] = fmin3(fmin3(best[q], a01.x, a01.y), a23.x, a23.y);
                best[q + 1] = fmin3(fmin3(best[q + 1], b01.x, b01.y), b23.x, b23.y);
                c0 = fmin3(c0, a01.x, b01.x);
                c1 = fmin3(c1, a01.y, b01.y);
                c2 = fmin3(c2, a23.x, b23.x);
                c3 = fmin3(c3, a23.y, b23.y);
            }
            // ---- column side: filtered publish of the warp minimum per reference ----
            // Fast path: each lane tests its own column minimum against the filter; some lane
            // passes iff the warp minimum passes.  NOTE: the filter words are updated
            // concurrently by other warps and an LDS.128 is served in several passes, so lanes
            // of ONE load may see different values: every decision is made warp-uniform with
            // a vote before any *_sync primitive.
            const uint4 W = *reinterpret_cast<const uint4 *>(sW + jj);
            const bool p0 = __float_as_uint(c0) <= W.x, p1 = __float_as_uint(c1) <= W.y;
            const bool p2 = __float_as_uint(c2) <= W.z, p3 = __float_as_uint(c3) <= W.w;
            if (__any_sync(FULL_MASK, p0 | p1 | p2 | p3)) {
                const bool pp_[4] = {p0, p1, p2, p3};
                const float cc[4] = {c0, c1, c2, c3};
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    if (__any_sync(FULL_MASK, pp_[r])) {
                        const unsigned mine = __float_as_uint(cc[r]);
                        const unsigned mn = __reduce_min_sync(FULL_MASK, mine);
                        if (ELECT) {
                            // lowest lane holding the minimum == lowest query group in this warp
                            const unsigned hit = __ballot_sync(FULL_MASK, mine == mn);
                            publish_column_min(lane, __ffs(hit) - 1, k2 + ref_begin + jj + r,
                                               ((unsigned long long)mn << 32) | (unsigned)group,
                                               sW + jj + r, mn);
                        } else {
                            publish_column_min_all(mine, mn, k2 + ref_begin + jj + r,
                                                   ((unsigned long long)mn << 32) | (unsigned)group,
                                                   sW + jj + r);
                        }
                    }
                }
            }
            // ---- row side: remember the granule of the last strict improvement ----
            if ((jj & (CH_GR - 1)) == CH_GR - 4) {
                const int g = (ref_begin + jj) / CH_GR;
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    if (best[q] < prev[q]) {
                        prev[q] = best[q];
                        granule[q] = g;
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const int i = q0 + q;
            if (i < q_end)
                atomicMin(k1 + i, ((unsigned long long)__float_as_uint(best[q]) << 32) |
                                      (unsigned)granule[q]);
        }
    }
}

// Resolve (value, granule) keys into (dist, idx): re-evaluate the candidates of the granule in
// ascending index order and take the first whose distance equals the minimum bit for bit.
// Also restores the key workspace to its all-ones "clean" state for the next call.
//   rows   : a warp takes 32 queries; for each, its 32 lanes test the 32 references of the
//            recorded granule at once (coalesced 384-byte read, ballot, find-first-set);
//   columns: a thread re-evaluates the Q queries of the recorded group, fully unrolled.
// FUSE_BWD (pp_chamfer_fwd_bwd_uniform): the backward for a loss that sees dist1/dist2 only through
// their sums runs right here -- the thread that resolves a point's neighbour also forms
// v = 2*gw[side]*(point - neighbour) (_ext/nmdistance_cuda.cu:176-181, same rounding steps as
// chamfer_bwd_kernel) and adds +v to its own gradient and -v to the neighbour's with RED.ADD.F32
// on arrays the forward kernel cleared.  No idx round trip, no extra launches.
template <int Q, bool LABELED, bool FUSE_BWD>
__global__ void __launch_bounds__(256)
chamfer_finalize_kernel(const float *__restrict__ xyz1, const float *__restrict__ xyz2, int B, int N,
                        int M, unsigned long long *__restrict__ key1,
                        unsigned long long *__restrict__ key2, float *__restrict__ dist1,
                        float *__restrict__ dist2, int *__restrict__ idx1, int *__restrict__ idx2,
                        float *__restrict__ sums, int row_blocks, const float *__restrict__ label1,
                        const float *__restrict__ label2, const float *__restrict__ gw,
                        float *__restrict__ g1, float *__restrict__ g2) {
    constexpr unsigned INF_BITS = 0x7f800000u;  // LABELED: no same-label partner -> (dist 0, idx -1)
    const int lane = threadIdx.x & 31;
    float s1 = 0.f, s2 = 0.f;
    pdl_wait();                // keys are complete and visible
    pdl_launch_dependents();   // a following backward kernel may queue up behind us
    if ((int)blockIdx.x < row_blocks) {
        // ---- rows: 256 queries per block, 32 per warp, never straddling a cloud's end badly:
        // t indexes the flattened (B*N) query array; lanes past the end idle.
        const long long total1 = (long long)B * N;
        const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
        unsigned want = 0u;
        int g = 0, b = 0;
        float qx = 0.f, qy = 0.f, qz = 0.f, ql = 0.f;
        const bool valid = t < total1;
        if (valid) {
            const unsigned long long key = key1[t];
            key1[t] = KEY_INIT;
            want = (unsigned)(key >> 32);
            g = (int)(unsigned)key;
            b = (int)(t / N);
            const float *q = xyz1 + (size_t)t * 3;
            qx = q[0]; qy = q[1]; qz = q[2];
            if (LABELED) ql = label1[t];
        }
        int found = g * CH_GR;
#pragma unroll 4
        for (int s = 0; s < 32; s++) {
            const unsigned w_s = __shfl_sync(FULL_MASK, want, s);
            const int g_s = __shfl_sync(FULL_MASK, g, s);
            const int b_s = __shfl_sync(FULL_MASK, b, s);
            const float x_s = __shfl_sync(FULL_MASK, qx, s);
            const float y_s = __shfl_sync(FULL_MASK, qy, s);
            const float z_s = __shfl_sync(FULL_MASK, qz, s);
            const float l_s = LABELED ? __shfl_sync(FULL_MASK, ql, s) : 0.f;
            const int j = g_s * CH_GR + lane;
            bool match = false;
            if (j < M) {
                const float *r = xyz2 + ((size_t)b_s * M + j) * 3;
                const float d = sqdist_xyz(__ldg(r), __ldg(r + 1), __ldg(r + 2), x_s, y_s, z_s);
                match = __float_as_uint(d) == w_s;
                if (LABELED) match = match && __ldg(label2 + (size_t)b_s * M + j) == l_s;
            }
            const unsigned hit = __ballot_sync(FULL_MASK, match);
            if (lane == s && hit != 0u) found = g_s * CH_GR + __ffs(hit) - 1;
        }
        if (valid) {
            if (LABELED && want == INF_BITS) {
                want = 0u;
                found = -1;
            }
            dist1[t] = __uint_as_float(want);
            idx1[t] = found;
            s1 = __uint_as_float(want);
            if (FUSE_BWD && found >= 0) {
                const float g = __fmul_rn(__ldg(gw + 0), 2.f);
                const size_t nb = ((size_t)b * M + found) * 3;
                const float vx = __fmul_rn(g, __fsub_rn(qx, __ldg(xyz2 + nb + 0)));
                const float vy = __fmul_rn(g, __fsub_rn(qy, __ldg(xyz2 + nb + 1)));
                const float vz = __fmul_rn(g, __fsub_rn(qz, __ldg(xyz2 + nb + 2)));
                float *own = g1 + (size_t)t * 3, *other = g2 + nb;
                atomicAdd(own + 0, vx); atomicAdd(own + 1, vy); atomicAdd(own + 2, vz);
                atomicAdd(other + 0, -vx); atomicAdd(other + 1, -vy); atomicAdd(other + 2, -vz);
            }
        }
    } else {
        const long long total2 = (long long)B * M;
        const long long u = (long long)((int)blockIdx.x - row_blocks) * 256 + threadIdx.x;
        if (u < total2) {
            const int b = (int)(u / M);
            const unsigned long long key = key2[u];
            key2[u] = KEY_INIT;
            unsigned want = (unsigned)(key >> 32);
            const int g = (int)(unsigned)key;
            const float *r = xyz2 + (size_t)u * 3;
            const float rx = r[0], ry = r[1], rz = r[2];
            const float rl = LABELED ? label2[u] : 0.f;
            const float *q = xyz1 + (size_t)b * N * 3;
            const int i0 = g * Q;
            int found = i0;
#pragma unroll
            for (int e = Q - 1; e >= 0; e--) {  // descending so the lowest match wins
                const int i = i0 + e;
                if (i < N) {
                    // same operand roles as the hot loop: (cloud-2 point) - (cloud-1 point)
                    const float d = sqdist_xyz(rx, ry, rz, __ldg(q + (size_t)i * 3), __ldg(q + (size_t)i * 3 + 1),
                                               __ldg(q + (size_t)i * 3 + 2));
                    bool same = __float_as_uint(d) == want;
                    if (LABELED) same = same && __ldg(label1 + (size_t)b * N + i) == rl;
                    if (same) found = i;
                }
            }
            if (LABELED && want == INF_BITS) {
                want = 0u;
                found = -1;
            }
            dist2[u] = __uint_as_float(want);
            idx2[u] = found;
            s2 = __uint_as_float(want);
            if (FUSE_BWD && found >= 0) {
                const float g = __fmul_rn(__ldg(gw + 1), 2.f);
                const size_t nb = ((size_t)b * N + found) * 3;
                const float vx = __fmul_rn(g, __fsub_rn(rx, __ldg(xyz1 + nb + 0)));
                const float vy = __fmul_rn(g, __fsub_rn(ry, __ldg(xyz1 + nb + 1)));
                const float vz = __fmul_rn(g, __fsub_rn(rz, __ldg(xyz1 + nb + 2)));
                float *own = g2 + (size_t)u * 3, *other = g1 + nb;
                atomicAdd(own + 0, vx); atomicAdd(own + 1, vy); atomicAdd(own + 2, vz);
                atomicAdd(other + 0, -vx); atomicAdd(other + 1, -vy); atomicAdd(other + 2, -vz);
            }
        }
    }
    if (sums != nullptr) {
        // fused loss partial sums: warp shuffle -> shared -> one atomicAdd per block
        __shared__ float sh1[8], sh2[8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(FULL_MASK, s1, o);
            s2 += __shfl_xor_sync(FULL_MASK, s2, o);
        }
        const int w = threadIdx.x >> 5;
        if (lane == 0) {
            sh1[w] = s1;
            sh2[w] = s2;
        }
        __syncthreads();
        if (w == 0) {
            s1 = lane < 8 ? sh1[lane] : 0.f;
            s2 = lane < 8 ? sh2[lane] : 0.f;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                s1 += __shfl_xor_sync(FULL_MASK, s1, o);
                s2 += __shfl_xor_sync(FULL_MASK, s2, o);
            }
            if (lane == 0) {
                if ((int)blockIdx.x < row_blocks) atomicAdd(sums + 0, s1);
                else atomicAdd(sums + 1, s2);
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
// gd1/gd2 == nullptr selects the uniform form used by the fused sum/mean loss: every point of
// side s has the upstream gradient gw[s] (a 2-float DEVICE vector).
template <int PHASE>
__global__ void __launch_bounds__(256)
chamfer_bwd_kernel(const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                   const float *__restrict__ gd1, const float *__restrict__ gd2,
                   const int *__restrict__ idx1, const int *__restrict__ idx2, int B, int N, int M,
                   int c, float *__restrict__ g1, float *__restrict__ g2,
                   const float *__restrict__ gw) {
    const long long total1 = (long long)B * N, total = total1 + (long long)B * M;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_wait();  // indices (phase 0) / the stored own terms (phase 1) are complete and visible
    pdl_launch_dependents();
    if (t >= total) return;
    const float *a, *bb, *gd;
    const int *idx;
    float *ga, *gb;
    long long u;
    int na, nb;
    int side;
    if (t < total1) {
        u = t; a = xyz1; bb = xyz2; gd = gd1; idx = idx1; ga = g1; gb = g2; na = N; nb = M; side = 0;
    } else {
        u = t - total1; a = xyz2; bb = xyz1; gd = gd2; idx = idx2; ga = g2; gb = g1; na = M; nb = N; side = 1;
    }
    const int b = (int)(u / na);
    const int j2 = idx[u];
    const float up = gd != nullptr ? gd[u] : __ldg(gw + side);
    const float g = __fmul_rn(up, 2.f);
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

// Debug aid (option "chamfer_ws_check"): verifies the promise a caller makes with PP_CHAMFER_WS_CLEAN.
// Synchronises the stream; returns the number of 64-bit words that are not all-ones, or -1 on a CUDA error.
__global__ void keys_check_kernel(const unsigned long long *k, size_t n, unsigned *bad) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (k[i] != 0xffffffffffffffffull) atomicAdd(bad, 1u);
}

static long long keys_all_ones(const void *workspace, size_t bytes, cudaStream_t st) {
    unsigned *bad = nullptr, host = 0;
    if (cudaMalloc(&bad, sizeof(unsigned)) != cudaSuccess) return -1;
    cudaMemsetAsync(bad, 0, sizeof(unsigned), st);
    keys_check_kernel<<<296, 256, 0, st>>>((const unsigned long long *)workspace, bytes / 8, bad);
    const cudaError_t e = cudaMemcpyAsync(&host, bad, sizeof(unsigned), cudaMemcpyDeviceToHost, st);
    const cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(bad);
    return (e != cudaSuccess || e2 != cudaSuccess) ? -1 : (long long)host;
}

static size_t chamfer_keys_bytes(int B, int N, int M) {
    return sizeof(unsigned long long) * ((size_t)B * N + (size_t)B * M);
}

extern "C" size_t pp_chamfer_fwd_workspace_bytes(int B, int N, int M) {
    if (B <= 0 || N < 0 || M < 0) return 0;
    // the sweep path's layout (chamfer_sweep.cu) contains more than the packed keys of the exact one-pass kernel
    const size_t a = chamfer_keys_bytes(B, N, M), b = chamfer_sweep_workspace_bytes(B, N, M);
    return a > b ? a : b;
}

template <int Q, int THREADS, int RB, int MINB, bool ROTATE = true, bool LABELED = false, bool STAGEQ = false,
          bool ELECT = true>
static int launch_chamfer_fwd(const float *xyz1, const float *xyz2, int B, int N, int M,
                              unsigned long long *key1, unsigned long long *key2, float *dist1,
                              float *dist2, int *idx1, int *idx2, float *sums, cudaStream_t st,
                              const float *label1 = nullptr, const float *label2 = nullptr,
                              const float *gw = nullptr, float *g1 = nullptr, float *g2 = nullptr) {
    constexpr int TQ = Q * THREADS;
    const int ref_blocks = ceil_div(M, RB);
    // Query splits: enough CTAs for several waves (tail effect), but as few as possible so
    // that every CTA sweeps many query tiles past its resident references (filter depth).
    const long long want_blocks = (long long)NUM_SMS_B200 * get_option("chamfer_blocks_per_sm", 24);
    int splits = (int)ceil_div_ll(want_blocks, (long long)B * ref_blocks);
    const int max_splits = ceil_div(N, TQ);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    const int queries_per_split = ceil_div(ceil_div(N, splits), TQ) * TQ;
    splits = ceil_div(N, queries_per_split);
    PP_REQUIRE(B <= 65535 && splits <= 65535, "chamfer: grid too large (B=%d)", B);
    dim3 grid(ref_blocks, B, splits);
    {
        KernelTimer timer("chamfer_fwd", st);
        chamfer_fwd_kernel<Q, THREADS, RB, MINB, ROTATE, LABELED, STAGEQ, ELECT><<<grid, THREADS, 0, st>>>(
            xyz1, xyz2, N, M, key1, key2, queries_per_split, label1, label2, g1, g2);
    }
    PP_LAUNCH_CHECK();
    const int row_blocks = (int)ceil_div_ll((long long)B * N, 256);
    const int col_blocks = (int)ceil_div_ll((long long)B * M, 256);
    {
        KernelTimer timer("chamfer_finalize", st);
        if (!LABELED && gw != nullptr)
            PP_CUDA(launch_pdl(chamfer_finalize_kernel<Q, false, true>, dim3(row_blocks + col_blocks), dim3(256), 0,
                               st, xyz1, xyz2, B, N, M, key1, key2, dist1, dist2, idx1, idx2, sums, row_blocks,
                               label1, label2, gw, g1, g2));
        else
            PP_CUDA(launch_pdl(chamfer_finalize_kernel<Q, LABELED, false>, dim3(row_blocks + col_blocks), dim3(256),
                               0, st, xyz1, xyz2, B, N, M, key1, key2, dist1, dist2, idx1, idx2, sums, row_blocks,
                               label1, label2, gw, g1, g2));
    }
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

static int chamfer_bwd_impl(const float *xyz1, const float *xyz2, const float *graddist1, const float *graddist2,
                            const float *gw, const int32_t *idx1, const int32_t *idx2, int B, int N, int M, int c,
                            float *gradxyz1, float *gradxyz2, int device, void *stream);

// which forward the last pp_chamfer_fwd / pp_chamfer_fwd_bwd_uniform call on this thread ran (pp_chamfer_last_path):
// 0 = exact FFMA one-pass kernel (or the generic kernel), 1 = tensor-core sweep + exact resolution
static thread_local int g_chamfer_last_path = 0;

// gw/g1/g2 != nullptr: the fused forward + uniform backward (c == 3 only).
static int chamfer_fwd_impl(const float *xyz1, const float *xyz2, int B, int N, int M, int c,
                            float *dist1, float *dist2, int32_t *idx1, int32_t *idx2, float *sums,
                            void *workspace, size_t workspace_bytes, int flags, int device,
                            void *stream, const float *gw, float *g1, float *g2) {
    PP_REQUIRE(B >= 0 && N >= 0 && M >= 0 && c >= 1, "chamfer_fwd: bad sizes B=%d N=%d M=%d c=%d", B, N, M, c);
    PP_REQUIRE((long long)B * N * c < (1ll << 31) && (long long)B * M * c < (1ll << 31),
               "chamfer_fwd: B*N*c must fit int32 indexing like the reference");
    if (B == 0 || (N == 0 && M == 0)) return PP_OK;
    // empty tensors legitimately carry null data pointers
    PP_REQUIRE((N == 0 || (xyz1 && dist1 && idx1)) && (M == 0 || (xyz2 && dist2 && idx2)), "chamfer_fwd: null pointer");
    DeviceGuard guard(device);
    PP_CUDA(guard.err);
    cudaStream_t st = (cudaStream_t)stream;
    if (sums) PP_CUDA(cudaMemsetAsync(sums, 0, 2 * sizeof(float), st));
    if (N == 0 || M == 0) {
        // the reference's loops never run and leave the Python-side zero fill in place
        if (N) { PP_CUDA(cudaMemsetAsync(dist1, 0, sizeof(float) * (size_t)B * N, st)); PP_CUDA(cudaMemsetAsync(idx1, 0, sizeof(int) * (size_t)B * N, st)); }
        if (M) { PP_CUDA(cudaMemsetAsync(dist2, 0, sizeof(float) * (size_t)B * M, st)); PP_CUDA(cudaMemsetAsync(idx2, 0, sizeof(int) * (size_t)B * M, st)); }
        // nothing to match against: gradients are zero (as in chamfer_bwd_impl)
        if (gw && N) PP_CUDA(cudaMemsetAsync(g1, 0, sizeof(float) * (size_t)B * N * c, st));
        if (gw && M) PP_CUDA(cudaMemsetAsync(g2, 0, sizeof(float) * (size_t)B * M * c, st));
        return PP_OK;
    }
    if (c != 3 || get_option("chamfer_generic", 0)) {
        PP_REQUIRE(gw == nullptr, "chamfer_fwd_bwd_uniform: only available for c == 3");
        PP_REQUIRE(sums == nullptr, "chamfer_fwd: fused sums are only available for c == 3");
        return launch_generic(false, xyz1, xyz2, nullptr, nullptr, B, N, M, c, dist1, dist2, idx1, idx2, st);
    }
    PP_REQUIRE(workspace != nullptr && ((uintptr_t)workspace & 15) == 0, "chamfer_fwd: workspace null or misaligned");
    int pick = get_option("chamfer_variant", 0);
    // Tensor-core path (chamfer_sweep.cu): approximate sweep with tcgen05.mma + exact resolution, bit-identical
    // results.  Automatic choice: it wins once the clouds are large enough for its fixed costs (preparation,
    // resolution and rescan launches) to disappear behind the sweep; 51 forces it, 1..35 force the exact FFMA kernel.
    const long long pairs_per_cloud = (long long)N * M;
    g_chamfer_last_path = 0;
    if (pick == 51 || (pick == 0 && pairs_per_cloud >= get_option("chamfer_tc_min_pairs", 2048 * 2048))) {
        g_chamfer_last_path = 1;
        // Very large batches (from ~3M points): the backward folded into the resolving kernels loses to the two
        // streaming backward kernels (its scattered RED.ADDs stop fitting the L2 next to the operand tiles:
        // B=256 N=M=8192: 3.22 vs 3.12 ms); below that the folded form wins (B=128: 1.57 vs 1.59 ms).
        const bool split_bwd = gw != nullptr && (long long)B * ((long long)N + M) >= get_option("chamfer_split_bwd_points", 3 << 20);
        const int rc = chamfer_sweep_launch(xyz1, xyz2, B, N, M, dist1, dist2, idx1, idx2, sums, workspace,
                                            workspace_bytes, split_bwd ? nullptr : gw, g1, g2, st);
        if (rc != PP_OK || !split_bwd) return rc;
        return chamfer_bwd_impl(xyz1, xyz2, nullptr, nullptr, gw, idx1, idx2, B, N, M, 3, g1, g2, device, stream);
    }
    const size_t need = chamfer_keys_bytes(B, N, M);
    if (workspace_bytes < need) {
        set_error("chamfer_fwd: workspace %zu < %zu bytes", workspace_bytes, need);
        return PP_ENOSPC;
    }
    unsigned long long *key1 = (unsigned long long *)workspace;
    unsigned long long *key2 = key1 + (size_t)B * N;
    // The finalize kernel leaves the keys all-ones again; callers that own a persistent
    // workspace say so with PP_CHAMFER_WS_CLEAN and save the fill.
    if (!(flags & PP_CHAMFER_WS_CLEAN)) PP_CUDA(cudaMemsetAsync(workspace, 0xff, need, st));
    else if (get_option("chamfer_ws_check", 0)) PP_REQUIRE(keys_all_ones(workspace, need, st) == 0,
                                                           "chamfer_fwd: PP_CHAMFER_WS_CLEAN passed but the key workspace is not all-ones");
    if (pick == 0) {
        // Smaller reference blocks keep small clouds spread over all SMs.  A warp takes 32*Q = 256
        // queries: when 4-warp CTAs would leave two or more warp slots of the last query split
        // without work (N = 2500: 10 warps in 3 CTAs), 2-warp CTAs waste none.
        const int warps = ceil_div(N, 256);
        const bool narrow = ceil_div(warps, 2) * 2 < ceil_div(warps, 4) * 4;
        // Large clouds keep the direct query loads: staging gains 0.5 % there with a warm L2 and
        // loses 2-6 % when the clouds come from DRAM (a CTA-wide barrier behind the loads).
        pick = (M <= 4096) ? (narrow ? 25 : 22) : 1;
        if (pick != 1 && get_option("chamfer_noelect", 1)) pick += 10;  // election-free column publish
    }
    switch (pick) {
        case 1: return launch_chamfer_fwd<8, 128, 256, 5>(xyz1, xyz2, B, N, M, key1, key2, dist1, dist2, idx1, idx2, sums, st, nullptr, nullptr, gw, g1, g2);
        case 2: return launch_chamfer_fwd<8, 128, 128, 5>(xyz1, xyz2, B, N, M, key1, key2, dist1, dist2, idx1, idx2, sums, st, nullptr, nullptr, gw, g1, g2);
        case 5: return launch_chamfer_fwd<8, 64, 128, 10>(xyz1, xyz2, B, N, M, key1, key2, dist1, dist2, idx1, idx2, sums, st, nullptr, nullptr, gw, g1, g2);
        case 21: return launch_chamfer_fwd<8, 128, 256, 5, true, false, true>(xyz1, xyz2, B, N, M, key1, key2, dist1, dist2, idx1, idx2, sums, st, nullptr, nullptr, gw, g1, g2);
        case 22: return launch_chamfer_fwd<8, 128, 128, 5, true, false, true>(xyz1, xyz2, B, N, M, key1, key2, dist1, dist2, idx1, idx2, sums, st, nullptr, nullptr, gw, g1, g2);
        case 25: return launch_chamfer_fwd<8, 64, 128, 10, true, false, true>(xyz1, xyz2, B, N, M, key1, key2, dist1, dist2, idx1, idx2, sums, st, nullptr, nullptr, gw, g1, g2);
        case 31: return launch_chamfer_fwd<8, 128, 256, 5, true, false, true, false>(xyz1, xyz2, B, N, M, key1, key2, dist1, dist2, idx1, idx2, sums, st, nullptr, nullptr, gw, g1, g2);
        case 32: return launch_chamfer_fwd<8, 128, 128, 5, true, false, true, false>(xyz1, xyz2, B, N, M, key1, key2, dist1, dist2, idx1, idx2, sums, st, nullptr, nullptr, gw, g1, g2);
        case 35: return launch_chamfer_fwd<8, 64, 128, 10, true, false, true, false>(xyz1, xyz2, B, N, M, key1, key2, dist1, dist2, idx1, idx2, sums, st, nullptr, nullptr, gw, g1, g2);
        case 13: return launch_chamfer_fwd<8, 128, 256, 5, false>(xyz1, xyz2, B, N, M, key1, key2, dist1, dist2, idx1, idx2, sums, st, nullptr, nullptr, gw, g1, g2);
        case 14: return launch_chamfer_fwd<8, 128, 128, 5, false>(xyz1, xyz2, B, N, M, key1, key2, dist1, dist2, idx1, idx2, sums, st, nullptr, nullptr, gw, g1, g2);
        default: break;
    }
    set_error("chamfer_fwd: unknown variant %d", pick);
    return PP_EINVAL;
}

extern "C" int pp_chamfer_last_path(void) { return g_chamfer_last_path; }

extern "C" int pp_chamfer_fwd(const float *xyz1, const float *xyz2, int B, int N, int M, int c,
                              float *dist1, float *dist2, int32_t *idx1, int32_t *idx2, float *sums,
                              void *workspace, size_t workspace_bytes, int flags, int device,
                              void *stream) {
    return chamfer_fwd_impl(xyz1, xyz2, B, N, M, c, dist1, dist2, idx1, idx2, sums, workspace,
                            workspace_bytes, flags, device, stream, nullptr, nullptr, nullptr);
}

extern "C" int pp_chamfer_fwd_bwd_uniform(const float *xyz1, const float *xyz2, const float *gw, int B,
                                          int N, int M, float *dist1, float *dist2, int32_t *idx1,
                                          int32_t *idx2, float *sums, float *gradxyz1, float *gradxyz2,
                                          void *workspace, size_t workspace_bytes, int flags,
                                          int device, void *stream) {
    PP_REQUIRE(gw != nullptr || B == 0, "chamfer_fwd_bwd_uniform: null weight vector");
    PP_REQUIRE(B == 0 || ((N == 0 || gradxyz1) && (M == 0 || gradxyz2)), "chamfer_fwd_bwd_uniform: null gradient pointer");
    return chamfer_fwd_impl(xyz1, xyz2, B, N, M, 3, dist1, dist2, idx1, idx2, sums, workspace,
                            workspace_bytes, flags, device, stream, gw, gradxyz1, gradxyz2);
}

extern "C" int pp_chamfer_labeled_fwd(const float *xyz1, const float *xyz2, const float *label1,
                                      const float *label2, int B, int N, int M, int c, float *dist1,
                                      float *dist2, int32_t *idx1, int32_t *idx2, void *workspace,
                                      size_t workspace_bytes, int flags, int device, void *stream) {
    PP_REQUIRE(B >= 0 && N >= 0 && M >= 0 && c >= 1, "chamfer_labeled_fwd: bad sizes");
    if (B == 0 || (N == 0 && M == 0)) return PP_OK;
    PP_REQUIRE((N == 0 || (xyz1 && label1 && dist1 && idx1)) && (M == 0 || (xyz2 && label2 && dist2 && idx2)), "chamfer_labeled_fwd: null pointer");
    DeviceGuard guard(device);
    PP_CUDA(guard.err);
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0 || M == 0) {
        if (N) { PP_CUDA(cudaMemsetAsync(dist1, 0, sizeof(float) * (size_t)B * N, st)); PP_CUDA(cudaMemsetAsync(idx1, 0, sizeof(int) * (size_t)B * N, st)); }
        if (M) { PP_CUDA(cudaMemsetAsync(dist2, 0, sizeof(float) * (size_t)B * M, st)); PP_CUDA(cudaMemsetAsync(idx2, 0, sizeof(int) * (size_t)B * M, st)); }
        return PP_OK;
    }
    if (c != 3 || workspace == nullptr || get_option("chamfer_generic", 0))
        return launch_generic(true, xyz1, xyz2, label1, label2, B, N, M, c, dist1, dist2, idx1, idx2, st);
    // fast path: the one-pass kernel with the label mask (same key workspace protocol as pp_chamfer_fwd)
    PP_REQUIRE((long long)B * N * c < (1ll << 31) && (long long)B * M * c < (1ll << 31),
               "chamfer_labeled_fwd: B*N*c must fit int32 indexing like the reference");
    const size_t need = chamfer_keys_bytes(B, N, M);
    PP_REQUIRE(((uintptr_t)workspace & 7) == 0, "chamfer_labeled_fwd: workspace misaligned");
    if (workspace_bytes < need) {
        set_error("chamfer_labeled_fwd: workspace %zu < %zu bytes", workspace_bytes, need);
        return PP_ENOSPC;
    }
    unsigned long long *key1 = (unsigned long long *)workspace;
    unsigned long long *key2 = key1 + (size_t)B * N;
    if (!(flags & PP_CHAMFER_WS_CLEAN)) PP_CUDA(cudaMemsetAsync(workspace, 0xff, need, st));
    if (M <= 4096)
        return launch_chamfer_fwd<8, 128, 128, 4, true, true>(xyz1, xyz2, B, N, M, key1, key2, dist1, dist2, idx1, idx2,
                                                              nullptr, st, label1, label2);
    return launch_chamfer_fwd<8, 128, 256, 4, true, true>(xyz1, xyz2, B, N, M, key1, key2, dist1, dist2, idx1, idx2,
                                                          nullptr, st, label1, label2);
}

static int chamfer_bwd_impl(const float *xyz1, const float *xyz2, const float *graddist1,
                            const float *graddist2, const float *gw, const int32_t *idx1, const int32_t *idx2, int B, int N, int M, int c,
                            float *gradxyz1, float *gradxyz2, int device, void *stream) {
    PP_REQUIRE(B >= 0 && N >= 0 && M >= 0 && c >= 1, "chamfer_bwd: bad sizes");
    if (B == 0 || (N == 0 && M == 0)) return PP_OK;
    PP_REQUIRE((N == 0 || (xyz1 && idx1 && gradxyz1)) && (M == 0 || (xyz2 && idx2 && gradxyz2)), "chamfer_bwd: null pointer");
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
    KernelTimer timer("chamfer_bwd", st);
    PP_CUDA(launch_pdl(chamfer_bwd_kernel<0>, dim3(blocks), dim3(256), 0, st, xyz1, xyz2, graddist1, graddist2, idx1,
                       idx2, B, N, M, c, gradxyz1, gradxyz2, gw));
    PP_LAUNCH_CHECK();
    PP_CUDA(launch_pdl(chamfer_bwd_kernel<1>, dim3(blocks), dim3(256), 0, st, xyz1, xyz2, graddist1, graddist2, idx1,
                       idx2, B, N, M, c, gradxyz1, gradxyz2, gw));
    PP_LAUNCH_CHECK();
    return PP_OK;
}

extern "C" int pp_chamfer_bwd(const float *xyz1, const float *xyz2, const float *graddist1,
                              const float *graddist2, const int32_t *idx1, const int32_t *idx2, int B,
                              int N, int M, int c, float *gradxyz1, float *gradxyz2, int device,
                              void *stream) {
    PP_REQUIRE((N == 0 || B == 0 || graddist1) && (M == 0 || B == 0 || graddist2), "chamfer_bwd: null graddist");
    return chamfer_bwd_impl(xyz1, xyz2, graddist1, graddist2, nullptr, idx1, idx2, B, N, M, c,
                            gradxyz1, gradxyz2, device, stream);
}

extern "C" int pp_chamfer_bwd_uniform(const float *xyz1, const float *xyz2, const float *gw,
                                      const int32_t *idx1, const int32_t *idx2, int B, int N, int M,
                                      int c, float *gradxyz1, float *gradxyz2, int device,
                                      void *stream) {
    PP_REQUIRE(gw != nullptr || B == 0, "chamfer_bwd_uniform: null weight vector");
    return chamfer_bwd_impl(xyz1, xyz2, nullptr, nullptr, gw, idx1, idx2, B, N, M, c,
                            gradxyz1, gradxyz2, device, stream);
}
