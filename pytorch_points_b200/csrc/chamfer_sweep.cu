// chamfer_sweep.cu -- Chamfer / nndistance forward for c == 3 clouds: an approximate sweep on the
// packed FP32 pipe followed by an exact resolution of the few surviving candidates.
//
// Replaces NmDistanceKernel x2 (_ext/nmdistance_cuda.cu:8-49,127-128).  The reference's distance
//      d = fma(tz,tz, fma(ty,ty, rn(tx*tx))),  t = rn(ref - query)                    (6 lane-ops)
// has to be reproduced bit for bit, and so does its lowest-index tie rule -- but only for the pair
// that WINS.  Every other pair merely has to be shown to lose.  So the sweep orders the pairs of a
// query by the expansion  e = |r|^2 - 2 q.r  on centred coordinates (3 FFMA per pair: the
// "GEMM-expansion distance path" on the FFMA pipe; |q|^2 is a per-query constant), whose error
// against the reference's value is bounded by EPS = 64 u R^2 (u = 2^-24, R = largest centred norm;
// derivation in DESIGN.md §3.1).  Per query it keeps the best value, the 32-reference granule it
// occurred in and the runner-up value over all OTHER granules:
//   runner-up > best + TAU (TAU = 2.5 EPS)  =>  every true minimiser (ties included) lies inside the
//       recorded granule, whose 32 pairs are re-evaluated with the exact chain;
//   otherwise the point is "ambiguous" (about 1 % of a uniform cloud, every point of a lattice): a
//       rescan walks all partners, exact-evaluates those within TAU of the best value and takes the
//       lowest index among the exact minima.
// Results are therefore identical to the reference's on every input; only the time depends on how
// many points are ambiguous.
//
// Both directions (dist1/idx1 and dist2/idx2) run the same one-sided pass with the roles of the
// clouds swapped, like the reference's two launches -- but at 3 lane-ops per pair instead of 6 and
// without any cross-thread traffic in the loop: a thread owns 8 queries for a whole chunk of
// references, so there is no column side, no filter, no atomics and no barrier in the hot loop.
//
// Kernels (one stream, programmatic dependent launch):
//   cs_prep_kernel      centre (mean of the leading points), centred coordinates, norms, R^2; every
//                       point in query form float4 {-2x,-2y,-2z,|p|^2} and in reference form (SoA
//                       blocks of 128: x[128] y[128] z[128] |p|^2[128]); resets keys, gradients.
//   cs_rowpass_kernel   CTA = (direction, cloud, tile of WARPS*256 queries, chunk of reference
//                       blocks).  The reference blocks stream through a 4-stage shared-memory ring
//                       filled by cp.async.bulk (UBLKCP) on full/empty mbarriers; warps drift apart by
//                       up to three blocks, nobody waits at a CTA barrier.  Hot loop per 4 references
//                       x 8 queries: 4 LDS.128 (broadcast), 48 FFMA2, 16 FMNMX3.
//   cs_finalize_kernel  exact resolution of the recorded granule (dist / idx, loss sums, fused
//                       backward); ambiguous points go to per-cloud lists.
//   cs_rescan_kernel    the ambiguous points, 32 at a time per CTA against their whole partner cloud.
#include "pp_common.cuh"

namespace pp {
namespace {

constexpr int CS_RB = 128;     // references per block
constexpr int CS_WT = 256;     // queries per warp (8 per lane, lane-interleaved)
constexpr int CS_GR = 32;      // granule (references)
constexpr int CS_STAGES = 4;   // reference blocks in flight per CTA
constexpr int CS_RS = 32;      // entries a rescan CTA handles at once
constexpr unsigned CS_INF_BITS = 0x7f800000u;
constexpr unsigned long long CS_KEY_INIT = 0x7f800000ffffffffull;
// TAU = 2.5 * EPS.  FFMA sweep: EPS = 64 u R^2 (u = 2^-24).  Tensor-core sweep (3xTF32 split): the dropped
// low*low products and the tensor core's accumulation add to the budget: EPS = 128 u R^2.
constexpr float CS_TAU_PER_R2 = 160.f * 5.9604644775390625e-8f;
constexpr float CS_TAU_PER_R2_TC = 320.f * 5.9604644775390625e-8f;

// One direction of the problem: `nq` query points (cloud A) against `nr` reference points (cloud B).
struct CsDir {
    const float *qxyz, *rxyz;        // original clouds (B,nq,3) / (B,nr,3)
    const float4 *qform;             // (B, npad) prepared queries
    const float *rform;              // (B, rblk, 4, 128) prepared references
    const float *aform;              // (B, ceil(nq/128), 4 chunks, 128 rows, 4) tensor-core A operand of the queries
    const float *bform;              // (B, rblk, 4 chunks, 128 rows, 4) tensor-core B operand of the references
    unsigned long long *key;         // (B, nq) value bits << 32 | granule
    unsigned *sec;                   // (B, nq) runner-up value bits
    uint2 *list;                     // (B, nq) ambiguous points: (index in cloud, best value bits)
    unsigned *count;                 // (B) entries in list
    float *dist;                     // outputs (B, nq)
    int *idx;
    float *gq, *gr;                  // fused backward: gradient of the query / reference cloud (or null)
    int nq, nr, npad, rblk;
    int nchunks, chunk_blocks, tiles;
};

struct CsArgs {
    CsDir d[2];
    const unsigned *r2bits;  // (B) bits of the largest centred squared norm
    float tau_per_r2;        // TAU = tau_per_r2 * R^2
    float rescan_per_r2;     // the rescan's window above the recorded best value (it re-approximates with FFMA)
    float *sums;             // [sum(dist1), sum(dist2)] or null
    const float *gw;         // fused backward weights or null
    int B;
};

struct CsLayout {
    size_t ctrl, qform[2], rform[2], key[2], sec[2], list[2], aform[2], bform[2], total;
    int npad[2], blk[2];
};

CsLayout cs_layout(int B, int N, int M, bool tc = false) {
    CsLayout L;
    const int n[2] = {N, M};
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t o = 0;
    L.ctrl = o; o += up(4 * 3 * (size_t)B);  // r2bits[B], count[2][B]
    for (int s = 0; s < 2; s++) {
        L.npad[s] = ceil_div(n[s], CS_WT) * CS_WT;
        L.blk[s] = ceil_div(n[s], CS_RB);
        L.qform[s] = o; o += up(16 * (size_t)B * L.npad[s]);
        L.rform[s] = o; o += up(16 * (size_t)B * L.blk[s] * CS_RB);
        L.key[s] = o;   o += up(8 * (size_t)B * n[s]);
        L.sec[s] = o;   o += up(4 * (size_t)B * n[s]);
        L.list[s] = o;  o += up(8 * (size_t)B * n[s]);
        L.aform[s] = L.bform[s] = 0;
        if (tc) {  // 16 tf32 values per point and role
            L.aform[s] = o; o += up(64 * (size_t)B * L.blk[s] * CS_RB);
            L.bform[s] = o; o += up(64 * (size_t)B * L.blk[s] * CS_RB);
        }
    }
    L.total = o;
    return L;
}

// ---- small PTX helpers ------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "CS_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra CS_DONE;\n\t"
        "bra CS_WAIT;\n\t"
        "CS_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// {a.x*q + c.x, a.y*q + c.y}: FFMA2 with the .F32 broadcast operand
__device__ __forceinline__ float2 fma2_bcast(float2 a, float q, float2 c) {
    float2 d;
    // volatile: the hot loop's issue order is the source order (see cs_rowpass_kernel)
    asm volatile("{\n\t.reg .b64 ra, rq, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rq, {%4, %4};\n\tmov.b64 rc, {%5, %6};\n\t"
        "fma.rn.f32x2 rd, ra, rq, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(q), "f"(c.x), "f"(c.y));
    return d;
}

// FMNMX3 pinned in source order like fma2_bcast (an FMNMX3 between two FFMA2 of one level would cost
// them the shared operand in the reuse cache)
__device__ __forceinline__ float fmin3_pinned(float a, float b, float c) {
    float r;
    asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// the value the sweep orders by: e = fma(z,qz', fma(y,qy', fma(x,qx', |r|^2))), q' = -2 q
__device__ __forceinline__ float approx_e(float x, float y, float z, float rr, float qx, float qy, float qz) {
    return __fmaf_rn(z, qz, __fmaf_rn(y, qy, __fmaf_rn(x, qx, rr)));
}

__device__ __forceinline__ void bwd_term(float gg, float px, float py, float pz, const float *__restrict__ nbr,
                                         float *own, float *oth) {
    const float vx = __fmul_rn(gg, __fsub_rn(px, __ldg(nbr + 0)));
    const float vy = __fmul_rn(gg, __fsub_rn(py, __ldg(nbr + 1)));
    const float vz = __fmul_rn(gg, __fsub_rn(pz, __ldg(nbr + 2)));
    atomicAdd(own + 0, vx); atomicAdd(own + 1, vy); atomicAdd(own + 2, vz);
    atomicAdd(oth + 0, -vx); atomicAdd(oth + 1, -vy); atomicAdd(oth + 2, -vz);
}

// ---- preparation --------------------------------------------------------------------------------
// grid (chunks, B, 2): blockIdx.z = cloud.  Cloud s is the query side of direction s and the
// reference side of direction 1 - s.
__global__ void __launch_bounds__(256)
cs_prep_kernel(const float *__restrict__ xyz1, const float *__restrict__ xyz2, int N, int M, float4 *__restrict__ qf0,
               float4 *__restrict__ qf1, float *__restrict__ rf0, float *__restrict__ rf1,
               unsigned long long *__restrict__ key0, unsigned long long *__restrict__ key1,
               unsigned *__restrict__ sec0, unsigned *__restrict__ sec1, unsigned *__restrict__ r2bits,
               float *__restrict__ g1, float *__restrict__ g2) {
    pdl_launch_dependents();
    const int b = blockIdx.y, s = blockIdx.z;
    const int lane = threadIdx.x & 31;
    __shared__ float s_c[3];
    if (threadIdx.x < 32) {
        // centre = mean of the (up to) 32 leading points of each cloud: the same instruction
        // sequence in every CTA of this cloud pair, hence the same bits
        float sx = 0.f, sy = 0.f, sz = 0.f;
        if (lane < N) {
            const float *p = xyz1 + ((size_t)b * N + lane) * 3;
            sx += p[0]; sy += p[1]; sz += p[2];
        }
        if (lane < M) {
            const float *p = xyz2 + ((size_t)b * M + lane) * 3;
            sx += p[0]; sy += p[1]; sz += p[2];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sx += __shfl_xor_sync(FULL_MASK, sx, o);
            sy += __shfl_xor_sync(FULL_MASK, sy, o);
            sz += __shfl_xor_sync(FULL_MASK, sz, o);
        }
        const float inv = 1.f / (float)(min(32, N) + min(32, M));
        if (lane == 0) { s_c[0] = sx * inv; s_c[1] = sy * inv; s_c[2] = sz * inv; }
    }
    __syncthreads();
    const float cx = s_c[0], cy = s_c[1], cz = s_c[2];
    const float *xyz = s ? xyz2 : xyz1;
    const int n = s ? M : N;
    float4 *qf = s ? qf1 : qf0;                  // query form of cloud s (direction s)
    float *rf = s ? rf1 : rf0;                   // reference form of cloud s (used by direction 1 - s)
    unsigned long long *key = s ? key1 : key0;
    unsigned *sec = s ? sec1 : sec0;
    float *g = s ? g2 : g1;
    const int npad = ceil_div(n, CS_WT) * CS_WT;  // a multiple of CS_RB as well
    const int rblk = ceil_div(n, CS_RB);
    float r2 = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npad; i += gridDim.x * blockDim.x) {
        float x = 0.f, y = 0.f, z = 0.f, nn = PP_INF;  // padding: |p|^2 = +inf, never a minimum
        if (i < n) {
            const size_t t = (size_t)b * n + i;
            const float *p = xyz + t * 3;
            x = __fsub_rn(p[0], cx); y = __fsub_rn(p[1], cy); z = __fsub_rn(p[2], cz);
            nn = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
            r2 = fmaxf(r2, nn);
            key[t] = CS_KEY_INIT;
            sec[t] = CS_INF_BITS;
            if (g != nullptr) { g[t * 3 + 0] = 0.f; g[t * 3 + 1] = 0.f; g[t * 3 + 2] = 0.f; }
        }
        qf[(size_t)b * npad + i] = make_float4(-2.f * x, -2.f * y, -2.f * z, nn);
        if (i < rblk * CS_RB) {
            float *blk = rf + ((size_t)b * rblk + i / CS_RB) * (4 * CS_RB);
            const int o = i % CS_RB;
            blk[o] = x; blk[CS_RB + o] = y; blk[2 * CS_RB + o] = z; blk[3 * CS_RB + o] = nn;
        }
    }
    const unsigned rb = __reduce_max_sync(FULL_MASK, __float_as_uint(r2));  // r2 >= 0: bits order like values
    if (lane == 0 && rb != 0u) atomicMax(r2bits + b, rb);
}

// ---- the one-sided pass ---------------------------------------------------------------------------
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 4 ? 4 : 8)
cs_rowpass_kernel(const CsArgs args) {
    __shared__ __align__(128) float sRef[CS_STAGES][4 * CS_RB];  // ring of reference blocks: x | y | z | |r|^2
    __shared__ __align__(8) unsigned long long sBar[2 * CS_STAGES];  // full[stage], empty[stage]

    pdl_launch_dependents();
    const CsDir &D = args.d[blockIdx.z];
    const int b = blockIdx.y;
    if ((int)blockIdx.x >= D.tiles * D.nchunks) return;
    const int tile = blockIdx.x / D.nchunks, chunk = blockIdx.x % D.nchunks;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int blk0 = chunk * D.chunk_blocks;
    const int nblk = min(D.rblk, blk0 + D.chunk_blocks) - blk0;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < CS_STAGES; i++) {
            mbar_init(smem_u32(sBar + i), 1);
            mbar_init(smem_u32(sBar + CS_STAGES + i), WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();  // the prepared arrays, R^2 and the reset keys are complete and visible

    const float *rsrc = D.rform + ((size_t)b * D.rblk + blk0) * (4 * CS_RB);
    if (threadIdx.x == 0) {
        for (int i = 0; i < min(CS_STAGES, nblk); i++) {
            mbar_expect_tx(smem_u32(sBar + i), 16 * CS_RB);
            bulk_g2s(smem_u32(&sRef[i][0]), rsrc + (size_t)i * (4 * CS_RB), 16 * CS_RB, smem_u32(sBar + i));
        }
    }
    // a thread's 8 queries: i = first + q * 32 (lane-interleaved: coalesced 512-byte rows)
    const int first = (tile * WARPS + w) * CS_WT + lane;
    const bool live = (tile * WARPS + w) * CS_WT < D.nq;  // warp-uniform: some query of this warp exists
    float qx[8], qy[8], qz[8], best[8], second[8];
    int gran[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) v = __ldg(D.qform + (size_t)b * D.npad + first + q * 32);
        qx[q] = v.x; qy[q] = v.y; qz[q] = v.z;
        best[q] = PP_INF; second[q] = PP_INF; gran[q] = 0;
    }

#pragma unroll 1
    for (int it = 0; it < nblk; it++) {
        const int st = it % CS_STAGES;
        const unsigned par = (unsigned)(it / CS_STAGES) & 1u;
        if (threadIdx.x == 0 && it >= 1 && it - 1 + CS_STAGES < nblk) {
            // refill the stage the CTA finished one block ago (the other warps have had a whole block
            // of time to release it, so this wait is normally over before it starts)
            const int ps = (it - 1) % CS_STAGES;
            mbar_wait(smem_u32(sBar + CS_STAGES + ps), (unsigned)((it - 1) / CS_STAGES) & 1u);
            mbar_expect_tx(smem_u32(sBar + ps), 16 * CS_RB);
            bulk_g2s(smem_u32(&sRef[ps][0]), rsrc + (size_t)(it - 1 + CS_STAGES) * (4 * CS_RB), 16 * CS_RB,
                     smem_u32(sBar + ps));
        }
        __syncwarp();
        mbar_wait(smem_u32(sBar + st), par);
        if (live) {
            const float *sX = sRef[st], *sY = sX + CS_RB, *sZ = sX + 2 * CS_RB, *sR = sX + 3 * CS_RB;
#pragma unroll 1
            for (int gi = 0; gi < CS_RB / CS_GR; gi++) {
                float gm[8];
#pragma unroll
                for (int q = 0; q < 8; q++) gm[q] = PP_INF;
#pragma unroll 2
                for (int s4 = 0; s4 < CS_GR; s4 += 4) {
                    const int jj = gi * CS_GR + s4;
                    const float4 X = *reinterpret_cast<const float4 *>(sX + jj);
                    const float4 Y = *reinterpret_cast<const float4 *>(sY + jj);
                    const float4 Z = *reinterpret_cast<const float4 *>(sZ + jj);
                    const float4 R = *reinterpret_cast<const float4 *>(sR + jj);
                    const float2 x01 = make_float2(X.x, X.y), x23 = make_float2(X.z, X.w);
                    const float2 y01 = make_float2(Y.x, Y.y), y23 = make_float2(Y.z, Y.w);
                    const float2 z01 = make_float2(Z.x, Z.y), z23 = make_float2(Z.z, Z.w);
                    const float2 r01 = make_float2(R.x, R.y), r23 = make_float2(R.z, R.w);
                    // Level by level over the 8 queries: consecutive FFMA2 share the reference pair
                    // (operand reuse cache: one register-file read of the query scalar per instruction,
                    // no bank conflict) and the 8 chains are independent (latency hidden without a
                    // second warp).
                    float2 acc[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) acc[q] = fma2_bcast(x01, qx[q], r01);
#pragma unroll
                    for (int q = 0; q < 8; q++) acc[q] = fma2_bcast(y01, qy[q], acc[q]);
#pragma unroll
                    for (int q = 0; q < 8; q++) acc[q] = fma2_bcast(z01, qz[q], acc[q]);
#pragma unroll
                    for (int q = 0; q < 8; q++) gm[q] = fmin3_pinned(gm[q], acc[q].x, acc[q].y);
#pragma unroll
                    for (int q = 0; q < 8; q++) acc[q] = fma2_bcast(x23, qx[q], r23);
#pragma unroll
                    for (int q = 0; q < 8; q++) acc[q] = fma2_bcast(y23, qy[q], acc[q]);
#pragma unroll
                    for (int q = 0; q < 8; q++) acc[q] = fma2_bcast(z23, qz[q], acc[q]);
#pragma unroll
                    for (int q = 0; q < 8; q++) gm[q] = fmin3_pinned(gm[q], acc[q].x, acc[q].y);
                }
                // granule done: best / runner-up over granules, granule of the best (equal minima in two
                // granules leave runner-up == best: the point is ambiguous, as it must be)
                const int gid = (blk0 + it) * (CS_RB / CS_GR) + gi;
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    second[q] = fminf(second[q], fmaxf(best[q], gm[q]));
                    if (gm[q] < best[q]) gran[q] = gid;
                    best[q] = fminf(best[q], gm[q]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(sBar + CS_STAGES + st));
    }
    if (!live) return;

    // ---- publish (value + |q|^2 + TAU keeps every published value positive: its bits order like values)
    const float tau = __uint_as_float(__ldcg(args.r2bits + b)) * args.tau_per_r2;
    unsigned long long *K = D.key + (size_t)b * D.nq;
    unsigned *S = D.sec + (size_t)b * D.nq;
    unsigned long long key[8], old[8];
    unsigned sb[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const int i = first + q * 32;
        const float qq = i < D.nq ? __fadd_rn(__ldg(&D.qform[(size_t)b * D.npad + i].w), tau) : 0.f;
        key[q] = ((unsigned long long)__float_as_uint(__fadd_rn(best[q], qq)) << 32) | (unsigned)gran[q];
        sb[q] = __float_as_uint(__fadd_rn(second[q], qq));  // +inf stays +inf
    }
    if (D.nchunks == 1) {
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int i = first + q * 32;
            if (i < D.nq) { K[i] = key[q]; S[i] = sb[q]; }
        }
    } else {
        // all eight exchanges in flight before the first dependent one: the loser of every comparison
        // at the key is a runner-up candidate
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int i = first + q * 32;
            old[q] = i < D.nq ? atomicMin(K + i, key[q]) : 0ull;
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int i = first + q * 32;
            if (i < D.nq) atomicMin(S + i, min((unsigned)((old[q] > key[q] ? old[q] : key[q]) >> 32), sb[q]));
        }
    }
}

// ---- the one-sided pass on the tensor cores (chamfer_variant 51) ----------------------------------
// Same contract as cs_rowpass_kernel, but e = |r|^2 - 2 q.r comes out of tcgen05.mma (kind::tf32, fp32
// accumulators in TMEM) instead of 3 FFMA per pair.  fp32 accuracy from TF32 operands by splitting every
// coordinate into hi + lo (each exactly representable in TF32) and spending 12 of K = 16 slots:
//      A row (query):     qh.x qh.y qh.z qh.x | qh.y qh.z ql.x ql.y | ql.z 1 1 1 | 0 0 0 0        (q = -2 * centred point)
//      B row (reference): rh.x rh.y rh.z rl.x | rl.y rl.z rh.x rh.y | rh.z n0 n1 n2 | 0 0 0 0    (|r|^2 = n0 + n1 + n2)
// -> sum = qh.rh + qh.rl + ql.rh + |r|^2; only ql.rl (<= 2^-22 |q||r|) is dropped.  Every product of two
// TF32 values is exact in fp32; what the tensor core's adder tree loses is covered by EPS.
// Operands are K-major without swizzle: 16-byte chunk c of row r sits at c * 2048 + r * 16 of an 8 KB
// tile (8 rows x 16 B core matrices, SBO = 128 B, LBO = 2048 B), so ONE cp.async.bulk moves a whole block.
// CTA = 128 queries (TMEM lane = query) x a chunk of 128-reference blocks; 6 warps: 0-3 epilogue (each
// reads its 32 lanes with tcgen05.ld 32x32b.x32 and keeps one query per thread), 4 = copy issuer + TMEM
// owner, 5 = MMA issuer.  Two 128-column accumulators alternate, so the MMA of block i+1 runs while
// block i is scanned; 256 TMEM columns per CTA -> two CTAs per SM.
constexpr int TC_STAGES = 3;         // B blocks in flight
constexpr int TC_TILE_BYTES = 8192;  // 128 rows x 16 tf32

__device__ __forceinline__ float to_tf32(float v) {  // round to nearest TF32; the low 13 bits come out zero
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// grid (ceil(npad/256), B, 2): blockIdx.z = cloud.  Reads the query form written by cs_prep_kernel.
__global__ void __launch_bounds__(256)
cs_prep_tc_kernel(const float4 *__restrict__ qf0, const float4 *__restrict__ qf1, int N, int M,
                  float *__restrict__ af0, float *__restrict__ af1, float *__restrict__ bf0, float *__restrict__ bf1) {
    pdl_wait();
    pdl_launch_dependents();
    const int b = blockIdx.y, s = blockIdx.z;
    const int n = s ? M : N;
    const int npad = ceil_div(n, CS_WT) * CS_WT, rows = ceil_div(n, CS_RB) * CS_RB;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const float4 q = (s ? qf1 : qf0)[(size_t)b * npad + i];  // {-2x, -2y, -2z, |p|^2}; padding: {0,0,0,inf}
    float *at = (s ? af1 : af0) + ((size_t)b * (rows / CS_RB) + i / CS_RB) * (TC_TILE_BYTES / 4) + (i % CS_RB) * 4;
    float *bt = (s ? bf1 : bf0) + ((size_t)b * (rows / CS_RB) + i / CS_RB) * (TC_TILE_BYTES / 4) + (i % CS_RB) * 4;
    const bool pad = i >= n;
    const float qhx = to_tf32(q.x), qhy = to_tf32(q.y), qhz = to_tf32(q.z);
    const float qlx = to_tf32(q.x - qhx), qly = to_tf32(q.y - qhy), qlz = to_tf32(q.z - qhz);
    const float x = -0.5f * q.x, y = -0.5f * q.y, z = -0.5f * q.z;  // exact
    const float rhx = to_tf32(x), rhy = to_tf32(y), rhz = to_tf32(z);
    const float rlx = to_tf32(x - rhx), rly = to_tf32(y - rhy), rlz = to_tf32(z - rhz);
    // padding references carry a huge (finite, TF32-exact) norm: never a minimum, no inf * 0 in the MMA
    const float nn = pad ? 1.0e30f : q.w;
    const float n0 = to_tf32(nn), n1 = pad ? 0.f : to_tf32(nn - n0), n2 = pad ? 0.f : to_tf32(nn - n0 - n1);
    constexpr int CH = TC_TILE_BYTES / 16;  // floats between two 16-byte chunks of one row
    *reinterpret_cast<float4 *>(at) = make_float4(qhx, qhy, qhz, qhx);
    *reinterpret_cast<float4 *>(at + CH) = make_float4(qhy, qhz, qlx, qly);
    *reinterpret_cast<float4 *>(at + 2 * CH) = make_float4(qlz, 1.f, 1.f, 1.f);
    *reinterpret_cast<float4 *>(at + 3 * CH) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4 *>(bt) = make_float4(rhx, rhy, rhz, rlx);
    *reinterpret_cast<float4 *>(bt + CH) = make_float4(rly, rlz, rhx, rhy);
    *reinterpret_cast<float4 *>(bt + 2 * CH) = make_float4(rhz, n0, n1, n2);
    *reinterpret_cast<float4 *>(bt + 3 * CH) = make_float4(0.f, 0.f, 0.f, 0.f);
}

// shared-memory matrix descriptor of a K-major, un-swizzled 128-row operand tile (layout above)
__device__ __forceinline__ unsigned long long tc_smem_desc(unsigned smem_addr) {
    return (unsigned long long)((smem_addr >> 4) & 0x3fffu)        // start address
           | ((unsigned long long)((CS_RB * 16) >> 4) << 16)      // leading byte offset: next 16-byte K chunk
           | ((unsigned long long)(128 >> 4) << 32)               // stride byte offset: next 8-row core matrix
           | (1ull << 46);                                        // descriptor version (Blackwell); no swizzle
}

__device__ __forceinline__ void tc_mma_tf32(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc,
                                            unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(unsigned bar) {  // the mbarrier completes when every MMA issued so far has
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_ld32(unsigned taddr, float (&v)[32]) {
    unsigned r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(192, 2)
cs_rowpass_tc_kernel(const CsArgs args) {
    __shared__ __align__(128) unsigned char sA[TC_TILE_BYTES];
    __shared__ __align__(128) unsigned char sB[TC_STAGES][TC_TILE_BYTES];
    // a_full | b_full[S] | b_empty[S] | t_full[2] | t_empty[2]
    __shared__ __align__(8) unsigned long long sBar[1 + 2 * TC_STAGES + 4];
    __shared__ unsigned sTmem;

    pdl_launch_dependents();
    const CsDir &D = args.d[blockIdx.z];
    const int b = blockIdx.y;
    if ((int)blockIdx.x >= D.tiles * D.nchunks) return;
    const int tile = blockIdx.x / D.nchunks, chunk = blockIdx.x % D.nchunks;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int blk0 = chunk * D.chunk_blocks;
    const int nblk = min(D.rblk, blk0 + D.chunk_blocks) - blk0;
    const unsigned bar_a = smem_u32(sBar), bar_bf = smem_u32(sBar + 1), bar_be = smem_u32(sBar + 1 + TC_STAGES);
    const unsigned bar_tf = smem_u32(sBar + 1 + 2 * TC_STAGES), bar_te = smem_u32(sBar + 3 + 2 * TC_STAGES);

    if (threadIdx.x == 0) {
        mbar_init(bar_a, 1);
#pragma unroll
        for (int i = 0; i < TC_STAGES; i++) { mbar_init(bar_bf + 8 * i, 1); mbar_init(bar_be + 8 * i, 1); }
#pragma unroll
        for (int i = 0; i < 2; i++) { mbar_init(bar_tf + 8 * i, 1); mbar_init(bar_te + 8 * i, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (w == 4) {  // one warp owns the tensor-memory allocation: 256 columns = two 128-column accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&sTmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = sTmem;
    pdl_wait();  // the prepared operands, R^2 and the reset keys are complete and visible

    if (w == 4) {
        if (lane == 0) {  // ---- copy issuer: the query tile once, then the reference blocks through the ring
            mbar_expect_tx(bar_a, TC_TILE_BYTES);
            bulk_g2s(smem_u32(sA), D.aform + ((size_t)b * D.tiles + tile) * (TC_TILE_BYTES / 4), TC_TILE_BYTES, bar_a);
            const float *src = D.bform + ((size_t)b * D.rblk + blk0) * (TC_TILE_BYTES / 4);
            for (int i = 0; i < nblk; i++) {
                const int st = i % TC_STAGES;
                if (i >= TC_STAGES) mbar_wait(bar_be + 8 * st, (unsigned)(i / TC_STAGES - 1) & 1u);
                mbar_expect_tx(bar_bf + 8 * st, TC_TILE_BYTES);
                bulk_g2s(smem_u32(sB[st]), src + (size_t)i * (TC_TILE_BYTES / 4), TC_TILE_BYTES, bar_bf + 8 * st);
            }
        }
    } else if (w == 5) {
        if (lane == 0) {  // ---- MMA issuer: D[128 x 128] = A[128 x 16] * B[128 x 16]^T as two K = 8 instructions
            // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128
            constexpr unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
            const unsigned long long adesc = tc_smem_desc(smem_u32(sA));
            mbar_wait(bar_a, 0);
            for (int i = 0; i < nblk; i++) {
                const int st = i % TC_STAGES, acc = i & 1;
                mbar_wait(bar_bf + 8 * st, (unsigned)(i / TC_STAGES) & 1u);
                if (i >= 2) mbar_wait(bar_te + 8 * acc, (unsigned)(i / 2 - 1) & 1u);  // the epilogue drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned long long bdesc = tc_smem_desc(smem_u32(sB[st]));
                const unsigned d = tmem + (unsigned)acc * 128u;
                // descriptor start addresses count 16-byte units: K chunks 2 and 3 start 2 * 2048 B further on
                tc_mma_tf32(d, adesc, bdesc, idesc, 0u);
                tc_mma_tf32(d, adesc + (2 * 2048 >> 4), bdesc + (2 * 2048 >> 4), idesc, 1u);
                tc_commit(bar_be + 8 * st);   // the block's shared-memory stage is free once both MMAs have read it
                tc_commit(bar_tf + 8 * acc);  // ... and the accumulator is complete
            }
        }
    } else {
        // ---- epilogue warps: thread = query = TMEM lane.  Per block four granules of 32 references.
        const int i_q = tile * CS_RB + (int)threadIdx.x;
        float best = PP_INF, second = PP_INF;
        int gran = 0;
        for (int i = 0; i < nblk; i++) {
            const int acc = i & 1;
            mbar_wait(bar_tf + 8 * acc, (unsigned)(i / 2) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned t0 = tmem + ((unsigned)(w * 32) << 16) + (unsigned)acc * 128u;
#pragma unroll
            for (int gi = 0; gi < CS_RB / CS_GR; gi++) {
                float v[32];
                tc_ld32(t0 + gi * CS_GR, v);
                float m0 = fmin3(v[0], v[1], v[2]), m1 = fmin3(v[3], v[4], v[5]);
                float m2 = fmin3(v[6], v[7], v[8]), m3 = fmin3(v[9], v[10], v[11]);
                m0 = fmin3(m0, v[12], v[13]); m1 = fmin3(m1, v[14], v[15]);
                m2 = fmin3(m2, v[16], v[17]); m3 = fmin3(m3, v[18], v[19]);
                m0 = fmin3(m0, v[20], v[21]); m1 = fmin3(m1, v[22], v[23]);
                m2 = fmin3(m2, v[24], v[25]); m3 = fmin3(m3, v[26], v[27]);
                m0 = fmin3(m0, v[28], v[29]); m1 = fmin3(m1, v[30], v[31]);
                const float gm = fminf(fmin3(m0, m1, m2), m3);
                const int gid = (blk0 + i) * (CS_RB / CS_GR) + gi;
                second = fminf(second, fmaxf(best, gm));
                if (gm < best) gran = gid;
                best = fminf(best, gm);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_te + 8 * acc);
        }
        if (i_q < D.nq) {
            const float tau = __uint_as_float(__ldcg(args.r2bits + b)) * args.tau_per_r2;
            const float qq = __fadd_rn(__ldg(&D.qform[(size_t)b * D.npad + i_q].w), tau);
            const unsigned long long key = ((unsigned long long)__float_as_uint(__fadd_rn(best, qq)) << 32) | (unsigned)gran;
            const unsigned sb = __float_as_uint(__fadd_rn(second, qq));
            unsigned long long *K = D.key + (size_t)b * D.nq + i_q;
            unsigned *S = D.sec + (size_t)b * D.nq + i_q;
            if (D.nchunks == 1) {
                *K = key; *S = sb;
            } else {
                const unsigned long long old = atomicMin(K, key);
                atomicMin(S, min((unsigned)((old > key ? old : key) >> 32), sb));
            }
        }
    }
    // every tcgen05 operation of this CTA has completed (the epilogue waited for the last accumulator)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

// ---- exact resolution ------------------------------------------------------------------------------
// One launch for both directions: blocks [0, blocks0) take the points of cloud 1 (dist1 / idx1), the
// rest those of cloud 2.  A warp takes 32 points; for each, its 32 lanes evaluate the 32 references of
// the recorded granule at once (coalesced 384-byte read, REDUX.MIN, ballot, find-first-set).
__global__ void __launch_bounds__(256)
cs_finalize_kernel(const CsArgs args, int blocks0) {
    const int lane = threadIdx.x & 31;
    pdl_wait();
    pdl_launch_dependents();
    const int dir = (int)blockIdx.x >= blocks0 ? 1 : 0;
    const CsDir &D = args.d[dir];
    const long long total = (long long)args.B * D.nq;
    const long long t = (long long)((int)blockIdx.x - (dir ? blocks0 : 0)) * 256 + threadIdx.x;
    bool todo = false;
    int gran = 0, b = 0;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (t < total) {
        const unsigned long long key = __ldcg(D.key + t);
        const unsigned sec = __ldcg(D.sec + t);
        const unsigned vb = (unsigned)(key >> 32);
        gran = (int)(unsigned)key;
        b = (int)(t / D.nq);
        const float tau = __uint_as_float(__ldcg(args.r2bits + b)) * args.tau_per_r2;
        if (sec <= __float_as_uint(__fadd_rn(__uint_as_float(vb), tau))) {
            const unsigned pos = atomicAdd(D.count + b, 1u);
            D.list[(size_t)b * D.nq + pos] = make_uint2((unsigned)(t - (long long)b * D.nq), vb);
        } else {
            todo = true;
            const float *q = D.qxyz + (size_t)t * 3;
            qx = q[0]; qy = q[1]; qz = q[2];
        }
    }
    unsigned want = 0u;
    int found = 0;
    const unsigned livemask = __ballot_sync(FULL_MASK, todo);
#pragma unroll 4
    for (int s = 0; s < 32; s++) {
        if (!((livemask >> s) & 1u)) continue;  // warp-uniform
        const int g_s = __shfl_sync(FULL_MASK, gran, s);
        const int b_s = __shfl_sync(FULL_MASK, b, s);
        const float x_s = __shfl_sync(FULL_MASK, qx, s);
        const float y_s = __shfl_sync(FULL_MASK, qy, s);
        const float z_s = __shfl_sync(FULL_MASK, qz, s);
        const int j = g_s * CS_GR + lane;
        unsigned db = 0xffffffffu;
        if (j < D.nr) {
            // (the squares make the operand order irrelevant: both of the reference's launches give these bits)
            const float *r = D.rxyz + ((size_t)b_s * D.nr + j) * 3;
            db = __float_as_uint(sqdist_xyz(__ldg(r), __ldg(r + 1), __ldg(r + 2), x_s, y_s, z_s));
        }
        const unsigned mn = __reduce_min_sync(FULL_MASK, db);
        const unsigned hit = __ballot_sync(FULL_MASK, db == mn);
        if (lane == s) {
            want = mn;
            found = g_s * CS_GR + __ffs(hit) - 1;  // lowest index among the exact minima
        }
    }
    float s1 = 0.f;
    if (todo) {
        D.dist[t] = __uint_as_float(want);
        D.idx[t] = found;
        s1 = __uint_as_float(want);
        if (args.gw != nullptr)
            bwd_term(__fmul_rn(__ldg(args.gw + dir), 2.f), qx, qy, qz, D.rxyz + ((size_t)b * D.nr + found) * 3,
                     D.gq + (size_t)t * 3, D.gr + ((size_t)b * D.nr + found) * 3);
    }
    if (args.sums != nullptr) {
        __shared__ float sh[8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(FULL_MASK, s1, o);
        if (lane == 0) sh[threadIdx.x >> 5] = s1;
        __syncthreads();
        if (threadIdx.x == 0) {
            float a = 0.f;
#pragma unroll
            for (int i = 0; i < 8; i++) a += sh[i];
            atomicAdd(args.sums + dir, a);
        }
    }
}

// ---- ambiguous points -------------------------------------------------------------------------------
// grid (groups, B, 2).  A CTA takes up to 32 listed points of one cloud, keeps their prepared form and
// thresholds in shared memory and walks the whole partner cloud once: every thread loads a partner
// point and tests it against the 32 entries (3 FFMA + compare each, the sweep's own operation order, so
// the partner that produced the recorded best value passes again); a partner within TAU of an entry's
// best value is evaluated exactly and offered to that entry's (exact value, index) key.
__global__ void __launch_bounds__(256)
cs_rescan_kernel(const CsArgs args) {
    __shared__ float4 sE[CS_RS];                 // {-2x, -2y, -2z, threshold on e}
    __shared__ float sP[CS_RS][3];               // original coordinates
    __shared__ unsigned sI[CS_RS];               // index in its cloud
    __shared__ unsigned long long sK[CS_RS];     // exact value bits << 32 | partner index
    pdl_wait();
    const CsDir &D = args.d[blockIdx.z];
    const int b = blockIdx.y;
    const unsigned n = __ldcg(D.count + b);
    const float tau = __uint_as_float(__ldcg(args.r2bits + b)) * args.tau_per_r2;
    const float *rf = D.rform + (size_t)b * D.rblk * (4 * CS_RB);
    const float *rx = D.rxyz + (size_t)b * D.nr * 3;
    for (unsigned e0 = blockIdx.x * CS_RS; e0 < n; e0 += gridDim.x * CS_RS) {
        const int ne = (int)min((unsigned)CS_RS, n - e0);
        __syncthreads();
        if ((int)threadIdx.x < CS_RS) {
            float4 v = make_float4(0.f, 0.f, 0.f, -PP_INF);  // unused slot: nothing passes
            if ((int)threadIdx.x < ne) {
                const uint2 ent = D.list[(size_t)b * D.nq + e0 + threadIdx.x];
                const float4 q = D.qform[(size_t)b * D.npad + ent.x];
                // recorded value = rn(e_best + qq), qq = rn(|q|^2 + TAU).  Candidates: rn(e + qq) <= value + TAU.
                // Tested as e <= thr_e with thr_e rounded up generously (a superset is harmless).
                const float qq = __fadd_rn(q.w, tau);
                const float lim = __fadd_rn(__uint_as_float(ent.y), __uint_as_float(__ldcg(args.r2bits + b)) * args.rescan_per_r2);
                float thr_e = __fsub_ru(lim, qq);
                thr_e = __fadd_ru(thr_e, fmaxf(fabsf(lim), fabsf(qq)) * 2.4e-7f);
                v = make_float4(q.x, q.y, q.z, thr_e);
                const float *p = D.qxyz + ((size_t)b * D.nq + ent.x) * 3;
                sP[threadIdx.x][0] = p[0]; sP[threadIdx.x][1] = p[1]; sP[threadIdx.x][2] = p[2];
                sI[threadIdx.x] = ent.x;
            }
            sE[threadIdx.x] = v;
            sK[threadIdx.x] = 0xffffffffffffffffull;
        }
        __syncthreads();
        for (int j = threadIdx.x; j < D.nr; j += 256) {
            const float *blk = rf + (size_t)(j / CS_RB) * (4 * CS_RB) + (j % CS_RB);
            const float x = blk[0], y = blk[CS_RB], z = blk[2 * CS_RB], rr = blk[3 * CS_RB];
            unsigned hits = 0u;
#pragma unroll 8
            for (int e = 0; e < CS_RS; e++) {
                const float4 q = sE[e];
                hits |= (approx_e(x, y, z, rr, q.x, q.y, q.z) <= q.w ? 1u : 0u) << e;
            }
            while (hits) {
                const int e = __ffs(hits) - 1;
                hits &= hits - 1;
                const float d = sqdist_xyz(__ldg(rx + (size_t)j * 3), __ldg(rx + (size_t)j * 3 + 1),
                                           __ldg(rx + (size_t)j * 3 + 2), sP[e][0], sP[e][1], sP[e][2]);
                atomicMin(&sK[e], ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)j);
            }
        }
        __syncthreads();
        if ((int)threadIdx.x < ne) {
            const unsigned long long k = sK[threadIdx.x];
            const size_t t = (size_t)b * D.nq + sI[threadIdx.x];
            const float d = __uint_as_float((unsigned)(k >> 32));
            const int found = (int)(unsigned)k;
            D.dist[t] = d;
            D.idx[t] = found;
            if (found >= 0 && found < D.nr) {  // always, by the error bound; never index out of range
                if (args.sums != nullptr) atomicAdd(args.sums + blockIdx.z, d);
                if (args.gw != nullptr)
                    bwd_term(__fmul_rn(__ldg(args.gw + blockIdx.z), 2.f), sP[threadIdx.x][0], sP[threadIdx.x][1],
                             sP[threadIdx.x][2], rx + (size_t)found * 3, D.gq + t * 3,
                             D.gr + ((size_t)b * D.nr + found) * 3);
            }
        }
    }
}

}  // namespace

size_t chamfer_sweep_workspace_bytes(int B, int N, int M) { return cs_layout(B, N, M, true).total; }

// gw / g1 / g2 != nullptr: fused uniform backward.  `sums` has been cleared by the caller.
int chamfer_sweep_launch(const float *xyz1, const float *xyz2, int B, int N, int M, float *dist1, float *dist2,
                         int *idx1, int *idx2, float *sums, void *workspace, size_t workspace_bytes,
                         const float *gw, float *g1, float *g2, cudaStream_t st, bool tc) {
    const CsLayout L = cs_layout(B, N, M, tc);
    if (workspace_bytes < L.total) {
        set_error("chamfer_fwd: workspace %zu < %zu bytes", workspace_bytes, L.total);
        return PP_ENOSPC;
    }
    PP_REQUIRE(B <= 65535, "chamfer: B=%d too large", B);
    PP_REQUIRE((long long)B * N < (1ll << 31) && (long long)B * M < (1ll << 31), "chamfer: B*N too large");
    char *ws = (char *)workspace;
    unsigned *r2bits = (unsigned *)(ws + L.ctrl);
    CsArgs A;
    A.r2bits = r2bits; A.sums = sums; A.gw = gw; A.B = B;
    A.tau_per_r2 = tc ? CS_TAU_PER_R2_TC : CS_TAU_PER_R2;
    // FFMA sweep: the rescan recomputes the sweep's own values -> the same window.  Tensor-core sweep: the
    // rescan's FFMA value of the true minimiser may sit EPS_tc + EPS_ffma above its tensor-core value, which
    // itself is within 2 EPS_tc of the recorded best: 3 * 128 u + 64 u, rounded up to 560 u.
    A.rescan_per_r2 = tc ? 560.f * 5.9604644775390625e-8f : CS_TAU_PER_R2;
    const int n[2] = {N, M};
    const float *xyz[2] = {xyz1, xyz2};
    float *dist[2] = {dist1, dist2};
    int *idx[2] = {idx1, idx2};
    float *g[2] = {gw ? g1 : nullptr, gw ? g2 : nullptr};
    // 2-warp CTAs when 4-warp CTAs would leave warp slots of the last tile idle (a warp takes 256 queries)
    int warps = get_option("chamfer_sweep_warps", 0);
    if (warps != 2 && warps != 4) {
        const int w4 = ceil_div(ceil_div(N, CS_WT), 4) * 4 + ceil_div(ceil_div(M, CS_WT), 4) * 4;
        const int w2 = ceil_div(ceil_div(N, CS_WT), 2) * 2 + ceil_div(ceil_div(M, CS_WT), 2) * 2;
        warps = w2 < w4 ? 2 : 4;
    }
    // enough CTAs for several waves: split the reference cloud into chunks when the tiles alone are too few
    const long long want_ctas = (long long)NUM_SMS_B200 * get_option("chamfer_sweep_ctas_per_sm", warps == 4 ? 10 : 20);
    int grid_x = 1;
    for (int s = 0; s < 2; s++) {
        CsDir &D = A.d[s];
        D.qxyz = xyz[s]; D.rxyz = xyz[1 - s];
        D.qform = (const float4 *)(ws + L.qform[s]);
        D.rform = (const float *)(ws + L.rform[1 - s]);
        D.aform = (const float *)(ws + L.aform[s]);
        D.bform = (const float *)(ws + L.bform[1 - s]);
        D.key = (unsigned long long *)(ws + L.key[s]);
        D.sec = (unsigned *)(ws + L.sec[s]);
        D.list = (uint2 *)(ws + L.list[s]);
        D.count = r2bits + B * (1 + s);
        D.dist = dist[s]; D.idx = idx[s];
        D.gq = g[s]; D.gr = g[1 - s];
        D.nq = n[s]; D.nr = n[1 - s];
        D.npad = L.npad[s]; D.rblk = L.blk[1 - s];
        D.tiles = tc ? ceil_div(n[s], CS_RB) : ceil_div(n[s], CS_WT * warps);
        const long long base = 2ll * B * D.tiles;
        int chunks = (int)ceil_div_ll(want_ctas, base);
        const int max_chunks = max(1, D.rblk / 4);  // at least four blocks (512 references) per chunk
        if (chunks > max_chunks) chunks = max_chunks;
        if (chunks < 1) chunks = 1;
        D.chunk_blocks = ceil_div(D.rblk, chunks);
        D.nchunks = ceil_div(D.rblk, D.chunk_blocks);
        grid_x = max(grid_x, D.tiles * D.nchunks);
    }

    PP_CUDA(cudaMemsetAsync(r2bits, 0, 4 * 3 * (size_t)B, st));
    {
        KernelTimer timer("chamfer_prep", st);
        const int per_cloud = max(L.npad[0], L.npad[1]);
        int chunks = ceil_div(per_cloud, 256 * 4);
        const int want = ceil_div(NUM_SMS_B200, B);  // at least two CTAs per SM over the whole grid
        if (chunks < want) chunks = min(want, ceil_div(per_cloud, 256));
        cs_prep_kernel<<<dim3(chunks, B, 2), 256, 0, st>>>(
            xyz1, xyz2, N, M, (float4 *)(ws + L.qform[0]), (float4 *)(ws + L.qform[1]), (float *)(ws + L.rform[0]),
            (float *)(ws + L.rform[1]), (unsigned long long *)(ws + L.key[0]), (unsigned long long *)(ws + L.key[1]),
            (unsigned *)(ws + L.sec[0]), (unsigned *)(ws + L.sec[1]), r2bits, g[0], g[1]);
        PP_LAUNCH_CHECK();
    }
    if (tc) {
        KernelTimer timer("chamfer_prep_tc", st);
        const int rows = max(L.blk[0], L.blk[1]) * CS_RB;
        PP_CUDA(launch_pdl(cs_prep_tc_kernel, dim3(ceil_div(rows, 256), B, 2), dim3(256), 0, st,
                           (const float4 *)(ws + L.qform[0]), (const float4 *)(ws + L.qform[1]), N, M,
                           (float *)(ws + L.aform[0]), (float *)(ws + L.aform[1]), (float *)(ws + L.bform[0]),
                           (float *)(ws + L.bform[1])));
        PP_LAUNCH_CHECK();
    }
    {
        KernelTimer timer("chamfer_fwd", st);
        if (tc)
            PP_CUDA(launch_pdl(cs_rowpass_tc_kernel, dim3(grid_x, B, 2), dim3(192), 0, st, A));
        else if (warps == 4)
            PP_CUDA(launch_pdl(cs_rowpass_kernel<4>, dim3(grid_x, B, 2), dim3(128), 0, st, A));
        else
            PP_CUDA(launch_pdl(cs_rowpass_kernel<2>, dim3(grid_x, B, 2), dim3(64), 0, st, A));
        PP_LAUNCH_CHECK();
    }
    {
        KernelTimer timer("chamfer_finalize", st);
        const int blocks0 = (int)ceil_div_ll((long long)B * N, 256), blocks1 = (int)ceil_div_ll((long long)B * M, 256);
        PP_CUDA(launch_pdl(cs_finalize_kernel, dim3(blocks0 + blocks1), dim3(256), 0, st, A, blocks0));
        PP_LAUNCH_CHECK();
    }
    {
        KernelTimer timer("chamfer_rescan", st);
        const int groups = max(1, min(64, ceil_div(2 * NUM_SMS_B200, B)));
        PP_CUDA(launch_pdl(cs_rescan_kernel, dim3(groups, B, 2), dim3(256), 0, st, A));
        PP_LAUNCH_CHECK();
    }
    return PP_OK;
}

}  // namespace pp
