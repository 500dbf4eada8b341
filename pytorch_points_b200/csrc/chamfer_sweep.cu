// chamfer_sweep.cu -- Chamfer / nndistance forward for c == 3 clouds: an approximate one-pass
// sweep on the packed FP32 pipe followed by an exact resolution of the few surviving candidates.
//
// Replaces NmDistanceKernel x2 (_ext/nmdistance_cuda.cu:8-49,127-128).  The reference's distance
//      d = fma(tz,tz, fma(ty,ty, rn(tx*tx))),  t = rn(ref - query)                    (6 lane-ops)
// has to be reproduced bit for bit, and so does its lowest-index tie rule -- but only for the pair
// that WINS.  Every other pair merely has to be shown to lose.  So the sweep evaluates the
// expansion  |r|^2 - 2 q.r + |q|^2  on centred coordinates (3 FFMA + 1 FADD per pair: the
// "GEMM-expansion distance path" on the FFMA pipe), whose error against the reference's value is
// bounded by EPS = 64 u R^2 (u = 2^-24, R = largest centred norm; derivation in DESIGN.md §3.1),
// and keeps, per point, the best approximate value, where it occurred (a 32-reference granule for
// rows, an 8-query lane group for columns) and the runner-up value of every OTHER granule/group.
//   runner-up > best + TAU (TAU = 2.5 EPS)  =>  every true minimiser (ties included) lies inside
//       the recorded granule/group: it is re-evaluated with the exact chain (32 resp. 8 pairs);
//   otherwise the point is "ambiguous" (1-2 % of uniform clouds, every point of a lattice): a
//       rescan kernel walks all partners, exact-evaluates those within TAU and takes the lowest
//       index among the exact minima.
// Results are therefore identical to the reference's on every input; only the time depends on
// how many points are ambiguous.
//
// Kernels (one stream, programmatic dependent launch):
//   cs_prep_kernel      centre (mean of the leading points), centred coordinates, norms, R^2;
//                       queries as float4 {-2x,-2y,-2z,|q|^2}, references as SoA blocks of 128
//                       {x[128],y[128],z[128],|r|^2[128]}; resets keys / gradients / counters.
//   cs_sweep_kernel     one CTA per (cloud, 128-reference block); each warp streams 256-query
//                       tiles through its own double-buffered shared-memory slots with
//                       cp.async.bulk + mbarrier (no CTA barrier in the loop) and sweeps them past
//                       the resident block.  Rows: per query the minima of the block's four
//                       granules, pushed to the global (best, granule) key / runner-up word only
//                       when they pass the row's current threshold.  Columns: per-warp
//                       (best, group, runner-up) records in shared memory behind a shared filter;
//                       the CTA sees every query of its references, so it resolves its columns
//                       itself at the end (dist2 / idx2, loss sum, fused backward).
//   cs_finalize_rows_kernel   exact resolution of the rows (dist1 / idx1, loss sum, fused backward).
//   cs_rescan_kernel    the ambiguous rows and columns.
#include "pp_common.cuh"

namespace pp {
namespace {

constexpr int CS_RB = 128;   // references per block
constexpr int CS_WT = 256;   // queries per warp tile (8 per lane, lane-interleaved)
constexpr int CS_GR = 32;    // row-side granule (references)
constexpr unsigned CS_INF_BITS = 0x7f800000u;
constexpr unsigned long long CS_KEY_INIT = 0x7f800000ffffffffull;
// TAU = 2.5 * EPS, EPS = 64 u R^2, u = 2^-24
constexpr float CS_TAU_PER_R2 = 160.f * 5.9604644775390625e-8f;

struct CsLayout {
    size_t ctrl, prepq, prepr, key1, sec1, rowlist, collist, total;
    int npad, mblk;
};

CsLayout cs_layout(int B, int N, int M) {
    CsLayout L;
    L.npad = ceil_div(N, CS_WT) * CS_WT;
    L.mblk = ceil_div(M, CS_RB);
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t o = 0;
    L.ctrl = o;    o += up(64 + 4 * (size_t)B);
    L.prepq = o;   o += up(16 * (size_t)B * L.npad);
    L.prepr = o;   o += up(16 * (size_t)B * L.mblk * CS_RB);
    L.key1 = o;    o += up(8 * (size_t)B * N);
    L.sec1 = o;    o += up(4 * (size_t)B * N);
    L.rowlist = o; o += up(8 * (size_t)B * N);
    L.collist = o; o += up(8 * (size_t)B * M);
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
    asm("{\n\t.reg .b64 ra, rq, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rq, {%4, %4};\n\tmov.b64 rc, {%5, %6};\n\t"
        "fma.rn.f32x2 rd, ra, rq, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(q), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 add2_bcast(float2 a, float q) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rq, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rq, {%4, %4};\n\t"
        "add.rn.f32x2 rd, ra, rq;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(q));
    return d;
}

// the approximate value both sides order by: f = fma(z,qz', fma(y,qy', fma(x,qx', |r|^2))) + (|q|^2 + TAU)
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
__global__ void __launch_bounds__(256)
cs_prep_kernel(const float *__restrict__ xyz1, const float *__restrict__ xyz2, int N, int M, int npad, int mblk,
               float4 *__restrict__ prepq, float *__restrict__ prepr, unsigned long long *__restrict__ key1,
               unsigned *__restrict__ sec1, unsigned *__restrict__ r2bits, float *__restrict__ g1,
               float *__restrict__ g2) {
    pdl_launch_dependents();
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    __shared__ float s_c[3];
    if (threadIdx.x < 32) {
        // centre = mean of the (up to) 32 leading points of each cloud: the same instruction
        // sequence in every CTA of this cloud, hence the same bits
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
    float r2 = 0.f;
    const int stride = gridDim.x * blockDim.x;
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = t0; i < npad; i += stride) {
        float4 v = make_float4(0.f, 0.f, 0.f, PP_INF);  // padding: never a column minimum
        if (i < N) {
            const size_t t = (size_t)b * N + i;
            const float *p = xyz1 + t * 3;
            const float x = __fsub_rn(p[0], cx), y = __fsub_rn(p[1], cy), z = __fsub_rn(p[2], cz);
            const float n = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
            v = make_float4(-2.f * x, -2.f * y, -2.f * z, n);
            r2 = fmaxf(r2, n);
            key1[t] = CS_KEY_INIT;
            sec1[t] = CS_INF_BITS;
            if (g1 != nullptr) { g1[t * 3 + 0] = 0.f; g1[t * 3 + 1] = 0.f; g1[t * 3 + 2] = 0.f; }
        }
        prepq[(size_t)b * npad + i] = v;
    }
    for (int j = t0; j < mblk * CS_RB; j += stride) {
        float x = 0.f, y = 0.f, z = 0.f, n = PP_INF;  // padding: never a row minimum
        if (j < M) {
            const size_t t = (size_t)b * M + j;
            const float *p = xyz2 + t * 3;
            x = __fsub_rn(p[0], cx); y = __fsub_rn(p[1], cy); z = __fsub_rn(p[2], cz);
            n = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
            r2 = fmaxf(r2, n);
            if (g2 != nullptr) { g2[t * 3 + 0] = 0.f; g2[t * 3 + 1] = 0.f; g2[t * 3 + 2] = 0.f; }
        }
        float *blk = prepr + ((size_t)b * mblk + j / CS_RB) * (4 * CS_RB);
        const int o = j % CS_RB;
        blk[o] = x; blk[CS_RB + o] = y; blk[2 * CS_RB + o] = z; blk[3 * CS_RB + o] = n;
    }
    const unsigned rb = __reduce_max_sync(FULL_MASK, __float_as_uint(r2));  // r2 >= 0: bits order like values
    if (lane == 0 && rb != 0u) atomicMax(r2bits + b, rb);
}

// ---- the sweep ------------------------------------------------------------------------------------
struct CsOut {
    float *dist2;
    int *idx2;
    float *sums;
    const float *gw;
    float *g1, *g2;
    unsigned *ctrl;
    uint2 *collist;
};

constexpr int CS_CAP = 384;  // candidate entries a warp can park between two drains

template <int WARPS>
struct CsSmem {  // byte offsets into the sweep kernel's dynamic shared memory
    static constexpr int Q = 0;
    static constexpr int REF = Q + WARPS * 2 * CS_WT * 16;
    static constexpr int BUF = REF + 4 * CS_RB * 4;
    static constexpr int KEY = BUF + WARPS * CS_CAP * 8;
    static constexpr int BAR = KEY + CS_RB * 8;
    static constexpr int THR = BAR + (1 + 2 * WARPS) * 8 + 8;  // keeps 16-byte alignment below
    static constexpr int W = THR + WARPS * CS_WT * 4;
    static constexpr int SEC = W + CS_RB * 4;
    static constexpr int SUM = SEC + CS_RB * 4;
    static constexpr int BYTES = SUM + WARPS * 4;
};

// Column candidates a warp parked during one tile -> the CTA's (best, group) key and runner-up word.
// entry.x = value bits, entry.y = column << 25 | group
__device__ __noinline__ void cs_drain(const uint2 *buf, int cnt, unsigned long long *sKey, unsigned *sSec, int lane) {
    for (int e = lane; e < cnt; e += 32) {
        const uint2 en = buf[e];
        const int j = (int)(en.y >> 25);
        const unsigned long long key = ((unsigned long long)en.x << 32) | (en.y & 0x1ffffffu);
        const unsigned long long old = atomicMin(sKey + j, key);
        atomicMin(sSec + j, (unsigned)((old > key ? old : key) >> 32));  // the loser of every comparison
    }
    __syncwarp();
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 4 ? 4 : 8)
cs_sweep_kernel(const float *__restrict__ xyz1, const float *__restrict__ xyz2, int N, int M, int npad, int mblk,
                const float4 *__restrict__ prepq, const float *__restrict__ prepr,
                unsigned long long *__restrict__ key1, unsigned *__restrict__ sec1,
                const unsigned *__restrict__ r2bits, CsOut out) {
    constexpr int THREADS = WARPS * 32;
    // dynamic shared memory (above the 48 KB static limit), carved by hand; see CsSmem
    extern __shared__ __align__(128) unsigned char cs_smem[];
    typedef CsSmem<WARPS> S;
    float4(*sQ)[2][CS_WT] = reinterpret_cast<float4(*)[2][CS_WT]>(cs_smem + S::Q);   // per-warp double-buffered query tiles
    float *sRef = reinterpret_cast<float *>(cs_smem + S::REF);                        // x | y | z | |r|^2
    uint2(*sBuf)[CS_CAP] = reinterpret_cast<uint2(*)[CS_CAP]>(cs_smem + S::BUF);      // per-warp parked column candidates
    unsigned long long *sKey = reinterpret_cast<unsigned long long *>(cs_smem + S::KEY);  // column record: value bits << 32 | group
    unsigned long long *sBar = reinterpret_cast<unsigned long long *>(cs_smem + S::BAR);
    unsigned(*sThr)[CS_WT] = reinterpret_cast<unsigned(*)[CS_WT]>(cs_smem + S::THR);  // per-warp row thresholds of the current tile
    unsigned *sW = reinterpret_cast<unsigned *>(cs_smem + S::W);                      // column filter: bits(best so far + TAU)
    unsigned *sSec = reinterpret_cast<unsigned *>(cs_smem + S::SEC);                  // column record: runner-up value bits
    float *sSum = reinterpret_cast<float *>(cs_smem + S::SUM);

    pdl_launch_dependents();
    const int b = blockIdx.x, blk = blockIdx.y;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const float *sX = sRef, *sY = sRef + CS_RB, *sZ = sRef + 2 * CS_RB, *sR = sRef + 3 * CS_RB;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < 1 + 2 * WARPS; i++) mbar_init(smem_u32(sBar + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    for (int t = threadIdx.x; t < CS_RB; t += THREADS) {
        sKey[t] = 0xffffffffffffffffull;
        sSec[t] = 0xffffffffu;
    }
    __syncthreads();
    pdl_wait();  // the prepared arrays, R^2 and the reset keys are complete and visible

    const float tau = __uint_as_float(__ldcg(r2bits + b)) * CS_TAU_PER_R2;
    const int nwt = npad / CS_WT;
    // warp w takes the tiles w, w + WARPS, ...; CTAs of different reference blocks start at
    // different tiles, so a row meets its reference blocks one after the other and the row
    // threshold it reads has already been tightened by the earlier ones
    const int cnt = nwt > w ? (nwt - w + WARPS - 1) / WARPS : 0;
    const int rot = cnt > 0 ? blk % cnt : 0;
    const float4 *qsrc = prepq + (size_t)b * npad;
    const unsigned bar_ref = smem_u32(sBar);
    const unsigned bar_q0 = smem_u32(sBar + 1 + 2 * w), bar_q1 = smem_u32(sBar + 2 + 2 * w);
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar_ref, 16 * CS_RB);
        bulk_g2s(smem_u32(sRef), prepr + ((size_t)b * mblk + blk) * (4 * CS_RB), 16 * CS_RB, bar_ref);
    }
    if (cnt > 0 && lane == 0) {
        const int wt = w + WARPS * rot;
        mbar_expect_tx(bar_q0, 16 * CS_WT);
        bulk_g2s(smem_u32(&sQ[w][0][0]), qsrc + (size_t)wt * CS_WT, 16 * CS_WT, bar_q0);
    }
    mbar_wait(bar_ref, 0);
    // ---- filter seed: every reference against the leading queries of warp 0's first tile, so that the
    // sweep starts with a finite threshold (with +inf every lane of the first visit would be a candidate)
    mbar_wait(smem_u32(sBar + 1), 0);
    for (int t = threadIdx.x; t < CS_RB; t += THREADS) {
        const float x = sX[t], y = sY[t], z = sZ[t], rr = sR[t];
        float m = PP_INF;
#pragma unroll 8
        for (int i = 0; i < 128; i++) {
            const float4 q = sQ[0][0][i];  // broadcast read; padding queries carry |q|^2 = +inf
            m = fminf(m, __fadd_rn(approx_e(x, y, z, rr, q.x, q.y, q.z), __fadd_rn(q.w, tau)));
        }
        // padding references (|r|^2 = +inf) get a filter nothing passes
        sW[t] = blk * CS_RB + t < M ? __float_as_uint(__fadd_rn(m, tau)) : 0u;
    }
    __syncthreads();

    unsigned long long *k1 = key1 + (size_t)b * N;
    unsigned *s1 = sec1 + (size_t)b * N;

#pragma unroll 1
    for (int it = 0; it < cnt; it++) {
        int kk = it + rot;
        if (kk >= cnt) kk -= cnt;
        const int wt = w + WARPS * kk;
        const int s = it & 1;
        if (it + 1 < cnt && lane == 0) {  // prefetch the next tile into the other slot
            int kn = it + 1 + rot;
            if (kn >= cnt) kn -= cnt;
            const unsigned bar_n = s ? bar_q0 : bar_q1;
            mbar_expect_tx(bar_n, 16 * CS_WT);
            bulk_g2s(smem_u32(&sQ[w][s ^ 1][0]), qsrc + (size_t)(w + WARPS * kn) * CS_WT, 16 * CS_WT, bar_n);
        }
        const int ibase = wt * CS_WT + lane;
        // the rows' current thresholds (upper word of their keys) travel to shared memory while the
        // tile is swept (LDGSTS); they are read at the end of the tile
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int i = ibase + q * 32;
            if (i < N)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&sThr[w][q * 32 + lane])),
                             "l"(reinterpret_cast<const unsigned *>(k1 + i) + 1)
                             : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        mbar_wait(s ? bar_q1 : bar_q0, (unsigned)(it >> 1) & 1u);
        float qx[8], qy[8], qz[8], qq[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const float4 v = sQ[w][s][q * 32 + lane];
            qx[q] = v.x; qy[q] = v.y; qz[q] = v.z;
            qq[q] = __fadd_rn(v.w, tau);  // + TAU keeps every ordered value positive (uint ordering)
        }
        __syncwarp();  // the slot may be refilled from the next iteration on
        const unsigned group = (unsigned)wt * 32u + (unsigned)lane;
        int parked = 0;  // warp-uniform: entries in sBuf[w]

        float g[4][8];  // per granule and query: min over the granule of e = f - qq
#pragma unroll
        for (int sgi = 0; sgi < 4; sgi++) {
            // warps sweep the block's four granules in rotated order: a column meets the CTA's warps one
            // after the other and the filter a warp reads already holds what its predecessors found
            const int sg = (sgi + w) & 3;
#pragma unroll
            for (int q = 0; q < 8; q++) g[sgi][q] = PP_INF;
#pragma unroll 1
            for (int st = 0; st < CS_GR; st += 4) {
                const int jj = sg * CS_GR + st;
                const float4 X = *reinterpret_cast<const float4 *>(sX + jj);
                const float4 Y = *reinterpret_cast<const float4 *>(sY + jj);
                const float4 Z = *reinterpret_cast<const float4 *>(sZ + jj);
                const float4 R = *reinterpret_cast<const float4 *>(sR + jj);
                const float2 x01 = make_float2(X.x, X.y), x23 = make_float2(X.z, X.w);
                const float2 y01 = make_float2(Y.x, Y.y), y23 = make_float2(Y.z, Y.w);
                const float2 z01 = make_float2(Z.x, Z.y), z23 = make_float2(Z.z, Z.w);
                const float2 r01 = make_float2(R.x, R.y), r23 = make_float2(R.z, R.w);
                float c0 = PP_INF, c1 = PP_INF, c2 = PP_INF, c3 = PP_INF;
#pragma unroll
                for (int q = 0; q < 8; q += 2) {
                    const float2 a01 = fma2_bcast(z01, qz[q], fma2_bcast(y01, qy[q], fma2_bcast(x01, qx[q], r01)));
                    const float2 a23 = fma2_bcast(z23, qz[q], fma2_bcast(y23, qy[q], fma2_bcast(x23, qx[q], r23)));
                    const float2 b01 = fma2_bcast(z01, qz[q + 1], fma2_bcast(y01, qy[q + 1], fma2_bcast(x01, qx[q + 1], r01)));
                    const float2 b23 = fma2_bcast(z23, qz[q + 1], fma2_bcast(y23, qy[q + 1], fma2_bcast(x23, qx[q + 1], r23)));
                    g[sgi][q] = fmin3(fmin3(g[sgi][q], a01.x, a01.y), a23.x, a23.y);
                    g[sgi][q + 1] = fmin3(fmin3(g[sgi][q + 1], b01.x, b01.y), b23.x, b23.y);
                    const float2 fa01 = add2_bcast(a01, qq[q]), fa23 = add2_bcast(a23, qq[q]);
                    const float2 fb01 = add2_bcast(b01, qq[q + 1]), fb23 = add2_bcast(b23, qq[q + 1]);
                    c0 = fmin3(c0, fa01.x, fb01.x);
                    c1 = fmin3(c1, fa01.y, fb01.y);
                    c2 = fmin3(c2, fa23.x, fb23.x);
                    c3 = fmin3(c3, fa23.y, fb23.y);
                }
                // ---- column side: a lane whose group minimum passes the filter parks (value, column,
                // group) in the warp's buffer -- no reduction, no election; the buffer is merged into the
                // CTA's column records at the end of the tile.  The filter words are updated concurrently
                // by the other warps; every value ever stored is an observed value + TAU, hence
                // >= final minimum + TAU: a stale or lost update only lets extra candidates through.
                const uint4 W = *reinterpret_cast<const uint4 *>(sW + jj);
                const bool p0 = __float_as_uint(c0) <= W.x, p1 = __float_as_uint(c1) <= W.y;
                const bool p2 = __float_as_uint(c2) <= W.z, p3 = __float_as_uint(c3) <= W.w;
                if (__any_sync(FULL_MASK, p0 | p1 | p2 | p3)) {
                    const bool pp_[4] = {p0, p1, p2, p3};
                    const float cc[4] = {c0, c1, c2, c3};
                    const unsigned ww[4] = {W.x, W.y, W.z, W.w};
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const unsigned mask = __ballot_sync(FULL_MASK, pp_[r]);
                        if (mask != 0u) {
                            if (pp_[r]) {
                                const int j = jj + r;
                                const int slot = parked + __popc(mask & lt_mask);
                                if (slot < CS_CAP)
                                    sBuf[w][slot] = make_uint2(__float_as_uint(cc[r]), ((unsigned)j << 25) | group);
                                else
                                    sSec[j] = 0u;  // no room: the column goes to the rescan
                                const unsigned nt = __float_as_uint(__fadd_rn(cc[r], tau));
                                if (nt < ww[r]) sW[j] = nt;
                            }
                            parked += __popc(mask);
                        }
                    }
                }
            }
        }
        if (parked > 0) cs_drain(sBuf[w], min(parked, CS_CAP), sKey, sSec, lane);
        // ---- row side: the tile's four granule minima of each query against the row's threshold ----
        asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int i = ibase + q * 32;
            const float m = fminf(fmin3(g[0][q], g[1][q], g[2][q]), g[3][q]);
            const unsigned vb = __float_as_uint(__fadd_rn(m, qq[q]));
            const bool pass = i < N && vb <= __float_as_uint(__fadd_rn(__uint_as_float(sThr[w][q * 32 + lane]), tau));
            if (pass) {
                // a granule holding the tile minimum, runner-up among the other three (two granules
                // holding it => runner-up == minimum => the row is ambiguous and goes to the rescan)
                int a = 3;
                float second = fmin3(g[0][q], g[1][q], g[2][q]);
                if (g[2][q] == m) { a = 2; second = fmin3(g[0][q], g[1][q], g[3][q]); }
                if (g[1][q] == m) { a = 1; second = fmin3(g[0][q], g[2][q], g[3][q]); }
                if (g[0][q] == m) { a = 0; second = fmin3(g[1][q], g[2][q], g[3][q]); }
                const unsigned long long key =
                    ((unsigned long long)vb << 32) | (unsigned)(blk * (CS_RB / CS_GR) + ((a + w) & 3));
                const unsigned long long old = atomicMin(k1 + i, key);
                const unsigned loser = (unsigned)((old > key ? old : key) >> 32);
                const unsigned sb = __float_as_uint(__fadd_rn(second, qq[q]));  // +inf stays +inf
                atomicMin(s1 + i, min(loser, sb));
            }
        }
        __syncwarp();  // sThr[w] is rewritten by the next iteration's copies
    }

    // ---- columns: this CTA has seen every query of its 128 references ----
    __syncthreads();
    float s2 = 0.f;
    for (int t = threadIdx.x; t < CS_RB; t += THREADS) {
        const int j = blk * CS_RB + t;
        if (j >= M) continue;
        const unsigned long long key = sKey[t];
        const unsigned best = (unsigned)(key >> 32), grp = (unsigned)key;
        const size_t u = (size_t)b * M + j;
        // A column whose candidates did not all fit the buffer carries runner-up 0; its recorded best is
        // still an observed value, so best + TAU bounds the minimum (+inf if nothing was recorded at all).
        if (key == 0xffffffffffffffffull || sSec[t] <= __float_as_uint(__fadd_rn(__uint_as_float(best), tau))) {
            const unsigned pos = atomicAdd(out.ctrl + 1, 1u);
            out.collist[pos] = make_uint2((unsigned)u, key == 0xffffffffffffffffull ? CS_INF_BITS : best);
            continue;
        }
        const float *r = xyz2 + u * 3;
        const float rx = __ldg(r), ry = __ldg(r + 1), rz = __ldg(r + 2);
        const float *q = xyz1 + (size_t)b * N * 3;
        const int i0 = (int)(grp >> 5) * CS_WT + (int)(grp & 31u);
        float bd = PP_INF;
        int bi = i0;
#pragma unroll
        for (int e = 0; e < 8; e++) {  // ascending index, strict '<': lowest index on ties
            const int i = i0 + e * 32;
            if (i < N) {
                // same operand roles as the reference's second launch: (cloud-1 point) - (cloud-2 point)
                const float d = sqdist_xyz(__ldg(q + (size_t)i * 3), __ldg(q + (size_t)i * 3 + 1),
                                           __ldg(q + (size_t)i * 3 + 2), rx, ry, rz);
                if (d < bd) { bd = d; bi = i; }
            }
        }
        out.dist2[u] = bd;
        out.idx2[u] = bi;
        s2 += bd;
        if (out.gw != nullptr)
            bwd_term(__fmul_rn(__ldg(out.gw + 1), 2.f), rx, ry, rz, xyz1 + ((size_t)b * N + bi) * 3, out.g2 + u * 3,
                     out.g1 + ((size_t)b * N + bi) * 3);
    }
    if (out.sums != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(FULL_MASK, s2, o);
        if (lane == 0) sSum[w] = s2;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
#pragma unroll
            for (int ww = 0; ww < WARPS; ww++) t += sSum[ww];
            atomicAdd(out.sums + 1, t);
        }
    }
}

// ---- rows: exact resolution --------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cs_finalize_rows_kernel(const float *__restrict__ xyz1, const float *__restrict__ xyz2, int B, int N, int M,
                        const unsigned long long *__restrict__ key1, const unsigned *__restrict__ sec1,
                        const unsigned *__restrict__ r2bits, float *__restrict__ dist1, int *__restrict__ idx1,
                        float *__restrict__ sums, const float *__restrict__ gw, float *__restrict__ g1,
                        float *__restrict__ g2, unsigned *__restrict__ ctrl, uint2 *__restrict__ rowlist) {
    const int lane = threadIdx.x & 31;
    pdl_wait();
    pdl_launch_dependents();
    const long long total = (long long)B * N;
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    bool todo = false;
    int gran = 0, b = 0;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (t < total) {
        const unsigned long long key = __ldcg(key1 + t);
        const unsigned sec = __ldcg(sec1 + t);
        const unsigned vb = (unsigned)(key >> 32);
        gran = (int)(unsigned)key;
        b = (int)(t / N);
        const float tau = __uint_as_float(__ldcg(r2bits + b)) * CS_TAU_PER_R2;
        if (sec <= __float_as_uint(__fadd_rn(__uint_as_float(vb), tau))) {
            const unsigned pos = atomicAdd(ctrl + 0, 1u);
            rowlist[pos] = make_uint2((unsigned)t, vb);
        } else {
            todo = true;
            const float *q = xyz1 + (size_t)t * 3;
            qx = q[0]; qy = q[1]; qz = q[2];
        }
    }
    unsigned want = 0u;
    int found = 0;
    const unsigned live = __ballot_sync(FULL_MASK, todo);
#pragma unroll 4
    for (int s = 0; s < 32; s++) {
        if (!((live >> s) & 1u)) continue;  // warp-uniform
        const int g_s = __shfl_sync(FULL_MASK, gran, s);
        const int b_s = __shfl_sync(FULL_MASK, b, s);
        const float x_s = __shfl_sync(FULL_MASK, qx, s);
        const float y_s = __shfl_sync(FULL_MASK, qy, s);
        const float z_s = __shfl_sync(FULL_MASK, qz, s);
        const int j = g_s * CS_GR + lane;
        unsigned db = CS_INF_BITS;
        if (j < M) {
            const float *r = xyz2 + ((size_t)b_s * M + j) * 3;
            db = __float_as_uint(sqdist_xyz(__ldg(r), __ldg(r + 1), __ldg(r + 2), x_s, y_s, z_s));
        }
        const unsigned mn = __reduce_min_sync(FULL_MASK, db);
        const unsigned hit = __ballot_sync(FULL_MASK, db == mn);
        if (lane == s) {
            want = mn;
            found = g_s * CS_GR + __ffs(hit) - 1;
        }
    }
    float s1 = 0.f;
    if (todo) {
        dist1[t] = __uint_as_float(want);
        idx1[t] = found;
        s1 = __uint_as_float(want);
        if (gw != nullptr)
            bwd_term(__fmul_rn(__ldg(gw + 0), 2.f), qx, qy, qz, xyz2 + ((size_t)b * M + found) * 3, g1 + (size_t)t * 3,
                     g2 + ((size_t)b * M + found) * 3);
    }
    if (sums != nullptr) {
        __shared__ float sh[8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(FULL_MASK, s1, o);
        if (lane == 0) sh[threadIdx.x >> 5] = s1;
        __syncthreads();
        if (threadIdx.x == 0) {
            float a = 0.f;
#pragma unroll
            for (int i = 0; i < 8; i++) a += sh[i];
            atomicAdd(sums + 0, a);
        }
    }
}

// ---- ambiguous rows and columns: walk every partner, exact-evaluate those within TAU ----------
// One CTA per entry; the approximate values are recomputed with the sweep's own operation order,
// so the partner that produced the recorded best value passes the test again.
__global__ void __launch_bounds__(256)
cs_rescan_kernel(const float *__restrict__ xyz1, const float *__restrict__ xyz2, int N, int M, int npad, int mblk,
                 const float4 *__restrict__ prepq, const float *__restrict__ prepr,
                 const unsigned *__restrict__ r2bits, const unsigned *__restrict__ ctrl,
                 const uint2 *__restrict__ rowlist, const uint2 *__restrict__ collist, float *__restrict__ dist1,
                 float *__restrict__ dist2, int *__restrict__ idx1, int *__restrict__ idx2, float *__restrict__ sums,
                 const float *__restrict__ gw, float *__restrict__ g1, float *__restrict__ g2) {
    __shared__ unsigned long long sh[8];
    pdl_wait();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned nrow = __ldcg(ctrl + 0), ncol = __ldcg(ctrl + 1);
    for (unsigned e = blockIdx.x; e < nrow + ncol; e += gridDim.x) {
        const bool is_row = e < nrow;
        const uint2 ent = is_row ? rowlist[e] : collist[e - nrow];
        const int na = is_row ? N : M;   // own cloud
        const int nb = is_row ? M : N;   // partner cloud
        const int b = (int)(ent.x / (unsigned)na), a = (int)(ent.x % (unsigned)na);
        const float tau = __uint_as_float(__ldcg(r2bits + b)) * CS_TAU_PER_R2;
        const float thr = __fadd_rn(__uint_as_float(ent.y), tau);
        const float *pa = (is_row ? xyz1 : xyz2) + (size_t)ent.x * 3;
        const float *pb = (is_row ? xyz2 : xyz1) + (size_t)b * nb * 3;
        const float ax = __ldg(pa), ay = __ldg(pa + 1), az = __ldg(pa + 2);
        unsigned long long bestkey = 0xffffffffffffffffull;
        if (is_row) {
            const float4 q = prepq[(size_t)b * npad + a];
            const float qq = __fadd_rn(q.w, tau);
            const float *rb = prepr + (size_t)b * mblk * (4 * CS_RB);
#pragma unroll 4
            for (int j = threadIdx.x; j < M; j += 256) {
                const float *blk = rb + (size_t)(j / CS_RB) * (4 * CS_RB) + (j % CS_RB);
                const float f = __fadd_rn(approx_e(blk[0], blk[CS_RB], blk[2 * CS_RB], blk[3 * CS_RB], q.x, q.y, q.z), qq);
                if (f <= thr) {
                    const float d = sqdist_xyz(__ldg(pb + (size_t)j * 3), __ldg(pb + (size_t)j * 3 + 1),
                                               __ldg(pb + (size_t)j * 3 + 2), ax, ay, az);
                    const unsigned long long k = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)j;
                    bestkey = k < bestkey ? k : bestkey;
                }
            }
        } else {
            const float *blk = prepr + ((size_t)b * mblk + a / CS_RB) * (4 * CS_RB) + (a % CS_RB);
            const float rx = blk[0], ry = blk[CS_RB], rz = blk[2 * CS_RB], rr = blk[3 * CS_RB];
            const float4 *qb = prepq + (size_t)b * npad;
#pragma unroll 4
            for (int i = threadIdx.x; i < N; i += 256) {
                const float4 q = qb[i];
                const float f = __fadd_rn(approx_e(rx, ry, rz, rr, q.x, q.y, q.z), __fadd_rn(q.w, tau));
                if (f <= thr) {
                    const float d = sqdist_xyz(__ldg(pb + (size_t)i * 3), __ldg(pb + (size_t)i * 3 + 1),
                                               __ldg(pb + (size_t)i * 3 + 2), ax, ay, az);
                    const unsigned long long k = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)i;
                    bestkey = k < bestkey ? k : bestkey;
                }
            }
        }
        // (exact value, index) ascending: lowest index among the exact minima
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(FULL_MASK, bestkey, o);
            bestkey = other < bestkey ? other : bestkey;
        }
        if (lane == 0) sh[w] = bestkey;
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 1; i < 8; i++) bestkey = sh[i] < bestkey ? sh[i] : bestkey;
            const float d = __uint_as_float((unsigned)(bestkey >> 32));
            const int found = (int)(unsigned)bestkey;
            (is_row ? dist1 : dist2)[ent.x] = d;
            (is_row ? idx1 : idx2)[ent.x] = found;
            if (found >= 0 && found < nb) {  // always, by the error bound; never index out of range
                if (sums != nullptr) atomicAdd(sums + (is_row ? 0 : 1), d);
                if (gw != nullptr) {
                    float *ga = is_row ? g1 : g2, *gb = is_row ? g2 : g1;
                    bwd_term(__fmul_rn(__ldg(gw + (is_row ? 0 : 1)), 2.f), ax, ay, az, pb + (size_t)found * 3,
                             ga + (size_t)ent.x * 3, gb + ((size_t)b * nb + found) * 3);
                }
            }
        }
        __syncthreads();  // sh[] is reused by the next entry
    }
}

}  // namespace

size_t chamfer_sweep_workspace_bytes(int B, int N, int M) { return cs_layout(B, N, M).total; }

// gw / g1 / g2 != nullptr: fused uniform backward.  `sums` has been cleared by the caller.
int chamfer_sweep_launch(const float *xyz1, const float *xyz2, int B, int N, int M, float *dist1, float *dist2,
                         int *idx1, int *idx2, float *sums, void *workspace, size_t workspace_bytes,
                         const float *gw, float *g1, float *g2, cudaStream_t st) {
    const CsLayout L = cs_layout(B, N, M);
    if (workspace_bytes < L.total) {
        set_error("chamfer_fwd: workspace %zu < %zu bytes", workspace_bytes, L.total);
        return PP_ENOSPC;
    }
    PP_REQUIRE(B <= 65535 && L.mblk <= 65535, "chamfer: grid too large (B=%d, M=%d)", B, M);
    PP_REQUIRE((long long)B * N < (1ll << 32) && (long long)B * M < (1ll << 32), "chamfer: B*N too large");
    char *ws = (char *)workspace;
    unsigned *ctrl = (unsigned *)(ws + L.ctrl);
    unsigned *r2bits = ctrl + 16;
    float4 *prepq = (float4 *)(ws + L.prepq);
    float *prepr = (float *)(ws + L.prepr);
    unsigned long long *key1 = (unsigned long long *)(ws + L.key1);
    unsigned *sec1 = (unsigned *)(ws + L.sec1);
    uint2 *rowlist = (uint2 *)(ws + L.rowlist), *collist = (uint2 *)(ws + L.collist);

    PP_CUDA(cudaMemsetAsync(ctrl, 0, 64 + 4 * (size_t)B, st));
    {
        KernelTimer timer("chamfer_prep", st);
        const int per_cloud = max(L.npad, L.mblk * CS_RB);
        int chunks = ceil_div(per_cloud, 256 * 4);
        const int want = ceil_div(2 * NUM_SMS_B200, B);  // at least two CTAs per SM over the whole grid
        if (chunks < want) chunks = min(want, ceil_div(per_cloud, 256));
        cs_prep_kernel<<<dim3(chunks, B), 256, 0, st>>>(xyz1, xyz2, N, M, L.npad, L.mblk, prepq, prepr, key1, sec1,
                                                       r2bits, gw ? g1 : nullptr, gw ? g2 : nullptr);
        PP_LAUNCH_CHECK();
    }
    CsOut out;
    out.dist2 = dist2; out.idx2 = idx2; out.sums = sums; out.gw = gw; out.g1 = g1; out.g2 = g2;
    out.ctrl = ctrl; out.collist = collist;
    {
        KernelTimer timer("chamfer_fwd", st);
        // a warp takes 256 queries: 2-warp CTAs when 4-warp CTAs would leave warp slots idle
        const int nwt = L.npad / CS_WT;
        int warps = get_option("chamfer_sweep_warps", 0);
        if (warps != 2 && warps != 4) warps = (ceil_div(nwt, 2) * 2 < ceil_div(nwt, 4) * 4) ? 2 : 4;
        static bool smem_set_dev[64] = {};  // opt in to > 48 KB of dynamic shared memory, once per device
        int dev = 0;
        PP_CUDA(cudaGetDevice(&dev));
        bool &smem_set = smem_set_dev[dev & 63];
        if (!smem_set) {
            PP_CUDA(cudaFuncSetAttribute(cs_sweep_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, CsSmem<4>::BYTES));
            PP_CUDA(cudaFuncSetAttribute(cs_sweep_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, CsSmem<2>::BYTES));
            smem_set = true;
        }
        if (warps == 4)
            PP_CUDA(launch_pdl(cs_sweep_kernel<4>, dim3(B, L.mblk), dim3(128), CsSmem<4>::BYTES, st, xyz1, xyz2, N, M, L.npad, L.mblk,
                               (const float4 *)prepq, (const float *)prepr, key1, sec1, (const unsigned *)r2bits, out));
        else
            PP_CUDA(launch_pdl(cs_sweep_kernel<2>, dim3(B, L.mblk), dim3(64), CsSmem<2>::BYTES, st, xyz1, xyz2, N, M, L.npad, L.mblk,
                               (const float4 *)prepq, (const float *)prepr, key1, sec1, (const unsigned *)r2bits, out));
        PP_LAUNCH_CHECK();
    }
    {
        KernelTimer timer("chamfer_finalize", st);
        const unsigned blocks = (unsigned)ceil_div_ll((long long)B * N, 256);
        PP_CUDA(launch_pdl(cs_finalize_rows_kernel, dim3(blocks), dim3(256), 0, st, xyz1, xyz2, B, N, M,
                           (const unsigned long long *)key1, (const unsigned *)sec1, (const unsigned *)r2bits, dist1,
                           idx1, sums, gw, g1, g2, ctrl, rowlist));
        PP_LAUNCH_CHECK();
    }
    {
        KernelTimer timer("chamfer_rescan", st);
        PP_CUDA(launch_pdl(cs_rescan_kernel, dim3(NUM_SMS_B200 * 8), dim3(256), 0, st, xyz1, xyz2, N, M, L.npad,
                           L.mblk, (const float4 *)prepq, (const float *)prepr, (const unsigned *)r2bits,
                           (const unsigned *)ctrl, (const uint2 *)rowlist, (const uint2 *)collist, dist1, dist2, idx1,
                           idx2, sums, gw, g1, g2));
        PP_LAUNCH_CHECK();
    }
    return PP_OK;
}

}  // namespace pp
