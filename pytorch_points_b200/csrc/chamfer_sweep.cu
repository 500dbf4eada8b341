// chamfer_sweep.cu -- Chamfer / nndistance forward for c == 3 clouds on the 5th-generation tensor
// cores: an approximate sweep (tcgen05.mma, accumulators in TMEM) followed by an exact resolution
// of the few surviving candidates.  Results are bit-identical to the reference's.
//
// Replaces NmDistanceKernel x2 (_ext/nmdistance_cuda.cu:8-49,127-128).  The reference's distance
//      d = fma(tz,tz, fma(ty,ty, rn(tx*tx))),  t = rn(ref - query)                (6 FP32 lane-ops)
// has to be reproduced bit for bit, and so does its lowest-index tie rule -- but only for the pair
// that WINS.  Every other pair merely has to be shown to lose.  So the sweep orders the pairs of a
// query by the expansion  e = |r|^2 - 2 q.r  on centred coordinates (|q|^2 is a per-query constant)
// -- a GEMM with K = 16 -- whose error against the reference's value is bounded by EPS = 128 u R^2
// (u = 2^-24, R = largest centred norm; derivation in DESIGN.md §3.1).  Per query it keeps the best
// value, the 32-reference granule it occurred in, the runner-up value over all OTHER granules, and a
// 64-bit mask of the reference blocks that came within TAU = 2.5 EPS of the running best:
//   runner-up > best + TAU  =>  every true minimiser (ties included) lies inside the recorded granule,
//       whose 32 pairs are re-evaluated with the exact chain;
//   otherwise the point is "ambiguous" (2 % of a uniform cloud, 9 % of a sphere surface, every point
//       of a lattice): the blocks in its mask -- a superset of wherever a minimiser can be -- are
//       re-evaluated exactly and the lowest index among the exact minima is taken.
// Results therefore equal the reference's on every input; only the time depends on the data.
//
// Both directions (dist1/idx1 and dist2/idx2) run the same one-sided pass with the roles of the
// clouds swapped, like the reference's two launches.
//
// fp32 accuracy from TF32 operands: every coordinate is split into hi + lo (each exactly representable
// in TF32) and 12 of the K = 16 slots are spent on
//      A row (query):     qh.x qh.y qh.z qh.x | qh.y qh.z ql.x ql.y | ql.z 1 1 1 | 0 0 0 0      (q = -2 * centred point)
//      B row (reference): rh.x rh.y rh.z rl.x | rl.y rl.z rh.x rh.y | rh.z n0 n1 n2 | 0 0 0 0  (|r|^2 = n0 + n1 + n2)
// -> sum = qh.rh + qh.rl + ql.rh + |r|^2; only ql.rl (<= 2^-22 |q||r|) is dropped.  Every product of two
// TF32 values is exact in fp32; what the tensor core's adder tree loses is inside EPS.
// Operands are K-major without swizzle: 16-byte chunk c of row r sits at c * 2048 + r * 16 of an 8 KB tile
// (8-row x 16-byte core matrices, SBO = 128 B, LBO = 2048 B), so ONE cp.async.bulk (UBLKCP) moves a block.
//
// Kernels (one stream, programmatic dependent launch):
//   cs_prep_kernel        centre (mean of the leading points), centred coordinates, norms, R^2; every point
//                         as an A row (query role) and a B row (reference role); resets keys, gradients.  No
//                         cleared memory is assumed anywhere: the scratch buffer carries no state between calls.
//   cs_rowpass_tc_kernel  CTA = 128 queries (TMEM lane = query) x a chunk of 128-reference blocks.  6 warps:
//                         0-3 epilogue (each reads its 32 lanes with tcgen05.ld 32x32b.x32, one query per
//                         thread: FMNMX3 trees + granule bookkeeping), 4 = copy issuer + TMEM owner (ring of
//                         8 reference blocks, one cp.async.bulk each), 5 = MMA issuer (two kind::tf32 M128 N128
//                         K8 instructions per block).  Two 128-column accumulators alternate; an accumulator
//                         goes back to the issuer as soon as its block sits in registers.  256 TMEM columns
//                         per CTA -> two CTAs per SM.  What paces it (profiles/r02_chamfer_tc_pipeline_
//                         experiments.txt): the accumulator hand-over between the issuing thread and the
//                         epilogue (1.2 ms of 2.4 at 256 x 8192^2 with all work switched off) and the ALU pipe
//                         (FMNMX3 issues every second cycle); the tcgen05.ld themselves run at a third of
//                         the rate the same loads reach alone.
//   cs_finalize_kernel    exact resolution: the recorded granule, or for ambiguous points the blocks of the mask
//                         (dist / idx, loss sums, fused backward).
#include <algorithm>
#include <atomic>

#include "pp_common.cuh"
#include "tc_common.cuh"

namespace pp {
namespace {

constexpr int CS_GR = 32;            // granule (references)
constexpr int TC_STAGES = 8;         // reference blocks in flight per CTA
constexpr unsigned CS_INF_BITS = 0x7f800000u;
constexpr unsigned long long CS_KEY_INIT = 0x7f800000ffffffffull;
// TAU = 2.5 * EPS, EPS = 128 u R^2, u = 2^-24
constexpr float CS_TAU_PER_R2 = 320.f * 5.9604644775390625e-8f;
constexpr int CS_R2_SLOTS = 64;      // preparation CTAs per cloud pair (32 per cloud): each leaves its own maximum

// One direction of the problem: `nq` query points (cloud A) against `nr` reference points (cloud B).
struct CsDir {
    const float *qxyz, *rxyz;        // original clouds (B,nq,3) / (B,nr,3)
    const float *aform;              // (B, tiles, 4 chunks, 128 rows, 4) tensor-core A operand of the queries
    const float *bform;              // (B, rblk, 4 chunks, 128 rows, 4) tensor-core B operand of the references
    const float *norm;               // (B, tiles * 128) |q|^2 of the centred queries
    unsigned long long *key;         // (B, nq) value bits << 32 | granule
    unsigned *sec;                   // (B, nq) runner-up value bits
    unsigned long long *mask;        // (B, nq) reference blocks (groups of 2^mask_shift blocks) within TAU of the best
    float *dist;                     // outputs (B, nq)
    int *idx;
    float *gq, *gr;                  // fused backward: gradient of the query / reference cloud (or null)
    int nq, nr, rblk, tiles;
    int nchunks, chunk_blocks, mask_shift;
};

struct CsArgs {
    CsDir d[2];
    const unsigned *r2part;  // (B, CS_R2_SLOTS) bits of the preparation CTAs' largest centred squared norms
    unsigned *taubits;       // (B) bits of TAU, written by the sweep for the resolving kernel
    float *sums;             // [sum(dist1), sum(dist2)] or null
    unsigned *qctr;          // work-item counter of the persistent sweep (zeroed by the preparation kernel)
    const float *gw;         // fused backward weights or null
    int B;
};

struct CsLayout {
    size_t ctrl, aform[2], bform[2], norm[2], key[2], sec[2], mask[2], total;
    int blk[2];
};

CsLayout cs_layout(int B, int N, int M) {
    CsLayout L;
    const int n[2] = {N, M};
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t o = 0;
    L.ctrl = o; o += up(4 * ((size_t)B * (CS_R2_SLOTS + 1) + 1));  // r2part[B][CS_R2_SLOTS], taubits[B], item counter
    for (int s = 0; s < 2; s++) {
        L.blk[s] = ceil_div(n[s], CS_RB);
        const size_t rows = (size_t)B * L.blk[s] * CS_RB;
        L.aform[s] = o; o += up(64 * rows);
        L.bform[s] = o; o += up(64 * rows);
        L.norm[s] = o;  o += up(4 * rows);
        L.key[s] = o;   o += up(8 * (size_t)B * n[s]);
        L.sec[s] = o;   o += up(4 * (size_t)B * n[s]);
        L.mask[s] = o;  o += up(8 * (size_t)B * n[s]);
    }
    L.total = o;
    return L;
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
cs_prep_kernel(const float *__restrict__ xyz1, const float *__restrict__ xyz2, int N, int M, float *__restrict__ af0,
               float *__restrict__ af1, float *__restrict__ bf0, float *__restrict__ bf1, float *__restrict__ nm0,
               float *__restrict__ nm1, unsigned long long *__restrict__ key0, unsigned long long *__restrict__ key1,
               unsigned *__restrict__ sec0, unsigned *__restrict__ sec1, unsigned long long *__restrict__ msk0,
               unsigned long long *__restrict__ msk1, unsigned *__restrict__ r2part, float *__restrict__ g1,
               float *__restrict__ g2, unsigned *__restrict__ qctr) {
    pdl_launch_dependents();
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) *qctr = 0u;
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
    float *af = s ? af1 : af0, *bf = s ? bf1 : bf0, *nm = s ? nm1 : nm0;
    unsigned long long *key = s ? key1 : key0, *msk = s ? msk1 : msk0;
    unsigned *sec = s ? sec1 : sec0;
    float *g = s ? g2 : g1;
    const int rblk = ceil_div(n, CS_RB);
    constexpr int CH = TC_TILE_BYTES / 16;  // floats between two 16-byte chunks of one operand row
    float r2 = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rblk * CS_RB; i += gridDim.x * blockDim.x) {
        float x = 0.f, y = 0.f, z = 0.f, nn = 0.f;
        const bool pad = i >= n;
        if (!pad) {
            const size_t t = (size_t)b * n + i;
            const float *p = xyz + t * 3;
            x = __fsub_rn(p[0], cx); y = __fsub_rn(p[1], cy); z = __fsub_rn(p[2], cz);
            nn = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
            r2 = fmaxf(r2, nn);
            key[t] = CS_KEY_INIT;
            sec[t] = CS_INF_BITS;
            msk[t] = 0ull;
            if (g != nullptr) { g[t * 3 + 0] = 0.f; g[t * 3 + 1] = 0.f; g[t * 3 + 2] = 0.f; }
        }
        nm[(size_t)b * rblk * CS_RB + i] = nn;
        // the point as row (i % 128) of an A tile (query role, q = -2 p) and of a B tile (reference role)
        const size_t tile = ((size_t)b * rblk + i / CS_RB) * (TC_TILE_BYTES / 4) + (size_t)(i % CS_RB) * 4;
        const float qx = -2.f * x, qy = -2.f * y, qz = -2.f * z;
        const float qhx = to_tf32(qx), qhy = to_tf32(qy), qhz = to_tf32(qz);
        const float qlx = to_tf32(qx - qhx), qly = to_tf32(qy - qhy), qlz = to_tf32(qz - qhz);
        const float rhx = to_tf32(x), rhy = to_tf32(y), rhz = to_tf32(z);
        const float rlx = to_tf32(x - rhx), rly = to_tf32(y - rhy), rlz = to_tf32(z - rhz);
        // padding references carry a huge (finite, TF32-exact) norm: never a minimum, no inf * 0 in the MMA
        const float nv = pad ? 1.0e30f : nn;
        const float n0 = to_tf32(nv), n1 = pad ? 0.f : to_tf32(nv - n0), n2 = pad ? 0.f : to_tf32(nv - n0 - n1);
        float *at = af + tile, *bt = bf + tile;
        *reinterpret_cast<float4 *>(at) = make_float4(qhx, qhy, qhz, qhx);
        *reinterpret_cast<float4 *>(at + CH) = make_float4(qhy, qhz, qlx, qly);
        *reinterpret_cast<float4 *>(at + 2 * CH) = make_float4(qlz, 1.f, 1.f, 1.f);
        *reinterpret_cast<float4 *>(at + 3 * CH) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4 *>(bt) = make_float4(rhx, rhy, rhz, rlx);
        *reinterpret_cast<float4 *>(bt + CH) = make_float4(rly, rlz, rhx, rhy);
        *reinterpret_cast<float4 *>(bt + 2 * CH) = make_float4(rhz, n0, n1, n2);
        *reinterpret_cast<float4 *>(bt + 3 * CH) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // this CTA's largest squared norm into its own slot (plain store: nothing has to be cleared beforehand);
    // the consumers take the maximum over the CS_R2_SLOTS slots of the cloud pair
    __shared__ unsigned s_r2[8];
    const unsigned rb = __reduce_max_sync(FULL_MASK, __float_as_uint(r2));  // r2 >= 0: bits order like values
    if (lane == 0) s_r2[threadIdx.x >> 5] = rb;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned m = 0u;
#pragma unroll
        for (int i = 0; i < 8; i++) m = max(m, s_r2[i]);
        r2part[(size_t)b * CS_R2_SLOTS + s * (CS_R2_SLOTS / 2) + blockIdx.x] = m;
    }
    // slots of CTAs that do not exist
    if (blockIdx.x == 0 && threadIdx.x >= 32 && (int)threadIdx.x - 32 + (int)gridDim.x < CS_R2_SLOTS / 2)
        r2part[(size_t)b * CS_R2_SLOTS + s * (CS_R2_SLOTS / 2) + gridDim.x + threadIdx.x - 32] = 0u;
}

// ---- the one-sided pass on the tensor cores ---------------------------------------------------------
// Warp roles: four epilogue warps (warp w reads TMEM lanes 32 w ..., the hardware's lane window of that warp;
// thread = query), one copy issuer that also owns the TMEM allocation, one MMA issuer.
constexpr int TC_EPI = 4;
constexpr int TC_THREADS = (TC_EPI + 2) * 32;
constexpr int TC_GRAN_PER_WARP = CS_RB / CS_GR;

__global__ void __launch_bounds__(TC_THREADS, 2)
cs_rowpass_tc_kernel(const CsArgs args) {
    extern __shared__ __align__(128) unsigned char cs_dyn_smem[];  // A tile | TC_STAGES B tiles
    unsigned char *sA = cs_dyn_smem;
    unsigned char (*sB)[TC_TILE_BYTES] = reinterpret_cast<unsigned char (*)[TC_TILE_BYTES]>(cs_dyn_smem + TC_TILE_BYTES);
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
        for (int i = 0; i < 2; i++) { mbar_init(bar_tf + 8 * i, 1); mbar_init(bar_te + 8 * i, TC_EPI); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (w == TC_EPI) {  // one warp owns the tensor-memory allocation: 256 columns = two 128-column accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&sTmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = sTmem;
    pdl_wait();  // the prepared operands, R^2 and the reset keys are complete and visible

    if (w == TC_EPI) {
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
    } else if (w == TC_EPI + 1) {
        if (lane == 0) {  // ---- MMA issuer: D[128 x 128] = A[128 x 16] * B[128 x 16]^T as two K = 8 instructions
            // This lone thread sets the pace of the CTA (every instruction of it is exposed latency), so its loop
            // is unrolled over the ring: stage, accumulator and both wait parities are constants of the slot.
            static_assert(TC_STAGES % 4 == 0, "the slot decides accumulator and its wait parity");
            // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128
            constexpr unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
            const unsigned long long adesc = tc_smem_desc(smem_u32(sA));
            const unsigned long long bdesc0 = tc_smem_desc(smem_u32(sB[0]));
            mbar_wait(bar_a, 0);
            unsigned ring_parity = 0u;
            for (int i0 = 0; i0 < nblk; i0 += TC_STAGES, ring_parity ^= 1u) {
#pragma unroll
                for (int u = 0; u < TC_STAGES; u++) {
                    if (i0 + u >= nblk) break;
                    const int acc = u & 1;
                    mbar_wait(bar_bf + 8 * u, ring_parity);
                    // the epilogue drained this accumulator: its use number (i0 + u) / 2 - 1 has the parity of u / 2 - 1
                    if (i0 + u >= 2) mbar_wait(bar_te + 8 * acc, (unsigned)((u >> 1) + 1) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const unsigned long long bdesc = bdesc0 + (unsigned long long)(u * (TC_TILE_BYTES >> 4));
                    const unsigned d = tmem + (unsigned)acc * 128u;
                    // descriptor start addresses count 16-byte units: K chunks 2 and 3 start 2 * 2048 B further on
                    tc_mma_tf32(d, adesc, bdesc, idesc, 0u);
                    tc_mma_tf32(d, adesc + (2 * 2048 >> 4), bdesc + (2 * 2048 >> 4), idesc, 1u);
                    tc_commit(bar_be + 8 * u);    // the block's shared-memory stage is free once both MMAs have read it
                    tc_commit(bar_tf + 8 * acc);  // ... and the accumulator is complete
                }
            }
        }
    } else {
        // ---- epilogue warps: thread = query = TMEM lane.  Per block four granules of 32 references.
        const int i_q = tile * CS_RB + (int)threadIdx.x;
        const unsigned rb = __reduce_max_sync(FULL_MASK, max(__ldcg(args.r2part + (size_t)b * CS_R2_SLOTS + lane),
                                                             __ldcg(args.r2part + (size_t)b * CS_R2_SLOTS + 32 + lane)));
        const float tau = __uint_as_float(rb) * CS_TAU_PER_R2;
        if (blockIdx.x == 0 && blockIdx.z == 0 && threadIdx.x == 0) args.taubits[b] = __float_as_uint(tau);
        float best = PP_INF, second = PP_INF;
        int gran = 0;
        unsigned long long mask = 0ull;
        for (int i = 0; i < nblk; i++) {
            const int acc = i & 1;
            mbar_wait(bar_tf + 8 * acc, (unsigned)(i / 2) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned t0 = tmem + ((unsigned)(w * 32) << 16) + (unsigned)acc * 128u;
            // the whole block into registers, then the accumulator goes back to the MMA issuer at once: the
            // next-but-one block is computed while this one is still being reduced
            unsigned raw[TC_GRAN_PER_WARP][32];
#pragma unroll
            for (int gi = 0; gi < TC_GRAN_PER_WARP; gi++) tc_ld32_issue(t0 + gi * CS_GR, raw[gi]);
#pragma unroll
            for (int gi = 0; gi < TC_GRAN_PER_WARP; gi++) tc_ld_wait(raw[gi]);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_te + 8 * acc);
            float bm = PP_INF;
#pragma unroll
            for (int gi = 0; gi < TC_GRAN_PER_WARP; gi++) {
                float v[32];
#pragma unroll
                for (int k = 0; k < 32; k++) v[k] = __uint_as_float(raw[gi][k]);
                float m0 = fmin3(v[0], v[1], v[2]), m1 = fmin3(v[3], v[4], v[5]);
                float m2 = fmin3(v[6], v[7], v[8]), m3 = fmin3(v[9], v[10], v[11]);
                m0 = fmin3(m0, v[12], v[13]); m1 = fmin3(m1, v[14], v[15]);
                m2 = fmin3(m2, v[16], v[17]); m3 = fmin3(m3, v[18], v[19]);
                m0 = fmin3(m0, v[20], v[21]); m1 = fmin3(m1, v[22], v[23]);
                m2 = fmin3(m2, v[24], v[25]); m3 = fmin3(m3, v[26], v[27]);
                m0 = fmin3(m0, v[28], v[29]); m1 = fmin3(m1, v[30], v[31]);
                const float gm = fminf(fmin3(m0, m1, m2), m3);
                // granule done: best / runner-up over granules, granule of the best (equal minima in two
                // granules leave runner-up == best: the point is ambiguous, as it must be)
                const int gid = (blk0 + i) * (CS_RB / CS_GR) + gi;
                second = fminf(second, fmaxf(best, gm));
                if (gm < best) gran = gid;
                best = fminf(best, gm);
                bm = fminf(bm, gm);
            }
            // the block is flagged when one of its granules lies within TAU of the best value seen so far
            // (>= the final one, so the flagged blocks are a superset of those that can hold a minimiser)
            if (bm <= __fadd_rn(best, tau)) mask |= 1ull << ((blk0 + i) >> D.mask_shift);
        }
        if (i_q < D.nq) {
            // published value = e + |q|^2 + TAU: positive, so its bits order like the values
            const float qq = __fadd_rn(__ldg(D.norm + (size_t)b * D.tiles * CS_RB + i_q), tau);
            const unsigned long long key = ((unsigned long long)__float_as_uint(__fadd_rn(best, qq)) << 32) | (unsigned)gran;
            const unsigned sb = __float_as_uint(__fadd_rn(second, qq));  // +inf stays +inf
            const size_t t = (size_t)b * D.nq + i_q;
            if (D.nchunks == 1) {
                D.key[t] = key; D.sec[t] = sb; D.mask[t] = mask;
            } else {  // the loser of every comparison at the key is a runner-up candidate
                const unsigned long long old = atomicMin(D.key + t, key);
                atomicMin(D.sec + t, min((unsigned)((old > key ? old : key) >> 32), sb));
                atomicOr(D.mask + t, mask);
            }
        }
    }
    // every tcgen05 operation of this CTA has completed (the epilogue waited for the last accumulator)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == TC_EPI) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

// ---- the same pass as a persistent kernel with a work queue --------------------------------------------------
// One CTA per accumulator pair (two per SM) takes work items (direction, cloud, query tile, chunk) off a global
// counter.  What a fresh CTA pays per tile -- launch, tensor-memory allocation behind the CTA that just left,
// barrier set-up, the first operand copies with nothing to overlap them, the last accumulator drained with the
// tensor pipe idle -- is paid once per CTA: the copy issuer fetches the next item and runs ahead into it (the
// query tile is double buffered, the reference ring never drains), the MMA issuer follows, and the epilogue warps
// only store one item's results and clear their registers between two accumulators.  The item number travels
// with the query tile's slot (written before the slot's full barrier is armed); -1 ends the kernel.  Barrier
// parities follow from use counters that all roles advance identically.
constexpr int TCP_SMEM = (2 + TC_STAGES) * TC_TILE_BYTES;

struct CsItem { int dir, b, tile, chunk; };
__device__ __forceinline__ CsItem cs_item(const CsArgs &args, int wi, int items0) {
    CsItem it;
    it.dir = wi >= items0 ? 1 : 0;
    const CsDir &D = args.d[it.dir];
    const int local = wi - (it.dir ? items0 : 0);
    const int per_cloud = D.tiles * D.nchunks;
    it.b = local / per_cloud;
    const int rem = local - it.b * per_cloud;
    it.tile = rem / D.nchunks;
    it.chunk = rem - it.tile * D.nchunks;
    return it;
}

__global__ void __launch_bounds__(TC_THREADS, 2)
cs_rowpass_tc_persistent_kernel(const CsArgs args, int items0, int items_total) {
    extern __shared__ __align__(128) unsigned char cs_dyn_smem[];  // 2 A tiles | TC_STAGES B tiles
    unsigned char (*sA)[TC_TILE_BYTES] = reinterpret_cast<unsigned char (*)[TC_TILE_BYTES]>(cs_dyn_smem);
    unsigned char (*sB)[TC_TILE_BYTES] = reinterpret_cast<unsigned char (*)[TC_TILE_BYTES]>(cs_dyn_smem + 2 * TC_TILE_BYTES);
    // a_full[2] | a_empty[2] | b_full[S] | b_empty[S] | t_full[2] | t_empty[2]
    __shared__ __align__(8) unsigned long long sBar[4 + 2 * TC_STAGES + 4];
    __shared__ unsigned sTmem;
    __shared__ int sItem[2];

    pdl_launch_dependents();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned bar_af = smem_u32(sBar), bar_ae = smem_u32(sBar + 2);
    const unsigned bar_bf = smem_u32(sBar + 4), bar_be = smem_u32(sBar + 4 + TC_STAGES);
    const unsigned bar_tf = smem_u32(sBar + 4 + 2 * TC_STAGES), bar_te = smem_u32(sBar + 6 + 2 * TC_STAGES);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
            mbar_init(bar_af + 8 * i, 1);
            mbar_init(bar_ae + 8 * i, 1 + TC_EPI);  // the item's MMAs are through AND every epilogue warp has read the item number
        }
#pragma unroll
        for (int i = 0; i < TC_STAGES; i++) { mbar_init(bar_bf + 8 * i, 1); mbar_init(bar_be + 8 * i, 1); }
#pragma unroll
        for (int i = 0; i < 2; i++) { mbar_init(bar_tf + 8 * i, 1); mbar_init(bar_te + 8 * i, TC_EPI); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (w == TC_EPI) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&sTmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = sTmem;
    pdl_wait();  // the prepared operands, R^2, the reset keys and the cleared item counter are complete and visible

    if (w == TC_EPI) {
        if (lane == 0) {  // ---- copy issuer: takes the items off the queue
            unsigned used = 0u, uses = 0u;  // per stage: used before / parity of its use count
            for (int ai = 0;; ai++) {
                const int slot = ai & 1;
                // the slot's previous item (ai - 2) is through: completion number (ai >> 1) - 1 of its empty barrier
                if (ai >= 2) mbar_wait(bar_ae + 8 * slot, (unsigned)((ai >> 1) - 1) & 1u);
                const int wi = (int)atomicAdd(args.qctr, 1u);
                if (wi >= items_total) {
                    sItem[slot] = -1;
                    mbar_arrive(bar_af + 8 * slot);
                    break;
                }
                sItem[slot] = wi;
                const CsItem it = cs_item(args, wi, items0);
                const CsDir &D = args.d[it.dir];
                const int blk0 = it.chunk * D.chunk_blocks;
                const int nblk = min(D.rblk, blk0 + D.chunk_blocks) - blk0;
                mbar_expect_tx(bar_af + 8 * slot, TC_TILE_BYTES);
                bulk_g2s(smem_u32(sA[slot]), D.aform + ((size_t)it.b * D.tiles + it.tile) * (TC_TILE_BYTES / 4), TC_TILE_BYTES,
                         bar_af + 8 * slot);
                const float *src = D.bform + ((size_t)it.b * D.rblk + blk0) * (TC_TILE_BYTES / 4);
                for (int i = 0; i < nblk; i++) {
                    const int st = i % TC_STAGES;
                    if ((used >> st) & 1u) mbar_wait(bar_be + 8 * st, ((uses >> st) & 1u) ^ 1u);
                    used |= 1u << st;
                    uses ^= 1u << st;
                    mbar_expect_tx(bar_bf + 8 * st, TC_TILE_BYTES);
                    bulk_g2s(smem_u32(sB[st]), src + (size_t)i * (TC_TILE_BYTES / 4), TC_TILE_BYTES, bar_bf + 8 * st);
                }
            }
        }
    } else if (w == TC_EPI + 1) {
        if (lane == 0) {  // ---- MMA issuer
            constexpr unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
            const unsigned long long bdesc0 = tc_smem_desc(smem_u32(sB[0]));
            // parity of the next completion to wait for, per stage and per accumulator: registers with constant
            // indices in the unrolled loop.  The accumulators start out "drained" (pre-arrival of the epilogue warps).
            unsigned pf[TC_STAGES], pt[2] = {0u, 0u};
#pragma unroll
            for (int u = 0; u < TC_STAGES; u++) pf[u] = 0u;
            int ai = 0;
            for (;; ai++) {
                const int slot = ai & 1;
                mbar_wait(bar_af + 8 * slot, (unsigned)(ai >> 1) & 1u);
                const int wi = sItem[slot];
                if (wi < 0) break;
                const CsItem it = cs_item(args, wi, items0);
                const CsDir &D = args.d[it.dir];
                const int blk0 = it.chunk * D.chunk_blocks;
                const int nblk = min(D.rblk, blk0 + D.chunk_blocks) - blk0;
                const unsigned long long adesc = tc_smem_desc(smem_u32(sA[slot]));
                for (int i0 = 0; i0 < nblk; i0 += TC_STAGES) {
#pragma unroll
                    for (int u = 0; u < TC_STAGES; u++) {
                        if (i0 + u >= nblk) break;
                        const int acc = u & 1;
                        mbar_wait(bar_bf + 8 * u, pf[u]);
                        pf[u] ^= 1u;
                        mbar_wait(bar_te + 8 * acc, pt[acc]);
                        pt[acc] ^= 1u;
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const unsigned long long bdesc = bdesc0 + (unsigned long long)(u * (TC_TILE_BYTES >> 4));
                        const unsigned d = tmem + (unsigned)acc * 128u;
                        tc_mma_tf32(d, adesc, bdesc, idesc, 0u);
                        tc_mma_tf32(d, adesc + (2 * 2048 >> 4), bdesc + (2 * 2048 >> 4), idesc, 1u);
                        tc_commit(bar_be + 8 * u);
                        tc_commit(bar_tf + 8 * acc);
                    }
                }
                tc_commit(bar_ae + 8 * slot);
            }
            // nothing of this CTA may still be in flight towards its shared memory when it exits
            if (ai > 0) mbar_wait(bar_ae + 8 * ((ai - 1) & 1), (unsigned)((ai - 1) >> 1) & 1u);
        }
    } else {
        // ---- epilogue warps: thread = query = TMEM lane
        if (lane == 0) { mbar_arrive(bar_te); mbar_arrive(bar_te + 8); }  // both accumulators start out drained
        unsigned tfp[2] = {0u, 0u};  // parity of the next full-barrier completion to wait for, per accumulator
        for (int ai = 0;; ai++) {
            const int slot = ai & 1;
            mbar_wait(bar_af + 8 * slot, (unsigned)(ai >> 1) & 1u);
            const int wi = sItem[slot];
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ae + 8 * slot);  // this warp has the item number
            if (wi < 0) break;
            const CsItem it = cs_item(args, wi, items0);
            const CsDir &D = args.d[it.dir];
            const int b = it.b;
            const int blk0 = it.chunk * D.chunk_blocks;
            const int nblk = min(D.rblk, blk0 + D.chunk_blocks) - blk0;
            const int i_q = it.tile * CS_RB + (int)threadIdx.x;
            const unsigned rb = __reduce_max_sync(FULL_MASK, max(__ldcg(args.r2part + (size_t)b * CS_R2_SLOTS + lane),
                                                                 __ldcg(args.r2part + (size_t)b * CS_R2_SLOTS + 32 + lane)));
            const float tau = __uint_as_float(rb) * CS_TAU_PER_R2;
            if (it.dir == 0 && it.tile == 0 && it.chunk == 0 && threadIdx.x == 0) args.taubits[b] = __float_as_uint(tau);
            float best = PP_INF, second = PP_INF;
            int gran = 0;
            unsigned long long mask = 0ull;
#pragma unroll 2
            for (int i = 0; i < nblk; i++) {  // (unrolled by two: the accumulator index is a constant of each copy)
                const int acc = i & 1;
                mbar_wait(bar_tf + 8 * acc, tfp[acc]);
                tfp[acc] ^= 1u;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned t0 = tmem + ((unsigned)(w * 32) << 16) + (unsigned)acc * 128u;
                unsigned raw[TC_GRAN_PER_WARP][32];
#pragma unroll
                for (int gi = 0; gi < TC_GRAN_PER_WARP; gi++) tc_ld32_issue(t0 + gi * CS_GR, raw[gi]);
#pragma unroll
                for (int gi = 0; gi < TC_GRAN_PER_WARP; gi++) tc_ld_wait(raw[gi]);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_te + 8 * acc);
                float bm = PP_INF;
#pragma unroll
                for (int gi = 0; gi < TC_GRAN_PER_WARP; gi++) {
                    float v[32];
#pragma unroll
                    for (int k = 0; k < 32; k++) v[k] = __uint_as_float(raw[gi][k]);
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
                    bm = fminf(bm, gm);
                }
                if (bm <= __fadd_rn(best, tau)) mask |= 1ull << ((blk0 + i) >> D.mask_shift);
            }
            if (i_q < D.nq) {
                const float qq = __fadd_rn(__ldg(D.norm + (size_t)b * D.tiles * CS_RB + i_q), tau);
                const unsigned long long key = ((unsigned long long)__float_as_uint(__fadd_rn(best, qq)) << 32) | (unsigned)gran;
                const unsigned sb = __float_as_uint(__fadd_rn(second, qq));
                const size_t t = (size_t)b * D.nq + i_q;
                if (D.nchunks == 1) {
                    D.key[t] = key; D.sec[t] = sb; D.mask[t] = mask;
                } else {
                    const unsigned long long old = atomicMin(D.key + t, key);
                    atomicMin(D.sec + t, min((unsigned)((old > key ? old : key) >> 32), sb));
                    atomicOr(D.mask + t, mask);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == TC_EPI) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

// ---- exact resolution ------------------------------------------------------------------------------
// One launch for both directions: blocks [0, blocks0) take the points of cloud 1 (dist1 / idx1), the
// rest those of cloud 2.  A warp takes 32 points, one after the other with all 32 lanes:
//   decided point   -- the 32 references of its recorded granule at once (coalesced 384-byte read,
//                      REDUX.MIN, ballot, find-first-set);
//   ambiguous point -- every reference block of its mask (a superset of the blocks that can hold a
//                      minimiser), four blocks per trip; the smallest (exact value, index) pair wins.
// Either way the result is the lowest index among the exact minima, like the reference's strict '<' scan.
__global__ void __launch_bounds__(256)
cs_finalize_kernel(const CsArgs args, int blocks0) {
    const int lane = threadIdx.x & 31;
    pdl_wait();
    pdl_launch_dependents();
    const int dir = (int)blockIdx.x >= blocks0 ? 1 : 0;
    const CsDir &D = args.d[dir];
    const long long total = (long long)args.B * D.nq;
    const long long t = (long long)((int)blockIdx.x - (dir ? blocks0 : 0)) * 256 + threadIdx.x;
    int state = 0;  // 0 = no point, 1 = decided, 2 = ambiguous
    int gran = 0, b = 0;
    unsigned long long mask = 0ull;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (t < total) {
        const unsigned long long key = __ldcg(D.key + t);
        const unsigned sec = __ldcg(D.sec + t);
        const unsigned vb = (unsigned)(key >> 32);
        gran = (int)(unsigned)key;
        b = (int)(t / D.nq);
        const float tau = __uint_as_float(__ldcg(args.taubits + b));
        state = 1;
        if (sec <= __float_as_uint(__fadd_rn(__uint_as_float(vb), tau))) {
            state = 2;
            mask = __ldcg(D.mask + t);
        }
        const float *q = D.qxyz + (size_t)t * 3;
        qx = q[0]; qy = q[1]; qz = q[2];
    }
    unsigned want = 0u, hit_mine = 1u;  // decided points: the exact minimum and the lanes of the granule that reach it
    int found = 0;
    const unsigned decided = __ballot_sync(FULL_MASK, state == 1), ambiguous = __ballot_sync(FULL_MASK, state == 2);
    const int span = CS_RB << D.mask_shift;  // references per mask bit
    // the 32 points of a warp almost always belong to one cloud: its reference base is kept across points
    int b_cur = __shfl_sync(FULL_MASK, b, 0);
    const float *rx = D.rxyz + (size_t)b_cur * D.nr * 3;
    const int last = D.nr - 1;
#pragma unroll 4
    for (int s = 0; s < 32; s++) {
        if (!(((decided | ambiguous) >> s) & 1u)) continue;  // warp-uniform
        const int b_s = __shfl_sync(FULL_MASK, b, s);
        const float x_s = __shfl_sync(FULL_MASK, qx, s);
        const float y_s = __shfl_sync(FULL_MASK, qy, s);
        const float z_s = __shfl_sync(FULL_MASK, qz, s);
        if (b_s != b_cur) {  // warp-uniform
            b_cur = b_s;
            rx = D.rxyz + (size_t)b_s * D.nr * 3;
        }
        if ((decided >> s) & 1u) {
            const int g_s = __shfl_sync(FULL_MASK, gran, s);
            // lanes past the end of the cloud repeat its last point: an equal value in a HIGHER lane than the
            // real one, which the find-first-set below never picks
            const float *r = rx + (size_t)min(g_s * CS_GR + lane, last) * 3;
            // (the squares make the operand order irrelevant: both of the reference's launches give these bits)
            const unsigned db = __float_as_uint(sqdist_xyz(__ldg(r), __ldg(r + 1), __ldg(r + 2), x_s, y_s, z_s));
            const unsigned mn = __reduce_min_sync(FULL_MASK, db);
            const unsigned hit = __ballot_sync(FULL_MASK, db == mn);
            if (lane == s) {
                want = mn;
                hit_mine = hit;
            }
        } else {
            unsigned long long m = ((unsigned long long)__shfl_sync(FULL_MASK, (unsigned)(mask >> 32), s) << 32) |
                                   __shfl_sync(FULL_MASK, (unsigned)mask, s);
            unsigned long long bestkey = 0xffffffffffffffffull;
            while (m) {
                // four flagged blocks per trip: their loads are independent, so they travel together
                int start[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    start[k] = -1;
                    if (m) {
                        start[k] = (__ffsll((long long)m) - 1) * span;
                        m &= m - 1;
                    }
                }
#pragma unroll 2
                for (int off = lane; off < span; off += 32) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int j = start[k] + off;
                        if (start[k] >= 0 && j < D.nr) {
                            const float d = sqdist_xyz(__ldg(rx + (size_t)j * 3), __ldg(rx + (size_t)j * 3 + 1),
                                                       __ldg(rx + (size_t)j * 3 + 2), x_s, y_s, z_s);
                            const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)j;
                            bestkey = key < bestkey ? key : bestkey;
                        }
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(FULL_MASK, bestkey, o);
                bestkey = other < bestkey ? other : bestkey;
            }
            if (lane == s) {
                want = (unsigned)(bestkey >> 32);
                found = (int)(unsigned)bestkey;
            }
        }
    }
    if (state == 1) found = gran * CS_GR + __ffs(hit_mine) - 1;
    float s1 = 0.f;
    if (state != 0) {
        D.dist[t] = __uint_as_float(want);
        D.idx[t] = found;
        s1 = __uint_as_float(want);
        // (found is always in range: the block of the recorded best value is in the mask)
        if (args.gw != nullptr && found >= 0 && found < D.nr)
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

// ---- tensor-memory read probe (pp_microbench 7 / 8) ---------------------------------------------------
// What the sweep's epilogue can reach at best: the same 32x32b.x32 loads from the same CTA shape (four
// warps per CTA, two CTAs of 256 columns per SM), mode 0 with nothing behind them, mode 1 with the granule
// minimum tree (16 three-input minima per 32 values) and nothing else.
__global__ void __launch_bounds__(128, 2) tmem_probe_kernel(float *out, int iters, int mode) {
    __shared__ unsigned sTmem;
    const int w = threadIdx.x >> 5;
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&sTmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = sTmem, t0 = tmem + ((unsigned)(w * 32) << 16);
    float acc = PP_INF;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int gi = 0; gi < 8; gi++) {
            float v[32];
            tc_ld32(t0 + gi * 32, v);
            if (mode == 0) {
                acc = fminf(acc, v[gi]);
            } else {
                float m0 = fmin3(v[0], v[1], v[2]), m1 = fmin3(v[3], v[4], v[5]);
                float m2 = fmin3(v[6], v[7], v[8]), m3 = fmin3(v[9], v[10], v[11]);
                m0 = fmin3(m0, v[12], v[13]); m1 = fmin3(m1, v[14], v[15]);
                m2 = fmin3(m2, v[16], v[17]); m3 = fmin3(m3, v[18], v[19]);
                m0 = fmin3(m0, v[20], v[21]); m1 = fmin3(m1, v[22], v[23]);
                m2 = fmin3(m2, v[24], v[25]); m3 = fmin3(m3, v[26], v[27]);
                m0 = fmin3(m0, v[28], v[29]); m1 = fmin3(m1, v[30], v[31]);
                acc = fminf(acc, fminf(fmin3(m0, m1, m2), m3));
            }
        }
    }
    if (acc == 12345.678f) out[0] = acc;  // (never: keeps the values alive)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

}  // namespace

cudaError_t chamfer_sweep_tmem_probe(int mode, int iters, float *out, double *bytes) {
    const int ctas = NUM_SMS_B200 * 2;
    tmem_probe_kernel<<<ctas, 128>>>(out, iters, mode);
    *bytes = (double)ctas * 128 * iters * 8 * 32 * 4.0;
    return cudaGetLastError();
}

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
    PP_REQUIRE(B <= 65535, "chamfer: B=%d too large", B);
    PP_REQUIRE((long long)B * N < (1ll << 31) && (long long)B * M < (1ll << 31), "chamfer: B*N too large");
    char *ws = (char *)workspace;
    unsigned *r2part = (unsigned *)(ws + L.ctrl);
    CsArgs A;
    A.r2part = r2part; A.taubits = r2part + (size_t)B * CS_R2_SLOTS; A.sums = sums; A.gw = gw; A.B = B;
    A.qctr = A.taubits + B;
    const int n[2] = {N, M};
    const float *xyz[2] = {xyz1, xyz2};
    float *dist[2] = {dist1, dist2};
    int *idx[2] = {idx1, idx2};
    float *g[2] = {gw ? g1 : nullptr, gw ? g2 : nullptr};
    // enough CTAs for several waves (two CTAs per SM): split the reference cloud into chunks when the query
    // tiles alone are too few
    const long long want_ctas = (long long)NUM_SMS_B200 * get_option("chamfer_sweep_ctas_per_sm", 8);
    int grid_x = 1;
    for (int s = 0; s < 2; s++) {
        CsDir &D = A.d[s];
        D.qxyz = xyz[s]; D.rxyz = xyz[1 - s];
        D.aform = (const float *)(ws + L.aform[s]);
        D.bform = (const float *)(ws + L.bform[1 - s]);
        D.norm = (const float *)(ws + L.norm[s]);
        D.key = (unsigned long long *)(ws + L.key[s]);
        D.sec = (unsigned *)(ws + L.sec[s]);
        D.mask = (unsigned long long *)(ws + L.mask[s]);
        D.dist = dist[s]; D.idx = idx[s];
        D.gq = g[s]; D.gr = g[1 - s];
        D.nq = n[s]; D.nr = n[1 - s];
        D.tiles = L.blk[s]; D.rblk = L.blk[1 - s];
        D.mask_shift = 0;
        while ((64 << D.mask_shift) < D.rblk) D.mask_shift++;  // 64 mask bits cover all blocks
        const long long base = 2ll * B * D.tiles;
        int chunks = (int)ceil_div_ll(want_ctas, base);
        const int max_chunks = max(1, D.rblk / 4);  // at least four blocks (512 references) per chunk
        if (chunks > max_chunks) chunks = max_chunks;
        if (chunks < 1) chunks = 1;
        D.chunk_blocks = ceil_div(D.rblk, chunks);
        D.nchunks = ceil_div(D.rblk, D.chunk_blocks);
        grid_x = max(grid_x, D.tiles * D.nchunks);
    }

    {
        KernelTimer timer("chamfer_prep", st);
        const int rows = max(L.blk[0], L.blk[1]) * CS_RB;
        int chunks = ceil_div(rows, 256 * 2);
        const int want = ceil_div(NUM_SMS_B200, B);  // at least two CTAs per SM over the whole grid
        if (chunks < want) chunks = min(want, ceil_div(rows, 256));
        chunks = min(chunks, CS_R2_SLOTS / 2);      // every CTA owns one R^2 slot; the kernel strides over the rest
        cs_prep_kernel<<<dim3(chunks, B, 2), 256, 0, st>>>(
            xyz1, xyz2, N, M, (float *)(ws + L.aform[0]), (float *)(ws + L.aform[1]), (float *)(ws + L.bform[0]),
            (float *)(ws + L.bform[1]), (float *)(ws + L.norm[0]), (float *)(ws + L.norm[1]),
            (unsigned long long *)(ws + L.key[0]), (unsigned long long *)(ws + L.key[1]), (unsigned *)(ws + L.sec[0]),
            (unsigned *)(ws + L.sec[1]), (unsigned long long *)(ws + L.mask[0]), (unsigned long long *)(ws + L.mask[1]),
            r2part, g[0], g[1], A.qctr);
        PP_LAUNCH_CHECK();
    }
    {
        KernelTimer timer("chamfer_fwd", st);
        constexpr size_t smem = (size_t)(1 + TC_STAGES) * TC_TILE_BYTES;
        static std::atomic<bool> opted_in[64];  // the opt-in to more than 48 KB is per device
        int dev = 0;
        PP_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !opted_in[dev].load(std::memory_order_acquire)) {
            PP_CUDA(cudaFuncSetAttribute(cs_rowpass_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (dev >= 0 && dev < 64) opted_in[dev].store(true, std::memory_order_release);
        }
        // Persistent form with a work queue: measured 4-5 % faster on clouds of up to ~4096 points (20-32 blocks per
        // tile: the per-CTA set-up it removes matters there), 1 % slower at 8192 (64 blocks per tile: the hardware's
        // CTA turnover is already hidden behind the other CTA of the SM) -- profiles/r02_chamfer_tc_pipeline_experiments.txt
        const int persistent_opt = get_option("chamfer_persistent", -1);
        if (persistent_opt >= 0 ? persistent_opt != 0 : std::max(L.blk[0], L.blk[1]) <= 32) {
            static std::atomic<bool> opted_in_p[64];
            if (dev < 0 || dev >= 64 || !opted_in_p[dev].load(std::memory_order_acquire)) {
                PP_CUDA(cudaFuncSetAttribute(cs_rowpass_tc_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM));
                if (dev >= 0 && dev < 64) opted_in_p[dev].store(true, std::memory_order_release);
            }
            const long long items0 = (long long)B * A.d[0].tiles * A.d[0].nchunks;
            const long long items = items0 + (long long)B * A.d[1].tiles * A.d[1].nchunks;
            PP_REQUIRE(items < (1ll << 30), "chamfer: too many tiles");
            const int grid = (int)std::min<long long>(items, 2ll * NUM_SMS_B200);
            PP_CUDA(launch_pdl(cs_rowpass_tc_persistent_kernel, dim3(grid), dim3(TC_THREADS), (size_t)TCP_SMEM, st, A, (int)items0,
                               (int)items));
        } else {
            PP_CUDA(launch_pdl(cs_rowpass_tc_kernel, dim3(grid_x, B, 2), dim3(TC_THREADS), smem, st, A));
        }
        PP_LAUNCH_CHECK();
    }
    {
        KernelTimer timer("chamfer_finalize", st);
        const int blocks0 = (int)ceil_div_ll((long long)B * N, 256), blocks1 = (int)ceil_div_ll((long long)B * M, 256);
        PP_CUDA(launch_pdl(cs_finalize_kernel, dim3(blocks0 + blocks1), dim3(256), 0, st, A, blocks0));
        PP_LAUNCH_CHECK();
    }
    return PP_OK;
}

}  // namespace pp
