// sampling.cu -- farthest point sampling, gather, ball_query, group, three_nn for sm_100a.
//
// Replaces _ext/sampling_cuda.cu + _ext/interpolate_gpu.cu (three_nn) of the reference.
#include <cooperative_groups.h>

#include "ball_scan.cuh"
#include "pp_common.cuh"

namespace cg = cooperative_groups;

namespace pp {
namespace {

// ===========================================================================
// Farthest point sampling (_ext/sampling_cuda.cu:162-233).
//
// The reference runs ONE 512-thread block per cloud, re-reads 97% of the points
// and all running minima from global memory every round, and reduces with a
// 9-level shared-memory tree.  FPS is a latency chain (m-1 dependent rounds), so
// the B200 design minimises the per-round critical path instead:
//   * a cloud is spread over a thread-block CLUSTER of C CTAs (C <= 8) x 512
//     threads; every thread keeps P points and their running minima in REGISTERS
//     for the whole kernel (no per-round memory traffic at all);
//   * per round: register update -> 2 REDUX per warp (max value, then min
//     tie-key among the maxima) -> one bar.sync -> 2 REDUX over the warp records
//     -> the CTA winner (value, tie-key, xyz) is pushed into every CTA of the
//     cluster through distributed shared memory -> one cluster barrier -> every
//     thread picks the cluster winner from C records.
//   * tie-break is the reference's: among equal maxima the smallest
//     (k mod bs, k), bs = min(2^floor(log2 N), 512)  (SURVEY.md D3, A.2).
// Thread t of the cluster owns points k = t + p*T_total (p = 0..P-1): T_total is a
// multiple of bs, so (k mod bs) is constant per thread and the tie-key grows with
// p; a strict > scan in p order therefore keeps the lowest tie-key per thread.
// ===========================================================================
constexpr int FPS_T = 512;
constexpr int FPS_MAXC = 8;
constexpr int FPS_W = FPS_T / 32;

struct FpsRec {  // 32 bytes: (value, tie-key) in the first half, the point itself in the second
    int v;       // running-min value bits (>= 0 for real points, -1.0f for padding)
    unsigned t;  // tie-key
    int pad0, pad1;
    float x, y, z;
    int k;  // point index
};

// ---- cluster point-to-point signalling (mbarrier in the destination CTA's shared memory) ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned map_to_cta(unsigned local_addr, int cta) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// Asynchronous remote store: the 16 bytes land in the destination CTA's shared memory and, as
// part of the same operation, complete 16 transaction bytes on that CTA's mbarrier.  Fire and
// forget for the sender -- no release fence (a .release.cluster arrive costs a full
// MEMBAR.GPU per round, measured ~600 cycles).
__device__ __forceinline__ void st_async_v4(unsigned dst, float4 v, unsigned mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst),
                 "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)),
                 "r"(__float_as_uint(v.w)), "r"(mbar)
                 : "memory");
}
// One local thread arms the next phase: one arrival (this one) + `bytes` of remote stores.
__device__ __forceinline__ void mbar_arm(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    const unsigned a = smem_u32(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t"
        "}" ::"r"(a), "r"(parity)
        : "memory");
}

__device__ __forceinline__ unsigned fps_tiekey(int k, int bs_log2) {
    return ((unsigned)(k & ((1 << bs_log2) - 1)) << 22) | (unsigned)(k >> bs_log2);
}

template <int P, int C>
__global__ void __launch_bounds__(FPS_T, 1)
fps_cluster_kernel(const float *__restrict__ xyz, int N, int m, int seed, float *__restrict__ temp,
                   int *__restrict__ idx, float *__restrict__ new_xyz, int bs_log2,
                   long long *__restrict__ prof) {
    // -DPP_FPS_PROFILE: thread 0 of cluster 0 accumulates clock64() deltas per phase into `prof`
    // (how the per-round breakdown in DESIGN.md was measured); compiled out otherwise.
#ifdef PP_FPS_PROFILE
    long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long pc = 0;
#define FPS_STAMP(i)                                   \
    if (prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0) { \
        const long long now = clock64();               \
        pt[i] += now - pc;                             \
        pc = now;                                      \
    }
#else
#define FPS_STAMP(i)
#endif
    extern __shared__ __align__(16) unsigned char fps_smem[];
    float4 *sPts = reinterpret_cast<float4 *>(fps_smem);  // [P][FPS_T] this CTA's points
    __shared__ __align__(8) int2 wrec[2][FPS_W];           // per-warp (value, tie-key)
    __shared__ __align__(16) FpsRec crec[2][FPS_MAXC];     // per-CTA winners, written by peers
    __shared__ __align__(8) unsigned long long cbar[2];     // "all C records of this parity arrived"

    const int rank = (C > 1) ? (int)cg::this_cluster().block_rank() : 0;
    const int b = blockIdx.x / C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tg = rank * FPS_T + tid;  // thread index within the cluster
    constexpr int TT = FPS_T * C;

    const float *pts = xyz + (size_t)b * N * 3;
    float *tp = temp + (size_t)b * N;
    int *out = idx + (size_t)b * m;
    // optional fused gather: the winner's coordinates are already in hand every round
    float *oxyz = new_xyz != nullptr ? new_xyz + (size_t)b * m * 3 : nullptr;

    float px[P], py[P], pz[P], td[P];
#pragma unroll
    for (int p = 0; p < P; p++) {
        const int k = tg + p * TT;
        if (k < N) {
            px[p] = __ldg(pts + (size_t)k * 3 + 0);
            py[p] = __ldg(pts + (size_t)k * 3 + 1);
            pz[p] = __ldg(pts + (size_t)k * 3 + 2);
            td[p] = tp[k];
        } else {
            px[p] = py[p] = pz[p] = 0.f;
            td[p] = -1.f;  // padding can never be the maximum (real minima are >= 0)
        }
        sPts[p * FPS_T + tid] = make_float4(px[p], py[p], pz[p], 0.f);
    }
    const unsigned tk0 = fps_tiekey(tg, bs_log2);  // tie-key of point p: tk0 + p * (TT >> bs_log2)
    const unsigned tkstep = (unsigned)(TT >> bs_log2);

    // coordinates of the seed point, first output
    float ox = __ldg(pts + (size_t)seed * 3 + 0), oy = __ldg(pts + (size_t)seed * 3 + 1),
          oz = __ldg(pts + (size_t)seed * 3 + 2);
    if (tg == 0) {
        out[0] = seed;
        if (oxyz != nullptr) { oxyz[0] = ox; oxyz[1] = oy; oxyz[2] = oz; }
    }
    if (C > 1 && tid == 0) {
        mbar_init(&cbar[0], 1);
        mbar_init(&cbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_arm(&cbar[0], C * (unsigned)sizeof(FpsRec));  // first use of each barrier
        mbar_arm(&cbar[1], C * (unsigned)sizeof(FpsRec));
    }
    __syncthreads();
    if (C > 1) cg::this_cluster().sync();  // peers' shared memory and barriers are live before any remote store

#ifdef PP_FPS_PROFILE
    if (prof != nullptr) pc = clock64();
#endif
    for (int j = 1; j < m; j++) {
        const int par = j & 1;
        float best = -1.f;
        int bp = 0;
        FPS_STAMP(7)
        if (P >= 2) {
            // packed pipe: two points per FADD2/FMUL2/FFMA2 (same y-first rounding order), the
            // running maximum with FMNMX3, then the first p holding it (== strict '>' scan)
            const float nox = -ox, noy = -oy, noz = -oz;
#pragma unroll
            for (int p = 0; p < P; p += 2) {
                const float2 d = sqdist2_yxz(make_float2(px[p], px[p + 1]), make_float2(py[p], py[p + 1]),
                                             make_float2(pz[p], pz[p + 1]), nox, noy, noz);
                td[p] = fminf(d.x, td[p]);
                td[p + 1] = fminf(d.y, td[p + 1]);
                best = fmax3(best, td[p], td[p + 1]);
            }
#pragma unroll
            for (int p = P - 1; p >= 0; p--) bp = (td[p] == best) ? p : bp;
        } else {
#pragma unroll
            for (int p = 0; p < P; p++) {
                const float d = sqdist_yxz(__fsub_rn(px[p], ox), __fsub_rn(py[p], oy), __fsub_rn(pz[p], oz));
                const float d2 = fminf(d, td[p]);
                td[p] = d2;
                if (d2 > best) {
                    best = d2;
                    bp = p;
                }
            }
        }
        FPS_STAMP(0)
        // warp: max value, then min tie-key among the lanes holding it (two REDUX; measured
        // faster than one REDUX + vote + rare tie path)
        const int myv = __float_as_int(best);
        const unsigned mytk = tk0 + (unsigned)bp * tkstep;
        const int wv = __reduce_max_sync(FULL_MASK, myv);
        const unsigned wt = __reduce_min_sync(FULL_MASK, myv == wv ? mytk : 0xffffffffu);
        if (lane == 0) wrec[par][warp] = make_int2(wv, (int)wt);
        FPS_STAMP(1)
        __syncthreads();
        FPS_STAMP(2)
        if (C == 1) {
            // every warp reduces the FPS_W warp records redundantly (no second barrier)
            int2 r = lane < FPS_W ? wrec[par][lane] : make_int2(INT_MIN, -1);
            const int cv = __reduce_max_sync(FULL_MASK, r.x);
            const unsigned ct = __reduce_min_sync(FULL_MASK, r.x == cv ? (unsigned)r.y : 0xffffffffu);
            // decode the winner: k = (k div bs) * bs + (k mod bs)
            const int ck = (int)((ct & 0x3fffffu) << bs_log2) | (int)(ct >> 22);
            FPS_STAMP(3)
            const int slot = (ck / TT) * FPS_T + (ck % TT);
            const float4 w = sPts[slot];
            ox = w.x; oy = w.y; oz = w.z;
            if (tid == 0) {
                out[j] = ck;
                if (oxyz != nullptr) { oxyz[j * 3] = ox; oxyz[j * 3 + 1] = oy; oxyz[j * 3 + 2] = oz; }
            }
        } else {
            if (warp == 0) {
                // only warp 0 needs this CTA's winner: it ships it to every CTA of the cluster
                int2 r = lane < FPS_W ? wrec[par][lane] : make_int2(INT_MIN, -1);
                const int cv = __reduce_max_sync(FULL_MASK, r.x);
                const unsigned ct = __reduce_min_sync(FULL_MASK, r.x == cv ? (unsigned)r.y : 0xffffffffu);
                const int ck = (int)((ct & 0x3fffffu) << bs_log2) | (int)(ct >> 22);
                FPS_STAMP(3)
                if (lane < C) {
                    const int slot = (ck / TT) * FPS_T + (ck % TT - rank * FPS_T);
                    const float4 w = sPts[slot];
                    // lane c delivers the record to CTA c; each 16-byte st.async also completes
                    // 16 transaction bytes on CTA c's barrier (point-to-point: no cluster-wide
                    // barrier and no fence on the critical path)
                    const unsigned dst = map_to_cta(smem_u32(&crec[par][rank]), lane);
                    const unsigned bar = map_to_cta(smem_u32(&cbar[par]), lane);
                    st_async_v4(dst, make_float4(__int_as_float(cv), __uint_as_float(ct), 0.f, 0.f), bar);
                    st_async_v4(dst + 16, make_float4(w.x, w.y, w.z, __int_as_float(ck)), bar);
                }
            }
            FPS_STAMP(4)
            mbar_wait(&cbar[par], (unsigned)((j - 1) >> 1) & 1u);  // u-th use of this barrier
            if (tid == 0) mbar_arm(&cbar[par], C * (unsigned)sizeof(FpsRec));  // arm its next use
            FPS_STAMP(5)
            // cluster winner, lane-parallel: lane c looks at CTA c's (value, tie-key)
            int v = INT_MIN;
            unsigned t = 0xffffffffu;
            if (lane < C) {
                const int2 vt = *reinterpret_cast<const int2 *>(&crec[par][lane]);
                v = vt.x;
                t = (unsigned)vt.y;
            }
            const int bv = __reduce_max_sync(FULL_MASK, v);
            const unsigned bt = __reduce_min_sync(FULL_MASK, v == bv ? t : 0xffffffffu);
            const unsigned who = __ballot_sync(FULL_MASK, v == bv && t == bt);  // tie-keys are unique
            const float4 w = *(reinterpret_cast<const float4 *>(&crec[par][__ffs(who) - 1]) + 1);
            ox = w.x; oy = w.y; oz = w.z;
            if (tg == 0) {
                out[j] = __float_as_int(w.w);
                if (oxyz != nullptr) { oxyz[j * 3] = ox; oxyz[j * 3 + 1] = oy; oxyz[j * 3 + 2] = oz; }
            }
            FPS_STAMP(6)
        }
    }
#ifdef PP_FPS_PROFILE
    if (prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0)
        for (int i = 0; i < 8; i++) prof[i] = pt[i];
#endif
#undef FPS_STAMP
#pragma unroll
    for (int p = 0; p < P; p++) {
        const int k = tg + p * TT;
        if (k < N) tp[k] = td[p];
    }
    if (C > 1) cg::this_cluster().sync();  // nobody exits while a peer may still write into it
}

// Fallback for clouds too large for the register-resident kernel: one 1024-thread CTA per
// cloud streaming points and running minima from global/L2 each round.  Same arithmetic,
// same tie-key; only meant to keep every size correct.
__global__ void __launch_bounds__(1024, 1)
fps_stream_kernel(const float *__restrict__ xyz, int N, int m, int seed, float *__restrict__ temp,
                  int *__restrict__ idx, float *__restrict__ new_xyz, int bs_log2) {
    __shared__ __align__(8) int2 wrec[2][32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *pts = xyz + (size_t)b * N * 3;
    float *tp = temp + (size_t)b * N;
    int *out = idx + (size_t)b * m;
    float *oxyz = new_xyz != nullptr ? new_xyz + (size_t)b * m * 3 : nullptr;
    int old = seed;
    if (tid == 0) out[0] = seed;
    for (int j = 1; j <= m; j++) {
        const int par = j & 1;
        const float ox = __ldg(pts + (size_t)old * 3), oy = __ldg(pts + (size_t)old * 3 + 1),
                    oz = __ldg(pts + (size_t)old * 3 + 2);
        if (tid == 0 && oxyz != nullptr) {  // coordinates of sample j-1
            oxyz[(j - 1) * 3] = ox; oxyz[(j - 1) * 3 + 1] = oy; oxyz[(j - 1) * 3 + 2] = oz;
        }
        if (j == m) break;
        float best = -1.f;
        unsigned btk = 0xffffffffu;
        // 1024 is a multiple of bs, so (k mod bs) is constant per thread and k ascends
        for (int k = tid; k < N; k += 1024) {
            const float d = sqdist_yxz(__fsub_rn(__ldg(pts + (size_t)k * 3), ox),
                                       __fsub_rn(__ldg(pts + (size_t)k * 3 + 1), oy),
                                       __fsub_rn(__ldg(pts + (size_t)k * 3 + 2), oz));
            const float t0 = tp[k];
            const float d2 = fminf(d, t0);
            if (d2 != t0) tp[k] = d2;
            if (d2 > best) {
                best = d2;
                btk = fps_tiekey(k, bs_log2);
            }
        }
        const int myv = __float_as_int(best);
        const int wv = __reduce_max_sync(FULL_MASK, myv);
        const unsigned wt = __reduce_min_sync(FULL_MASK, myv == wv ? btk : 0xffffffffu);
        if (lane == 0) wrec[par][warp] = make_int2(wv, (int)wt);
        __syncthreads();
        const int2 r = wrec[par][lane];
        const int cv = __reduce_max_sync(FULL_MASK, r.x);
        const unsigned ct = __reduce_min_sync(FULL_MASK, r.x == cv ? (unsigned)r.y : 0xffffffffu);
        old = (int)((ct & 0x3fffffu) << bs_log2) | (int)(ct >> 22);
        if (tid == 0) out[j] = old;
    }
}

// ===========================================================================
// gather / group (_ext/sampling_cuda.cu:9-84, 447-514): flat one-element-per-thread copies.
// ===========================================================================
__global__ void __launch_bounds__(256)
gather_fwd_kernel(const float *__restrict__ points, const int *__restrict__ idx, int C, int N,
                  int npoint, long long total, float *__restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int j = (int)(t % npoint);
    const long long bc = t / npoint;  // b*C + c
    const int b = (int)(bc / C);
    out[t] = __ldg(points + bc * N + __ldg(idx + (size_t)b * npoint + j));
}

__global__ void __launch_bounds__(256)
gather_bwd_kernel(const float *__restrict__ grad_out, const int *__restrict__ idx, int C, int N,
                  int npoint, long long total, float *__restrict__ grad_points) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int j = (int)(t % npoint);
    const long long bc = t / npoint;
    const int b = (int)(bc / C);
    atomicAdd(grad_points + bc * N + __ldg(idx + (size_t)b * npoint + j), __ldg(grad_out + t));
}

__global__ void __launch_bounds__(256)
group_fwd_kernel(const float *__restrict__ points, const int *__restrict__ idx, int C, int N,
                 int per_cloud /* npoint*nsample */, long long total, float *__restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int js = (int)(t % per_cloud);
    const long long bc = t / per_cloud;
    const int b = (int)(bc / C);
    out[t] = __ldg(points + bc * N + __ldg(idx + (size_t)b * per_cloud + js));
}

__global__ void __launch_bounds__(256)
group_bwd_kernel(const float *__restrict__ grad_out, const int *__restrict__ idx, int C, int N,
                 int per_cloud, long long total, float *__restrict__ grad_points) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int js = (int)(t % per_cloud);
    const long long bc = t / per_cloud;
    const int b = (int)(bc / C);
    atomicAdd(grad_points + bc * N + __ldg(idx + (size_t)b * per_cloud + js), __ldg(grad_out + t));
}

// ===========================================================================
// ball_query (_ext/sampling_cuda.cu:340-376): the reference lets ONE thread scan all N
// points serially per centre.  Here a WARP owns a centre: 32 lanes test 32 consecutive
// points per step, a ballot + prefix-popcount compacts the hits in ascending index
// order, and the scan stops as soon as nsample hits exist.  First-hit padding and the
// all-zero empty ball are written by the same warp, so no pre-zeroed output is needed.
// ===========================================================================
constexpr int BQ_WARPS = 8;

__global__ void __launch_bounds__(BQ_WARPS * 32)
ball_query_kernel(const float *__restrict__ new_xyz, const float *__restrict__ xyz, int N, int M,
                  float r2, int nsample, int *__restrict__ idx) {
    const int b = blockIdx.y;
    const int j = blockIdx.x * BQ_WARPS + (threadIdx.x >> 5);
    if (j >= M) return;
    const int lane = threadIdx.x & 31;
    const float *q = new_xyz + ((size_t)b * M + j) * 3;
    const float nx = __ldg(q), ny = __ldg(q + 1), nz = __ldg(q + 2);
    int *out = idx + ((size_t)b * M + j) * nsample;
    int first;
    int cnt = ball_scan(xyz + (size_t)b * N * 3, N, nx, ny, nz, r2, nsample, first,
                        [&](int pos, int k) { out[pos] = k; });
    if (cnt > nsample) cnt = nsample;
    // slots cnt.. hold the first hit (:366-370); an empty ball stays all zero (sampling.cpp:93-94)
    for (int l = cnt + lane; l < nsample; l += 32) out[l] = first;
}

// ===========================================================================
// three_nn (_ext/interpolate_gpu.cu:9-52): thread per unknown point, known points staged
// through shared memory.  The reference keeps its three bests as double initialised to
// 1e40; every stored value is an exact float and (double)d < 1e40 <=> d < +inf for all
// non-NaN d, so float bests initialised to +inf give identical results and outputs.
// ===========================================================================
// Four known points per step: their coordinates come as three broadcast LDS.128, the four distances as
// packed FADD2 / FMUL2 / FFMA2 (same roundings per lane as the scalar chain; (k - u)^2 == (u - k)^2 bit for
// bit), and ONE comparison of their minimum against the current third best decides whether any of them can
// enter the list -- only then are they inserted one by one, in index order, with the reference's strict '<'.
constexpr int NN3_TILE = 1024;

__global__ void __launch_bounds__(256)
three_nn_kernel(const float *__restrict__ unknown, const float *__restrict__ known, int N, int M,
                float *__restrict__ dist2, int *__restrict__ idx) {
    __shared__ __align__(16) float sx[NN3_TILE], sy[NN3_TILE], sz[NN3_TILE];
    const int b = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = j < N;
    const float *u = unknown + ((size_t)b * N + (active ? j : 0)) * 3;
    const float ux = u[0], uy = u[1], uz = u[2];
    const float *kn = known + (size_t)b * M * 3;
    float b1 = PP_INF, b2 = PP_INF, b3 = PP_INF;
    int i1 = 0, i2 = 0, i3 = 0;
    auto insert = [&](float d, int k) {
        if (d < b1) {
            b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k;
        } else if (d < b2) {
            b3 = b2; i3 = i2; b2 = d; i2 = k;
        } else if (d < b3) {
            b3 = d; i3 = k;
        }
    };
    for (int k0 = 0; k0 < M; k0 += NN3_TILE) {
        const int cnt = min(NN3_TILE, M - k0), cnt4 = (cnt + 3) & ~3;
        __syncthreads();
        for (int t = threadIdx.x; t < cnt4; t += blockDim.x) {
            const bool in = t < cnt;  // the padding sits at +inf: its distance is +inf, never below a best
            sx[t] = in ? __ldg(kn + (size_t)(k0 + t) * 3) : PP_INF;
            sy[t] = in ? __ldg(kn + (size_t)(k0 + t) * 3 + 1) : 0.f;
            sz[t] = in ? __ldg(kn + (size_t)(k0 + t) * 3 + 2) : 0.f;
        }
        __syncthreads();
#pragma unroll 2
        for (int k = 0; k < cnt4; k += 4) {
            const float4 x = *reinterpret_cast<const float4 *>(sx + k), y = *reinterpret_cast<const float4 *>(sy + k),
                         z = *reinterpret_cast<const float4 *>(sz + k);
            const float2 da = sqdist2_yxz(make_float2(x.x, x.y), make_float2(y.x, y.y), make_float2(z.x, z.y), -ux, -uy, -uz);
            const float2 db = sqdist2_yxz(make_float2(x.z, x.w), make_float2(y.z, y.w), make_float2(z.z, z.w), -ux, -uy, -uz);
            if (fminf(fmin3(da.x, da.y, db.x), db.y) < b3) {
                insert(da.x, k0 + k); insert(da.y, k0 + k + 1); insert(db.x, k0 + k + 2); insert(db.y, k0 + k + 3);
            }
        }
    }
    if (active) {
        float *od = dist2 + ((size_t)b * N + j) * 3;
        int *oi = idx + ((size_t)b * N + j) * 3;
        od[0] = b1; od[1] = b2; od[2] = b3;
        oi[0] = i1; oi[1] = i2; oi[2] = i3;
    }
}

// ===========================================================================
// three_interpolate forward / backward (_ext/interpolate_gpu.cu:77-97,120-142).  The reference runs one
// thread per output element (index arithmetic, three index and three weight reads for every channel).
// Here a thread owns a point: its three indices and weights are read once (a warp reads 384 contiguous
// bytes of each), then it walks TI_CH channels -- three gathers from a channel row that sits in L1
// (m floats), one coalesced store per channel.  HBM-bound on the (B,C,n) output.
// Forward rounding order = what nvcc makes of w0*p0 + w1*p1 + w2*p2:  fma(w2,p2, fma(w0,p0, rn(w1*p1)))
// -- middle product first, the same pattern as the 3-term distances (verified bit for bit against the
// reference kernel).
// ===========================================================================
constexpr int TI_CH = 16;  // channels per thread

__global__ void __launch_bounds__(256)
three_interpolate_fwd_kernel(const float *__restrict__ points, const int *__restrict__ idx,
                             const float *__restrict__ weight, int C, int M, int N, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int b = blockIdx.z, c0 = blockIdx.y * TI_CH;
    const int *id = idx + ((size_t)b * N + i) * 3;
    const float *w = weight + ((size_t)b * N + i) * 3;
    const int i0 = __ldg(id), i1 = __ldg(id + 1), i2 = __ldg(id + 2);
    const float w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    const float *p = points + ((size_t)b * C + c0) * M;
    float *o = out + ((size_t)b * C + c0) * N + i;
    const int nc = min(TI_CH, C - c0);
    if (nc == TI_CH) {
#pragma unroll
        for (int c = 0; c < TI_CH; c++)
            __stcs(o + (size_t)c * N, __fmaf_rn(w2, __ldg(p + (size_t)c * M + i2),
                                                __fmaf_rn(w0, __ldg(p + (size_t)c * M + i0), __fmul_rn(w1, __ldg(p + (size_t)c * M + i1)))));
    } else {
        for (int c = 0; c < nc; c++)
            o[(size_t)c * N] = __fmaf_rn(w2, __ldg(p + (size_t)c * M + i2),
                                         __fmaf_rn(w0, __ldg(p + (size_t)c * M + i0), __fmul_rn(w1, __ldg(p + (size_t)c * M + i1))));
    }
}

__global__ void __launch_bounds__(256)
three_interpolate_bwd_kernel(const float *__restrict__ grad_out, const int *__restrict__ idx,
                             const float *__restrict__ weight, int C, int M, int N, float *__restrict__ grad_points) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int b = blockIdx.z, c0 = blockIdx.y * TI_CH;
    const int *id = idx + ((size_t)b * N + i) * 3;
    const float *w = weight + ((size_t)b * N + i) * 3;
    const int i0 = __ldg(id), i1 = __ldg(id + 1), i2 = __ldg(id + 2);
    const float w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    float *g = grad_points + ((size_t)b * C + c0) * M;
    const float *go = grad_out + ((size_t)b * C + c0) * N + i;
    const int nc = min(TI_CH, C - c0);
#pragma unroll 4
    for (int c = 0; c < nc; c++) {
        const float v = __ldg(go + (size_t)c * N);
        atomicAdd(g + (size_t)c * M + i0, __fmul_rn(v, w0));
        atomicAdd(g + (size_t)c * M + i1, __fmul_rn(v, w1));
        atomicAdd(g + (size_t)c * M + i2, __fmul_rn(v, w2));
    }
}

template <int P, int C>
int launch_fps_cluster(const float *xyz, int B, int N, int m, int seed, float *temp, int *idx,
                       float *new_xyz, int bs_log2, cudaStream_t st, int *max_clusters) {
    const size_t smem = sizeof(float4) * P * FPS_T;
    auto kern = fps_cluster_kernel<P, C>;
    PP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * C);
    cfg.blockDim = dim3(FPS_T);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (max_clusters) {
        // query only: how many such clusters can be co-resident with ONE CTA per SM (two CTAs
        // sharing an SM's issue slots lengthen every round of the latency chain, measured
        // 0.72 vs 0.65 ms).  Asking with more than half an SM's shared memory enforces that.
        if (C == 1) {
            *max_clusters = NUM_SMS_B200;
        } else {
            const size_t solo = smem > (size_t)120 * 1024 ? smem : (size_t)120 * 1024;
            PP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)solo));
            cfg.dynamicSmemBytes = solo;
            PP_CUDA(cudaOccupancyMaxActiveClusters(max_clusters, kern, &cfg));
            PP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        return PP_OK;
    }
    KernelTimer timer("fps", st);
    long long *prof = (long long *)(uintptr_t)(unsigned long long)get_option("fps_prof_ptr_lo", 0);
    if (prof != nullptr)
        prof = (long long *)(((unsigned long long)(unsigned)get_option("fps_prof_ptr_hi", 0) << 32) |
                             (unsigned long long)(unsigned)get_option("fps_prof_ptr_lo", 0));
    PP_CUDA(cudaLaunchKernelEx(&cfg, kern, xyz, N, m, seed, temp, idx, new_xyz, bs_log2, prof));
    return PP_OK;
}

template <int C>
int dispatch_fps_p(int P, const float *xyz, int B, int N, int m, int seed, float *temp, int *idx,
                   float *new_xyz, int bs_log2, cudaStream_t st, int *max_clusters) {
    if (P <= 1) return launch_fps_cluster<1, C>(xyz, B, N, m, seed, temp, idx, new_xyz, bs_log2, st, max_clusters);
    if (P <= 2) return launch_fps_cluster<2, C>(xyz, B, N, m, seed, temp, idx, new_xyz, bs_log2, st, max_clusters);
    if (P <= 4) return launch_fps_cluster<4, C>(xyz, B, N, m, seed, temp, idx, new_xyz, bs_log2, st, max_clusters);
    if (P <= 8) return launch_fps_cluster<8, C>(xyz, B, N, m, seed, temp, idx, new_xyz, bs_log2, st, max_clusters);
    return launch_fps_cluster<16, C>(xyz, B, N, m, seed, temp, idx, new_xyz, bs_log2, st, max_clusters);
}

int dispatch_fps(int C, int P, const float *xyz, int B, int N, int m, int seed, float *temp, int *idx,
                 float *new_xyz, int bs_log2, cudaStream_t st, int *max_clusters) {
    switch (C) {
        case 1: return dispatch_fps_p<1>(P, xyz, B, N, m, seed, temp, idx, new_xyz, bs_log2, st, max_clusters);
        case 2: return dispatch_fps_p<2>(P, xyz, B, N, m, seed, temp, idx, new_xyz, bs_log2, st, max_clusters);
        case 4: return dispatch_fps_p<4>(P, xyz, B, N, m, seed, temp, idx, new_xyz, bs_log2, st, max_clusters);
        case 8: return dispatch_fps_p<8>(P, xyz, B, N, m, seed, temp, idx, new_xyz, bs_log2, st, max_clusters);
        default: break;
    }
    set_error("fps: unsupported cluster width %d", C);
    return PP_EINVAL;
}

}  // namespace
}  // namespace pp

using namespace pp;

// launch shape of the last FPS call on this thread (pp_fps_last_plan): cluster width (0 = streaming
// fallback) and points per thread
static thread_local int g_fps_last_cluster = 0, g_fps_last_points = 0;

static int fps_impl(const float *xyz, int B, int N, int m, int seed, float *temp, int32_t *idx,
                    float *new_xyz, int device, void *stream) {
    PP_REQUIRE(B >= 0 && N >= 1, "fps: bad sizes B=%d N=%d", B, N);
    if (m <= 0 || B == 0) return PP_OK;  // _ext/sampling_cuda.cu:166
    PP_REQUIRE(xyz && temp && idx, "fps: null pointer");
    PP_REQUIRE(seed >= 0 && seed < N, "fps: seedIdx %d outside [0,%d)", seed, N);
    PP_REQUIRE(N < (1 << 22), "fps: N=%d too large", N);
    DeviceGuard guard(device);
    PP_CUDA(guard.err);
    cudaStream_t st = (cudaStream_t)stream;
    // bs = max(min(2^floor(log2 N), 512), 1)   (_ext/cuda_utils.h:11-16)
    int bs_log2 = 0;
    while ((2 << bs_log2) <= N && bs_log2 < 9) bs_log2++;
    // Cluster width: FPS is a latency chain, so spread each cloud over as many SMs as possible
    // -- but only as long as ALL clouds' clusters are co-resident (a second wave doubles the
    // time).  The driver knows how many clusters of each width fit (GPC topology).
    int C = get_option("fps_cluster", 0);
    if (C == 0) {
        C = 1;
        for (int c = 8; c > 1; c >>= 1) {
            if (N < c * FPS_T) continue;  // not enough points to give every thread one
            const int p = ceil_div(N, c * FPS_T);
            if (p > 16) continue;
            // occupancy answers are cached per (device, width, points-per-thread bucket)
            static int fit_cache[16][4][17];
            int pb = p <= 1 ? 1 : p <= 2 ? 2 : p <= 4 ? 4 : p <= 8 ? 8 : 16;
            int ci = c == 8 ? 3 : c == 4 ? 2 : 1;
            int &slot = fit_cache[device & 15][ci][pb];
            if (slot == 0) {
                int fit = 0;
                const int rc = dispatch_fps(c, p, xyz, B, N, m, seed, temp, idx, new_xyz, bs_log2, st, &fit);
                if (rc != PP_OK) return rc;
                slot = fit + 1;
            }
            const int fit = slot - 1;
            if (get_option("fps_verbose", 0)) fprintf(stderr, "[pp_fps] cluster %d x P %d: %d co-resident\n", c, p, fit);
            if (fit >= B) {
                C = c;
                break;
            }
        }
        if (C == 1 && ceil_div(N, FPS_T) > 16) {
            // too large for one CTA's registers: take the widest cluster that holds the cloud
            for (int c = 2; c <= 8; c <<= 1)
                if (ceil_div(N, c * FPS_T) <= 16) { C = c; break; }
        }
    }
    const int P = ceil_div(N, C * FPS_T);
    g_fps_last_cluster = (P > 16 || get_option("fps_stream", 0)) ? 0 : C;
    g_fps_last_points = P;
    if (P > 16 || get_option("fps_stream", 0)) {
        KernelTimer timer("fps", st);
        fps_stream_kernel<<<B, 1024, 0, st>>>(xyz, N, m, seed, temp, idx, new_xyz, bs_log2);
        PP_LAUNCH_CHECK();
        return PP_OK;
    }
    return dispatch_fps(C, P, xyz, B, N, m, seed, temp, idx, new_xyz, bs_log2, st, nullptr);
}

extern "C" int pp_fps_last_plan(int *cluster_width, int *points_per_thread) {
    if (cluster_width) *cluster_width = g_fps_last_cluster;
    if (points_per_thread) *points_per_thread = g_fps_last_points;
    return PP_OK;
}

extern "C" int pp_fps(const float *xyz, int B, int N, int m, int seed, float *temp, int32_t *idx,
                      int device, void *stream) {
    return fps_impl(xyz, B, N, m, seed, temp, idx, nullptr, device, stream);
}

extern "C" int pp_fps_gather(const float *xyz, int B, int N, int m, int seed, float *temp, int32_t *idx,
                             float *new_xyz, int device, void *stream) {
    PP_REQUIRE(new_xyz || m <= 0 || B == 0, "fps_gather: null new_xyz");
    return fps_impl(xyz, B, N, m, seed, temp, idx, new_xyz, device, stream);
}

extern "C" int pp_gather_fwd(const float *points, const int32_t *idx, int B, int C, int N, int npoint,
                             float *out, int device, void *stream) {
    PP_REQUIRE(B >= 0 && C >= 0 && N >= 0 && npoint >= 0, "gather_fwd: bad sizes");
    const long long total = (long long)B * C * npoint;
    if (total == 0) return PP_OK;
    PP_REQUIRE(points && idx && out && N > 0, "gather_fwd: null pointer or empty source");
    DeviceGuard guard(device);
    PP_CUDA(guard.err);
    KernelTimer timer("gather_fwd", (cudaStream_t)stream);
    gather_fwd_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(points, idx, C, N, npoint, total, out);
    PP_LAUNCH_CHECK();
    return PP_OK;
}

extern "C" int pp_gather_bwd(const float *grad_out, const int32_t *idx, int B, int C, int N, int npoint,
                             float *grad_points, int device, void *stream) {
    PP_REQUIRE(B >= 0 && C >= 0 && N >= 0 && npoint >= 0, "gather_bwd: bad sizes");
    const long long total = (long long)B * C * npoint;
    if (total == 0) return PP_OK;
    PP_REQUIRE(grad_out && idx && grad_points && N > 0, "gather_bwd: null pointer or empty target");
    DeviceGuard guard(device);
    PP_CUDA(guard.err);
    gather_bwd_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(grad_out, idx, C, N, npoint, total, grad_points);
    PP_LAUNCH_CHECK();
    return PP_OK;
}

extern "C" int pp_group_fwd(const float *points, const int32_t *idx, int B, int C, int N, int npoint,
                            int nsample, float *out, int device, void *stream) {
    PP_REQUIRE(B >= 0 && C >= 0 && N >= 0 && npoint >= 0 && nsample >= 0, "group_fwd: bad sizes");
    PP_REQUIRE((long long)npoint * nsample < (1ll << 31), "group_fwd: npoint*nsample too large");
    const long long total = (long long)B * C * npoint * nsample;
    if (total == 0) return PP_OK;
    PP_REQUIRE(points && idx && out && N > 0, "group_fwd: null pointer or empty source");
    DeviceGuard guard(device);
    PP_CUDA(guard.err);
    group_fwd_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(points, idx, C, N, npoint * nsample, total, out);
    PP_LAUNCH_CHECK();
    return PP_OK;
}

extern "C" int pp_group_bwd(const float *grad_out, const int32_t *idx, int B, int C, int N, int npoint,
                            int nsample, float *grad_points, int device, void *stream) {
    PP_REQUIRE(B >= 0 && C >= 0 && N >= 0 && npoint >= 0 && nsample >= 0, "group_bwd: bad sizes");
    PP_REQUIRE((long long)npoint * nsample < (1ll << 31), "group_bwd: npoint*nsample too large");
    const long long total = (long long)B * C * npoint * nsample;
    if (total == 0) return PP_OK;
    PP_REQUIRE(grad_out && idx && grad_points && N > 0, "group_bwd: null pointer or empty target");
    DeviceGuard guard(device);
    PP_CUDA(guard.err);
    group_bwd_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(grad_out, idx, C, N, npoint * nsample, total, grad_points);
    PP_LAUNCH_CHECK();
    return PP_OK;
}

extern "C" int pp_ball_query(const float *new_xyz, const float *xyz, int B, int N, int M, float radius,
                             int nsample, int32_t *idx, int device, void *stream) {
    PP_REQUIRE(B >= 0 && N >= 0 && M >= 0 && nsample >= 0, "ball_query: bad sizes");
    if (B == 0 || M == 0 || nsample == 0) return PP_OK;
    PP_REQUIRE(new_xyz && idx && (xyz || N == 0), "ball_query: null pointer");
    PP_REQUIRE(B <= 65535, "ball_query: B=%d too large", B);
    DeviceGuard guard(device);
    PP_CUDA(guard.err);
    const float r2 = radius * radius;  // rn(r*r) in fp32 (:354); host float multiply is the same single rounding
    dim3 grid(ceil_div(M, BQ_WARPS), B);
    KernelTimer timer("ball_query", (cudaStream_t)stream);
    ball_query_kernel<<<grid, BQ_WARPS * 32, 0, (cudaStream_t)stream>>>(new_xyz, xyz, N, M, r2, nsample, idx);
    PP_LAUNCH_CHECK();
    return PP_OK;
}

extern "C" int pp_three_nn(const float *unknown, const float *known, int B, int N, int M, float *dist2,
                           int32_t *idx, int device, void *stream) {
    PP_REQUIRE(B >= 0 && N >= 0 && M >= 0, "three_nn: bad sizes");
    if (B == 0 || N == 0) return PP_OK;
    PP_REQUIRE(unknown && dist2 && idx && (known || M == 0), "three_nn: null pointer");
    PP_REQUIRE(B <= 65535, "three_nn: B=%d too large", B);
    DeviceGuard guard(device);
    PP_CUDA(guard.err);
    dim3 grid(ceil_div(N, 256), B);
    three_nn_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(unknown, known, N, M, dist2, idx);
    PP_LAUNCH_CHECK();
    return PP_OK;
}

extern "C" int pp_three_interpolate_fwd(const float *points, const int32_t *idx, const float *weight, int B,
                                        int C, int M, int N, float *out, int device, void *stream) {
    PP_REQUIRE(B >= 0 && C >= 0 && M >= 0 && N >= 0, "three_interpolate_fwd: bad sizes");
    const long long total = (long long)B * C * N;
    if (total == 0) return PP_OK;
    PP_REQUIRE(points && idx && weight && out && M > 0, "three_interpolate_fwd: null pointer or empty source");
    DeviceGuard guard(device);
    PP_CUDA(guard.err);
    PP_REQUIRE(B <= 65535 && ceil_div(C, TI_CH) <= 65535, "three_interpolate_fwd: B or C too large");
    {
        KernelTimer timer("three_interpolate", (cudaStream_t)stream);
        three_interpolate_fwd_kernel<<<dim3(ceil_div(N, 256), ceil_div(C, TI_CH), B), 256, 0, (cudaStream_t)stream>>>(points, idx, weight, C, M, N, out);
    }
    PP_LAUNCH_CHECK();
    return PP_OK;
}

extern "C" int pp_three_interpolate_bwd(const float *grad_out, const int32_t *idx, const float *weight, int B,
                                        int C, int N, int M, float *grad_points, int device, void *stream) {
    PP_REQUIRE(B >= 0 && C >= 0 && M >= 0 && N >= 0, "three_interpolate_bwd: bad sizes");
    const long long total = (long long)B * C * N;
    if (total == 0) return PP_OK;
    PP_REQUIRE(grad_out && idx && weight && grad_points && M > 0, "three_interpolate_bwd: null pointer or empty target");
    DeviceGuard guard(device);
    PP_CUDA(guard.err);
    PP_REQUIRE(B <= 65535 && ceil_div(C, TI_CH) <= 65535, "three_interpolate_bwd: B or C too large");
    three_interpolate_bwd_kernel<<<dim3(ceil_div(N, 256), ceil_div(C, TI_CH), B), 256, 0, (cudaStream_t)stream>>>(grad_out, idx, weight, C, M, N, grad_points);
    PP_LAUNCH_CHECK();
    return PP_OK;
}
