// Warp-wide ball query scan shared by pp_ball_query (sampling.cu) and the fused QueryAndGroup
// kernel (sa_group.cu).  Semantics of _ext/sampling_cuda.cu:346-375: ascending index scan,
// keep the first `nsample` points with d2 < r2 (strict; d2 in the reference's y-first rounding
// order), stop as soon as they are found.
#pragma once
#include "pp_common.cuh"

namespace pp {

constexpr int BS_UNROLL = 4;  // 32 * 4 points per trip

// One warp, all 32 lanes call.  `put(pos, k)` stores hit number `pos` (< nsample), called by the
// lane that owns point k.  Returns the number of hits (may exceed nsample by less than a trip)
// and the first hit in `first`.  The next trip's coordinates are requested before the current
// trip is evaluated, so the L2 round trip overlaps the ballots.
template <typename Put>
__device__ __forceinline__ int ball_scan(const float *__restrict__ p, int N, float nx, float ny, float nz,
                                         float r2, int nsample, int &first, Put put) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int cnt = 0;
    first = 0;
    float cx[BS_UNROLL], cy[BS_UNROLL], cz[BS_UNROLL];
    auto fetch = [&](int base, float (&x)[BS_UNROLL], float (&y)[BS_UNROLL], float (&z)[BS_UNROLL]) {
#pragma unroll
        for (int u = 0; u < BS_UNROLL; u++) {
            const int k = base + u * 32 + lane;
            x[u] = y[u] = z[u] = PP_INF;  // beyond the cloud: d2 = inf, never a hit
            if (k < N) {
                x[u] = __ldg(p + (size_t)k * 3);
                y[u] = __ldg(p + (size_t)k * 3 + 1);
                z[u] = __ldg(p + (size_t)k * 3 + 2);
            }
        }
    };
    fetch(0, cx, cy, cz);
    for (int base = 0; base < N && cnt < nsample; base += 32 * BS_UNROLL) {
        float fx[BS_UNROLL], fy[BS_UNROLL], fz[BS_UNROLL];
        if (base + 32 * BS_UNROLL < N) fetch(base + 32 * BS_UNROLL, fx, fy, fz);
#pragma unroll
        for (int u = 0; u < BS_UNROLL; u++) {
            const float d2 = sqdist_yxz(__fsub_rn(nx, cx[u]), __fsub_rn(ny, cy[u]), __fsub_rn(nz, cz[u]));
            const bool hit = d2 < r2;  // strict, NaN never matches (:365)
            const unsigned mask = __ballot_sync(FULL_MASK, hit);
            if (mask != 0u && cnt < nsample) {
                if (cnt == 0) first = base + u * 32 + __ffs(mask) - 1;
                const int pos = cnt + __popc(mask & lt);
                if (hit && pos < nsample) put(pos, base + u * 32 + lane);
                cnt += __popc(mask);
            }
        }
#pragma unroll
        for (int u = 0; u < BS_UNROLL; u++) {
            cx[u] = fx[u]; cy[u] = fy[u]; cz[u] = fz[u];
        }
    }
    return cnt;
}

}  // namespace pp
