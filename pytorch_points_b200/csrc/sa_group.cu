// Set-abstraction grouping stage, fused ("next" rows N1/N2 of SURVEY.md section 8f).
//
// The reference's QueryAndGroup (network/operations.py:166-213) runs, per SA scale:
//   ball_query -> idx (B,M,ns)           _ext/sampling_cuda.cu:340-398
//   transpose xyz -> (B,3,N) contiguous  operations.py:196
//   group_points(xyz^T, idx)             _ext/sampling_cuda.cu:447-479
//   grouped_xyz -= new_xyz^T[..., None]  operations.py:198
//   group_points(features, idx)          operations.py:201
//   cat([grouped_xyz, grouped_features]) operations.py:203
// i.e. six kernels, and every grouped value crosses HBM three times (write, read for the
// cat, write).  Here ONE kernel does all of it: a warp owns a centre, finds its ball with
// the ball_query scan (same arithmetic, same ascending order, same first-hit padding), keeps
// the nsample indices in shared memory, and streams the (3 + C) x nsample output tile
// straight into its final place -- each value is written exactly once, nsample contiguous
// floats per channel.  idx is also written (the backward pass scatters through it).
#include "ball_scan.cuh"
#include "pp_common.cuh"

namespace pp {
namespace {

constexpr int QG_WARPS = 8;

// POINT_MAJOR: the features come as (B, N, C) -- the C channels of one source point are contiguous, so a ball
// member is read as full 128-byte lines (one channel per lane) and a 32 x 32 tile is turned in shared memory
// into nsample-contiguous output rows.  With the reference's (B, C, N) layout every gathered value drags a
// 32-byte sector of its own through L2 (8x over-fetch); a multi-scale level stages its features once
// (pp_channels_to_points) and runs every scale on the staged copy.
constexpr int QG_TILE_PITCH = 33;

template <bool POINT_MAJOR>
__global__ void __launch_bounds__(QG_WARPS * 32)
query_group_kernel(const float *__restrict__ new_xyz, const float *__restrict__ xyz,
                   const float *__restrict__ features, int N, int M, int C, float r2, int nsample,
                   int use_xyz, int *__restrict__ idx, float *__restrict__ out) {
    extern __shared__ int s_idx[];  // [QG_WARPS][nsample] (+ [QG_WARPS][32][33] floats if POINT_MAJOR)
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * QG_WARPS + warp;
    if (j >= M) return;  // whole warps leave; no block-wide barrier below
    int *mine = s_idx + warp * nsample;
    const float *q = new_xyz + ((size_t)b * M + j) * 3;
    const float nx = __ldg(q), ny = __ldg(q + 1), nz = __ldg(q + 2);
    const float *p = xyz + (size_t)b * N * 3;

    // ---- ball query (_ext/sampling_cuda.cu:346-375): ascending scan, strict d2 < r2 ----
    int first;
    int cnt = ball_scan(p, N, nx, ny, nz, r2, nsample, first, [&](int pos, int k) { mine[pos] = k; });
    if (cnt > nsample) cnt = nsample;
    for (int l = cnt + lane; l < nsample; l += 32) mine[l] = first;  // first-hit padding / empty ball -> 0
    __syncwarp();

    // ---- emit: idx, centred coordinates, features ----
    const size_t plane = (size_t)M * nsample;  // one output channel of one cloud
    const int ch_total = (use_xyz ? 3 : 0) + C;
    int *oi = idx + ((size_t)b * M + j) * nsample;
    float *o = out + (size_t)b * ch_total * plane + (size_t)j * nsample;
    for (int s0 = 0; s0 < nsample; s0 += 32) {
        const int s = s0 + lane;
        const bool ok = s < nsample;
        const int k = ok ? mine[s] : 0;
        if (ok) oi[s] = k;
        float *os = o + s;
        if (use_xyz) {
            // grouped_xyz - new_xyz: one rounded subtraction per component (operations.py:198)
            const float x = __ldg(p + (size_t)k * 3), y = __ldg(p + (size_t)k * 3 + 1),
                        z = __ldg(p + (size_t)k * 3 + 2);
            if (ok) {
                __stcs(os, __fsub_rn(x, nx));
                __stcs(os + plane, __fsub_rn(y, ny));
                __stcs(os + 2 * plane, __fsub_rn(z, nz));
            }
            os += 3 * plane;
        }
        if (POINT_MAJOR) continue;  // (features below, tile by tile)
        if (C > 0 && ok) {
            const float *f = features + (size_t)b * C * N + k;
            int c = 0;
            for (; c + 4 <= C; c += 4) {  // four independent gathers in flight
                const float v0 = __ldg(f + (size_t)(c + 0) * N), v1 = __ldg(f + (size_t)(c + 1) * N),
                            v2 = __ldg(f + (size_t)(c + 2) * N), v3 = __ldg(f + (size_t)(c + 3) * N);
                __stcs(os + (size_t)(c + 0) * plane, v0);
                __stcs(os + (size_t)(c + 1) * plane, v1);
                __stcs(os + (size_t)(c + 2) * plane, v2);
                __stcs(os + (size_t)(c + 3) * plane, v3);
            }
            for (; c < C; c++) __stcs(os + (size_t)c * plane, __ldg(f + (size_t)c * N));
        }
    }
    if (POINT_MAJOR && C > 0) {
        float *tile = reinterpret_cast<float *>(s_idx + QG_WARPS * nsample) + warp * 32 * QG_TILE_PITCH;
        const float *f = features + (size_t)b * N * C;
        float *of = o + (use_xyz ? 3 : 0) * plane;
        for (int s0 = 0; s0 < nsample; s0 += 32) {
            const int ns = min(32, nsample - s0);
            for (int c0 = 0; c0 < C; c0 += 32) {
                const int nc = min(32, C - c0);
                __syncwarp();
                // member s of the chunk: channels c0 .. c0+31 of its point, one per lane
#pragma unroll 8
                for (int s = 0; s < ns; s++) {
                    const int k = mine[s0 + s];
                    if (lane < nc) tile[s * QG_TILE_PITCH + lane] = __ldg(f + (size_t)k * C + c0 + lane);
                }
                __syncwarp();
                // channel c of the chunk: its ns samples, one per lane, contiguous in the output
#pragma unroll 8
                for (int c = 0; c < nc; c++)
                    if (lane < ns) __stcs(of + (size_t)(c0 + c) * plane + s0 + lane, tile[lane * QG_TILE_PITCH + c]);
            }
        }
    }
}

// (B, C, N) -> (B, N, C) through 32 x 32 shared-memory tiles: both sides move full lines
__global__ void __launch_bounds__(256)
channels_to_points_kernel(const float *__restrict__ in, int C, int N, float *__restrict__ out) {
    __shared__ float t[32][33];
    const int b = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float *src = in + (size_t)b * C * N;
    float *dst = out + (size_t)b * N * C;
#pragma unroll
    for (int r = ty; r < 32; r += 8)
        if (c0 + r < C && n0 + tx < N) t[r][tx] = __ldg(src + (size_t)(c0 + r) * N + n0 + tx);
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8)
        if (n0 + r < N && c0 + tx < C) dst[(size_t)(n0 + r) * C + c0 + tx] = t[tx][r];
}

// Backward of the fused stage.  A warp owns a centre again; lane s owns sample s.
//   grad_features[b,c,idx[s]] += go[b,x+c,j,s]                     (group_points_grad, :482-514)
//   grad_xyz[b,idx[s],:]      += go[b,0:3,j,s]                     (same op on xyz^T, then transposed back)
//   grad_new_xyz[b,j,:]        = -sum_s go[b,0:3,j,s]              (the broadcast subtraction, operations.py:198)
__global__ void __launch_bounds__(QG_WARPS * 32)
query_group_bwd_kernel(const float *__restrict__ grad_out, const int *__restrict__ idx, int N, int M,
                       int C, int nsample, int use_xyz, float *__restrict__ grad_features,
                       float *__restrict__ grad_xyz, float *__restrict__ grad_new_xyz) {
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * QG_WARPS + warp;
    if (j >= M) return;
    const size_t plane = (size_t)M * nsample;
    const int ch_total = (use_xyz ? 3 : 0) + C;
    const int *ii = idx + ((size_t)b * M + j) * nsample;
    const float *g = grad_out + (size_t)b * ch_total * plane + (size_t)j * nsample;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int s = lane; s < nsample; s += 32) {
        const int k = __ldg(ii + s);
        const float *gs = g + s;
        if (use_xyz) {
            const float gx = __ldg(gs), gy = __ldg(gs + plane), gz = __ldg(gs + 2 * plane);
            if (grad_xyz != nullptr) {
                float *t = grad_xyz + ((size_t)b * N + k) * 3;
                atomicAdd(t, gx);
                atomicAdd(t + 1, gy);
                atomicAdd(t + 2, gz);
            }
            sx += gx; sy += gy; sz += gz;
            gs += 3 * plane;
        }
        if (grad_features != nullptr) {
            float *t = grad_features + (size_t)b * C * N + k;
            for (int c = 0; c < C; c++) atomicAdd(t + (size_t)c * N, __ldg(gs + (size_t)c * plane));
        }
    }
    if (use_xyz && grad_new_xyz != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sx += __shfl_xor_sync(FULL_MASK, sx, o);
            sy += __shfl_xor_sync(FULL_MASK, sy, o);
            sz += __shfl_xor_sync(FULL_MASK, sz, o);
        }
        if (lane == 0) {
            float *t = grad_new_xyz + ((size_t)b * M + j) * 3;
            t[0] = -sx; t[1] = -sy; t[2] = -sz;
        }
    }
}

}  // namespace
}  // namespace pp

using namespace pp;

static int query_group_fwd_impl(const float *new_xyz, const float *xyz, const float *features, int B, int N, int M,
                                int C, float radius, int nsample, int use_xyz, int32_t *idx, float *out, int device,
                                void *stream, bool point_major) {
    PP_REQUIRE(B >= 0 && N >= 0 && M >= 0 && C >= 0 && nsample >= 0, "query_group: bad sizes");
    PP_REQUIRE(use_xyz || C > 0, "query_group: nothing to group (use_xyz=0 and no features)");
    if (B == 0 || M == 0 || nsample == 0) return PP_OK;
    PP_REQUIRE(new_xyz && idx && out && (xyz || N == 0), "query_group: null pointer");
    PP_REQUIRE(C == 0 || features, "query_group: C=%d but features is null", C);
    PP_REQUIRE(N > 0, "query_group: empty source cloud");
    PP_REQUIRE(B <= 65535, "query_group: B=%d too large", B);
    const size_t smem = (size_t)QG_WARPS * nsample * sizeof(int) +
                        (point_major ? (size_t)QG_WARPS * 32 * QG_TILE_PITCH * sizeof(float) : 0);
    PP_REQUIRE(smem <= 48 * 1024, "query_group: nsample=%d too large", nsample);
    DeviceGuard guard(device);
    PP_CUDA(guard.err);
    const float r2 = radius * radius;  // rn(r*r) in fp32 (_ext/sampling_cuda.cu:354)
    dim3 grid(ceil_div(M, QG_WARPS), B);
    KernelTimer timer("query_group", (cudaStream_t)stream);
    if (point_major)
        query_group_kernel<true><<<grid, QG_WARPS * 32, smem, (cudaStream_t)stream>>>(
            new_xyz, xyz, features, N, M, C, r2, nsample, use_xyz ? 1 : 0, idx, out);
    else
        query_group_kernel<false><<<grid, QG_WARPS * 32, smem, (cudaStream_t)stream>>>(
            new_xyz, xyz, features, N, M, C, r2, nsample, use_xyz ? 1 : 0, idx, out);
    PP_LAUNCH_CHECK();
    return PP_OK;
}

extern "C" int pp_query_group_fwd(const float *new_xyz, const float *xyz, const float *features, int B,
                                  int N, int M, int C, float radius, int nsample, int use_xyz,
                                  int32_t *idx, float *out, int device, void *stream) {
    return query_group_fwd_impl(new_xyz, xyz, features, B, N, M, C, radius, nsample, use_xyz, idx, out, device, stream,
                                false);
}

extern "C" int pp_query_group_fwd_pm(const float *new_xyz, const float *xyz, const float *features_pm, int B,
                                     int N, int M, int C, float radius, int nsample, int use_xyz,
                                     int32_t *idx, float *out, int device, void *stream) {
    return query_group_fwd_impl(new_xyz, xyz, features_pm, B, N, M, C, radius, nsample, use_xyz, idx, out, device,
                                stream, true);
}

extern "C" int pp_channels_to_points(const float *in, int B, int C, int N, float *out, int device, void *stream) {
    PP_REQUIRE(B >= 0 && C >= 0 && N >= 0, "channels_to_points: bad sizes");
    if (B == 0 || C == 0 || N == 0) return PP_OK;
    PP_REQUIRE(in && out, "channels_to_points: null pointer");
    PP_REQUIRE(B <= 65535 && ceil_div(C, 32) <= 65535, "channels_to_points: B=%d or C=%d too large", B, C);
    DeviceGuard guard(device);
    PP_CUDA(guard.err);
    KernelTimer timer("channels_to_points", (cudaStream_t)stream);
    channels_to_points_kernel<<<dim3(ceil_div(N, 32), ceil_div(C, 32), B), 256, 0, (cudaStream_t)stream>>>(in, C, N, out);
    PP_LAUNCH_CHECK();
    return PP_OK;
}

extern "C" int pp_query_group_bwd(const float *grad_out, const int32_t *idx, int B, int N, int M, int C,
                                  int nsample, int use_xyz, float *grad_features, float *grad_xyz,
                                  float *grad_new_xyz, int device, void *stream) {
    PP_REQUIRE(B >= 0 && N >= 0 && M >= 0 && C >= 0 && nsample >= 0, "query_group_bwd: bad sizes");
    if (B == 0 || M == 0 || nsample == 0) return PP_OK;
    PP_REQUIRE(grad_out && idx && N > 0, "query_group_bwd: null pointer or empty source");
    PP_REQUIRE(B <= 65535, "query_group_bwd: B=%d too large", B);
    DeviceGuard guard(device);
    PP_CUDA(guard.err);
    dim3 grid(ceil_div(M, QG_WARPS), B);
    KernelTimer timer("query_group_bwd", (cudaStream_t)stream);
    query_group_bwd_kernel<<<grid, QG_WARPS * 32, 0, (cudaStream_t)stream>>>(
        grad_out, idx, N, M, C, nsample, use_xyz ? 1 : 0, C > 0 ? grad_features : nullptr, grad_xyz, grad_new_xyz);
    PP_LAUNCH_CHECK();
    return PP_OK;
}
