// pp_core.cu -- error reporting, options and the roofline micro-benchmarks.
#include <stdarg.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "pp_common.cuh"

namespace pp {

static thread_local char g_err[512] = "";
static std::mutex g_opt_mu;
static std::map<std::string, int> g_opts;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int get_option(const char *name, int dflt) {
    std::lock_guard<std::mutex> lk(g_opt_mu);
    auto it = g_opts.find(name);
    return it == g_opts.end() ? dflt : it->second;
}

struct TimingRec {
    cudaEvent_t e0, e1;
};
static std::mutex g_time_mu;
static std::map<std::string, std::vector<TimingRec>> g_timings;

KernelTimer::KernelTimer(const char *name_, cudaStream_t st_) : name(name_), st(st_) {
    on = get_option("timing", 0) != 0;
    if (on) {
        if (cudaEventCreate(&e0) != cudaSuccess) { on = false; return; }
        cudaEventRecord(e0, st);
    }
}

KernelTimer::~KernelTimer() {
    if (!on) return;
    cudaEvent_t e1;
    if (cudaEventCreate(&e1) != cudaSuccess) { cudaEventDestroy(e0); return; }
    cudaEventRecord(e1, st);
    std::lock_guard<std::mutex> lk(g_time_mu);
    g_timings[name].push_back({e0, e1});
}

namespace {

// ---- micro-benchmarks: measure the peaks MEASURED_PEAKS.json does not hold ----
// 0: scalar FFMA chain        (flop = 2 / FFMA)
// 1: packed FFMA2 chain       (flop = 4 / FFMA2)
// 2: Chamfer op mix, scalar   (3 FADD + FMUL + 2 FFMA + FMNMX per pair; work = pairs)
// 3: Chamfer op mix, packed   (3 FADD2 + FMUL2 + 2 FFMA2 + FMNMX3 per 2 pairs; work = pairs)
// 4: shared-memory read bandwidth (LDS.128, conflict free; work = bytes)
// 5: L2 read bandwidth (32 MB buffer re-read; work = bytes)
// 6: REDUX.MIN throughput (work = warp instructions)
template <int ILP>
__global__ void __launch_bounds__(256) mb_ffma(float *out, int iters) {
    float a[ILP];
    const float x = 1.0000001f + threadIdx.x * 1e-9f, y = 1e-7f;
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = i * 0.5f + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) a[i] = __fmaf_rn(a[i], x, y);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += a[i];
    if (s == 12345.678f) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) mb_ffma2(float *out, int iters) {
    float2 a[ILP];
    const float x = 1.0000001f + threadIdx.x * 1e-9f, y = 1e-7f;
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = make_float2(i * 0.5f + threadIdx.x, i * 0.25f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            asm("{\n\t.reg .b64 ra, rx, ry;\n\t"
                "mov.b64 ra, {%0, %1};\n\tmov.b64 rx, {%2, %2};\n\tmov.b64 ry, {%3, %3};\n\t"
                "fma.rn.f32x2 ra, ra, rx, ry;\n\tmov.b64 {%0, %1}, ra;\n\t}"
                : "+f"(a[i].x), "+f"(a[i].y)
                : "f"(x), "f"(y));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += a[i].x + a[i].y;
    if (s == 12345.678f) out[0] = s;
}

template <int Q>
__global__ void __launch_bounds__(256) mb_mix_scalar(float *out, int iters) {
    float qx[Q], qy[Q], qz[Q], best[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) {
        qx[q] = threadIdx.x * 0.001f + q;
        qy[q] = threadIdx.x * 0.002f - q;
        qz[q] = threadIdx.x * 0.003f;
        best[q] = PP_INF;
    }
    float rx = 0.1f, ry = 0.2f, rz = 0.3f;
    for (int it = 0; it < iters; it++) {
        rx += 0.001f;  // 3 extra FADD per Q pairs (stands in for the LDS feed)
        ry += 0.002f;
        rz += 0.003f;
#pragma unroll
        for (int q = 0; q < Q; q++) best[q] = fminf(best[q], sqdist_xyz(rx, ry, rz, qx[q], qy[q], qz[q]));
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < Q; q++) s += best[q];
    if (s == 12345.678f) out[0] = s;
}

template <int Q>
__global__ void __launch_bounds__(256) mb_mix_packed(float *out, int iters) {
    float qx[Q], qy[Q], qz[Q], best[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) {
        qx[q] = threadIdx.x * 0.001f + q;
        qy[q] = threadIdx.x * 0.002f - q;
        qz[q] = threadIdx.x * 0.003f;
        best[q] = PP_INF;
    }
    float2 rx = make_float2(0.1f, 0.4f), ry = make_float2(0.2f, 0.5f), rz = make_float2(0.3f, 0.6f);
    for (int it = 0; it < iters; it++) {
        rx.x += 0.001f;
        ry.y += 0.002f;
        rz.x += 0.003f;
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const float2 d = sqdist2_xyz(rx, ry, rz, qx[q], qy[q], qz[q]);
            best[q] = fmin3(best[q], d.x, d.y);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < Q; q++) s += best[q];
    if (s == 12345.678f) out[0] = s;
}

__global__ void __launch_bounds__(256) mb_smem(float *out, int iters) {
    __shared__ __align__(16) float buf[8192];
    for (int i = threadIdx.x; i < 8192; i += 256) buf[i] = i;
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    int idx = threadIdx.x * 4;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const float4 v = *reinterpret_cast<const float4 *>(buf + ((idx + u * 1024) & 8191));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        idx = (idx + 4) & 8191;
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
}

__global__ void __launch_bounds__(256) mb_l2(const float4 *__restrict__ src, float *out, size_t n4, int iters) {
    float4 acc = make_float4(0, 0, 0, 0);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int it = 0; it < iters; it++) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            const float4 v = __ldcg(src + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
}

__global__ void __launch_bounds__(256) mb_redux(float *out, int iters) {
    unsigned v = threadIdx.x * 2654435761u, acc = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            acc += __reduce_min_sync(FULL_MASK, v + u + acc);
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// three-input / two-input float minima with all-distinct source registers, eight independent chains
template <int KIND>
__global__ void __launch_bounds__(256) mb_fmnmx(float *out, int iters) {
    float a[8], x[8], y[8];
#pragma unroll
    for (int u = 0; u < 8; u++) { a[u] = 1e30f - threadIdx.x; x[u] = 3.f + u + threadIdx.x; y[u] = 5.f + 2 * u + threadIdx.x; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (KIND == 0) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[u]) : "f"(x[u]), "f"(y[u]));
            else if (KIND == 1) asm volatile("min.f32 %0, %0, %1;" : "+f"(a[u]) : "f"(x[u]));
            else asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(a[u]) : "f"(a[(u + 1) & 7]), "f"(x[u]), "f"(y[u]));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < 8; u++) s += a[u];
    if (s == 12345.678f) out[0] = s;
}

}  // namespace
}  // namespace pp

using namespace pp;

extern "C" int pp_version(void) { return 1; }

extern "C" const char *pp_last_error_string(void) { return g_err; }

extern "C" int pp_set_option(const char *name, int value) {
    if (!name) return PP_EINVAL;
    std::lock_guard<std::mutex> lk(g_opt_mu);
    g_opts[name] = value;
    return PP_OK;
}

extern "C" int pp_timing_collect(const char *name, double *total_ms, int *count) {
    PP_REQUIRE(name && total_ms && count, "timing_collect: null argument");
    std::vector<TimingRec> recs;
    {
        std::lock_guard<std::mutex> lk(g_time_mu);
        auto it = g_timings.find(name);
        if (it != g_timings.end()) {
            recs.swap(it->second);
            g_timings.erase(it);
        }
    }
    *total_ms = 0.0;
    *count = 0;
    for (auto &r : recs) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
            *total_ms += ms;
            *count += 1;
        }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    return PP_OK;
}

namespace pp {
namespace {
__global__ void dot2_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ out) {
    pdl_wait();
    out[0] = __fmaf_rn(a[1], b[1], __fmul_rn(a[0], b[0]));
}
}  // namespace
}  // namespace pp

extern "C" int pp_dot2(const float *a, const float *b, float *out, int device, void *stream) {
    PP_REQUIRE(a && b && out, "dot2: null pointer");
    DeviceGuard guard(device);
    PP_CUDA(guard.err);
    PP_CUDA(launch_pdl(pp::dot2_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, a, b, out));
    return PP_OK;
}

extern "C" int pp_memcpy_async(void *dst, const void *src, size_t bytes, int device, void *stream) {
    PP_REQUIRE(dst && src, "memcpy_async: null pointer");
    if (bytes == 0) return PP_OK;
    DeviceGuard guard(device);
    PP_CUDA(guard.err);
    PP_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return PP_OK;
}

extern "C" int pp_microbench(int which, int iters, float *ms, double *work, int device) {
    PP_REQUIRE(ms && work && iters > 0, "microbench: bad arguments");
    DeviceGuard guard(device);
    PP_CUDA(guard.err);
    float *out = nullptr;
    PP_CUDA(cudaMalloc(&out, 256));
    cudaEvent_t e0, e1;
    PP_CUDA(cudaEventCreate(&e0));
    PP_CUDA(cudaEventCreate(&e1));
    const int blocks = NUM_SMS_B200 * 8, threads = 256;
    const double lanes = (double)blocks * threads;
    float4 *big = nullptr;
    const size_t l2_bytes = 32u << 20;
    if (which == 5) {
        PP_CUDA(cudaMalloc(&big, l2_bytes));
        PP_CUDA(cudaMemset(big, 0, l2_bytes));
    }
    int rc = PP_OK;
    for (int rep = 0; rep < 2; rep++) {  // rep 0 warms up
        PP_CUDA(cudaEventRecord(e0));
        switch (which) {
            case 0: mb_ffma<16><<<blocks, threads>>>(out, iters); *work = lanes * iters * 16 * 2.0; break;
            case 1: mb_ffma2<8><<<blocks, threads>>>(out, iters); *work = lanes * iters * 8 * 4.0; break;
            case 2: mb_mix_scalar<8><<<blocks, threads>>>(out, iters); *work = lanes * iters * 8.0; break;
            case 3: mb_mix_packed<8><<<blocks, threads>>>(out, iters); *work = lanes * iters * 16.0; break;
            case 4: mb_smem<<<blocks, threads>>>(out, iters); *work = lanes * iters * 8 * 16.0; break;
            case 5: mb_l2<<<blocks, threads>>>(big, out, l2_bytes / 16, iters); *work = (double)l2_bytes * iters; break;
            case 6: mb_redux<<<blocks, threads>>>(out, iters); *work = lanes / 32 * iters * 8.0; break;
            case 9: mb_fmnmx<0><<<blocks, threads>>>(out, iters); *work = lanes / 32 * iters * 8.0; break;
            case 10: mb_fmnmx<1><<<blocks, threads>>>(out, iters); *work = lanes / 32 * iters * 8.0; break;
            case 11: mb_fmnmx<2><<<blocks, threads>>>(out, iters); *work = lanes / 32 * iters * 8.0; break;
            case 7: case 8: {
                const cudaError_t pe = chamfer_sweep_tmem_probe(which - 7, iters, out, work);
                if (pe != cudaSuccess) { set_error("microbench: %s", cudaGetErrorString(pe)); rc = (int)pe; }
                break;
            }
            default: set_error("microbench: unknown probe %d", which); rc = PP_EINVAL; break;
        }
        if (rc != PP_OK) break;
        PP_CUDA(cudaEventRecord(e1));
        PP_CUDA(cudaEventSynchronize(e1));
        PP_CUDA(cudaEventElapsedTime(ms, e0, e1));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (big) cudaFree(big);
    return rc;
}
