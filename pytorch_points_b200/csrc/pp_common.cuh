// pp_common.cuh -- shared helpers for the sm_100a kernels of libpp_b200.so.
// No torch headers anywhere in csrc/: the library is a plain C-ABI CUDA .so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pp_b200.h"

namespace pp {

constexpr unsigned FULL_MASK = 0xffffffffu;
constexpr int NUM_SMS_B200 = 148;

void set_error(const char *fmt, ...);
int get_option(const char *name, int dflt);

// knn_morton.cu
size_t knn_morton_workspace_bytes(int B, int M, int N);
struct KmSorted {  // both clouds in Morton order (device pointers into the caller's workspace)
    const float *sp, *sq;                // sorted coordinates (B,N,3) / (B,M,3)
    const int *spi, *sqi;                // original index of every sorted point / query
    const unsigned long long *pk, *qk;   // sorted keys (batch << 32 | Morton code)
    const float4 *boxes;                 // bounding boxes of the 64-point tiles of the points
    unsigned long long *counter;         // a 64-bit scratch counter (statistics)
};
int knn_morton_prepare(const float *query, const float *points, int B, int M, int N, void *workspace,
                       size_t workspace_bytes, cudaStream_t st, KmSorted *out);
extern double g_knn_tiles_visited, g_knn_tiles_total;  // pp_knn_stats
// knn_tc.cu
bool knn_tc_supported(int B, int M, int N, int k);
size_t knn_tc_workspace_bytes(int B, int M, int N);
int knn_tc_launch(const float *query, const float *points, int B, int M, int N, int k, float *dist, int *idx,
                  void *workspace, size_t workspace_bytes, cudaStream_t st);
int knn_morton_launch(const float *query, const float *points, int B, int M, int N, int k, float *dist,
                      int *idx, void *workspace, size_t workspace_bytes, cudaStream_t st);

// chamfer_sweep.cu
cudaError_t chamfer_sweep_tmem_probe(int mode, int iters, float *out, double *bytes);  // pp_microbench 7 / 8
size_t chamfer_sweep_workspace_bytes(int B, int N, int M);
int chamfer_sweep_launch(const float *xyz1, const float *xyz2, int B, int N, int M, float *dist1, float *dist2,
                         int *idx1, int *idx2, float *sums, void *workspace, size_t workspace_bytes,
                         const float *gw, float *g1, float *g2, cudaStream_t st);

// Optional per-kernel timing (option "timing" = 1): CUDA events recorded on the launching
// stream right around one kernel; read back with pp_timing_collect().  Used by bench.py to
// measure the dominant kernel's duration without a profiler.
struct KernelTimer {
    const char *name;
    cudaStream_t st;
    cudaEvent_t e0 = nullptr;
    bool on;
    KernelTimer(const char *name, cudaStream_t st);
    ~KernelTimer();
};

// RAII device switch: every entry point runs on the device the caller names and
// restores the previous one, so a mismatched "current device" can never send a
// launch to the wrong GPU (the reference has no device guard).
struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) {
            err = cudaSetDevice(device);
            changed = (err == cudaSuccess);
        }
    }
    ~DeviceGuard() {
        if (changed) cudaSetDevice(prev);
    }
};

#define PP_REQUIRE(cond, ...)                  \
    do {                                       \
        if (!(cond)) {                         \
            ::pp::set_error(__VA_ARGS__);      \
            return PP_EINVAL;                  \
        }                                      \
    } while (0)

#define PP_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            ::pp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                            __LINE__);                                                     \
            return (int)_e;                                                                \
        }                                                                                  \
    } while (0)

// ---- programmatic dependent launch (PDL) ----------------------------------------------------
// The hot path at the headline shape is four short kernels back to back (75 + 17 + 8 + 8 us).  A
// kernel launched with launch_pdl() may become resident while its predecessor in the stream is
// still draining; it must call pdl_wait() before touching anything the predecessor wrote (the
// wait returns once all prerequisite grids have completed and their memory is visible), and a
// predecessor lets it in early by calling pdl_launch_dependents().  Without either call the
// behaviour is that of an ordinary launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    // with per-kernel timing on, kernels are launched fully serialised: CUDA events around a kernel that was allowed
    // to become resident during its predecessor do not bracket that kernel alone
    cfg.numAttrs = (get_option("pdl", 1) && !get_option("timing", 0)) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define PP_LAUNCH_CHECK()                                                                     \
    do {                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            ::pp::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, \
                            __LINE__);                                                        \
            return (int)_e;                                                                   \
        }                                                                                     \
    } while (0)

__host__ __device__ static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------
// fp32 building blocks with the rounding order pinned by explicit intrinsics /
// PTX (never left to the compiler's contraction), SURVEY.md D4 + Appendix A.
// ---------------------------------------------------------------------------

// Chamfer / KNN order (_ext/nmdistance_cuda.cu:31-35): t = ref - query,
// d = fma(tz,tz, fma(ty,ty, rn(tx*tx))).
__device__ __forceinline__ float sqdist_xyz(float rx, float ry, float rz, float qx, float qy,
                                            float qz) {
    const float tx = __fsub_rn(rx, qx), ty = __fsub_rn(ry, qy), tz = __fsub_rn(rz, qz);
    return __fmaf_rn(tz, tz, __fmaf_rn(ty, ty, __fmul_rn(tx, tx)));
}

// FPS / ball_query / three_nn order (_ext/sampling_cuda.cu:202,364; interpolate_gpu.cu:36):
// the 3-term expression contracts to fma(dz,dz, fma(dx,dx, rn(dy*dy))) -- y first.
__device__ __forceinline__ float sqdist_yxz(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// Packed fp32x2 (Blackwell FADD2/FMUL2/FFMA2): two Chamfer-order distances per
// instruction stream.  (rx,ry,rz) hold two reference points, nq* the NEGATED query
// coordinate: rn(r + (-q)) == rn(r - q) bit for bit.  ptxas folds the {nq,nq}
// pair into the .F32 broadcast operand of FADD2.
__device__ __forceinline__ float2 sqdist2_xyz(float2 rx, float2 ry, float2 rz, float nqx, float nqy,
                                              float nqz) {
    float2 d;
    asm("{\n\t"
        ".reg .b64 bx, by, bz, qx, qy, qz, tx, ty, tz, dd;\n\t"
        "mov.b64 bx, {%2, %3};\n\t"
        "mov.b64 by, {%4, %5};\n\t"
        "mov.b64 bz, {%6, %7};\n\t"
        "mov.b64 qx, {%8, %8};\n\t"
        "mov.b64 qy, {%9, %9};\n\t"
        "mov.b64 qz, {%10, %10};\n\t"
        "add.rn.f32x2 tx, bx, qx;\n\t"
        "add.rn.f32x2 ty, by, qy;\n\t"
        "add.rn.f32x2 tz, bz, qz;\n\t"
        "mul.rn.f32x2 dd, tx, tx;\n\t"
        "fma.rn.f32x2 dd, ty, ty, dd;\n\t"
        "fma.rn.f32x2 dd, tz, tz, dd;\n\t"
        "mov.b64 {%0, %1}, dd;\n\t"
        "}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(rx.x), "f"(rx.y), "f"(ry.x), "f"(ry.y), "f"(rz.x), "f"(rz.y), "f"(nqx), "f"(nqy),
          "f"(nqz));
    return d;
}

// Same but in the FPS / ball_query order (y first): d = fma(tz,tz, fma(tx,tx, rn(ty*ty))).
__device__ __forceinline__ float2 sqdist2_yxz(float2 rx, float2 ry, float2 rz, float nqx, float nqy,
                                              float nqz) {
    float2 d;
    asm("{\n\t"
        ".reg .b64 bx, by, bz, qx, qy, qz, tx, ty, tz, dd;\n\t"
        "mov.b64 bx, {%2, %3};\n\t"
        "mov.b64 by, {%4, %5};\n\t"
        "mov.b64 bz, {%6, %7};\n\t"
        "mov.b64 qx, {%8, %8};\n\t"
        "mov.b64 qy, {%9, %9};\n\t"
        "mov.b64 qz, {%10, %10};\n\t"
        "add.rn.f32x2 tx, bx, qx;\n\t"
        "add.rn.f32x2 ty, by, qy;\n\t"
        "add.rn.f32x2 tz, bz, qz;\n\t"
        "mul.rn.f32x2 dd, ty, ty;\n\t"
        "fma.rn.f32x2 dd, tx, tx, dd;\n\t"
        "fma.rn.f32x2 dd, tz, tz, dd;\n\t"
        "mov.b64 {%0, %1}, dd;\n\t"
        "}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(rx.x), "f"(rx.y), "f"(ry.x), "f"(ry.y), "f"(rz.x), "f"(rz.y), "f"(nqx), "f"(nqy),
          "f"(nqz));
    return d;
}

// Three-input min/max (FMNMX3 on sm_100).
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

constexpr float PP_INF = __builtin_huge_valf();

}  // namespace pp
