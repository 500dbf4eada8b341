/*
 * pp_oracle.c -- CPU restatement of the pytorch_points hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for the CUDA kernels in
 * pytorch_points_b200/csrc; it is never the thing shipped or measured as the
 * product.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.
 *
 * Every function restates, operation by operation, what the reference's CUDA
 * kernels compute once nvcc (default --fmad=true, -O2) has contracted them;
 * the fp32 rounding sequence is spelled out with fmaf() and the file must be
 * compiled with -ffp-contract=off so the host compiler adds no contraction of
 * its own.  Citations are into /root/reference/pytorch_points/.
 *
 * Parity status: the reference ships no golden vectors.  This
 * restatement is pinned against outputs of the reference's own kernels built
 * for sm_100a (oracle/build_ref.sh -> oracle/_ref/*.so) and run on a B200;
 * the captured vectors live in tests/golden/ (see tests/golden/make_golden.py).
 * KNN has no reference implementation in the snapshot: its
 * contract is defined by this repo and is "parity unpinned".
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PP_CHUNK 512 /* _ext/nmdistance_cuda.cu:5  const int BATCH = 512 */

/* ------------------------------------------------------------------------- */
/* Chamfer / nndistance forward, one direction.                               */
/* _ext/nmdistance_cuda.cu:8-49 (NmDistanceKernel)                            */
/*   tmp = ref[k] - query[j]   (:33)                                          */
/*   d   = fma(tmp,tmp,d), d starting at 0, components in order (:31-35)      */
/*   inside a 512-chunk: k==0 || d<best (:36); across chunks: k2==0 ||        */
/*   result>best (:41)  => lowest index wins ties.                            */
/* ------------------------------------------------------------------------- */
static void nm_one_direction(int n, int c, const float *q, int m, const float *r,
                             float *result, int32_t *result_i)
{
#pragma omp parallel
    {
        float dbuf[PP_CHUNK];
#pragma omp for schedule(static)
        for (int j = 0; j < n; j++) {
            const float *qj = q + (size_t)j * c;
            float res = 0.f;
            int32_t res_i = 0;
            for (int k2 = 0; k2 < m; k2 += PP_CHUNK) {
                int end_k = (m < k2 + PP_CHUNK ? m : k2 + PP_CHUNK) - k2;
                const float *rc = r + (size_t)k2 * c;
                if (c == 3) {
                    const float qx = qj[0], qy = qj[1], qz = qj[2];
                    for (int k = 0; k < end_k; k++) {
                        float tx = rc[k * 3 + 0] - qx;
                        float ty = rc[k * 3 + 1] - qy;
                        float tz = rc[k * 3 + 2] - qz;
                        dbuf[k] = fmaf(tz, tz, fmaf(ty, ty, fmaf(tx, tx, 0.f)));
                    }
                } else {
                    for (int k = 0; k < end_k; k++) {
                        float d = 0.f;
                        for (int cc = 0; cc < c; cc++) {
                            float tmp = rc[k * c + cc] - qj[cc];
                            d = fmaf(tmp, tmp, d);
                        }
                        dbuf[k] = d;
                    }
                }
                int best_i = 0;
                float best = 0.f;
                for (int k = 0; k < end_k; k++) {
                    float d = dbuf[k];
                    if (k == 0 || d < best) { best = d; best_i = k + k2; }
                }
                if (k2 == 0 || res > best) { res = best; res_i = best_i; }
            }
            result[j] = res;
            result_i[j] = res_i;
        }
    }
}

/* _ext/nmdistance_cuda.cu:118-140 (chamfer_cuda_forward): two launches, roles swapped */
int oracle_chamfer_fwd(const float *xyz1, const float *xyz2, int B, int N, int M, int c,
                       float *dist1, float *dist2, int32_t *idx1, int32_t *idx2)
{
    for (int b = 0; b < B; b++) {
        const float *a = xyz1 + (size_t)b * N * c, *bb = xyz2 + (size_t)b * M * c;
        nm_one_direction(N, c, a, M, bb, dist1 + (size_t)b * N, idx1 + (size_t)b * N);
        nm_one_direction(M, c, bb, N, a, dist2 + (size_t)b * M, idx2 + (size_t)b * M);
    }
    return 1;
}

/* ------------------------------------------------------------------------- */
/* Labeled Chamfer, one direction. _ext/nmdistance_cuda.cu:56-115             */
/* per chunk best=1e10f,best_i=-1 (:78-80); candidate only if labels equal    */
/* (:89) with the k==0||d<best quirk inside (:96); chunk 0 always writes      */
/* (:101); finally idx<0 => dist=0 (:110-113).                                */
/* ------------------------------------------------------------------------- */
static void labeled_one_direction(int n, int c, const float *q, const float *ql, int m,
                                  const float *r, const float *rl, float *result,
                                  int32_t *result_i)
{
#pragma omp parallel for schedule(static)
    for (int j = 0; j < n; j++) {
        const float *qj = q + (size_t)j * c;
        const float l1 = ql[j];
        float res = 0.f;
        int32_t res_i = 0;
        for (int k2 = 0; k2 < m; k2 += PP_CHUNK) {
            int end_k = (m < k2 + PP_CHUNK ? m : k2 + PP_CHUNK) - k2;
            int best_i = -1;
            float best = 1e10f;
            for (int k = 0; k < end_k; k++) {
                if (l1 == rl[k2 + k]) {
                    float d = 0.f;
                    for (int cc = 0; cc < c; cc++) {
                        float tmp = r[(size_t)(k2 + k) * c + cc] - qj[cc];
                        d = fmaf(tmp, tmp, d);
                    }
                    if (k == 0 || d < best) { best = d; best_i = k + k2; }
                }
            }
            if (k2 == 0 || res > best) { res = best; res_i = best_i; }
        }
        if (res_i < 0) res = 0.f;
        result[j] = res;
        result_i[j] = res_i;
    }
}

/* _ext/nmdistance_cuda.cu:142-166 */
int oracle_chamfer_labeled_fwd(const float *xyz1, const float *xyz2, const float *label1,
                               const float *label2, int B, int N, int M, int c, float *dist1,
                               float *dist2, int32_t *idx1, int32_t *idx2)
{
    for (int b = 0; b < B; b++) {
        const float *a = xyz1 + (size_t)b * N * c, *bb = xyz2 + (size_t)b * M * c;
        const float *la = label1 + (size_t)b * N, *lb = label2 + (size_t)b * M;
        labeled_one_direction(N, c, a, la, M, bb, lb, dist1 + (size_t)b * N, idx1 + (size_t)b * N);
        labeled_one_direction(M, c, bb, lb, N, a, la, dist2 + (size_t)b * M, idx2 + (size_t)b * M);
    }
    return 1;
}

/* ------------------------------------------------------------------------- */
/* Chamfer backward. _ext/nmdistance_cuda.cu:169-221                          */
/*   g = gd*2; v = g*(a[j]-b[idx[j]]); grad_a[j] += v; grad_b[idx[j]] += -v   */
/*   (:176-181), skipped when idx<0 (:175); both sides accumulate into the    */
/*   same zeroed arrays (:204-210).  The GPU adds atomically in unspecified   */
/*   order; this restatement adds side 1 then side 2 in index order, so       */
/*   comparisons against it are tolerance based.              */
/* ------------------------------------------------------------------------- */
static void grad_one_side(int n, int c, const float *a, int m, const float *b, const float *gd,
                          const int32_t *idx, float *ga, float *gb)
{
    (void)m;
    for (int j = 0; j < n; j++) {
        int j2 = idx[j];
        if (j2 < 0) continue;
        float g = gd[j] * 2.f;
        for (int cc = 0; cc < c; cc++) {
            float diff = a[(size_t)j * c + cc] - b[(size_t)j2 * c + cc];
            float v = g * diff;
            ga[(size_t)j * c + cc] += v;
            gb[(size_t)j2 * c + cc] += -v;
        }
    }
}

int oracle_chamfer_bwd(const float *xyz1, const float *xyz2, const float *gd1, const float *gd2,
                       const int32_t *idx1, const int32_t *idx2, int B, int N, int M, int c,
                       float *g1, float *g2)
{
    memset(g1, 0, sizeof(float) * (size_t)B * N * c);
    memset(g2, 0, sizeof(float) * (size_t)B * M * c);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; b++) {
        const float *a = xyz1 + (size_t)b * N * c, *bb = xyz2 + (size_t)b * M * c;
        float *ga = g1 + (size_t)b * N * c, *gb = g2 + (size_t)b * M * c;
        grad_one_side(N, c, a, M, bb, gd1 + (size_t)b * N, idx1 + (size_t)b * N, ga, gb);
        grad_one_side(M, c, bb, N, a, gd2 + (size_t)b * M, idx2 + (size_t)b * M, gb, ga);
    }
    return 1;
}

/* ------------------------------------------------------------------------- */
/* Farthest point sampling. _ext/sampling_cuda.cu:162-233                     */
/*   block size bs = max(min(2^floor(log2 n),512),1)  (_ext/cuda_utils.h:11-16)*/
/*   d = (x2-x1)^2+(y2-y1)^2+(z2-z1)^2 contracted by nvcc to                  */
/*       fma(dz,dz, fma(dx,dx, rn(dy*dy)))   (:202; SASS FMUL(dy) FFMA FFMA)  */
/*   d2 = min(d,temp[k]); temp[k]=d2 (:203-205)                               */
/*   per thread t: candidates k=t,t+bs,.. strict > from best=-1 (:189,206-209)*/
/*   tree over thread slots keeps the lower slot on ties (:214-226)           */
/*   => argmax with tie key (k mod bs, k).                                    */
/* ------------------------------------------------------------------------- */
static int fps_block_size(int n)
{
    /* std::log(double(n))/std::log(2.0) truncated to int, as cuda_utils.h:14 */
    int pow_2 = (int)(log((double)n) / log(2.0));
    int v = 1 << pow_2;
    if (v > 512) v = 512;
    if (v < 1) v = 1;
    return v;
}

int oracle_fps_block_size(int n) { return fps_block_size(n); }

int oracle_fps(const float *xyz, int B, int N, int m, int seed, float *temp, int32_t *idx)
{
    if (m <= 0) return 1; /* :166 */
    const int bs = fps_block_size(N);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; b++) {
        const float *p = xyz + (size_t)b * N * 3;
        float *tp = temp + (size_t)b * N;
        int32_t *out = idx + (size_t)b * m;
        float *tbest = (float *)malloc(sizeof(float) * bs);
        int *tbesti = (int *)malloc(sizeof(int) * bs);
        int old = seed;
        out[0] = old;
        for (int j = 1; j < m; j++) {
            const float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
            for (int t = 0; t < bs; t++) { tbest[t] = -1.f; tbesti[t] = 0; }
            for (int k = 0; k < N; k++) {
                float dx = p[k * 3 + 0] - x1, dy = p[k * 3 + 1] - y1, dz = p[k * 3 + 2] - z1;
                float d = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
                float td = tp[k];
                float d2 = fminf(d, td);
                if (d2 != td) tp[k] = d2;
                int t = k % bs; /* ascending k per slot == the strided per-thread loop */
                if (d2 > tbest[t]) { tbest[t] = d2; tbesti[t] = k; }
            }
            /* tree: slot i1 replaced by i2 only if strictly smaller => lowest slot wins ties */
            for (int u = 0; (1 << u) < bs; u++) {
                for (int t = 0; t < (bs >> (u + 1)); t++) {
                    int i1 = (t * 2) << u, i2 = (t * 2 + 1) << u;
                    if (tbest[i1] < tbest[i2]) { tbest[i1] = tbest[i2]; tbesti[i1] = tbesti[i2]; }
                }
            }
            old = tbesti[0];
            out[j] = old;
        }
        free(tbest);
        free(tbesti);
    }
    return 1;
}

/* ------------------------------------------------------------------------- */
/* ball_query. _ext/sampling_cuda.cu:340-376; output zero-filled by the       */
/* wrapper (_ext/sampling.cpp:93-94).                                         */
/*   r2 = rn(radius*radius) (:354)                                            */
/*   d2 = (nx-x)^2+(ny-y)^2+(nz-z)^2 -> fma(dz,dz, fma(dx,dx, rn(dy*dy)))     */
/*   first hit fills all nsample slots (:366-370); strict < (:365)            */
/* ------------------------------------------------------------------------- */
int oracle_ball_query(const float *new_xyz, const float *xyz, int B, int N, int M, float radius,
                      int nsample, int32_t *idx)
{
    const float r2 = radius * radius;
    memset(idx, 0, sizeof(int32_t) * (size_t)B * M * nsample);
#pragma omp parallel for schedule(static) collapse(2)
    for (int b = 0; b < B; b++) {
        for (int j = 0; j < M; j++) {
            const float *q = new_xyz + ((size_t)b * M + j) * 3;
            const float *p = xyz + (size_t)b * N * 3;
            int32_t *o = idx + ((size_t)b * M + j) * nsample;
            const float nx = q[0], ny = q[1], nz = q[2];
            int cnt = 0;
            for (int k = 0; k < N; k++) {
                float dx = nx - p[k * 3 + 0], dy = ny - p[k * 3 + 1], dz = nz - p[k * 3 + 2];
                float d2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
                if (d2 < r2) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; l++) o[l] = k;
                    o[cnt] = k;
                    ++cnt;
                    if (cnt >= nsample) break;
                }
            }
        }
    }
    return 1;
}

/* gather_points fwd/bwd. _ext/sampling_cuda.cu:9-25,47-64 */
int oracle_gather_fwd(const float *points, const int32_t *idx, int B, int C, int N, int npoint,
                      float *out)
{
    for (int b = 0; b < B; b++)
        for (int c = 0; c < C; c++)
            for (int j = 0; j < npoint; j++)
                out[((size_t)b * C + c) * npoint + j] =
                    points[((size_t)b * C + c) * N + idx[(size_t)b * npoint + j]];
    return 1;
}

int oracle_gather_bwd(const float *grad_out, const int32_t *idx, int B, int C, int N, int npoint,
                      float *grad_points)
{
    memset(grad_points, 0, sizeof(float) * (size_t)B * C * N);
    for (int b = 0; b < B; b++)
        for (int c = 0; c < C; c++)
            for (int j = 0; j < npoint; j++)
                grad_points[((size_t)b * C + c) * N + idx[(size_t)b * npoint + j]] +=
                    grad_out[((size_t)b * C + c) * npoint + j];
    return 1;
}

/* group_points fwd/bwd ("next" row N1). _ext/sampling_cuda.cu:447-467,482-503 */
int oracle_group_fwd(const float *points, const int32_t *idx, int B, int C, int N, int npoint,
                     int nsample, float *out)
{
    for (int b = 0; b < B; b++)
        for (int c = 0; c < C; c++)
            for (int j = 0; j < npoint; j++)
                for (int s = 0; s < nsample; s++)
                    out[(((size_t)b * C + c) * npoint + j) * nsample + s] =
                        points[((size_t)b * C + c) * N +
                               idx[((size_t)b * npoint + j) * nsample + s]];
    return 1;
}

int oracle_group_bwd(const float *grad_out, const int32_t *idx, int B, int C, int N, int npoint,
                     int nsample, float *grad_points)
{
    memset(grad_points, 0, sizeof(float) * (size_t)B * C * N);
    for (int b = 0; b < B; b++)
        for (int c = 0; c < C; c++)
            for (int j = 0; j < npoint; j++)
                for (int s = 0; s < nsample; s++)
                    grad_points[((size_t)b * C + c) * N +
                                idx[((size_t)b * npoint + j) * nsample + s]] +=
                        grad_out[(((size_t)b * C + c) * npoint + j) * nsample + s];
    return 1;
}

/* ------------------------------------------------------------------------- */
/* group_knn.  The reference snapshot has NO implementation (SURVEY.md D1;    */
/* README.md:12 names it, callers use pytorch3d.ops.knn_points e.g.           */
/* network/layers.py:52).  Contract defined by this repo: squared distance    */
/* in the Chamfer op order (ref - query, fma chain from 0), neighbours sorted */
/* ascending by (distance, index).  "parity unpinned".                        */
/* ------------------------------------------------------------------------- */
int oracle_knn(const float *query, const float *points, int B, int M, int N, int c, int k,
               float *dist, int32_t *idx)
{
    if (k > N) return 0;
#pragma omp parallel for schedule(static) collapse(2)
    for (int b = 0; b < B; b++) {
        for (int j = 0; j < M; j++) {
            const float *q = query + ((size_t)b * M + j) * c;
            const float *p = points + (size_t)b * N * c;
            float *od = dist + ((size_t)b * M + j) * k;
            int32_t *oi = idx + ((size_t)b * M + j) * k;
            int cnt = 0;
            for (int i = 0; i < N; i++) {
                float d = 0.f;
                for (int cc = 0; cc < c; cc++) {
                    float tmp = p[(size_t)i * c + cc] - q[cc];
                    d = fmaf(tmp, tmp, d);
                }
                /* ascending i: strict < keeps the lower index on equal distance */
                if (cnt < k || d < od[k - 1]) {
                    int pos = cnt < k ? cnt : k - 1;
                    while (pos > 0 && d < od[pos - 1]) {
                        od[pos] = od[pos - 1];
                        oi[pos] = oi[pos - 1];
                        pos--;
                    }
                    od[pos] = d;
                    oi[pos] = i;
                    if (cnt < k) cnt++;
                }
            }
        }
    }
    return 1;
}

/* ------------------------------------------------------------------------- */
/* three_nn ("next" row N3). _ext/interpolate_gpu.cu:9-52: d as ball_query    */
/* with (unknown - known); bests kept as double 1e40, strict < cascade.       */
/* ------------------------------------------------------------------------- */
int oracle_three_nn(const float *unknown, const float *known, int B, int N, int M, float *dist2,
                    int32_t *idx)
{
#pragma omp parallel for schedule(static) collapse(2)
    for (int b = 0; b < B; b++) {
        for (int j = 0; j < N; j++) {
            const float *u = unknown + ((size_t)b * N + j) * 3;
            const float *kn = known + (size_t)b * M * 3;
            double best1 = 1e40, best2 = 1e40, best3 = 1e40;
            int b1 = 0, b2 = 0, b3 = 0;
            for (int k = 0; k < M; k++) {
                float dx = u[0] - kn[k * 3 + 0], dy = u[1] - kn[k * 3 + 1], dz = u[2] - kn[k * 3 + 2];
                float d = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
                if (d < best1) {
                    best3 = best2; b3 = b2; best2 = best1; b2 = b1; best1 = d; b1 = k;
                } else if (d < best2) {
                    best3 = best2; b3 = b2; best2 = d; b2 = k;
                } else if (d < best3) {
                    best3 = d; b3 = k;
                }
            }
            float *od = dist2 + ((size_t)b * N + j) * 3;
            int32_t *oi = idx + ((size_t)b * N + j) * 3;
            od[0] = (float)best1; od[1] = (float)best2; od[2] = (float)best3;
            oi[0] = b1; oi[1] = b2; oi[2] = b3;
        }
    }
    return 1;
}

/* three_interpolate fwd/bwd ("next" row N3). _ext/interpolate_gpu.cu:77-97,120-142.
 * w0*p0 + w1*p1 + w2*p2 is contracted by nvcc to fma(w2,p2, fma(w0,p0, rn(w1*p1))) -- middle
 * product first, like the 3-term distances; pinned by tests/golden/three_interpolate.npz. */
int oracle_three_interpolate_fwd(const float *points, const int32_t *idx, const float *weight, int B, int C,
                                 int M, int N, float *out)
{
    for (int b = 0; b < B; b++)
        for (int c = 0; c < C; c++)
            for (int i = 0; i < N; i++) {
                const int32_t *id = idx + ((size_t)b * N + i) * 3;
                const float *w = weight + ((size_t)b * N + i) * 3;
                const float *p = points + ((size_t)b * C + c) * M;
                out[((size_t)b * C + c) * N + i] = fmaf(w[2], p[id[2]], fmaf(w[0], p[id[0]], w[1] * p[id[1]]));
            }
    return 1;
}

int oracle_three_interpolate_bwd(const float *grad_out, const int32_t *idx, const float *weight, int B, int C,
                                 int N, int M, float *grad_points)
{
    memset(grad_points, 0, sizeof(float) * (size_t)B * C * M);
    for (int b = 0; b < B; b++)
        for (int c = 0; c < C; c++)
            for (int i = 0; i < N; i++) {
                const int32_t *id = idx + ((size_t)b * N + i) * 3;
                const float *w = weight + ((size_t)b * N + i) * 3;
                float *g = grad_points + ((size_t)b * C + c) * M;
                const float go = grad_out[((size_t)b * C + c) * N + i];
                g[id[0]] += go * w[0];
                g[id[1]] += go * w[1];
                g[id[2]] += go * w[2];
            }
    return 1;
}

int oracle_version(void) { return 1; }

/* Thread count of the OpenMP loops above (bench.py's CPU arms: torch.distributed.run exports
 * OMP_NUM_THREADS=1 and libgomp has usually read it before this library is loaded). */
#ifdef _OPENMP
#include <omp.h>
void oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int oracle_get_max_threads(void) { return omp_get_max_threads(); }
#else
void oracle_set_threads(int n) { (void)n; }
int oracle_get_max_threads(void) { return 1; }
#endif
