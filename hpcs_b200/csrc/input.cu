// Device side of the per-step input pipeline.
//
// The reference rotates every batch on the HOST (points.cpu() -> pytorch3d Rotate / RotateAxisAngle ->
// transform_points), uploads it, then hands the backbone a transposed view, and builds the one-hot category vector
// with torch.eye on the host followed by .cuda() (hpcs/models/shapenet_hyp_hc.py:63-73,84, partnet_hyp_hc.py:82-105,
// hpcs/utils/data.py:24-29).  Here the host only draws the rotation PARAMETERS from torch's CPU generator, in the
// reference's order (randn(B,4) for 'so3', rand(B) for 'z': 16 or 4 bytes per cloud, RNG parity kept), and one kernel
// turns them into matrices, rotates, and writes the backbone's [B,3,N] layout directly.
//
// Arithmetic restated from pytorch3d 0.7.2 (hpcs-env.yaml:287; not vendored in the reference):
//   random_rotations: q = o / copysign(|o|, o_0) for o ~ N(0,1)^4, then R from the unit quaternion (r,i,j,k) with
//                     two_s = 2 / |q|^2:  [[1-ts(jj+kk), ts(ij-kr), ts(ik+jr)], [ts(ij+kr), 1-ts(ii+kk), ts(jk-ir)],
//                                          [ts(ik-jr), ts(jk+ir), 1-ts(ii+jj)]]
//   RotateAxisAngle(angle, 'Z', degrees): a = angle/180*pi, column-vector matrix [[c,-s,0],[s,c,0],[0,0,1]] TRANSPOSED
//   transform_points: row vectors, out = p @ R.
#include "common.cuh"

namespace hpcs {

__device__ __forceinline__ void rotation_of(const float* __restrict__ params, int mode, int b, float R[9]) {
    if (mode == 1) {                                   // R[B,3,3] given
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = params[(size_t)b * 9 + i];
    } else if (mode == 2) {                            // four N(0,1) draws per cloud -> unit quaternion -> matrix
        const float o0 = params[b * 4 + 0], o1 = params[b * 4 + 1], o2 = params[b * 4 + 2], o3 = params[b * 4 + 3];
        float nrm = sqrtf(o0 * o0 + o1 * o1 + o2 * o2 + o3 * o3);
        if ((nrm < 0.f) != (o0 < 0.f)) nrm = -nrm;
        const float r = o0 / nrm, i = o1 / nrm, j = o2 / nrm, k = o3 / nrm;
        const float ts = 2.0f / (r * r + i * i + j * j + k * k);
        R[0] = 1.f - ts * (j * j + k * k); R[1] = ts * (i * j - k * r);       R[2] = ts * (i * k + j * r);
        R[3] = ts * (i * j + k * r);       R[4] = 1.f - ts * (i * i + k * k); R[5] = ts * (j * k - i * r);
        R[6] = ts * (i * k - j * r);       R[7] = ts * (j * k + i * r);       R[8] = 1.f - ts * (i * i + j * j);
    } else if (mode == 3) {                            // one U(0,1) draw per cloud -> rotation about z by 360 u degrees
        const float a = params[b] * 360.f / 180.0f * 3.14159265358979323846f;
        float s, c;
        sincosf(a, &s, &c);
        R[0] = c;   R[1] = s;   R[2] = 0.f;            // transpose of [[c,-s,0],[s,c,0],[0,0,1]]
        R[3] = -s;  R[4] = c;   R[5] = 0.f;
        R[6] = 0.f; R[7] = 0.f; R[8] = 1.f;
    } else {
        R[0] = R[4] = R[8] = 1.f;
        R[1] = R[2] = R[3] = R[5] = R[6] = R[7] = 0.f;
    }
}

// pts[B,N,3] -> out[B,3,N] = (pts @ R_b)^T.  One thread per point: 12-byte reads (the three floats of a point are
// consecutive, a warp reads 384 contiguous bytes), three coalesced 4-byte writes.
__global__ void __launch_bounds__(256)
rotate_points_kernel(const float* __restrict__ pts, const float* __restrict__ params, int mode, int B, int N,
                     float* __restrict__ out, float* __restrict__ rot_out) {
    const int b = blockIdx.y;
    float R[9];
    rotation_of(params, mode, b, R);
    if (rot_out && blockIdx.x == 0 && threadIdx.x < 9) rot_out[(size_t)b * 9 + threadIdx.x] = R[threadIdx.x];
    const float* src = pts + (size_t)b * N * 3;
    float* dst = out + (size_t)b * 3 * N;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const float x = src[3 * n], y = src[3 * n + 1], z = src[3 * n + 2];
        if (mode == 0) {
            dst[n] = x; dst[N + n] = y; dst[2 * N + n] = z;
        } else {
            dst[n]         = fmaf(z, R[6], fmaf(y, R[3], x * R[0]));
            dst[N + n]     = fmaf(z, R[7], fmaf(y, R[4], x * R[1]));
            dst[2 * N + n] = fmaf(z, R[8], fmaf(y, R[5], x * R[2]));
        }
    }
}

__global__ void __launch_bounds__(256)
one_hot_kernel(const int64_t* __restrict__ y, int64_t rows, int classes, float* __restrict__ out) {
    const int64_t total = rows * classes;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / classes;
        out[i] = (y[r] == (i - r * classes)) ? 1.f : 0.f;
    }
}

// logits[i][c] = scale * (e_i . w_c / (max(|e_i|, eps) max(|w_c|, eps)) - margin * [labels[i] == c]): the metrics-side CosFace logits.
// A block stages the normalised class directions once (shared memory, odd row stride) and walks 64-row groups of embeddings;
// thread -> (row, class) pairs with the class index fastest, so the logits leave in coalesced rows.
constexpr int kLgRows = 64;
__global__ void __launch_bounds__(256)
cosface_logits_kernel(const float* __restrict__ emb, const float* __restrict__ W, const int64_t* __restrict__ labels, int64_t n, int D,
                      int classes, float margin, float scale, float* __restrict__ logits) {
    extern __shared__ float lg_sm[];
    const int S = D + 1;
    float* wn = lg_sm;                                   // [classes][S] normalised columns of W
    float* es = wn + (size_t)classes * S;                // [kLgRows][S] normalised embedding rows
    for (int c = threadIdx.x; c < classes; c += blockDim.x) {
        float s2 = 0.f;
        for (int d = 0; d < D; ++d) { const float w = W[(size_t)d * classes + c]; s2 = fmaf(w, w, s2); }
        const float inv = 1.f / fmaxf(sqrtf(s2), 1e-12f);           // F.normalize(p=2, eps=1e-12)
        for (int d = 0; d < D; ++d) wn[c * S + d] = W[(size_t)d * classes + c] * inv;
    }
    const int64_t groups = (n + kLgRows - 1) / kLgRows;
    for (int64_t g = blockIdx.x; g < groups; g += gridDim.x) {
        const int64_t r0 = g * kLgRows;
        const int nr = (int)((n - r0) < kLgRows ? (n - r0) : kLgRows);
        __syncthreads();
        for (int i = threadIdx.x; i < nr * D; i += blockDim.x) es[(i / D) * S + i % D] = emb[r0 * D + i];
        __syncthreads();
        for (int r = threadIdx.x >> 5; r < nr; r += blockDim.x >> 5) {      // one warp per row: its norm
            float s2 = 0.f;
            for (int d = threadIdx.x & 31; d < D; d += 32) s2 = fmaf(es[r * S + d], es[r * S + d], s2);
            s2 = warp_sum(s2);
            const float inv = 1.f / fmaxf(sqrtf(s2), 1e-12f);
            for (int d = threadIdx.x & 31; d < D; d += 32) es[r * S + d] *= inv;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < nr * classes; t += blockDim.x) {
            const int r = t / classes, c = t - r * classes;
            const float* e = es + r * S;
            const float* w = wn + c * S;
            float acc = 0.f;
            for (int d = 0; d < D; ++d) acc = fmaf(e[d], w[d], acc);
            const float m = labels[r0 + r] == c ? margin : 0.f;
            logits[(r0 + r) * classes + c] = (acc - m) * scale;
        }
    }
}

}  // namespace hpcs

extern "C" int hpcs_cosface_logits_f32(const float* emb, const float* W, const int64_t* labels, int64_t n, int D, int classes, float margin,
                                       float scale, float* logits, void* stream) {
    using namespace hpcs;
    if (n == 0) return HPCS_OK;
    if (!emb || !W || !labels || !logits || n < 0 || D <= 0 || classes <= 0) return fail(HPCS_ERR_ARG, "cosface_logits: bad arguments");
    const size_t smem = ((size_t)classes + kLgRows) * (D + 1) * sizeof(float);
    if (smem > 200 * 1024) return fail(HPCS_ERR_ARG, "cosface_logits: classes=%d x D=%d does not fit shared memory", classes, D);
    static thread_local size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        cudaFuncSetAttribute(cosface_logits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = smem;
    }
    int64_t groups = (n + kLgRows - 1) / kLgRows;
    const int64_t cap = 4 * (int64_t)sm_count();
    cosface_logits_kernel<<<(int)(groups < cap ? groups : cap), 256, smem, as_stream(stream)>>>(emb, W, labels, n, D, classes, margin, scale,
                                                                                             logits);
    return check_launch("cosface_logits_kernel");
}

extern "C" int hpcs_rotate_points_f32(const float* pts, const float* params, int mode, int B, int N, float* out,
                                      float* rot_out, void* stream) {
    using namespace hpcs;
    if (B == 0 || N == 0) return HPCS_OK;
    if (!pts || !out) return fail(HPCS_ERR_ARG, "rotate_points: null pointer");
    if (mode < 0 || mode > 3 || (mode != 0 && !params)) return fail(HPCS_ERR_ARG, "rotate_points: mode %d needs params", mode);
    if (B < 0 || B > 65535 || N < 0) return fail(HPCS_ERR_ARG, "rotate_points: bad shape B=%d N=%d", B, N);
    int gx = (N + 255) / 256;
    if (gx > 64) gx = 64;
    rotate_points_kernel<<<dim3(gx, B), 256, 0, as_stream(stream)>>>(pts, params, mode, B, N, out, rot_out);
    return check_launch("rotate_points_kernel");
}

extern "C" int hpcs_one_hot_f32(const int64_t* y, int64_t rows, int num_classes, float* out, void* stream) {
    using namespace hpcs;
    if (rows == 0) return HPCS_OK;
    if (!y || !out || rows < 0 || num_classes <= 0) return fail(HPCS_ERR_ARG, "one_hot: bad arguments");
    int64_t blocks = (rows * num_classes + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    one_hot_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(y, rows, num_classes, out);
    return check_launch("one_hot_kernel");
}
