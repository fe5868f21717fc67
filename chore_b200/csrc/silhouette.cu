// Silhouette rasteriser of the 'sil' fitting phase: forward + the neural-mesh-renderer gradient, for sm_100a.
//
// The reference renders the object template with the vendored neural_renderer (JiangWenPL fork of NMR v1.1.3):
// SilLossROI.forward (recon/obj_pose_roi.py:159-172) -> Renderer.render_silhouettes -> rasterize_silhouettes ->
// forward_face_index_map_cuda_kernel_1/2 and backward_pixel_map_cuda_kernel
// (external/neural_renderer/neural_renderer/cuda/rasterize_cuda_kernel.cu:25-216,291-550) -- scalar SIMT code without an
// sm_100 build.  This file restates the two operations the silhouette term needs (alpha only, no textures / depth output):
//
//   sil_setup_kernel     per face: back-face test, pixel-space inverse (barycentric) matrix, pixel bounding box
//   sil_forward_kernel   one CTA per 16 x 16 pixel tile: faces whose bounding box touches the tile are compacted into shared
//                        memory in rounds of 256 (no cap: the reference drops faces beyond 512 per 4 x 4 block), every pixel
//                        z-tests them (inside test in NDC, clamped + renormalised barycentrics, 1/z interpolation, near/far)
//   sil_backward_kernel  NMR's hand-designed gradient: for every edge of every front face and both scan axes, walk the pixels
//                        the edge crosses and turn "what would this pixel's alpha become if the edge moved across it" into a
//                        gradient on the two edge vertices.  One WARP per face (the reference: one thread), lanes over the scan
//                        position, fixed-order shuffle reduction at the end.
// Arithmetic follows the reference expression by expression (fp32, with its double literals) so that the face index maps agree
// except on razor-edge pixels where FMA contraction differs.
#include "common.cuh"

namespace {

constexpr int kTile = 16;
constexpr int kRound = 256;

struct FaceAux {          // per (b, face), written by the set-up kernel
    float inv[9];
    int xmin, xmax, ymin, ymax;   // pixel bounding box (empty when xmin > xmax); back faces get an empty box
};

__device__ __forceinline__ bool backside(const float *f) {
    return (f[7] - f[1]) * (f[3] - f[0]) < (f[4] - f[1]) * (f[6] - f[0]);
}

__global__ void __launch_bounds__(256) sil_setup_kernel(const float *__restrict__ faces, int n, int is, FaceAux *__restrict__ aux) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float *face = faces + (size_t)i * 9;
    FaceAux a;
    a.xmin = 1; a.xmax = 0; a.ymin = 1; a.ymax = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) a.inv[k] = 0.f;
    if (!backside(face)) {
        float p[3][2];
#pragma unroll
        for (int num = 0; num < 3; ++num)
#pragma unroll
            for (int dim = 0; dim < 2; ++dim) p[num][dim] = (float)(0.5 * (double)(face[3 * num + dim] * is + is - 1));
        float inv[9] = {p[1][1] - p[2][1], p[2][0] - p[1][0], p[1][0] * p[2][1] - p[2][0] * p[1][1],
                        p[2][1] - p[0][1], p[0][0] - p[2][0], p[2][0] * p[0][1] - p[0][0] * p[2][1],
                        p[0][1] - p[1][1], p[1][0] - p[0][0], p[0][0] * p[1][1] - p[1][0] * p[0][1]};
        const float den = p[2][0] * (p[0][1] - p[1][1]) + p[0][0] * (p[1][1] - p[2][1]) + p[1][0] * (p[2][1] - p[0][1]);
#pragma unroll
        for (int k = 0; k < 9; ++k) a.inv[k] = inv[k] / den;
        a.xmin = (int)fmax(ceil((double)fminf(fminf(p[0][0], p[1][0]), p[2][0])), 0.);
        a.xmax = (int)fmin((double)fmaxf(fmaxf(p[0][0], p[1][0]), p[2][0]), is - 1.);
        a.ymin = (int)fmax(ceil((double)fminf(fminf(p[0][1], p[1][1]), p[2][1])), 0.);
        a.ymax = (int)fmin((double)fmaxf(fmaxf(p[0][1], p[1][1]), p[2][1]), is - 1.);
    }
    aux[i] = a;
}

__global__ void __launch_bounds__(kTile * kTile) sil_forward_kernel(const float *__restrict__ faces, const FaceAux *__restrict__ aux,
                                                                   int F, int is, float near, float far, float *__restrict__ alpha,
                                                                   int32_t *__restrict__ face_index) {
    __shared__ float s_face[kRound][9];
    __shared__ float s_inv[kRound][9];
    __shared__ int s_id[kRound];
    __shared__ int s_warp[8];
    __shared__ int s_n;
    const int b = blockIdx.z, tid = threadIdx.y * kTile + threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int xi = blockIdx.x * kTile + threadIdx.x, yi = blockIdx.y * kTile + threadIdx.y;
    const int tx0 = blockIdx.x * kTile, ty0 = blockIdx.y * kTile, tx1 = tx0 + kTile - 1, ty1 = ty0 + kTile - 1;
    const float yp = (float)((2. * yi + 1 - is) / is), xp = (float)((2. * xi + 1 - is) / is);
    float depth_min = far;
    int best = -1;
    const float *fb = faces + (size_t)b * F * 9;
    const FaceAux *ab = aux + (size_t)b * F;
    for (int f0 = 0; f0 < F; f0 += kRound) {
        // ---- compact the faces of this round whose bounding box touches the tile (index order kept) ----
        const int f = f0 + tid;
        bool hit = false;
        if (f < F) {
            const FaceAux &a = ab[f];
            hit = a.xmin <= a.xmax && a.ymin <= a.ymax && a.xmin <= tx1 && a.xmax >= tx0 && a.ymin <= ty1 && a.ymax >= ty0;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            for (int w = 0; w < 8; ++w) { const int c = s_warp[w]; s_warp[w] = acc; acc += c; }
            s_n = acc;
        }
        __syncthreads();
        if (hit) {
            const int pos = s_warp[warp] + __popc(bal & ((1u << lane) - 1u));
            s_id[pos] = f;
#pragma unroll
            for (int k = 0; k < 9; ++k) { s_face[pos][k] = fb[(size_t)f * 9 + k]; s_inv[pos][k] = ab[f].inv[k]; }
        }
        __syncthreads();
        const int n = s_n;
        if (xi < is && yi < is) {
            for (int j = 0; j < n; ++j) {
                const float *face = s_face[j], *inv = s_inv[j];
                // inside test in normalised coordinates (three edge functions)
                if (((yp - face[1]) * (face[3] - face[0]) < (xp - face[0]) * (face[4] - face[1])) ||
                    ((yp - face[4]) * (face[6] - face[3]) < (xp - face[3]) * (face[7] - face[4])) ||
                    ((yp - face[7]) * (face[0] - face[6]) < (xp - face[6]) * (face[1] - face[7])))
                    continue;
                float w[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) w[k] = inv[3 * k] * xi + inv[3 * k + 1] * yi + inv[3 * k + 2];
                float wsum = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) { w[k] = (float)fmin(fmax((double)w[k], 0.), 1.); wsum += w[k]; }
#pragma unroll
                for (int k = 0; k < 3; ++k) w[k] /= wsum;
                const float zp = (float)(1. / (double)(w[0] / face[2] + w[1] / face[5] + w[2] / face[8]));
                if (zp <= near || far <= zp) continue;
                if (zp < depth_min) { depth_min = zp; best = s_id[j]; }
            }
        }
        __syncthreads();
    }
    if (xi < is && yi < is) {
        const size_t o = ((size_t)b * is + yi) * is + xi;
        face_index[o] = best;
        alpha[o] = best >= 0 ? 1.f : 0.f;
    }
}

// one warp per (image, face)
__global__ void __launch_bounds__(256) sil_backward_kernel(const float *__restrict__ faces, const int32_t *__restrict__ face_index,
                                                           const float *__restrict__ alpha, const float *__restrict__ g_alpha, int n, int F,
                                                           int is, float eps, float *__restrict__ g_faces) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    const int bn = i / F, fn = i - bn * F;
    const float *face = faces + (size_t)i * 9;
    float g[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) g[k] = 0.f;
    if (!backside(face)) {
        const size_t img = (size_t)bn * is * is;
        for (int edge = 0; edge < 3; ++edge) {
            const int pi0 = edge, pi1 = (edge + 1) % 3, pi2 = (edge + 2) % 3;
            float pp[3][2];
            const int pis[3] = {pi0, pi1, pi2};
#pragma unroll
            for (int num = 0; num < 3; ++num)
#pragma unroll
                for (int dim = 0; dim < 2; ++dim) pp[num][dim] = (float)(0.5 * (double)(face[3 * pis[num] + dim] * is + is - 1));
            for (int axis = 0; axis < 2; ++axis) {
                float p[3][2];
#pragma unroll
                for (int num = 0; num < 3; ++num) { p[num][0] = pp[num][axis]; p[num][1] = pp[num][1 - axis]; }
                const int direction = (axis == 0) ? (p[0][0] < p[1][0] ? -1 : 1) : (p[0][0] < p[1][0] ? 1 : -1);
                const int d0_from = (int)fmax(ceil((double)fminf(p[0][0], p[1][0])), 0.);
                const int d0_to = (int)fmin((double)fmaxf(p[0][0], p[1][0]), is - 1.);
                const int stride = axis == 0 ? is : 1;                 // step of d1 in the maps
                float ga = 0.f, gb = 0.f;                              // gradient on vertex pi0 / pi1, component (1 - axis)
                for (int d0 = d0_from + lane; d0 <= d0_to; d0 += 32) {
                    const float d1_cross = (p[1][1] - p[0][1]) / (p[1][0] - p[0][0]) * (d0 - p[0][0]) + p[0][1];
                    const int d1_in = direction > 0 ? (int)floorf(d1_cross) : (int)ceilf(d1_cross);
                    const int d1_out = d1_in + direction;
                    if (d1_in < 0 || is <= d1_in || d1_out < 0 || is <= d1_out) continue;
                    const size_t col = axis == 0 ? (size_t)d0 : (size_t)d0 * is;      // fixed part of the map index
                    const size_t idx_in = img + col + (size_t)d1_in * stride, idx_out = img + col + (size_t)d1_out * stride;
                    const float alpha_in = alpha[idx_in], alpha_out = alpha[idx_out];
                    const bool c0 = p[1][0] != d0, c1 = p[0][0] != d0;
                    const float k0 = c0 ? (p[1][0] - p[0][0]) / (p[1][0] - d0) : 0.f, k1 = c1 ? (p[1][0] - p[0][0]) / (d0 - p[0][0]) : 0.f;
                    // ---- pixels outside the face, on the far side of the edge ----
                    if (face_index[idx_in] == fn) {
                        const int lim = direction > 0 ? is - 1 : 0;
                        const int from = max(min(d1_out, lim), 0), to = min(max(d1_out, lim), is - 1);
                        for (int d1 = from; d1 <= to; ++d1) {
                            const size_t o = img + col + (size_t)d1 * stride;
                            const float diff = (alpha[o] - alpha_in) * g_alpha[o];
                            if (diff <= 0.f) continue;
                            if (c0) { float dist = (float)((double)(k0 * (d1 - d1_cross)) * 2. / is); dist = 0.f < dist ? dist + eps : dist - eps; ga -= diff / dist; }
                            if (c1) { float dist = (float)((double)(k1 * (d1 - d1_cross)) * 2. / is); dist = 0.f < dist ? dist + eps : dist - eps; gb -= diff / dist; }
                        }
                    }
                    // ---- pixels inside the face ----
                    {
                        float d0_cross2;
                        if ((d0 - p[0][0]) * (d0 - p[2][0]) < 0.f) d0_cross2 = (p[2][1] - p[0][1]) / (p[2][0] - p[0][0]) * (d0 - p[0][0]) + p[0][1];
                        else d0_cross2 = (p[1][1] - p[2][1]) / (p[1][0] - p[2][0]) * (d0 - p[2][0]) + p[2][1];
                        const int lim = direction > 0 ? (int)ceilf(d0_cross2) : (int)floorf(d0_cross2);
                        const int from = max(min(d1_in, lim), 0), to = min(max(d1_in, lim), is - 1);
                        for (int d1 = from; d1 <= to; ++d1) {
                            const size_t o = img + col + (size_t)d1 * stride;
                            if (face_index[o] != fn) continue;
                            const float diff = (alpha[o] - alpha_out) * g_alpha[o];
                            if (diff <= 0.f) continue;
                            if (c0) { float dist = (float)((double)(k0 * (d1 - d1_cross)) * 2. / is); dist = 0.f < dist ? dist + eps : dist - eps; ga -= diff / dist; }
                            if (c1) { float dist = (float)((double)(k1 * (d1 - d1_cross)) * 2. / is); dist = 0.f < dist ? dist + eps : dist - eps; gb -= diff / dist; }
                        }
                    }
                }
                // the two vertices of this edge, component (1 - axis)
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (k == pi0) g[k * 3 + (1 - axis)] += ga;
                    if (k == pi1) g[k * 3 + (1 - axis)] += gb;
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) g[k] = warp_sum(g[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) g_faces[(size_t)i * 9 + k] = g[k];
    }
}

}   // namespace

extern "C" size_t chore_silhouette_workspace_bytes(int B, int F) { return (size_t)B * F * sizeof(FaceAux) + 256; }

extern "C" int chore_silhouette_fwd(chore_handle *h, const float *faces, int B, int F, int image_size, float near, float far,
                                    float *alpha, int32_t *face_index, void *workspace, size_t workspace_bytes, void *stream) {
    CHORE_CHECK(h && faces && alpha && face_index && workspace && B > 0 && F > 0 && image_size > 0, "bad arguments");
    CHORE_CHECK(workspace_bytes >= chore_silhouette_workspace_bytes(B, F), "workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FaceAux *aux = static_cast<FaceAux *>(workspace);
    const int n = B * F;
    CHORE_LAUNCH(sil_setup_kernel, (n + 255) / 256, 256, 0, st, faces, n, image_size, aux);
    const int t = (image_size + kTile - 1) / kTile;
    CHORE_LAUNCH(sil_forward_kernel, dim3(t, t, B), dim3(kTile, kTile), 0, st, faces, aux, F, image_size, near, far, alpha, face_index);
    return CHORE_OK;
}

extern "C" int chore_silhouette_bwd(chore_handle *h, const float *faces, const int32_t *face_index, const float *alpha,
                                    const float *g_alpha, int B, int F, int image_size, float eps, float *g_faces, void *stream) {
    CHORE_CHECK(h && faces && face_index && alpha && g_alpha && g_faces && B > 0 && F > 0 && image_size > 0, "bad arguments");
    const int n = B * F;
    CHORE_LAUNCH(sil_backward_kernel, (n + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream), faces, face_index, alpha, g_alpha, n, F,
                 image_size, eps, g_faces);
    return CHORE_OK;
}
