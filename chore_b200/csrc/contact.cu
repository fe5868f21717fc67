// Joint-phase contact term: ReconFitterBase.compute_contact_loss (recon/recon_fit_base.py:553-608) on the device.
//
// The reference thresholds the cross distance fields at 0.08 m (object distance at the SMPL vertices, human distance at the
// object points), splits the contact points of every image by SMPL part (the vertices carry fixed part labels, the object points
// the argmax of the part head), collects one (human cloud, object cloud) pair per image and part present on both sides, and
// takes pytorch3d.loss.chamfer_distance(Pointclouds(h), Pointclouds(o)) with its defaults over that list: squared distances,
// mean over the points of each cloud, mean over the pairs, both directions added.  (pytorch3d is a third-party dependency that
// is not vendored -- requirements.txt:22 pins no version -- so this restates its documented default reduction.)  Python loops
// over images and parts, boolean indexing and a ragged Pointclouds structure become four launches:
//   contact_flags_kernel   per image: contact counts, per-point (active, part) code, per-part counts of both sides
//   contact_nn_kernel      brute-force nearest neighbour inside the same (image, part), both directions (<= 6890 x 20000 pairs)
//   contact_loss_kernel    the chamfer value (fixed-order reduction) and the number of pairs
//   contact_grad_kernel    closed-form gradient to both point sets (the nearest-neighbour side through fp32 atomics)
#include "common.cuh"

namespace {

constexpr int kParts = 14;

struct ContactWs {
    int32_t *code_h, *code_o;      // (B,Nh) / (B,No): part label, or -1 when the point takes no part in any pair
    int32_t *cnt_h, *cnt_o;        // (B,14): active points per part
    int32_t *nn_h, *nn_o;          // nearest neighbour on the other side (index), -1 = none
    float *d_h, *d_o;              // squared distance to it
    float *partial;                // per-block partial sums of the loss kernel
};

__global__ void __launch_bounds__(1024) contact_flags_kernel(const float *__restrict__ df_hum_o, const float *__restrict__ df_obj_h,
                                                             const float *__restrict__ part_o, const int32_t *__restrict__ part_labels,
                                                             int Nh, int No, float thresh, ContactWs w) {
    __shared__ int ch, co, nh[kParts], no[kParts];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) { ch = 0; co = 0; }
    if (tid < kParts) { nh[tid] = 0; no[tid] = 0; }
    __syncthreads();
    int c = 0;
    for (int i = tid; i < Nh; i += 1024) c += df_hum_o[(size_t)b * Nh + i] < thresh;
    if (c) atomicAdd(&ch, c);
    c = 0;
    for (int i = tid; i < No; i += 1024) c += df_obj_h[(size_t)b * No + i] < thresh;
    if (c) atomicAdd(&co, c);
    __syncthreads();
    const int n_ch = ch, n_co = co;
    const bool skip = n_ch + n_co == 0;                  // no contact on either side: the image contributes nothing (:575-577)
    // without contact points on one side ALL its points are pulled (:578-589)
    for (int i = tid; i < Nh; i += 1024) {
        const bool act = !skip && (n_ch == 0 || df_hum_o[(size_t)b * Nh + i] < thresh);
        const int lab = part_labels[i];
        w.code_h[(size_t)b * Nh + i] = act ? lab : -1;
        if (act) atomicAdd(&nh[lab], 1);
    }
    for (int i = tid; i < No; i += 1024) {
        const bool act = !skip && (n_co == 0 || df_obj_h[(size_t)b * No + i] < thresh);
        int best = 0;
        float bv = part_o[((size_t)b * kParts) * No + i];
#pragma unroll
        for (int p = 1; p < kParts; ++p) {               // torch.argmax: first maximum
            const float v = part_o[((size_t)b * kParts + p) * No + i];
            if (v > bv) { bv = v; best = p; }
        }
        w.code_o[(size_t)b * No + i] = act ? best : -1;
        if (act) atomicAdd(&no[best], 1);
    }
    __syncthreads();
    if (tid < kParts) { w.cnt_h[b * kParts + tid] = nh[tid]; w.cnt_o[b * kParts + tid] = no[tid]; }
}

// query points q (with codes cq) against target points t (codes ct) of the same image: nearest target with the same part code
__global__ void __launch_bounds__(256) contact_nn_kernel(const float *__restrict__ qp, const int32_t *__restrict__ cq, int Nq,
                                                         const float *__restrict__ tp, const int32_t *__restrict__ ct, int Nt,
                                                         const int32_t *__restrict__ cnt_t, int32_t *__restrict__ nn, float *__restrict__ d2) {
    __shared__ float4 tile[256];
    const int b = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
    int code = -1;
    float x = 0.f, y = 0.f, z = 0.f;
    if (i < Nq) {
        code = cq[(size_t)b * Nq + i];
        if (code >= 0 && cnt_t[b * kParts + code] == 0) code = -1;      // the part is absent on the other side: no pair (:593-594)
        const float *p = qp + ((size_t)b * Nq + i) * 3;
        x = p[0]; y = p[1]; z = p[2];
    }
    float best = 3.4e38f;
    int arg = -1;
    for (int t0 = 0; t0 < Nt; t0 += 256) {
        const int t = t0 + threadIdx.x;
        float4 v = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
        if (t < Nt) {
            const float *p = tp + ((size_t)b * Nt + t) * 3;
            v = make_float4(p[0], p[1], p[2], __int_as_float(ct[(size_t)b * Nt + t]));
        }
        __syncthreads();
        tile[threadIdx.x] = v;
        __syncthreads();
        if (code >= 0) {
            const int n = min(256, Nt - t0);
            for (int k = 0; k < n; ++k) {
                const float4 q = tile[k];
                if (__float_as_int(q.w) != code) continue;
                const float dx = x - q.x, dy = y - q.y, dz = z - q.z;
                const float d = dx * dx + dy * dy + dz * dz;
                if (d < best) { best = d; arg = t0 + k; }            // first minimum wins, like a torch.min over the row
            }
        }
    }
    if (i < Nq) { nn[(size_t)b * Nq + i] = arg; d2[(size_t)b * Nq + i] = arg >= 0 ? best : 0.f; }
}

// loss = (1 / n_pairs) sum_pairs [ mean_h d2 + mean_o d2 ]; one block, fixed summation order
__global__ void __launch_bounds__(1024) contact_loss_kernel(int B, int Nh, int No, ContactWs w, float *__restrict__ loss,
                                                            int32_t *__restrict__ n_pairs) {
    __shared__ float red[1024];
    __shared__ int pairs;
    const int tid = threadIdx.x;
    if (tid == 0) {
        int p = 0;
        for (int i = 0; i < B * kParts; ++i) p += (w.cnt_h[i] > 0 && w.cnt_o[i] > 0);
        pairs = p;
    }
    __syncthreads();
    float s = 0.f;
    for (size_t i = tid; i < (size_t)B * Nh; i += 1024)
        if (w.nn_h[i] >= 0) s += w.d_h[i] / (float)w.cnt_h[(i / Nh) * kParts + w.code_h[i]];
    for (size_t i = tid; i < (size_t)B * No; i += 1024)
        if (w.nn_o[i] >= 0) s += w.d_o[i] / (float)w.cnt_o[(i / No) * kParts + w.code_o[i]];
    red[tid] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (tid < o) red[tid] += red[tid + o];
        __syncthreads();
    }
    if (tid == 0) {
        n_pairs[0] = pairs;
        loss[0] = pairs > 0 ? red[0] / (float)pairs : 0.f;
    }
}

// d loss / d points: 2 (p - nn(p)) / (n_pairs * n_side) at p and minus that at nn(p)
__global__ void __launch_bounds__(256) contact_grad_kernel(const float *__restrict__ qp, const float *__restrict__ tp, int Nq, int Nt,
                                                           const int32_t *__restrict__ code, const int32_t *__restrict__ cnt,
                                                           const int32_t *__restrict__ nn, const int32_t *__restrict__ n_pairs,
                                                           float *__restrict__ g_q, float *__restrict__ g_t) {
    const int b = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
    if (i >= Nq) return;
    const int j = nn[(size_t)b * Nq + i];
    if (j < 0 || n_pairs[0] == 0) return;
    const float c = 2.f / ((float)n_pairs[0] * (float)cnt[b * kParts + code[(size_t)b * Nq + i]]);
    const float *p = qp + ((size_t)b * Nq + i) * 3, *t = tp + ((size_t)b * Nt + j) * 3;
    const float gx = c * (p[0] - t[0]), gy = c * (p[1] - t[1]), gz = c * (p[2] - t[2]);
    if (g_q) {
        float *g = g_q + ((size_t)b * Nq + i) * 3;
        atomicAdd(g, gx); atomicAdd(g + 1, gy); atomicAdd(g + 2, gz);
    }
    if (g_t) {
        float *g = g_t + ((size_t)b * Nt + j) * 3;
        atomicAdd(g, -gx); atomicAdd(g + 1, -gy); atomicAdd(g + 2, -gz);
    }
}

size_t ws_layout(int B, int Nh, int No, ContactWs *w, char *base) {
    size_t off = 0;
    auto take = [&](size_t bytes) { char *p = base ? base + off : nullptr; off += (bytes + 255) / 256 * 256; return p; };
    ContactWs t{};
    t.code_h = (int32_t *)take((size_t)B * Nh * 4); t.code_o = (int32_t *)take((size_t)B * No * 4);
    t.cnt_h = (int32_t *)take((size_t)B * kParts * 4); t.cnt_o = (int32_t *)take((size_t)B * kParts * 4);
    t.nn_h = (int32_t *)take((size_t)B * Nh * 4); t.nn_o = (int32_t *)take((size_t)B * No * 4);
    t.d_h = (float *)take((size_t)B * Nh * 4); t.d_o = (float *)take((size_t)B * No * 4);
    t.partial = (float *)take(4096);
    if (w) *w = t;
    return off;
}

}   // namespace

extern "C" size_t chore_contact_workspace_bytes(int B, int Nh, int No) { return ws_layout(B, Nh, No, nullptr, nullptr); }

extern "C" int chore_contact_loss(chore_handle *h, const float *smpl_verts, const float *object, const float *df_hum_o,
                                  const float *df_obj_h, const float *part_o, const int32_t *part_labels, int B, int Nh, int No,
                                  float thresh, float *loss, int32_t *n_pairs, float *g_smpl, float *g_obj, void *workspace,
                                  size_t workspace_bytes, void *stream) {
    CHORE_CHECK(h && smpl_verts && object && df_hum_o && df_obj_h && part_o && part_labels && loss && n_pairs && workspace &&
                B > 0 && Nh > 0 && No > 0, "bad arguments");
    CHORE_CHECK(workspace_bytes >= ws_layout(B, Nh, No, nullptr, nullptr), "workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ContactWs w;
    ws_layout(B, Nh, No, &w, static_cast<char *>(workspace));
    CHORE_LAUNCH(contact_flags_kernel, B, 1024, 0, st, df_hum_o, df_obj_h, part_o, part_labels, Nh, No, thresh, w);
    CHORE_LAUNCH(contact_nn_kernel, dim3((Nh + 255) / 256, B), 256, 0, st, smpl_verts, w.code_h, Nh, object, w.code_o, No, w.cnt_o, w.nn_h, w.d_h);
    CHORE_LAUNCH(contact_nn_kernel, dim3((No + 255) / 256, B), 256, 0, st, object, w.code_o, No, smpl_verts, w.code_h, Nh, w.cnt_h, w.nn_o, w.d_o);
    CHORE_LAUNCH(contact_loss_kernel, 1, 1024, 0, st, B, Nh, No, w, loss, n_pairs);
    if (g_smpl || g_obj) {
        if (g_smpl) CHORE_CUDA(cudaMemsetAsync(g_smpl, 0, (size_t)B * Nh * 3 * sizeof(float), st));
        if (g_obj) CHORE_CUDA(cudaMemsetAsync(g_obj, 0, (size_t)B * No * 3 * sizeof(float), st));
        CHORE_LAUNCH(contact_grad_kernel, dim3((Nh + 255) / 256, B), 256, 0, st, smpl_verts, object, Nh, No, w.code_h, w.cnt_h, w.nn_h, n_pairs, g_smpl, g_obj);
        CHORE_LAUNCH(contact_grad_kernel, dim3((No + 255) / 256, B), 256, 0, st, object, smpl_verts, No, Nh, w.code_o, w.cnt_o, w.nn_o, n_pairs, g_obj, g_smpl);
    }
    return CHORE_OK;
}
