// Loss values and closed-form loss gradients of the two fitting steps, and a single-launch Adam update.
//
// A fit iteration of the reference (recon/recon_fit_behave.py:165-222,293-337 + torch.optim.Adam) is, outside the
// field query and the LBS, a few hundred tiny elementwise / reduction ops on tensors of 3..20 000 elements.  With
// the autograd graph gone (fitter.py:FusedFitSteps) they were still ~200 eager torch launches per iteration, about
// 40 % of its time.  These kernels evaluate the same formulas in 8 launches:
//   fit_smpl_field_kernel      df_h + part cross-entropy terms   -> g_df, g_parts             (recon_fit_base.py:537-542)
//   fit_landmark_kernel        smplz + 2-D keypoint terms        -> g_landmarks                (:230-231, :653-676)
//   fit_pose_prior_kernel      Mahalanobis priors + pinit        -> g_pose (accumulated)       (:522-535, behave :317-319)
//   fit_obj_reduce/finish/grad object, ocent terms               -> g_df, g_centers, dvec      (:513-520, behave :175-198)
//   add_rowvec_kernel          g_obj += coef * dvec
//   adam_step_kernel           torch.optim.Adam (no weight decay, no amsgrad) for <= 8 small tensors
// Loss values are accumulated into one device float in a fixed order (deterministic).
#include "common.cuh"

#include <cstring>

namespace {

constexpr int kRedThreads = 256;

__device__ __forceinline__ float block_sum(float v, float *red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        t = warp_sum(t);
    }
    return t;     // valid in warp 0
}

// ---- SMPL step: field terms ---------------------------------------------------------------------------------
// L = wd * sum_{b,n} min(df_h, 0.1) + wp * sum_{b,n} CE(parts[b,:,n], labels[b,n]);  wd = W_dfh k/(B N), wp = W_part k/B
__global__ void __launch_bounds__(kRedThreads) fit_smpl_field_kernel(const float *__restrict__ df, const float *__restrict__ parts,
                                                                     const long long *__restrict__ labels, int B, int N, float wd, float wp,
                                                                     float *__restrict__ g_df, float *__restrict__ g_parts,
                                                                     float *__restrict__ partials) {
    __shared__ float red[32];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float loss = 0.f;
    if (i < (long long)B * N) {
        const int b = (int)(i / N), n = (int)(i - (long long)b * N);
        const float dfh = df[((size_t)b * 2) * N + n];
        g_df[((size_t)b * 2) * N + n] = dfh <= 0.1f ? wd : 0.f;
        g_df[((size_t)b * 2 + 1) * N + n] = 0.f;
        loss = wd * fminf(dfh, 0.1f);
        float z[CHORE_NUM_PARTS], m = -INFINITY;
#pragma unroll
        for (int c = 0; c < CHORE_NUM_PARTS; ++c) { z[c] = parts[((size_t)b * CHORE_NUM_PARTS + c) * N + n]; m = fmaxf(m, z[c]); }
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < CHORE_NUM_PARTS; ++c) s += expf(z[c] - m);
        const float lse = m + logf(s);
        const int lab = (int)labels[(size_t)b * N + n];
#pragma unroll
        for (int c = 0; c < CHORE_NUM_PARTS; ++c) {
            const float p = expf(z[c] - lse);
            g_parts[((size_t)b * CHORE_NUM_PARTS + c) * N + n] = wp * (p - (c == lab ? 1.f : 0.f));
            if (c == lab) loss += wp * (lse - z[c]);
        }
    }
    const float t = block_sum(loss, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// ordered sum of the block partials -> loss[0] += sum  (one warp; fixed order => reproducible)
__global__ void fit_sum_partials_kernel(const float *__restrict__ partials, int n, float *__restrict__ loss) {
    float v = 0.f;
    for (int i = threadIdx.x; i < n; i += 32) v += partials[i];
    v = warp_sum(v);
    if (threadIdx.x == 0) loss[0] += v;
}

// ---- SMPL step: landmark terms -------------------------------------------------------------------------------
struct CamConsts { float fx, fy, cx, cy, half_crop, scale; };
// one thread per (b, landmark); smplz on body25 joint 8, j2d on the first nJ landmarks
__global__ void fit_landmark_kernel(const float *__restrict__ lm, const float *__restrict__ kpts, const float *__restrict__ cc, int B, int L,
                                    int nJ, float z0, float cz, float cj, CamConsts cam, float *__restrict__ g_lm, float *__restrict__ partials) {
    __shared__ float red[32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float l = 0.f;
    if (i < B * L) {
        const int b = i / L, j = i - b * L;
        const float x = lm[(size_t)i * 3], y = lm[(size_t)i * 3 + 1], z = lm[(size_t)i * 3 + 2];
        float gx = 0.f, gy = 0.f, gz = 0.f;
        if (j == 8) { const float dz = z - z0; l += cz * dz * dz; gz += 2.f * cz * dz; }        // cz = W_smplz k / B
        if (kpts != nullptr && j < nJ) {                                                           // cj = W_j2d k / (B nJ)
            const float px = (cam.fx * x / z + cam.cx + cam.half_crop - cc[b * 2]) * cam.scale;
            const float py = (cam.fy * y / z + cam.cy + cam.half_crop - cc[b * 2 + 1]) * cam.scale;
            const float *kp = kpts + ((size_t)b * nJ + j) * 3;
            const float ex = px - kp[0], ey = py - kp[1], conf = kp[2];
            l += cj * conf * (ex * ex + ey * ey);
            const float c = 2.f * cj * conf * cam.scale;
            const float ax = c * ex * cam.fx / z, ay = c * ey * cam.fy / z;
            gx += ax; gy += ay; gz += -(ax * x + ay * y) / z;
        }
        g_lm[(size_t)i * 3] = gx; g_lm[(size_t)i * 3 + 1] = gy; g_lm[(size_t)i * 3 + 2] = gz;
    }
    const float t = block_sum(l, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// ---- SMPL step: pose priors + initial-pose term ------------------------------------------------------------------
// one block (128 threads) per batch row; body: ||(pose[3:66] - mu) P||^2 * cb,  hands: ||(pose[66:111] - mu_l) P_l||^2 * ch (+ right),
// pinit: cp * ||pose[3:72] - pose_init||^2;  gradients ADDED to g_pose (B,156).  Rows are summed into loss[0] in order by block 0
// of a second launch (fit_sum_partials_kernel).
__global__ void __launch_bounds__(128) fit_pose_prior_kernel(const float *__restrict__ pose, const float *__restrict__ pose_init,
                                                             const float *__restrict__ bmean, const float *__restrict__ bprec,
                                                             const float *__restrict__ hmean, const float *__restrict__ lprec,
                                                             const float *__restrict__ rprec, int npose, float cb, float ch, float cp,
                                                             float *__restrict__ g_pose, float *__restrict__ partials) {
    __shared__ float d[96], t[96], red[32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float *p = pose + (size_t)b * npose;
    float *g = g_pose + (size_t)b * npose;
    float l = 0.f;
    if (bmean != nullptr) {
        // body prior: 63 x 63
        if (tid < 63) d[tid] = p[3 + tid] - bmean[tid];
        __syncthreads();
        if (tid < 63) {
            float a = 0.f;
            for (int i = 0; i < 63; ++i) a = fmaf(d[i], bprec[i * 63 + tid], a);
            t[tid] = a; l += cb * a * a;
        }
        __syncthreads();
        if (tid < 63) {
            float a = 0.f;
            for (int j = 0; j < 63; ++j) a = fmaf(t[j], bprec[tid * 63 + j], a);
            g[3 + tid] += 2.f * cb * a;
        }
        __syncthreads();
        // hand priors: 2 x (45 x 45) on pose[66:156]
        if (tid < 90) d[tid] = p[66 + tid] - hmean[tid];
        __syncthreads();
        if (tid < 90) {
            const int hnd = tid / 45, c = tid - hnd * 45;
            const float *P = hnd ? rprec : lprec;
            float a = 0.f;
            for (int i = 0; i < 45; ++i) a = fmaf(d[hnd * 45 + i], P[i * 45 + c], a);
            t[tid] = a; l += ch * a * a;
        }
        __syncthreads();
        if (tid < 90) {
            const int hnd = tid / 45, c = tid - hnd * 45;
            const float *P = hnd ? rprec : lprec;
            float a = 0.f;
            for (int j = 0; j < 45; ++j) a = fmaf(t[hnd * 45 + j], P[c * 45 + j], a);
            g[66 + tid] += 2.f * ch * a;
        }
    }
    __syncthreads();   // g[66..71] is touched by the hand-prior threads 0..5 above and by the pinit threads 63..68 below
    if (pose_init != nullptr && tid < 69) {
        const float df = p[3 + tid] - pose_init[(size_t)b * 69 + tid];
        l += cp * df * df;
        g[3 + tid] += 2.f * cp * df;
    }
    const float s = block_sum(l, red);
    if (tid == 0) partials[b] = s;
}

// ---- object step -----------------------------------------------------------------------------------------------
// per (block, b): partial sums of obj (3), centers[3:6] (3) and min(df_o, 0.8)
__global__ void __launch_bounds__(kRedThreads) fit_obj_reduce_kernel(const float *__restrict__ obj, const float *__restrict__ df,
                                                                     const float *__restrict__ cen, int N, float *__restrict__ partials) {
    __shared__ float red[32];
    const int b = blockIdx.y;
    float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const float *o = obj + ((size_t)b * N + n) * 3;
        acc[0] += o[0]; acc[1] += o[1]; acc[2] += o[2];
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[3 + c] += cen[((size_t)b * 6 + 3 + c) * N + n];
        acc[6] += fminf(df[((size_t)b * 2 + 1) * N + n], 0.8f);
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const float t = block_sum(acc[k], red);
        if (threadIdx.x == 0) partials[((size_t)b * gridDim.x + blockIdx.x) * 7 + k] = t;
    }
}
// one thread per b, then thread 0 adds the loss terms in order:
//   dvec = mean(obj) - smpl_center - mean(centers[3:6]);  L += wo * sum min(df_o, .8) + wc * sum_b |dvec_b|^2 + ws * sum_b (s_b - s0)^2
__global__ void fit_obj_finish_kernel(const float *__restrict__ partials, int nblk, const float *__restrict__ smpl_center,
                                      const float *__restrict__ s, int B, int N, float s0, float wo, float wc, float ws,
                                      float *__restrict__ dvec, float *__restrict__ loss) {
    __shared__ float lb[64];
    const int b = threadIdx.x;
    if (b < B) {
        float a[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < nblk; ++k)
            for (int c = 0; c < 7; ++c) a[c] += partials[((size_t)b * nblk + k) * 7 + c];
        float l = wo * a[6];
        for (int c = 0; c < 3; ++c) {
            const float dv = a[c] / (float)N - smpl_center[b * 3 + c] - a[3 + c] / (float)N;
            dvec[b * 3 + c] = dv;
            l += wc * dv * dv;
        }
        const float ds = s[b] - s0;
        lb[b] = l + ws * ds * ds;
    }
    __syncthreads();
    if (b == 0) {
        float t = 0.f;
        for (int i = 0; i < B; ++i) t += lb[i];
        loss[0] += t;
    }
}
// g_df[:,0] = 0, g_df[:,1] = (df_o <= 0.8) wo;  g_cen[:, :3] = 0, g_cen[:, 3:] = -coef dvec     (coef = 2 wc / N)
__global__ void __launch_bounds__(256) fit_obj_grad_kernel(const float *__restrict__ df, const float *__restrict__ dvec, int B, int N, float wo,
                                                           float coef, float *__restrict__ g_df, float *__restrict__ g_cen) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * N) return;
    const int b = (int)(i / N), n = (int)(i - (long long)b * N);
    g_df[((size_t)b * 2) * N + n] = 0.f;
    g_df[((size_t)b * 2 + 1) * N + n] = df[((size_t)b * 2 + 1) * N + n] <= 0.8f ? wo : 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        g_cen[((size_t)b * 6 + c) * N + n] = 0.f;
        g_cen[((size_t)b * 6 + 3 + c) * N + n] = -coef * dvec[b * 3 + c];
    }
}
// x[b, n, :] += alpha * v[b, :]
__global__ void __launch_bounds__(256) add_rowvec_kernel(float *__restrict__ x, const float *__restrict__ v, int B, int N, float alpha) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * N * 3) return;
    const int b = (int)(i / ((long long)N * 3)), c = (int)(i % 3);
    x[i] += alpha * v[b * 3 + c];
}

// ---- surface projection step (Generator.approx_surface, recon/generator.py:50-79) --------------------------------
// t = clamp(df[:, k], max=thr); t.sum().backward()  =>  g_df[:, k] = (df[:, k] <= thr), the other channel 0
__global__ void __launch_bounds__(256) surface_clamp_grad_kernel(const float *__restrict__ df, int k, float thr, int B, int N,
                                                                  float *__restrict__ g_df) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * N) return;
    const int b = (int)(i / N), n = (int)(i - (long long)b * N);
    g_df[((size_t)b * 2 + k) * N + n] = df[((size_t)b * 2 + k) * N + n] <= thr ? 1.f : 0.f;
    g_df[((size_t)b * 2 + (1 - k)) * N + n] = 0.f;
}
// p <- p - normalize(g, eps = 1e-12) * min(df[:, k], thr)     (F.normalize: g / max(||g||_2, eps))
__global__ void __launch_bounds__(256) surface_step_kernel(const float *__restrict__ pts, const float *__restrict__ g, const float *__restrict__ df,
                                                            int k, float thr, int B, int N, float *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * N) return;
    const int b = (int)(i / N), n = (int)(i - (long long)b * N);
    const float gx = g[i * 3], gy = g[i * 3 + 1], gz = g[i * 3 + 2];
    const float nrm = fmaxf(sqrtf(gx * gx + gy * gy + gz * gz), 1e-12f);
    const float t = fminf(df[((size_t)b * 2 + k) * N + n], thr);
    out[i * 3] = pts[i * 3] - gx / nrm * t;
    out[i * 3 + 1] = pts[i * 3 + 1] - gy / nrm * t;
    out[i * 3 + 2] = pts[i * 3 + 2] - gz / nrm * t;
}

// ---- Adam ------------------------------------------------------------------------------------------------------
struct AdamArgs {
    chore_adam_entry e[CHORE_ADAM_MAX_ENTRIES];
    int n;
};
// torch.optim.Adam(betas=(b1,b2), eps, weight_decay=0, amsgrad=False), single block so that the device-side step
// counter (needed under CUDA-graph replay: kernel arguments are frozen) is read by every thread before it is bumped.
__global__ void __launch_bounds__(256) adam_step_kernel(const AdamArgs a, float lr, float b1, float b2, float eps, int *__restrict__ step,
                                                        const float *__restrict__ gscale, float *__restrict__ loss_inout) {
    const int t = step[0] + 1;
    const float gs = gscale != nullptr ? gscale[0] : 1.f;
    const float bc1 = 1.f - powf(b1, (float)t), bc2 = 1.f - powf(b2, (float)t);
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    for (int k = 0; k < a.n; ++k) {
        const chore_adam_entry e = a.e[k];
        for (int i = threadIdx.x; i < e.rows * e.cols; i += blockDim.x) {
            const int r = i / e.cols, c = i - r * e.cols;
            float g = gs * e.grad[(size_t)r * e.grad_ld + c];
            if (e.grad_acc != nullptr) { g += e.grad_acc[i]; e.grad_acc[i] = g; }   // sum since the last zero_grad()
            const float m = b1 * e.exp_avg[i] + (1.f - b1) * g;
            const float v = b2 * e.exp_avg_sq[i] + (1.f - b2) * g * g;
            e.exp_avg[i] = m; e.exp_avg_sq[i] = v;
            e.param[i] -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + eps);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        step[0] = t;
        if (loss_inout != nullptr) loss_inout[0] *= gs;
    }
}

}   // namespace

extern "C" size_t chore_fit_workspace_floats(int B, int N) {
    const size_t blocks = ((size_t)B * N + kRedThreads - 1) / kRedThreads;
    return blocks + (size_t)B * 64 * 7 + 256;
}

extern "C" int chore_fit_smpl_field_grads(chore_handle *h, const float *df, const float *parts, const int64_t *labels, int B, int N,
                                          float wd, float wp, float *g_df, float *g_parts, float *loss, float *workspace, void *stream) {
    CHORE_CHECK(h && df && parts && labels && g_df && g_parts && loss && workspace && B > 0 && N > 0, "bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int blocks = (int)(((long long)B * N + kRedThreads - 1) / kRedThreads);
    CHORE_LAUNCH(fit_smpl_field_kernel, blocks, kRedThreads, 0, st, df, parts, reinterpret_cast<const long long *>(labels), B, N, wd, wp, g_df,
                 g_parts, workspace);
    CHORE_LAUNCH(fit_sum_partials_kernel, 1, 32, 0, st, workspace, blocks, loss);
    return CHORE_OK;
}

extern "C" int chore_fit_landmark_grads(chore_handle *h, const float *landmarks, const float *body_kpts, const float *crop_center, int B, int L,
                                        int n_joints, float z0, float cz, float cj, const float cam[6], float *g_landmarks, float *loss,
                                        float *workspace, void *stream) {
    CHORE_CHECK(h && landmarks && crop_center && cam && g_landmarks && loss && workspace && B > 0 && L > 8 && n_joints <= L, "bad arguments");
    CamConsts c{cam[0], cam[1], cam[2], cam[3], cam[4], cam[5]};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int blocks = (B * L + kRedThreads - 1) / kRedThreads;
    CHORE_CHECK(blocks <= 256, "too many landmarks (%d x %d)", B, L);
    CHORE_LAUNCH(fit_landmark_kernel, blocks, kRedThreads, 0, st, landmarks, body_kpts, crop_center, B, L, n_joints, z0, cz, cj, c,
                 g_landmarks, workspace);
    CHORE_LAUNCH(fit_sum_partials_kernel, 1, 32, 0, st, workspace, blocks, loss);
    return CHORE_OK;
}

extern "C" int chore_fit_pose_prior_grads(chore_handle *h, const float *pose, const float *pose_init, const float *body_mean,
                                          const float *body_prec, const float *hand_mean, const float *lhand_prec, const float *rhand_prec, int B,
                                          int n_pose, float cb, float ch, float cp, float *g_pose, float *loss, float *workspace, void *stream) {
    CHORE_CHECK(h && pose && g_pose && loss && workspace && B > 0 && n_pose == 156, "bad arguments (SMPL-H pose of 156 expected)");
    CHORE_CHECK(body_mean == nullptr || (body_prec && hand_mean && lhand_prec && rhand_prec), "incomplete priors");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CHORE_LAUNCH(fit_pose_prior_kernel, B, 128, 0, st, pose, pose_init, body_mean, body_prec, hand_mean, lhand_prec, rhand_prec, n_pose, cb, ch,
                 cp, g_pose, workspace);
    CHORE_LAUNCH(fit_sum_partials_kernel, 1, 32, 0, st, workspace, B, loss);
    return CHORE_OK;
}

extern "C" int chore_fit_obj_field_grads(chore_handle *h, const float *obj, const float *df, const float *centers, const float *smpl_center,
                                         const float *s, int B, int N, float s0, float wo, float wc, float ws, float *g_df, float *g_centers,
                                         float *dvec, float *loss, float *workspace, void *stream) {
    CHORE_CHECK(h && obj && df && centers && smpl_center && s && g_df && g_centers && dvec && loss && workspace && B > 0 && B <= 64 && N > 0,
                "bad arguments (B <= 64)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int nblk = (N + kRedThreads * 4 - 1) / (kRedThreads * 4);
    nblk = nblk < 1 ? 1 : (nblk > 64 ? 64 : nblk);
    CHORE_LAUNCH(fit_obj_reduce_kernel, dim3(nblk, B), kRedThreads, 0, st, obj, df, centers, N, workspace);
    CHORE_LAUNCH(fit_obj_finish_kernel, 1, 64, 0, st, workspace, nblk, smpl_center, s, B, N, s0, wo, wc, ws, dvec, loss);
    const int blocks = (int)(((long long)B * N + 255) / 256);
    CHORE_LAUNCH(fit_obj_grad_kernel, blocks, 256, 0, st, df, dvec, B, N, wo, 2.f * wc / (float)N, g_df, g_centers);
    return CHORE_OK;
}

extern "C" int chore_add_rowvec(chore_handle *h, float *x, const float *v, int B, int N, float alpha, void *stream) {
    CHORE_CHECK(h && x && v && B > 0 && N > 0, "bad arguments");
    const long long n = (long long)B * N * 3;
    CHORE_LAUNCH(add_rowvec_kernel, (unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream), x, v, B, N, alpha);
    return CHORE_OK;
}

extern "C" int chore_adam_step(chore_handle *h, const chore_adam_entry *entries, int n, float lr, float beta1, float beta2, float eps,
                               int32_t *step, const float *gscale, float *loss_inout, void *stream) {
    CHORE_CHECK(h && entries && step && n > 0 && n <= CHORE_ADAM_MAX_ENTRIES, "bad arguments (1..%d entries)", CHORE_ADAM_MAX_ENTRIES);
    AdamArgs a{};
    a.n = n;
    for (int i = 0; i < n; ++i) {
        CHORE_CHECK(entries[i].param && entries[i].grad && entries[i].exp_avg && entries[i].exp_avg_sq && entries[i].rows > 0 &&
                    entries[i].cols > 0 && entries[i].grad_ld >= entries[i].cols, "entry %d is malformed", i);
        a.e[i] = entries[i];
    }
    CHORE_LAUNCH(adam_step_kernel, 1, 256, 0, static_cast<cudaStream_t>(stream), a, lr, beta1, beta2, eps, step, gscale, loss_inout);
    return CHORE_OK;
}

extern "C" int chore_surface_clamp_grad(chore_handle *h, const float *df, int df_idx, float threshold, int B, int N, float *g_df,
                                        void *stream) {
    CHORE_CHECK(h && df && g_df && B > 0 && N > 0 && (df_idx == 0 || df_idx == 1), "bad arguments");
    CHORE_LAUNCH(surface_clamp_grad_kernel, (unsigned)(((long long)B * N + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream), df, df_idx,
                 threshold, B, N, g_df);
    return CHORE_OK;
}

extern "C" int chore_surface_step(chore_handle *h, const float *points, const float *g_points, const float *df, int df_idx, float threshold,
                                  int B, int N, float *out_points, void *stream) {
    CHORE_CHECK(h && points && g_points && df && out_points && B > 0 && N > 0 && (df_idx == 0 || df_idx == 1), "bad arguments");
    CHORE_LAUNCH(surface_step_kernel, (unsigned)(((long long)B * N + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream), points, g_points, df,
                 df_idx, threshold, B, N, out_points);
    return CHORE_OK;
}
