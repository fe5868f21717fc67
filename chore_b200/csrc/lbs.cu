// SMPL-H linear blend skinning (forward + analytic backward), rigid object transform and SO(3)
// projection.
//
// Replaces SMPL_Layer.forward (lib_smpl/smplpytorch/smplpytorch/pytorch/smpl_layer.py:72-175:
// ~600 eager launches per call: 52 Rodrigues in a Python loop, 51 chained 4x4 matmuls, three
// dense matmuls and the skinning product), th_posemap_axisang (tensutils.py:6-19),
// batch_rodrigues/quat2mat (rodrigues_layer.py:13-52), ReconFitterBase.transform_obj_verts
// and project_so3 (recon/recon_fit_base.py:167-188,367-371) by
//   lbs_pose_kernel    1 CTA / body : Rodrigues, joints, kinematic chain -> A[52][3x4], jtr
//   lbs_vertex_kernel  1 warp / vertex: shape + pose blend shapes (streams posedirs), skinning
// and their adjoints.  Nothing is saved by the forward; the backward recomputes.
#include "common.cuh"

#include <cstring>

namespace {

constexpr int kMaxJ = 64;        // joints held in shared memory (SMPL-H: 52)
constexpr int kMaxBetas = 16;
constexpr int kVertsPerCta = 32; // 8 warps x 4 vertices

struct Rod {                     // intermediates of one Rodrigues evaluation
    float ang, s, c, m;          // |r + 1e-8|, sin, cos of ang/2, |q|
    float n[3];                  // r / ang
    float q[4];                  // normalised quaternion (w, x, y, z)
};

// rodrigues_layer.py:13-52: theta = ||r + 1e-8||, q = (cos t/2, sin t/2 * r/theta) renormalised
__device__ __forceinline__ void rodrigues(const float r[3], float R[9], Rod &o) {
    const float a0 = r[0] + 1e-8f, a1 = r[1] + 1e-8f, a2 = r[2] + 1e-8f;
    o.ang = sqrtf(a0 * a0 + a1 * a1 + a2 * a2);
    o.n[0] = r[0] / o.ang; o.n[1] = r[1] / o.ang; o.n[2] = r[2] / o.ang;
    const float half = o.ang * 0.5f;
    o.c = cosf(half); o.s = sinf(half);
    float q0 = o.c, q1 = o.s * o.n[0], q2 = o.s * o.n[1], q3 = o.s * o.n[2];
    o.m = sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    const float w = q0 / o.m, x = q1 / o.m, y = q2 / o.m, z = q3 / o.m;
    o.q[0] = w; o.q[1] = x; o.q[2] = y; o.q[3] = z;
    const float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
    const float wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
    R[0] = w2 + x2 - y2 - z2; R[1] = 2 * xy - 2 * wz;   R[2] = 2 * wy + 2 * xz;
    R[3] = 2 * wz + 2 * xy;   R[4] = w2 - x2 + y2 - z2; R[5] = 2 * yz - 2 * wx;
    R[6] = 2 * xz - 2 * wy;   R[7] = 2 * wx + 2 * yz;   R[8] = w2 - x2 - y2 + z2;
}

// adjoint of rodrigues(): gR (9) -> gr (3)
__device__ __forceinline__ void rodrigues_bwd(const float r[3], const Rod &o, const float g[9], float gr[3]) {
    const float w = o.q[0], x = o.q[1], y = o.q[2], z = o.q[3];
    float gq[4];
    gq[0] = 2 * w * (g[0] + g[4] + g[8]) + 2 * (-z * g[1] + y * g[2] + z * g[3] - x * g[5] - y * g[6] + x * g[7]);
    gq[1] = 2 * x * (g[0] - g[4] - g[8]) + 2 * (y * g[1] + z * g[2] + y * g[3] - w * g[5] + z * g[6] + w * g[7]);
    gq[2] = 2 * y * (-g[0] + g[4] - g[8]) + 2 * (x * g[1] + w * g[2] + x * g[3] + z * g[5] - w * g[6] + z * g[7]);
    gq[3] = 2 * z * (-g[0] - g[4] + g[8]) + 2 * (-w * g[1] + x * g[2] + w * g[3] + y * g[5] + x * g[6] + y * g[7]);
    // through q / |q|
    const float dot = gq[0] * w + gq[1] * x + gq[2] * y + gq[3] * z;
    float gu[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) gu[i] = (gq[i] - o.q[i] * dot) / o.m;
    // u = (c, s*n)
    const float gc = gu[0];
    const float gs = gu[1] * o.n[0] + gu[2] * o.n[1] + gu[3] * o.n[2];
    const float gn[3] = {gu[1] * o.s, gu[2] * o.s, gu[3] * o.s};
    const float ghalf = -o.s * gc + o.c * gs;
    // n = r / ang ; ang = |r + 1e-8|
    const float inv = 1.f / o.ang;
    float gang = 0.5f * ghalf - (gn[0] * r[0] + gn[1] * r[1] + gn[2] * r[2]) * inv * inv;
#pragma unroll
    for (int i = 0; i < 3; ++i) gr[i] = gn[i] * inv + gang * (r[i] + 1e-8f) * inv;
}

struct LbsParams {
    int V, J, nb, P;   // P = (J-1)*9
    const float *v_template, *shapedirs, *posedirs, *weights, *j_template, *j_shapedirs;
    const int32_t *parents;
    const float *pose, *betas, *trans, *offsets;
    int B;
    float *posemap;    // [B][P]
    float *A;          // [B][J][12]  skinning transforms (rest pose removed), row-major 3x4
    float *G;          // [B][J][12]  global joint transforms
    float *joints;     // [B][J][3]   rest joints
    float *verts, *jtr, *v_posed, *naked;
    // backward
    const float *g_verts, *g_jtr;
    float *gA;         // [B][J][12]
    float *gpm;        // [B][P]
    float *gbetas_acc; // [B][nb]
    float *gtrans_acc; // [B][3]
    float *g_pose, *g_betas, *g_trans, *g_offsets;
};

// ---- per body: Rodrigues, rest joints, kinematic chain --------------------------------------------
__global__ void __launch_bounds__(64) lbs_pose_kernel(const LbsParams p) {
    __shared__ float Rs[kMaxJ][9], Js[kMaxJ][3], Gs[kMaxJ][12];
    const int b = blockIdx.x, j = threadIdx.x;
    if (j < p.J) {
        float r[3] = {p.pose[(size_t)b * p.J * 3 + j * 3], p.pose[(size_t)b * p.J * 3 + j * 3 + 1],
                      p.pose[(size_t)b * p.J * 3 + j * 3 + 2]};
        float R[9];
        Rod tmp;
        rodrigues(r, R, tmp);
#pragma unroll
        for (int i = 0; i < 9; ++i) Rs[j][i] = R[i];
        if (j >= 1) {   // pose_map = R - I (tensutils.py:41-53)
#pragma unroll
            for (int i = 0; i < 9; ++i) p.posemap[(size_t)b * p.P + (j - 1) * 9 + i] = R[i] - ((i % 4 == 0) ? 1.f : 0.f);
        }
        // J = J_regressor . (v_template + shapedirs . beta)  (smpl_layer.py:102-103), regressor pre-applied
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float s = 0.f;
            for (int k = 0; k < p.nb; ++k) s = fmaf(p.j_shapedirs[((size_t)j * 3 + c) * p.nb + k], p.betas[(size_t)b * p.nb + k], s);
            Js[j][c] = p.j_template[j * 3 + c] + s;
            p.joints[((size_t)b * p.J + j) * 3 + c] = Js[j][c];
        }
    }
    __syncthreads();
    if (j == 0) {   // 51 dependent 3x4 products (smpl_layer.py:116-130)
#pragma unroll
        for (int i = 0; i < 9; ++i) Gs[0][(i / 3) * 4 + (i % 3)] = Rs[0][i];
        Gs[0][3] = Js[0][0]; Gs[0][7] = Js[0][1]; Gs[0][11] = Js[0][2];
        for (int i = 1; i < p.J; ++i) {
            const int par = p.parents[i];
            const float d[3] = {Js[i][0] - Js[par][0], Js[i][1] - Js[par][1], Js[i][2] - Js[par][2]};
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float g0 = Gs[par][r * 4], g1 = Gs[par][r * 4 + 1], g2 = Gs[par][r * 4 + 2], g3 = Gs[par][r * 4 + 3];
#pragma unroll
                for (int c = 0; c < 3; ++c) Gs[i][r * 4 + c] = (g0 * Rs[i][c] + g1 * Rs[i][3 + c]) + g2 * Rs[i][6 + c];
                Gs[i][r * 4 + 3] = ((g0 * d[0] + g1 * d[1]) + g2 * d[2]) + g3;
            }
        }
    }
    __syncthreads();
    if (j < p.J) {
        float *Ao = p.A + ((size_t)b * p.J + j) * 12, *Go = p.G + ((size_t)b * p.J + j) * 12;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            // remove the rest pose: A = G - pack(G . [J;0])  (smpl_layer.py:135-142)
            const float t = (Gs[j][r * 4] * Js[j][0] + Gs[j][r * 4 + 1] * Js[j][1]) + Gs[j][r * 4 + 2] * Js[j][2];
#pragma unroll
            for (int c = 0; c < 3; ++c) Ao[r * 4 + c] = Gs[j][r * 4 + c];
            Ao[r * 4 + 3] = Gs[j][r * 4 + 3] - t;
#pragma unroll
            for (int c = 0; c < 4; ++c) Go[r * 4 + c] = Gs[j][r * 4 + c];
            if (p.jtr) p.jtr[((size_t)b * p.J + j) * 3 + r] = Gs[j][r * 4 + 3] + p.trans[b * 3 + r];
        }
    }
}

// v_posed of vertex v (warp-cooperative): returns the same value on every lane
__device__ __forceinline__ void blend_vertex(const LbsParams &p, int b, int v, const float *pm_s, const float *beta_s,
                                             int lane, float vp[3], float nk[3]) {
    float acc[3] = {0.f, 0.f, 0.f};
    const float *prow = p.posedirs + (size_t)v * 3 * p.P;
#pragma unroll 4
    for (int k = lane; k < p.P; k += 32) {
        const float m = pm_s[k];
        acc[0] = fmaf(__ldg(prow + k), m, acc[0]);
        acc[1] = fmaf(__ldg(prow + p.P + k), m, acc[1]);
        acc[2] = fmaf(__ldg(prow + 2 * p.P + k), m, acc[2]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        acc[c] = warp_sum(acc[c]);
        float s = 0.f;
        const float *srow = p.shapedirs + ((size_t)v * 3 + c) * p.nb;
        for (int k = 0; k < p.nb; ++k) s = fmaf(__ldg(srow + k), beta_s[k], s);
        const float vs = __ldg(p.v_template + (size_t)v * 3 + c) + s;
        nk[c] = vs + acc[c];
        vp[c] = nk[c] + (p.offsets ? __ldg(p.offsets + ((size_t)b * p.V + v) * 3 + c) : 0.f);
    }
}

// blended skinning transform T = sum_j w[v][j] A_j (3x4), same value on every lane
__device__ __forceinline__ void blend_transform(const LbsParams &p, int v, const float *A_s, int lane, float T[12]) {
#pragma unroll
    for (int i = 0; i < 12; ++i) T[i] = 0.f;
    for (int j = lane; j < p.J; j += 32) {
        const float w = __ldg(p.weights + (size_t)v * p.J + j);
        if (w != 0.f) {
#pragma unroll
            for (int i = 0; i < 12; ++i) T[i] = fmaf(w, A_s[j * 12 + i], T[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) T[i] = warp_sum(T[i]);
}

__global__ void __launch_bounds__(256) lbs_vertex_kernel(const LbsParams p) {
    extern __shared__ float sm[];
    float *pm_s = sm;                    // [P]
    float *A_s = pm_s + p.P;             // [J][12]
    float *beta_s = A_s + p.J * 12;      // [nb]
    const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < p.P; i += 256) pm_s[i] = p.posemap[(size_t)b * p.P + i];
    for (int i = tid; i < p.J * 12; i += 256) A_s[i] = p.A[(size_t)b * p.J * 12 + i];
    if (tid < p.nb) beta_s[tid] = p.betas[(size_t)b * p.nb + tid];
    __syncthreads();
    const int v0 = blockIdx.x * kVertsPerCta;
    for (int v = v0 + warp; v < min(p.V, v0 + kVertsPerCta); v += 8) {
        float vp[3], nk[3], T[12];
        blend_vertex(p, b, v, pm_s, beta_s, lane, vp, nk);
        blend_transform(p, v, A_s, lane, T);
        if (lane < 3) {
            const int r = lane;
            const float out = ((T[r * 4] * vp[0] + T[r * 4 + 1] * vp[1]) + T[r * 4 + 2] * vp[2]) + T[r * 4 + 3];
            const size_t o = ((size_t)b * p.V + v) * 3 + r;
            p.verts[o] = out + p.trans[b * 3 + r];
            if (p.v_posed) p.v_posed[o] = vp[r];
            if (p.naked) p.naked[o] = nk[r];
        }
    }
}

// ---- backward, vertex side: g_verts -> gA, g_posemap, g_betas (shape part), g_trans, g_offsets --------
__global__ void __launch_bounds__(256) lbs_vertex_bwd_kernel(const LbsParams p) {
    extern __shared__ float sm[];
    float *pm_s = sm;                      // [P]
    float *A_s = pm_s + p.P;               // [J][12]
    float *beta_s = A_s + p.J * 12;        // [nb] (padded to 16)
    float *gA_s = beta_s + kMaxBetas;      // [J][12]
    float *gpm_s = gA_s + p.J * 12;        // [P]
    float *gb_s = gpm_s + p.P;             // [16]
    float *gt_s = gb_s + kMaxBetas;        // [4]
    const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < p.P; i += 256) { pm_s[i] = p.posemap[(size_t)b * p.P + i]; gpm_s[i] = 0.f; }
    for (int i = tid; i < p.J * 12; i += 256) { A_s[i] = p.A[(size_t)b * p.J * 12 + i]; gA_s[i] = 0.f; }
    if (tid < kMaxBetas) { beta_s[tid] = tid < p.nb ? p.betas[(size_t)b * p.nb + tid] : 0.f; gb_s[tid] = 0.f; }
    if (tid < 4) gt_s[tid] = 0.f;
    __syncthreads();
    const int v0 = blockIdx.x * kVertsPerCta;
    float gt[3] = {0.f, 0.f, 0.f};
    for (int v = v0 + warp; v < min(p.V, v0 + kVertsPerCta); v += 8) {
        float vp[3], nk[3], T[12];
        blend_transform(p, v, A_s, lane, T);
        const float *gv = p.g_verts + ((size_t)b * p.V + v) * 3;
        const float g[3] = {__ldg(gv), __ldg(gv + 1), __ldg(gv + 2)};
        gt[0] += g[0]; gt[1] += g[1]; gt[2] += g[2];
        // g_vposed = T.R^T g (does not depend on the vertex position: known before the pose blend shapes are read)
        float gvp[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) gvp[c] = T[c] * g[0] + T[4 + c] * g[1] + T[8 + c] * g[2];
        if (p.g_offsets && lane < 3) p.g_offsets[((size_t)b * p.V + v) * 3 + lane] = gvp[lane];
        // ONE pass over the 3 x P pose blend shapes of the vertex (5.5 KB, L2 resident): the forward recompute
        // v_posed += P[v][c][k] posemap[k] and the adjoint g_posemap[k] += sum_c P[v][c][k] gvp[c] share every load
        {
            float acc[3] = {0.f, 0.f, 0.f};
            const float *prow = p.posedirs + (size_t)v * 3 * p.P;
#pragma unroll 4
            for (int k = lane; k < p.P; k += 32) {
                const float p0 = __ldg(prow + k), p1 = __ldg(prow + p.P + k), p2 = __ldg(prow + 2 * p.P + k);
                const float m = pm_s[k];
                acc[0] = fmaf(p0, m, acc[0]); acc[1] = fmaf(p1, m, acc[1]); acc[2] = fmaf(p2, m, acc[2]);
                atomicAdd(&gpm_s[k], p0 * gvp[0] + p1 * gvp[1] + p2 * gvp[2]);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {                 // same operation order as blend_vertex
                acc[c] = warp_sum(acc[c]);
                float sb = 0.f;
                const float *srow = p.shapedirs + ((size_t)v * 3 + c) * p.nb;
                for (int k = 0; k < p.nb; ++k) sb = fmaf(__ldg(srow + k), beta_s[k], sb);
                const float vs = __ldg(p.v_template + (size_t)v * 3 + c) + sb;
                nk[c] = vs + acc[c];
                vp[c] = nk[c] + (p.offsets ? __ldg(p.offsets + ((size_t)b * p.V + v) * 3 + c) : 0.f);
            }
        }
        // gA_j += w_vj * g (x) [vp;1]
        for (int j = lane; j < p.J; j += 32) {
            const float w = __ldg(p.weights + (size_t)v * p.J + j);
            if (w != 0.f) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    atomicAdd(&gA_s[j * 12 + r * 4 + 0], w * g[r] * vp[0]);
                    atomicAdd(&gA_s[j * 12 + r * 4 + 1], w * g[r] * vp[1]);
                    atomicAdd(&gA_s[j * 12 + r * 4 + 2], w * g[r] * vp[2]);
                    atomicAdd(&gA_s[j * 12 + r * 4 + 3], w * g[r]);
                }
            }
        }
        // g_betas[k] += sum_c S[v][c][k] gvp[c]
        if (lane < p.nb) {
            const float *srow = p.shapedirs + (size_t)v * 3 * p.nb;
            const float s = __ldg(srow + lane) * gvp[0] + __ldg(srow + p.nb + lane) * gvp[1] + __ldg(srow + 2 * p.nb + lane) * gvp[2];
            atomicAdd(&gb_s[lane], s);
        }
    }
    if (lane == 0) { atomicAdd(&gt_s[0], gt[0]); atomicAdd(&gt_s[1], gt[1]); atomicAdd(&gt_s[2], gt[2]); }
    __syncthreads();
    for (int i = tid; i < p.P; i += 256) atomicAdd(&p.gpm[(size_t)b * p.P + i], gpm_s[i]);
    for (int i = tid; i < p.J * 12; i += 256) atomicAdd(&p.gA[(size_t)b * p.J * 12 + i], gA_s[i]);
    if (tid < p.nb) atomicAdd(&p.gbetas_acc[(size_t)b * p.nb + tid], gb_s[tid]);
    if (tid < 3) atomicAdd(&p.gtrans_acc[b * 3 + tid], gt_s[tid]);
}

// ---- backward, body side: chain adjoint, Rodrigues adjoint -------------------------------------------
__global__ void __launch_bounds__(64) lbs_pose_bwd_kernel(const LbsParams p) {
    __shared__ float Rs[kMaxJ][9], Js[kMaxJ][3], Gs[kMaxJ][12], gG[kMaxJ][12], gR[kMaxJ][9], gJ[kMaxJ][3];
    const int b = blockIdx.x, j = threadIdx.x;
    float r[3] = {0.f, 0.f, 0.f};
    Rod rod{};
    if (j < p.J) {
        r[0] = p.pose[(size_t)b * p.J * 3 + j * 3]; r[1] = p.pose[(size_t)b * p.J * 3 + j * 3 + 1];
        r[2] = p.pose[(size_t)b * p.J * 3 + j * 3 + 2];
        float R[9];
        rodrigues(r, R, rod);
#pragma unroll
        for (int i = 0; i < 9; ++i) Rs[j][i] = R[i];
        const float *Gi = p.G + ((size_t)b * p.J + j) * 12;
        const float *gAi = p.gA + ((size_t)b * p.J + j) * 12;
#pragma unroll
        for (int c = 0; c < 3; ++c) Js[j][c] = p.joints[((size_t)b * p.J + j) * 3 + c];
#pragma unroll
        for (int i = 0; i < 12; ++i) Gs[j][i] = Gi[i];
        // A.R = G.R ; A.t = G.t - G.R J ; jtr = G.t (+trans)
        float gAt[3], gj[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) gAt[rr] = gAi[rr * 4 + 3];
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                gG[j][rr * 4 + c] = gAi[rr * 4 + c] - gAt[rr] * Js[j][c];
                gj[c] -= Gi[rr * 4 + c] * gAt[rr];
            }
            gG[j][rr * 4 + 3] = gAt[rr] + (p.g_jtr ? p.g_jtr[((size_t)b * p.J + j) * 3 + rr] : 0.f);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) gJ[j][c] = gj[c];
    }
    __syncthreads();
    if (j == 0) {
        for (int i = p.J - 1; i >= 1; --i) {
            const int par = p.parents[i];
            const float d[3] = {Js[i][0] - Js[par][0], Js[i][1] - Js[par][1], Js[i][2] - Js[par][2]};
            // G_i.R = Gp.R R_i ; G_i.t = Gp.R d + Gp.t
            float gd[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int a = 0; a < 3; ++a) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    // gR_i[a][c] = sum_r Gp.R[r][a] gG_i.R[r][c]
                    gR[i][a * 3 + c] = Gs[par][0 * 4 + a] * gG[i][0 * 4 + c] + Gs[par][1 * 4 + a] * gG[i][1 * 4 + c] +
                                       Gs[par][2 * 4 + a] * gG[i][2 * 4 + c];
                    // gGp.R[a][c] += sum_k gG_i.R[a][k] R_i[c][k] + gG_i.t[a] d[c]
                    gG[par][a * 4 + c] += gG[i][a * 4 + 0] * Rs[i][c * 3 + 0] + gG[i][a * 4 + 1] * Rs[i][c * 3 + 1] +
                                          gG[i][a * 4 + 2] * Rs[i][c * 3 + 2] + gG[i][a * 4 + 3] * d[c];
                }
                gd[a] = Gs[par][0 * 4 + a] * gG[i][3] + Gs[par][1 * 4 + a] * gG[i][7] + Gs[par][2 * 4 + a] * gG[i][11];
                gG[par][a * 4 + 3] += gG[i][a * 4 + 3];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) { gJ[i][c] += gd[c]; gJ[par][c] -= gd[c]; }
        }
        // root: G_0.R = R_0 ; G_0.t = J_0
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int c = 0; c < 3; ++c) gR[0][a * 3 + c] = gG[0][a * 4 + c];
            gJ[0][a] += gG[0][a * 4 + 3];
        }
    }
    __syncthreads();
    if (j < p.J) {
        float g[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) g[i] = gR[j][i] + (j >= 1 ? p.gpm[(size_t)b * p.P + (j - 1) * 9 + i] : 0.f);
        float gr[3];
        rodrigues_bwd(r, rod, g, gr);
#pragma unroll
        for (int c = 0; c < 3; ++c) p.g_pose[(size_t)b * p.J * 3 + j * 3 + c] = gr[c];
    }
    if (j < p.nb) {   // betas: shape blend (accumulated by the vertex kernel) + joint regressor path
        float s = p.gbetas_acc[(size_t)b * p.nb + j];
        for (int i = 0; i < p.J; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) s = fmaf(p.j_shapedirs[((size_t)i * 3 + c) * p.nb + j], gJ[i][c], s);
        p.g_betas[(size_t)b * p.nb + j] = s;
    }
    if (j < 3) {
        float s = p.gtrans_acc[b * 3 + j];
        if (p.g_jtr)
            for (int i = 0; i < p.J; ++i) s += p.g_jtr[((size_t)b * p.J + i) * 3 + j];
        p.g_trans[b * 3 + j] = s;
    }
}

// ---- rigid transform (recon_fit_base.py:367-371): out = (v R + t) * s ----------------------------------
__global__ void __launch_bounds__(256) rigid_fwd_kernel(const float *__restrict__ verts, const float *__restrict__ R,
                                                        const float *__restrict__ t, const float *__restrict__ s,
                                                        int N, float *__restrict__ out) {
    const int b = blockIdx.y, n = blockIdx.x * 256 + threadIdx.x;
    if (n >= N) return;
    const float *Rb = R + b * 9;
    const float *v = verts + ((size_t)b * N + n) * 3;
    const float x = v[0], y = v[1], z = v[2], sc = s[b];
    float *o = out + ((size_t)b * N + n) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = ((((x * Rb[c] + y * Rb[3 + c]) + z * Rb[6 + c])) + t[b * 3 + c]) * sc;
}

__global__ void __launch_bounds__(256) rigid_bwd_kernel(const float *__restrict__ verts, const float *__restrict__ R,
                                                        const float *__restrict__ t, const float *__restrict__ s,
                                                        int N, const float *__restrict__ g_out, float *g_R, float *g_t,
                                                        float *g_s, float *g_verts) {
    __shared__ float red[13];
    const int b = blockIdx.y, n = blockIdx.x * 256 + threadIdx.x, tid = threadIdx.x;
    if (tid < 13) red[tid] = 0.f;
    __syncthreads();
    float acc[13];
#pragma unroll
    for (int i = 0; i < 13; ++i) acc[i] = 0.f;
    if (n < N) {
        const float *Rb = R + b * 9;
        const float *v = verts + ((size_t)b * N + n) * 3;
        const float *g = g_out + ((size_t)b * N + n) * 3;
        const float x[3] = {v[0], v[1], v[2]}, sc = s[b];
        float u[3], gu[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            u[c] = ((x[0] * Rb[c] + x[1] * Rb[3 + c]) + x[2] * Rb[6 + c]) + t[b * 3 + c];
            gu[c] = g[c] * sc;
            acc[12] += g[c] * u[c];
            acc[9 + c] = gu[c];
        }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) acc[i * 3 + c] = x[i] * gu[c];
        if (g_verts) {
            float *gv = g_verts + ((size_t)b * N + n) * 3;
#pragma unroll
            for (int i = 0; i < 3; ++i) gv[i] = gu[0] * Rb[i * 3] + gu[1] * Rb[i * 3 + 1] + gu[2] * Rb[i * 3 + 2];
        }
    }
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        const float v = warp_sum(acc[i]);
        if ((tid & 31) == 0) atomicAdd(&red[i], v);
    }
    __syncthreads();
    if (tid < 9) atomicAdd(&g_R[b * 9 + tid], red[tid]);
    else if (tid < 12) atomicAdd(&g_t[b * 3 + tid - 9], red[tid]);
    else if (tid == 12) atomicAdd(&g_s[b], red[12]);
}

// ---- SO(3) projection (recon_fit_base.py:167-188) in fp64, one thread per matrix ---------------------
// R = U diag(1,1,det(U V^T)) V^T  ==  u1 v1^T + u2 v2^T + det(V) (u1 x u2) v3^T   (sigma_1 >= sigma_2 >= sigma_3)
struct So3 {
    double v[3][3];     // right singular vectors (rows)
    double u[3][3];     // u1, u2, u1 x u2 (rows)
    double lam[3];      // sigma_1, sigma_2, det(V) * (u3 . M v3): eigenvalues of the symmetric factor R^T M
    double sg;          // sign(det V)
    double R[3][3];
};

__device__ void so3_decompose(const float *__restrict__ mat, So3 &o) {
    double M[3][3], S[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[i][j] = (double)mat[i * 3 + j];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) S[i][j] = M[0][i] * M[0][j] + M[1][i] * M[1][j] + M[2][i] * M[2][j];
    // cyclic Jacobi on the symmetric S = M^T M
    for (int sweep = 0; sweep < 30; ++sweep) {
        const double off = fabs(S[0][1]) + fabs(S[0][2]) + fabs(S[1][2]);
        // converged to fp64 round-off relative to the spectrum (Jacobi converges quadratically: 5-6 sweeps); the result is
        // rounded to fp32 afterwards.  Sweeping on to an exact zero cost 30 sweeps = 40 us of a 480 us fit iteration.
        if (off <= 1e-17 * (fabs(S[0][0]) + fabs(S[1][1]) + fabs(S[2][2])) || off < 1e-300) break;
        for (int pq = 0; pq < 3; ++pq) {
            const int pi = pq == 2 ? 1 : 0, qi = pq == 0 ? 1 : 2;
            if (fabs(S[pi][qi]) < 1e-300) continue;
            const double theta = (S[qi][qi] - S[pi][pi]) / (2.0 * S[pi][qi]);
            const double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double cc = 1.0 / sqrt(tt * tt + 1.0), sn = tt * cc;
            for (int k = 0; k < 3; ++k) {   // S <- S J
                const double a = S[k][pi], bq = S[k][qi];
                S[k][pi] = cc * a - sn * bq; S[k][qi] = sn * a + cc * bq;
            }
            for (int k = 0; k < 3; ++k) {   // S <- J^T S
                const double a = S[pi][k], bq = S[qi][k];
                S[pi][k] = cc * a - sn * bq; S[qi][k] = sn * a + cc * bq;
            }
            for (int k = 0; k < 3; ++k) {   // V <- V J
                const double a = V[k][pi], bq = V[k][qi];
                V[k][pi] = cc * a - sn * bq; V[k][qi] = sn * a + cc * bq;
            }
        }
    }
    int idx[3] = {0, 1, 2};
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2 - i; ++j)
            if (S[idx[j]][idx[j]] < S[idx[j + 1]][idx[j + 1]]) { const int t = idx[j]; idx[j] = idx[j + 1]; idx[j + 1] = t; }
    for (int k = 0; k < 3; ++k)
        for (int i = 0; i < 3; ++i) o.v[k][i] = V[i][idx[k]];
    double Mv[3][3];
    for (int k = 0; k < 3; ++k)
        for (int i = 0; i < 3; ++i) Mv[k][i] = M[i][0] * o.v[k][0] + M[i][1] * o.v[k][1] + M[i][2] * o.v[k][2];
    for (int k = 0; k < 2; ++k) {
        double nrm = 0;
        for (int i = 0; i < 3; ++i) nrm += Mv[k][i] * Mv[k][i];
        nrm = sqrt(nrm);
        for (int i = 0; i < 3; ++i) o.u[k][i] = Mv[k][i] / (nrm > 0 ? nrm : 1.0);
    }
    // re-orthogonalise u2 against u1 (guards a nearly rank-1 input)
    const double d12 = o.u[0][0] * o.u[1][0] + o.u[0][1] * o.u[1][1] + o.u[0][2] * o.u[1][2];
    double n2 = 0;
    for (int i = 0; i < 3; ++i) { o.u[1][i] -= d12 * o.u[0][i]; n2 += o.u[1][i] * o.u[1][i]; }
    n2 = sqrt(n2);
    for (int i = 0; i < 3; ++i) o.u[1][i] /= (n2 > 0 ? n2 : 1.0);
    o.u[2][0] = o.u[0][1] * o.u[1][2] - o.u[0][2] * o.u[1][1];
    o.u[2][1] = o.u[0][2] * o.u[1][0] - o.u[0][0] * o.u[1][2];
    o.u[2][2] = o.u[0][0] * o.u[1][1] - o.u[0][1] * o.u[1][0];
    const double detV = o.v[0][0] * (o.v[1][1] * o.v[2][2] - o.v[1][2] * o.v[2][1]) -
                        o.v[0][1] * (o.v[1][0] * o.v[2][2] - o.v[1][2] * o.v[2][0]) +
                        o.v[0][2] * (o.v[1][0] * o.v[2][1] - o.v[1][1] * o.v[2][0]);
    o.sg = detV >= 0 ? 1.0 : -1.0;
    for (int k = 0; k < 3; ++k) o.lam[k] = o.u[k][0] * Mv[k][0] + o.u[k][1] * Mv[k][1] + o.u[k][2] * Mv[k][2];
    o.lam[2] *= o.sg;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            o.R[i][j] = o.u[0][i] * o.v[0][j] + o.u[1][i] * o.v[1][j] + o.sg * o.u[2][i] * o.v[2][j];
}

__global__ void project_so3_kernel(const float *__restrict__ mats, int B, float *__restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    So3 o;
    so3_decompose(mats + b * 9, o);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) out[b * 9 + i * 3 + j] = (float)o.R[i][j];
}

// adjoint of the projection.  With M = R P (P = V diag(lam) V^T symmetric), R^T dR = X is
// antisymmetric and solves X P + P X = B - B^T, B = R^T dM, hence
//   g_M = R V Y V^T,  Y_ij = (Gt_ij - Gt_ji) / (lam_i + lam_j),  Gt = V^T R^T g_R V.
__global__ void project_so3_bwd_kernel(const float *__restrict__ mats, const float *__restrict__ g_out, int B,
                                       float *__restrict__ g_mats) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    So3 o;
    so3_decompose(mats + b * 9, o);
    double RtG[3][3], Gt[3][3], Y[3][3], T1[3][3], T2[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += o.R[k][i] * (double)g_out[b * 9 + k * 3 + j];
            RtG[i][j] = s;
        }
    for (int i = 0; i < 3; ++i)       // Gt = V^T (R^T g) V ; o.v rows are the columns of V
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k)
                for (int l = 0; l < 3; ++l) s += o.v[i][k] * RtG[k][l] * o.v[j][l];
            Gt[i][j] = s;
        }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double den = o.lam[i] + o.lam[j];
            if (fabs(den) < 1e-12) den = den >= 0 ? 1e-12 : -1e-12;
            Y[i][j] = i == j ? 0.0 : (Gt[i][j] - Gt[j][i]) / den;
        }
    for (int i = 0; i < 3; ++i)       // T1 = V Y V^T
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k)
                for (int l = 0; l < 3; ++l) s += o.v[k][i] * Y[k][l] * o.v[l][j];
            T1[i][j] = s;
        }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += o.R[i][k] * T1[k][j];
            T2[i][j] = s;
        }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) g_mats[b * 9 + i * 3 + j] = (float)T2[i][j];
}

int lbs_fill(chore_handle *h, LbsParams &p, int B) {
    const LbsModel &m = h->lbs;
    if (!m.loaded) {
        chore_set_error("body model not loaded (chore_lbs_load_model)");
        return CHORE_ERR_NO_WEIGHTS;
    }
    p.V = m.V; p.J = m.J; p.nb = m.nb; p.P = (m.J - 1) * 9;
    p.v_template = m.v_template; p.shapedirs = m.shapedirs; p.posedirs = m.posedirs; p.weights = m.weights;
    p.j_template = m.j_template; p.j_shapedirs = m.j_shapedirs; p.parents = m.parents;
    p.B = B;
    // workspace: posemap, A, G, joints, gA, gpm, gbetas, gtrans
    const size_t per_b = (size_t)p.P * 2 + (size_t)p.J * 12 * 3 + (size_t)p.J * 3 + kMaxBetas + 4;
    if (int rc = chore_lbs_ws_reserve(h, per_b * B * sizeof(float))) return rc;
    float *w = static_cast<float *>(h->lbs_ws);
    p.posemap = w; w += (size_t)B * p.P;
    p.A = w; w += (size_t)B * p.J * 12;
    p.G = w; w += (size_t)B * p.J * 12;
    p.joints = w; w += (size_t)B * p.J * 3;
    p.gA = w; w += (size_t)B * p.J * 12;
    p.gpm = w; w += (size_t)B * p.P;
    p.gbetas_acc = w; w += (size_t)B * kMaxBetas;
    p.gtrans_acc = w;
    return CHORE_OK;
}

// ---------------------------------------------------------------------------------------------
// landmark regressors: Y[b, r, :] (+)= sum_k val[k] * X[b, col[k], :] over the CSR row r.
// Replaces batch_sparse_dense_matmul (lib_smpl/torch_functions.py:52-76), which the reference calls three
// times per get_landmarks() with a Python loop over the batch (lib_smpl/wrapper_pytorch.py:78-90).
// Forward: one warp per landmark (hundreds of non-zeros per row).  Adjoint: one thread per vertex over the
// transposed CSR (a handful of non-zeros per row), so no atomics and a deterministic sum order.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) spmm3_warp_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                         const float *__restrict__ val, int rows, int cols, int B,
                                                         const float *__restrict__ X, float *__restrict__ Y) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= rows * B) return;
    const int b = w / rows, r = w - b * rows;
    const float *x = X + (size_t)b * cols * 3;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int k = rowptr[r] + lane; k < rowptr[r + 1]; k += 32) {
        const float v = __ldg(val + k);
        const float *p = x + (size_t)__ldg(col + k) * 3;
        a0 = fmaf(v, __ldg(p), a0); a1 = fmaf(v, __ldg(p + 1), a1); a2 = fmaf(v, __ldg(p + 2), a2);
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane == 0) {
        float *y = Y + ((size_t)b * rows + r) * 3;
        y[0] = a0; y[1] = a1; y[2] = a2;
    }
}

__global__ void __launch_bounds__(256) spmm3_thread_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                           const float *__restrict__ val, int rows, int cols, int B,
                                                           const float *__restrict__ X, float *__restrict__ Y, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * B) return;
    const int b = i / rows, r = i - b * rows;
    const float *x = X + (size_t)b * cols * 3;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) {
        const float v = __ldg(val + k);
        const float *p = x + (size_t)__ldg(col + k) * 3;
        a0 = fmaf(v, __ldg(p), a0); a1 = fmaf(v, __ldg(p + 1), a1); a2 = fmaf(v, __ldg(p + 2), a2);
    }
    float *y = Y + ((size_t)b * rows + r) * 3;
    if (accumulate) { a0 += y[0]; a1 += y[1]; a2 += y[2]; }
    y[0] = a0; y[1] = a1; y[2] = a2;
}

}   // namespace

extern "C" int chore_lbs_load_model(chore_handle *h, const float *v_template, const float *shapedirs,
                                    const float *posedirs, const float *J_regressor, const float *weights,
                                    const int32_t *parents, int V, int J, int n_betas, int on_device) {
    CHORE_CHECK(h && v_template && shapedirs && posedirs && J_regressor && weights && parents, "null argument");
    CHORE_CHECK(V > 0 && J > 1 && J <= kMaxJ && n_betas > 0 && n_betas <= kMaxBetas, "unsupported body model V=%d J=%d betas=%d", V, J, n_betas);
    CHORE_CUDA(cudaSetDevice(h->device));
    const int P = (J - 1) * 9;
    auto to_host = [&](const void *src, size_t bytes, void *dst) -> int {
        if (on_device) CHORE_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
        else memcpy(dst, src, bytes);
        return CHORE_OK;
    };
    std::vector<float> vt((size_t)V * 3), sd((size_t)V * 3 * n_betas), pd((size_t)V * 3 * P), jr((size_t)J * V), wt((size_t)V * J);
    std::vector<int32_t> par(J);
    if (to_host(v_template, vt.size() * 4, vt.data()) || to_host(shapedirs, sd.size() * 4, sd.data()) ||
        to_host(posedirs, pd.size() * 4, pd.data()) || to_host(J_regressor, jr.size() * 4, jr.data()) ||
        to_host(weights, wt.size() * 4, wt.data()) || to_host(parents, par.size() * 4, par.data()))
        return CHORE_ERR_CUDA;
    for (int j = 1; j < J; ++j) CHORE_CHECK(par[j] >= 0 && par[j] < j, "parents[%d]=%d is not an earlier joint", j, par[j]);
    // fold the (dense, mostly zero) joint regressor into the template and the shape basis
    std::vector<float> jt((size_t)J * 3), js((size_t)J * 3 * n_betas);
    for (int j = 0; j < J; ++j) {
        std::vector<double> acc(3 + 3 * n_betas, 0.0);
        for (int v = 0; v < V; ++v) {
            const double w = jr[(size_t)j * V + v];
            if (w == 0.0) continue;
            for (int c = 0; c < 3; ++c) {
                acc[c] += w * vt[(size_t)v * 3 + c];
                for (int k = 0; k < n_betas; ++k) acc[3 + c * n_betas + k] += w * sd[((size_t)v * 3 + c) * n_betas + k];
            }
        }
        for (int c = 0; c < 3; ++c) {
            jt[j * 3 + c] = (float)acc[c];
            for (int k = 0; k < n_betas; ++k) js[((size_t)j * 3 + c) * n_betas + k] = (float)acc[3 + c * n_betas + k];
        }
    }
    LbsModel &m = h->lbs;
    auto up = [&](float **dst, const std::vector<float> &src) -> int {
        if (int rc = chore_dev_alloc(h, reinterpret_cast<void **>(dst), src.size() * 4)) return rc;
        CHORE_CUDA(cudaMemcpy(*dst, src.data(), src.size() * 4, cudaMemcpyHostToDevice));
        return CHORE_OK;
    };
    if (up(&m.v_template, vt) || up(&m.shapedirs, sd) || up(&m.posedirs, pd) || up(&m.weights, wt) ||
        up(&m.j_template, jt) || up(&m.j_shapedirs, js))
        return CHORE_ERR_CUDA;
    if (int rc = chore_dev_alloc(h, reinterpret_cast<void **>(&m.parents), J * sizeof(int32_t))) return rc;
    CHORE_CUDA(cudaMemcpy(m.parents, par.data(), J * sizeof(int32_t), cudaMemcpyHostToDevice));
    m.parents_h.assign(par.begin(), par.end());
    m.V = V; m.J = J; m.nb = n_betas;
    m.loaded = true;
    return CHORE_OK;
}

extern "C" int chore_lbs_fwd(chore_handle *h, const float *pose, const float *betas, const float *trans,
                             const float *offsets, int B, float *verts, float *jtr, float *v_posed, float *naked,
                             void *stream) {
    CHORE_CHECK(h && pose && betas && trans && verts && B > 0, "null argument");
    LbsParams p{};
    if (int rc = lbs_fill(h, p, B)) return rc;
    p.pose = pose; p.betas = betas; p.trans = trans; p.offsets = offsets;
    p.verts = verts; p.jtr = jtr; p.v_posed = v_posed; p.naked = naked;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CHORE_LAUNCH(lbs_pose_kernel, B, 64, 0, st, p);
    const size_t smem = (size_t)(p.P + p.J * 12 + kMaxBetas) * sizeof(float);
    dim3 grid((p.V + kVertsPerCta - 1) / kVertsPerCta, B);
    CHORE_LAUNCH(lbs_vertex_kernel, grid, 256, smem, st, p);
    return CHORE_OK;
}

extern "C" int chore_lbs_bwd(chore_handle *h, const float *pose, const float *betas, const float *trans,
                             const float *offsets, int B, const float *g_verts, const float *g_jtr, float *g_pose,
                             float *g_betas, float *g_trans, float *g_offsets, void *stream) {
    CHORE_CHECK(h && pose && betas && trans && g_verts && g_pose && g_betas && g_trans && B > 0, "null argument");
    LbsParams p{};
    if (int rc = lbs_fill(h, p, B)) return rc;
    p.pose = pose; p.betas = betas; p.trans = trans; p.offsets = offsets;
    p.g_verts = g_verts; p.g_jtr = g_jtr;
    p.g_pose = g_pose; p.g_betas = g_betas; p.g_trans = g_trans; p.g_offsets = g_offsets;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // recompute the body-side forward (posemap, A, G, joints), then the two adjoint kernels
    CHORE_LAUNCH(lbs_pose_kernel, B, 64, 0, st, p);
    const size_t acc_floats = (size_t)B * (p.J * 12 + p.P + kMaxBetas + 4);
    CHORE_CUDA(cudaMemsetAsync(p.gA, 0, acc_floats * sizeof(float), st));
    const size_t smem = (size_t)(2 * p.P + 2 * p.J * 12 + 2 * kMaxBetas + 4) * sizeof(float);
    dim3 grid((p.V + kVertsPerCta - 1) / kVertsPerCta, B);
    CHORE_LAUNCH(lbs_vertex_bwd_kernel, grid, 256, smem, st, p);
    CHORE_LAUNCH(lbs_pose_bwd_kernel, B, 64, 0, st, p);
    return CHORE_OK;
}

extern "C" int chore_rigid_fwd(chore_handle *h, const float *verts, const float *R, const float *t, const float *s,
                               int B, int N, float *out, void *stream) {
    CHORE_CHECK(h && verts && R && t && s && out && B > 0 && N >= 0, "null argument");
    if (N == 0) return CHORE_OK;
    dim3 grid((N + 255) / 256, B);
    CHORE_LAUNCH(rigid_fwd_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), verts, R, t, s, N, out);
    return CHORE_OK;
}

extern "C" int chore_rigid_bwd(chore_handle *h, const float *verts, const float *R, const float *t, const float *s,
                               int B, int N, const float *g_out, float *g_R, float *g_t, float *g_s, float *g_verts,
                               void *stream) {
    CHORE_CHECK(h && verts && R && t && s && g_out && g_R && g_t && g_s && B > 0 && N >= 0, "null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CHORE_CUDA(cudaMemsetAsync(g_R, 0, (size_t)B * 9 * sizeof(float), st));
    CHORE_CUDA(cudaMemsetAsync(g_t, 0, (size_t)B * 3 * sizeof(float), st));
    CHORE_CUDA(cudaMemsetAsync(g_s, 0, (size_t)B * sizeof(float), st));
    if (N == 0) return CHORE_OK;
    dim3 grid((N + 255) / 256, B);
    CHORE_LAUNCH(rigid_bwd_kernel, grid, 256, 0, st, verts, R, t, s, N, g_out, g_R, g_t, g_s, g_verts);
    return CHORE_OK;
}

extern "C" int chore_project_so3(chore_handle *h, const float *mats, int B, float *out, void *stream) {
    CHORE_CHECK(h && mats && out && B > 0, "null argument");
    CHORE_LAUNCH(project_so3_kernel, (B + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream), mats, B, out);
    return CHORE_OK;
}

extern "C" int chore_project_so3_bwd(chore_handle *h, const float *mats, const float *g_out, int B, float *g_mats,
                                     void *stream) {
    CHORE_CHECK(h && mats && g_out && g_mats && B > 0, "null argument");
    CHORE_LAUNCH(project_so3_bwd_kernel, (B + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream), mats, g_out, B, g_mats);
    return CHORE_OK;
}

// ---- landmark regressors ---------------------------------------------------------------------------
extern "C" int chore_landmarks_load(chore_handle *h, const int32_t *rowptr, const int32_t *col, const float *val,
                                    int L, int V, int nnz) {
    CHORE_CHECK(h && rowptr && col && val && L > 0 && V > 0 && nnz > 0, "bad landmark regressor (L=%d V=%d nnz=%d)", L, V, nnz);
    CHORE_CHECK(rowptr[0] == 0 && rowptr[L] == nnz, "rowptr does not span [0, nnz]");
    CHORE_CUDA(cudaSetDevice(h->device));
    for (int r = 0; r < L; ++r) CHORE_CHECK(rowptr[r] <= rowptr[r + 1], "rowptr is not monotone at row %d", r);
    for (int k = 0; k < nnz; ++k) CHORE_CHECK(col[k] >= 0 && col[k] < V, "column index %d out of range at entry %d", col[k], k);
    // transpose (counting sort by column; rows stay ascending inside a column => deterministic adjoint order)
    std::vector<int32_t> trow(V + 1, 0), tcol(nnz);
    std::vector<float> tval(nnz);
    for (int k = 0; k < nnz; ++k) ++trow[col[k] + 1];
    for (int v = 0; v < V; ++v) trow[v + 1] += trow[v];
    std::vector<int32_t> fill(trow.begin(), trow.end() - 1);
    for (int r = 0; r < L; ++r)
        for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) {
            const int d = fill[col[k]]++;
            tcol[d] = r; tval[d] = val[k];
        }
    LandmarkModel &m = h->lmk;
    auto up = [&](void **dst, const void *src, size_t bytes) -> int {
        if (int rc = chore_dev_alloc(h, dst, bytes)) return rc;
        CHORE_CUDA(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
        return CHORE_OK;
    };
    if (up(reinterpret_cast<void **>(&m.rowptr), rowptr, (size_t)(L + 1) * 4) || up(reinterpret_cast<void **>(&m.col), col, (size_t)nnz * 4) ||
        up(reinterpret_cast<void **>(&m.val), val, (size_t)nnz * 4) || up(reinterpret_cast<void **>(&m.t_rowptr), trow.data(), (size_t)(V + 1) * 4) ||
        up(reinterpret_cast<void **>(&m.t_col), tcol.data(), (size_t)nnz * 4) || up(reinterpret_cast<void **>(&m.t_val), tval.data(), (size_t)nnz * 4))
        return CHORE_ERR_CUDA;
    m.L = L; m.V = V; m.nnz = nnz; m.loaded = true;
    return CHORE_OK;
}

extern "C" int chore_landmarks_fwd(chore_handle *h, const float *verts, int B, float *out, void *stream) {
    CHORE_CHECK(h && verts && out && B > 0, "bad arguments");
    const LandmarkModel &m = h->lmk;
    if (!m.loaded) { chore_set_error("landmark regressors not loaded (chore_landmarks_load)"); return CHORE_ERR_NO_WEIGHTS; }
    const int warps = m.L * B;
    CHORE_LAUNCH(spmm3_warp_kernel, (warps * 32 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream), m.rowptr, m.col, m.val,
                 m.L, m.V, B, verts, out);
    return CHORE_OK;
}

extern "C" int chore_landmarks_bwd(chore_handle *h, const float *g_out, int B, float *g_verts, int accumulate, void *stream) {
    CHORE_CHECK(h && g_out && g_verts && B > 0, "bad arguments");
    const LandmarkModel &m = h->lmk;
    if (!m.loaded) { chore_set_error("landmark regressors not loaded (chore_landmarks_load)"); return CHORE_ERR_NO_WEIGHTS; }
    const int n = m.V * B;
    CHORE_LAUNCH(spmm3_thread_kernel, (n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream), m.t_rowptr, m.t_col, m.t_val,
                 m.V, m.L, B, g_out, g_verts, accumulate);
    return CHORE_OK;
}
