// Fused pixel-aligned point query (forward and gradient-to-points), fp32 SIMT version.
//
// One launch replaces, per query of B x N points (reference paths relative to /root/reference):
//   model/camera.py:44-88      KinectColorCamera.project_points  (~12 pointwise launches)
//   model/geometry.py:4-14     index() = grid_sample(bilinear, zeros, align_corners=True), twice
//   model/chore.py:139-143     torch.cat -> (B,323,N) (never materialised here)
//   model/chore.py:74-85,156-167  4 heads x 4 conv1d(k=1) (+ReLU)
//   model/chore.py:147-150     df[~in_img] = OUT_DIST
//
// Layout: feature maps are NHWC, so each bilinear tap is one contiguous 1 KB (256 ch) or 256 B
// (64 ch) read.  A CTA owns a tile of P points: the 323-value feature column of every point
// is staged once in shared memory (Xt[k][p]) and the four heads run as register-tiled fp32
// GEMMs against weight panels streamed with cp.async (L2-resident, 1.2 MB in total).
#include "common.cuh"
#include <cstdlib>
#include <cstring>

namespace {

constexpr int NT = 256;   // threads per CTA
constexpr int KC = 32;    // weight rows staged per cp.async stage

// model/camera.py:26-40: python doubles, cast to fp32 when they meet an fp32 tensor
constexpr float kFx = static_cast<float>(979.7844 / 2048. * 2048);
constexpr float kFy = static_cast<float>(979.840 / 2048. * 2048);
constexpr float kCx = static_cast<float>(1018.952 / 2048. * 2048);
constexpr float kCy = static_cast<float>(779.486 / 2048. * 2048);
constexpr float kHalfCrop = 600.0f;   // loadSize / 2
constexpr float kCrop = 1200.0f;
constexpr float kZ0 = 2.2f;
constexpr float kOutDist = 5.0f;      // model/chore.py:65

__host__ __device__ __forceinline__ int head_out(int h) { return h == 0 ? 2 : (h == 1 ? 9 : (h == 2 ? 14 : 6)); }

struct QueryParams {
    const float *feat, *skip;       // (B,fh,fw,256), (B,2fh,2fw,64) NHWC
    int fh, fw;
    const float *points;            // (B,N,3) or null when the grid generator is used
    const float *crop_center;       // (B,2)
    int B;
    long long N;                    // points per batch element (row length of the outputs)
    long long n_start;              // first point evaluated (grid mode), 0 otherwise
    long long n_count;              // points evaluated per batch element
    // dense-grid generator (model/sdf.py:4-27)
    int grid_mode;
    int ry, rz;
    double step[3], bmin[3];
    int batch_index;                // grid mode: which image
    unsigned head_mask;
    float *out[kNumHeads];          // df, pca, parts, centers  (B,nout,N)
    unsigned char *in_img;          // (B,N) or null
    const float *g_out[kNumHeads];  // backward only
    float *g_points;                // backward only (B,N,3)
    // weights
    const float *w1t, *w1o, *b1, *w2t, *w2o, *b2, *w3t, *w3o, *b3, *w4, *b4;
};

// ---- projection (exact fp32 op order of model/camera.py:64-65,75-78; no FMA contraction) ----
__device__ __forceinline__ void project(float x, float y, float z, float ccx, float ccy, float &nx,
                                        float &ny) {
    float px = __fadd_rn(__fdiv_rn(__fmul_rn(kFx, x), z), kCx);
    float py = __fadd_rn(__fdiv_rn(__fmul_rn(kFy, y), z), kCy);
    px = __fsub_rn(__fadd_rn(kHalfCrop, px), ccx);
    py = __fsub_rn(__fadd_rn(kHalfCrop, py), ccy);
    nx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, px), kCrop), 1.0f);
    ny = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, py), kCrop), 1.0f);
}

// bilinear taps of grid_sample(align_corners=True, zeros padding) on a (H,W) map
struct Taps {
    int x0, y0;        // north-west texel
    float wx, wy;      // distance to the west / north texel
    unsigned valid;    // bit0 nw, bit1 ne, bit2 sw, bit3 se
};
__device__ __forceinline__ Taps make_taps(float nx, float ny, int H, int W) {
    Taps t;
    float ix = __fmul_rn(__fadd_rn(nx, 1.0f), 0.5f * (float)(W - 1));
    float iy = __fmul_rn(__fadd_rn(ny, 1.0f), 0.5f * (float)(H - 1));
    t.valid = 0;
    t.x0 = 0; t.y0 = 0; t.wx = 0.f; t.wy = 0.f;
    // also rejects NaN / inf (a point on the camera plane)
    if (ix > -1.0f && ix < (float)W && iy > -1.0f && iy < (float)H) {
        float fx0 = floorf(ix), fy0 = floorf(iy);
        t.x0 = (int)fx0; t.y0 = (int)fy0;
        t.wx = ix - fx0; t.wy = iy - fy0;
        bool xl = t.x0 >= 0, xr = t.x0 + 1 < W, yt = t.y0 >= 0, yb = t.y0 + 1 < H;
        t.valid = (unsigned)(xl && yt) | ((unsigned)(xr && yt) << 1) | ((unsigned)(xl && yb) << 2) |
                  ((unsigned)(xr && yb) << 3);
    }
    return t;
}

// ---- register-tiled GEMM: acc[p][o] (+)= sum_k Xs[k][p] * W[k][o], 128 output columns ----
// thread (tx = tid&15, ty = tid>>4) owns points ty*RP..+RP and outputs tx*4..+4, 64+tx*4..+4
template <int RP, int LD>
__device__ __forceinline__ void gemm_tile(const float *__restrict__ Wg, int ldw, int K,
                                          const float *Xs, float *Ws, float (&acc)[RP][8], int tid) {
    const int tx = tid & 15, ty = tid >> 4;
    const int nchunks = (K + KC - 1) / KC;
    auto issue = [&](int chunk) {
        const int k0 = chunk * KC;
        const int rows = min(KC, K - k0);
        float *dst = Ws + (chunk & 1) * KC * 128;
#pragma unroll
        for (int i = 0; i < (KC * 32) / NT; ++i) {
            int idx = tid + i * NT;
            int row = idx >> 5, c4 = (idx & 31) * 4;
            if (row < rows) cp_async16(dst + row * 128 + c4, Wg + (size_t)(k0 + row) * ldw + c4);
        }
        cp_async_commit();
    };
    issue(0);
    for (int chunk = 0; chunk < nchunks; ++chunk) {
        if (chunk + 1 < nchunks) {
            issue(chunk + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const int k0 = chunk * KC;
        const int rows = min(KC, K - k0);
        const float *wb = Ws + (chunk & 1) * KC * 128;
        const float *xb = Xs + (size_t)k0 * LD + ty * RP;
#pragma unroll 4
        for (int kk = 0; kk < rows; ++kk) {
            const float4 w0 = *reinterpret_cast<const float4 *>(wb + kk * 128 + tx * 4);
            const float4 w1 = *reinterpret_cast<const float4 *>(wb + kk * 128 + 64 + tx * 4);
            float xv[RP];
            if constexpr (RP == 4) {
                const float4 t = *reinterpret_cast<const float4 *>(xb + kk * LD);
                xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
            } else {
                const float2 t = *reinterpret_cast<const float2 *>(xb + kk * LD);
                xv[0] = t.x; xv[1] = t.y;
            }
#pragma unroll
            for (int i = 0; i < RP; ++i) {
                acc[i][0] = fmaf(xv[i], w0.x, acc[i][0]);
                acc[i][1] = fmaf(xv[i], w0.y, acc[i][1]);
                acc[i][2] = fmaf(xv[i], w0.z, acc[i][2]);
                acc[i][3] = fmaf(xv[i], w0.w, acc[i][3]);
                acc[i][4] = fmaf(xv[i], w1.x, acc[i][4]);
                acc[i][5] = fmaf(xv[i], w1.y, acc[i][5]);
                acc[i][6] = fmaf(xv[i], w1.z, acc[i][6]);
                acc[i][7] = fmaf(xv[i], w1.w, acc[i][7]);
            }
        }
        __syncthreads();
    }
}

__device__ __forceinline__ int out_col(int tx, int j) { return j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4); }

template <int RP>
__device__ __forceinline__ void init_acc(float (&acc)[RP][8], const float *__restrict__ bias, int tx) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float b = bias ? __ldg(bias + out_col(tx, j)) : 0.f;
#pragma unroll
        for (int i = 0; i < RP; ++i) acc[i][j] = b;
    }
}

// ReLU + store to Hs[o][p]; optionally records the ReLU mask
template <int RP, int LD, bool MASK>
__device__ __forceinline__ void store_relu(const float (&acc)[RP][8], float *Hs, unsigned char *mask,
                                           int tid) {
    const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int o = out_col(tx, j);
#pragma unroll
        for (int i = 0; i < RP; ++i) {
            float v = acc[i][j];
            if (MASK) mask[o * LD + ty * RP + i] = v > 0.f;
            Hs[o * LD + ty * RP + i] = fmaxf(v, 0.f);
        }
    }
}

// store acc * mask (backward through ReLU)
template <int RP, int LD>
__device__ __forceinline__ void store_masked(const float (&acc)[RP][8], float *Hs,
                                             const unsigned char *mask, int tid) {
    const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int o = out_col(tx, j);
#pragma unroll
        for (int i = 0; i < RP; ++i)
            Hs[o * LD + ty * RP + i] = mask[o * LD + ty * RP + i] ? acc[i][j] : 0.f;
    }
}

// per-point data kept in shared memory
template <int P>
struct PointSmem {
    float x[P], y[P], z[P];
    float nx[P], ny[P];
    unsigned char in_img[P];
    unsigned char live[P];   // point index < n_count
};

template <int P>
__device__ __forceinline__ void load_points(const QueryParams &q, int b, long long n0, PointSmem<P> &ps,
                                            float *Xt, int LD, int tid) {
    if (tid < P) {
        const long long n = n0 + tid;
        const bool live = n < q.n_count;
        float x = 0.f, y = 0.f, z = 1.f;
        if (live) {
            if (q.grid_mode) {
                // model/sdf.py:4-27: coord = b_min + (b_max-b_min)/res * idx in float64, x-major
                const long long g = q.n_start + n;
                const long long iz = g % q.rz, t = g / q.rz;
                const long long iy = t % q.ry, ix = t / q.ry;
                x = (float)__dadd_rn(__dmul_rn(q.step[0], (double)ix), q.bmin[0]);
                y = (float)__dadd_rn(__dmul_rn(q.step[1], (double)iy), q.bmin[1]);
                z = (float)__dadd_rn(__dmul_rn(q.step[2], (double)iz), q.bmin[2]);
            } else {
                const float *p = q.points + ((size_t)b * q.N + q.n_start + n) * 3;
                x = p[0]; y = p[1]; z = p[2];
            }
        }
        float nx, ny;
        project(x, y, z, q.crop_center[b * 2 + 0], q.crop_center[b * 2 + 1], nx, ny);
        ps.x[tid] = x; ps.y[tid] = y; ps.z[tid] = z;
        ps.nx[tid] = nx; ps.ny[tid] = ny;
        ps.in_img[tid] = (nx >= -1.0f) && (nx <= 1.0f) && (ny >= -1.0f) && (ny <= 1.0f);
        ps.live[tid] = live;
        // z_feat = [x, y, z - 2.2] (model/chore.py:128-129), channels 256..258; 323 is padding
        Xt[256 * LD + tid] = x;
        Xt[257 * LD + tid] = y;
        Xt[258 * LD + tid] = __fsub_rn(z, kZ0);
        Xt[323 * LD + tid] = 0.f;
    }
}

// bilinear gather of every point of the tile into Xt (warp per point)
template <int P, int LD>
__device__ __forceinline__ void gather_tile(const QueryParams &q, int b, const PointSmem<P> &ps, float *Xt,
                                            int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    const float *F = q.feat + (size_t)b * q.fh * q.fw * kFeatC;
    const float *S = q.skip + (size_t)b * (2 * q.fh) * (2 * q.fw) * kSkipC;
    for (int p = warp; p < P; p += NT / 32) {
        const float nx = ps.nx[p], ny = ps.ny[p];
        {   // hourglass feature, 256 channels: lane owns c = lane*4..+3 and 128+lane*4..+3
            const Taps t = make_taps(nx, ny, q.fh, q.fw);
            const float wgt[4] = {(1.f - t.wy) * (1.f - t.wx), (1.f - t.wy) * t.wx, t.wy * (1.f - t.wx),
                                  t.wy * t.wx};
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (t.valid & (1u << k)) {
                    const float *src = F + ((size_t)(t.y0 + (k >> 1)) * q.fw + (t.x0 + (k & 1))) * kFeatC;
                    const float4 v0 = __ldg(reinterpret_cast<const float4 *>(src) + lane);
                    const float4 v1 = __ldg(reinterpret_cast<const float4 *>(src + 128) + lane);
                    a0.x = fmaf(v0.x, wgt[k], a0.x); a0.y = fmaf(v0.y, wgt[k], a0.y);
                    a0.z = fmaf(v0.z, wgt[k], a0.z); a0.w = fmaf(v0.w, wgt[k], a0.w);
                    a1.x = fmaf(v1.x, wgt[k], a1.x); a1.y = fmaf(v1.y, wgt[k], a1.y);
                    a1.z = fmaf(v1.z, wgt[k], a1.z); a1.w = fmaf(v1.w, wgt[k], a1.w);
                }
            }
            float *d0 = Xt + (lane * 4) * LD + p;
            d0[0] = a0.x; d0[LD] = a0.y; d0[2 * LD] = a0.z; d0[3 * LD] = a0.w;
            float *d1 = Xt + (128 + lane * 4) * LD + p;
            d1[0] = a1.x; d1[LD] = a1.y; d1[2 * LD] = a1.z; d1[3 * LD] = a1.w;
        }
        {   // stem skip feature, 64 channels at twice the resolution: lane owns c = lane*2, +1
            const Taps t = make_taps(nx, ny, 2 * q.fh, 2 * q.fw);
            const float wgt[4] = {(1.f - t.wy) * (1.f - t.wx), (1.f - t.wy) * t.wx, t.wy * (1.f - t.wx),
                                  t.wy * t.wx};
            float2 a = make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (t.valid & (1u << k)) {
                    const float *src =
                        S + ((size_t)(t.y0 + (k >> 1)) * (2 * q.fw) + (t.x0 + (k & 1))) * kSkipC;
                    const float2 v = __ldg(reinterpret_cast<const float2 *>(src) + lane);
                    a.x = fmaf(v.x, wgt[k], a.x);
                    a.y = fmaf(v.y, wgt[k], a.y);
                }
            }
            float *d = Xt + (259 + lane * 2) * LD + p;
            d[0] = a.x; d[LD] = a.y;
        }
    }
}

template <int P>
constexpr size_t fwd_smem_floats() {
    return (size_t)kPointCPad * (P + 4) + 2 * kHidden * (P + 4) + 2 * KC * 128;
}

// =========================================== forward ===========================================
template <int P>
__global__ void __launch_bounds__(NT, 1) query_fwd_kernel(const QueryParams q) {
    constexpr int LD = P + 4, RP = P / 16;
    extern __shared__ __align__(16) float smem[];
    float *Xt = smem;                       // [324][LD]
    float *Ha = Xt + kPointCPad * LD;       // [128][LD]
    float *Hb = Ha + kHidden * LD;          // [128][LD]
    float *Ws = Hb + kHidden * LD;          // [2][KC][128]
    __shared__ PointSmem<P> ps;

    const int tid = threadIdx.x, tx = tid & 15;
    const int b = q.grid_mode ? q.batch_index : blockIdx.y;
    const long long n0 = (long long)blockIdx.x * P;

    load_points<P>(q, b, n0, ps, Xt, LD, tid);
    __syncthreads();
    gather_tile<P, LD>(q, b, ps, Xt, tid);
    __syncthreads();

#pragma unroll 1
    for (int h = 0; h < kNumHeads; ++h) {
        if (!(q.head_mask & (1u << h))) continue;
        float acc[RP][8];
        init_acc<RP>(acc, q.b1 + h * kHidden, tx);
        gemm_tile<RP, LD>(q.w1t + h * kHidden, 4 * kHidden, kPointCPad, Xt, Ws, acc, tid);
        store_relu<RP, LD, false>(acc, Ha, nullptr, tid);
        init_acc<RP>(acc, q.b2 + h * kHidden, tx);
        gemm_tile<RP, LD>(q.w2t + (size_t)h * kHidden * kHidden, kHidden, kHidden, Ha, Ws, acc, tid);
        store_relu<RP, LD, false>(acc, Hb, nullptr, tid);
        init_acc<RP>(acc, q.b3 + h * kHidden, tx);
        gemm_tile<RP, LD>(q.w3t + (size_t)h * kHidden * kHidden, kHidden, kHidden, Hb, Ws, acc, tid);
        store_relu<RP, LD, false>(acc, Ha, nullptr, tid);
        __syncthreads();
        // last layer: n_out <= 14 outputs per point
        const int nout = head_out(h);
        float *outp = q.out[h];
        for (int idx = tid; idx < nout * P; idx += NT) {
            const int o = idx / P, p = idx - o * P;
            const float *w = q.w4 + ((size_t)h * 16 + o) * kHidden;
            float s = __ldg(q.b4 + h * 16 + o);
#pragma unroll 8
            for (int k = 0; k < kHidden; ++k) s = fmaf(Ha[k * LD + p], __ldg(w + k), s);
            if (h == 0 && !ps.in_img[p]) s = kOutDist;   // model/chore.py:147-150
            if (ps.live[p]) outp[((size_t)b * nout + o) * q.N + q.n_start + n0 + p] = s;
        }
        // the next head's first epilogue write to Ha happens after several barriers
    }
    if (q.in_img && tid < P && ps.live[tid]) q.in_img[(size_t)b * q.N + q.n_start + n0 + tid] = ps.in_img[tid];
}

// =========================================== backward ==========================================
template <int P>
constexpr size_t bwd_smem_bytes() {
    return ((size_t)2 * kPointCPad * (P + 4) + 2 * kHidden * (P + 4) + 2 * KC * 128 + 16 * P) * sizeof(float) +
           (size_t)3 * kHidden * (P + 4);
}

// d(sum of bilinear samples weighted by g)/d(ix, iy) for one map; lanes split the channels.
// Mirrors grid_sampler_2d_backward (zeros padding: out-of-range taps contribute nothing).
template <int C, int LD>
__device__ __forceinline__ void bilinear_grad(const float *__restrict__ map, int H, int W, float nx, float ny,
                                              const float *gX /* &gX[c0][p] */, int lane, float &gnx,
                                              float &gny) {
    const Taps t = make_taps(nx, ny, H, W);
    float gix = 0.f, giy = 0.f;
    if (t.valid) {
        constexpr int V = C / 32;   // channels per lane: 8 (feat) or 2 (skip)
        const float e = 1.f - t.wx, w = t.wx, s = 1.f - t.wy, n = t.wy;
#pragma unroll
        for (int part = 0; part < (V >= 4 ? V / 4 : 1); ++part) {
            constexpr int VW = V >= 4 ? 4 : V;
            const int c0 = (V >= 4 ? part * 128 + lane * 4 : lane * VW);
            float val[4][VW];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int j = 0; j < VW; ++j) val[k][j] = 0.f;
                if (t.valid & (1u << k)) {
                    const float *src = map + ((size_t)(t.y0 + (k >> 1)) * W + (t.x0 + (k & 1))) * C + c0;
                    if constexpr (VW == 4) {
                        const float4 v = __ldg(reinterpret_cast<const float4 *>(src));
                        val[k][0] = v.x; val[k][1] = v.y; val[k][2] = v.z; val[k][3] = v.w;
                    } else {
                        const float2 v = __ldg(reinterpret_cast<const float2 *>(src));
                        val[k][0] = v.x; val[k][1] = v.y;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < VW; ++j) {
                const float g = gX[(size_t)(c0 + j) * LD];
                gix += g * (s * (val[1][j] - val[0][j]) + n * (val[3][j] - val[2][j]));
                giy += g * (e * (val[2][j] - val[0][j]) + w * (val[3][j] - val[1][j]));
            }
        }
    }
    gix = warp_sum(gix);
    giy = warp_sum(giy);
    gnx += gix * (0.5f * (float)(W - 1));
    gny += giy * (0.5f * (float)(H - 1));
}

template <int P>
__global__ void __launch_bounds__(NT, 1) query_bwd_kernel(const QueryParams q) {
    constexpr int LD = P + 4, RP = P / 16;
    extern __shared__ __align__(16) float smem[];
    float *Xt = smem;                        // [324][LD]
    float *gX = Xt + kPointCPad * LD;        // [324][LD]
    float *Ha = gX + kPointCPad * LD;        // [128][LD]
    float *Hb = Ha + kHidden * LD;           // [128][LD]
    float *Ws = Hb + kHidden * LD;           // [2][KC][128]
    float *go = Ws + 2 * KC * 128;           // [16][P]
    unsigned char *mask = reinterpret_cast<unsigned char *>(go + 16 * P);   // [3][128][LD]
    __shared__ PointSmem<P> ps;

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int b = blockIdx.y;
    const long long n0 = (long long)blockIdx.x * P;

    load_points<P>(q, b, n0, ps, Xt, LD, tid);
    for (int i = tid; i < kPointCPad * LD; i += NT) gX[i] = 0.f;
    __syncthreads();
    gather_tile<P, LD>(q, b, ps, Xt, tid);
    __syncthreads();

#pragma unroll 1
    for (int h = 0; h < kNumHeads; ++h) {
        const float *gh = q.g_out[h];
        if (gh == nullptr) continue;
        const int nout = head_out(h);
        unsigned char *m1 = mask, *m2 = mask + kHidden * LD, *m3 = mask + 2 * kHidden * LD;
        float acc[RP][8];
        // ---- forward recompute, keeping only the ReLU masks ----
        init_acc<RP>(acc, q.b1 + h * kHidden, tx);
        gemm_tile<RP, LD>(q.w1t + h * kHidden, 4 * kHidden, kPointCPad, Xt, Ws, acc, tid);
        store_relu<RP, LD, true>(acc, Ha, m1, tid);
        init_acc<RP>(acc, q.b2 + h * kHidden, tx);
        gemm_tile<RP, LD>(q.w2t + (size_t)h * kHidden * kHidden, kHidden, kHidden, Ha, Ws, acc, tid);
        store_relu<RP, LD, true>(acc, Hb, m2, tid);
        init_acc<RP>(acc, q.b3 + h * kHidden, tx);
        gemm_tile<RP, LD>(q.w3t + (size_t)h * kHidden * kHidden, kHidden, kHidden, Hb, Ws, acc, tid);
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int i = 0; i < RP; ++i) m3[out_col(tx, j) * LD + ty * RP + i] = acc[i][j] > 0.f;
        // ---- upstream gradient of this head ----
        for (int idx = tid; idx < nout * P; idx += NT) {
            const int o = idx / P, p = idx - o * P;
            float g = 0.f;
            if (ps.live[p] && !(h == 0 && !ps.in_img[p]))
                g = gh[((size_t)b * nout + o) * q.N + q.n_start + n0 + p];
            go[o * P + p] = g;
        }
        __syncthreads();
        // gH3 = W4^T g  (.) mask3 -> Hb
        for (int idx = tid; idx < kHidden * P; idx += NT) {
            const int k = idx / P, p = idx - k * P;
            float s = 0.f;
            for (int o = 0; o < nout; ++o) s = fmaf(__ldg(q.w4 + ((size_t)h * 16 + o) * kHidden + k), go[o * P + p], s);
            Hb[k * LD + p] = m3[k * LD + p] ? s : 0.f;
        }
        __syncthreads();
        // gH2 = W3^T gH3 (.) mask2 -> Ha
        init_acc<RP>(acc, nullptr, tx);
        gemm_tile<RP, LD>(q.w3o + (size_t)h * kHidden * kHidden, kHidden, kHidden, Hb, Ws, acc, tid);
        store_masked<RP, LD>(acc, Ha, m2, tid);
        // gH1 = W2^T gH2 (.) mask1 -> Hb
        init_acc<RP>(acc, nullptr, tx);
        gemm_tile<RP, LD>(q.w2o + (size_t)h * kHidden * kHidden, kHidden, kHidden, Ha, Ws, acc, tid);
        store_masked<RP, LD>(acc, Hb, m1, tid);
        // gX += W1^T gH1, 324 columns in three 128-wide tiles
#pragma unroll 1
        for (int t = 0; t < 3; ++t) {
            init_acc<RP>(acc, nullptr, tx);
            gemm_tile<RP, LD>(q.w1o + (size_t)h * kHidden * kW1oLd + t * 128, kW1oLd, kHidden, Hb, Ws, acc, tid);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = t * 128 + out_col(tx, j);
                if (c < kPointCPad) {
#pragma unroll
                    for (int i = 0; i < RP; ++i) gX[c * LD + ty * RP + i] += acc[i][j];
                }
            }
        }
        __syncthreads();
    }
    __syncthreads();

    // ---- features -> image coordinates -> 3D point ----
    const int warp = tid >> 5, lane = tid & 31;
    const float *F = q.feat + (size_t)b * q.fh * q.fw * kFeatC;
    const float *S = q.skip + (size_t)b * (2 * q.fh) * (2 * q.fw) * kSkipC;
    for (int p = warp; p < P; p += NT / 32) {
        if (!ps.live[p]) continue;
        float gnx = 0.f, gny = 0.f;
        bilinear_grad<kFeatC, LD>(F, q.fh, q.fw, ps.nx[p], ps.ny[p], gX + p, lane, gnx, gny);
        bilinear_grad<kSkipC, LD>(S, 2 * q.fh, 2 * q.fw, ps.nx[p], ps.ny[p], gX + 259 * LD + p, lane, gnx, gny);
        if (lane == 0) {
            const float x = ps.x[p], y = ps.y[p], z = ps.z[p];
            // nx = (2*px')/1200 - 1 ; px' = 600 + px - cc ; px = (fx*x)/z + cx
            const float gpx = (gnx / kCrop) * 2.0f, gpy = (gny / kCrop) * 2.0f;
            const float ux = kFx * x, uy = kFy * y;
            float gx = kFx * (gpx / z), gy = kFy * (gpy / z);
            float gz = -gpx * ((ux / z) / z) - gpy * ((uy / z) / z);
            gx += gX[256 * LD + p];
            gy += gX[257 * LD + p];
            gz += gX[258 * LD + p];
            float *d = q.g_points + ((size_t)b * q.N + q.n_start + n0 + p) * 3;
            d[0] = gx; d[1] = gy; d[2] = gz;
        }
    }
}

int check_maps(const chore_handle *h, const float *feat, const float *skip, int fh, int fw) {
    CHORE_CHECK(h != nullptr, "null handle");
    if (!h->mlp.loaded) {
        chore_set_error("decoder weights not loaded (chore_load_weights)");
        return CHORE_ERR_NO_WEIGHTS;
    }
    CHORE_CHECK(feat && skip && fh > 1 && fw > 1, "bad feature maps (%p, %p, %d x %d)", (const void *)feat,
                (const void *)skip, fh, fw);
    return CHORE_OK;
}

void fill_weights(QueryParams &q, const MlpWeights &m) {
    q.w1t = m.w1t; q.w1o = m.w1o; q.b1 = m.b1;
    q.w2t = m.w2t; q.w2o = m.w2o; q.b2 = m.b2;
    q.w3t = m.w3t; q.w3o = m.w3o; q.b3 = m.b3;
    q.w4 = m.w4; q.b4 = m.b4;
}

template <int P>
int launch_fwd(const QueryParams &q, int grid_y, cudaStream_t st) {
    constexpr size_t smem = fwd_smem_floats<P>() * sizeof(float);
    CHORE_ONCE_PER_DEVICE(cudaFuncSetAttribute(query_fwd_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long tiles = (q.n_count + P - 1) / P;
    CHORE_CHECK(tiles < (1ll << 31), "too many points per launch");
    dim3 grid((unsigned)tiles, (unsigned)grid_y);
    CHORE_LAUNCH(query_fwd_kernel<P>, grid, NT, smem, st, q);
    return CHORE_OK;
}

}   // namespace

// ---------------------------------------------------------------------------------------------
// weight repacking (reference keys: {df,pca_predictor,part_predictor,center_predictor}.{0,2,4,6})
// ---------------------------------------------------------------------------------------------
static int fetch_host(const chore_tensor_desc *t, std::vector<float> &dst) {
    size_t n = 1;
    for (int i = 0; i < t->ndim; ++i) n *= (size_t)t->shape[i];
    dst.resize(n);
    if (t->on_device)
        CHORE_CUDA(cudaMemcpy(dst.data(), t->data, n * sizeof(float), cudaMemcpyDeviceToHost));
    else
        memcpy(dst.data(), t->data, n * sizeof(float));
    return CHORE_OK;
}

static int upload(chore_handle *h, float **dst, const std::vector<float> &src) {
    int rc = chore_dev_alloc(h, reinterpret_cast<void **>(dst), src.size() * sizeof(float));
    if (rc) return rc;
    CHORE_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(float), cudaMemcpyHostToDevice));
    return CHORE_OK;
}

int query_load_weights(chore_handle *h, const std::map<std::string, const chore_tensor_desc *> &t) {
    static const char *head_names[kNumHeads] = {"df", "pca_predictor", "part_predictor", "center_predictor"};
    for (int hd = 0; hd < kNumHeads; ++hd)
        for (int li = 0; li < 8; li += 2) {
            std::string base = std::string(head_names[hd]) + "." + std::to_string(li);
            if (!t.count(base + ".weight") || !t.count(base + ".bias")) return CHORE_OK;   // decoder absent
        }
    std::vector<float> w1t((size_t)kPointCPad * 512, 0.f), w1o((size_t)512 * kW1oLd, 0.f), b1(512, 0.f);
    std::vector<float> w2t((size_t)4 * 128 * 128), w2o(w2t.size()), b2(512), w3t(w2t.size()), w3o(w2t.size()), b3(512);
    std::vector<float> w4((size_t)4 * 16 * 128, 0.f), b4(64, 0.f), tmp;
    std::vector<float> raw1((size_t)4 * 128 * kPointC), raw2((size_t)4 * 128 * 128), raw3((size_t)4 * 128 * 128);
    for (int hd = 0; hd < kNumHeads; ++hd) {
        const std::string hn = head_names[hd];
        const chore_tensor_desc *d = t.at(hn + ".0.weight");
        CHORE_CHECK(d->ndim >= 2 && d->shape[0] == kHidden && d->shape[1] == kPointC, "%s.0.weight: bad shape", hn.c_str());
        if (int rc = fetch_host(d, tmp)) return rc;
        for (int o = 0; o < kHidden; ++o)
            for (int k = 0; k < kPointC; ++k) {
                const float v = tmp[(size_t)o * kPointC + k];
                raw1[((size_t)hd * kHidden + o) * kPointC + k] = v;
                w1t[(size_t)k * 512 + hd * kHidden + o] = v;
                w1o[(size_t)(hd * kHidden + o) * kW1oLd + k] = v;
            }
        if (int rc = fetch_host(t.at(hn + ".0.bias"), tmp)) return rc;
        for (int o = 0; o < kHidden; ++o) b1[hd * kHidden + o] = tmp[o];
        for (int layer = 0; layer < 2; ++layer) {
            const std::string ln = hn + (layer == 0 ? ".2" : ".4");
            d = t.at(ln + ".weight");
            CHORE_CHECK(d->shape[0] == kHidden && d->shape[1] == kHidden, "%s.weight: bad shape", ln.c_str());
            if (int rc = fetch_host(d, tmp)) return rc;
            std::vector<float> &wt = layer == 0 ? w2t : w3t, &wo = layer == 0 ? w2o : w3o;
            for (int o = 0; o < kHidden; ++o)
                for (int k = 0; k < kHidden; ++k) {
                    const float v = tmp[(size_t)o * kHidden + k];
                    (layer == 0 ? raw2 : raw3)[((size_t)hd * kHidden + o) * kHidden + k] = v;
                    wt[((size_t)hd * kHidden + k) * kHidden + o] = v;
                    wo[((size_t)hd * kHidden + o) * kHidden + k] = v;
                }
            if (int rc = fetch_host(t.at(ln + ".bias"), tmp)) return rc;
            std::vector<float> &bb = layer == 0 ? b2 : b3;
            for (int o = 0; o < kHidden; ++o) bb[hd * kHidden + o] = tmp[o];
        }
        d = t.at(hn + ".6.weight");
        CHORE_CHECK(d->shape[0] == kHeadOut[hd] && d->shape[1] == kHidden, "%s.6.weight: bad shape", hn.c_str());
        if (int rc = fetch_host(d, tmp)) return rc;
        for (int o = 0; o < kHeadOut[hd]; ++o)
            for (int k = 0; k < kHidden; ++k) w4[((size_t)hd * 16 + o) * kHidden + k] = tmp[(size_t)o * kHidden + k];
        if (int rc = fetch_host(t.at(hn + ".6.bias"), tmp)) return rc;
        for (int o = 0; o < kHeadOut[hd]; ++o) b4[hd * 16 + o] = tmp[o];
    }
    MlpWeights &m = h->mlp;
    int rc = 0;
    rc |= upload(h, &m.w1t, w1t); rc |= upload(h, &m.w1o, w1o); rc |= upload(h, &m.b1, b1);
    rc |= upload(h, &m.w2t, w2t); rc |= upload(h, &m.w2o, w2o); rc |= upload(h, &m.b2, b2);
    rc |= upload(h, &m.w3t, w3t); rc |= upload(h, &m.w3o, w3o); rc |= upload(h, &m.b3, b3);
    rc |= upload(h, &m.w4, w4); rc |= upload(h, &m.b4, b4);
    if (rc) return CHORE_ERR_CUDA;
    if (int rc2 = query_tc_pack_weights(h, raw1, raw2, raw3, w4)) return rc2;
    if (int rc2 = query_g_pack_weights(h, raw1)) return rc2;
    m.loaded = true;
    return CHORE_OK;
}

bool query_use_tensor_cores() {
    static const bool simt = [] {
        const char *e = getenv("CHORE_B200_QUERY");
        return e != nullptr && strcmp(e, "simt") == 0;
    }();
    return !simt;
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" int chore_query_fwd(chore_handle *h, const float *feat, const float *skip, int fh, int fw,
                               const float *points, const float *crop_center, int B, int N, uint32_t head_mask,
                               float *df, float *pca, float *parts, float *centers, uint8_t *in_img,
                               void *stream) {
    if (int rc = check_maps(h, feat, skip, fh, fw)) return rc;
    CHORE_CHECK(points && crop_center && B > 0 && N >= 0, "bad points / crop_center / sizes");
    head_mask &= CHORE_HEAD_ALL;
    CHORE_CHECK(head_mask != 0, "head_mask selects no head");
    float *outs[kNumHeads] = {df, pca, parts, centers};
    for (int i = 0; i < kNumHeads; ++i)
        CHORE_CHECK(!(head_mask & (1u << i)) || outs[i], "output %d requested by head_mask is NULL", i);
    if (N == 0) return CHORE_OK;
    QueryParams q{};
    q.feat = feat; q.skip = skip; q.fh = fh; q.fw = fw;
    q.points = points; q.crop_center = crop_center;
    q.B = B; q.N = N; q.n_start = 0; q.n_count = N; q.grid_mode = 0;
    q.head_mask = head_mask;
    for (int i = 0; i < kNumHeads; ++i) q.out[i] = outs[i];
    q.in_img = in_img;
    fill_weights(q, h->mlp);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (query_use_tensor_cores() && head_mask == CHORE_HEAD_ALL && query_tc2_enabled())     // all heads: CTA-pair kernel
        return query_tc2_launch(h, feat, skip, fh, fw, points, crop_center, B, N, 0, N, 0, 0, nullptr, nullptr, nullptr, outs, in_img, st);
    if (query_use_tensor_cores())
        return query_tc_launch(h, feat, skip, fh, fw, points, crop_center, B, N, 0, N, 0, 0, nullptr, nullptr, nullptr,
                               head_mask, outs, in_img, st);
    // small problems: 32-point tiles fill more SMs
    const long long tiles64 = ((long long)N + 63) / 64 * B;
    if (tiles64 < 2ll * h->sm_count) return launch_fwd<32>(q, B, st);
    return launch_fwd<64>(q, B, st);
}

extern "C" int chore_query_grid(chore_handle *h, const float *feat, const float *skip, int fh, int fw,
                                const float *crop_center, int b, const int res[3], const double b_min[3],
                                const double b_max[3], int64_t start, int64_t count, uint32_t head_mask,
                                float *df, float *pca, float *parts, float *centers, void *stream) {
    if (int rc = check_maps(h, feat, skip, fh, fw)) return rc;
    CHORE_CHECK(crop_center && res && b_min && b_max && b >= 0, "bad grid arguments");
    const int64_t total = (int64_t)res[0] * res[1] * res[2];
    CHORE_CHECK(res[0] > 0 && res[1] > 0 && res[2] > 0 && start >= 0 && count >= 0 && start + count <= total,
                "grid range [%lld, %lld) outside %lld points", (long long)start, (long long)(start + count),
                (long long)total);
    head_mask &= CHORE_HEAD_ALL;
    CHORE_CHECK(head_mask != 0, "head_mask selects no head");
    float *outs[kNumHeads] = {df, pca, parts, centers};
    for (int i = 0; i < kNumHeads; ++i)
        CHORE_CHECK(!(head_mask & (1u << i)) || outs[i], "output %d requested by head_mask is NULL", i);
    if (count == 0) return CHORE_OK;
    QueryParams q{};
    q.feat = feat; q.skip = skip; q.fh = fh; q.fw = fw;
    q.points = nullptr; q.crop_center = crop_center;
    q.B = 1; q.N = total; q.n_start = start; q.n_count = count;
    q.grid_mode = 1; q.ry = res[1]; q.rz = res[2]; q.batch_index = b;
    for (int i = 0; i < 3; ++i) {
        q.step[i] = (b_max[i] - b_min[i]) / (double)res[i];
        q.bmin[i] = b_min[i];
    }
    q.head_mask = head_mask;
    // outputs are (nout, total) rows of image b; the kernel indexes them with b = batch_index,
    // so rebase the pointers to make "b * nout * N" vanish
    for (int i = 0; i < kNumHeads; ++i) q.out[i] = outs[i] ? outs[i] - (size_t)b * kHeadOut[i] * total : nullptr;
    q.in_img = nullptr;
    fill_weights(q, h->mlp);
    if (query_use_tensor_cores()) {
        if (query_g_mode() == 1)
            return query_g_launch(h, feat, skip, fh, fw, crop_center, total, start, count, b, res, q.step, q.bmin, head_mask, q.out,
                                  static_cast<cudaStream_t>(stream));
    }
    if (query_use_tensor_cores() && head_mask == CHORE_HEAD_ALL && query_tc2_enabled())
        return query_tc2_launch(h, feat, skip, fh, fw, nullptr, crop_center, 1, total, start, count, 1, b, res, q.step, q.bmin, q.out, nullptr,
                                static_cast<cudaStream_t>(stream));
    if (query_use_tensor_cores())
        return query_tc_launch(h, feat, skip, fh, fw, nullptr, crop_center, 1, total, start, count, 1, b, res, q.step, q.bmin,
                               head_mask, q.out, nullptr, static_cast<cudaStream_t>(stream));
    return launch_fwd<64>(q, 1, static_cast<cudaStream_t>(stream));
}

extern "C" size_t chore_query_bwd_workspace_bytes(int B, int N) {
    return query_use_tensor_cores() ? (size_t)B * (size_t)N * 384 * sizeof(float) : 0;
}

extern "C" int chore_query_bwd_ws(chore_handle *h, const float *feat, const float *skip, int fh, int fw,
                                  const float *points, const float *crop_center, int B, int N, const float *g_df,
                                  const float *g_pca, const float *g_parts, const float *g_centers, float *g_points,
                                  void *workspace, size_t workspace_bytes, void *stream) {
    if (int rc = check_maps(h, feat, skip, fh, fw)) return rc;
    CHORE_CHECK(points && crop_center && g_points && B > 0 && N >= 0, "bad points / crop_center / g_points");
    if (N == 0) return CHORE_OK;
    if (query_use_tensor_cores() && getenv("CHORE_B200_QUERY_BWD_SIMT") == nullptr) {
        const float *const gh[4] = {g_df, g_pca, g_parts, g_centers};
        return query_bwd_tc_launch(h, feat, skip, fh, fw, points, crop_center, B, N, gh, g_points, workspace, workspace_bytes,
                                   static_cast<cudaStream_t>(stream));
    }
    constexpr int P = 32;
    QueryParams q{};
    q.feat = feat; q.skip = skip; q.fh = fh; q.fw = fw;
    q.points = points; q.crop_center = crop_center;
    q.B = B; q.N = N; q.n_start = 0; q.n_count = N; q.grid_mode = 0;
    q.g_out[0] = g_df; q.g_out[1] = g_pca; q.g_out[2] = g_parts; q.g_out[3] = g_centers;
    q.g_points = g_points;
    fill_weights(q, h->mlp);
    constexpr size_t smem = bwd_smem_bytes<P>();
    CHORE_ONCE_PER_DEVICE(cudaFuncSetAttribute(query_bwd_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((N + P - 1) / P), (unsigned)B);
    CHORE_LAUNCH(query_bwd_kernel<P>, grid, NT, smem, static_cast<cudaStream_t>(stream), q);
    return CHORE_OK;
}

extern "C" int chore_query_bwd(chore_handle *h, const float *feat, const float *skip, int fh, int fw,
                               const float *points, const float *crop_center, int B, int N, const float *g_df,
                               const float *g_pca, const float *g_parts, const float *g_centers,
                               float *g_points, void *stream) {
    // handle-owned scratch (grows on demand; use chore_query_bwd_ws under CUDA-graph capture)
    return chore_query_bwd_ws(h, feat, skip, fh, fw, points, crop_center, B, N, g_df, g_pca, g_parts, g_centers, g_points,
                              nullptr, 0, stream);
}
