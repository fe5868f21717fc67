// Stacked-hourglass encoder, second generation: the graph walk of HGFilter.forward (model/HGFilters.py:144-185),
// HourGlass._forward (:26-50) and ConvBlock.forward (model/net_util.py:374-396) over conv_hx_kernel (conv_hx.cu).
//
// Differences to the first version (encoder.cu + conv_tc.cu, 457 launches per image):
//   * every producer accumulates the GroupNorm statistics of what it writes (conv epilogue, avg-pool, upsample+add,
//     stem) -- there is no statistics pass (135 launches gone);
//   * the GroupNorm affine + ReLU + fp16 hi/lo split happen in the consumer's prologue in shared memory -- there are no
//     fp16 activation planes in HBM and no act_split launches (149 gone);
//   * the ~175 remaining launches of one image are captured once per (buffers, shape) in a CUDA graph and replayed.
#include "common.cuh"

#include <cstring>
#include <stdexcept>
#include <vector>

namespace {

constexpr int kGroups = 32;

// ------------------------------------------------------------------------------------------
// block-level GroupNorm statistics of element-wise producers: every thread owns a fixed float4 channel
// chunk (the grid stride is a multiple of the channel count), accumulates locally, then shared fp32
// atomics per block and one fp64 global atomic per (group, moment) and block.
// ------------------------------------------------------------------------------------------
struct EwStats {
    float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
    __device__ __forceinline__ void add(const float4 &v, int cpg) {
        if (cpg >= 4) {
            s0 += (v.x + v.y) + (v.z + v.w);
            q0 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        } else {   // cpg == 2: channels (c, c+1) and (c+2, c+3) are two groups
            s0 += v.x + v.y; q0 += v.x * v.x + v.y * v.y;
            s1 += v.z + v.w; q1 += v.z * v.z + v.w * v.w;
        }
    }
    // acc: shared float[64]; c = first channel of this thread's chunk
    __device__ __forceinline__ void flush(float *acc, int c, int cpg, double *gst, int b) {
        const int g = c / cpg;
        atomicAdd(&acc[g * 2], s0); atomicAdd(&acc[g * 2 + 1], q0);
        if (cpg < 4) { atomicAdd(&acc[g * 2 + 2], s1); atomicAdd(&acc[g * 2 + 3], q1); }
        __syncthreads();
        if (threadIdx.x < kGroups * 2 && acc[threadIdx.x] != 0.f)
            atomicAdd(&gst[(size_t)b * kGroups * 2 + threadIdx.x], (double)acc[threadIdx.x]);
    }
};

__device__ __forceinline__ void gn_affine_hx(const double *__restrict__ sums, int b, int c, int cpg, double inv_n,
                                             const float *__restrict__ gamma, const float *__restrict__ beta, float &scale,
                                             float &shift) {
    const int g = c / cpg;
    const double mean = __ldcg(sums + (size_t)b * kGroups * 2 + g * 2) * inv_n;
    double var = __ldcg(sums + (size_t)b * kGroups * 2 + g * 2 + 1) * inv_n - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const float rstd = (float)(1.0 / sqrt(var + 1e-5));
    scale = __ldg(gamma + c) * rstd;
    shift = __ldg(beta + c) - (float)mean * scale;
}

// out = relu(gn(in)) materialised (tmpx is an output of the encoder) + statistics of out
__global__ void __launch_bounds__(256) gn_apply_stats_kernel(const float *__restrict__ in, float *__restrict__ out, int C, int HW,
                                                             const double *__restrict__ sums, const float *__restrict__ gamma,
                                                             const float *__restrict__ beta, double *__restrict__ st_out) {
    chore_pdl_launch_dependents();
    chore_pdl_wait();
    __shared__ float acc[kGroups * 2];
    const int b = blockIdx.y, tid = threadIdx.x, c4n = C / 4, cpg = C / kGroups;
    if (tid < kGroups * 2) acc[tid] = 0.f;
    __syncthreads();
    const size_t n4 = (size_t)HW * c4n;
    const int c = (int)((blockIdx.x * 256 + tid) % c4n) * 4;
    const double inv_n = 1.0 / ((double)HW * cpg);
    float sc[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) gn_affine_hx(sums, b, c + j, cpg, inv_n, gamma, beta, sc[j], sh[j]);
    EwStats st;
    const float4 *src = reinterpret_cast<const float4 *>(in) + (size_t)b * n4;
    float4 *dst = reinterpret_cast<float4 *>(out) + (size_t)b * n4;
    for (size_t i = (size_t)blockIdx.x * 256 + tid; i < n4; i += (size_t)gridDim.x * 256) {
        float4 v = __ldcg(src + i);
        v.x = fmaxf(fmaf(v.x, sc[0], sh[0]), 0.f); v.y = fmaxf(fmaf(v.y, sc[1], sh[1]), 0.f);
        v.z = fmaxf(fmaf(v.z, sc[2], sh[2]), 0.f); v.w = fmaxf(fmaf(v.w, sc[3], sh[3]), 0.f);
        dst[i] = v;
        st.add(v, cpg);
    }
    st.flush(acc, c, cpg, st_out, b);
}

// avg-pool 2x2 stride 2 (F.avg_pool2d, model/HGFilters.py:32,152) + statistics of the result
__global__ void __launch_bounds__(256) avgpool_stats_kernel(const float *__restrict__ in, float *__restrict__ out, int H, int W, int C,
                                                            double *__restrict__ st_out) {
    chore_pdl_launch_dependents();
    chore_pdl_wait();
    __shared__ float acc[kGroups * 2];
    const int b = blockIdx.y, tid = threadIdx.x, c4n = C / 4, cpg = C / kGroups, OW = W / 2, OH = H / 2;
    if (tid < kGroups * 2) acc[tid] = 0.f;
    __syncthreads();
    const size_t n4 = (size_t)OH * OW * c4n;
    const int c4 = (int)((blockIdx.x * 256 + tid) % c4n);
    EwStats st;
    const float4 *src = reinterpret_cast<const float4 *>(in) + (size_t)b * H * W * c4n;
    float4 *dst = reinterpret_cast<float4 *>(out) + (size_t)b * n4;
    for (size_t i = (size_t)blockIdx.x * 256 + tid; i < n4; i += (size_t)gridDim.x * 256) {
        const size_t t = i / c4n;
        const int ox = (int)(t % OW), oy = (int)(t / OW);
        const float4 *p = src + ((size_t)(oy * 2) * W + ox * 2) * c4n + c4;
        const float4 v00 = __ldcg(p), v01 = __ldcg(p + c4n), v10 = __ldcg(p + (size_t)W * c4n), v11 = __ldcg(p + (size_t)W * c4n + c4n);
        float4 o;
        o.x = (((v00.x + v01.x) + v10.x) + v11.x) / 4.f;
        o.y = (((v00.y + v01.y) + v10.y) + v11.y) / 4.f;
        o.z = (((v00.z + v01.z) + v10.z) + v11.z) / 4.f;
        o.w = (((v00.w + v01.w) + v10.w) + v11.w) / 4.f;
        dst[i] = o;
        st.add(o, cpg);
    }
    if (st_out) st.flush(acc, c4 * 4, cpg, st_out, b);
}

// up = up + bicubic_x2(low), align_corners=True, A = -0.75, clamped taps (F.interpolate(..., mode='bicubic',
// align_corners=True) + add, model/HGFilters.py:47-49) + statistics of the result
__device__ __forceinline__ void cubic_coeffs_hx(float t, float (&c)[4]) {
    const float A = -0.75f;
    const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
    c[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
    c[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
    c[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
    c[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

// One thread = one channel quad of a 2 x 2 block of outputs.  The four outputs read 4 x 4 taps each out of a shared 5 x 5
// window of the low-resolution map (scale (I - 1) / (O - 1) < 1/2: neighbouring outputs move by at most one input pixel), so
// the window is streamed row by row -- 25 loads instead of 64 -- and every output accumulates its taps in exactly the order
// of the one-output-per-thread formulation (rows top to bottom, columns left to right, fmaf chains from 0): same bits.
__global__ void __launch_bounds__(256) upadd_stats_kernel(const float *__restrict__ low, float *__restrict__ up, int IH, int IW, int C,
                                                          double *__restrict__ st_out) {
    chore_pdl_launch_dependents();
    chore_pdl_wait();
    __shared__ float acc[kGroups * 2];
    const int b = blockIdx.y, tid = threadIdx.x, c4n = C / 4, cpg = C / kGroups, OW = IW * 2, OH = IH * 2;
    if (tid < kGroups * 2) acc[tid] = 0.f;
    __syncthreads();
    const size_t nblk = (size_t)IH * IW * c4n;             // 2 x 2 output blocks x channel quads
    const int c4 = (int)((blockIdx.x * 256 + tid) % c4n);
    const float sy = OH > 1 ? (float)(IH - 1) / (float)(OH - 1) : 0.f;
    const float sx = OW > 1 ? (float)(IW - 1) / (float)(OW - 1) : 0.f;
    EwStats st;
    const float4 *lb = reinterpret_cast<const float4 *>(low) + (size_t)b * IH * IW * c4n + c4;
    float4 *ub = reinterpret_cast<float4 *>(up) + (size_t)b * OH * OW * c4n + c4;
    for (size_t i = (size_t)blockIdx.x * 256 + tid; i < nblk; i += (size_t)gridDim.x * 256) {
        const size_t t = i / c4n;
        const int ox = 2 * (int)(t % IW), oy = 2 * (int)(t / IW);
        int iy[2], ix[2];
        float cy[2][4], cx[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float ry = sy * (oy + r), rx = sx * (ox + r);
            iy[r] = (int)floorf(ry); ix[r] = (int)floorf(rx);
            cubic_coeffs_hx(ry - iy[r], cy[r]);
            cubic_coeffs_hx(rx - ix[r], cx[r]);
        }
        const bool dy = iy[1] != iy[0], dx = ix[1] != ix[0];          // the second output row / column starts one input pixel later
        float4 o[2][2];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c) o[r][c] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 5; ++k) {                                  // window row iy[0] - 1 + k
            if (k == 4 && !dy) break;                                  // nobody reads the fifth row
            const int yy = min(max(iy[0] - 1 + k, 0), IH - 1);
            float4 v[5];
#pragma unroll
            for (int m = 0; m < 5; ++m) {
                const int xx = min(max(ix[0] - 1 + m, 0), IW - 1);
                v[m] = (m < 4 || dx) ? __ldg(lb + ((size_t)yy * IW + xx) * c4n) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float4 rs[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const bool sh = c == 1 && dx;
                float4 rsum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const float4 a = sh ? v[m + 1] : v[m];
                    rsum.x = fmaf(a.x, cx[c][m], rsum.x); rsum.y = fmaf(a.y, cx[c][m], rsum.y);
                    rsum.z = fmaf(a.z, cx[c][m], rsum.z); rsum.w = fmaf(a.w, cx[c][m], rsum.w);
                }
                rs[c] = rsum;
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const bool sh = r == 1 && dy;
                // tap index of this window row for output row r: k (not shifted, k <= 3) or k - 1 (shifted, k >= 1)
                if (sh ? k == 0 : k == 4) continue;
                const float w = sh ? cy[r][k > 0 ? k - 1 : 0] : cy[r][k < 4 ? k : 3];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    o[r][c].x = fmaf(rs[c].x, w, o[r][c].x); o[r][c].y = fmaf(rs[c].y, w, o[r][c].y);
                    o[r][c].z = fmaf(rs[c].z, w, o[r][c].z); o[r][c].w = fmaf(rs[c].w, w, o[r][c].w);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float4 *dst = ub + ((size_t)(oy + r) * OW + (ox + c)) * c4n;
                float4 u = __ldcg(dst);
                u.x += o[r][c].x; u.y += o[r][c].y; u.z += o[r][c].z; u.w += o[r][c].w;
                *dst = u;
                st.add(u, cpg);
            }
    }
    if (st_out) st.flush(acc, c4 * 4, cpg, st_out, b);
}

// ------------------------------------------------------------------------------------------
// stem: conv 7x7 stride 2 pad 3, 5 -> 64, + bias; NCHW in, NHWC out (model/HGFilters.py:149) + statistics
// ------------------------------------------------------------------------------------------
constexpr int kStemTile = 16, kStemPatch = kStemTile * 2 + 5;   // 37
constexpr int kStemPatchFloats = (CHORE_IN_CH * kStemPatch * (kStemPatch + 1) + 3) / 4 * 4;

__global__ void __launch_bounds__(256) stem_hx_kernel(const float *__restrict__ img, int H, int W, const float *__restrict__ w /*[5*49][64]*/,
                                                      const float *__restrict__ bias, float *__restrict__ out, double *__restrict__ st_out) {
    chore_pdl_launch_dependents();
    extern __shared__ __align__(16) float smem[];
    __shared__ double acc[kGroups * 2];
    float *patch = smem;                                           // [5][37][38]
    float *ws = smem + kStemPatchFloats;                            // [245][64]
    const int OH = H / 2, OW = W / 2;
    const int b = blockIdx.z, oy0 = blockIdx.y * kStemTile, ox0 = blockIdx.x * kStemTile;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < kGroups * 2) acc[tid] = 0.0;
    for (int i = tid; i < CHORE_IN_CH * 49 * 64 / 4; i += 256)
        reinterpret_cast<float4 *>(ws)[i] = __ldg(reinterpret_cast<const float4 *>(w) + i);
    const int iy0 = oy0 * 2 - 3, ix0 = ox0 * 2 - 3;
    for (int i = tid; i < CHORE_IN_CH * kStemPatch * kStemPatch; i += 256) {
        const int c = i / (kStemPatch * kStemPatch), r = i % (kStemPatch * kStemPatch);
        const int py = r / kStemPatch, px = r % kStemPatch;
        const int gy = iy0 + py, gx = ix0 + px;
        float v = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(img + (((size_t)b * CHORE_IN_CH + c) * H + gy) * W + gx);
        patch[(c * kStemPatch + py) * (kStemPatch + 1) + px] = v;
    }
    __syncthreads();
    const int ty = tid / kStemTile, tx = tid % kStemTile;
    const int oy = oy0 + ty, ox = ox0 + tx;
    const bool valid = oy < OH && ox < OW;
    chore_pdl_wait();                  // `out` / the statistics slot may still be in use by the previous kernel of the stream
#pragma unroll 1
    for (int pass = 0; pass < 4; ++pass) {
        float a[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = __ldg(bias + pass * 16 + j);
        for (int c = 0; c < CHORE_IN_CH; ++c)
            for (int ky = 0; ky < 7; ++ky) {
                const float *prow = patch + (c * kStemPatch + ty * 2 + ky) * (kStemPatch + 1) + tx * 2;
                const float *wrow = ws + ((c * 7 + ky) * 7) * 64 + pass * 16;
#pragma unroll
                for (int kx = 0; kx < 7; ++kx) {
                    const float x = prow[kx];
                    const float4 *wv = reinterpret_cast<const float4 *>(wrow + kx * 64);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 q = wv[j];
                        a[j * 4 + 0] = fmaf(x, q.x, a[j * 4 + 0]);
                        a[j * 4 + 1] = fmaf(x, q.y, a[j * 4 + 1]);
                        a[j * 4 + 2] = fmaf(x, q.z, a[j * 4 + 2]);
                        a[j * 4 + 3] = fmaf(x, q.w, a[j * 4 + 3]);
                    }
                }
            }
        if (valid) {
            float4 *dst = reinterpret_cast<float4 *>(out + (((size_t)b * OH + oy) * OW + ox) * 64 + pass * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_float4(a[j * 4], a[j * 4 + 1], a[j * 4 + 2], a[j * 4 + 3]);
        }
        // GroupNorm(32, 64): 2 channels per group -> 8 groups per pass
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            float s = valid ? a[2 * g] + a[2 * g + 1] : 0.f;
            float q = valid ? a[2 * g] * a[2 * g] + a[2 * g + 1] * a[2 * g + 1] : 0.f;
            s = warp_sum(s); q = warp_sum(q);
            if (lane == g) { atomicAdd(&acc[(pass * 8 + g) * 2], (double)s); atomicAdd(&acc[(pass * 8 + g) * 2 + 1], (double)q); }
        }
    }
    __syncthreads();
    if (tid < kGroups * 2) atomicAdd(&st_out[(size_t)b * kGroups * 2 + tid], acc[tid]);
}

// ------------------------------------------------------------------------------------------
// host-side graph walk
// ------------------------------------------------------------------------------------------
struct ActH {
    float *p = nullptr;
    int C = 0, H = 0, W = 0;
    double *st = nullptr;     // [B][32][2] sums of this tensor (filled by its producers), or null
};

struct CtxH {
    chore_handle *h;
    cudaStream_t st;
    int B;
    bool dry;
    char *base = nullptr;
    size_t top = 0, peak = 0;
    char *zbase = nullptr;       // arena that is zeroed at the start of every encode: statistics slots, K-split counters
    size_t ztop = 0;
    int rc = 0;
    // fork / join of the hourglass branches (see hourglass()): side streams and events, owned by the EncoderPlan
    cudaStream_t side[2] = {nullptr, nullptr};
    cudaEvent_t *events = nullptr;
    int n_events = 0, ev_next = 0;
    int hold = 0;                // > 0: temporaries stay allocated (their kernels run concurrently with what is recorded next)

    float *alloc(size_t floats) {
        const size_t bytes = (floats * sizeof(float) + 255) / 256 * 256;
        float *p = reinterpret_cast<float *>(base + top);
        top += bytes;
        if (top > peak) peak = top;
        return p;
    }
    void *zalloc(size_t bytes) {
        void *p = zbase ? zbase + ztop : reinterpret_cast<void *>(16);
        ztop += (bytes + 15) / 16 * 16;
        return p;
    }
    double *slot() { return static_cast<double *>(zalloc((size_t)B * kGroups * 2 * sizeof(double))); }
    ActH act(int C, int H, int W, bool stats) {
        ActH a;
        a.C = C; a.H = H; a.W = W;
        a.p = alloc((size_t)B * H * W * C);
        if (stats) a.st = slot();
        return a;
    }
};

#define HX_LAUNCH(ctx, kernel, grid, block, smem, ...)                                       \
    do {                                                                                     \
        if (!(ctx).dry && (ctx).rc == 0) {                                                   \
            cudaError_t e_ = chore_launch_pdl(kernel, dim3(grid), dim3(block), (smem), (ctx).st, __VA_ARGS__); \
            g_launch_count.fetch_add(1, std::memory_order_relaxed);                          \
            if (e_ != cudaSuccess) {                                                         \
                chore_set_error("%s:%d: launch of %s -> %s", __FILE__, __LINE__, #kernel, cudaGetErrorString(e_)); \
                (ctx).rc = CHORE_ERR_CUDA;                                                   \
            }                                                                                \
        }                                                                                    \
    } while (0)

const ConvW &cw(CtxH &c, const std::string &k) { return c.h->enc.conv.at(k); }
const NormW &nw(CtxH &c, const std::string &k) { return c.h->enc.norm.at(k); }

int ew_grid(CtxH &c, size_t n4) {
    const size_t blocks = (n4 + 255) / 256;
    const size_t cap = (size_t)c.h->sm_count * 2;
    return (int)(blocks < cap ? blocks : cap);
}

struct ConvSpec {
    const ActH *in = nullptr;
    const NormW *gn = nullptr;           // GroupNorm + ReLU on the input (needs in->st), or null
    float *out = nullptr; int ld_out = 0, off_out = 0;
    const float *res = nullptr; int ld_res = 0, off_res = 0;
    ActH *raw = nullptr;                 // optional raw copy (with its statistics slot)
    double *st_out = nullptr; int c_out_total = 0;   // statistics of out: groups of c_out_total / 32 channels
};

void conv(CtxH &c, const ConvW &w, const ConvSpec &s) {
    if (c.rc) return;
    ConvHxArgs a{};
    a.in = s.in->p; a.ld_in = s.in->C; a.off_in = 0; a.Cin = w.cin;
    a.B = c.B; a.H = s.in->H; a.W = s.in->W; a.KS = w.kh; a.N = w.cout;
    a.w = w.whx; a.bias = w.bias;
    if (s.gn) { a.gn_in = s.in->st; a.gamma = s.gn->gamma; a.beta = s.gn->beta; a.relu = 1; }
    a.out = s.out; a.ld_out = s.ld_out; a.off_out = s.off_out;
    a.res = s.res; a.ld_res = s.ld_res; a.off_res = s.off_res;
    if (s.raw) {
        a.raw = s.raw->p; a.ld_raw = s.raw->C; a.off_raw = 0;
        a.st_raw = s.raw->st; a.cpg_raw = s.raw->C / kGroups;
    }
    if (s.st_out) { a.st_out = s.st_out; a.cpg_out = s.c_out_total / kGroups; }
    if (!w.whx || (s.gn && !s.in->st)) {
        chore_set_error("encoder (hx): layer is missing packed weights or input statistics");
        c.rc = CHORE_ERR_INVALID;
        return;
    }
    ConvHxPlan pl{};
    if ((c.rc = conv_hx_plan(c.h, a, &pl)) != CHORE_OK) return;
    const size_t mark = c.top;
    float *part = pl.part_floats ? c.alloc(pl.part_floats) : nullptr;
    int *cnt = pl.counters ? static_cast<int *>(c.zalloc((size_t)pl.counters * sizeof(int))) : nullptr;
    c.top = mark;                        // scratch of this launch only (stream order makes reuse safe)
    if (c.dry) return;
    c.rc = conv_hx_launch(c.h, a, pl, part, cnt, c.st);
    g_launch_count.fetch_add(0, std::memory_order_relaxed);
}

// ConvBlock.forward (model/net_util.py:374-396).  `out` may be preallocated (C = cout).
ActH conv_block(CtxH &c, const std::string &p, const ActH &x, int cout, bool want_stats, ActH out = ActH()) {
    const int cin = x.C, H = x.H, W = x.W;
    if (!out.p) out = c.act(cout, H, W, false);
    out.st = want_stats ? c.slot() : nullptr;
    const size_t mark = c.top;
    ActH t1 = c.act(cout / 2, H, W, true), t2 = c.act(cout / 4, H, W, true);
    const float *res = x.p;
    int ld_res = cin;
    if (cin != cout) {   // downsample = Sequential(bn4, ReLU, conv1x1): the residual lives in `out`
        ConvSpec s;
        s.in = &x; s.gn = &nw(c, p + ".bn4");
        s.out = out.p; s.ld_out = cout;
        conv(c, cw(c, p + ".downsample.2"), s);
        res = out.p;
        ld_res = cout;
    }
    {   // conv1: x -> [0, cout/2)
        ConvSpec s;
        s.in = &x; s.gn = &nw(c, p + ".bn1");
        s.out = out.p; s.ld_out = cout; s.off_out = 0;
        s.res = res; s.ld_res = ld_res; s.off_res = 0;
        s.raw = &t1;
        s.st_out = out.st; s.c_out_total = cout;
        conv(c, cw(c, p + ".conv1"), s);
    }
    {   // conv2: o1 -> [cout/2, 3cout/4)
        ConvSpec s;
        s.in = &t1; s.gn = &nw(c, p + ".bn2");
        s.out = out.p; s.ld_out = cout; s.off_out = cout / 2;
        s.res = res; s.ld_res = ld_res; s.off_res = cout / 2;
        s.raw = &t2;
        s.st_out = out.st; s.c_out_total = cout;
        conv(c, cw(c, p + ".conv2"), s);
    }
    {   // conv3: o2 -> [3cout/4, cout)
        ConvSpec s;
        s.in = &t2; s.gn = &nw(c, p + ".bn3");
        s.out = out.p; s.ld_out = cout; s.off_out = 3 * cout / 4;
        s.res = res; s.ld_res = ld_res; s.off_res = 3 * cout / 4;
        s.st_out = out.st; s.c_out_total = cout;
        conv(c, cw(c, p + ".conv3"), s);
    }
    if (c.hold == 0) c.top = mark;   // t1, t2 are dead (stream order makes reuse safe; a forked branch keeps them until the join)
    return out;
}

ActH avgpool(CtxH &c, const ActH &x, bool want_stats, float *dst = nullptr) {
    ActH o;
    if (dst) { o.p = dst; o.C = x.C; o.H = x.H / 2; o.W = x.W / 2; if (want_stats) o.st = c.slot(); }
    else o = c.act(x.C, x.H / 2, x.W / 2, want_stats);
    const size_t n4 = (size_t)o.H * o.W * o.C / 4;
    HX_LAUNCH(c, avgpool_stats_kernel, dim3(ew_grid(c, n4), c.B), 256, 0, x.p, o.p, x.H, x.W, x.C, o.st);
    return o;
}

// HourGlass._forward (model/HGFilters.py:26-50); the result carries statistics (its consumer normalises it)
// The skip branch (up1 = ConvBlock at this resolution) and the low-resolution path (pool, ConvBlocks / inner hourglass) are
// independent until the upsample + add.  The low-resolution convolutions occupy 48 - 96 SMs for a few microseconds each, so the
// skip branch is recorded on a side stream (an edge-free pair of chains in the captured graph) and joined before the add:
// its four launches then run beside the low path instead of in front of it.
static int fork_levels() {
    static const int n = [] {
        const char *e = getenv("CHORE_B200_HX_FORK");
        return e != nullptr ? atoi(e) : 2;
    }();
    return n;
}

ActH hourglass(CtxH &c, const std::string &p, int level, const ActH &x) {
    const std::string L = std::to_string(level);
    const bool fork = level <= fork_levels() && level <= 2 && (c.dry || (c.side[level - 1] != nullptr && c.ev_next + 2 <= c.n_events));
    ActH up1;
    size_t mark;
    cudaEvent_t ev_join = nullptr;
    if (fork) {
        up1 = c.act(x.C, x.H, x.W, false);
        mark = c.top;                                    // everything above up1 is released at the end, the side branch's temporaries too
        cudaStream_t main_st = c.st;
        if (!c.dry && c.rc == 0) {
            cudaEvent_t ev_fork = c.events[c.ev_next++];
            ev_join = c.events[c.ev_next++];
            if (cudaEventRecord(ev_fork, main_st) != cudaSuccess || cudaStreamWaitEvent(c.side[level - 1], ev_fork, 0) != cudaSuccess) c.rc = CHORE_ERR_CUDA;
            c.st = c.side[level - 1];
        }
        ++c.hold;
        conv_block(c, p + ".b1_" + L, x, x.C, false, up1);
        --c.hold;
        if (!c.dry) {
            if (c.rc == 0 && cudaEventRecord(ev_join, c.st) != cudaSuccess) c.rc = CHORE_ERR_CUDA;
            c.st = main_st;
        }
    } else {
        up1 = conv_block(c, p + ".b1_" + L, x, x.C, false);
        mark = c.top;
    }
    ActH low1 = conv_block(c, p + ".b2_" + L, avgpool(c, x, true), x.C, true);
    ActH low2 = level > 1 ? hourglass(c, p, level - 1, low1) : conv_block(c, p + ".b2_plus_" + L, low1, x.C, true);
    ActH low3 = conv_block(c, p + ".b3_" + L, low2, x.C, false);
    up1.st = c.slot();
    if (fork && !c.dry && c.rc == 0 && cudaStreamWaitEvent(c.st, ev_join, 0) != cudaSuccess) c.rc = CHORE_ERR_CUDA;
    const size_t n4 = (size_t)low3.H * low3.W * low3.C / 4;          // one thread per 2 x 2 output block and channel quad
    HX_LAUNCH(c, upadd_stats_kernel, dim3(ew_grid(c, n4), c.B), 256, 0, low3.p, up1.p, low3.H, low3.W, low3.C, up1.st);
    c.top = mark;
    return up1;
}

constexpr int kNumStack = 5, kDepth = 2;

// im2col of the 7x7 stride-2 pad-3 stem (model/HGFilters.py:149): NCHW image -> (B, H/2 * W/2, 256) columns, k = (c, ky, kx)
// (245 values, zero padded), so that the stem runs as a 1x1 convolution on the tensor cores (conv_hx_kernel<1>, K = 256)
// with bias and GroupNorm statistics in its epilogue.  One float4 of columns per thread; the image stays in L1 / L2.
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float *__restrict__ img, int H, int W, float *__restrict__ col, size_t total4) {
    chore_pdl_launch_dependents();
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total4) return;
    const int OH = H / 2, OW = W / 2;
    const int q = (int)(i & 63);
    const size_t px = i >> 6;
    const int ox = (int)(px % OW);
    const size_t t = px / OW;
    const int oy = (int)(t % OH), b = (int)(t / OH);
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k = q * 4 + j;
        const int c = k / 49, r = k - c * 49, ky = r / 7, kx = r - ky * 7;
        const int iy = 2 * oy + ky - 3, ix = 2 * ox + kx - 3;
        const bool ok = k < CHORE_IN_CH * 49 && iy >= 0 && iy < H && ix >= 0 && ix < W;
        v[j] = ok ? __ldg(img + (((size_t)b * CHORE_IN_CH + c) * H + iy) * W + ix) : 0.f;
    }
    reinterpret_cast<float4 *>(col)[i] = make_float4(v[0], v[1], v[2], v[3]);
}

static bool stem_on_tensor_cores() {
    static const bool on = [] {
        const char *e = getenv("CHORE_B200_STEM");
        return !(e != nullptr && strcmp(e, "simt") == 0);
    }();
    return on;
}

void run_graph(CtxH &c, const float *images, int H, int W, float *feat, float *skip, float *normx) {
    const std::string p = "image_filter";
    const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4;
    ActH s0 = c.act(64, H2, W2, true);
    if (stem_on_tensor_cores() && cw(c, p + ".conv1").whx != nullptr) {
        const size_t mark = c.top;
        ActH col = c.act(256, H2, W2, false);            // 67 MB per 512^2 image, released right after the convolution
        const size_t total4 = (size_t)c.B * H2 * W2 * 64;
        HX_LAUNCH(c, stem_im2col_kernel, (unsigned)((total4 + 255) / 256), 256, 0, images, H, W, col.p, total4);
        ConvW w = cw(c, p + ".conv1");
        w.kh = w.kw = 1; w.cin = 256;
        ConvSpec s;
        s.in = &col;
        s.out = s0.p; s.ld_out = 64;
        s.st_out = s0.st; s.c_out_total = 64;
        conv(c, w, s);
        c.top = mark;
    } else {
        dim3 grid((W2 + kStemTile - 1) / kStemTile, (H2 + kStemTile - 1) / kStemTile, c.B);
        const size_t smem = (size_t)(kStemPatchFloats + CHORE_IN_CH * 49 * 64) * sizeof(float);
        const ConvW &w = cw(c, p + ".conv1");
        HX_LAUNCH(c, stem_hx_kernel, grid, 256, smem, images, H, W, w.w, w.bias, s0.p, s0.st);
    }
    ActH tmpx;
    tmpx.p = skip; tmpx.C = 64; tmpx.H = H2; tmpx.W = W2; tmpx.st = c.slot();
    {
        const size_t n4 = (size_t)H2 * W2 * 64 / 4;
        const NormW &n = nw(c, p + ".bn1");
        HX_LAUNCH(c, gn_apply_stats_kernel, dim3(ew_grid(c, n4), c.B), 256, 0, s0.p, skip, 64, H2 * W2, s0.st, n.gamma, n.beta, tmpx.st);
    }
    ActH x = conv_block(c, p + ".conv2", tmpx, 128, false);
    ActH nx = avgpool(c, x, true, normx);
    x = conv_block(c, p + ".conv3", nx, 128, true);
    ActH previous = conv_block(c, p + ".conv4", x, 256, true);
    for (int i = 0; i < kNumStack; ++i) {
        const std::string si = std::to_string(i);
        const size_t mark = c.top;
        ActH hg = hourglass(c, p + ".m" + si, kDepth, previous);
        ActH ll = conv_block(c, p + ".top_m_" + si, hg, 256, false);
        ActH ll2 = c.act(256, H4, W4, true);
        {   // conv_last (1x1 + bias) on the raw block output; its GroupNorm+ReLU (bn_end) is applied by the consumers
            ConvSpec s;
            s.in = &ll;
            s.out = ll2.p; s.ld_out = 256;
            s.st_out = ll2.st; s.c_out_total = 256;
            conv(c, cw(c, p + ".conv_last" + si), s);
        }
        const NormW &ne = nw(c, p + ".bn_end" + si);
        const bool last = i == kNumStack - 1;
        ActH out;
        out.C = 256; out.H = H4; out.W = W4;
        out.p = last ? feat : c.alloc((size_t)c.B * H4 * W4 * 256);
        {   // l_i: the stack output
            ConvSpec s;
            s.in = &ll2; s.gn = &ne;
            s.out = out.p; s.ld_out = 256;
            conv(c, cw(c, p + ".l" + si), s);
        }
        if (!last) {   // previous = previous + bl(ll) + al(out)   (model/HGFilters.py:180-183)
            ConvSpec s;
            s.in = &ll2; s.gn = &ne;
            s.out = previous.p; s.ld_out = 256;
            s.res = previous.p; s.ld_res = 256;
            conv(c, cw(c, p + ".bl" + si), s);
            previous.st = c.slot();
            ConvSpec s2;
            s2.in = &out;
            s2.out = previous.p; s2.ld_out = 256;
            s2.res = previous.p; s2.ld_res = 256;
            s2.st_out = previous.st; s2.c_out_total = 256;
            conv(c, cw(c, p + ".al" + si), s2);
        }
        c.top = mark;
    }
}

}   // namespace

// ---------------------------------------------------------------------------------------------
// cached CUDA graphs of one encode call
// ---------------------------------------------------------------------------------------------
struct EncoderPlan {
    struct Entry {
        const void *images; void *feat, *skip, *normx, *ws, *ws2;
        int B, H, W;
        cudaGraphExec_t exec;
        uint64_t kernels;
    };
    std::vector<Entry> entries;
    // callers whose buffers come from a cycling allocator would make every call a new address combination (= a new capture of
    // ~2 ms); after kMaxSpecific address-specific graphs per shape the encode runs a generic graph on handle-owned staging
    // buffers and copies the image in and the maps out (42 MB of device-to-device copies per 512 x 512 image, ~1 % of the encode)
    static constexpr int kMaxSpecific = 4;
    float *gen_buf = nullptr;
    size_t gen_floats = 0;
    cudaStream_t cap_stream = nullptr;
    cudaStream_t side[2] = {nullptr, nullptr};      // hourglass skip branches (fork / join)
    std::vector<cudaEvent_t> events;
    bool stem_configured = false;
};

void encoder_plan_destroy(chore_handle *h) {
    if (!h->enc_plan) return;
    for (auto &e : h->enc_plan->entries) cudaGraphExecDestroy(e.exec);
    if (h->enc_plan->cap_stream) cudaStreamDestroy(h->enc_plan->cap_stream);
    if (h->enc_plan->gen_buf) cudaFree(h->enc_plan->gen_buf);
    for (cudaStream_t s : h->enc_plan->side) if (s) cudaStreamDestroy(s);
    for (cudaEvent_t e : h->enc_plan->events) cudaEventDestroy(e);
    delete h->enc_plan;
    h->enc_plan = nullptr;
}

static bool graphs_enabled() {
    static const bool on = [] {
        const char *e = getenv("CHORE_B200_ENCODER_GRAPH");
        return !(e != nullptr && strcmp(e, "0") == 0);
    }();
    return on;
}

int encode_hx(chore_handle *h, const float *images, int B, int H, int W, float *feat, float *skip, float *normx, cudaStream_t st) {
    if (!h->enc_plan) h->enc_plan = new EncoderPlan();
    EncoderPlan &plan = *h->enc_plan;
    CtxH dry{};
    dry.h = h; dry.B = B; dry.dry = true; dry.st = nullptr;
    try {
        run_graph(dry, images, H, W, feat, skip, normx);
    } catch (const std::out_of_range &) {
        chore_set_error("encoder weights incomplete: a tensor of the reference state_dict is missing");
        return CHORE_ERR_NO_WEIGHTS;
    }
    const size_t gn_bytes = dry.ztop + 16;
    if (h->ws_bytes < dry.peak + 256 || h->ws2_bytes < gn_bytes) {
        // the arenas move: every captured graph holds stale pointers
        for (auto &e : plan.entries) cudaGraphExecDestroy(e.exec);
        plan.entries.clear();
    }
    if (int rc = chore_ws_reserve(h, dry.peak + 256)) return rc;
    if (int rc = chore_ws2_reserve(h, gn_bytes)) return rc;
    if (int rc = conv_hx_configure(h)) return rc;
    if (!plan.stem_configured) {
        const size_t smem = (size_t)(kStemPatchFloats + CHORE_IN_CH * 49 * 64) * sizeof(float);
        CHORE_CUDA(cudaFuncSetAttribute(stem_hx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        plan.stem_configured = true;
    }
    if (plan.events.empty()) {
        for (cudaStream_t &s : plan.side) CHORE_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        plan.events.resize(2 * kNumStack * kDepth);
        for (cudaEvent_t &e : plan.events) CHORE_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    auto record = [&](cudaStream_t s, const float *im, float *ft, float *sk, float *nx) -> int {
        CtxH c{};
        c.h = h; c.B = B; c.dry = false; c.st = s;
        c.base = static_cast<char *>(h->ws);
        c.zbase = static_cast<char *>(h->ws2);
        c.side[0] = plan.side[0]; c.side[1] = plan.side[1];
        c.events = plan.events.data(); c.n_events = (int)plan.events.size();
        CHORE_CUDA(cudaMemsetAsync(h->ws2, 0, gn_bytes, s));
        run_graph(c, im, H, W, ft, sk, nx);
        return c.rc;
    };
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (st != nullptr && cudaStreamIsCapturing(st, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
    if (!graphs_enabled() || cap != cudaStreamCaptureStatusNone) return record(st, images, feat, skip, normx);

    auto find = [&](const void *im, void *ft, void *sk, void *nx) -> EncoderPlan::Entry * {
        for (auto &e : plan.entries)
            if (e.images == im && e.feat == ft && e.skip == sk && e.normx == nx && e.ws == h->ws && e.ws2 == h->ws2 && e.B == B && e.H == H && e.W == W)
                return &e;
        return nullptr;
    };
    auto capture = [&](const float *im, float *ft, float *sk, float *nx, EncoderPlan::Entry **out) -> int {
        if (!plan.cap_stream) CHORE_CUDA(cudaStreamCreateWithFlags(&plan.cap_stream, cudaStreamNonBlocking));
        const uint64_t before = g_launch_count.load();
        CHORE_CUDA(cudaStreamBeginCapture(plan.cap_stream, cudaStreamCaptureModeThreadLocal));
        const int rc = record(plan.cap_stream, im, ft, sk, nx);
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(plan.cap_stream, &graph);
        const uint64_t kernels = g_launch_count.load() - before;
        g_launch_count.fetch_sub(kernels, std::memory_order_relaxed);   // recorded, not executed
        if (rc != CHORE_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess) {
            chore_set_error("encoder graph capture failed: %s", cudaGetErrorString(ce));
            return CHORE_ERR_CUDA;
        }
        cudaGraphExec_t exec = nullptr;
        const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) {
            chore_set_error("encoder graph instantiation failed: %s", cudaGetErrorString(ie));
            return CHORE_ERR_CUDA;
        }
        if (plan.entries.size() >= 16) { cudaGraphExecDestroy(plan.entries.front().exec); plan.entries.erase(plan.entries.begin()); }
        plan.entries.push_back({im, ft, sk, nx, h->ws, h->ws2, B, H, W, exec, kernels});
        *out = &plan.entries.back();
        return CHORE_OK;
    };
    EncoderPlan::Entry *e = find(images, feat, skip, normx);
    if (e == nullptr) {
        int specific = 0;
        for (auto &x : plan.entries)
            if (x.B == B && x.H == H && x.W == W && x.ws == h->ws && x.ws2 == h->ws2 && x.images != plan.gen_buf) ++specific;
        if (specific < EncoderPlan::kMaxSpecific) {
            if (int rc = capture(images, feat, skip, normx, &e)) return rc;
        } else {
            // generic graph on the staging buffers: [image | feat | skip | normx]
            const size_t n_img = (size_t)B * CHORE_IN_CH * H * W, n_feat = (size_t)B * (H / 4) * (W / 4) * 256,
                         n_skip = (size_t)B * (H / 2) * (W / 2) * 64, n_nx = (size_t)B * (H / 4) * (W / 4) * 128;
            const size_t need = n_img + n_feat + n_skip + n_nx;
            if (plan.gen_floats < need) {
                for (size_t i = 0; i < plan.entries.size();)      // graphs on the old staging buffer are stale
                    if (plan.entries[i].images == plan.gen_buf && plan.gen_buf != nullptr) { cudaGraphExecDestroy(plan.entries[i].exec); plan.entries.erase(plan.entries.begin() + i); }
                    else ++i;
                if (plan.gen_buf) CHORE_CUDA(cudaFree(plan.gen_buf));
                plan.gen_buf = nullptr; plan.gen_floats = 0;
                CHORE_CUDA(cudaMalloc(&plan.gen_buf, need * sizeof(float)));
                plan.gen_floats = need;
            }
            float *g_img = plan.gen_buf, *g_feat = g_img + n_img, *g_skip = g_feat + n_feat, *g_nx = g_skip + n_skip;
            e = find(g_img, g_feat, g_skip, g_nx);
            if (e == nullptr)
                if (int rc = capture(g_img, g_feat, g_skip, g_nx, &e)) return rc;
            CHORE_CUDA(cudaMemcpyAsync(g_img, images, n_img * sizeof(float), cudaMemcpyDeviceToDevice, st));
            CHORE_CUDA(cudaGraphLaunch(e->exec, st));
            g_launch_count.fetch_add(e->kernels, std::memory_order_relaxed);
            CHORE_CUDA(cudaMemcpyAsync(feat, g_feat, n_feat * sizeof(float), cudaMemcpyDeviceToDevice, st));
            CHORE_CUDA(cudaMemcpyAsync(skip, g_skip, n_skip * sizeof(float), cudaMemcpyDeviceToDevice, st));
            if (normx) CHORE_CUDA(cudaMemcpyAsync(normx, g_nx, n_nx * sizeof(float), cudaMemcpyDeviceToDevice, st));
            return CHORE_OK;
        }
    }
    CHORE_CUDA(cudaGraphLaunch(e->exec, st));
    g_launch_count.fetch_add(e->kernels, std::memory_order_relaxed);
    return CHORE_OK;
}
