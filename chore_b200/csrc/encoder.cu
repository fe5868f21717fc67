// Stacked-hourglass image encoder (fp32 SIMT version), channels-last.
//
// Replaces HGFilter.forward (model/HGFilters.py:144-185), HourGlass._forward (:26-50) and
// ConvBlock.forward (model/net_util.py:374-396) of the reference: 150 conv2d + 137 GroupNorm(32)
// + 11 avg-pools + 10 bicubic upsamples run as separate ATen ops there.  Here
//   - activations are NHWC so a pixel's channels are contiguous (what the point query wants),
//   - GroupNorm is split into a statistics reduction and an affine+ReLU that is applied in
//     the consuming convolution's prologue (the normalised tensor is never written),
//   - channel concat and the residual add of a ConvBlock are the conv epilogue (the conv
//     writes its slice of the block output with the residual already added),
//   - the bicubic x2 upsample is fused with the hourglass skip add.
#include "common.cuh"

#include <cstring>
#include <stdexcept>

namespace {

// ------------------------------------------------------------------------------------------
// GroupNorm statistics: per (image, group) sum and sum of squares in fp64
// ------------------------------------------------------------------------------------------
constexpr int kGroups = 32;

__global__ void __launch_bounds__(256) gn_stats_kernel(const float *__restrict__ in, int ld, int off, int C,
                                                       int HW, int pix_per_cta, double *__restrict__ sums) {
    __shared__ double acc[kGroups * 2];
    const int b = blockIdx.y, tid = threadIdx.x;
    if (tid < kGroups * 2) acc[tid] = 0.0;
    __syncthreads();
    const int c4n = C / 4;                       // float4 lanes per pixel
    const int lanes = 256 / c4n;                 // pixels processed concurrently (C <= 1024)
    const int c4 = tid % c4n, pl = tid / c4n;
    const int p0 = blockIdx.x * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
    float s = 0.f, ss = 0.f;
    if (pl < lanes) {
        const float *base = in + (size_t)b * HW * ld + off + c4 * 4;
        for (int p = p0 + pl; p < p1; p += lanes) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(base + (size_t)p * ld));
            s += (v.x + v.y) + (v.z + v.w);
            ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
        const int g = (c4 * 4) / (C / kGroups);   // C/32 is a multiple of 4 or {1,2}: see host check
        atomicAdd(&acc[g * 2], (double)s);
        atomicAdd(&acc[g * 2 + 1], (double)ss);
    }
    __syncthreads();
    if (tid < kGroups * 2) atomicAdd(&sums[(size_t)b * kGroups * 2 + tid], acc[tid]);
}

// groups narrower than 4 channels (C = 64: 2 channels per group): scalar variant
__global__ void __launch_bounds__(256) gn_stats_narrow_kernel(const float *__restrict__ in, int ld, int off, int C,
                                                              int HW, int pix_per_cta, double *__restrict__ sums) {
    __shared__ double acc[kGroups * 2];
    const int b = blockIdx.y, tid = threadIdx.x;
    if (tid < kGroups * 2) acc[tid] = 0.0;
    __syncthreads();
    const int lanes = 256 / C;
    const int c = tid % C, pl = tid / C;
    const int p0 = blockIdx.x * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
    float s = 0.f, ss = 0.f;
    if (pl < lanes) {
        const float *base = in + (size_t)b * HW * ld + off + c;
        for (int p = p0 + pl; p < p1; p += lanes) {
            const float v = __ldg(base + (size_t)p * ld);
            s += v;
            ss += v * v;
        }
        const int g = c / (C / kGroups);
        atomicAdd(&acc[g * 2], (double)s);
        atomicAdd(&acc[g * 2 + 1], (double)ss);
    }
    __syncthreads();
    if (tid < kGroups * 2) atomicAdd(&sums[(size_t)b * kGroups * 2 + tid], acc[tid]);
}

// scale/shift of channel c from the group sums: y = relu(x * scale + shift)
__device__ __forceinline__ void gn_affine(const double *__restrict__ sums, int b, int c, int cpg, double inv_n,
                                          const float *__restrict__ gamma, const float *__restrict__ beta,
                                          float &scale, float &shift) {
    const int g = c / cpg;
    const double mean = sums[(size_t)b * kGroups * 2 + g * 2] * inv_n;
    double var = sums[(size_t)b * kGroups * 2 + g * 2 + 1] * inv_n - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const float rstd = (float)(1.0 / sqrt(var + 1e-5));
    scale = __ldg(gamma + c) * rstd;
    shift = __ldg(beta + c) - (float)mean * scale;
}

// materialise relu(gn(x)) (only needed for tmpx, which is an output of the encoder)
__global__ void __launch_bounds__(256) gn_relu_apply_kernel(const float *__restrict__ in, float *__restrict__ out,
                                                            int C, int HW, const double *__restrict__ sums,
                                                            const float *__restrict__ gamma,
                                                            const float *__restrict__ beta, size_t total4) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total4) return;
    const int c4n = C / 4;
    const int c = (int)(i % c4n) * 4;
    const int b = (int)(i / ((size_t)c4n * HW));
    const double inv_n = 1.0 / ((double)HW * (C / kGroups));
    float4 v = __ldg(reinterpret_cast<const float4 *>(in) + i);
    float sc, sh;
    gn_affine(sums, b, c + 0, C / kGroups, inv_n, gamma, beta, sc, sh); v.x = fmaxf(fmaf(v.x, sc, sh), 0.f);
    gn_affine(sums, b, c + 1, C / kGroups, inv_n, gamma, beta, sc, sh); v.y = fmaxf(fmaf(v.y, sc, sh), 0.f);
    gn_affine(sums, b, c + 2, C / kGroups, inv_n, gamma, beta, sc, sh); v.z = fmaxf(fmaf(v.z, sc, sh), 0.f);
    gn_affine(sums, b, c + 3, C / kGroups, inv_n, gamma, beta, sc, sh); v.w = fmaxf(fmaf(v.w, sc, sh), 0.f);
    reinterpret_cast<float4 *>(out)[i] = v;
}

// ------------------------------------------------------------------------------------------
// stem: conv 7x7 stride 2 pad 3, 5 -> 64, + bias; NCHW in, NHWC out (model/HGFilters.py:149)
// ------------------------------------------------------------------------------------------
constexpr int kStemTile = 16, kStemPatch = kStemTile * 2 + 5;   // 37
constexpr int kStemPatchFloats = (CHORE_IN_CH * kStemPatch * (kStemPatch + 1) + 3) / 4 * 4;   // keeps ws 16 B aligned

__global__ void __launch_bounds__(256) stem_conv_kernel(const float *__restrict__ img, int H, int W,
                                                        const float *__restrict__ w /*[5*49][64]*/,
                                                        const float *__restrict__ bias, float *__restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    float *patch = smem;                                           // [5][37][38]
    float *ws = smem + kStemPatchFloats;                            // [245][64]
    const int OH = H / 2, OW = W / 2;
    const int b = blockIdx.z, oy0 = blockIdx.y * kStemTile, ox0 = blockIdx.x * kStemTile;
    const int tid = threadIdx.x;
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
#pragma unroll 1
    for (int pass = 0; pass < 4; ++pass) {
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = __ldg(bias + pass * 16 + j);
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
                        acc[j * 4 + 0] = fmaf(x, q.x, acc[j * 4 + 0]);
                        acc[j * 4 + 1] = fmaf(x, q.y, acc[j * 4 + 1]);
                        acc[j * 4 + 2] = fmaf(x, q.z, acc[j * 4 + 2]);
                        acc[j * 4 + 3] = fmaf(x, q.w, acc[j * 4 + 3]);
                    }
                }
            }
        if (oy < OH && ox < OW) {
            float4 *dst = reinterpret_cast<float4 *>(out + (((size_t)b * OH + oy) * OW + ox) * 64 + pass * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_float4(acc[j * 4], acc[j * 4 + 1], acc[j * 4 + 2], acc[j * 4 + 3]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// conv KSxKS (stride 1, pad KS/2) as an implicit GEMM with a fused GroupNorm+ReLU prologue and
// a concat / residual epilogue
// ------------------------------------------------------------------------------------------
struct ConvArgs {
    const float *in; int ld_in, off_in, Cin;
    int B, H, W;
    const float *w;            // [KS*KS][Cin][Cout]
    const float *bias;         // [Cout] or null
    int Cout;
    const double *gn_sums;     // [B][32][2] of the input tensor, or null (no prologue)
    const float *gamma, *beta;
    float *out; int ld_out, off_out;          // out = conv (+bias) (+res)
    const float *res; int ld_res, off_res;    // optional residual (may alias out)
    float *raw; int ld_raw, off_raw;          // optional copy of conv (+bias) without the residual
};

constexpr int kTW = 16, kTH = 8, kCK = 16;

template <int KS, int COT>
struct ConvSmem {
    static constexpr int HW_ = kTW + KS - 1, HH = kTH + KS - 1;
    static constexpr int HLD = (HW_ + 3) / 4 * 4;   // 20 for KS=3, 16 for KS=1
    static constexpr int xs_floats = kCK * HH * HLD;
    static constexpr int ws_floats = KS * KS * kCK * COT;
    static constexpr size_t bytes = (size_t)(xs_floats + ws_floats + 512) * sizeof(float);
};

template <int KS, int COT>
__global__ void __launch_bounds__(256) conv_kernel(const ConvArgs a) {
    using S = ConvSmem<KS, COT>;
    constexpr int TCO = COT / 16, PAD = KS / 2, NX = 8 + KS - 1;
    extern __shared__ __align__(16) float smem[];
    float *xs = smem;                    // [CK][HH][HLD]
    float *ws = xs + S::xs_floats;       // [KS*KS][CK][COT]
    float *sc = ws + S::ws_floats;       // [256] GroupNorm scale per input channel
    float *sh = sc + 256;                // [256] shift

    const int tid = threadIdx.x;
    const int tiles_x = (a.W + kTW - 1) / kTW;
    const int tx0 = (blockIdx.x % tiles_x) * kTW, ty0 = (blockIdx.x / tiles_x) * kTH;
    const int co0 = blockIdx.y * COT;
    const int b = blockIdx.z;
    const int cg = tid & 15, pg = tid >> 4;
    const int r = pg >> 1, px0 = (pg & 1) * 8;
    const bool has_gn = a.gn_sums != nullptr;

    if (has_gn) {
        const double inv_n = 1.0 / ((double)a.H * a.W * (a.Cin / kGroups));
        for (int c = tid; c < a.Cin; c += 256) gn_affine(a.gn_sums, b, c, a.Cin / kGroups, inv_n, a.gamma, a.beta, sc[c], sh[c]);
    }

    float acc[8][TCO];
#pragma unroll
    for (int p = 0; p < 8; ++p)
#pragma unroll
        for (int c = 0; c < TCO; ++c) acc[p][c] = 0.f;

    const float *inb = a.in + (size_t)b * a.H * a.W * a.ld_in + a.off_in;
    for (int ci0 = 0; ci0 < a.Cin; ci0 += kCK) {
        __syncthreads();   // previous chunk fully consumed (and sc/sh visible on the first pass)
        // ---- stage the activated input halo tile, transposed to [ci][y][x] ----
        for (int idx = tid; idx < S::HH * S::HW_ * (kCK / 4); idx += 256) {
            const int c4 = idx & 3, hp = idx >> 2;
            const int hx = hp % S::HW_, hy = hp / S::HW_;
            const int gy = ty0 + hy - PAD, gx = tx0 + hx - PAD;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
                const int c = ci0 + c4 * 4;
                v = __ldg(reinterpret_cast<const float4 *>(inb + ((size_t)gy * a.W + gx) * a.ld_in + c));
                if (has_gn) {
                    v.x = fmaxf(fmaf(v.x, sc[c + 0], sh[c + 0]), 0.f);
                    v.y = fmaxf(fmaf(v.y, sc[c + 1], sh[c + 1]), 0.f);
                    v.z = fmaxf(fmaf(v.z, sc[c + 2], sh[c + 2]), 0.f);
                    v.w = fmaxf(fmaf(v.w, sc[c + 3], sh[c + 3]), 0.f);
                }
            }
            float *d = xs + ((c4 * 4) * S::HH + hy) * S::HLD + hx;
            d[0] = v.x; d[S::HH * S::HLD] = v.y; d[2 * S::HH * S::HLD] = v.z; d[3 * S::HH * S::HLD] = v.w;
        }
        // ---- stage the weight panel [tap][ci][co] ----
        for (int idx = tid; idx < KS * KS * kCK * (COT / 4); idx += 256) {
            const int c4 = idx % (COT / 4), rest = idx / (COT / 4);
            const int ci = rest % kCK, tap = rest / kCK;
            const float4 v = __ldg(reinterpret_cast<const float4 *>(a.w + ((size_t)tap * a.Cin + ci0 + ci) * a.Cout + co0) + c4);
            reinterpret_cast<float4 *>(ws + (tap * kCK + ci) * COT)[c4] = v;
        }
        __syncthreads();
#pragma unroll 2
        for (int ci = 0; ci < kCK; ++ci) {
#pragma unroll
            for (int ky = 0; ky < KS; ++ky) {
                float xr[12];
                const float *xp = xs + (ci * S::HH + r + ky) * S::HLD + px0;
                {
                    const float4 t0 = *reinterpret_cast<const float4 *>(xp);
                    const float4 t1 = *reinterpret_cast<const float4 *>(xp + 4);
                    xr[0] = t0.x; xr[1] = t0.y; xr[2] = t0.z; xr[3] = t0.w;
                    xr[4] = t1.x; xr[5] = t1.y; xr[6] = t1.z; xr[7] = t1.w;
                    if (KS == 3) {
                        const float2 t2 = *reinterpret_cast<const float2 *>(xp + 8);
                        xr[8] = t2.x; xr[9] = t2.y;
                    }
                }
#pragma unroll
                for (int kx = 0; kx < KS; ++kx) {
                    const float *wp = ws + ((ky * KS + kx) * kCK + ci) * COT + cg * TCO;
                    float wv[TCO];
                    if constexpr (TCO == 4) {
                        const float4 t = *reinterpret_cast<const float4 *>(wp);
                        wv[0] = t.x; wv[1] = t.y; wv[2] = t.z; wv[3] = t.w;
                    } else {
                        const float2 t = *reinterpret_cast<const float2 *>(wp);
                        wv[0] = t.x; wv[1] = t.y;
                    }
#pragma unroll
                    for (int p = 0; p < 8; ++p)
#pragma unroll
                        for (int c = 0; c < TCO; ++c) acc[p][c] = fmaf(xr[p + kx], wv[c], acc[p][c]);
                }
            }
        }
        (void)NX;
    }

    // ---- epilogue ----
    const int gy = ty0 + r;
    if (gy >= a.H) return;
    const int co = co0 + cg * TCO;
    float bv[TCO];
#pragma unroll
    for (int c = 0; c < TCO; ++c) bv[c] = a.bias ? __ldg(a.bias + co + c) : 0.f;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const int gx = tx0 + px0 + p;
        if (gx >= a.W) continue;
        const size_t pix = ((size_t)b * a.H + gy) * a.W + gx;
        float v[TCO];
#pragma unroll
        for (int c = 0; c < TCO; ++c) v[c] = acc[p][c] + bv[c];
        if (a.raw) {
            float *d = a.raw + pix * a.ld_raw + a.off_raw + co;
#pragma unroll
            for (int c = 0; c < TCO; ++c) d[c] = v[c];
        }
        if (a.res) {
            const float *rp = a.res + pix * a.ld_res + a.off_res + co;
#pragma unroll
            for (int c = 0; c < TCO; ++c) v[c] += rp[c];
        }
        float *d = a.out + pix * a.ld_out + a.off_out + co;
#pragma unroll
        for (int c = 0; c < TCO; ++c) d[c] = v[c];
    }
}

// ------------------------------------------------------------------------------------------
// avg-pool 2x2 stride 2 (F.avg_pool2d, model/HGFilters.py:32,152)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) avgpool2_kernel(const float *__restrict__ in, float *__restrict__ out, int H,
                                                       int W, int C, size_t total4) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total4) return;
    const int c4n = C / 4, OW = W / 2, OH = H / 2;
    const int c4 = (int)(i % c4n);
    size_t t = i / c4n;
    const int ox = (int)(t % OW); t /= OW;
    const int oy = (int)(t % OH);
    const int b = (int)(t / OH);
    const float4 *p = reinterpret_cast<const float4 *>(in) + (((size_t)b * H + oy * 2) * W + ox * 2) * c4n + c4;
    const float4 v00 = __ldg(p), v01 = __ldg(p + c4n), v10 = __ldg(p + (size_t)W * c4n), v11 = __ldg(p + (size_t)W * c4n + c4n);
    float4 o;
    o.x = (((v00.x + v01.x) + v10.x) + v11.x) / 4.f;
    o.y = (((v00.y + v01.y) + v10.y) + v11.y) / 4.f;
    o.z = (((v00.z + v01.z) + v10.z) + v11.z) / 4.f;
    o.w = (((v00.w + v01.w) + v10.w) + v11.w) / 4.f;
    reinterpret_cast<float4 *>(out)[i] = o;
}

// ------------------------------------------------------------------------------------------
// out = up1 + bicubic_x2(low), align_corners=True, A = -0.75, clamped taps
// (F.interpolate(..., mode='bicubic', align_corners=True) + add, model/HGFilters.py:47-49)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cubic_coeffs(float t, float (&c)[4]) {
    const float A = -0.75f;
    const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
    c[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
    c[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
    c[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
    c[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

__global__ void __launch_bounds__(256) upsample_add_kernel(const float *__restrict__ low, float *__restrict__ up,
                                                           int IH, int IW, int C, size_t total4) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total4) return;
    const int c4n = C / 4, OW = IW * 2, OH = IH * 2;
    const int c4 = (int)(i % c4n);
    size_t t = i / c4n;
    const int ox = (int)(t % OW); t /= OW;
    const int oy = (int)(t % OH);
    const int b = (int)(t / OH);
    const float sy = OH > 1 ? (float)(IH - 1) / (float)(OH - 1) : 0.f;
    const float sx = OW > 1 ? (float)(IW - 1) / (float)(OW - 1) : 0.f;
    const float ry = sy * oy, rx = sx * ox;
    const int iy = (int)floorf(ry), ix = (int)floorf(rx);
    float cy[4], cx[4];
    cubic_coeffs(ry - iy, cy);
    cubic_coeffs(rx - ix, cx);
    const float4 *lb = reinterpret_cast<const float4 *>(low) + (size_t)b * IH * IW * c4n + c4;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int yy = min(max(iy - 1 + j, 0), IH - 1);
        float4 rsum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int xx = min(max(ix - 1 + k, 0), IW - 1);
            const float4 v = __ldg(lb + ((size_t)yy * IW + xx) * c4n);
            rsum.x = fmaf(v.x, cx[k], rsum.x); rsum.y = fmaf(v.y, cx[k], rsum.y);
            rsum.z = fmaf(v.z, cx[k], rsum.z); rsum.w = fmaf(v.w, cx[k], rsum.w);
        }
        o.x = fmaf(rsum.x, cy[j], o.x); o.y = fmaf(rsum.y, cy[j], o.y);
        o.z = fmaf(rsum.z, cy[j], o.z); o.w = fmaf(rsum.w, cy[j], o.w);
    }
    float4 u = reinterpret_cast<float4 *>(up)[i];
    u.x += o.x; u.y += o.y; u.z += o.z; u.w += o.w;
    reinterpret_cast<float4 *>(up)[i] = u;
}

// ------------------------------------------------------------------------------------------
// host-side graph walk
// ------------------------------------------------------------------------------------------
struct Act {
    float *p = nullptr;
    int C = 0, H = 0, W = 0;
};

struct Ctx {
    chore_handle *h;
    cudaStream_t st;
    int B;
    bool dry;                 // sizing pass: bump the arena, launch nothing
    char *base = nullptr;
    size_t top = 0, peak = 0;
    double *gn_base = nullptr;   // arena of [B][32][2] slots, zeroed once per encode
    int gn_next = 0, gn_slots = 0;
    int rc = 0;

    float *alloc(size_t floats) {
        const size_t bytes = (floats * sizeof(float) + 255) / 256 * 256;
        float *p = reinterpret_cast<float *>(base + top);
        top += bytes;
        if (top > peak) peak = top;
        return p;
    }
    Act act(int C, int H, int W) {
        Act a;
        a.C = C; a.H = H; a.W = W;
        a.p = alloc((size_t)B * H * W * C);
        return a;
    }
    double *gn_slot() {
        double *p = gn_base ? gn_base + (size_t)gn_next * B * kGroups * 2 : nullptr;
        ++gn_next;
        return p;
    }
};

#define ENC_LAUNCH(ctx, kernel, grid, block, smem, ...)                                      \
    do {                                                                                     \
        if (!(ctx).dry && (ctx).rc == 0) {                                                   \
            kernel<<<(grid), (block), (smem), (ctx).st>>>(__VA_ARGS__);                      \
            g_launch_count.fetch_add(1, std::memory_order_relaxed);                          \
            cudaError_t e_ = cudaGetLastError();                                             \
            if (e_ != cudaSuccess) {                                                         \
                chore_set_error("%s:%d: launch of %s -> %s", __FILE__, __LINE__, #kernel, cudaGetErrorString(e_)); \
                (ctx).rc = CHORE_ERR_CUDA;                                                   \
            }                                                                                \
        }                                                                                    \
    } while (0)

// statistics of channels [off, off+C) of an NHWC tensor with row length ld
double *gn_stats(Ctx &c, const float *in, int ld, int off, int C, int H, int W) {
    double *slot = c.gn_slot();
    const int HW = H * W;
    int ctas = (HW + 63) / 64;              // >= 64 pixels per CTA: enough CTAs to cover 148 SMs at 128^2
    if (ctas > 1184) ctas = 1184;
    const int ppc = (HW + ctas - 1) / ctas;
    dim3 grid((HW + ppc - 1) / ppc, c.B);
    if ((C / kGroups) % 4 == 0)
        ENC_LAUNCH(c, gn_stats_kernel, grid, 256, 0, in, ld, off, C, HW, ppc, slot);
    else
        ENC_LAUNCH(c, gn_stats_narrow_kernel, grid, 256, 0, in, ld, off, C, HW, ppc, slot);
    return slot;
}

template <int KS, int COT>
void launch_conv(Ctx &c, const ConvArgs &a) {
    using S = ConvSmem<KS, COT>;
    if (!c.dry && c.rc == 0)
        c.rc = []() -> int { CHORE_ONCE_PER_DEVICE(cudaFuncSetAttribute(conv_kernel<KS, COT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::bytes)); return (int)CHORE_OK; }();
    const int tiles = ((a.W + kTW - 1) / kTW) * ((a.H + kTH - 1) / kTH);
    dim3 grid(tiles, a.Cout / COT, a.B);
    auto kern = conv_kernel<KS, COT>;
    ENC_LAUNCH(c, kern, grid, 256, S::bytes, a);
}

void conv(Ctx &c, const ConvW &w, const ConvArgs &a0) {
    ConvArgs a = a0;
    a.w = w.w; a.bias = w.bias; a.Cin = w.cin; a.Cout = w.cout; a.B = c.B;
    if (encoder_use_tensor_cores() && w.wtc && a.H >= 8 && a.W >= 8 && w.cout % 32 == 0 && w.cout <= 256) {
        const size_t mark = c.top;
        const int cp = (w.cin + 63) / 64 * 64;
        float *planes = c.alloc((size_t)c.B * a.H * a.W * cp);      // 2 fp16 planes = 4 bytes per element
        if (!c.dry && c.rc == 0) {
            ConvTcArgs t{};
            t.in = a.in; t.ld_in = a.ld_in; t.off_in = a.off_in; t.Cin = w.cin;
            t.B = c.B; t.H = a.H; t.W = a.W; t.KS = w.kh; t.Cout = w.cout;
            t.wstream = w.wtc; t.bias = w.bias;
            t.gn_sums = a.gn_sums; t.gamma = a.gamma; t.beta = a.beta;
            t.out = a.out; t.ld_out = a.ld_out; t.off_out = a.off_out;
            t.res = a.res; t.ld_res = a.ld_res; t.off_res = a.off_res;
            t.raw = a.raw; t.ld_raw = a.ld_raw; t.off_raw = a.off_raw;
            t.planes = planes;
            c.rc = conv_tc_launch(t, c.st);
        }
        c.top = mark;
        return;
    }
    const bool narrow = (w.cout % 64) != 0;
    if (w.kh == 3) {
        if (narrow) launch_conv<3, 32>(c, a); else launch_conv<3, 64>(c, a);
    } else {
        if (narrow) launch_conv<1, 32>(c, a); else launch_conv<1, 64>(c, a);
    }
}

const ConvW &cw(Ctx &c, const std::string &k) { return c.h->enc.conv.at(k); }
const NormW &nw(Ctx &c, const std::string &k) { return c.h->enc.norm.at(k); }

// ConvBlock.forward (model/net_util.py:374-396).  `out` may be preallocated (C = cout).
Act conv_block(Ctx &c, const std::string &p, const Act &x, int cout, Act out = Act()) {
    const int cin = x.C, H = x.H, W = x.W;
    if (!out.p) out = c.act(cout, H, W);
    const size_t mark = c.top;
    Act t1 = c.act(cout / 2, H, W), t2 = c.act(cout / 4, H, W);
    double *sx = gn_stats(c, x.p, cin, 0, cin, H, W);
    const float *res = x.p;
    int ld_res = cin;
    if (cin != cout) {   // downsample = Sequential(bn4, ReLU, conv1x1) -> residual lives in `out`
        ConvArgs a{};
        a.in = x.p; a.ld_in = cin; a.H = H; a.W = W;
        a.gn_sums = sx; a.gamma = nw(c, p + ".bn4").gamma; a.beta = nw(c, p + ".bn4").beta;
        a.out = out.p; a.ld_out = cout;
        conv(c, cw(c, p + ".downsample.2"), a);
        res = out.p;
        ld_res = cout;
    }
    {   // conv1: x -> [0, cout/2)
        ConvArgs a{};
        a.in = x.p; a.ld_in = cin; a.H = H; a.W = W;
        a.gn_sums = sx; a.gamma = nw(c, p + ".bn1").gamma; a.beta = nw(c, p + ".bn1").beta;
        a.out = out.p; a.ld_out = cout; a.off_out = 0;
        a.res = res; a.ld_res = ld_res; a.off_res = 0;
        a.raw = t1.p; a.ld_raw = cout / 2;
        conv(c, cw(c, p + ".conv1"), a);
    }
    double *s1 = gn_stats(c, t1.p, cout / 2, 0, cout / 2, H, W);
    {   // conv2: o1 -> [cout/2, 3cout/4)
        ConvArgs a{};
        a.in = t1.p; a.ld_in = cout / 2; a.H = H; a.W = W;
        a.gn_sums = s1; a.gamma = nw(c, p + ".bn2").gamma; a.beta = nw(c, p + ".bn2").beta;
        a.out = out.p; a.ld_out = cout; a.off_out = cout / 2;
        a.res = res; a.ld_res = ld_res; a.off_res = cout / 2;
        a.raw = t2.p; a.ld_raw = cout / 4;
        conv(c, cw(c, p + ".conv2"), a);
    }
    double *s2 = gn_stats(c, t2.p, cout / 4, 0, cout / 4, H, W);
    {   // conv3: o2 -> [3cout/4, cout)
        ConvArgs a{};
        a.in = t2.p; a.ld_in = cout / 4; a.H = H; a.W = W;
        a.gn_sums = s2; a.gamma = nw(c, p + ".bn3").gamma; a.beta = nw(c, p + ".bn3").beta;
        a.out = out.p; a.ld_out = cout; a.off_out = 3 * cout / 4;
        a.res = res; a.ld_res = ld_res; a.off_res = 3 * cout / 4;
        conv(c, cw(c, p + ".conv3"), a);
    }
    c.top = mark;   // t1, t2 are dead (stream order makes reuse safe)
    return out;
}

Act avgpool(Ctx &c, const Act &x) {
    Act o = c.act(x.C, x.H / 2, x.W / 2);
    const size_t total4 = (size_t)c.B * o.H * o.W * o.C / 4;
    ENC_LAUNCH(c, avgpool2_kernel, (unsigned)((total4 + 255) / 256), 256, 0, x.p, o.p, x.H, x.W, x.C, total4);
    return o;
}

// HourGlass._forward (model/HGFilters.py:26-50)
Act hourglass(Ctx &c, const std::string &p, int level, const Act &x) {
    const std::string L = std::to_string(level);
    Act up1 = conv_block(c, p + ".b1_" + L, x, x.C);
    const size_t mark = c.top;
    Act low1 = conv_block(c, p + ".b2_" + L, avgpool(c, x), x.C);
    Act low2 = level > 1 ? hourglass(c, p, level - 1, low1) : conv_block(c, p + ".b2_plus_" + L, low1, x.C);
    Act low3 = conv_block(c, p + ".b3_" + L, low2, x.C);
    const size_t total4 = (size_t)c.B * up1.H * up1.W * up1.C / 4;
    ENC_LAUNCH(c, upsample_add_kernel, (unsigned)((total4 + 255) / 256), 256, 0, low3.p, up1.p, low3.H, low3.W,
               low3.C, total4);
    c.top = mark;
    return up1;
}

constexpr int kNumStack = 5, kDepth = 2;

void run_graph(Ctx &c, const float *images, int H, int W, float *feat, float *skip, float *normx) {
    const std::string p = "image_filter";
    const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4;
    // stem
    Act s0 = c.act(64, H2, W2);
    {
        dim3 grid((W2 + kStemTile - 1) / kStemTile, (H2 + kStemTile - 1) / kStemTile, c.B);
        const size_t smem = (size_t)(kStemPatchFloats + CHORE_IN_CH * 49 * 64) * sizeof(float);
        if (!c.dry && c.rc == 0)
            c.rc = [smem]() -> int { CHORE_ONCE_PER_DEVICE(cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); return (int)CHORE_OK; }();
        const ConvW &w = cw(c, p + ".conv1");
        ENC_LAUNCH(c, stem_conv_kernel, grid, 256, smem, images, H, W, w.w, w.bias, s0.p);
    }
    double *ss = gn_stats(c, s0.p, 64, 0, 64, H2, W2);
    Act tmpx;
    tmpx.p = skip; tmpx.C = 64; tmpx.H = H2; tmpx.W = W2;
    {
        const size_t total4 = (size_t)c.B * H2 * W2 * 64 / 4;
        const NormW &n = nw(c, p + ".bn1");
        ENC_LAUNCH(c, gn_relu_apply_kernel, (unsigned)((total4 + 255) / 256), 256, 0, s0.p, skip, 64, H2 * W2, ss,
                   n.gamma, n.beta, total4);
    }
    Act x = conv_block(c, p + ".conv2", tmpx, 128);
    Act nx;
    if (normx) {
        nx.p = normx; nx.C = 128; nx.H = H4; nx.W = W4;
        const size_t total4 = (size_t)c.B * H4 * W4 * 128 / 4;
        ENC_LAUNCH(c, avgpool2_kernel, (unsigned)((total4 + 255) / 256), 256, 0, x.p, nx.p, x.H, x.W, x.C, total4);
    } else {
        nx = avgpool(c, x);
    }
    x = conv_block(c, p + ".conv3", nx, 128);
    Act previous = conv_block(c, p + ".conv4", x, 256);
    for (int i = 0; i < kNumStack; ++i) {
        const std::string si = std::to_string(i);
        const size_t mark = c.top;
        Act hg = hourglass(c, p + ".m" + si, kDepth, previous);
        Act ll = conv_block(c, p + ".top_m_" + si, hg, 256);
        Act ll2 = c.act(256, H4, W4);
        {   // conv_last (1x1 + bias); its GroupNorm+ReLU (bn_end) is applied by the consumers
            ConvArgs a{};
            a.in = ll.p; a.ld_in = 256; a.H = H4; a.W = W4;
            a.out = ll2.p; a.ld_out = 256;
            conv(c, cw(c, p + ".conv_last" + si), a);
        }
        double *se = gn_stats(c, ll2.p, 256, 0, 256, H4, W4);
        const NormW &ne = nw(c, p + ".bn_end" + si);
        const bool last = i == kNumStack - 1;
        Act out;
        out.C = 256; out.H = H4; out.W = W4;
        out.p = last ? feat : c.alloc((size_t)c.B * H4 * W4 * 256);
        {   // l_i: the stack output
            ConvArgs a{};
            a.in = ll2.p; a.ld_in = 256; a.H = H4; a.W = W4;
            a.gn_sums = se; a.gamma = ne.gamma; a.beta = ne.beta;
            a.out = out.p; a.ld_out = 256;
            conv(c, cw(c, p + ".l" + si), a);
        }
        if (!last) {   // previous = previous + bl(ll) + al(out)   (model/HGFilters.py:180-183)
            ConvArgs a{};
            a.in = ll2.p; a.ld_in = 256; a.H = H4; a.W = W4;
            a.gn_sums = se; a.gamma = ne.gamma; a.beta = ne.beta;
            a.out = previous.p; a.ld_out = 256;
            a.res = previous.p; a.ld_res = 256;
            conv(c, cw(c, p + ".bl" + si), a);
            ConvArgs a2{};
            a2.in = out.p; a2.ld_in = 256; a2.H = H4; a2.W = W4;
            a2.out = previous.p; a2.ld_out = 256;
            a2.res = previous.p; a2.ld_res = 256;
            conv(c, cw(c, p + ".al" + si), a2);
        }
        c.top = mark;
    }
}

}   // namespace

// ---------------------------------------------------------------------------------------------
// weights: conv (Cout,Cin,kh,kw) -> [kh][kw][Cin][Cout]; GroupNorm affine as is
// ---------------------------------------------------------------------------------------------
int encoder_load_weights(chore_handle *h, const std::map<std::string, const chore_tensor_desc *> &t) {
    if (!t.count("image_filter.conv1.weight")) return CHORE_OK;   // encoder absent
    EncoderWeights &e = h->enc;
    std::vector<float> src, dst;
    for (const auto &kv : t) {
        const std::string &name = kv.first;
        if (name.rfind("image_filter.", 0) != 0) continue;
        const chore_tensor_desc *d = kv.second;
        size_t n = 1;
        for (int i = 0; i < d->ndim; ++i) n *= (size_t)d->shape[i];
        src.resize(n);
        if (d->on_device)
            CHORE_CUDA(cudaMemcpy(src.data(), d->data, n * sizeof(float), cudaMemcpyDeviceToHost));
        else
            memcpy(src.data(), d->data, n * sizeof(float));
        const size_t dot = name.rfind('.');
        const std::string base = name.substr(0, dot), leaf = name.substr(dot + 1);
        float *dev = nullptr;
        if (d->ndim == 4) {
            const int co = (int)d->shape[0], ci = (int)d->shape[1], kh = (int)d->shape[2], kw = (int)d->shape[3];
            dst.resize(n);
            for (int o = 0; o < co; ++o)
                for (int i = 0; i < ci; ++i)
                    for (int y = 0; y < kh; ++y)
                        for (int x = 0; x < kw; ++x)
                            dst[(((size_t)y * kw + x) * ci + i) * co + o] = src[(((size_t)o * ci + i) * kh + y) * kw + x];
            if (base == "image_filter.conv1") {
                // stem layout is [ci][ky][kx][co] (the kernel walks channels outermost)
                for (int o = 0; o < co; ++o)
                    for (int i = 0; i < ci; ++i)
                        for (int y = 0; y < kh; ++y)
                            for (int x = 0; x < kw; ++x)
                                dst[((((size_t)i * kh + y) * kw) + x) * co + o] = src[(((size_t)o * ci + i) * kh + y) * kw + x];
            }
            if (int rc = chore_dev_alloc(h, reinterpret_cast<void **>(&dev), n * sizeof(float))) return rc;
            CHORE_CUDA(cudaMemcpy(dev, dst.data(), n * sizeof(float), cudaMemcpyHostToDevice));
            ConvW &w = e.conv[base];
            w.w = dev; w.kh = kh; w.kw = kw; w.cin = ci; w.cout = co;
            if (base == "image_filter.conv1" && encoder_mode() == "hx" && ci * kh * kw <= 256 && co % 32 == 0) {
                // stem as a 1x1 convolution over im2col columns (encoder_hx.cu): [co][k = (ci, ky, kx)], zero padded to 256
                std::vector<float> wst((size_t)co * 256, 0.f);
                for (int o = 0; o < co; ++o)
                    for (int k = 0; k < ci * kh * kw; ++k) wst[(size_t)o * 256 + k] = src[(size_t)o * ci * kh * kw + k];
                if (int rc = conv_hx_pack_weights(h, wst.data(), co, 256, 1, 1, &w.whx)) return rc;
            }
            if (base != "image_filter.conv1" && (kh == 1 || kh == 3) && co % 32 == 0 && co <= 256) {
                const std::string mode = encoder_mode();
                if (mode == "hx") {
                    if (int rc = conv_hx_pack_weights(h, src.data(), co, ci, kh, kw, &w.whx)) return rc;
                } else if (mode == "tc1") {
                    if (int rc = conv_tc_pack_weights(h, src.data(), co, ci, kh, kw, &w.wtc)) return rc;
                }
            }
        } else if (d->ndim == 1) {
            if (int rc = chore_dev_alloc(h, reinterpret_cast<void **>(&dev), n * sizeof(float))) return rc;
            CHORE_CUDA(cudaMemcpy(dev, src.data(), n * sizeof(float), cudaMemcpyHostToDevice));
            const bool is_norm = base.find(".bn") != std::string::npos || base.find(".downsample.0") != std::string::npos;
            if (is_norm) {
                NormW &nrm = e.norm[base];
                nrm.c = (int)n;
                if (leaf == "weight") nrm.gamma = dev; else nrm.beta = dev;
            } else if (leaf == "bias") {
                e.conv[base].bias = dev;
            }
        }
    }
    e.loaded = true;
    return CHORE_OK;
}

extern "C" int chore_encode(chore_handle *h, const float *images, int B, int H, int W, float *feat, float *skip,
                            float *normx, void *stream) {
    CHORE_CHECK(h && images && feat && skip, "null argument");
    CHORE_CHECK(B > 0 && H >= 32 && W >= 32 && H % 16 == 0 && W % 16 == 0, "image size %dx%d must be a multiple of 16 (>= 32)", H, W);
    if (!h->enc.loaded) {
        chore_set_error("encoder weights not loaded (chore_load_weights)");
        return CHORE_ERR_NO_WEIGHTS;
    }
    if (strcmp(encoder_mode(), "hx") == 0) return encode_hx(h, images, B, H, W, feat, skip, normx, static_cast<cudaStream_t>(stream));
    Ctx dry{};
    dry.h = h; dry.B = B; dry.dry = true; dry.st = nullptr;
    try {
        run_graph(dry, images, H, W, feat, skip, normx);
    } catch (const std::out_of_range &) {
        chore_set_error("encoder weights incomplete: a tensor of the reference state_dict is missing");
        return CHORE_ERR_NO_WEIGHTS;
    }
    const size_t gn_bytes = (size_t)dry.gn_next * B * kGroups * 2 * sizeof(double);
    if (int rc = chore_ws_reserve(h, dry.peak + 256)) return rc;
    if (int rc = chore_ws2_reserve(h, gn_bytes)) return rc;
    Ctx c{};
    c.h = h; c.B = B; c.dry = false; c.st = static_cast<cudaStream_t>(stream);
    c.base = static_cast<char *>(h->ws);
    c.gn_base = static_cast<double *>(h->ws2);
    c.gn_slots = dry.gn_next;
    CHORE_CUDA(cudaMemsetAsync(h->ws2, 0, gn_bytes, c.st));
    run_graph(c, images, H, W, feat, skip, normx);
    return c.rc;
}
