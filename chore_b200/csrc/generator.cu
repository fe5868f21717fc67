// Neural point-cloud generation loop on the device: the bookkeeping of Generator.gen_pc_batch (recon/generator.py:123-217)
// around the field queries -- surface filter + ordered compaction, resampling with perturbation, min-count truncation
// and the final mean / argmax reductions -- which the reference does with boolean indexing, Python lists and a host
// round trip (.cpu(), .item()) per image and outer iteration.
//
//   gen_compact_kernel   mask = min(df_k, threshold) < filter_val; STABLE compaction (index order, like boolean-mask
//                        indexing): the pre-projection samples of the hits are packed for the resampling step and, from
//                        the second outer iteration on, the projected points + their predictions (part label = argmax of
//                        the 14 logits, 9 PCA values, 6 centre values) are appended to the per-image output buffers.
//   gen_resample_kernel  next samples: a hit drawn uniformly + N(0, (threshold/3)^2) noise, or -- when an image has at
//                        most one hit -- an initial sample + N(0, 0.5^2) noise (generator.py:166-176).  Random numbers come
//                        from Philox4x32-10 (curand), stream = image * sample_num + j, offset = outer iteration, or from
//                        caller-provided tensors (tests / replaying the reference's draws).
//   gen_total_kernel     samples_count += min over images of this iteration's hit count (generator.py:160).
//   gen_finalize_kernel  pca_axis / centers = mean over the first samples_count kept points (compose_outdict :190-217),
//                        summed in index order by one block per image: deterministic.
#include "common.cuh"

#include <curand_kernel.h>

namespace {

constexpr int kCompactThreads = 1024;

__global__ void __launch_bounds__(kCompactThreads) gen_compact_kernel(
    const float *__restrict__ df, int df_idx, float threshold, float filter_val, const float *__restrict__ surf,
    const float *__restrict__ samples, const float *__restrict__ pca, const float *__restrict__ parts,
    const float *__restrict__ centers, int N, int cap, int append, float *__restrict__ out_points,
    int32_t *__restrict__ out_labels, float *__restrict__ out_pca, float *__restrict__ out_centers,
    int32_t *__restrict__ out_count, float *__restrict__ packed, int32_t *__restrict__ iter_count) {
    __shared__ int warp_tot[32];
    __shared__ int base_s;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *dfb = df + ((size_t)b * 2 + df_idx) * N;
    const int base_out = append ? out_count[b] : 0;
    int running = 0;                                       // hits of this image before the current sweep position
    for (int n0 = 0; n0 < N; n0 += kCompactThreads) {
        const int n = n0 + tid;
        bool hit = false;
        if (n < N) hit = fminf(dfb[n], threshold) < filter_val;       // clamp(df, max=threshold) < filter_val (generator.py:150-151)
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        const int in_warp = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        if (warp == 0) {
            int v = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += t;
            }
            warp_tot[lane] = v;                            // inclusive prefix over the 32 warps
            if (lane == 31) base_s = v;
        }
        __syncthreads();
        if (hit) {
            const int pos = running + (warp ? warp_tot[warp - 1] : 0) + in_warp;      // rank among this image's hits
            const float *s = samples + ((size_t)b * N + n) * 3;
            float *pk = packed + ((size_t)b * N + pos) * 3;
            pk[0] = s[0]; pk[1] = s[1]; pk[2] = s[2];
            const int o = base_out + pos;
            if (append && o < cap) {
                const float *sp = surf + ((size_t)b * N + n) * 3;
                float *op = out_points + ((size_t)b * cap + o) * 3;
                op[0] = sp[0]; op[1] = sp[1]; op[2] = sp[2];
                int best = 0;
                float bv = parts[((size_t)b * 14) * N + n];
#pragma unroll
                for (int c = 1; c < 14; ++c) {             // torch.argmax: first maximum wins
                    const float v = parts[((size_t)b * 14 + c) * N + n];
                    if (v > bv) { bv = v; best = c; }
                }
                out_labels[(size_t)b * cap + o] = best;
#pragma unroll
                for (int c = 0; c < 9; ++c) out_pca[((size_t)b * cap + o) * 9 + c] = pca[((size_t)b * 9 + c) * N + n];
#pragma unroll
                for (int c = 0; c < 6; ++c) out_centers[((size_t)b * cap + o) * 6 + c] = centers[((size_t)b * 6 + c) * N + n];
            }
        }
        running += base_s;
        __syncthreads();
    }
    if (tid == 0) {
        iter_count[b] = running;
        if (append) out_count[b] = min(cap, base_out + running);
    }
}

__global__ void __launch_bounds__(256) gen_resample_kernel(const float *__restrict__ packed, const int32_t *__restrict__ iter_count,
                                                           const float *__restrict__ samples_init, int N, int Ninit, int sample_num,
                                                           float sigma_hit, float sigma_miss, unsigned long long seed,
                                                           unsigned long long offset, const float *__restrict__ uniforms,
                                                           const float *__restrict__ normals, float *__restrict__ out) {
    const int b = blockIdx.y;
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= sample_num) return;
    float u, n0, n1, n2;
    if (uniforms != nullptr) {
        u = uniforms[(size_t)b * sample_num + j];
        const float *nn = normals + ((size_t)b * sample_num + j) * 3;
        n0 = nn[0]; n1 = nn[1]; n2 = nn[2];
    } else {
        curandStatePhilox4_32_10_t st;
        curand_init(seed, (unsigned long long)b * sample_num + j, offset, &st);
        u = curand_uniform(&st);                           // (0, 1]
        const float4 g = curand_normal4(&st);
        n0 = g.x; n1 = g.y; n2 = g.z;
        u = 1.0f - u;                                      // [0, 1)
    }
    const int hits = iter_count[b];
    const float *src;
    float sigma;
    if (hits > 1) {                                        // generator.py:168-172
        int idx = (int)(u * (float)hits);
        idx = idx < hits ? idx : hits - 1;
        src = packed + ((size_t)b * N + idx) * 3;
        sigma = sigma_hit;
    } else {                                               // generator.py:174-176: restart from the initial samples
        int idx = (int)(u * (float)Ninit);
        idx = idx < Ninit ? idx : Ninit - 1;
        src = samples_init + ((size_t)b * Ninit + idx) * 3;
        sigma = sigma_miss;
    }
    float *o = out + ((size_t)b * sample_num + j) * 3;
    // two roundings like `samples + sigma * torch.randn(...)` (no FMA contraction): bit-identical to the eager formula
    o[0] = __fadd_rn(src[0], __fmul_rn(sigma, n0)); o[1] = __fadd_rn(src[1], __fmul_rn(sigma, n1)); o[2] = __fadd_rn(src[2], __fmul_rn(sigma, n2));
}

__global__ void gen_total_kernel(const int32_t *__restrict__ iter_count, int B, int32_t *__restrict__ samples_count) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int m = iter_count[0];
        for (int b = 1; b < B; ++b) m = min(m, iter_count[b]);
        samples_count[0] += m;
    }
}

__global__ void __launch_bounds__(256) gen_finalize_kernel(const float *__restrict__ out_pca, const float *__restrict__ out_centers,
                                                           int cap, const int32_t *__restrict__ samples_count,
                                                           float *__restrict__ pca_mean, float *__restrict__ centers_mean) {
    __shared__ float red[256];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n = min(samples_count[0], cap);
    for (int c = 0; c < 15; ++c) {
        const float *src = c < 9 ? out_pca + (size_t)b * cap * 9 + c : out_centers + (size_t)b * cap * 6 + (c - 9);
        const int ld = c < 9 ? 9 : 6;
        float s = 0.f;
        for (int i = tid; i < n; i += 256) s += src[(size_t)i * ld];
        red[tid] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (tid < o) red[tid] += red[tid + o];
            __syncthreads();
        }
        if (tid == 0) {
            const float m = n > 0 ? red[0] / (float)n : 0.f;
            if (c < 9) pca_mean[b * 9 + c] = m; else centers_mean[b * 6 + c - 9] = m;
        }
        __syncthreads();
    }
}

}   // namespace

extern "C" int chore_gen_compact(chore_handle *h, const float *df, int df_idx, float threshold, float filter_val, const float *surf,
                                 const float *samples, const float *pca, const float *parts, const float *centers, int B, int N,
                                 int cap, int append, float *out_points, int32_t *out_labels, float *out_pca, float *out_centers,
                                 int32_t *out_count, float *packed, int32_t *iter_count, void *stream) {
    CHORE_CHECK(h && df && samples && packed && iter_count && B > 0 && N > 0 && (df_idx == 0 || df_idx == 1), "bad arguments");
    CHORE_CHECK(!append || (surf && pca && parts && centers && out_points && out_labels && out_pca && out_centers && out_count && cap > 0),
                "append needs the prediction tensors and the output buffers");
    CHORE_LAUNCH(gen_compact_kernel, B, kCompactThreads, 0, static_cast<cudaStream_t>(stream), df, df_idx, threshold, filter_val, surf,
                 samples, pca, parts, centers, N, cap, append, out_points, out_labels, out_pca, out_centers, out_count, packed,
                 iter_count);
    return CHORE_OK;
}

extern "C" int chore_gen_resample(chore_handle *h, const float *packed, const int32_t *iter_count, const float *samples_init, int B,
                                  int N, int Ninit, int sample_num, float sigma_hit, float sigma_miss, uint64_t seed, uint64_t offset,
                                  const float *uniforms, const float *normals, float *out, void *stream) {
    CHORE_CHECK(h && packed && iter_count && samples_init && out && B > 0 && N > 0 && Ninit > 0 && sample_num > 0, "bad arguments");
    CHORE_CHECK((uniforms == nullptr) == (normals == nullptr), "uniforms and normals come together");
    dim3 grid((sample_num + 255) / 256, B);
    CHORE_LAUNCH(gen_resample_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), packed, iter_count, samples_init, N, Ninit,
                 sample_num, sigma_hit, sigma_miss, (unsigned long long)seed, (unsigned long long)offset, uniforms, normals, out);
    return CHORE_OK;
}

extern "C" int chore_gen_total(chore_handle *h, const int32_t *iter_count, int B, int32_t *samples_count, void *stream) {
    CHORE_CHECK(h && iter_count && samples_count && B > 0, "bad arguments");
    CHORE_LAUNCH(gen_total_kernel, 1, 32, 0, static_cast<cudaStream_t>(stream), iter_count, B, samples_count);
    return CHORE_OK;
}

extern "C" int chore_gen_finalize(chore_handle *h, const float *out_pca, const float *out_centers, int B, int cap,
                                  const int32_t *samples_count, float *pca_mean, float *centers_mean, void *stream) {
    CHORE_CHECK(h && out_pca && out_centers && samples_count && pca_mean && centers_mean && B > 0 && cap > 0, "bad arguments");
    CHORE_LAUNCH(gen_finalize_kernel, B, 256, 0, static_cast<cudaStream_t>(stream), out_pca, out_centers, cap, samples_count, pca_mean,
                 centers_mean);
    return CHORE_OK;
}
