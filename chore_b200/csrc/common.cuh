// Shared declarations of libchore_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <map>
#include <string>
#include <vector>

#include "../../include/chore_b200.h"

// ---- error plumbing ---------------------------------------------------------------------
void chore_set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launch_count;

#define CHORE_CUDA(call)                                                                     \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            chore_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return CHORE_ERR_CUDA;                                                           \
        }                                                                                    \
    } while (0)

#define CHORE_CHECK(cond, ...)                                                               \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            chore_set_error(__VA_ARGS__);                                                    \
            return CHORE_ERR_INVALID;                                                        \
        }                                                                                    \
    } while (0)

// cudaFuncSetAttribute is per device (context): run `call` once per device of this process, not once per process.
// The bit test / set is atomic; two threads racing on the same device at worst both make the (idempotent) call.
#define CHORE_ONCE_PER_DEVICE(call)                                                          \
    do {                                                                                     \
        static std::atomic<uint64_t> done_{0};                                               \
        int dev_ = 0;                                                                        \
        cudaGetDevice(&dev_);                                                                \
        if (!((done_.load(std::memory_order_acquire) >> (dev_ & 63)) & 1ull)) {              \
            CHORE_CUDA(call);                                                                \
            done_.fetch_or(1ull << (dev_ & 63), std::memory_order_release);                  \
        }                                                                                    \
    } while (0)

// every kernel launch goes through this so chore_launch_count() is exact
#define CHORE_LAUNCH(kernel, grid, block, smem, stream, ...)                                 \
    do {                                                                                     \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                          \
        g_launch_count.fetch_add(1, std::memory_order_relaxed);                              \
        CHORE_CUDA(cudaGetLastError());                                                      \
    } while (0)

// Programmatic dependent launch: the kernel may start while its predecessor in the stream drains (its CTAs then block
// in griddepcontrol.wait -- chore_pdl_wait() -- until the predecessor's memory is visible).  Every kernel launched this
// way MUST call chore_pdl_wait() before it reads or writes anything a predecessor touches and before it exits.
// Opt-in with CHORE_B200_PDL=1 (ordinary stream order otherwise: replayed from a CUDA graph the encoder gained nothing).
bool chore_pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t chore_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = chore_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define CHORE_LAUNCH_PDL(kernel, grid, block, smem, stream, ...)                             \
    do {                                                                                     \
        g_launch_count.fetch_add(1, std::memory_order_relaxed);                              \
        CHORE_CUDA(chore_launch_pdl(kernel, (grid), (block), (smem), (stream), __VA_ARGS__)); \
    } while (0)
#ifdef __CUDACC__
__device__ __forceinline__ void chore_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void chore_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

// ---- constants of the chore-release configuration ------------------------------------------
constexpr int kFeatC = CHORE_FEAT_CH;     // 256
constexpr int kSkipC = CHORE_SKIP_CH;     // 64
constexpr int kPointC = CHORE_POINT_CH;   // 323
constexpr int kPointCPad = 324;           // padded to a multiple of 4
constexpr int kHidden = 128;
constexpr int kNumHeads = 4;              // kernel order: df, pca, parts, centers
constexpr int kHeadOut[kNumHeads] = {2, 9, 14, 6};
constexpr int kW1oLd = 384;               // leading dimension of the [512][323] layout (3 x 128)

// ---- MLP weights in kernel layouts ------------------------------------------------------------
struct MlpWeights {
    // "t" layouts are [K][N] (input-major), used by the forward GEMMs;
    // "o" layouts are the PyTorch [out][in] layouts, used as [K][N] by the backward GEMMs.
    float *w1t = nullptr;   // [324][512]   heads concatenated along N (head h at cols 128h)
    float *w1o = nullptr;   // [512][384]   zero padded beyond col 323
    float *b1 = nullptr;    // [512]
    float *w2t = nullptr, *w2o = nullptr, *b2 = nullptr;   // [4][128][128], [4][128]
    float *w3t = nullptr, *w3o = nullptr, *b3 = nullptr;
    float *w4 = nullptr;    // [4][16][128]  rows >= n_out zero
    float *b4 = nullptr;    // [4][16]
    unsigned char *wstream = nullptr;   // tensor-core path: per-tile stream of pre-swizzled fp16 hi/lo panels
    unsigned char *wstream_bwd = nullptr;   // 4 per-head streams for the tensor-core backward
    bool loaded = false;
};

// ---- encoder weights ------------------------------------------------------------------------------
struct ConvW {
    unsigned char *wtc = nullptr;   // first tensor-core path: [tap][kb][hi|lo] fp16 panels (conv_tc.cu), packed on demand
    unsigned char *whx = nullptr;   // halo/transform path: [kb][tap][hi|lo] fp16 panels (conv_hx.cu)
    float *w = nullptr;     // [kh][kw][cin][cout]  (HWIO)
    float *bias = nullptr;  // [cout] or null
    int kh = 0, kw = 0, cin = 0, cout = 0;
};
struct NormW {
    float *gamma = nullptr, *beta = nullptr;
    int c = 0;
};
struct EncoderWeights {
    std::map<std::string, ConvW> conv;
    std::map<std::string, NormW> norm;
    bool loaded = false;
};

// ---- SMPL-H body model ----------------------------------------------------------------------------
struct LbsModel {
    int V = 0, J = 0, nb = 0;
    float *v_template = nullptr;   // [V][3]
    float *shapedirs = nullptr;    // [V][3][nb]
    float *posedirs = nullptr;     // [V][3][(J-1)*9]
    float *weights = nullptr;      // [V][J]
    float *j_template = nullptr;   // [J][3]      = J_regressor . v_template
    float *j_shapedirs = nullptr;  // [J][3][nb]  = J_regressor . shapedirs
    int32_t *parents = nullptr;    // device [J]
    std::vector<int> parents_h;
    bool loaded = false;
};

// ---- landmark regressors (body25 | face | hand rows stacked), CSR and its transpose ------------------
struct LandmarkModel {
    int L = 0, V = 0, nnz = 0;
    int32_t *rowptr = nullptr, *col = nullptr;     // [L+1], [nnz]   landmarks <- vertices
    float *val = nullptr;
    int32_t *t_rowptr = nullptr, *t_col = nullptr; // [V+1], [nnz]   vertices <- landmarks (adjoint, no atomics)
    float *t_val = nullptr;
    bool loaded = false;
};

// dense-grid query with layer 1 folded into the feature maps (query_g.cu)
struct QueryGWeights {
    unsigned char *wf[2] = {nullptr, nullptr};   // conv_hx 1x1 weights: W1[:, 0:256] rows of heads {0,1} / {2,3}
    unsigned char *ws[2] = {nullptr, nullptr};   // same for the 64 stem-skip channels W1[:, 259:323]
    float *wz = nullptr;                         // [3][512] columns of x, y, z - 2.2
    float *gf = nullptr, *gs = nullptr;          // projected maps of ONE image: [fh*fw][512], [4*fh*fw][512]
    size_t map_px = 0;                           // fh*fw the maps were allocated for
};

struct chore_handle {
    int device = 0;
    int sm_count = 148;
    MlpWeights mlp;
    QueryGWeights qg;
    EncoderWeights enc;
    LbsModel lbs;
    LandmarkModel lmk;
    // growable scratch (encoder activations, LBS intermediates)
    void *ws = nullptr;
    size_t ws_bytes = 0;
    void *ws2 = nullptr;
    size_t ws2_bytes = 0;
    void *bwd_ws = nullptr;      // query backward: gX scratch (B*N, 384)
    size_t bwd_ws_bytes = 0;
    void *lbs_ws = nullptr;      // LBS / rigid intermediates (own buffer: may interleave with encode)
    size_t lbs_ws_bytes = 0;
    std::vector<void *> owned;   // device allocations released by chore_destroy
    bool hx_configured = false;  // per-handle (= per-device) kernel attributes set (conv_hx.cu)
    struct EncoderPlan *enc_plan = nullptr;   // cached CUDA graphs of chore_encode (encoder_hx.cu)
};
void encoder_plan_destroy(chore_handle *h);

int chore_ws_reserve(chore_handle *h, size_t bytes);    // (re)allocates h->ws
int chore_ws2_reserve(chore_handle *h, size_t bytes);   // (re)allocates h->ws2
int chore_lbs_ws_reserve(chore_handle *h, size_t bytes);
int chore_dev_alloc(chore_handle *h, void **p, size_t bytes);

// tensor-core query path (query_tc.cu)
int query_tc_pack_weights(chore_handle *h, const std::vector<float> &w1, const std::vector<float> &w2,
                          const std::vector<float> &w3, const std::vector<float> &w4);
int query_tc_launch(chore_handle *h, const float *feat, const float *skip, int fh, int fw, const float *points,
                    const float *crop_center, int B, long long N, long long n_start, long long n_count, int grid_mode,
                    int batch_index, const int *res, const double *step, const double *bmin, unsigned head_mask,
                    float *const outs[4], unsigned char *in_img, cudaStream_t st);
int query_bwd_tc_launch(chore_handle *h, const float *feat, const float *skip, int fh, int fw, const float *points,
                        const float *crop_center, int B, long long N, const float *const g_heads[4], float *g_points,
                        void *workspace, size_t workspace_bytes, cudaStream_t st);
bool query_use_tensor_cores();   // CHORE_B200_QUERY=simt selects the fp32 SIMT kernel
// dense-grid forward through per-pixel projected maps (query_g.cu); experimental, opt-in with CHORE_B200_QUERY_PRE=1
int query_g_pack_weights(chore_handle *h, const std::vector<float> &w1);
int query_g_mode();
int query_g_launch(chore_handle *h, const float *feat, const float *skip, int fh, int fw, const float *crop_center, long long N,
                   long long n_start, long long n_count, int batch_index, const int *res, const double *step, const double *bmin,
                   unsigned head_mask, float *const outs[4], cudaStream_t st);
// cta_group::2 forward for all four heads (query_tc2.cu); experimental, opt-in with CHORE_B200_QUERY_2CTA=1
bool query_tc2_enabled();
int query_tc2_launch(chore_handle *h, const float *feat, const float *skip, int fh, int fw, const float *points,
                     const float *crop_center, int B, long long N, long long n_start, long long n_count, int grid_mode,
                     int batch_index, const int *res, const double *step, const double *bmin, float *const outs[4],
                     unsigned char *in_img, cudaStream_t st);

// tensor-core encoder convolutions (conv_tc.cu)
struct ConvTcArgs {
    const float *in; int ld_in, off_in, Cin;
    int B, H, W, KS, Cout;
    const unsigned char *wstream;
    const float *bias;
    const double *gn_sums; const float *gamma, *beta;   // GroupNorm+ReLU prologue (null = identity)
    float *out; int ld_out, off_out;
    const float *res; int ld_res, off_res;
    float *raw; int ld_raw, off_raw;
    void *planes;                                       // scratch: 2 * B*H*W*Cp halves (Cp = Cin rounded up to 64)
};
bool encoder_use_tensor_cores();   // CHORE_B200_ENCODER=simt selects the fp32 SIMT convolutions
int conv_tc_pack_weights(chore_handle *h, const float *w, int cout, int cin, int kh, int kw, unsigned char **dev);
int conv_tc_launch(const ConvTcArgs &a, cudaStream_t st);

// halo/transform tensor-core convolutions with fused GroupNorm prologue and statistics epilogue (conv_hx.cu)
struct ConvHxArgs {
    const float *in; int ld_in, off_in, Cin;     // fp32 NHWC input, channels [off_in, off_in + Cin) of rows of ld_in
    int B, H, W, KS, N;                          // N = Cout
    const unsigned char *w;                      // ConvW::whx
    const float *bias;                           // [N] or null
    const double *gn_in;                         // [B][32][2] sums of the input (GroupNorm(32, Cin)), null = identity
    const float *gamma, *beta; int relu;
    float *out; int ld_out, off_out;             // out = conv (+bias) (+res)
    const float *res; int ld_res, off_res;       // optional residual (may alias out)
    float *raw; int ld_raw, off_raw;             // optional copy of conv (+bias) without the residual
    double *st_raw; int cpg_raw;                 // optional statistics of raw / out: [B][32][2], channels per group
    double *st_out; int cpg_out;
};
int conv_hx_pack_weights(chore_handle *h, const float *w, int cout, int cin, int kh, int kw, unsigned char **dev);
int conv_hx_configure(chore_handle *h);   // per-device kernel attributes (idempotent)
struct ConvHxPlan {
    int tiles, n_splits, k_splits;
    size_t part_floats;      // fp32 scratch for the partial tiles of a K split (0 = none)
    int counters;            // zero-initialised ints the launch needs (0 = none)
};
int conv_hx_plan(const chore_handle *h, const ConvHxArgs &a, ConvHxPlan *plan);
int conv_hx_launch(chore_handle *h, const ConvHxArgs &a, const ConvHxPlan &plan, float *part, int *part_cnt, cudaStream_t st);
int encode_hx(chore_handle *h, const float *images, int B, int H, int W, float *feat, float *skip, float *normx, cudaStream_t st);
const char *encoder_mode();   // CHORE_B200_ENCODER: "hx" (default), "tc1" (first tensor-core path), "simt"

// implemented per translation unit
int query_load_weights(chore_handle *h, const std::map<std::string, const chore_tensor_desc *> &t);
int encoder_load_weights(chore_handle *h, const std::map<std::string, const chore_tensor_desc *> &t);

// ---- small device helpers ------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
