/* chore_b200.h -- C ABI of libchore_b200.so, the sm_100a implementation of the CHORE hot path.
 *
 * The reference (xiexh20/CHORE) has no FFI: its hot path is Python duck-typing on
 * BasePIFuNet.filter/query/get_preds, the SMPL wrapper's forward() and three ReconFitterBase
 * helpers.  This header is the boundary a maintainer binds instead of the
 * torch ops listed beside every entry point (paths relative to the reference tree).  The
 * ctypes binding that the Python host side uses is chore_b200/_lib.py; INTEGRATION.md shows the
 * reference-side stub.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / C++ types.
 *   - every data pointer is a DEVICE pointer to fp32 unless the comment says otherwise.
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it.
 *   - return value: 0 = ok, otherwise a chore_status; chore_last_error() gives the text
 *     (thread-local).  Nothing throws across the ABI.
 *   - the caller owns every buffer; the handle owns the repacked weights and its workspace.
 *   - one handle per device; calls on one handle must not be issued concurrently.
 */
#ifndef CHORE_B200_H
#define CHORE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct chore_handle chore_handle;

enum chore_status {
    CHORE_OK = 0,
    CHORE_ERR_INVALID = 1,      /* bad argument (null pointer, bad shape)          */
    CHORE_ERR_CUDA = 2,         /* a CUDA runtime call failed                       */
    CHORE_ERR_NO_WEIGHTS = 3,   /* weights / body model not loaded yet              */
    CHORE_ERR_ARCH = 4          /* device is not sm_100                             */
};

/* constants of config/chore-release.json that the kernels are compiled for */
#define CHORE_IN_CH 5        /* RGBM3 input: RGB + person mask + object mask         */
#define CHORE_FEAT_CH 256    /* hourglass_dim                                        */
#define CHORE_SKIP_CH 64     /* stem width (tmpx skip feature)                       */
#define CHORE_POINT_CH 323   /* 256 + (x, y, z-2.2) + 64  (model/chore.py:139-143)   */
#define CHORE_NUM_PARTS 14
#define CHORE_SMPLH_VERTS 6890
#define CHORE_SMPLH_JOINTS 52

/* head_mask bits for chore_query_fwd / chore_query_bwd */
#define CHORE_HEAD_DF 1u
#define CHORE_HEAD_PCA 2u
#define CHORE_HEAD_PARTS 4u
#define CHORE_HEAD_CENTERS 8u
#define CHORE_HEAD_ALL 15u

/* one named fp32 tensor of the reference state_dict (checkpoint['model_state_dict'],
 * recon/generator.py:243-267; an optional "module." prefix is accepted) */
typedef struct {
    const char *name;      /* e.g. "image_filter.conv1.weight", "df.0.weight"        */
    const float *data;     /* contiguous fp32, PyTorch layout                        */
    int ndim;
    int64_t shape[4];
    int on_device;         /* 0: host pointer, 1: device pointer                     */
} chore_tensor_desc;

/* ---- lifetime ---------------------------------------------------------------------- */
int chore_create(int device, chore_handle **out);
void chore_destroy(chore_handle *h);
const char *chore_last_error(void);
/* ABI version and the number of kernels launched through this library since load
 * (bench.py reports the latter as gpu_launches). */
int chore_abi_version(void);
uint64_t chore_launch_count(void);

/* ---- weights: replaces nn.Module.load_state_dict (recon/generator.py:264) ------------- */
/* Repacks the reference tensors into the kernel layouts.  Tensors it does not know are
 * ignored; missing ones make the later calls fail with CHORE_ERR_NO_WEIGHTS. */
int chore_load_weights(chore_handle *h, const chore_tensor_desc *tensors, int n);

/* ---- encoder: replaces HGFilter.forward (model/HGFilters.py:144-185) called from
 *      CHORE.filter (model/chore.py:87-96), eval mode (last stack output only) -------- */
/* images: (B,5,H,W) NCHW in [0,1], H and W multiples of 16.
 * feat:   (B,H/4,W/4,256) NHWC  = im_feat_list[-1] of the reference, channels-last.
 * skip:   (B,H/2,W/2,64)  NHWC  = tmpx (post-ReLU stem), channels-last.
 * normx:  (B,H/4,W/4,128) NHWC or NULL. */
int chore_encode(chore_handle *h, const float *images, int B, int H, int W,
                 float *feat, float *skip, float *normx, void *stream);

/* ---- point query: replaces CHORE.query + decode (model/chore.py:107-167), i.e.
 *      KinectColorCamera.project_points (model/camera.py:44-88), index() twice
 *      (model/geometry.py:4-14), torch.cat, 16 conv1d, and the OUT_DIST masked write ---- */
/* feat (B,fh,fw,256) NHWC, skip (B,2fh,2fw,64) NHWC, points (B,N,3), crop_center (B,2).
 * Outputs in the reference layout: df (B,2,N), pca (B,9,N) (viewed (B,3,3,N) by the
 * caller), parts (B,14,N), centers (B,6,N); any output of a head not in head_mask may be
 * NULL.  in_img: (B,N) uint8 or NULL. */
int chore_query_fwd(chore_handle *h, const float *feat, const float *skip, int fh, int fw,
                    const float *points, const float *crop_center, int B, int N,
                    uint32_t head_mask, float *df, float *pca, float *parts, float *centers,
                    uint8_t *in_img, void *stream);

/* gradient of sum_k <g_k, head_k> w.r.t. the points: what the reference's callers get from
 * loss.backward() through conv1d/grid_sampler/projection (recon/generator.py:63-70,
 * recon/recon_fit_behave.py:149-152).  g_* have the forward output layouts; NULL = zero.
 * g_points: (B,N,3), overwritten. */
int chore_query_bwd(chore_handle *h, const float *feat, const float *skip, int fh, int fw,
                    const float *points, const float *crop_center, int B, int N,
                    const float *g_df, const float *g_pca, const float *g_parts,
                    const float *g_centers, float *g_points, void *stream);

/* Same with caller-owned scratch of chore_query_bwd_workspace_bytes(B, N) bytes (may be 0): required when the
 * call is captured in a CUDA graph, because the scratch pointer is baked into the captured launches.  A scratch of
 * k times that size lets the k heads with a non-NULL gradient be evaluated concurrently in ONE launch (each head
 * writes its own buffer, summed in head order afterwards: deterministic); with less, the heads run one after the
 * other.  Small queries (a fit step: 54..157 tiles, two heads) are latency bound and gain ~1.3x from this. */
size_t chore_query_bwd_workspace_bytes(int B, int N);
int chore_query_bwd_ws(chore_handle *h, const float *feat, const float *skip, int fh, int fw,
                       const float *points, const float *crop_center, int B, int N,
                       const float *g_df, const float *g_pca, const float *g_parts,
                       const float *g_centers, float *g_points, void *workspace,
                       size_t workspace_bytes, void *stream);

/* dense grid of model/sdf.py:4-48 (create_grid + batch_eval) evaluated without
 * materialising the coordinates: point i = (ix*ry + iy)*rz + iz (np.mgrid order),
 * coord = float32(b_min + (b_max-b_min)/res * idx) evaluated in float64 like numpy does there -- the bounds are
 * DOUBLES: float32 bounds move ~1/3 of the coordinates by one ulp, which at z = 0.2 m is 4e-5 texels and showed as
 * 2e-4 relative field error against the reference on the 256^3 grid.  Evaluates points [start, start+count). */
int chore_query_grid(chore_handle *h, const float *feat, const float *skip, int fh, int fw,
                     const float *crop_center, int b, const int res[3], const double b_min[3],
                     const double b_max[3], int64_t start, int64_t count, uint32_t head_mask,
                     float *df, float *pca, float *parts, float *centers, void *stream);

/* ---- SMPL-H linear blend skinning: replaces SMPL_Layer.forward
 *      (lib_smpl/smplpytorch/smplpytorch/pytorch/smpl_layer.py:72-175) ----------------- */
/* body-model buffers in the reference's registered-buffer layouts (smpl_layer.py:49-64):
 * v_template (V,3), shapedirs (V,3,10), posedirs (V,3,(J-1)*9), J_regressor (J,V) dense,
 * weights (V,J), parents (J) int32 (kintree_table row 0).  Host or device pointers. */
int chore_lbs_load_model(chore_handle *h, const float *v_template, const float *shapedirs,
                         const float *posedirs, const float *J_regressor, const float *weights,
                         const int32_t *parents, int V, int J, int n_betas, int on_device);
/* pose (B,J*3) axis-angle, betas (B,n_betas), trans (B,3), offsets (B,V,3) or NULL ->
 * verts (B,V,3), jtr (B,J,3), v_posed (B,V,3) or NULL, naked (B,V,3) or NULL. */
int chore_lbs_fwd(chore_handle *h, const float *pose, const float *betas, const float *trans,
                  const float *offsets, int B, float *verts, float *jtr, float *v_posed,
                  float *naked, void *stream);
/* analytic backward of the same call (inputs are re-read; nothing is saved by fwd).
 * g_verts (B,V,3), g_jtr (B,J,3) or NULL -> g_pose (B,J*3), g_betas (B,n_betas),
 * g_trans (B,3), g_offsets (B,V,3) or NULL. */
int chore_lbs_bwd(chore_handle *h, const float *pose, const float *betas, const float *trans,
                  const float *offsets, int B, const float *g_verts, const float *g_jtr,
                  float *g_pose, float *g_betas, float *g_trans, float *g_offsets, void *stream);

/* ---- landmark regressors: replaces batch_sparse_dense_matmul (lib_smpl/torch_functions.py:52-76)
 *      as used by SMPLPyTorchWrapperBatch*.get_landmarks (lib_smpl/wrapper_pytorch.py:78-90): the
 *      body25 (25), face (70) and hand (42) regressors of lib_smpl/body_landmark.py:16-28, stacked
 *      row-wise into one (L,V) CSR matrix ------------------------------------------------- */
/* rowptr (L+1), col (nnz), val (nnz): HOST pointers, landmarks <- vertices. */
int chore_landmarks_load(chore_handle *h, const int32_t *rowptr, const int32_t *col, const float *val,
                         int L, int V, int nnz);
/* verts (B,V,3) -> out (B,L,3) */
int chore_landmarks_fwd(chore_handle *h, const float *verts, int B, float *out, void *stream);
/* adjoint: g_out (B,L,3) -> g_verts (B,V,3), overwritten or accumulated into (accumulate != 0) */
int chore_landmarks_bwd(chore_handle *h, const float *g_out, int B, float *g_verts, int accumulate,
                        void *stream);

/* ---- rigid object transform: replaces ReconFitterBase.transform_obj_verts
 *      (recon/recon_fit_base.py:367-371): out = (verts @ R + t) * s --------------------- */
/* verts (B,N,3), R (B,3,3), t (B,3), s (B) -> out (B,N,3) */
int chore_rigid_fwd(chore_handle *h, const float *verts, const float *R, const float *t,
                    const float *s, int B, int N, float *out, void *stream);
/* g_out (B,N,3) -> g_R (B,3,3), g_t (B,3), g_s (B), g_verts (B,N,3) or NULL */
int chore_rigid_bwd(chore_handle *h, const float *verts, const float *R, const float *t,
                    const float *s, int B, int N, const float *g_out, float *g_R, float *g_t,
                    float *g_s, float *g_verts, void *stream);

/* SO(3) projection R = U diag(1,1,det(U V^T)) V^T: replaces ReconFitterBase.project_so3
 * (recon/recon_fit_base.py:167-188; torch.svd of a 3x3).  mats, out: (B,3,3). */
int chore_project_so3(chore_handle *h, const float *mats, int B, float *out, void *stream);
/* adjoint of the projection (what autograd derives through torch.svd in the reference):
 * g_out (B,3,3) -> g_mats (B,3,3) */
int chore_project_so3_bwd(chore_handle *h, const float *mats, const float *g_out, int B,
                          float *g_mats, void *stream);

/* ---- fit-step losses and optimiser: the non-field arithmetic of one optimisation step of
 *      ReconFitterBehave.forward_smpl / forward_step('object only') (recon/recon_fit_behave.py:165-222,
 *      293-337) with get_loss_weights (:339-358) folded into the coefficients, as closed-form loss
 *      values + gradients w.r.t. the network outputs / SMPL parameters, and torch.optim.Adam.step.
 *      All pointers are device pointers; `loss` is ONE device float that every call adds its terms to;
 *      `workspace` holds chore_fit_workspace_floats(B, N) floats. ------------------------------ */
size_t chore_fit_workspace_floats(int B, int N);
/* wd * sum min(df_h, 0.1) + wp * sum CE(parts, labels): df (B,2,N), parts (B,14,N), labels (B,N) int64
 * -> g_df (B,2,N), g_parts (B,14,N)   (recon_fit_base.py:537-542, recon_fit_behave.py:313) */
int chore_fit_smpl_field_grads(chore_handle *h, const float *df, const float *parts, const int64_t *labels,
                               int B, int N, float wd, float wp, float *g_df, float *g_parts, float *loss,
                               float *workspace, void *stream);
/* cz * sum_b (J[b,8].z - z0)^2 (+ cj * sum conf |proj(J) - kpts|^2 when body_kpts (B,n_joints,3) is given):
 * landmarks (B,L,3) -> g_landmarks (B,L,3).  cam = {fx_px, fy_px, cx_px, cy_px, crop_size/2,
 * net_in_size/crop_size}   (recon_fit_base.py:230-231, 653-676; model/camera.py:51-71) */
int chore_fit_landmark_grads(chore_handle *h, const float *landmarks, const float *body_kpts,
                             const float *crop_center, int B, int L, int n_joints, float z0, float cz, float cj,
                             const float cam[6], float *g_landmarks, float *loss, float *workspace, void *stream);
/* cb * sum_b |(pose[3:66]-mean) P|^2 + ch * sum_b,hands |(pose[66:]-mean_h) P_h|^2 + cp * sum_b |pose[3:72]-pose_init|^2;
 * gradients are ADDED to g_pose (B,156).  Priors / pose_init may be NULL (term skipped)
 * (lib_smpl/th_smpl_prior.py:32-39, lib_smpl/th_hand_prior.py:69-78, recon_fit_behave.py:317-319) */
int chore_fit_pose_prior_grads(chore_handle *h, const float *pose, const float *pose_init, const float *body_mean,
                               const float *body_prec, const float *hand_mean, const float *lhand_prec,
                               const float *rhand_prec, int B, int n_pose, float cb, float ch, float cp, float *g_pose,
                               float *loss, float *workspace, void *stream);
/* object step: dvec_b = mean(obj_b) - smpl_center_b - mean(centers_b[3:6]);
 * loss += wo * sum min(df_o, 0.8) + wc * sum_b |dvec_b|^2 + ws * sum_b (s_b - s0)^2;
 * g_df (B,2,N), g_centers (B,6,N), dvec (B,3) out   (recon_fit_base.py:513-520, recon_fit_behave.py:175-198) */
int chore_fit_obj_field_grads(chore_handle *h, const float *obj, const float *df, const float *centers,
                              const float *smpl_center, const float *s, int B, int N, float s0, float wo, float wc,
                              float ws, float *g_df, float *g_centers, float *dvec, float *loss, float *workspace,
                              void *stream);
/* x (B,N,3) += alpha * v (B,3) broadcast over N */
int chore_add_rowvec(chore_handle *h, float *x, const float *v, int B, int N, float alpha, void *stream);

/* ---- surface projection step of Generator.approx_surface (recon/generator.py:50-79): the two elementwise
 *      stages around the field query and its gradient.  df (B,2,N), df_idx 0 = human / 1 = object. ------ */
/* gradient of clamp(df[:, df_idx], max=threshold).sum() w.r.t. df: g_df (B,2,N) */
int chore_surface_clamp_grad(chore_handle *h, const float *df, int df_idx, float threshold, int B, int N,
                             float *g_df, void *stream);
/* out = points - F.normalize(g_points, dim=-1, eps=1e-12) * min(df[:, df_idx], threshold): points, g_points, out (B,N,3);
 * out may alias points */
int chore_surface_step(chore_handle *h, const float *points, const float *g_points, const float *df, int df_idx,
                       float threshold, int B, int N, float *out_points, void *stream);

/* ---- silhouette rasteriser of the 'sil' fitting phase: replaces neural_renderer's rasterize_silhouettes as used by
 *      SilLossROI.forward (recon/obj_pose_roi.py:159-172; external/neural_renderer/neural_renderer/cuda/
 *      rasterize_cuda_kernel.cu:25-216 forward_face_index_map kernels, :291-550 backward_pixel_map kernel).
 *      faces (B,F,3,3): per face 3 vertices (x, y in normalised image coordinates [-1,1], z = depth); image rows are NOT
 *      flipped here (the reference flips after rasterising, rasterize.py:318-322).  alpha (B,S,S) in {0,1},
 *      face_index (B,S,S) int32 (-1 = background); g_faces (B,F,3,3): only the x / y entries are non-zero. ------------ */
size_t chore_silhouette_workspace_bytes(int B, int F);
int chore_silhouette_fwd(chore_handle *h, const float *faces, int B, int F, int image_size, float near, float far,
                         float *alpha, int32_t *face_index, void *workspace, size_t workspace_bytes, void *stream);
int chore_silhouette_bwd(chore_handle *h, const float *faces, const int32_t *face_index, const float *alpha,
                         const float *g_alpha, int B, int F, int image_size, float eps, float *g_faces, void *stream);

/* ---- joint-phase contact term: ReconFitterBase.compute_contact_loss (recon/recon_fit_base.py:553-608): contact points =
 *      cross distance field < thresh (0.08 m; all points of a side that has none), split by SMPL part (fixed vertex labels
 *      `part_labels` (Nh), argmax of `part_o` (B,14,No) for the object points), one cloud pair per (image, part) present on
 *      both sides, pytorch3d chamfer_distance defaults over the pairs (squared distances, mean per cloud, mean over pairs,
 *      both directions).  loss: 1 float, n_pairs: 1 int32 (0 = "no contact": the reference adds no term);
 *      g_smpl (B,Nh,3) / g_obj (B,No,3): d loss / d points, optional. ---------------------------------------------------- */
size_t chore_contact_workspace_bytes(int B, int Nh, int No);
int chore_contact_loss(chore_handle *h, const float *smpl_verts, const float *object, const float *df_hum_o,
                       const float *df_obj_h, const float *part_o, const int32_t *part_labels, int B, int Nh, int No,
                       float thresh, float *loss, int32_t *n_pairs, float *g_smpl, float *g_obj, void *workspace,
                       size_t workspace_bytes, void *stream);

/* ---- bookkeeping of Generator.gen_pc_batch (recon/generator.py:123-217) on the device: what the reference does with
 *      boolean indexing, Python lists and .cpu() / .item() round trips per image and outer iteration.  B images, N samples
 *      of this outer iteration; `cap` = capacity of the per-image output buffers. -------------------------------------- */
/* hit = min(df[:, df_idx], threshold) < filter_val.  STABLE (index-order) compaction per image: the pre-projection
 * `samples` of the hits are packed into `packed` (B,N,3) and `iter_count[b]` = number of hits; with append != 0 the
 * projected points `surf`, the part label (argmax of the 14 logits, first maximum), the 9 PCA values and the 6 centre
 * values of the hits are appended at out_count[b] (entries beyond cap are dropped) -- parse_preds (:88-100). */
int chore_gen_compact(chore_handle *h, const float *df, int df_idx, float threshold, float filter_val, const float *surf,
                      const float *samples, const float *pca, const float *parts, const float *centers, int B, int N,
                      int cap, int append, float *out_points, int32_t *out_labels, float *out_pca, float *out_centers,
                      int32_t *out_count, float *packed, int32_t *iter_count, void *stream);
/* next samples (:164-177): image b with more than one hit draws sample_num of its packed hits uniformly and adds
 * N(0, sigma_hit^2) noise, otherwise it restarts from samples_init (B,Ninit,3) + N(0, sigma_miss^2).  Random numbers:
 * Philox4x32-10 with (seed, subsequence = b * sample_num + j, offset) -- or, when `uniforms` (B,sample_num) in [0,1) and
 * `normals` (B,sample_num,3) are given, those (index = floor(u * count)). */
int chore_gen_resample(chore_handle *h, const float *packed, const int32_t *iter_count, const float *samples_init, int B,
                       int N, int Ninit, int sample_num, float sigma_hit, float sigma_miss, uint64_t seed, uint64_t offset,
                       const float *uniforms, const float *normals, float *out, void *stream);
/* samples_count[0] += min_b iter_count[b]   (:160) */
int chore_gen_total(chore_handle *h, const int32_t *iter_count, int B, int32_t *samples_count, void *stream);
/* compose_outdict (:190-217): pca_mean (B,9) / centers_mean (B,6) = mean over the first samples_count[0] kept points */
int chore_gen_finalize(chore_handle *h, const float *out_pca, const float *out_centers, int B, int cap,
                       const int32_t *samples_count, float *pca_mean, float *centers_mean, void *stream);

/* torch.optim.Adam.step (weight_decay 0, amsgrad off) for up to CHORE_ADAM_MAX_ENTRIES small tensors in one
 * launch.  `step` is a device int32 step counter (read, then incremented by the kernel: graph-replay safe).
 * The reference fitting loops call zero_grad() once per OUTER iteration and loss.backward() in every inner step
 * (recon/recon_fit_behave.py:135-152,244-273), so the gradient Adam sees is the SUM of the inner-step gradients
 * since the last zero_grad().  `grad_acc` (optional, per entry) reproduces that: acc += gscale * grad, Adam reads
 * acc; the caller zeroes it where the reference calls zero_grad().  `gscale` (optional device scalar, default 1)
 * is the 1 / (1 + decay) factor of get_loss_weights (:339-358) that multiplies every loss term: a captured graph
 * follows the decay schedule by rewriting that scalar.  `loss_inout` (optional device scalar) is multiplied by
 * gscale in the same launch so the reported loss carries the same factor. */
#define CHORE_ADAM_MAX_ENTRIES 8
typedef struct {
    float *param;            /* (rows, cols) contiguous                                   */
    const float *grad;       /* (rows, cols) with leading dimension grad_ld (a column slice
                                of a wider gradient buffer is fine)                       */
    float *exp_avg;          /* (rows, cols) contiguous state                             */
    float *exp_avg_sq;
    float *grad_acc;         /* (rows, cols) contiguous accumulator, or NULL              */
    int rows, cols, grad_ld;
} chore_adam_entry;
int chore_adam_step(chore_handle *h, const chore_adam_entry *entries, int n, float lr, float beta1, float beta2,
                    float eps, int32_t *step, const float *gscale, float *loss_inout, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CHORE_B200_H */
