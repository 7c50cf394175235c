/* mickey_b200 — C ABI of the B200-native MicKey inference hot path.
 *
 * The reference (nianticlabs/mickey) has no native boundary: its hot path is the Python method
 * MickeyRelativePose.forward (lib/models/MicKey/compute_pose.py:20-37).  This library is what a
 * maintainer binds instead of the three Python stages that method calls; each entry point names the
 * reference interface it replaces.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions: extern "C", plain pointers and sizes, int status (0 = OK, < 0 = error, message via
 * mk_last_error()), no exceptions cross the ABI.  Every pointer named *_dev is a DEVICE pointer owned
 * by the caller (PyTorch on the Python side); kernels are enqueued asynchronously on `stream`
 * (a cudaStream_t passed as void*).  One handle per device; a handle is not re-entrant.
 */
#ifndef MICKEY_B200_H
#define MICKEY_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mk_handle mk_handle;

/* Mirrors the keys of the reference config that the hot path reads
 * (config/MicKey/curriculum_learning.yaml:4-32,89-96; config/default.py). */
typedef struct mk_config {
  int embed_dim, depth, heads;        /* DINOv2 variant (dinov2.py:306-342): 384/12/6, 768/12/12, 1024/24/16 */
  int down_factor;                    /* MICKEY.DINOV2.DOWN_FACTOR (14) */
  int block_dims[4];                  /* MICKEY.KP_HEADS.BLOCKS_DIM (512,256,128,64) */
  int desc_dim;                       /* MICKEY.DSC_HEAD.LAST_DIM (128) */
  int use_softmax;                    /* MICKEY.KP_HEADS.USE_SOFTMAX */
  int depth_sigmoid;                  /* MICKEY.KP_HEADS.USE_DEPTHSIGMOID */
  float max_depth;                    /* MICKEY.KP_HEADS.MAX_DEPTH */
  int kp_pos_enc, dsc_pos_enc;        /* *.POS_ENCODING */
  int norm_dsc;                       /* MICKEY.DSC_HEAD.NORM_DSC */
  float temperature;                  /* FEATURE_MATCHER.DUAL_SOFTMAX.TEMPERATURE */
  int use_dustbin;                    /* FEATURE_MATCHER.DUAL_SOFTMAX.USE_DUSTBIN */
  int it_matches, it_ransac;          /* PROCRUSTES.IT_MATCHES / IT_RANSAC */
  int num_sampled, num_corr, num_refine;   /* NUM_SAMPLED_MATCHES / NUM_CORR_3D_3D / NUM_REFINEMENTS */
  float th_inlier, th_soft_inlier;    /* TH_INLIER / TH_SOFT_INLIER */
} mk_config;

/* ---- lifecycle (replaces MickeyRelativePose.__init__, compute_pose.py:9-18, and builder.py:5-20) ---- */
int mk_create(int device, const mk_config* cfg, mk_handle** out);
int mk_destroy(mk_handle* h);
const char* mk_last_error(void);
const char* mk_version(void);
/* sizeof(mk_config) / sizeof(mk_gemm_args) as compiled into the library (binding self-check). */
int mk_sizeof_config(void);
int mk_sizeof_gemm_args(void);

/* Register a packed weight / table tensor that lives in device memory (replaces load_state_dict,
 * builder.py:11-13; the packing itself — fp16 cast, BatchNorm folding, per-head stacking — is done by
 * mickey_b200/engine.py and documented in DESIGN.md).  dtype: 0 = fp32, 1 = fp16. */
int mk_set_tensor(mk_handle* h, const char* name, const void* ptr_dev, int dtype, long long numel);
/* Declare the image geometry the size-dependent tables ("patch.posb", "patch.clspos", "head.pe") were
 * built for, and check that every tensor the pipeline needs has been registered. */
int mk_finalize(mk_handle* h, int img_h, int img_w);

long long mk_workspace_bytes(mk_handle* h, int n_pairs, int img_h, int img_w);
/* Byte offset of a named intermediate buffer inside the workspace (debugging / stage-wise tests; names and
 * layouts are listed in DESIGN.md §3: "X", "F", "CAT", "Y4d", ...). */
long long mk_workspace_offset(mk_handle* h, const char* name, int n_pairs, int img_h, int img_w);

/* ---- stage 1: feature extraction for 2*n_pairs images
 * replaces MicKey_Extractor.forward (mickey_extractor.py:43-58) for image0 and image1 plus
 * get_abs_kpts_coordinates / prepare_kpts_dsc (compute_correspondences.py:20-43).
 * images_dev fp32 [2*n_pairs, 3, H, W] (all image0 first, then all image1), values in [0,1].
 * kps [2n,2,N] px coords, depth [2n,1,N], scr [2n,1,N], dsc [2n,128,N]  (N = (H/14)*(W/14)). */
int mk_extract(mk_handle* h, const float* images_dev, int n_pairs, int img_h, int img_w, float* kps_dev,
               float* depth_dev, float* scr_dev, float* dsc_dev, void* ws_dev, long long ws_bytes, void* stream);

/* Same stage fed by uint8 images straight from the decoder: images_u8_dev uint8 [2*n_pairs, H, W, 3],
 * RGB, HWC — what lib/datasets/utils.py:61-71 (cv2.imread -> cvtColor -> resize) holds before `.float() / 255`
 * (:74); the division, the crop to multiples of 14 (mickey_extractor.py:46) and the patch gather happen in one kernel,
 * bit-identical to mk_extract on the reference's float tensor. */
int mk_extract_u8(mk_handle* h, const unsigned char* images_u8_dev, int n_pairs, int img_h, int img_w, float* kps_dev,
                  float* depth_dev, float* scr_dev, float* dsc_dev, void* ws_dev, long long ws_bytes, void* stream);

/* ---- stage 2: dual-softmax matcher
 * replaces featureMatcher/dualSoftmax.forward (feature_matcher.py:48-83), kp_matrix_scores
 * (compute_correspondences.py:46-50) and `final_scores = scores * kp_scores` (compute_pose.py:23).
 * Uses the descriptors/scores left in the workspace by mk_extract.  Outputs fp32 [n_pairs, N, N];
 * scores_dev AND kp_scores_dev may both be NULL ("lean" mode: only final_scores, the one matrix the solver reads, is
 * materialised: 17 instead of 47 MB per 720x540 pair); the same holds for mk_forward / mk_forward_u8.
 * nn_pitch: row pitch of the three outputs in floats.  N (or 0) = the reference's contiguous [n_pairs, N, N].  A pitch
 * that is a multiple of 4 (e.g. N rounded up to 32: 1952 for N = 1938) makes every row 16-byte aligned, which lets the
 * outputs leave through TMA tensor stores as full 128-byte lines; a caller then views the buffers as
 * [n_pairs, N, nn_pitch][:, :, :N].  N = 1938 itself cannot be described by a tensor map (7752-byte rows). */
int mk_match(mk_handle* h, int n_pairs, float* scores_dev, float* kp_scores_dev, float* final_scores_dev, long long nn_pitch,
             void* ws_dev, long long ws_bytes, void* stream);

/* ---- stage 3: probabilistic Procrustes RANSAC
 * replaces e2eProbabilisticProcrustesSolver.estimate_pose_vectorized (probabilisticProcrustes.py:183-348).
 * final_scores_dev fp32 [n_pairs, N, N] with row pitch nn_pitch floats (N or 0 = contiguous).
 * K0/K1 fp32 [n_pairs,3,3].  pose_dev fp32 [n_pairs,13] = R row-major (9) | t (3) | soft inlier count (1).
 * outer_idx_dev int32 [n_pairs*IT_MATCHES, NUM_SAMPLED] / inner_idx_dev int32 [n_pairs*IT_MATCHES*IT_RANSAC, 3]:
 * when non-NULL they replace the two random draws (:231, :251) — the parity tests inject the reference's.
 * seed: non-zero = (re)seed the solver's counter-based generator; 0 = continue the device-side sequence
 * (the state lives in device memory and advances after every solve, so a captured CUDA graph of this call
 * draws fresh numbers on every replay).
 * Optional outputs (NULL to skip): best_set_dev int32 [n_pairs] (index into the IT_MATCHES sampled sets),
 * inlier_mask_dev fp32 [n_pairs, NUM_SAMPLED] (hard inliers of the winning set at the final pose),
 * sampled_idx_out_dev int32 [n_pairs*IT_MATCHES, NUM_SAMPLED] (the cells that were drawn),
 * hyp_scores_out_dev fp32 [n_pairs, IT_MATCHES*IT_RANSAC].  status_dev int32[1]: bit0 = not enough non-zero
 * cells, bit1 = candidate overflow (selection truncated), bit2 = non-finite hypothesis; any bit gives the reference's
 * zero pose (R = 0, t = 0, inliers = 0 for the whole batch, probabilisticProcrustes.py:331-342). */
int mk_solve_pose(mk_handle* h, const float* final_scores_dev, long long nn_pitch, const float* kps_dev, const float* depth_dev,
                  const float* K0_dev, const float* K1_dev, int n_pairs, int n_kpts, unsigned long long seed,
                  const int* outer_idx_dev, const int* inner_idx_dev, float* pose_dev, int* best_set_dev,
                  float* inlier_mask_dev, int* sampled_idx_out_dev, float* hyp_scores_out_dev, int* status_dev,
                  void* ws_dev, long long ws_bytes, void* stream);

/* ---- whole path: replaces MickeyRelativePose.forward (compute_pose.py:20-37) ---- */
int mk_forward(mk_handle* h, const float* images_dev, const float* K0_dev, const float* K1_dev, int n_pairs,
               int img_h, int img_w, unsigned long long seed, float* kps_dev, float* depth_dev, float* scr_dev,
               float* dsc_dev, float* scores_dev, float* kp_scores_dev, float* final_scores_dev, long long nn_pitch,
               float* pose_dev, int* best_set_dev, float* inlier_mask_dev, int* sampled_idx_out_dev, int* status_dev,
               void* ws_dev, long long ws_bytes, void* stream);

int mk_forward_u8(mk_handle* h, const unsigned char* images_u8_dev, const float* K0_dev, const float* K1_dev, int n_pairs,
                  int img_h, int img_w, unsigned long long seed, float* kps_dev, float* depth_dev, float* scr_dev,
                  float* dsc_dev, float* scores_dev, float* kp_scores_dev, float* final_scores_dev, long long nn_pitch,
                  float* pose_dev, int* best_set_dev, float* inlier_mask_dev, int* sampled_idx_out_dev, int* status_dev,
                  void* ws_dev, long long ws_bytes, void* stream);

/* ---- after the path: submission records (replaces the per-pair loop of submission.py:43-59)
 * pose_dev fp32 [n_pairs,13] as written by mk_forward / mk_solve_pose -> out_dev fp64 [n_pairs, 9] =
 * qw qx qy qz | tx ty tz | inliers | valid.  Quaternion = transforms3d.quaternions.mat2quat(R) (principal eigenvector
 * of the symmetric 4x4 K(R), w >= 0) computed in fp64; valid = 0 where the reference skips the frame
 * (np.isnan(R).any() or np.isnan(t).any() or np.isinf(t).any(), submission.py:51-52).  One D2H copy per batch. */
int mk_pose_to_submission(const float* pose_dev, int n_pairs, double* out_dev, void* stream);

/* (Re)seed the solver's device-side generator on `stream` (used in front of a CUDA-graph replay of mk_forward
 * captured with seed = 0). */
int mk_set_seed(mk_handle* h, unsigned long long seed, void* stream);

/* Number of kernel launches issued by this library since the handle was created (for bench.py). */
long long mk_launch_count(mk_handle* h);
/* Per-kernel-class device timing: while enabled every launch issued by the stages is bracketed by CUDA
 * events on the launch stream; mk_profile_read synchronises and writes "<class> <scopes> <total ms>" lines. */
int mk_profile_enable(mk_handle* h, int enable);
int mk_profile_read(mk_handle* h, char* buf, int buf_bytes);

/* ---- operator-level entry points (unit tests of single kernels; not needed by an integrator) ---- */
typedef struct mk_gemm_args {
  int epi;                 /* 0 STORE_H, 1 RESID_F, 2 PATCH, 3 CONV, 4 STORE_F, 5 LN, 6 LSE (row + column partials), 7 DUAL,
                              8 RESID_LN (RESID_F, then out_h = LayerNorm(out_f row) * aux + beta; N <= 1024) */
  int impl;                /* 0 default (tcgen05), 1 tcgen05, 2 SIMT debug kernel */
  const void* a; long long a_rows, a_cols, a_ld;
  const void* b; long long b_rows, b_cols, b_ld;
  int M, N, k_chunks, chunks_per_tap, num_taps;
  int tap_shift[9];
  int groups, a_row_group_off, a_col_group_off, a_col_base, b_row_group_off;
  int act;                 /* 0 none, 1 GELU(erf), 2 ReLU */
  const float* bias; int bias_group_off;
  const float* gamma; const float* beta; int ln_group_off;
  float* out_f; long long out_f_ld, out_f_group_off;
  void* out_h; long long out_h_ld, out_h_group_off;
  const void* res_h; long long res_h_ld, res_h_group_off;
  const float* aux; int aux_group_mask;
  int pad_h2, pad_w2, tok_per_img;
  float eps;
  int n_valid; float inv_temp;
  const float* dustbin;
  float* part_row; float* part_col; int part_ld;   /* LSE out: float2 (max, sum) partials, [groups][part_ld/64 | part_ld/32 slots][part_ld];
                                                      part_ld = n_valid rounded up to 128 */
  const float* lse_r; const float* lse_c;          /* DUAL in: log2-domain log-sum-exp per row / column, [groups, part_ld] (mk_op_matcher_reduce) */
  const float* scr0; const float* scr1;
  float* scores; float* kp_scores; float* final_scores;
  float lse_bound;         /* LSE: > 0 = every |A.B| <= lse_bound (normalised descriptors): fixed-shift partials; 0 = true maxima */
  long long out_pitch;     /* DUAL: row pitch of the outputs in floats (0 = n_valid); % 4 == 0 selects the TMA-store path */
} mk_gemm_args;

int mk_op_gemm(const mk_gemm_args* args, void* stream);
int mk_op_patch_gather(const float* img, void* patches_h, int n_img, int H, int W, int kpad, float* x_f,
                       const float* cls_pos, int D, void* stream);
int mk_op_ingest_u8(const unsigned char* img_u8, void* patches_h, int n_img, int H, int W, int kpad, float* x_f,
                    const float* cls_pos, int D, void* stream);
int mk_op_layernorm(const float* x, const float* w, const float* b, void* out_h, int rows, int D, float eps, int mode,
                    int gh, int gw, void* stream);
/* impl: 0 = default (tcgen05), 1 = tcgen05/TMEM kernel, 2 = mma.sync kernel (cross-check) */
int mk_op_attention(const void* qkv_h, void* out_h, int n_img, int T, int D, int heads, int impl, void* stream);
/* kv_part_f: scratch [n_img, G, ceil(h2*w2/32), 8, 272] fp32 */
int mk_op_linattn(const float* qkv_f, float* kv_part_f, float* kv_f, void* msg_h, int n_img, int G, int h2, int w2, float eps,
                  void* stream);
int mk_op_matcher_reduce(const float* part_row, const float* part_col, const float* dustbin, int B, int N, int part_ld,
                         float* lse_r, float* lse_c, void* stream);
int mk_op_sample(const float* final_scores, int B, int N, long long pitch, int IM, int n_sample, unsigned long long seed, void* ws,
                 long long ws_bytes, int* idx_out, int* status, void* stream);
long long mk_op_sample_workspace_bytes(int B, int IM);

#ifdef __cplusplus
}
#endif
#endif /* MICKEY_B200_H */
