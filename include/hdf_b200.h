/* libhdf_b200.so -- C ABI of the B200-native H-DenseFormer 3D hot path.
 *
 * The reference (shijun18/H-DenseFormer) is pure Python over torch/ATen library kernels and has no
 * FFI of its own; each entry point below names the reference call it replaces
 * (file:line relative to the reference root).  Conventions:
 *   - every pointer is a DEVICE pointer unless the comment says "host";
 *   - activations are channels-last [N, D, H, W, C] with an explicit channel stride `ld*` (elements),
 *     so any operand may be a channel slice of a pre-allocated concat buffer (replaces torch.cat,
 *     models/HDenseFormer.py:230,247,249,251);
 *   - `dtype` selects the activation storage type (HDF_F32 exact path, HDF_BF16 fast path); parameters,
 *     statistics, token tensors and reductions are always fp32 (fp64 for cross-block sums);
 *   - kernels are launched asynchronously on `stream` (a cudaStream_t); nothing synchronises, nothing
 *     allocates: callers own every buffer, workspaces are sized by the *_workspace queries;
 *   - return value 0 = success, <0 = hdf_status; hdf_last_error_string() describes the last failure
 *     on the calling thread.  There is no CPU fallback anywhere.
 */
#ifndef HDF_B200_H
#define HDF_B200_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { HDF_OK = 0, HDF_ERR_ARG = -1, HDF_ERR_CUDA = -2, HDF_ERR_ARCH = -3, HDF_ERR_UNSUPPORTED = -4 } hdf_status;
typedef enum { HDF_F32 = 0, HDF_BF16 = 1 } hdf_dtype;

/* ---- library ---- */
int hdf_init(int device);                 /* checks compute capability 10.x, caches the SM count */
int hdf_version(void);
int hdf_sm_count(void);
unsigned long long hdf_launch_count(void);   /* kernels launched by this library so far (this process) */
const char* hdf_last_error_string(void);

/* ---- 3x3x3 convolutions (nn.Conv3d k3 p1: models/HDenseFormer.py:151,167; nn.ConvTranspose3d k3 s2 p1 op1:
 *      :211,215,219; their autograd backward).  mode 0: conv s1 p1; mode 1: transposed conv s2 (gather form);
 *      mode 2: conv s2 p1 (= input gradient of mode 1).  (Do,Ho,Wo) are OUTPUT dims; input dims follow from
 *      mode.  w_packed is [27][Cin][Cout] fp32 made by hdf_conv_pack_weights. ---- */
int hdf_conv_pack_weights(const float* w, float* packed, int A, int B, long long stride_a, long long stride_b, int flip,
                          void* stream); /* packed[tap][a][b] = w[a*stride_a + b*stride_b + (flip ? 26-tap : tap)] */
int hdf_conv3d_fwd(int dtype, int mode, const void* x, long long ldx, const float* w_packed, const float* bias, void* y,
                   long long ldy, int N, int Do, int Ho, int Wo, int Cin, int Cout, void* stream);
size_t hdf_conv3d_wgrad_workspace(int N, int Do, int Ho, int Wo, int Cin, int Cout);
/* dw[ci*stride_ci + co*stride_co + tap] (+)= sum_o x[in(o,tap)][ci] * dy[o][co]   (torch weight layouts) */
int hdf_conv3d_wgrad(int dtype, int mode, const void* x, long long ldx, const void* dy, long long ldy, float* dw,
                     long long stride_ci, long long stride_co, int N, int Do, int Ho, int Wo, int Cin, int Cout,
                     void* workspace, size_t ws_bytes, int accumulate, void* stream);

/* ---- tcgen05/TMEM tensor-core path for the same convolutions (bf16 activations, fp32 accumulate).
 *      Returns HDF_ERR_UNSUPPORTED for shapes it does not take; callers then use the SIMT entry above. ---- */
int hdf_tc_supported(int mode, int Cin, int Cout);
size_t hdf_tc_pack_bytes(int Cin, int Cout);
int hdf_tc_pack_weights(const float* w, void* packed_bf16, int Cin, int Cout, long long stride_ci, long long stride_co,
                        int flip, int cin_valid, void* stream); /* packed[tap][co][ci] bf16 (K-major B operand per tap);
                        channels ci >= cin_valid are zero (K padding for 2..4-channel inputs) */
int hdf_tc_conv3d_fwd(int mode, const void* x, long long ldx, const void* w_packed_bf16, const float* bias, void* y,
                      long long ldy, int N, int Do, int Ho, int Wo, int Cin, int Cout, void* stream);
/* weight-stationary variant for Cout = 32, Cin in {32, 64}, stride 1 (csrc/tc_conv_ws.cu): weights resident in TMEM as
 * the A operand, voxels on the MMA's N side, one TMA box per input plane shared by three output planes.  hdf_tc_conv3d_fwd
 * dispatches to it (HDF_TC_NO_WS=1 disables); exported so that tests can compare both kernels on the same operands. */
int hdf_tc_ws_supported(int mode, int Cin, int Cout);
int hdf_tc_ws_conv3d_fwd(const void* x, long long ldx, const void* w_packed_bf16, const float* bias, void* y, long long ldy,
                         int N, int D, int H, int W, int Cin, void* stream);
/* ... and with the InstanceNorm statistics of the output (mean, 1/sqrt(var + eps) per sample and channel, [N][32] fp32)
 * accumulated in the epilogue; workspace = hdf_tc_ws_stats_workspace(N) bytes, N <= 8 */
size_t hdf_tc_ws_stats_workspace(int N);
int hdf_tc_ws_conv3d_fwd_stats(const void* x, long long ldx, const void* w_packed_bf16, const float* bias, void* y, long long ldy,
                               int N, int D, int H, int W, int Cin, float eps, float* mean, float* rstd, void* workspace,
                               size_t ws_bytes, void* stream);
int hdf_instnorm_stats_finalize(const double* partial, int chunks, int N, int C, long long V, float eps, float* mean,
                                float* rstd, void* stream);
/* shift-major kernel for the big transposed convolution ConvTranspose3d(64 -> 32, k3, s2, p1, op1) = upconv_1
 * (reference models/HDenseFormer.py:215,249; csrc/tc_convt.cu): 8 shifted input boxes per tile instead of 27 tap boxes, the 8
 * output parity classes side by side in TMEM.  w is the torch weight [64][32][3][3][3] fp32; the packed buffer
 * (hdf_tc_convt_packed_bytes() bytes) is the kernel's swizzled shared-memory image.  x [N,D,H,W,64] bf16 -> y [N,2D,2H,2W,32]
 * bf16 (+ fp32 bias[32]).  HDF_TC_NO_CONVT=1 makes hdf_tc_convt_supported return 0 (callers fall back to mode 1 of
 * hdf_tc_conv3d_fwd). */
int hdf_tc_convt_supported(int Cin, int Cout);
size_t hdf_tc_convt_packed_bytes(void);
int hdf_tc_convt_pack_weights(const float* w, void* packed_bf16, void* stream);
int hdf_tc_convt_fwd(const void* x, long long ldx, const void* w_packed_bf16, const float* bias, void* y, long long ldy, int N,
                     int D, int H, int W, void* stream);
int hdf_tc_wgrad_supported(int mode, int Cin, int Cout);   /* mode 0 or 1 */
/* plane-ring weight gradient for stride-1 convs with a 32-channel operand (csrc/tc_wgrad_ws.cu): three w-shifted boxes of
 * the 32-channel operand stacked along M, three line-shifted sub-tiles of the other operand stacked along N, its planes in a
 * ring shared by three tiles.  hdf_tc_conv3d_wgrad dispatches to it (HDF_TC_NO_WGRAD_WS=1 disables). */
int hdf_tc_wgrad_ws_supported(int mode, int Cin, int Cout);
size_t hdf_tc_wgrad_ws_workspace(int N, int D, int H, int W, int Cin, int Cout);
int hdf_tc_wgrad_ws(const void* x, long long ldx, const void* dy, long long ldy, float* dw, long long stride_ci,
                    long long stride_co, int N, int D, int H, int W, int Cin, int Cout, void* workspace, size_t ws_bytes,
                    int accumulate, void* stream);
size_t hdf_tc_wgrad_workspace(int mode, int N, int Do, int Ho, int Wo, int Cin, int Cout);
int hdf_tc_conv3d_wgrad(int mode, const void* x, long long ldx, const void* dy, long long ldy, float* dw, long long stride_ci,
                        long long stride_co, int N, int Do, int Ho, int Wo, int Cin, int Cout, void* workspace,
                        size_t ws_bytes, int accumulate, void* stream);

/* ---- stem: the network's first BasicConv3d conv (models/HDenseFormer.py:198, Conv3d(in_channels -> nf, k3 p1, no bias),
 *      in_channels = 1..4).  The 27*Cin taps of every voxel are gathered once into a K-major bf16 matrix
 *      xcol [N*D*H*W, Kp] (Kp = hdf_stem_kp(Cin) = 27*Cin rounded up to 64, zero padded) straight from the caller's
 *      NCDHW fp32 batch; forward is then one tcgen05 GEMM over xcol, the weight gradient one over (xcol, dy). ---- */
int hdf_stem_kp(int Cin);                       /* 0 when unsupported */
int hdf_stem_supported(int Cin, int Cout);
int hdf_stem_im2col(const float* x_ncdhw, void* xcol_bf16, int N, int Cin, int D, int H, int W, void* stream);
int hdf_stem_pack_weights(const float* w, void* packed_bf16, int Cin, int Cout, void* stream); /* [Cout][Kp] bf16 from [Cout][Cin][27] */
int hdf_stem_conv_fwd(const void* xcol_bf16, const void* w_packed_bf16, void* y, long long ldy, int N, int D, int H, int W,
                      int Cin, int Cout, void* stream);
size_t hdf_stem_wgrad_workspace(int N, int D, int H, int W, int Cin, int Cout);
int hdf_stem_conv_wgrad(const void* xcol_bf16, const void* dy, long long ldy, float* dw, int N, int D, int H, int W, int Cin,
                        int Cout, void* workspace, size_t ws_bytes, int accumulate, void* stream);
/* the same first convolution WITHOUT the im2col matrix (csrc/stem_tc.cu): producer threads assemble the 128-voxel x Kp-tap
 * operand tile in shared memory straight from the fp32 NCDHW volume (gather -> bf16 -> swizzled UMMA layout); forward reads it
 * K-major against weights resident in shared memory, the weight gradient reads the same image MN-major against TMA-loaded dY
 * tiles.  x_ncdhw [N][Cin][D][H][W] fp32; y / dy [N*D*H*W, Cout] bf16 with row stride ldy; w_packed from
 * hdf_stem_pack_weights.  HDF_NO_STEM_FUSED=1 makes hdf_stem_fused_supported return 0. */
int hdf_stem_fused_supported(int Cin, int Cout);
int hdf_stem_fused_fwd(const float* x_ncdhw, const void* w_packed_bf16, void* y, long long ldy, int N, int Cin, int D, int H, int W,
                       int Cout, void* stream);
size_t hdf_stem_fused_wgrad_workspace(int Cin, int Cout);
int hdf_stem_fused_wgrad(const float* x_ncdhw, const void* dy, long long ldy, float* dw, int N, int Cin, int D, int H, int W, int Cout,
                         void* workspace, size_t ws_bytes, int accumulate, void* stream);

/* ---- patch embedding (nn.Conv3d k16 s16 + position_embeddings + Dropout: models/HDenseFormer.py:115-119,
 *      133-138).  img is the caller's NCDHW fp32 batch; tokens are [B*ntok, E] fp32 rows (ld = ldo). ---- */
size_t hdf_patch_embed_fwd_workspace(int B, int D, int H, int W, int E);
int hdf_patch_embed_fwd(const float* img, int B, int Mch, int modality, int D, int H, int W, const float* weight,
                        const float* bias, const float* pos, float* out, long long ldo, int E, float p,
                        const unsigned long long* seed_ptr, unsigned long long seed, unsigned call_id, void* workspace,
                        size_t ws_bytes, void* stream);
/* the same patch embedding as a tcgen05 implicit GEMM (csrc/patch_tc.cu): 128-token A tiles gathered from the fp32 volume into
 * the swizzled UMMA layout by producer threads, bf16 weights by TMA, fp32 accumulator in TMEM, K split 8 ways; bf16 operand
 * rounding, same arguments and output as hdf_patch_embed_fwd.  E in {64, 128, 256}. */
int hdf_patch_embed_tc_supported(int E);
size_t hdf_patch_embed_tc_workspace(int B, int D, int H, int W, int E);
int hdf_patch_embed_tc_fwd(const float* img, int B, int Mch, int modality, int D, int H, int W, const float* weight,
                           const float* bias, const float* pos, float* out, long long ldo, int E, float p,
                           const unsigned long long* seed_ptr, unsigned long long seed, unsigned call_id, void* workspace,
                           size_t ws_bytes, void* stream);
/* dropout masks are a pure function of (*seed_ptr + seed, call_id, element index): seed_ptr (nullable) is a device
 * counter the host advances once per forward, so captured CUDA graphs draw fresh masks on every replay */
size_t hdf_patch_embed_wgrad_workspace(int B, int D, int H, int W, int E);
int hdf_patch_embed_wgrad(const float* img, int B, int Mch, int modality, int D, int H, int W, const float* dtok,
                          long long ldd, float* dweight, int E, void* workspace, size_t ws_bytes, int accumulate,
                          void* stream);
int hdf_posemb_grad(const float* dtok, long long ld, float* dpos, int B, int ntok, int E, int accumulate, void* stream);

/* ---- token GEMMs (nn.Linear: models/HDenseFormer.py:37,40,57,60,85) with fused bias / exact GELU /
 *      dropout / residual epilogue.  C = epi(A[M,K] @ op(B)); op(B)=B^T for B [N,K] when b_is_nk. ---- */
int hdf_gemm_rowmajor(const float* A, long long lda, const float* Bm, long long ldb, int b_is_nk, float* C, long long ldc,
                      int M, int N, int K, const float* bias, const float* residual, long long ldr, float* pre, int act,
                      float p, const unsigned long long* seed_ptr, unsigned long long seed, unsigned call_id, int accumulate,
                      void* stream);
size_t hdf_gemm_at_b_workspace(int M, int N, int K);
int hdf_gemm_at_b(const float* A, long long lda, const float* Bm, long long ldb, float* C, int M, int N, int K,
                  void* workspace, size_t ws_bytes, int accumulate, void* stream); /* C[M,N] (+)= A[K,M]^T B[K,N] */
int hdf_act_dropout_bwd(const float* dy, long long ldd, const float* pre, float* dz, long long ldz, int M, int N, int act,
                        float p, const unsigned long long* seed_ptr, unsigned long long seed, unsigned call_id, void* stream);
int hdf_add_rows_f32(float* dst, long long ldd, const float* src, long long lds, long long rows, int C, int accumulate,
                     void* stream);

/* ---- LayerNorm (PreNorm, models/HDenseFormer.py:11-17) and attention (Dense_Attention, :47-75; heads of
 *      dim 4, qkv rows = [q | k | v], softmax never materialised) ---- */
int hdf_layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, float* out, long long ldo,
                      float* mean, float* rstd, int M, int C, float eps, void* stream);
size_t hdf_layernorm_bwd_workspace(int M, int C);
int hdf_layernorm_bwd(const float* dy, long long ldd, const float* x, long long ldx, const float* mean, const float* rstd,
                      const float* gamma, float* dx, long long ldo, int accumulate_dx, float* dgamma, float* dbeta,
                      int accumulate_params, int M, int C, void* workspace, size_t ws_bytes, void* stream);
int hdf_attention_fwd(const float* qkv, long long ld, float* o, long long ldo, float* lse, int B, int N, int H, float scale,
                      void* stream);
int hdf_attention_bwd(const float* qkv, long long ld, const float* o, long long ldo, const float* dout, long long lddo,
                      const float* lse, float* dqkv, long long ldg, int B, int N, int H, float scale, void* stream);

/* fused row-local chain of one DCT inner layer (models/HDenseFormer.py:95-98): attention out-projection + dropout +
 * residual, then the shared feed-forward applied twice (LN -> Linear -> GELU -> Dropout -> Linear -> Dropout) */
int hdf_dct_c_fwd(const float* o, const float* h0, float* h1, float* n2, float* z1, float* f1, float* h2, float* n3, float* z1b,
                  float* g1, float* m2, float* r2, float* m3, float* r3, float* fout, long long ldf, const float* Wo,
                  const float* bo, const float* gm, const float* bt, const float* W1, const float* b1, const float* W2,
                  const float* b2, int R, float p, const unsigned long long* seed_ptr, unsigned long long seed, unsigned ida,
                  unsigned idb, unsigned idc, unsigned idd, unsigned ide, void* stream);
int hdf_dct_a_fwd(const float* F, long long ldf, int Cl, const float* Wl, const float* bl, const float* gm, const float* bt,
                  const float* Wqkv, float* h0, float* n1, float* m1, float* r1, float* qkv, int R, void* stream);
/* bf16-path tensor-core forward of the same layer (csrc/tok_tc.cu): mma.sync bf16 Linears, tf32 m16n8k4 / m16n8k8 for
 * Q K^T / P V (head_dim 4 = the tf32 K atom), fp32 softmax statistics through warp shuffles.  hdf_tok_a_fwd has the operands
 * of hdf_dct_a_fwd; hdf_tok_c_fwd = attention (writes o, lse [B,8,N]) + the chain of hdf_dct_c_fwd, same saved tensors. */
int hdf_tok_a_fwd(const float* F, long long ldf, int Cl, const float* Wl, const float* bl, const float* gm, const float* bt,
                  const float* Wqkv, float* h0, float* n1, float* m1, float* r1, float* qkv, int R, void* stream);
int hdf_tok_c_fwd(const float* qkv, const float* h0, float* o, float* lse, float* h1, float* n2, float* z1, float* f1, float* h2,
                  float* n3, float* z1b, float* g1, float* m2, float* r2, float* m3, float* r3, float* fout, long long ldf,
                  const float* Wo, const float* bo, const float* gm, const float* bt, const float* W1, const float* b1,
                  const float* W2, const float* b2, int B, int N, float scale, float p, const unsigned long long* seed_ptr,
                  unsigned long long seed, unsigned ida, unsigned idb, unsigned idc, unsigned idd, unsigned ide, void* stream);
size_t hdf_dct_a_bwd_workspace(int R, int Cl);
int hdf_dct_a_bwd(const float* dqkv, const float* dh1, const float* h0, const float* n1, const float* m1, const float* r1,
                  const float* F, long long ldf, int Cl, const float* Wqkv, const float* gm, const float* Wl, float* dF,
                  long long lddf, float* dWqkv, float* dgm, float* dbt, float* dWl, float* dbl, int R, void* workspace,
                  size_t ws_bytes, void* stream);   /* head of the layer: Linear_l + LN1 + to_qkv, and its backward */
size_t hdf_dct_c_bwd_workspace(int R);
int hdf_dct_c_bwd(const float* dg2, long long ldg, const float* o, const float* h1, const float* n2, const float* z1,
                  const float* f1, const float* h2, const float* n3, const float* z1b, const float* g1, const float* m2,
                  const float* r2, const float* m3, const float* r3, const float* Wo, const float* gm, const float* W1,
                  const float* W2, float* d_o, float* dh1, float* dW2, float* db2, float* dW1, float* db1, float* dWo, float* dbo,
                  float* dgm, float* dbt, int R, float p, const unsigned long long* seed_ptr, unsigned long long seed,
                  unsigned ida, unsigned idb, unsigned idc, unsigned idd, unsigned ide, void* workspace, size_t ws_bytes,
                  void* stream);

/* ---- InstanceNorm3d (+affine) + ReLU (+ residual add) (BasicConv3d / UpConv: models/HDenseFormer.py:152-158,
 *      168-169; the "+ at3" adds at :238-244) ---- */
size_t hdf_reduce_workspace(int N, long long V, int C);
int hdf_instnorm_stats(int dtype, const void* y, long long ldy, int N, long long V, int C, float eps, float* mean,
                       float* rstd, void* workspace, size_t ws_bytes, void* stream);
int hdf_instnorm_apply(int dtype, const void* y, long long ldy, const float* mean, const float* rstd, const float* gamma,
                       const float* beta, const void* residual, long long ldr, void* out, long long ldo, int N,
                       long long V, int C, int relu, void* stream);
/* ... fused with the 1x1x1 head that consumes its output (reference models/HDenseFormer.py:253-255 after :152-158): out as
 * above (no residual) AND logits [N, ncls, V] bf16 = head_b + out . head_w^T from the same registers; bf16 only, ncls <= 4 */
int hdf_instnorm_apply_head_supported(int C, int ncls);
int hdf_instnorm_apply_head(const void* y, long long ldy, const float* mean, const float* rstd, const float* gamma, const float* beta,
                            void* out, long long ldo, int N, long long V, int C, int relu, const float* head_w, const float* head_b,
                            void* logits, int ncls, void* stream);
int hdf_instnorm_bwd(int dtype, const void* dout, long long ldd, const void* y, long long ldy, const float* mean,
                     const float* rstd, const float* gamma, const float* beta, void* dy, long long ldo, int N, long long V,
                     int C, int relu, float* s1, float* s2, float* dgamma, float* dbeta, int accumulate_params,
                     void* workspace, size_t ws_bytes, void* stream);
int hdf_colsum(int dtype, const void* x, long long ld, long long rows, int C, float* out, int accumulate, void* workspace,
               size_t ws_bytes, void* stream);

/* ---- pooling / upsampling / element-wise (nn.MaxPool3d(2,2): models/HDenseFormer.py:199,203,207;
 *      F.interpolate trilinear x2 align_corners=False: :174) ---- */
int hdf_maxpool2_fwd(int dtype, const void* x, long long ldx, void* out, long long ldo, int N, int Do, int Ho, int Wo,
                     int C, void* stream);
int hdf_maxpool2_bwd(int dtype, const void* x, long long ldx, const void* dpool, long long ldp, void* dx, long long lddx,
                     int N, int Do, int Ho, int Wo, int C, int accumulate, void* stream);
int hdf_upsample2_fwd(int dtype, const void* x, long long ldx, void* out, long long ldo, int N, int Di, int Hi, int Wi,
                      int C, void* stream);
int hdf_upsample2_bwd(int dtype, const void* dout, long long ldd, void* dx, long long lddx, int N, int Di, int Hi, int Wi,
                      int C, int accumulate, void* stream);
/* the same four entries with an explicit depth factor for FLAT volumes [N, 1, H, W, C] (the 2-D model,
 * reference models/HDenseFormer_2D.py: MaxPool2d(2), bilinear x2): pd / sd = 1 leaves the depth axis alone, 2 = the 3-D op */
int hdf_maxpool2_fwd_ex(int dtype, const void* x, long long ldx, void* out, long long ldo, int N, int Do, int Ho, int Wo,
                        int C, int pd, void* stream);
int hdf_maxpool2_bwd_ex(int dtype, const void* x, long long ldx, const void* dpool, long long ldp, void* dx, long long lddx,
                        int N, int Do, int Ho, int Wo, int C, int accumulate, int pd, void* stream);
int hdf_upsample2_fwd_ex(int dtype, const void* x, long long ldx, void* out, long long ldo, int N, int Di, int Hi, int Wi,
                         int C, int sd, void* stream);
int hdf_upsample2_bwd_ex(int dtype, const void* dout, long long ldd, void* dx, long long lddx, int N, int Di, int Hi, int Wi,
                         int C, int accumulate, int sd, void* stream);
int hdf_add_(int dtype, void* dst, long long ldd, const void* src, long long lds, long long rows, int C, void* stream);
int hdf_copy_rows(int dtype, void* dst, long long ldd, const void* src, long long lds, long long rows, int C, void* stream);
int hdf_cast_rows_from_f32(int dtype, const float* src, long long lds, void* dst, long long ldd, long long rows, int C,
                           void* stream);
int hdf_cast_rows_to_f32(int dtype, const void* src, long long lds, float* dst, long long ldd, long long rows, int C,
                         void* stream);
int hdf_ncdhw_to_cl(int dtype, const float* x, void* out, long long ldo, int N, int C, long long V, void* stream);

/* ---- 1x1x1 heads (nn.Conv3d k1: models/HDenseFormer.py:223-227,246-253); logits are NCDHW ---- */
int hdf_head_fwd(int dtype, const void* a, long long lda, const float* w, const float* b, void* out, int N, long long V,
                 int C, int ncls, void* stream);
size_t hdf_head_bwd_workspace(int N, long long V, int C, int ncls);
int hdf_head_bwd(int dtype, const void* g, const void* a, long long lda, const float* w, void* da, long long ldd,
                 float* dw, float* db, int N, long long V, int C, int ncls, int accumulate_da, int accumulate_params,
                 void* workspace, size_t ws_bytes, void* stream);

/* ---- loss (CEPlusDice / DeepSuperloss: loss/combine_loss.py:25-35,72-79; DiceLoss: loss/dice_loss.py:26-41,
 *      70-87; CrossentropyLoss: loss/cross_entropy.py:10-22).  One call per deep-supervision level. ---- */
size_t hdf_loss_sums_bytes(int B, int C);
int hdf_loss_level_fwd(int dtype, const void* logits, const float* target, const float* class_weight, int B, int C, int Dl,
                       int Hl, int Wl, int level_stride, int ignore_index, int has_ignore, float smooth, float level_weight,
                       float ce_weight, float dice_weight, double* sums, float* out_level /* [3]: loss, ce, dice */,
                       float* total, void* stream);
int hdf_loss_level_bwd(int dtype, const void* logits, const float* target, const float* class_weight, int B, int C, int Dl,
                       int Hl, int Wl, int level_stride, int ignore_index, int has_ignore, float smooth, float level_weight,
                       float ce_weight, float dice_weight, const double* sums, const float* grad_out, void* dlogits,
                       void* stream);
/* ... with a separate stride along depth for flat inputs (2-D model: logits [B, C, 1, H_l, W_l], target [B, C, 1, H, W]) */
int hdf_loss_level_fwd_ex(int dtype, const void* logits, const float* target, const float* class_weight, int B, int C, int Dl,
                          int Hl, int Wl, int level_stride, int depth_stride, int ignore_index, int has_ignore, float smooth,
                          float level_weight, float ce_weight, float dice_weight, double* sums, float* out_level, float* total,
                          void* stream);
int hdf_loss_level_bwd_ex(int dtype, const void* logits, const float* target, const float* class_weight, int B, int C, int Dl,
                          int Hl, int Wl, int level_stride, int depth_stride, int ignore_index, int has_ignore, float smooth,
                          float level_weight, float ce_weight, float dice_weight, const double* sums, const float* grad_out,
                          void* dlogits, void* stream);

/* ---- sliding-window inference (trainer.py:560-582; cal_steps :595-618 stays host code).
 *      steps_* are HOST arrays of window starts; the count map is analytic, never stored. ---- */
/* metric tail (trainer.py:382-398, 919-945; metrics.py:82-151): per-sample confusion counts of argmax(target) x
 * argmax(logits), accumulated into conf [B][C][C] (uint64, caller zeroes); no host synchronisation */
int hdf_confusion_update(int dtype, const void* logits, const float* target, int B, int C, long long V,
                         unsigned long long* conf, void* stream);
int hdf_sw_accumulate(int dtype, const void* logits, float* agg, int C, int X, int Y, int Z, int x0, int y0, int z0, int px,
                      int py, int pz, void* stream);
int hdf_sw_finalize(float* agg, long long* mask, int C, int X, int Y, int Z, const int* steps_x, int nx, const int* steps_y,
                    int ny, const int* steps_z, int nz, int patch_x, int patch_y, int patch_z, int normalise, void* stream);

/* ---- fused Adam / AdamW over the flat gradient arena (reference: trainer.py:793-840 optimizer selection with no weight
 *      decay for 1-D parameters and biases).  The table (one entry per parameter tensor) is packed on the host with
 *      hdf_adam_table_set and uploaded by the caller; chunks is a device int2 array (tensor index, chunk index), one CTA
 *      per hdf_adam_chunk() elements; hyper is a device float[2] {learning rate, step count} so that graph replays see
 *      scheduler updates; grad_scale multiplies the gradient (1/world for a summed all-reduce, 1 otherwise). ---- */
size_t hdf_adam_table_bytes(int ntensors);
int hdf_adam_chunk(void);
int hdf_adam_table_set(void* host_table, int index, float* param, long long offset, long long n, float weight_decay);
int hdf_adam_step(const void* table, const void* chunks, int nchunks, const float* grad_flat, float* m_flat, float* v_flat,
                  float* hyper, float beta1, float beta2, float eps, int adamw, float grad_scale, void* stream);

/* ---- GPU input pipeline (csrc/prep.cu): one sample of the reference's 3-D transform chain
 *   RandomCrop3D -> PETandCTNormalize | MRNormalize | Trunc_and_Normalize -> RandomTranslationRotationZoom3D ->
 *   RandomFlip3D -> To_Tensor   (reference data_utils/transformer_3d.py:7-169, data_utils/data_loader.py:16-68,126-159,
 *   composed at trainer.py:128-141), fused into a statistics pass over the crop window and one gather kernel.
 *   vol [M][Dv][Hv][Wv] fp32, lab [Dv][Hv][Wv] fp32 class ids (device); crop origin (d0,h0,w0) size (D,H,W);
 *   norm_mode 0 none | 1 PET/CT (p0 = CT window centre, p1 = half width; channel 1 z-scored over the crop) | 2 MR (divide
 *   by the channel maximum, negatives -> 0) | 3 truncate to [p0, p1] and scale to [0, 1];
 *   affine: DEVICE pointer to the 3 x 4 row-major double matrix (rows 0-2 of compose(T, R, Z)) or null = no warp; the warp
 *   is skimage.transform.warp(image, coords) = scipy map_coordinates(order 1, mode 'constant', cval 0), labels warped per
 *   class and thresholded at 0.5; flip_axis 0 none | 1 H | 2 W (applied after the warp);
 *   img_out [M][D][H][W] fp32; lab_out [num_class][D][H][W] fp32 one-hot, channel 0 = background, or null.
 *   M, num_class <= 8.  workspace: hdf_prep_workspace(M) bytes. */
size_t hdf_prep_workspace(int M);
int hdf_prep_sample(const float* vol, const float* lab, int M, int Dv, int Hv, int Wv, int d0, int h0, int w0, int D, int H, int W,
                    int norm_mode, float p0, float p1, const double* affine, int flip_axis, int num_class, float* img_out,
                    float* lab_out, void* workspace, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HDF_B200_H */
