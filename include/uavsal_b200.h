/*
 * uavsal_b200.h — C ABI of the B200 (sm_100a) UAVSal inference hot path.
 *
 * The reference (zhangkao/IIP_UAVSal_Saliency) has no FFI layer: its "operator API" for this path is the
 * Python surface of model.py / model_feature.py / model_convlstm.py / utils_data.py / utils_score_torch.py,
 * whose arithmetic runs in ATen/cuDNN/cuBLAS.  This header is the boundary the new build introduces
 * underneath that surface.  Each entry point names the reference call site(s) whose
 * arithmetic it replaces.  Conventions:
 *
 *   - every function returns 0 on success, a positive cudaError_t, or a negative UAVSAL_E* argument error;
 *     uavsal_last_error() returns a static description of the most recent failure on this thread;
 *   - all pointers are DEVICE pointers unless named host_*; nothing is allocated, nothing synchronises;
 *     work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - "act" tensors are the arena format: NHWC rows (pixels) x channels, stored as TWO bf16 planes
 *     (hi = bf16(x), lo = bf16(x - hi)) so that tensor-core GEMMs can run the error-compensated product
 *     hi*hi + hi*lo + lo*hi with fp32 accumulation (DESIGN.md "Precision").  An act argument is the triple
 *         (const uint16_t* p, int64_t plane, int ld)
 *     p      = hi plane base (already offset to the first channel of a concat slot),
 *     plane  = element offset from the hi plane to the lo plane (0 = no lo plane: bf16x1 "fast" mode),
 *     ld     = row pitch in elements (multiple of 8; >= channel count for concat slots);
 *   - BatchNorm (eval) is folded into weights/bias by the host (model.py:70,95; eps 1e-5).
 */
#ifndef UAVSAL_B200_H
#define UAVSAL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UAVSAL_EINVAL   (-1)   /* bad argument (shape, alignment, null pointer) */
#define UAVSAL_ENOTSUP  (-2)   /* shape not supported by this kernel */
#define UAVSAL_EDRIVER  (-3)   /* driver entry point (cuTensorMapEncodeTiled) unavailable */

/* epilogue flags for uavsal_pw_gemm* / uavsal_conv3x3* */
#define UAVSAL_F_RELU6    1    /* clamp to [0,6]            (nn.ReLU6, model.py:71) */
#define UAVSAL_F_RESIDUAL 2    /* out += res                (model.py:101, 247) */
#define UAVSAL_F_SIGMOID  4    /* out = sigmoid(out)        (model.py:373) */
#define UAVSAL_F_RELU     16   /* max(x, 0)                 (nn.ReLU of the ResNet / VGG backbones, model_feature.py:72-128; exclusive with RELU6) */
#define UAVSAL_F_OUT_F32  8    /* uavsal_pw_gemm only: `out` is a float* to fp32 rows [m][out_ld] (out_plane ignored); used for the
                                  hidden tensor between a dwBlock's expand conv and its depthwise conv (model.py:90-92) */
#define UAVSAL_F_OUT_Q16  32   /* uavsal_pw_gemm only, with UAVSAL_F_RELU6: `out` is a uint16_t* to "q16" rows [m][out_ld] (out_plane ignored),
                                  q = rne(relu6(v) * 65535 / 6), read back as q * 6 / 65535 (|error| <= 4.6e-5): half the bytes of the fp32 rows
                                  for the widest hidden tensors (the 256 -> 1536 class of dwBlocks, model.py:90-92) */
#define UAVSAL_F_HID_Q16  64   /* uavsal_dw_project only: `hid` points to q16 rows (uint16_t, hid_ld in elements) instead of fp32 rows */
/* `terms` arguments of the tcgen05 entry points: 1 (bf16 x 1, "fast") or 3 (hi*hi + hi*lo + lo*hi, "exact"), optionally ORed with
   UAVSAL_TERMS_GEN1 to run the first-generation one-tile-per-CTA tcgen05 kernel instead of the persistent one (an independent
   cross-check engine for the tests; per call, so two callers on concurrent streams cannot disturb each other) */
#define UAVSAL_TERMS_GEN1 0x100
/* plane value marking an activation argument as plain fp32 rows (uavsal_dw3x3 input) instead of split-bf16 planes */
#define UAVSAL_PLANE_F32  (-1)
/* plane value marking an activation argument as q16 rows (uint16_t fixed point of a ReLU6 output, see UAVSAL_F_OUT_Q16) */
#define UAVSAL_PLANE_Q16  (-2)

int         uavsal_version(void);                 /* ABI version, currently 1 */
const char* uavsal_arch(void);                    /* "sm_100a" */
const char* uavsal_last_error(void);
int         uavsal_device_ok(int device);         /* 0 if `device` is compute capability 10.x */
int         uavsal_set_option(int key, int value);/* (process-global developer knobs for A/B timing; the product path never calls this)
                                                     key 1: tcgen05 GEMM kernel version (2 = persistent, default; 1 = one tile per CTA; per call: UAVSAL_TERMS_GEN1);
                                                     key 2: depthwise path (2 = TMA-staged, default; 1 = sliding rows; 0 = generic);
                                                     key 3: timing-ablation bits (dev only; results invalid when non-zero);
                                                     key 4: CTAs per cluster of the persistent GEMM (2 = cta_group::2 pair mode, default; 1);
                                                     key 5: cap on the GEMM's shared-memory pipeline depth (dev);
                                                     key 6: programmatic dependent launch of the product-path kernels (1 = on, default; 0);
                                                     key 7: ConvTWA step kernel (1 = resident-A shifted-view kernel, default; 0 = generic implicit GEMM) */

/* ---- weight preparation (BasicConv2d / dwBlock / rnn_conv parameters -> what the kernels below read) --------------------
 * conv weight w (cout, cin, taps) fp32 [taps = kh*kw: 1 or 9; cin = input channels per group], optional BatchNorm2d (eval:
 * model.py:69-70, 94-95, eps 1e-5) folded in fp32: w' = w * gamma / sqrt(var + eps), b' = beta + (conv_bias - mean) * gamma /
 * sqrt(var + eps); without BatchNorm b' = conv_bias (or 0).  Destination element (row r, k = tap * cin + ci); rows >= cout and
 * k >= taps * cin are zero padding.  gates > 1: rows g * (cout/gates) + c are emitted at c * gates + g (ConvLSTM gate conv,
 * model_convlstm.py:111-117).  Layouts:
 *   UAVSAL_W_ROWS_SPLIT  uint16 [2][n_pad][k_pad]: bf16 hi / lo planes, K-major rows (tcgen05 B operand: uavsal_pw_gemm,
 *                        uavsal_conv3x3, uavsal_twa_sequence, uavsal_convlstm_sequence, uavsal_dw_project, uavsal_expand_dw3x3)
 *   UAVSAL_W_ROWS_F32    float [n_pad][k_pad]
 *   UAVSAL_W_COLS_F32    float [k_pad][n_pad] (depthwise [9][C], stem [27][32], the SIMT cross-check engine)
 * out_bias: float [n_pad] or NULL. */
#define UAVSAL_W_ROWS_SPLIT 0
#define UAVSAL_W_ROWS_F32   1
#define UAVSAL_W_COLS_F32   2
int uavsal_pack_weights(const float* w, int cout, int cin, int taps, const float* bn_weight, const float* bn_bias,
                        const float* bn_mean, const float* bn_var, float bn_eps, const float* conv_bias, int gates,
                        int layout, int n_pad, int k_pad, void* out_w, float* out_bias, void* stream);

/* ---- layout conversion at the module boundary (torch NCHW fp32 <-> arena) ---------------------- */
/* NCHW fp32 -> act NHWC with channels zero-padded to cpad (cb priors, Demo_Test.py:16,22; states).  */
int uavsal_pack_nchw_f32(const float* src, int n, int c, int h, int w,
                         uint16_t* dst, int64_t plane, int ld, int cpad, void* stream);
/* act NHWC -> NCHW fp32 (returned maps and hidden state, model.py:370-375). */
int uavsal_unpack_nchw_f32(const uint16_t* src, int64_t plane, int ld, int n, int c, int h, int w,
                           float* dst, void* stream);

/* ---- a1 + K1: normalize_data (utils_data.py:43-65) fused with the stem conv 3x3 s2 + BN + ReLU6
 *      (torchvision features[0], used at model_feature.py:63).
 *      x_kind: 0 = fp32 NCHW already normalised (UAVSal.forward input, model.py:341)
 *              1 = uint8 NCHW raw RGB (Demo_Test.py:70,77: normalisation fused)
 *              2 = uint8 NHWC raw RGB (preprocess_videos output, utils_data.py:255-287)
 *      w: [3][3][3][32] (ky,kx,ci,co) fp32 BN-folded; bias[32]. */
int uavsal_stem_conv3x3s2(const void* x, int x_kind, int n, int h, int w,
                          const float* wgt, const float* bias,
                          uint16_t* out, int64_t out_plane, int out_ld, void* stream);

/* Same operation with the folded weights [27][32] and bias [32] given as HOST pointers: they are copied into the kernel's
 * parameter block at launch (3.5 KB, read through the constant bank), so nothing has to stay alive after the call returns. */
int uavsal_stem_conv3x3s2_hw(const void* x, int x_kind, int n, int h, int w, const float* wgt_host, const float* bias_host,
                             uint16_t* out, int64_t out_plane, int out_ld, void* stream);

/* ---- K2: depthwise 3x3 + BN + ReLU6 (model.py:92 BasicConv2d(groups=hidden); torchvision InvertedResidual dw)
 *      stride 1|2, dilation >= 1, padding = dilation.  wgt: [9][c] fp32 BN-folded, bias[c].
 *      in_plane == UAVSAL_PLANE_F32 / UAVSAL_PLANE_Q16: `in` points to plain fp32 / q16 rows [n*h*w][in_ld] (the hidden tensor written by
 *      uavsal_pw_gemm with UAVSAL_F_OUT_F32 / UAVSAL_F_OUT_Q16); accepted for dilation 1, and for dilation > 1 (stride 1) on maps small
 *      enough for two whole images of a 64-channel block to fit shared memory (h*w <= 799 for q16, 399 for fp32), else UAVSAL_ENOTSUP. */
int uavsal_dw3x3(const uint16_t* in, int64_t in_plane, int in_ld, int n, int h, int w, int c,
                 int stride, int dilation, const float* wgt, const float* bias, int relu6,
                 uint16_t* out, int64_t out_plane, int out_ld, void* stream);

/* dwBlock conv[0] + conv[1] fused (model.py:90-92; torchvision InvertedResidual): 1x1 expand + BN + ReLU6 followed by the
 * depthwise 3x3 (pad 1, stride 1|2) + BN + ReLU6, without materialising the expanded tensor.  cin <= 32.
 * w1: bf16 planes [2][ceil(hidden/64)*64][kp] (K-major, kp = cin rounded up to 16, zero padded), b1: [ceil(hidden/64)*64]
 * fp32; wd [9][hidden], bd [hidden] as uavsal_dw3x3.  out = (n, ho, wo, hidden) split-bf16. */
int uavsal_expand_dw3x3(const uint16_t* x, int64_t x_plane, int x_ld, int n, int h, int w, int cin,
                        const uint16_t* w1, int kp, const float* b1, int hidden, int stride,
                        const float* wd, const float* bd,
                        uint16_t* out, int64_t out_plane, int out_ld, void* stream);

/* dwBlock conv[1] + conv[2] + conv[3] fused (model.py:92-101): depthwise 3x3 (pad 1, stride 1) + BN + ReLU6 on the fp32 hidden
 * tensor `hid` (n, h, w, hidden) [rows of hid_ld floats], immediately consumed as the A operand of the 1x1 project conv + BN
 * (+ residual) on the tensor cores - the depthwise output never reaches HBM.  hidden % 128 == 0, cout % 64 == 0, cout <= 256;
 * also hidden == 32 with cout == 16 and no residual (torchvision features[1]): an fp32 FFMA kernel, `terms` ignored.
 * wd [9][hidden], bd [hidden] as uavsal_dw3x3; wgt/kpad/bias/res/out as uavsal_pw_gemm (flags: UAVSAL_F_RESIDUAL, and
 * UAVSAL_F_HID_Q16 when `hid` holds q16 rows - tensor-core kernel only). */
int uavsal_dw_project(const float* hid, int hid_ld, int n, int h, int w, int hidden,
                      const float* wd, const float* bd,
                      const uint16_t* wgt, int kpad, int cout, const float* bias, int flags, int terms,
                      const uint16_t* res, int64_t res_plane, int res_ld,
                      uint16_t* out, int64_t out_plane, int out_ld, void* stream);

/* Whole dwBlock / InvertedResidual (model.py:74-103; stride 1, dilation 1) in one launch: 1x1 expand + BN + ReLU6 -> depthwise 3x3
 * + BN + ReLU6 -> 1x1 project + BN (+ residual x); the 6x hidden tensor lives in TMEM / shared memory only.
 * x act (n, h, w, cin), cin % 8 == 0, cin <= 64; w1: bf16 planes [2][hidden][kp1] (K-major, kp1 >= cin), b1 [hidden];
 * hidden % 64 == 0 (a block whose hidden width is not a multiple of 64 is passed with its weights zero-padded: padded hidden
 * channels are exactly 0 through both ReLU6s and meet zero project weights); wd [9][hidden], bd [hidden] as uavsal_dw3x3;
 * w2: bf16 planes [2][cout16][hidden], b2 [cout16] with cout16 = cout rounded up to 16 (rows >= cout zero), cout % 8 == 0,
 * cout <= 64; flags: UAVSAL_F_RESIDUAL only; res / out as uavsal_pw_gemm.  Results equal uavsal_pw_gemm(F_OUT_F32) ->
 * uavsal_dw3x3 -> uavsal_pw_gemm bit for bit. */
int uavsal_mbconv_fused(const uint16_t* x, int64_t x_plane, int x_ld, int n, int h, int w, int cin,
                        const uint16_t* w1, int kp1, const float* b1, int hidden,
                        const float* wd, const float* bd,
                        const uint16_t* w2, int cout, const float* b2, int flags, int terms,
                        const uint16_t* res, int64_t res_plane, int res_ld,
                        uint16_t* out, int64_t out_plane, int out_ld, void* stream);

/* The 32 -> 16 block (torchvision features[1]) with its weights as HOST arrays: wd [9][32], bd [32], wp [16][32] (cout, hidden)
 * fp32 and bias [16] are copied into the kernel's parameter block at launch and read as constant-bank operands. */
int uavsal_dw_project32_hw(const float* hid, int hid_ld, int n, int h, int w, const float* wd_host, const float* bd_host,
                           const float* wp_host, const float* bias_host, uint16_t* out, int64_t out_plane, int out_ld, void* stream);

/* ---- K3: pointwise 1x1 conv + BN (+ReLU6)(+residual)(+sigmoid) (model.py:89,94,120-128,181,184,230).
 *      out[m][n0..] = act( sum_k A[m][k] * W[n][k] + bias[n] ) (+ res[m][n])
 *      tcgen05 version: wgt = bf16 planes [2][n][kpad] (hi, lo), kpad % 8 == 0; terms = 1 (bf16x1) or 3.
 *      SIMT version:    wgt_f32 = [k][n] fp32. */
int uavsal_pw_gemm(const uint16_t* a, int64_t a_plane, int a_ld, int m, int k,
                   const uint16_t* wgt, int kpad, int n, const float* bias, int flags, int terms,
                   const uint16_t* res, int64_t res_plane, int res_ld,
                   uint16_t* out, int64_t out_plane, int out_ld, void* stream);
int uavsal_pw_gemm_simt(const uint16_t* a, int64_t a_plane, int a_ld, int m, int k,
                        const float* wgt_f32, int n, const float* bias, int flags,
                        const uint16_t* res, int64_t res_plane, int res_ld,
                        uint16_t* out, int64_t out_plane, int out_ld, void* stream);

/* ---- K4: dense 3x3 conv, pad 1, stride 1 + BN + ReLU6 (conv_last, model.py:131,156) as implicit GEMM.
 *      in: act (n,h,w,c) with c % 64 == 0 for the tcgen05 version.
 *      tcgen05: wgt = bf16 planes [2][cout][9*c], k = (ky*3+kx)*c + ci.   SIMT: wgt_f32 = [9*c][cout]. */
int uavsal_conv3x3(const uint16_t* in, int64_t in_plane, int in_ld, int n, int h, int w, int c,
                   const uint16_t* wgt, int cout, const float* bias, int flags, int terms,
                   uint16_t* out, int64_t out_plane, int out_ld, void* stream);
int uavsal_conv3x3_simt(const uint16_t* in, int64_t in_plane, int in_ld, int n, int h, int w, int c,
                        const float* wgt_f32, int cout, const float* bias, int flags,
                        uint16_t* out, int64_t out_plane, int out_ld, void* stream);

/* ---- K5: F.interpolate(bilinear, align_corners=True) written into a concat slot (model.py:152-153,360)
 *      with the context prior's repeat(T,1,1,1) folded in: output frame i reads source frame i % n_src
 *      (model.py:361, quirk Q3).  src_group / dst_group > 0 batch several reference calls in one launch: output frame
 *      i = g*dst_group + j reads source frame g*src_group + j % src_group (0, 0 = one call). */
int uavsal_bilinear_ac(const uint16_t* in, int64_t in_plane, int in_ld, int n_src, int hs, int ws, int c,
                       uint16_t* out, int64_t out_plane, int out_ld, int n_dst, int hd, int wd,
                       int src_group, int dst_group, void* stream);

/* ---- K6: teConv_sub neighbour differences over the call batch (model.py:194-200): x1 (n,hw,c) -> (n,hw,2c).
 *      group > 0: the n frames are consecutive reference calls of `group` frames (mirrored edges at each call boundary). */
int uavsal_tdiff_cat(const uint16_t* in, int64_t in_plane, int in_ld, int n, int hw, int c,
                     uint16_t* out, int64_t out_plane, int out_ld, int group, void* stream);

/* ---- K7: context prior T-sum: x.view(B,T,C,H,W).sum(1) (model.py:357-358): (b*t,hw,c) -> (b,hw,c) */
int uavsal_ctx_sum(const uint16_t* in, int64_t in_plane, int in_ld, int b, int t, int hw, int c,
                   uint16_t* out, int64_t out_plane, int out_ld, void* stream);

/* elementwise out = a + b (STBlock fu_type='sum', model.py:241) */
int uavsal_add(const uint16_t* a, int64_t a_plane, int a_ld, const uint16_t* b, int64_t b_plane, int b_ld,
               int64_t rows, int c, uint16_t* out, int64_t out_plane, int out_ld, void* stream);

/* ---- K8: ConvTWA sequence, batch 1 (model_convlstm.py:276-292 cell, 364-377 loop):
 *      for t: i = sigmoid(conv3x3([x_t, h])); h = i*x_t + (1-i)*h; seq_out[t] = h.
 *      x: act (t_steps, h, w, c); h0: act (1,h,w,c); seq_out: act (t_steps,h,w,c) (last frame = h_last).
 *      tcgen05: wgt = bf16 planes [2][c][9*2c] (input-channel order [x, h], model_convlstm.py:279). */
int uavsal_twa_sequence(const uint16_t* x, int64_t x_plane, int x_ld,
                        const uint16_t* h0, int64_t h0_plane, int h0_ld,
                        int t_steps, int h, int w, int c,
                        const uint16_t* wgt, const float* wgt_f32, int terms, float* gx_workspace,
                        uint16_t* seq_out, int64_t seq_plane, int seq_ld, int batch, void* sync_workspace, void* stream);
/*      batch independent sequences advance together (x, seq: batch*t_steps images, sequence-major; h0: batch images).
 *      gx_workspace (optional, batch*t_steps*h*w*c floats): when given, the input half W_x*x_t of the gate conv is hoisted out
 *      of the recurrence into one batched implicit GEMM and only W_h*h_{t-1} (K = 9c) runs per step.
 *      sync_workspace (optional, with gx_workspace; uavsal_twa_sync_bytes(batch, h, w) bytes of device memory, 4-byte aligned,
 *      zeroed by the call): when given and one step's grid fits the SMs, the whole `for t` loop (model_convlstm.py:364-377) is
 *      ONE launch whose CTAs hand h_{t-1} tiles to their neighbours through per-tile counters; otherwise one launch per step.
 *      Same results bit for bit. */
size_t uavsal_twa_sync_bytes(int batch, int h, int w);

/* ---- K9: ConvLSTM sequence, one layer, batch_first (model_convlstm.py:111-126 cell, 199-212 loop).
 *      x: act (b, t, h, w, cin) ; h state act (b,h,w,ch) updated in place through seq_out; c state fp32
 *      (b,h,w,ch) updated in place.  Gate order i,f,o,g; packed weights interleave the four gates per channel:
 *      packed row 4*ch_idx + g  <-  reference row g*ch + ch_idx.  bias (4*ch, same interleave) may be NULL.
 *      seq_out: act (b, t, h, w, ch). */
int uavsal_convlstm_sequence(const uint16_t* x, int64_t x_plane, int x_ld,
                             const uint16_t* h0, int64_t h0_plane, int h0_ld, float* c_state,
                             int b, int t_steps, int h, int w, int cin, int ch,
                             const uint16_t* wgt, const float* wgt_f32, const float* bias, int terms,
                             uint16_t* seq_out, int64_t seq_plane, int seq_ld, void* stream);

/* ---- K10b: readout fused: depthwise 3x3 + BN + ReLU6 of the fp32 hidden tensor (conv_out_st.conv.1) folded straight into
 *      the 1-output project conv + BN + sigmoid (conv_out_st.conv.2/.3, model.py:372-373): the 1536-channel depthwise
 *      output is never written.  partial_ws: n*h*w*ceil(c/64) floats of workspace. */
int uavsal_dw3x3_dot_sigmoid(const float* in, int in_ld, int n, int h, int w, int c, const float* wd, const float* bd,
                             const float* wproj, float bias_proj, float* partial_ws, float* out_f32, void* stream);
/*      the same on q16 rows (the hidden tensor written by uavsal_pw_gemm with UAVSAL_F_OUT_Q16), in_ld in elements */
int uavsal_dw3x3_dot_sigmoid_q16(const uint16_t* in, int in_ld, int n, int h, int w, int c, const float* wd, const float* bd,
                                 const float* wproj, float bias_proj, float* partial_ws, float* out_f32, void* stream);

/* ---- K10a: readout project conv 1536->1 + BN + sigmoid (conv_out_st.conv.2/.3 + model.py:373):
 *      out_f32[row] = sigmoid( dot(A[row][:k], wgt[:k]) + bias ). */
int uavsal_dot_sigmoid(const uint16_t* a, int64_t a_plane, int a_ld, int64_t rows, int k,
                       const float* wgt, float bias, float* out_f32, void* stream);

/* ---- K10b / a11: postprocess_predictions + np2mat (utils_data.py:289-303, 68-82):
 *      letterbox-inverse bilinear resize (cv2.resize INTER_LINEAR semantics) of each (hs,ws) map to
 *      (hd,wd), divide by the per-frame max, *255, round-half-even, uint8.  frame_max: scratch (n floats). */
int uavsal_post_u8(const float* maps, int n, int hs, int ws, int hd, int wd, float* frame_max,
                   uint8_t* out_u8, void* stream);

/* postprocess_predictions alone (utils_data.py:289-303): same resize and /max*255, float output, no uint8 cast. */
int uavsal_post_f32(const float* maps, int n, int hs, int ws, int hd, int wd, float* frame_max,
                    float* out_f32, void* stream);

/* ---- K11: utils_score_torch.metric_cc / metric_nss / metric_kl / metric_sim (180-218, helpers 20-50).
 *      pred (n,1,h,w), truth (n,2,h,w) (ch0 density, ch1 fixations), fp32 or uint8-valued (dtype 0=f32, 1=u8).
 *      out (n,4) fp32 columns CC, NSS, KLD, SIM.  scratch: n*16 doubles. */
int uavsal_metrics4(const void* pred, const void* truth, int dtype, int n, int h, int w,
                    double* scratch, float* out, void* stream);

/* ---- video front-end after decode: utils_data.padding (321-343) as called by preprocess_videos (255-287).
 *      src (n, sh, sw, 3) uint8 frames as cv2.VideoCapture delivers them (BGR); dst (n, dh, dw, 3) uint8: the frame resized
 *      with cv2.resize's 8-bit INTER_LINEAR arithmetic (bit-exact, incl. its exact-2x INTER_AREA shortcut) keeping the aspect
 *      ratio, centred on a zero canvas; swap_rb = 1 also applies the BGR -> RGB reorder of :270.  dw <= 4096. */
int uavsal_letterbox_u8(const uint8_t* src, int n, int sh, int sw, uint8_t* dst, int dh, int dw, int swap_rb, void* stream);

/* ---- alternative backbones (model_feature.ReResNet / ReVGG, model_feature.py:72-128; torchvision resnet.py / vgg.py) -------------
 * uavsal_conv_first: the first conv from the raw frame, 3 -> 64 channels: k = 7, stride 2, pad 3 (ResNet conv1 + bn1 + relu) or k = 3,
 *   stride 1, pad 1 (VGG features.0 + relu).  x / x_kind as uavsal_stem_conv3x3s2 (uint8 kinds normalise as utils_data.normalize_data);
 *   wgt fp32 [k*k*3][64] (UAVSAL_W_COLS_F32 of uavsal_pack_weights), bias [64] or NULL; flags: UAVSAL_F_RELU.
 * uavsal_maxpool: nn.MaxPool2d(k, stride, pad), floor mode (ResNet 3/2/1, VGG 2/2/0); k = 1, stride 2 subsamples rows (in front of a
 *   stride-2 1x1 conv; behind a stride-1 evaluation of a stride-2 3x3 conv).
 * uavsal_add_act: out = a + b, then ReLU when flags has UAVSAL_F_RELU (the tail of a ResNet block: relu(conv(x) + identity)). */
int uavsal_conv_first(const void* x, int x_kind, int n, int h, int w, int k, int stride, const float* wgt, const float* bias, int flags,
                      uint16_t* out, int64_t out_plane, int out_ld, void* stream);
int uavsal_maxpool(const uint16_t* in, int64_t in_plane, int in_ld, int n, int h, int w, int c, int k, int stride, int pad,
                   uint16_t* out, int64_t out_plane, int out_ld, void* stream);
int uavsal_add_act(const uint16_t* a, int64_t a_plane, int a_ld, const uint16_t* b, int64_t b_plane, int b_ld, int64_t rows, int c,
                   int flags, uint16_t* out, int64_t out_plane, int out_ld, void* stream);

/* ---- utils_score_torch.metric_auc_j (53-88), deterministic part: S = min-max normalised pred (fp32, as :83), fixations =
 *      truth channel 1 > 0.5; out[n] = AUC-Judd (NaN when the map has no positive value or the frame no fixation, :54).
 *      The reference's optional jitter (:82, global torch generator) is added to `pred` by the caller.  pred (n,1,h,w),
 *      truth (n,2,h,w) fp32.  Frames with up to 4096 fixations are sorted in shared memory; the reference has no cap
 *      (:53-74), so maps of more than 4096 pixels must come with `workspace` (16-byte aligned device memory of at least
 *      uavsal_auc_judd_workspace(n,h,w) bytes, contents irrelevant) in which denser frames are sorted instead; without it
 *      the call is refused (UAVSAL_EINVAL) rather than answering NaN for such a frame. */
int64_t uavsal_auc_judd_workspace(int n, int h, int w);
int uavsal_auc_judd(const float* pred, const float* truth, int n, int h, int w, float* out, void* workspace,
                    int64_t workspace_bytes, void* stream);

/* ---- auc_b (91-120) / auc_s (135-159): sampled AUC with thresholds k*step below the largest sample.  The random pixel
 *      indices are the CALLER's draw (the reference uses the global numpy generator, :103 / :143-144):
 *      rand_idx int32 (n, max_k, n_rep) = flat pixel index of sample k of repetition rep; n_k[n] = samples per repetition
 *      (Borji: n_fix; shuffled: min(n_fix, n_ind)) and the false-positive denominator.  n_rep <= 128, step >= 1/62.
 *      out[n] fp32 (NaN as above, or when n_k is 0). */
int uavsal_auc_sampled(const float* pred, const float* truth, int n, int h, int w, const int32_t* rand_idx,
                       const int32_t* n_k, int max_k, int n_rep, double step, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UAVSAL_B200_H */
