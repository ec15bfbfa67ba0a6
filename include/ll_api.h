/*
 * ll_api.h -- C ABI of the B200-native learned-lifting / tree-entropy hot path.
 *
 * Plain `extern "C"` entry points, plain pointers and sizes, no torch / C++ types.
 * The reference (uberkk/ImageCompressionLearnedLiftingandLearnedTreeBasedModels) is
 * pure Python/PyTorch and has no FFI layer of its own; each entry point below cites
 * the reference function (file:line under /root/reference) whose arithmetic it
 * replaces.  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to fp32 unless stated; the caller (PyTorch)
 *    owns all buffers, outputs and scratch included;
 *  - all work is enqueued asynchronously on `stream` (a cudaStream_t passed as
 *    void*); no hidden synchronisation, re-entrant for distinct streams;
 *  - return value: LL_OK or a negative code; ll_last_error() gives a message for
 *    the calling thread;
 *  - strides are in ELEMENTS.
 */
#ifndef LL_API_H
#define LL_API_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LL_OK 0
#define LL_EINVAL (-1) /* bad shape / alignment / argument           */
#define LL_EARCH (-2)  /* current device is not sm_100 (B200)        */
#define LL_ECUDA (-3)  /* CUDA launch/runtime failure, see last error */

typedef void* ll_stream_t; /* cudaStream_t */

const char* ll_last_error(void);
int ll_version(void);
/* 0 if the current CUDA device can run this library (compute capability 10.x). */
int ll_check_device(void);
int ll_sm_count(void);

/* ------------------------------------------------------------------------- */
/* Learned lifting (K2)                                                       */
/* ------------------------------------------------------------------------- */

/* A (batch, y, x) strided view of fp32 data. */
typedef struct ll_view3 {
  float* ptr;
  int64_t sb, sy, sx;
} ll_view3;

/* One lifting-step job: dout = din + sign * (skip + res_weight * CNN(skip)),
 * skip = 3-tap pre-filter of `src` along y.  All three views are (nb, ny, nx). */
typedef struct ll_lift_job {
  ll_view3 src, din, dout;
  int32_t nb, ny, nx;
} ll_lift_job;

#define LL_LIFT_BLOB_FLOATS 53336

/* Packs one lifting step's parameters into the kernel's blob layout (device to
 * device, on `stream`).  pre_w: (3,) taps of convBlock[k] (lifting_dwt_nets.py:784-827);
 * w1..b4: P_block_v2 conv1..conv4 weight/bias in torch layout
 * (graphs/layers/P_block_v2.py:16-36).  blob: LL_LIFT_BLOB_FLOATS floats. */
int ll_pack_lift_step(const float* pre_w, const float* w1, const float* b1, const float* w2, const float* b2,
                      const float* w3, const float* b3, const float* w4, const float* b4, float* blob,
                      ll_stream_t stream);

/* One lifting step over up to 2 independent jobs in a single launch.  Replaces one
 * "skip = convBlock[k](src); dst = dst +- (skip + P|U[j](skip) * w)" group of
 * lifting_forward_row_2_stage_lifting (graphs/layers/wavelet_forward_v2.py:58-74) /
 * lifting_inverse_row_2_stage_lifting (graphs/layers/wavelet_inverse_v2.py:76-90),
 * including P_block_v2.forward (graphs/layers/P_block_v2.py:40-55).
 * sign = +1 forward, -1 inverse, 0 = write the raw CNN output (stand-alone P_block_v2.forward,
 * pass pre-filter taps (0,1,0)); linear != 0 drops both tanh (linearity_flag == 0).
 * precision (per call, nothing process-wide):
 *   LL_LIFT_FP32  every layer on the FP32 FMA pipe (exact, ~45 % of the FFMA2 peak);
 *   LL_LIFT_TC    conv2 / conv3 (94 % of the MACs) on tcgen05 with the 3xTF32 hi/lo split and FP32
 *                 accumulation in tensor memory (fp32-level accuracy, see lift_tc.cu), the rest FP32;
 *   LL_LIFT_TC16  the same kernel with conv2 / conv3 as a 3xFP16 split (kind::f16 at twice the TF32 rate; hi = fp16(v),
 *                 lo = fp16(v - hi) of operands pre-scaled by 2^8 so that the lo parts stay normal: 22 significand bits
 *                 per operand like 3xTF32, one FP32 accumulator, exact 2^-16 in the epilogue).  Needs bounded
 *                 activations: with linear != 0 (no tanh) the call runs the 3xTF32 kernel instead. */
#define LL_LIFT_FP32 0
#define LL_LIFT_TC 1
#define LL_LIFT_TC16 2
int ll_lift_step(const ll_lift_job* jobs, int njobs, const float* blob, float sign, float res_weight,
                 int linear, int precision, ll_stream_t stream);

/* One full 2-D lifting level, forward: x (B,h,w) -> ll (B,h/2,w/2), yh (B,3,h/2,w/2)
 * ordered LH,HL,HH.  Replaces wavelet_forward_v2.one_level_lifting
 * (graphs/layers/wavelet_forward_v2.py:26-54).  blobs[4]: packed steps 1..4.
 * scratch: ll_lift_level_scratch_floats(B,h,w) floats.  x_sb/ll_sb/yh_sb: batch
 * strides (elements) so that outputs can be slices of larger tensors; rows are dense.
 * scale: 0, or 1 with nh/nl device scalars (wavelet_forward_v2.py:76-80). */
size_t ll_lift_level_scratch_floats(int B, int h, int w);
int ll_lift_level_fwd(const float* x, int64_t x_sb, float* ll, int64_t ll_sb, float* yh, int64_t yh_sb,
                      float* scratch, int B, int h, int w, const float* const* blobs, float res_weight,
                      int linear, int scale, const float* nh, const float* nl, int precision, ll_stream_t stream);
/* Inverse level: (ll, yh) -> x.  Replaces wavelet_inverse_v2.one_level_lifting +
 * reconstruct_fun (graphs/layers/wavelet_inverse_v2.py:20-56, 68-92). */
int ll_lift_level_inv(const float* ll, int64_t ll_sb, const float* yh, int64_t yh_sb, float* x, int64_t x_sb,
                      float* scratch, int B, int h, int w, const float* const* blobs, float res_weight,
                      int linear, int scale, const float* nh, const float* nl, int precision, ll_stream_t stream);

/* ------------------------------------------------------------------------- */
/* CDF 9/7 fixed-filter DWT (K1)                                              */
/* ------------------------------------------------------------------------- */

/* One level of pytorch_wavelets.DWTForward(mode='periodization', wave='bior4.4')
 * as used by DWTPytorchWaveletsLayer.encode (graphs/layers/lifting_dwt_nets.py:228-231,250):
 * x (N,h,w) planes -> ll (N,h/2,w/2), yh (N,3,h/2,w/2) ordered LH,HL,HH.
 * N = B*C planes; *_sn are plane strides in elements, rows dense. */
int ll_dwt97_fwd_level(const float* x, int64_t x_sn, float* ll, int64_t ll_sn, float* yh, int64_t yh_sn,
                       int N, int h, int w, ll_stream_t stream);
/* One level of DWTInverse (lifting_dwt_nets.py:231,274): (ll, yh) -> x (N,h,w), h,w = output size. */
int ll_dwt97_inv_level(const float* ll, int64_t ll_sn, const float* yh, int64_t yh_sn, float* x, int64_t x_sn,
                       int N, int h, int w, ll_stream_t stream);

/* All J levels in one call (DWTForward(J,...)(x) / DWTInverse(...)((Yl, Yh)), lifting_dwt_nets.py:250,274):
 * x (N,h,w) dense planes; yl (N,h>>J,w>>J); yh[j] (N,3,h>>(j+1),w>>(j+1)), finest first;
 * scratch: ll_dwt97_scratch_floats(N,h,w,J) floats for the intermediate LL planes. */
size_t ll_dwt97_scratch_floats(int N, int h, int w, int J);
int ll_dwt97_fwd(const float* x, float* yl, float* const* yh, float* scratch, int N, int h, int w, int J,
                 ll_stream_t stream);
int ll_dwt97_inv(const float* yl, const float* const* yh, float* x, float* scratch, int N, int h, int w, int J,
                 ll_stream_t stream);

/* ------------------------------------------------------------------------- */
/* Pointwise subband auto-encoder (v1) fused with the quantiser               */
/* ------------------------------------------------------------------------- */

#define LL_AE1_BLOB_FLOATS 2212 /* per channel: w0[32] b0[32] W1[32x32] b1[32] W2[32x32] b2[32] w3[32] b3[1] pad[3] */

/* SubbandAutoEncoder.encode / .decode (graphs/layers/lifting_dwt_nets.py:99-125):
 * per channel c of x (B,C,n): 1 -> 32 tanh -> 32 tanh -> 32 tanh -> 1.
 * blob: C * LL_AE1_BLOB_FLOATS packed by ll_pack_ae1.  If q != NULL also writes
 * q = rint(y) (torch.round, half to even) -- the quantised symbols. */
int ll_pack_ae1(const float* w0, const float* b0, const float* w1, const float* b1, const float* w2,
                const float* b2, const float* w3, const float* b3, int C, int transposed, float* blob,
                ll_stream_t stream);
int ll_ae1_apply(const float* x, float* y, float* q, const float* blob, int B, int C, int64_t n,
                 ll_stream_t stream);

/* ------------------------------------------------------------------------- */
/* Context CNNs and rate kernels of the tree-based entropy models             */
/* ------------------------------------------------------------------------- */

/* fp32 direct convolution (cross-correlation, stride 1, zero padding K/2, K in {1,3,5}):
 * y[b, map(co)] = act(bias[co] + sum w[co, ci, u, v] * x[b, ci, ...]).  x (B,Cin,H,W) -- or
 * (B,Cin,H/2,W/2) when upsample2 != 0 (nearest 2x upsampling fused into the load; replaces
 * repeat_interleave(2,2).repeat_interleave(2,3), LiftingBasedDWT_net.py:348,367); w in torch
 * layout (Cout, Cin/groups, K, K) with any MaskedConv2d mask (graphs/layers/masked_conv2d.py:
 * 5-21) already multiplied in; lrelu != 0 applies LeakyReLU(0.01).  Output channel remap:
 * map(co) = (co / co_group) * co_stride + co_off + co % co_group (co_group <= 0: identity), which
 * turns torch.cat((plc0,csc0,plc1,csc1,plc2,csc2)) (:357-359) into a write pattern.
 * Replaces nn.Conv2d / MaskedConv2d forward of csc_xe, csc_list, plc_list, cgp_out_xo_list
 * (:269-318), onlyEZWT.plc_list (:789-794) and the ZTBlock dep_* CNNs (:618-680). */
int ll_conv2d(const float* x, int64_t x_sb, const float* w, const float* b, float* y, int64_t y_sb, int B, int Cin,
              int H, int W, int Cout, int K, int groups, int upsample2, int lrelu, int co_group, int co_stride,
              int co_off, ll_stream_t stream);

/* Pointwise tail of a ZTBlock dependency CNN in one pass: out = W3 . lrelu(W2 . lrelu(W1 . x + b1) + b2) + b3 per
 * pixel, LeakyReLU(0.01).  x (B,32,hw) fp32 with batch stride x_sb (elements) and channel stride hw; w1, w2 in torch
 * layout (32,32,1,1), w3 (1,32,1,1), b3 may be NULL; out (B,1,hw) with batch stride out_sb.  C must be 32.
 * Replaces the last three nn.Conv2d (1x1) + two nn.LeakyReLU of every dep_{1..4}_list_{mu,sigma}[n]
 * (graphs/models/LiftingBasedDWT_net.py:618-680, applied at :727-740). */
int ll_pw_mlp3(const float* x, int64_t x_sb, const float* w1, const float* b1, const float* w2, const float* b2,
               const float* w3, const float* b3, float* out, int64_t out_sb, int B, int C, int64_t hw,
               ll_stream_t stream);

/* ------------------------------------------------------------------------- */
/* Dense context CNNs on the tcgen05 tensor cores (K3)                        */
/* ------------------------------------------------------------------------- */

/* Small-Cin convs of the context models, fp32 FMA, result rounded to BF16 and written channels-last
 * for the tensor-core layers: the first conv of plc_list[i] (Conv2d(3,243,3,padding=1) + LeakyReLU on
 * the nearest-2x-upsampled quantised parent, LiftingBasedDWT_net.py:271,348,355; onlyEZWT :789-790)
 * and csc_list[i] (MaskedConv2d 'A' 5x5, 3->243, groups=3, :274-277,353).  x fp32 (B,Cin,H,W) -- or
 * (B,Cin,H/2,W/2) when upsample2 != 0; w (Cout,Cin/groups,K,K) with the mask multiplied in; only the
 * first live_taps taps (row-major) are evaluated (12 for mask 'A' 5x5, K*K for a dense conv).
 * out bf16 (B,H,W,out_cstride): the `region` channels from out_coff are all written -- destination
 * channel d holds conv channel (d / co_gstride) * co_group + d % co_gstride when d % co_gstride <
 * co_group, else 0 (co_group <= 0: identity).  Channel counts/offsets are multiples of 8. */
int ll_ctx_conv_nhwc(const float* x, const float* w, const float* bias, void* out, int B, int Cin, int H, int W,
                     int Cout, int K, int groups, int live_taps, int upsample2, int lrelu, int out_cstride,
                     int out_coff, int co_group, int co_gstride, int region, ll_stream_t stream);

/* cgp layers 2-4 + Gaussian rate in one launch (reference: graphs/models/LiftingBasedDWT_net.py:362-365, the last three
 * grouped 1x1 convs of cgp_out_xo_list[i] and GaussianConditional + -log2): h1 (B,H,W,Cin_total) bf16 NHWC -> layer 2 as a
 * grouped 1x1 tcgen05 GEMM (wp bf16 [groups][1][64][Kpad] from ll_pack_igemm_weight, koff as in ll_igemm_conv, bias2
 * (groups*C2), LeakyReLU) whose accumulator feeds layers 3 (w3 (groups*C3, C2), b3 (groups*C3), LeakyReLU) and 4 (w4
 * (2*groups, C3): row 2g = sigma, 2g+1 = mu; b4 (2*groups)) and the rate of x (B,groups,H,W fp32, batch stride x_sb;
 * noise: training-mode additive noise or NULL) in the epilogue -> bits (B,groups,H,W; batch stride bits_sb); the sum of
 * the bits is added to *sum_out when non-NULL.  C2 <= 64, C3 <= 20.  Same arithmetic and operation order as
 * ll_igemm_conv followed by ll_cgp_tail_rate, without the fp32 map in between. */
int ll_igemm_cgp_tail(const void* h1, const void* wp, const float* bias2, int B, int H, int W, int Cin_total, int Kpad, int C2,
                      int groups, const int* koff, const float* w3, const float* b3, const float* w4, const float* b4, int C3,
                      const float* x, int64_t x_sb, const float* noise, float* bits, int64_t bits_sb, double* sum_out,
                      ll_stream_t stream);

/* The two small-Cin context convs as tensor-core GEMMs (reference: graphs/models/LiftingBasedDWT_net.py:271,274-277,353-355):
 * con (B,3,H/2,W/2) quantised parent, q (B,3,H,W) quantised child -> out (B,H,W,320) bf16 = per pixel the im2col row of
 * the plc head (3x3 on the nearest-2x-upsampled parent, K = 27) and of the masked csc (12 live taps per group), each
 * value split hi + lo in bf16: channels [0,81) head as [hi | lo | hi], [81,128) zero, 128 + 64 g + [0,36) csc group g as
 * [hi | lo | hi], rest zero.  Consumed by ll_igemm_conv with 1-tap weights packed as [W_hi | W_hi | W_lo]. */
int ll_ctx_im2col(const float* con, const float* q, void* out, int B, int H, int W, ll_stream_t stream);

/* Weight packing for ll_igemm_conv: torch layout (Co,Ci,R,S) fp32 (taps = R*S in {1,9}) ->
 * bf16 [taps][Npad][Kpad], zero padded; Npad % 16 == 0 (<= 256), Kpad % 64 == 0. */
int ll_pack_igemm_weight(const float* w, void* wp, int Co, int Ci, int taps, int Npad, int Kpad, ll_stream_t stream);

/* fp32 NCHW (B,C,H,W; batch stride x_sb) -> channels [out_coff, out_coff+C) of a bf16 NHWC tensor
 * with out_cstride channels per pixel (drops csc_list[i]'s output into the cgp input layout). */
int ll_nchw_to_nhwc_bf16(const float* x, int64_t x_sb, void* out, int B, int C, int H, int W, int out_cstride,
                         int out_coff, ll_stream_t stream);

/* Implicit-GEMM convolution on tcgen05 / TMEM (BF16 operands, FP32 accumulation), stride 1,
 * zero padding: taps == 9 -> 3x3 (the 243->243 conv of plc_list[i][2], :272,355; onlyEZWT :791),
 * taps == 1 -> 1x1 (dense per-group layers of cgp_out_xo_list, :280-290; onlyEZWT's head :792-794).
 * x_nhwc bf16 (B,H,W,Cin_total).  `groups` independent GEMMs share the launch (<= 3): group g reads
 * Kpad/64 blocks of 64 input channels starting at channel koff[g*(Kpad/64) + k] (koff == NULL:
 * groups == 1 and blocks 0,64,...), its weights are wp[g] (bf16 [groups][taps][Npad][Kpad] from
 * ll_pack_igemm_weight), its bias is bias[g*Cout ...]; lrelu != 0 applies LeakyReLU(0.01).
 * Outputs (either or both):
 *   out_f32  fp32 NCHW, batch stride out_sb, channel map(g*Cout + c) with the ll_conv2d remap
 *            (co / co_group) * co_stride + co_off + co % co_group (co_group <= 0: identity);
 *   out_bf16 bf16 NHWC with out_cstride channels per pixel: group g writes ALL Npad channels at
 *            out_coff + g*out_gstride (channels >= Cout are exact zeros). */
int ll_igemm_conv(const void* x_nhwc, const void* wp, const float* bias, int B, int H, int W, int Cin_total, int Kpad,
                  int Npad, int Cout, int taps, int groups, const int* koff, int lrelu, float* out_f32, int64_t out_sb,
                  int co_group, int co_stride, int co_off, void* out_bf16, int out_cstride, int out_coff,
                  int out_gstride, ll_stream_t stream);

/* ---- 3xTF32 tensor-core chain for SubbandAutoEncoderBerk (lifting_dwt_nets.py:126-165): 3x3 convs and the 1x1
 * GDN / inverse-GDN norms with fp32-level accuracy (the network feeds the quantiser).  Activations travel
 * channels-last in fp32 as [hi | lo] TF32 halves (hi = rna_tf32(v), lo = rna_tf32(v - hi)); every product is
 * A_hi*B_hi + A_lo*B_hi + A_hi*B_lo with FP32 accumulation in tensor memory. ---- */

/* (Co,Ci,R,S) fp32 -> [taps][Npad][2*Kpad] fp32 ([hi | lo] along K).  transposed != 0: w is a ConvTranspose2d weight
 * (stored (in, out, R, S); stride 1, padding 1) and is flipped / swapped into the equivalent conv weight. */
int ll_pack_tf32_weight(const float* w, float* wp, int Co, int Ci, int taps, int Npad, int Kpad, int transposed,
                        ll_stream_t stream);
/* a_nhwc (B,H,W,2C) [hi | lo]; wp from ll_pack_tf32_weight with Kpad == C; bias[Cout].  epi 1: conv -> y (B,H,W,Cout)
 * raw output and sz (B,H,W,2*Cout) = split of y^2 (the GDN norm input); epi 2: GDN -> sz = split of
 * y * rsqrt(acc + bias) (inverse != 0: * sqrt), y is READ; epi 3: conv -> y only.  C, Cout multiples of 32.
 * pair != 0 (the product's choice): CTA pairs (clusters of 2, tcgen05 cta_group::2, M = 256) that fetch every k-block once
 * and share the weight tile; pair == 0: the single-CTA kernel (kept for A/B measurements).  Same results bit for bit. */
int ll_igemm_tf32(const float* a_nhwc, const float* wp, const float* bias, int B, int H, int W, int C, int Npad, int Cout,
                  int taps, int epi, int inverse, float* y, float* sz, int pair, ll_stream_t stream);
/* Fused conv (taps 1 | 9) + GDN / inverse GDN of SubbandAutoEncoderBerk (lifting_dwt_nets.py:140-148, graphs/layers/gdn.py:54-92)
 * on CTA pairs: sz (B,H,W,2N) = [hi | lo] split of y * rsqrt(beta + gamma . y^2) (inverse != 0: * sqrt), y = conv(a) + bias.
 * The conv output, its square and the norm stay in tensor memory.  wp: ll_pack_tf32_weight of the conv (Npad == N,
 * Kpad == C); gp: ll_pack_tf32_weight of the reparametrised gamma (N,N,1,1); beta: N reparametrised values.
 * N in {32, 64, 96, 192}; C a multiple of 32 up to 256. */
int ll_igemm_tf32_gdn(const float* a_nhwc, const float* wp, const float* bias, const float* gp, const float* beta, int B, int H,
                      int W, int C, int N, int taps, int inverse, float* sz, ll_stream_t stream);
/* First layer of SubbandAutoEncoderBerk (Conv2d(iC, N, 3, padding=1), iC <= 3; for ae_up the equivalent conv of the
 * ConvTranspose2d) fused with its GDN / inverse GDN: x (B,iC,H,W) fp32 NCHW, w0 (N,iC,3,3); the conv (K = 9 iC padded to
 * 32) runs as a 3xTF32 split on the tensor cores inside the same CTA-pair kernel -- the input windows are split and staged
 * in tensor memory by the kernel, w0 is split and packed by the kernel -- like the norm; sz as ll_igemm_tf32_gdn. */
int ll_conv3_gdn_head(const float* x, const float* w0, const float* bias, const float* gp, const float* beta, int B, int iC, int H,
                      int W, int N, int inverse, float* sz, ll_stream_t stream);
/* fp32 NCHW (B,C,H,W) -> y NHWC raw (optional) and sz NHWC (B,H,W,2C) = split of x^2 (mode 0) or of x (mode 1). */
int ll_nchw_to_nhwc_split(const float* x, float* y, float* sz, int B, int C, int H, int W, int mode, ll_stream_t stream);
/* z NHWC (B,H,W,2C) [hi | lo] -> fp32 NCHW (B,C,H,W) = hi + lo. */
int ll_nhwc_split_to_nchw(const float* z, float* out, int B, int C, int H, int W, ll_stream_t stream);
/* Last 3x3 conv of SubbandAutoEncoderBerk.ae_down / ae_up (nn.Conv2d / ConvTranspose2d(iC*32, iC, 3, padding=1),
 * lifting_dwt_nets.py:133,143) on the chain's own activations: z NHWC (B,H,W,2C) [hi | lo], w (Cout,C,3,3) in
 * cross-correlation layout (ConvTranspose weights flipped / transposed by the caller), bias (Cout) or NULL ->
 * out fp32 NCHW (B,Cout,H,W).  Exact fp32 FMA (feeds the quantiser).  C % 32 == 0, Cout in {1, 3}. */
int ll_nhwc_split_conv3(const float* z, const float* w, const float* bias, float* out, int B, int C, int Cout, int H, int W,
                        ll_stream_t stream);
/* 1x1 head of onlyEZWT's parent context net (LeakyReLU + nn.Conv2d(243, 6, 1), LiftingBasedDWT_net.py:792-794) on
 * the raw channels-last output of ll_igemm_tf32 (epi 3): out[b][o][pix] = bias[o] + sum_c w[o][c] * lrelu(y[pix][c]).
 * y (B*hw, Cpad) fp32 (channels >= C ignored), w (Cout, C), out fp32 NCHW (B, Cout, hw); exact fp32 FMA.
 * C <= 256, Cpad % 4 == 0, Cout <= 8. */
int ll_nhwc_lrelu_conv1(const float* y, const float* w, const float* bias, float* out, int B, int64_t hw, int C, int Cpad,
                        int Cout, int lrelu, ll_stream_t stream);

/* Tail of the cgp MLP fused with the rate: per group g (= child subband) and pixel,
 * h = LeakyReLU(W3[g] h2 + b3[g]) (C2 -> C3), (sigma, mu) = W4[g] h + b4[g], then exactly
 * ll_gauss_rate on x[:, g] (:286-290,361-365).  h2 fp32 (B, G*C2, hw) batch stride h2_sb; w3
 * (G*C3, C2), w4 (G*2, C3) in torch layout; ms_out optional (B, 2G, hw) dense; other arguments as
 * ll_gauss_rate.  C2 <= 64, C3 <= 32. */
int ll_cgp_tail_rate(const float* h2, int64_t h2_sb, const float* w3, const float* b3, const float* w4, const float* b4,
                     const float* x, int64_t x_sb, const float* noise, float* bits, int64_t bits_sb, float* y,
                     float* ms_out, int B, int G, int C2, int C3, int64_t hw, double* sum_out, ll_stream_t stream);

/* ------------------------------------------------------------------------- */
/* Agent-side pointwise work around the codec                */
/* ------------------------------------------------------------------------- */

/* agents/liftingDWT_agent.py:170-171 (also :104-105 in the training loop), clrch == 1:
 * y = compressai.transforms.RGB2YCbCr()(x) (BT.709, full range, chroma offset +0.5); y[:, 0] -= 0.5.
 * rgb, ycc: planar fp32 (B, 3, hw). */
int ll_rgb_to_ycbcr_shift(const float* rgb, float* ycc, int B, int64_t hw, ll_stream_t stream);

/* agents/liftingDWT_agent.py:174-181 + the MSE of TrainRDLoss.forward3 (graphs/losses/rate_dist.py:36):
 * xhat = clamp(YCbCr2RGB(yhat + (0.5, 0, 0)) - 0.5, -0.5, 0.5) written to xhat (optional), and
 * sse[b] += sum over the image of ((rgb_ref - 0.5) - xhat)^2 (optional, double, one entry per image; the caller
 * zeroes it): mse = sum(sse) / (3 B hw), PSNR = 10 log10(1 / mse) (:186).  Planar fp32 (B, 3, hw). */
int ll_ycbcr_to_rgb_sse(const float* ycc_hat, const float* rgb_ref, float* xhat, int B, int64_t hw, double* sse,
                        ll_stream_t stream);

/* ------------------------------------------------------------------------- */
/* Parallel entropy coding of the quantised subbands         */
/* ------------------------------------------------------------------------- */

/* Interleaved rANS (32-bit state, 16-bit words, 2^16 probability resolution), one GPU thread per stream.  Replaces,
 * for the entropy layers whose contexts depend on already decoded levels only (factorized :182-231, onlyEZWT
 * :759-840, ZTBlock :558-757), the serial per-coefficient Python coder the reference has for its autoregressive model
 * (compress_ar / decompress_ar, LiftingBasedDWT_net.py:458-556, around compressai.ans BufferedRansEncoder /
 * RansDecoder).  y (B, C, hw) holds the DEQUANTISED values the forward pass returns (round(x - mu) + mu); image b owns
 * S streams, stream s codes samples s, s+S, ... of the image.  mode 0: GaussianConditional, par = ms (B, 2C, hw) with
 * channel 2c = sigma, 2c+1 = mu; mode 1: EntropyBottleneck, par = blob of ll_pack_eb (C, 64); mode 2: the same
 * Gaussian discretised on the integer grid, y = round(x) (DWTConditioned2EntropyLayerZTBlock decodes and conditions
 * on plain rounding, :719-724,754).
 * ll_rans_encode writes stream st = b*S + s backwards into scratch[st*cap .. (st+1)*cap) (cap = ll_rans_stream_cap)
 * and its length in 16-bit words into counts[st]; ll_rans_pack gathers the streams at offsets[st] (exclusive scan of
 * counts, computed by the caller); ll_rans_decode reads the packed words and returns y bit for bit. */
int64_t ll_rans_stream_cap(int64_t n_per_image, int S);
int ll_rans_encode(int mode, const float* y, const float* par, int B, int C, int64_t hw, int S, uint16_t* scratch,
                   int32_t* counts, ll_stream_t stream);
int ll_rans_pack(const uint16_t* scratch, const int32_t* counts, const int64_t* offsets, int64_t nstreams, int cap,
                 uint16_t* packed, ll_stream_t stream);
int ll_rans_decode(int mode, const uint16_t* packed, const int64_t* offsets, const float* par, int B, int C, int64_t hw,
                   int S, float* y, ll_stream_t stream);

/* EntropyModel.quantize (compressai 1.2.1; call sites :330,341,352,719): q = round-half-even(x)
 * when noise == NULL ("dequantize"), else q = x + noise ("noise"; the caller draws U(-1/2,1/2)). */
int ll_quantize(const float* x, const float* noise, float* q, int64_t n, ll_stream_t stream);

/* GaussianConditional.forward(x, sigma, means=mu) followed by -log2 (:334-335,345-346,364-365,
 * 752-754,832-833).  x (B,C,hw) with batch stride x_sb; ms (B,2C,hw): channel 2c = sigma, 2c+1 = mu
 * (the reference's [:,0::2] / [:,1::2] split); noise NULL (eval: y = round(x-mu)+mu) or (B,C,hw)
 * dense (training: y = x + noise).  bits (B,C,hw) with batch stride bits_sb; y optional dense
 * output of the (de)quantised tensor; sum_out optional double accumulator += sum(bits)
 * (TrainRDLoss.forward3's reductions, graphs/losses/rate_dist.py:35-42). */
int ll_gauss_rate(const float* x, int64_t x_sb, const float* ms, int64_t ms_sb, const float* noise, float* bits,
                  int64_t bits_sb, float* y, int B, int C, int64_t hw, double* sum_out, ll_stream_t stream);

#define LL_EB_BLOB_FLOATS 64
/* EntropyBottleneck (compressai 1.2.1, filters (3,3,3,3)); params = 15 device pointers in
 * registration order: _matrix0,_bias0,_factor0, ..., _matrix3,_bias3,_factor3, _matrix4,_bias4,
 * quantiles.  blob: C * LL_EB_BLOB_FLOATS (softplus / tanh of the shape parameters precomputed). */
int ll_pack_eb(const float* const* params, int C, float* blob, ll_stream_t stream);
/* EntropyBottleneck.forward + -log2 (:204-210,689-690,800-801) on dense (B,C,hw):
 * y = round(x - median_c) + median_c (noise NULL) or x + noise; bits = -log2 max(p, 1e-9). */
int ll_eb_rate(const float* x, const float* noise, const float* blob, float* y, float* bits, int B, int C, int64_t hw,
               double* sum_out, ll_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LL_API_H */
