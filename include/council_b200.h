/*
 * council_b200.h -- C ABI of libcouncil_b200.so: the sm_100a kernels behind the Council-GAN
 * training step (dis_update / dis_council_update / gen_update).
 *
 * The reference (Onr/Council-GAN) has no FFI: every device op on this path is a PyTorch library
 * call made from networks.py / trainer_council.py.  Each entry point below replaces one of those
 * call sites (cited per function, paths relative to the reference tree) so that a maintainer can
 * bind it with ctypes from the same Python code.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - all tensors are dense fp32; activations are channels-last, stacked over the council:
 *         act[G][B][H][W][C]      G = council members in the launch ("groups"), C % 4 == 0
 *     a source with `x_groups == 1` is shared (broadcast) by all G members;
 *   - convolution weights are stacked OHWI:  w[G][Cout][KH][KW][Cin]   (reference: OIHW per member);
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never allocates,
 *     and returns 0 on success or a negative cg_status; cg_last_error() describes the last failure
 *     of the calling thread;
 *   - `ws` / `ws_bytes` is caller-provided scratch; if it is too small the call fails with
 *     CG_ERR_WORKSPACE and cg_last_error() reports the size needed.
 */
#ifndef COUNCIL_B200_H
#define COUNCIL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    CG_OK = 0,
    CG_ERR_ARG = -1,        /* unsupported shape / argument */
    CG_ERR_WORKSPACE = -2,  /* workspace too small */
    CG_ERR_CUDA = -3,       /* CUDA runtime / driver error */
    CG_ERR_NO_DEVICE = -4   /* no sm_100 device */
} cg_status;

/* activation codes (Conv2dBlock activations, networks.py:494-507) */
enum { CG_ACT_NONE = 0, CG_ACT_RELU = 1, CG_ACT_LRELU = 2, CG_ACT_TANH = 3 };

/* Geometry of one grouped convolution.  Spatial sizes are those of the STORED input tensor;
 * with `ups` the convolution sees nearest-upsample-x2 of it (nn.Upsample(scale_factor=2),
 * networks.py:385, folded into the gather).  Ho/Wo are the output sizes. */
typedef struct {
    int32_t G, x_groups;         /* groups in w / y; groups in x (1 = shared input) */
    int32_t B, H, W, Cin;        /* stored input  [x_groups][B][H][W][Cin]  */
    int32_t Ho, Wo, Cout;        /* output        [G][B][Ho][Wo][Cout]      */
    int32_t KH, KW, stride, pad; /* zero padding (nn.ZeroPad2d, networks.py:473-474) */
    int32_t ups;                 /* 0/1 */
} cg_conv_geom;

const char* cg_last_error(void);
/* Library / device introspection: returns the SM count of the current device (>0) or a cg_status. */
int cg_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* 0 = SIMT fp32 kernels only; 1 (default) = tcgen05 TF32 tensor-core kernels wherever a layer qualifies;
 * other values are a bit mask for bring-up: 2 = data gradient only, 4 = weight gradient only, ...
 * (internally 1 forward | 2 dgrad | 4 wgrad).  Measurement switches on top of the mask: 8 = no CTA-pair (cta_group::2)
 * kernels, 16 = none for 128-wide tiles, 32 = CTA pairs for 64-wide tiles too, 64 = CTA pairs in the weight
 * gradient too (both measured slower), 1<<16 = one weight-gradient CTA per SM, 1<<18 = one x-on-M weight-gradient CTA per SM and 1<<17 =
 * one forward CTA per SM for tiles <= 64 channels wide (default: two co-resident CTAs interleave their MMAs), 1<<19 = image-side layers (<= 8 input lanes)
 * on the older TMA-im2col / explicit-patch paths instead of the shared-memory patch builders (csrc/conv_img.cu), 1<<20 = accumulator-layout epilogue
 * stores instead of the coalescing shared-memory patch (default: patch for tiles <= 64 channels wide; 1<<21 = for the wide tiles too),
 * 1<<22 = programmatic dependent launch between this library's kernels (every kernel carries the griddepcontrol pair; measured -4 % on the
 * launch-bound 128x128 configuration and +2 % at 256x256 batch 8, so the trainer turns it on for small maps only),
 * 1<<24 = programmatic dependent launch for the helper kernels only (those without dynamic shared memory),
 * 1<<23 = keep the widest N tile on small maps (default: narrower tiles when a launch has fewer tiles than SMs), bits 8..15 = cap on the CTA pairs launched.
 * The switches are per calling thread (like cg_last_error), not process-global.  Returns the previous mask. */
int cg_set_tensor_core_mode(int mode);
/* number of kernels launched by this library since load (bench.py reports it as gpu_launches) */
uint64_t cg_launch_count(void);
/* TMA descriptors are cached by (device pointer, geometry): hits / misses since load (bench.py reports them) */
void cg_tensor_map_cache_stats(uint64_t* hits, uint64_t* misses);

/* ---- convolution (Conv2dBlock conv + bias + activation, networks.py:513-520; nn.Linear with
 *      H=W=KH=KW=1, networks.py:531,563) ------------------------------------------------------- */
/* y = act(conv(x, w) + bias).  bias may be NULL. */
int cg_conv_fwd(const cg_conv_geom* g, const float* x, const float* w, const float* bias, float* y,
                int act, float slope, void* ws, size_t ws_bytes, void* stream);
/* y = conv(x, w) (no bias, no activation) and the instance-norm statistics of y in one call (Conv2d followed by
 * InstanceNorm2d / AdaIN, networks.py:516-518): mean, rstd [G][B][Cout].  On the tensor path the per-channel sums come
 * out of the convolution epilogue, saving a pass over y. */
int cg_conv_fwd_stats(const cg_conv_geom* g, const float* x, const float* w, float* y, float* mean, float* rstd,
                      float eps, void* ws, size_t ws_bytes, void* stream);
size_t cg_conv_fwd_stats_workspace_bytes(const cg_conv_geom* g);
/* dx = (conv_transpose(dy, w) [+ addend]) * act'(mask_src)      (autograd of the call above)
 * dx has the STORED input shape; with g->ups the 2x2 upsample fan-in is summed.
 * addend / mask_src (same shape as dx) may be NULL; act' = mask_src > 0 ? 1 : mask_slope. */
int cg_conv_dgrad(const cg_conv_geom* g, const float* dy, const float* w, float* dx,
                  const float* addend, const float* mask_src, float mask_slope,
                  void* ws, size_t ws_bytes, void* stream);
/* dw[G][Cout][KH][KW][Cin] = sum over pixels dy (x) im2col(x);  db[G][Cout] = sum dy (NULL: skip).
 * Deterministic (fixed split-K order). */
int cg_conv_wgrad(const cg_conv_geom* g, const float* x, const float* dy, float* dw, float* db,
                  void* ws, size_t ws_bytes, void* stream);
size_t cg_conv_workspace_bytes(const cg_conv_geom* g, int which /*0 fwd, 1 dgrad, 2 wgrad*/);

/* ---- instance norm / AdaIN (nn.InstanceNorm2d networks.py:483; AdaptiveInstanceNorm2d :640-653) */
/* mean, rstd [G][B][C] over H*W (biased variance, rstd = 1/sqrt(var+eps)). */
int cg_in_stats(const float* y, float* mean, float* rstd, int G, int B, int HW, int C, float eps,
                void* ws, size_t ws_bytes, void* stream);
/* z = act(gamma * (y-mean)*rstd + beta) [+ res];  AdaIN parameters come straight from the MLP
 * output adain[G][B][P]: beta = adain[.., off : off+C], gamma = adain[.., off+C : off+2C]
 * (assign_adain_params, networks.py:303-312).  adain == NULL: plain instance norm.
 * ups: z is written nearest-upsampled x2 ([G][B][2H][2W][C]).  res (shape of y) may be NULL. */
int cg_norm_act_fwd(const float* y, const float* mean, const float* rstd, const float* adain, int P,
                    int off, const float* res, float* z, int G, int B, int H, int W, int C, int act,
                    int ups, void* stream);
/* backward of the above w.r.t. y and the AdaIN parameters (d_adain[..][off:off+2C] is overwritten).
 * dz has the shape of z (upsampled if ups).  The residual branch gradient is dz itself. */
int cg_norm_act_bwd(const float* dz, const float* y, const float* mean, const float* rstd,
                    const float* adain, int P, int off, float* dy, float* d_adain, int G, int B,
                    int H, int W, int C, int act, int ups, void* ws, size_t ws_bytes, void* stream);

/* Single-launch forms of the three calls above (csrc/norm_coop.cu): statistics + normalise in ONE kernel, and the whole
 * backward in ONE kernel; a launch keeps only as many instances in flight as fit in L2, so the second pass over y (and dz)
 * is served from L2 and HBM sees y once.  Same results as cg_in_stats + cg_norm_act_fwd / cg_norm_act_bwd; mean / rstd
 * [G][B][C] are also written by the forward (the backward needs them).  ws >= cg_norm_fused_workspace_bytes(). */
int cg_norm_fused_fwd(const float* y, const float* adain, int P, int off, const float* res, float* z, float* mean,
                      float* rstd, int G, int B, int H, int W, int C, int act, int ups, float eps, void* ws,
                      size_t ws_bytes, void* stream);
int cg_norm_fused_bwd(const float* dz, const float* y, const float* mean, const float* rstd, const float* adain, int P,
                      int off, float* dy, float* d_adain, int G, int B, int H, int W, int C, int act, int ups, void* ws,
                      size_t ws_bytes, void* stream);
size_t cg_norm_fused_workspace_bytes(int G, int B, int C);

/* backward of nn.Upsample(scale_factor=2) (networks.py:385): dx[N][H][W][C] = 2x2 fan-in sum of d_up[N][2H][2W][C] */
int cg_upsample2x_bwd(const float* d_up, float* dx, int N, int H, int W, int C, void* stream);

/* ---- attention-mask head (Decoder_V2_atten.forward networks.py:398-407) ----------------------- */
/* h[G][B][HW][12] = tanh output of dec.model.9; x_in[B][HW][4] (shared); outputs padded to 4 ch. */
int cg_mask_head_fwd(const float* h, const float* x_in, float* x_fake, float* mask, int G, int B,
                     int HW, void* stream);
/* dh_pre[G][B][HW][12] = gradient w.r.t. the PRE-tanh output of dec.model.9. d_mask may be NULL. */
int cg_mask_head_bwd(const float* h, const float* x_in, const float* d_xfake, const float* d_mask,
                     float* dh_pre, int G, int B, int HW, void* stream);

/* The decoder tail of a no-grad pass in ONE launch (csrc/head_fused.cu): y[G][B][HW][64] = raw output of the last 3x3 block;
 * z = relu(AdaIN(y)) -> 1x1 64->64 + relu (w1, b1) -> 1x1 64->64 + relu (w2, b2) -> 1x1 64->12 + tanh (w3, b3) -> mask
 * compositing with x_in (cg_mask_head_fwd).  Weights [G][Cout][64], HW % 128 == 0.  Replaces cg_norm_act_fwd + 3 x cg_conv_fwd +
 * cg_mask_head_fwd (Decoder_V2_atten, networks.py:391-407) when nothing has to be kept for a backward pass. */
int cg_head_fused(const float* y, const float* mean, const float* rstd, const float* adain, int P, int off, const float* w1,
                  const float* b1, const float* w2, const float* b2, const float* w3, const float* b3, const float* x_in,
                  float* x_fake, float* mask, int G, int B, int HW, void* stream);

/* ---- image-space helpers ---------------------------------------------------------------------- */
/* nn.AvgPool2d(3, 2, padding=1, count_include_pad=False), networks.py:32,129 */
int cg_avgpool_fwd(const float* x, float* y, int N, int H, int W, int C, void* stream);
/* dx[N][H][W][Cx] (first nch lanes) = or += avgpool^T(dy[N][H/2][W/2][Cy] first nch lanes) */
int cg_avgpool_bwd(const float* dy, float* dx, int N, int H, int W, int Cy, int Cx, int nch,
                   int accumulate, void* stream);
/* dst[n][..][0:nch] += src[n][..][0:nch]  (pixel counts equal) */
int cg_acc_slice(float* dst, const float* src, long npix, int Cd, int Cs, int nch, void* stream);
/* y[g][n] = slot(idx[g*Bt+n]) (4 lanes) ++ x_in[n % B] (4 lanes, if x_in != NULL -> 8-lane output), where
 * slot(k) = k < n0 ? pool0[k] : pool1[k - n0]  (pool1 may be NULL when every index is < n0):
 * builds discriminator minibatches without materialising torch.cat((fake, real)) (networks.py:56-64) or
 * torch.cat((x, x_input), 1) (networks.py:152). */
int cg_gather_images(const float* pool0, int n0, const float* pool1, const int32_t* idx, const float* x_in, float* y,
                     int G, int Bt, int B, int HW, void* stream);
/* NCHW [N][C][HW] <-> channels-last [N][HW][Cp] (Cp >= C, pad lanes zeroed) */
int cg_nchw_to_nhwc(const float* x, float* y, int N, int C, int HW, int Cp, void* stream);
int cg_nhwc_to_nchw(const float* x, float* y, int N, int C, int HW, int Cp, void* stream);

/* ---- losses ----------------------------------------------------------------------------------- */
/* LSGAN (networks.py:64,90,166,194).  out[G][nseg][n_per_seg]; sums[G][nseg] = sum (out-target[seg])^2;
 * loss[g] (+)= sum_seg weights[g][seg] * mean_seg((out-target[seg])^2).  targets: device [nseg]; weights: device [G][nseg]. */
int cg_lsgan_fwd(const float* out, const float* targets, const float* weights, float* sums, float* loss,
                 int G, int nseg, int n_per_seg, int accumulate, void* stream);
/* dout = coef[g][seg] * (out - target[seg])   (coef device pointer, [G][nseg]) */
int cg_lsgan_bwd(const float* out, const float* targets, const float* coef, float* dout, int G,
                 int nseg, int n_per_seg, void* stream);
/* focus-loss sums over a mask [G][B][H][W][4] (3 live lanes), trainer_council.py:230-250:
 * sums[G][4] = { sum 1/(|m-c|+eps), sum m, sum |m[h+1]-m[h]|, sum |m[w+1]-m[w]| } */
int cg_focus_fwd(const float* mask, float* sums, int G, int B, int H, int W, float center, float eps,
                 void* ws, size_t ws_bytes, void* stream);
/* dmask = coef[g][0]*d(zero_one term) + coef[g][1]*d(sum m) + coef[g][2]*d(TV sums) */
int cg_focus_bwd(const float* mask, const float* coef, float* dmask, int G, int B, int H, int W,
                 float center, float eps, void* stream);

/* ---- fused losses (one launch per discriminator update and direction; two per gen_update and direction) ----------
 * These replace the four calls above on the training path: the loss of EVERY discriminator scale, the focus terms, the
 * loss-history matching and all their gradients, with no host round trip (trainer_council.py:518-524,576-586 run on the
 * device).  `ws`: >= cg_loss_workspace_bytes() bytes, ZERO-INITIALISED once by the caller, not shared between streams;
 * every call leaves its first 16 bytes (a ticket counter) zero again. */
#define CG_LOSS_MAX_MAPS 4
#define CG_LOSS_MAX_G 8
#define CG_LOSS_MAX_SEG 8
typedef struct {
    int32_t nmaps, G, nseg, _pad;             /* patch maps (discriminator scales), members, segments per member */
    const float* out[CG_LOSS_MAX_MAPS];       /* out[m][G][nseg][n_per_seg[m]]                                      */
    float* dout[CG_LOSS_MAX_MAPS];            /* gradient of the same shape (NULL: not wanted)                      */
    int32_t n_per_seg[CG_LOSS_MAX_MAPS];
    float target[CG_LOSS_MAX_SEG];            /* 0 for generated, 1 for real / less-style segments                  */
    float weight[CG_LOSS_MAX_G][CG_LOSS_MAX_SEG];
    float loss_scale, grad_scale;             /* 1/world under data parallelism                                     */
} cg_lsgan_desc;
/* loss_total[g] (+)= loss_scale * sum_m sum_s weight[g][s] * mean((out-target[s])^2);   loss_plain[g] (NULL: skip) = the
 * same with unit weights;   dout = grad_scale * weight[g][s] * 2/n_per_seg * (out - target[s]).
 * calc_dis_loss networks.py:56-64,158-166 for all scales at once, with its autograd. */
int cg_lsgan_fused(const cg_lsgan_desc* d, float* loss_total, int accumulate, float* loss_plain, void* ws,
                   size_t ws_bytes, void* stream);

typedef struct {
    int32_t G, B, H, W;                       /* mask[G][B][H][W][4]; B = this rank's batch                         */
    int32_t n_adv, n_cl;                      /* scales of MsImageDis / MsImageDisCouncil evaluated on x_fake (0..2) */
    const float* adv_out[2]; float* adv_dout[2];   /* [G][adv_n]; gradient written by pass 1 (constant coefficient)  */
    const float* cl_out[2];  float* cl_dout[2];    /* [G][cl_n];  gradient written by pass 2 (needs w_match)         */
    int32_t adv_n[2], cl_n[2];
    const float* mask;                        /* NULL: no focus terms                                               */
    float center, eps, adv_grad_scale, _pad;  /* d adv_out = adv_grad_scale * 2/adv_n * (out-1): gan_w / world       */
} cg_gen_loss_desc;
typedef struct {
    int32_t world, hist_size, head_gan, head_council;   /* history rings double[G][hist_size+1], window starts at head */
    int32_t gan_on, council_on, focus_on, matching, small_abs, small_square;
    double gan_w, council_w, w01, wtot, wtv;  /* loss weights (0 = term off)                                        */
    double numel;                             /* mask.numel() of the GLOBAL minibatch                               */
} cg_gen_loss_hp;
/* pass 1: scal[G][6] = { sum_scales mean (D(x)-1)^2, sum_scales mean (DC(x)-1)^2, sum 1/(|m-c|+eps), sum m,
 * sum |dh m|, sum |dw m| } for this rank, and the gradient of the adversarial maps (calc_gen_loss networks.py:84-90,188-194;
 * focus criteria trainer_council.py:230-250). */
int cg_gen_loss_fwd(const cg_gen_loss_desc* d, float* scal, void* ws, size_t ws_bytes, void* stream);
/* pass 2 (scal summed over ranks): assembles the generator loss of every member (trainer_council.py:392-451,497-529,
 * 559-634), appends to the loss histories and derives w_match (:518-524,576-586), publishes
 * pub[G][8] = { total of this direction, adv, zero_one, mask_total, TV, council loss, w_match, raw council },
 * total[g] (+)= direction total, and writes the council-map and mask gradients.  The caller advances head_gan /
 * head_council by one after a call that appended (gan_on && matching / council_on && matching). */
int cg_gen_loss_bwd(const cg_gen_loss_desc* d, const cg_gen_loss_hp* hp, const float* scal, double* hist_gan,
                    double* hist_council, float* total, int accumulate, float* pub, float* d_mask, void* ws,
                    size_t ws_bytes, void* stream);
size_t cg_loss_workspace_bytes(int G, int B, int H, int W);

/* plumbing: p[0:bytes] = 0 on `stream` (cudaMemsetAsync; keeps framework fill kernels out of the launch list) */
int cg_zero(void* p, size_t bytes, void* stream);

/* ---- input pipeline (the reference's per-image torchvision / Pillow transforms, utils.py:122-181, on a batch of decoded
 *      uint8 RGB images in device memory; bit-exact with Pillow 12 / torchvision 0.26, see csrc/augment.cu) ------------------ */
enum { CG_AUG_NONE = 0, CG_AUG_GRAY = 1, CG_AUG_BRIGHTNESS = 2, CG_AUG_CONTRAST = 3, CG_AUG_SATURATION = 4, CG_AUG_HUE = 5 };
/* One colour phase, in place: image b (desc[b] = {byte offset into imgs, H, W, 0}, RGB interleaved) gets opcode[b] with param[b]:
 * GRAY = RandomGrayscale; BRIGHTNESS / CONTRAST / SATURATION = ImageEnhance with factor param[b]; HUE = hue shift of
 * (int)param[b] steps of 1/255 (pass float(int(hue_factor * 255))).  lsum: B x uint64 scratch.  transforms.ColorJitter applies its
 * four ops in a random per-image order: call once per position of the permutation. */
int cg_aug_color(uint8_t* imgs, const int32_t* desc, const int32_t* opcode, const float* param, unsigned long long* lsum,
                 int B, int max_pixels, int any_contrast, void* stream);
/* n images of ONE source size H x W (src_off[i] = byte offset): optional horizontal flip, Pillow bilinear resize to oh x ow
 * (tables from Resample.c precompute_coeffs: bounds[out][2] = {first tap, count}, kk[out][ksize] 22-bit fixed point), crop window
 * crop[i] = {top, left} of size ch x cw, ToTensor + Normalize(0.5, 0.5): image i is written to batch slot slot[i] of
 * out_nhwc[.][ch][cw][4] (fp32, lane 3 = 0) and, if not NULL, of out_nchw[.][3][ch][cw].  tmp: n * H * ow * 3 bytes. */
int cg_aug_resize_crop(const uint8_t* imgs, const int32_t* src_off, const int32_t* flip, const int32_t* slot,
                       const int32_t* crop, int n, int H, int W, int oh, int ow, int ch, int cw, const int32_t* bounds_h,
                       const int32_t* kk_h, int ksize_h, const int32_t* bounds_v, const int32_t* kk_v, int ksize_v,
                       uint8_t* tmp, float* out_nhwc, float* out_nchw, void* stream);

/* ---- optimiser (torch.optim.Adam as used at trainer_council.py:170-179) ----------------------- */
int cg_adam_step(float* p, const float* g, float* m, float* v, long n, float lr, float beta1,
                 float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COUNCIL_B200_H */
