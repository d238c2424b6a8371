/*
 * pdes_b200.h — C-ABI of the B200-native backend for the physics-constrained
 * DenseED training hot path of cics-nd/pde-surrogate.
 *
 * The reference has no FFI: its "plugin interface" for this path is the Python
 * module namespace (models.codec.DenseED, models.darcy.conv_*, utils.image_gradient.
 * SobelFilter).  This header is the boundary a binding for that namespace talks to:
 * extern "C", plain pointers and sizes, no torch / C++ types.  All pointers are DEVICE
 * pointers unless a parameter is documented as host.  `stream` is a cudaStream_t passed
 * as void* (NULL = legacy default stream).  The caller owns every buffer; the training-path
 * entry points (pdes_darcy_loss_*, pdes_densenet_forward/backward, pdes_adam_*) never allocate,
 * free or synchronise the host.  Exceptions, all outside the training loop:
 *   - pdes_densenet_bind synchronises the device once (the caller's workspace fill may be in flight
 *     on another stream) and creates the executor's internal low-priority stream + two events;
 *   - the unit-test entry points pdes_conv2d_fwd/dgrad/wgrad take scratch for the packed operands from
 *     the stream-ordered allocator (cudaMallocAsync / cudaFreeAsync) and synchronise the stream once
 *     while uploading their descriptor table;
 *   - pdes_densenet_timing_report / _read synchronise the device (diagnostics).
 *
 * Every function returns PDES_OK (0) or a non-zero code; pdes_last_error() then holds a
 * human-readable message (thread-local).
 *
 * Reference citations are relative to the upstream repository root.
 */
#ifndef PDES_B200_H
#define PDES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDES_ABI_VERSION 3

#define PDES_OK 0
#define PDES_ERR_INVALID 1     /* bad argument (shape, null pointer, alignment)        */
#define PDES_ERR_CUDA 2        /* a CUDA runtime / driver call failed                  */
#define PDES_ERR_UNSUPPORTED 3 /* valid in the reference, not implemented by this path */
#define PDES_ERR_STATE 4       /* call order violated (e.g. backward before forward)   */

const char* pdes_last_error(void);
int pdes_abi_version(void);
/* sm count and compute capability of the current device (host out-pointers). */
int pdes_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------
 * Sobel finite-difference stencils.
 * Replaces SobelFilter.grad_h / SobelFilter.grad_v (utils/image_gradient.py:50-75,
 * 77-92) for filter_size=3: replicate-pad 1, 3x3 cross-correlation with VSOBEL/HSOBEL
 * /8, times the image width (height), then the 3-point one-sided boundary modifier
 * (image_gradient.py:43-46) when `correct` != 0.
 *   img, out : n_img planes of H*W fp32, contiguous.
 *   dir      : 0 = grad_h (d/dx, last dim), 1 = grad_v (d/dy).
 *   adjoint  : 0 = apply the operator, 1 = apply its exact transpose (autograd backward).
 * ---------------------------------------------------------------------------------- */
int pdes_sobel_grad(const float* img, float* out, int64_t n_img, int H, int W, int dir,
                    int correct, int adjoint, void* stream);

/* ------------------------------------------------------------------------------------
 * Fused Darcy mixed-residual loss.
 * Replaces, in ONE launch, conv_constitutive_constraint (models/darcy.py:162-176),
 * conv_continuity_constraint (models/darcy.py:210-224) and conv_boundary_condition
 * (models/darcy.py:226-233) evaluated on the same (input, output) pair.
 *   K      : (B,1,H,W) fp32 permeability ("input"); may be NULL -> constitutive term = 0.
 *   out    : (B,3,H,W) fp32 NCHW contiguous: u, sigma1, sigma2.
 *   loss4  : 4 floats written: [constitutive, continuity, dirichlet, neumann].
 *   use_tb : darcy.py:221-224 (0 drops rows 0 and H-1 from the continuity mean).
 *   ws     : pdes_darcy_loss_workspace_bytes() bytes, zero-initialised ONCE by the caller;
 *            the kernel leaves it zeroed again on completion.
 * ---------------------------------------------------------------------------------- */
size_t pdes_darcy_loss_workspace_bytes(void);
int pdes_darcy_loss_fwd(const float* K, const float* out, int B, int H, int W, int use_tb,
                        float* loss4, void* ws, void* stream);
/* Gradient of  sum_i gw4[i] * loss4[i]  w.r.t. `out`  (closed form).
 *   gw4  : 4 fp32 upstream gradients ON THE DEVICE (no host sync).
 *   dout : (B,3,H,W) fp32, overwritten. */
int pdes_darcy_loss_bwd(const float* K, const float* out, const float* gw4, int B, int H,
                        int W, int use_tb, float* dout, void* stream);
/* The same two calls for the NONLINEAR constitutive law of conv_constitutive_constraint_nonlinear
 * (models/darcy.py:179-191):  -K grad(u) = sigma + beta1 sqrt(K) sigma^2 + beta2 K sigma^3  (loss4[0]; the other
 * three terms are unchanged).  K is required.  Used by solve_conv_mixed_residual.py --nonlinear (lines 73-77, 134-137). */
int pdes_darcy_loss_nl_fwd(const float* K, const float* out, int B, int H, int W, int use_tb,
                           float beta1, float beta2, float* loss4, void* ws, void* stream);
int pdes_darcy_loss_nl_bwd(const float* K, const float* out, const float* gw4, int B, int H, int W,
                           int use_tb, float beta1, float beta2, float* dout, void* stream);
/* Selects the implementation of the two calls above: 0 = auto, 1 = force the generic
 * (any H,W) kernels, 2 / 3 = force the whole-image-in-shared-memory TMA kernels with 256 / 512
 * threads per CTA (unrolled 4- / 2-row strips when they tile the image exactly, else rolling
 * strips), 4 / 5 = the same thread counts with rolling strips always. Test / A-B hook. */
int pdes_darcy_loss_set_impl(int impl);

/* ------------------------------------------------------------------------------------
 * DenseED network executor.
 * Replaces DenseED.__init__/forward (models/codec.py:211-296) built from _DenseLayer
 * (43-75, non-bottleneck), _DenseBlock (78-86), _Transition (89-160, bottleneck=True,
 * upsample='nearest') and last_decoding (163-188), with nn.BatchNorm2d + ReLU fused into
 * the consuming convolution, plus the autograd backward of all of it.
 * ---------------------------------------------------------------------------------- */
typedef struct pdes_net pdes_net_t;

typedef struct pdes_densenet_config {
  int32_t in_channels;   /* DenseED(in_channels=...)           */
  int32_t out_channels;  /* DenseED(out_channels=...)          */
  int32_t imsize;        /* square input H = W                 */
  int32_t n_blocks;      /* len(blocks), odd                   */
  int32_t blocks[15];    /* layers per dense block             */
  int32_t growth_rate;   /* default 16                         */
  int32_t init_features; /* default 48                         */
  int32_t max_batch;     /* capacity the workspace is sized for */
  int32_t arch;          /* 2: cGlow coupling network _DenseCoupling (models/glow_msc.py:276-294): blocks[0] dense
                          * layers on a planar (B, in_channels, imsize, imsize) input, then BatchNorm -> ReLU ->
                          * Conv2dZeros (3x3 with bias and the exp(3*scale) gain, glow_msc.py:240-255) to out_channels.
                          * 0: DenseED (models/codec.py:210-318).  1: Decoder (models/codec.py:321-370): a plain 3x3
                          * conv0 on a planar (B, in_channels, imsize, imsize) latent, then decoding blocks only;
                          * the output is imsize * 2^n_blocks wide.  (ABI version 2) */
  int32_t dropout;       /* 1: the network was built with drop_rate > 0 (nn.Dropout2d behind its convolutions,
                          * models/codec.py:70-71, 110-149, 171-172): pdes_densenet_set_dropout may be used */
  int32_t upsample;      /* x2 upsampling of the decoding transitions (models/codec.py:139-150, 175-178):
                          * 0 'nearest' (default), 1 'bilinear' (align_corners=True), 2 None: the transitions use
                          * nn.ConvTranspose2d(k3, s2, p1, op1) named convT2 (weight (Cin, Cout, 3, 3), codec.py:139-142)
                          * and the last decoding does not upsample (codec.py:176-179): the output is half as wide */
  int32_t bottleneck;    /* bottleneck * growth_rate = the width above which a dense layer takes the bottleneck form
                          * norm1 -> relu -> conv1 (1x1, in -> bottleneck * growth) -> norm2 -> relu -> conv2 (3x3 -> growth)
                          * (models/codec.py:56-64: DenseED(bottleneck=True, bn_size=...)); 0 = plain dense layers.
                          * (ABI version 3) */
} pdes_densenet_config;

/* Host-side object (no device memory). */
int pdes_densenet_create(const pdes_densenet_config* cfg, pdes_net_t** net);
void pdes_densenet_destroy(pdes_net_t* net);

/* Parameter tensors in the reference's named_parameters() order, living in ONE flat fp32
 * buffer (conv weights logical OIHW, BatchNorm weight/bias).  kind: 0 conv weight,
 * 1 BN weight, 2 BN bias.  name is the reference state_dict key. */
int pdes_densenet_num_params(const pdes_net_t* net);
int64_t pdes_densenet_param_floats(const pdes_net_t* net); /* flat length incl. padding */
int pdes_densenet_param_info(const pdes_net_t* net, int idx, char* name, size_t name_cap,
                             int64_t* offset, int32_t* ndim, int64_t shape[4], int32_t* kind);
/* BatchNorm running statistics: one flat fp32 buffer [running_mean | running_var] per
 * BN layer, in module order. */
int pdes_densenet_num_bn(const pdes_net_t* net);
int64_t pdes_densenet_running_floats(const pdes_net_t* net);
int pdes_densenet_bn_info(const pdes_net_t* net, int idx, char* name, size_t name_cap,
                          int64_t* mean_offset, int64_t* var_offset, int32_t* channels);

/* nn.Dropout2d sites in execution order: fills channels[i] (the convolution's Cout) for up to `cap` sites and
 * returns their number.  Before a TRAINING forward the caller passes the masks of that pass, one contiguous
 * (B, channels[i]) fp32 block per site in the same order (values 0 or 1/(1-p); device memory that stays valid
 * until the matching backward has run); NULL switches dropout off (evaluation). */
int pdes_densenet_dropout_sites(const pdes_net_t* net, int32_t* channels, int cap);
int pdes_densenet_set_dropout(pdes_net_t* net, const float* masks);

/* Spatial size (H = W) of the network output: imsize for DenseED (imsize / 2 with upsample = 2), imsize * 2^n_blocks
 * for Decoder (imsize * 2^(n_blocks-1) with upsample = 2). */
int pdes_densenet_output_size(const pdes_net_t* net);

size_t pdes_densenet_workspace_bytes(const pdes_net_t* net);
/* Bind device buffers.  params/grads: pdes_densenet_param_floats() floats each;
 * running: pdes_densenet_running_floats() floats; workspace: zero-initialised by the caller
 * once.  grads may be NULL for inference-only use. */
int pdes_densenet_bind(pdes_net_t* net, float* params, float* grads, float* running,
                       void* workspace, size_t workspace_bytes);

/* x: (B,in_channels,imsize,imsize) NCHW fp32; out: (B,out_channels,imsize,imsize) NCHW.
 * training != 0: batch statistics, running-stat update (momentum 0.1, unbiased var,
 * torch.nn.BatchNorm2d semantics) and activations kept for backward.
 * training == 0: running statistics (model.eval()).
 * From the second call with the same (B, training) on, the launches of the pass are replayed as ONE CUDA graph
 * (captured once on a private stream, launched into `stream`; x / out are staged through the workspace); the same
 * holds for pdes_densenet_backward.  Not while `stream` itself is being captured (then the launches are recorded
 * into the caller's graph), not for dropout networks, not with PDES_EXEC_GRAPH=0 in the environment. */
int pdes_densenet_forward(pdes_net_t* net, const float* x, float* out, int B, int training,
                          void* stream);
/* dout: dL/d(out), (B,out_channels,imsize,imsize).  ADDS dL/d(param) into `grads`
 * (the caller zeroes it: model.zero_grad()).  Must follow a training forward of the
 * same B. */
int pdes_densenet_backward(pdes_net_t* net, const float* dout, void* stream);
/* The same for a coupling network (arch 2), also returning the gradient w.r.t. the network input:
 * dx is a planar (B, in_channels, H, W) fp32 buffer (overwritten) or NULL. */
int pdes_densenet_backward_dx(pdes_net_t* net, const float* dout, float* dx, void* stream);
/* Useful (2*MAC) FLOPs of one forward / one forward+backward at batch B. */
double pdes_densenet_flops(const pdes_net_t* net, int B, int training);
/* Number of kernel launches issued by the last forward / backward call. */
int pdes_densenet_last_launches(const pdes_net_t* net);
/* Implementation selector for the convolutions: 0 = auto (tensor-core kernels where
 * available), 1 = force the SIMT fp32 kernels everywhere. Test hook. */
int pdes_densenet_set_conv_impl(pdes_net_t* net, int impl);
/* Diagnostics: set_timing(net, 1) records one CUDA event after every eager launch of the executor;
 * timing_report synchronises the device and prints one "<us> <label>" line per launch to stderr. */
int pdes_densenet_set_timing(pdes_net_t* net, int on);
int pdes_densenet_timing_report(pdes_net_t* net);
/* The same measurements returned to the caller: us[i] / flops[i] (useful 2*MAC of a convolution launch,
 * 0 otherwise) of launch i and the '\n'-separated labels; returns the number of launches, or a
 * negative error code.  Clears the recorded marks. */
int pdes_densenet_timing_read(pdes_net_t* net, float* us, double* flops, int cap, char* labels,
                              size_t labels_cap);

/* ------------------------------------------------------------------------------------
 * Fused Adam over a flat buffer; replaces torch.optim.Adam.step() as used at
 * train_codec_mixed_residual.py:151,239 (amsgrad=False, L2 weight_decay added to the
 * gradient).  `step` is the 1-based step count, `grad_scale` multiplies g first (1/world
 * after a sum all-reduce).
 * ---------------------------------------------------------------------------------- */
int pdes_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                   int64_t step, void* stream);

/* Same update with the step-dependent scalars read from DEVICE memory, so that the launch can be
 * replayed from a CUDA graph: hyper = {lr/(1-beta1^t), 1/sqrt(1-beta2^t), beta1, beta2, eps,
 * weight_decay, grad_scale} (7 floats, see pdes_adam_hyper). */
int pdes_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper,
                       void* stream);
/* Fills the 7 host floats consumed by pdes_adam_step_dev for 1-based step `step`. */
int pdes_adam_hyper(float* hyper7_host, float lr, float beta1, float beta2, float eps,
                    float weight_decay, float grad_scale, int64_t step);

/* ------------------------------------------------------------------------------------
 * Single convolution entry points (the building blocks the executor uses), exposed for
 * unit tests against torch.nn.functional.conv2d.  NHWC activations.
 * ---------------------------------------------------------------------------------- */
typedef struct pdes_conv_desc {
  int32_t B, Hin, Win;       /* stored input spatial size (before upsampling)        */
  int32_t Cin, ld_in;        /* channels read, and the pixel stride of the buffer    */
  int32_t Hout, Wout;        /* output spatial size                                  */
  int32_t Cout, ld_out;      /* channels written, pixel stride of the output buffer  */
  int32_t c_off_out;         /* first channel written inside the output pixel        */
  int32_t KH, KW, stride, pad;
  int32_t upsample;          /* 1: nearest x2 of the input is folded into addressing */
  int32_t bn_relu;           /* 1: a = max(0, x*scale[c]+shift[c]) applied to input  */
  int32_t out_nchw;          /* 1: output written planar (B,Cout,Hout,Wout)          */
} pdes_conv_desc;

/* Precision of the tensor-core paths of the three entry points below (impl 2 / 4): 0 = fp32-accurate (every
 * operand as two fp16 pieces, three tensor-core products per useful product; default), 1 = one fp16 piece,
 * 2 = one bf16 piece (one product; the "bf16 tensor-core conv path" of the reference's benchmark configs).
 * The DenseED executor takes the same choice from the environment: PDES_CONV_DTYPE = fp32 | fp16 | bf16. */
int pdes_conv2d_set_precision(int mode);

/* w: OIHW fp32 (Cout,Cin,KH,KW); scale/shift: Cin floats or NULL; y as described.
 * ch_sum/ch_sumsq: Cout doubles accumulated (+=) with the per-channel sum and sum of
 * squares of the outputs, or NULL.  impl: 0/1 CUDA-core fp32, 2 tensor-core (two-piece fp16), 3 the dedicated
 * first-convolution kernels (x is then the PLANAR (B, Cin, H, W) network input; fwd and wgrad only),
 * 4 (fwd only) the fused thin-layer kernel: BatchNorm+ReLU+operand split inside the convolution
 * (3x3, stride 1, pad 1, Cout <= 16, Win in {8, 16, 32}). */
int pdes_conv2d_fwd(const pdes_conv_desc* d, const float* x, const float* w,
                    const float* scale, const float* shift, float* y, double* ch_sum,
                    double* ch_sumsq, int impl, void* stream);
/* dx_pre[p,ci] = d/d(a[p,ci]) of the same convolution given dy (NHWC, ld_out stride,
 * channel offset c_off_out), i.e. gradient w.r.t. the BN+ReLU'd (and upsampled) operand,
 * already summed over the 2x2 upsampling footprint; written densely as (B,Hin,Win,Cin). */
int pdes_conv2d_dgrad(const pdes_conv_desc* d, const float* dy, const float* w, float* da,
                      int impl, void* stream);
/* dw (OIHW) += sum_p a[p+tap,ci] * dy[p,co]. */
int pdes_conv2d_wgrad(const pdes_conv_desc* d, const float* x, const float* scale,
                      const float* shift, const float* dy, float* dw, int impl,
                      void* stream);

/* Diagnostics, host only: the tiling of the tensor-core convolution kernel for a GEMM-K operand of Cin_k
 * channels and N (multiple of 16, <= 256) output channels: out[0..10] = supported, channels per chunk,
 * chunks, accumulator groups, accumulator sets, accumulator stages, activation ring depth, filter ring
 * depth, taps per filter stage, dynamic shared memory bytes, TMEM columns. */
int pdes_conv_tc_plan(int KS, int Cin_k, int N, int64_t out[11]);

#ifdef __cplusplus
}
#endif
#endif /* PDES_B200_H */
