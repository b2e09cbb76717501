/*
 * gsd_b200.h -- C ABI of libgsd_b200.so, the B200 (sm_100a) implementation of the
 * gelslim_depth U-Net hot path.
 *
 * The reference (MMintLab/gelslim_depth) has NO FFI / plugin interface of its own: the seam is
 * Python duck typing.  Each entry point below therefore cites the reference
 * *Python* interface it replaces (paths relative to the reference checkout); the Python-side binding
 * a maintainer adds is the ctypes stub shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer owned by the caller unless the
 *     parameter name ends in `_host`; PyTorch (or any other allocator) stays the owner of all memory;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream, performs
 *     no allocation and no synchronisation, and is therefore CUDA-graph capturable;
 *   - return value 0 = success, negative = error (gsd_last_error() gives the message); nothing is
 *     thrown across the boundary;
 *   - one plan per (geometry, device); a plan is not thread-safe, distinct plans are.
 *   - there is NO CPU fallback: on a machine without an sm_100 device every compute call fails.
 */
#ifndef GSD_B200_H
#define GSD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSD_ABI_VERSION 1
#define GSD_MAX_DIMS 8

enum { GSD_DTYPE_BF16 = 0, GSD_DTYPE_FP32 = 1 };
enum { GSD_MODE_INFER = 0, GSD_MODE_TRAIN = 1 };

/* Geometry of one U-Net instance: the arguments of
 * gelslim_depth/models/unet.py:61  UNet(n_channels, n_classes, layer_dimensions, kernel_size=3,
 * maxpool_size=2, upconv_stride=2) plus the tensor shape of `forward(x)` (unet.py:79). */
typedef struct gsd_geometry {
  int32_t batch;        /* N of x                                     */
  int32_t in_channels;  /* n_channels (3 or 6)                        */
  int32_t height;       /* H of x (network resolution)                */
  int32_t width;        /* W of x                                     */
  int32_t n_classes;    /* 1 or 2                                     */
  int32_t n_dims;       /* len(layer_dimensions), 2..GSD_MAX_DIMS     */
  int32_t dims[GSD_MAX_DIMS]; /* layer_dimensions, each a multiple of 64 */
  int32_t dtype;        /* GSD_DTYPE_*: arithmetic of the conv path   */
  int32_t mode;         /* GSD_MODE_*                                 */
} gsd_geometry;

/* Pre/post-processing fused around the network, i.e. everything
 * gelslim_depth/processing_utils/complete_prediction.py:4-10 does besides `model(x=...)`, plus the
 * caller-side difference image (image_utils.py:6-10):
 *   x_net[c] = in_scale[c] * area_resample( use_diff ? (raw - base + 255)/2 : raw ) + in_shift[c]
 *   depth    = area_resample( y_net * out_scale + out_shift )
 * With raw_height == geometry.height (etc.) the resampling is the identity. */
typedef struct gsd_prepost {
  int32_t use_diff;             /* 1: subtract `base` (image_utils.py:7-9)                     */
  int32_t base_batch;           /* 1 (broadcast) or batch                                       */
  int32_t raw_height, raw_width;/* size of the raw frames (e.g. 320 x 427)                      */
  int32_t out_height, out_width;/* `output_size` of predict_depth_from_RGB                     */
  float in_scale[8];            /* normalize_tactile_image: scale/denominator, per channel      */
  float in_shift[8];            /*                         -scale*bias/denominator              */
  float out_scale;              /* denormalize_depth_image: denominator/scale                   */
  float out_shift;              /*                          bias                                */
  int32_t split_fingers;        /* 1: x holds batch/2 frame PAIRS with 2*in_channels channels; network sample b is
                                 *    finger b/(batch/2) of pair b%(batch/2) (general_dataset.py:71: all Left, then all Right) */
  int32_t input_u8;             /* 0: x is float 0..255 NCHW; 1: uint8 NCHW camera bytes (4x less host->device traffic);
                                   2: uint8 NHWC = interleaved frames (batch, raw_height, raw_width, channels) as a camera
                                   driver / cv2 delivers them (README.md:155: "in whatever way you acquire tactile images") */
} gsd_prepost;

typedef struct gsd_plan gsd_plan;

int gsd_abi_version(void);
const char* gsd_last_error(void);
/* Number of CUDA devices with compute capability 10.x; 0 means every compute call will fail. */
int gsd_device_count(void);

/* --- plan ------------------------------------------------------------------------------------ */
/* Replaces: UNet.__init__ (unet.py:61-77) + the implicit shape specialisation of the first call. */
int gsd_plan_create(gsd_plan** out, const gsd_geometry* g, int device);
void gsd_plan_destroy(gsd_plan* p);
/* Bytes of activation workspace / packed weights the caller must allocate (256-byte aligned). */
size_t gsd_plan_workspace_bytes(const gsd_plan* p);
size_t gsd_plan_packed_bytes(const gsd_plan* p);
/* Number of parameter tensors / BatchNorm buffer tensors gsd_pack_weights expects. */
int gsd_plan_num_params(const gsd_plan* p);
int gsd_plan_num_bn_buffers(const gsd_plan* p);
/* Kernel launches one gsd_forward issues (for bench.py's gpu_launches). */
int gsd_plan_forward_launches(const gsd_plan* p);
/* Process the batch as independent chunks of `frames_per_chunk` frames (default: the whole batch).
 * Chunks are what gsd_forward_host pipelines against the host<->device copies. */
int gsd_plan_set_chunk(gsd_plan* p, int frames_per_chunk);
/* gsd_forward_host pipelines upload | compute | download chunk by chunk, so only the first upload and the last
 * download are exposed: a smaller first / last chunk (0 = same as frames_per_chunk) shortens them. */
int gsd_plan_set_chunk_ramp(gsd_plan* p, int first_frames, int last_frames);
/* 2*M*N*K summed over the conv / transposed-conv GEMMs of one gsd_forward (valid after the first
 * forward); the denominator-free numerator of bench.py's tensor roofline. */
double gsd_plan_conv_flops(const gsd_plan* p);
/* 1 if the plan's last forward ran get_difference_image / the Left-Right split / normalize_tactile_image
 * (image_utils.py:6-10, general_dataset.py:71, normalization_utils.py:29-34) inside the first conv's producer warps
 * (raw frames already at network size), 0 if a separate prologue pass wrote the normalised 16-channel input first
 * (area down-sampling needed, fp32 mode, or GSD_NO_FUSED_PROLOGUE set). */
int gsd_plan_first_fused(const gsd_plan* p);

/* Replaces: model.load_state_dict / .to(device) on the reference module (test_depth_estimation.py:61-65).
 * params: fp32 device pointers in nn.Module.parameters() order of the reference UNet
 *         (inc.double_conv.0.weight, inc.double_conv.1.weight, inc.double_conv.1.bias, ...).
 * bn_buffers: fp32 device pointers, (running_mean, running_var) per BatchNorm in module order.
 * Folds eval-mode BatchNorm into a per-channel (scale, shift) applied in the conv epilogue and
 * re-lays conv weights as K-major bf16 GEMM operands. */
int gsd_pack_weights(gsd_plan* p, const void* const* params, const void* const* bn_buffers,
                     void* packed, void* stream);

/* The same, but only if the parameters / BatchNorm buffers CHANGED since `packed` was last written -- decided on the
 * device from a 64-bit content fingerprint, so writes the host cannot see are caught too: torch_ema 0.3's copy_to /
 * restore and the reference's weight init write through `param.data` (train_unet.py:248-250,389,428,480), which bypasses
 * torch's tensor version counters, and the training kernels of this library update parameters through raw pointers.
 * state: 4 x uint64 in device memory, zero-initialised by the caller whenever `packed` is (re)allocated; two launches
 * (fingerprint, gated pack), asynchronous on `stream`, no host synchronisation. */
int gsd_pack_weights_if_changed(gsd_plan* p, const void* const* params, const void* const* bn_buffers, void* packed,
                                unsigned long long* state, void* stream);

/* Replaces: UNet.forward (unet.py:79-88) and, with a non-trivial gsd_prepost, the whole of
 * predict_depth_from_RGB (complete_prediction.py:4-10).
 * x:    fp32 (or uint8, see gsd_prepost.input_u8) NCHW (batch, in_channels, raw_height, raw_width)
 * base: fp32 NCHW (base_batch, in_channels, raw_height, raw_width) or NULL
 * y:    fp32 NCHW (batch, n_classes, out_height, out_width) */
int gsd_forward(gsd_plan* p, const void* x, const float* base, const gsd_prepost* pp, float* y,
                void* workspace, const void* packed, void* stream);

/* Same computation with HOST buffers (pinned or pageable): copies the frames host->device in
 * `chunk`-frame pieces on a copy stream overlapped with compute, and the depth maps back.
 * `x_dev`/`y_dev` are caller-owned device staging buffers of the full batch size.
 * Blocks until y_host is complete. */
int gsd_forward_host(gsd_plan* p, const void* x_host, const float* base, const gsd_prepost* pp,
                     float* y_host, void* x_dev, float* y_dev, void* workspace, const void* packed,
                     void* stream);

/* Non-blocking form for a stream of batches (the serving loop around predict_depth_from_RGB,
 * test_depth_estimation.py:75-90, where batch k+1 is already waiting in host memory while batch k computes).
 * The caller rotates up to GSD_MAX_HOST_SLOTS staging sets (x_dev, y_dev, y_host): the call only enqueues
 * upload | compute | download and returns; the upload of the next call on another slot overlaps this call's
 * compute, and this call's download overlaps the next compute.  Re-using a slot waits (on the device, not the
 * host) until that slot's previous upload was consumed and its previous download finished.
 * gsd_forward_host_wait blocks the host until y_host of the slot's latest call is complete.
 * All calls of one plan must use the same `stream`; x_host must stay untouched until the wait returns. */
#define GSD_MAX_HOST_SLOTS 4
int gsd_forward_host_async(gsd_plan* p, const void* x_host, const float* base, const gsd_prepost* pp,
                           float* y_host, void* x_dev, float* y_dev, void* workspace, const void* packed,
                           void* stream, int slot);
int gsd_forward_host_wait(gsd_plan* p, int slot);

/* gsd_forward with CUDA events recorded between consecutive launches (synchronises `stream`):
 * ms_host[i] / flops_host[i] = device time and 2*M*N*K of launch i in network order (index 0 = input
 * prologue, 1..n-2 = conv / transposed-conv GEMMs, n-1 = 1x1 head [+ area resample]).  Measurement aid
 * for bench.py's live roofline; not on the product path. */
int gsd_forward_profiled(gsd_plan* p, const void* x, const float* base, const gsd_prepost* pp, float* y,
                         void* workspace, const void* packed, void* stream, float* ms_host,
                         double* flops_host, int capacity, int* n_out);

/* --- launch planning without a GPU (host-logic tests, tests/test_host_rules_cpu.py) -------------- */
/* Which kernel and configuration a conv3x3 layer (C0 [+ C1 concat] -> Cout at HxW, batch B) gets on a device with
 * `num_sms` SMs.  out[12] = {halo-resident (1) or tap-streaming (0) kernel, N per UMMA, M tiles per work item,
 * resident weights, epilogue warps, CTA pair (cta_group::2), halo ring depth, weight ring depth, dynamic smem bytes,
 * grid, tile height, tile width}.  Touches neither the GPU nor the driver. */
int gsd_debug_plan_conv3x3(int B, int H, int W, int C0, int C1, int Cout, int num_sms, int* out);
/* Chunk sizes gsd_forward_host would use for (batch, gsd_plan_set_chunk, gsd_plan_set_chunk_ramp); returns the
 * number of chunks written to out[capacity] (negative = error). */
int gsd_debug_chunk_schedule(int batch, int chunk, int first, int last, int* out, int capacity);

/* --- per-layer taps (parity tests localise an error to a layer: the oracle records the same tensors, oracle/unet_oracle.py
 * unet_forward_with_taps, as forward hooks on unet.py:7-57 would) -------------------------------------------------------- */
/* activations a plan keeps after gsd_forward: index 0 .. 2*n_dims-1 = encoder conv->BN->ReLU units in network order
 * (inc.double_conv.2, inc.double_conv.5, down.0...double_conv.2, ...), then per decoder block: up.i.up output,
 * up.i.conv.double_conv.2, up.i.conv.double_conv.5. */
int gsd_debug_num_activations(const gsd_plan* p);
int gsd_debug_activation_shape(const gsd_plan* p, int index, int* C, int* H, int* W);
/* -> dst fp32 NCHW (batch, C, H, W) from the workspace of the last (single-chunk) gsd_forward.  The very last unit is only
 * stored when the 1x1 head is not fused into it (environment GSD_NO_HEAD_FUSION at bind time). */
int gsd_debug_read_activation(const gsd_plan* p, int index, const void* workspace, float* dst, void* stream);

/* --- single operators (used by the parity tests; the plan is built from exactly these) ---------- */
/* conv KxK (taps given explicitly) as implicit GEMM on tcgen05, NHWC bf16.
 *   src0:(B,H,W,C0) [+ src1:(B,H1,W1,C1) placed at offset (off_y, off_x), zero elsewhere -> virtual
 *   F.pad + torch.cat of unet.py:43-48];  w: bf16 [Cout_total][ntaps*(C0+C1)] K-major;
 *   out = relu?(acc*scale[n] + shift[n]) as bf16 NHWC (B,H,W,Cout); optional 2x2 max-pooled copy.
 *   out_groups == 4 selects the transposed-conv scatter (unet.py:36): w rows are
 *   [(dy*2+dx)*Cout + co], out is (B,2H,2W,Cout). */
int gsd_op_conv_bf16(const void* src0, int C0, const void* src1, int C1, int H1, int W1, int off_y,
                     int off_x, int B, int H, int W, const void* w, int Cout, int ntaps,
                     const int8_t* tap_dy, const int8_t* tap_dx, int out_groups, const float* scale,
                     const float* shift, int relu, void* out, void* pooled, int block_n, int device,
                     void* stream);

/* conv3x3 / pad 1 through the halo-resident kernel (csrc/conv_halo.cuh): same tensors as
 * gsd_op_conv_bf16 with the 9 taps implied.  base_off_mode selects how the UMMA descriptor encodes the
 * swizzle phase of the shifted tap views (1 = PTX-ISA base-offset field; 0 = experiment). */
int gsd_op_conv3x3_halo_bf16(const void* src0, int C0, const void* src1, int C1, int H1, int W1, int off_y,
                             int off_x, int B, int H, int W, const void* w, int Cout, const float* scale,
                             const float* shift, int relu, void* out, void* pooled, int block_n,
                             int base_off_mode, int device, void* stream);

/* conv3x3 / pad 1 weight gradient (autograd of unet.py:11,14 -- train_unet.py:374) as a tcgen05 GEMM over pixels:
 *   dw[co][tap][ci] += sum_{b,y,x} dz[b,y,x,co] * x[b,y+dy,x+dx,ci]      (fp32, accumulated; caller zeroes dw)
 * x0/x1: NHWC bf16 sources of the (virtual) concat, dz: NHWC bf16 gradient of the conv output. */
int gsd_op_wgrad3x3_bf16(const void* x0, int C0, const void* x1, int C1, int H1, int W1, int off_y, int off_x,
                         const void* dz, int Cout, int B, int H, int W, float* dw, int device, void* stream);

/* --- training-step operators (train_utils/train_unet.py:346-377); one kernel launch each ------------------ */
/* conv with automatic kernel choice (halo-resident / tap-streaming) + optional batch statistics:
 * stats[0..N) += sum over valid pixels of the raw accumulator, stats[N..2N) += sum of squares (train-mode
 * BatchNorm2d, unet.py:12,15).  ntaps: 9 (3x3, pad 1) or 1; groups 4 = transposed-conv scatter.
 * scale == NULL means 1, shift == NULL means 0 (training forward / dgrad): the epilogue then skips the per-channel
 * constant loads, which cost as much shared-memory bandwidth as its bf16 transposition. */
int gsd_op_conv_auto_bf16(const void* src0, int C0, const void* src1, int C1, int H1, int W1, int off_y,
                          int off_x, int B, int H, int W, const void* w, int Cout, int ntaps, int groups,
                          const float* scale, const float* shift, int relu, void* out, void* pooled,
                          float* stats, int device, void* stream);
/* ConvTranspose2d(k=2,s=2) backward (unet.py:36): input gradient (tcgen05 GEMM over a 5-D space-to-depth TMA view of
 * dU) and weight gradient (tcgen05 GEMM over pixels, dw (Cin,Cout,2,2) fp32 accumulated).  dU is a dense
 * (B,Hf,Wf,C) tensor whose (2H x 2W) window at (off_y, off_x) is the gradient of the up-sampled map. */
int gsd_op_convt_dgrad_bf16(const void* du, int Cs, int Hf, int Wf, int off_y, int off_x, const void* w, int Cin,
                            int B, int H, int W, const float* scale, const float* shift, void* out, int device,
                            void* stream);
int gsd_op_convt_wgrad_bf16(const void* in, int Cin, const void* du, int Cout, int Hf, int Wf, int off_y, int off_x,
                            int B, int H, int W, float* dw, int device, void* stream);
/* fp32 NCHW frames -> NHWC bf16 padded to 16 channels (same arithmetic as the fused forward prologue) */
int gsd_op_prologue_bf16(const float* x, const float* base, int base_batch, int use_diff, int B, int C, int Hr,
                         int Wr, int H, int W, const float* scale8_host, const float* shift8_host, void* out16,
                         void* stream);
/* train-mode BatchNorm2d: batch mean / biased variance -> (scale, shift, mean, rstd); running stats updated in
 * place with momentum and the unbiased variance (running_mean may be NULL).
 * neg_center (or NULL): the conv epilogue stored z - center (center = running mean before this step, passed to the
 * conv as shift = -center) so that bf16 rounds relative to the fluctuation; outputs then refer to the stored tensor. */
int gsd_op_bn_finalize(const float* stats, double count, const float* gamma, const float* beta, float* running_mean,
                       float* running_var, float momentum, float eps, int C, const float* neg_center, float* scale,
                       float* shift, float* mean, float* rstd, long long* num_batches_tracked /* += 1, or NULL */,
                       void* stream);
int gsd_op_negate_f32(const float* in, int n, float* out, void* stream);
/* a = relu(z*scale + shift) (+ MaxPool2d(2) copy, unet.py:26) */
int gsd_op_bn_relu_apply(const void* z, const float* scale, const float* shift, int B, int H, int W, int C, void* a,
                         void* pooled, void* stream);
/* MSE_loss (train_unet.py:51-52): *loss += mean((y-t)^2), dy = 2(y-t)/n */
int gsd_op_mse(const float* y, const float* t, long long n, float* loss, float* dy, void* stream);
/* OutConv (unet.py:54) forward on a (B,H,W,64) bf16 activation and its backward (da bf16, dw/db fp32 accumulated) */
int gsd_op_head_fwd(const void* a, const float* w, const float* bias, int ncls, int B, int H, int W, float* y, void* stream);
int gsd_op_head_bwd(const void* a, const float* dy, const float* w, int ncls, int B, int H, int W, void* da, float* dw,
                    float* db, void* stream);
/* BatchNorm+ReLU backward: reduction pass (sums = [sum g, sum g*zhat] = [dbeta, dgamma]) and apply pass (dz) */
/* (the ReLU mask is recomputed as z*scale + shift > 0, so the post-ReLU tensor is not re-read; z == NULL in
 * gsd_op_bn_bwd_reduce gives a plain per-channel sum of `da`) */
int gsd_op_bn_bwd_reduce(const void* da, const float* scale, const float* shift, const void* z, const float* mean,
                         const float* rstd, long long npix, int C, float* sums, void* stream);
int gsd_op_bn_bwd_apply(const void* da, const float* scale, const float* shift, const void* z, const float* mean,
                        const float* rstd, const float* gamma, const float* sums, double count, long long npix, int C,
                        void* dz, void* stream);
/* Last unit of the network (up.3.conv.3 -> BatchNorm2d -> ReLU -> OutConv, unet.py:17,54-57) without materialising
 * the post-ReLU tensor: forward y = OutConv(relu(z*scale + shift)) from the raw conv output z ... */
int gsd_op_bn_relu_head_fwd(const void* z, const float* scale, const float* shift, const float* w, const float* bias,
                            int ncls, int B, int H, int W, float* y, void* stream);
/* ... and its backward in two passes over z: sums[128] (zeroed by the caller) = [dbeta | dgamma], dw/db (accumulated)
 * = OutConv weight / bias gradients, dz = gradient of the raw conv output.  Equals gsd_op_head_bwd followed by
 * gsd_op_bn_bwd_reduce / _apply on the materialised tensors (3 tensor passes instead of 7). */
int gsd_op_head_bn_bwd(const void* z, const float* dy, const float* w, const float* scale, const float* shift,
                       const float* mean, const float* rstd, const float* gamma, double count, int ncls, int B, int H,
                       int W, float* sums, float* dw, float* db, void* dz, void* stream);
/* MaxPool2d(2) backward fused with the skip-connection gradient add (dskip may be NULL) */
int gsd_op_maxpool_bwd(const void* a, const void* dpool, const void* dskip, int B, int H, int W, int C, void* dfull,
                       void* stream);
/* fp32 parameter -> bf16 GEMM operand; modes: 0 conv fwd, 1 conv dgrad (flipped taps), 2 convT fwd, 3 convT dgrad */
int gsd_op_pack_weight(int mode, const float* w, int O, int I, int Ipad, void* out, void* stream);
/* the same for every layer of the network in ONE launch: `items_dev` is a DEVICE array of n_items records (the pointers
 * are stable across steps because parameters alias a flat arena).  mode 0 items may carry `out_dgrad`: the dgrad
 * operand (mode 1 layout) is then written from the same staged tile.  `start` = running sum of gsd_pack_item_units()
 * of the preceding items, `total_units` = the sum over all items. */
typedef struct gsd_pack_item {
  const float* w;
  void* out;
  void* out_dgrad;            /* mode 0 only, may be NULL */
  int32_t mode, O, I, Ipad;
  int64_t start;
} gsd_pack_item;
long long gsd_pack_item_units(int mode, int O, int I, int Ipad);
int gsd_op_pack_weights_batched(const gsd_pack_item* items_dev, int n_items, long long total_units, void* stream);
/* wgrad arena [O][9][Ipad] fp32 -> Conv2d.weight.grad layout (O,I,3,3); clear != 0: the arena is zeroed as it is
 * read, ready for the next step's accumulation */
int gsd_op_unpack_wgrad(float* dwk, int O, int I, int Ipad, float* grad, int clear, void* stream);
/* torch.optim.Adam(lr, betas, eps, weight_decay) with coupled L2 (train_unet.py:306,375) fused with the
 * torch_ema==0.3 update (train_unet.py:309,376) over one flat fp32 arena; `step` is 1-based, shadow may be NULL */
int gsd_op_adam_ema(float* p, const float* g, float* m, float* v, float* shadow, long long n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, long long step, float ema_decay, long long ema_updates,
                    float grad_scale, void* stream);

/* CUDA-graph-replayable form: `counter` = 2 device int64 (Adam steps, EMA updates done so far), advanced on device */
int gsd_op_adam_ema_dev(float* p, const float* g, float* m, float* v, float* shadow, long long n, float lr, float beta1,
                        float beta2, float eps, float weight_decay, float ema_decay, long long* counter, float grad_scale,
                        void* stream);

/* --- training plan: the loop body of train_utils/train_unet.py:346-377 sequenced by the library --------------------------
 *     optimizer.zero_grad(); output = unet(x=input); loss = MSE_loss(output, target); loss.backward();
 *     optimizer.step(); ema.update()
 * One gsd_train_plan per (geometry with mode = GSD_MODE_TRAIN, device).  All memory is the caller's: the workspace
 * (gsd_train_plan_workspace_bytes: saved activations, gradient temporaries, bf16 operands, per-layer constants), the fp32
 * parameters / gradients / BatchNorm buffers (pointers in nn.Module.parameters() order, see gsd_pack_weights) and the
 * optimizer arenas.  Every call below only enqueues library kernels and memset / copy nodes on `stream` (plus a plan-owned
 * side stream that is forked from and joined back into it with events): no allocation, no host synchronisation, CUDA-graph
 * capturable.  gsd_train_plan_bind is the one exception (uploads small tables, synchronises). */
typedef struct gsd_train_plan gsd_train_plan;
/* Fires on the HOST while the step is being enqueued: bucket >= 0 -- the last gradient kernel of gradient bucket `bucket`
 * (flat-arena elements [lo, hi), see gsd_train_plan_set_buckets) has just been enqueued on main_stream / side_stream
 * (NULL if not in use): order the communication stream after both and enqueue the all-reduce of that range (replaces the
 * reducer hooks of torch DistributedDataParallel around train_unet.py:374).  bucket == -1 (gsd_train_step only): backward
 * is fully enqueued, the optimizer kernel comes next: make main_stream wait for the communication stream. */
typedef void (*gsd_bucket_cb)(void* user, int bucket, long long lo, long long hi, void* main_stream, void* side_stream);
typedef struct gsd_adam {
  float lr, beta1, beta2, eps, weight_decay;  /* torch.optim.Adam(lr=1e-3, weight_decay=1e-6), train_unet.py:306 */
  float ema_decay;                            /* torch_ema ExponentialMovingAverage(decay=0.995), train_unet.py:309 */
  float grad_scale;                           /* multiplies the gradient first: 1 / world_size after a sum all-reduce */
} gsd_adam;
typedef struct gsd_optimizer_state {
  float* params;        /* flat fp32 arena the bound parameter pointers alias, n elements */
  const float* grads;   /* flat arena the bound gradient pointers alias */
  float* m;             /* Adam first / second moments */
  float* v;
  float* ema;           /* EMA shadow, or NULL */
  long long n;
  long long* counter;   /* device int64[2]: Adam steps / EMA updates done so far (advanced on the device) */
  gsd_adam hp;
} gsd_optimizer_state;
/* Replaces: UNet(...).train() + the shape specialisation of the first step (train_unet.py:235,344).
   g->mode = GSD_MODE_TRAIN; g->dtype = GSD_DTYPE_BF16 (tcgen05 path, the measured one) or GSD_DTYPE_FP32 (FFMA parity path:
   the same entry points and protocol on fp32 activations, ~1/50 of the speed, for gradient / loss-curve comparisons with the
   fp32 reference; csrc/train_plan_f32.h). */
int gsd_train_plan_create(gsd_train_plan** out, const gsd_geometry* g, int device);
void gsd_train_plan_destroy(gsd_train_plan* p);
size_t gsd_train_plan_workspace_bytes(const gsd_train_plan* p);
int gsd_train_plan_num_params(const gsd_train_plan* p);
int gsd_train_plan_num_bn(const gsd_train_plan* p);
/* element count of every parameter in parameters() order (to lay out flat arenas); returns the number of parameters */
int gsd_train_plan_param_numel(const gsd_train_plan* p, long long* out, int capacity);
/* kernel launches one gsd_train_step enqueues */
int gsd_train_plan_launches(const gsd_train_plan* p);
/* params / grads: fp32 device pointers in parameters() order; bn_buffers: (running_mean, running_var) per BatchNorm in
 * module order; num_batches_tracked: int64 device pointers per BatchNorm, or NULL.  Re-bind whenever a pointer changes. */
int gsd_train_plan_bind(gsd_train_plan* p, const void* const* params, void* const* grads, void* const* bn_buffers,
                        long long* const* num_batches_tracked, void* workspace);
/* Gradient buckets for the data-parallel all-reduce: bucket_of_param[i] in [0, n_buckets), bucket b covers flat-arena
 * elements [lo[b], hi[b]).  n_buckets == 0 disables the callback. */
int gsd_train_plan_set_buckets(gsd_train_plan* p, int n_buckets, const int* bucket_of_param, const long long* lo,
                               const long long* hi);
/* Replaces: `output = unet(x=input)` in .train() mode (train_unet.py:347; UNet.forward unet.py:79-88 with batch-statistics
 * BatchNorm: running statistics updated in place with momentum 0.1 / unbiased variance, num_batches_tracked += 1). */
int gsd_train_forward(gsd_train_plan* p, const float* x, float* y, void* stream);
/* Replaces: `loss.backward()` (train_unet.py:374) given dy = d loss / d output (fp32 NCHW): dgrad / wgrad as tcgen05
 * GEMMs, BatchNorm / ReLU / max-pool / transposed-conv backward; every parameter gradient lands at its bound pointer. */
int gsd_backward(gsd_train_plan* p, const float* dy, void* stream, gsd_bucket_cb cb, void* user);
/* Replaces: `optimizer.step(); ema.update()` (train_unet.py:375-376) over flat arenas. */
int gsd_adam_ema_step(float* param_arena, const float* grad_arena, float* m, float* v, float* ema, long long n,
                      const gsd_adam* hp, long long* counter, void* stream);
/* Replaces: the whole loop body (train_unet.py:346-377).  loss: device float, overwritten with MSE_loss(output, target)
 * (train_unet.py:51-52); opt == NULL stops after backward (gradients only). */
int gsd_train_step(gsd_train_plan* p, const float* x, const float* target, float* loss, const gsd_optimizer_state* opt,
                   void* stream, gsd_bucket_cb cb, void* user);

/* Host-logic tests: the same plan without a GPU.  gsd_backward on it enqueues nothing but walks the real step structure,
 * i.e. reports complete gradient buckets through the callback in the real order (tests/test_ddp_cpu.py). */
int gsd_debug_train_plan_create(gsd_train_plan** out, const gsd_geometry* g);

/* Stand-alone processing helper: fp32 NCHW -> fp32 NCHW,
 *   out[:, c] = scale8[min(c,7)] * area_resample(use_diff ? (x - base + 255)/2 : x) + shift8[min(c,7)]
 * Replaces (when called outside the fused forward): get_difference_image (image_utils.py:6-10),
 * sample_multi_channel_image_to_desired_size(..., 'area') (image_utils.py:12-15),
 * normalize_tactile_image / normalize_depth_image / denormalize_depth_image
 * (normalization_utils.py:4-35, 70-130).  scale8/shift8 are HOST arrays of 8 floats.
 * split_fingers != 0: x is (B/2, 2C, Hr, Wr) Left|Right frame pairs and out is (B, C, H, W) = all Left fingers, then
 * all Right fingers (torch.cat((x[:, :C], x[:, C:]), dim=0), general_dataset.py:71-74); base likewise. */
int gsd_op_image_affine(const float* x, const float* base, int base_batch, int use_diff, int B, int C,
                        int Hr, int Wr, int H, int W, const float* scale8_host, const float* shift8_host,
                        float* out, int split_fingers, int device, void* stream);
/* blur_depth_images (image_utils.py:17-19) = torchvision gaussian_blur: depthwise kernel_size x kernel_size Gaussian,
 * reflect padding, on `planes` fp32 H x W planes; sigma <= 0 selects torchvision's default for the kernel size. */
int gsd_op_gaussian_blur(const float* x, int planes, int H, int W, int kernel_size, float sigma, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GSD_B200_H */
