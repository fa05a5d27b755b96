/*
 * loma_nerf_b200.h -- C ABI of libloma_nerf_b200.so
 *
 * A from-scratch sm_100a (B200) implementation of the one hot path loma-nerf runs through its
 * loma-compiled _code/nerf.so and _code/mlp_fit.so: positional encoding -> per-sample coordinate
 * MLP (forward + reverse-mode gradient) -> front-to-back alpha compositing (forward + backward)
 * -> SSE loss.  Two groups of entry points:
 *
 *  (1) COMPAT: the five symbols the reference hosts bind with ctypes
 *      (loma_public/compiler.py:262-276 sets their argtypes; train_nerf.py:212-213,
 *      fit_img.py:358-360 fetch them).  Same names, argument order, ragged float** and float***
 *      buffers, in-place outputs and accumulate-into-d_ semantics as the loma-generated C.
 *      Host pointers in, host pointers out, synchronous.
 *  (2) FLAT: contiguous-buffer entry points (device pointers, or pinned/pageable host pointers
 *      for the *_host variants) used by benchmarks, the Python mirror and multi-GPU sharding.
 *      The compat symbols are thin gather/scatter wrappers over these.
 *
 * No torch types, no C++ types: plain pointers, ints, floats and POD structs.
 * There is NO CPU fallback: every entry point fails (status != 0 / NaN loss) without a CUDA
 * device.  All arithmetic is IEEE fp32 unless lnb_step_args.path selects the tensor cores
 * (LNB_PATH_TC: bf16 operands, fp32 accumulation).
 */
#ifndef LOMA_NERF_B200_H
#define LOMA_NERF_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LNB_API __attribute__((visibility("default")))
#else
#define LNB_API
#endif

#define LNB_ABI_VERSION 4
#define LNB_MAX_LAYERS 16

/* ------------------------------------------------------------------------------------------
 * (1) COMPAT symbols
 * ------------------------------------------------------------------------------------------ */

/* replaces: loma function nerf_evaluate_and_march, /root/reference/scripts/nerf.py:1-304
 * (call site train_nerf.py:325-366, eval call site train_nerf.py:616-657).
 * Writes in place: intermediate_outputs (accumulated onto entry contents, then activated),
 * img_sample_rgba_arr, alpha, cumprod_alpha, weights_samples (overwritten), accumulated_color
 * (accumulated onto entry contents).  Returns the SSE loss; NaN on any CUDA failure. */
LNB_API float nerf_evaluate_and_march(float **layer_input, int layer_input_h, int layer_input_w,
                              float ***ws, float **bs, float **target_image, int target_image_h,
                              int target_image_w, int num_weights, int **weight_shapes,
                              int **bias_shapes, int **intermediate_output_shapes,
                              float ***intermediate_outputs, float ***img_sample_rgba_arr,
                              int num_samples, float **dists, float **alpha, float **cumprod_alpha,
                              float **weights_samples, float **accumulated_color);

/* replaces: grad_nerf_evaluate_and_march = rev_diff(nerf_evaluate_and_march),
 * scripts/nerf.py:306; signature rule loma_public/reverse_diff.py:504-517 (each In argument is
 * followed by its adjoint; int -> int*; trailing float _dreturn); call site train_nerf.py:395-478.
 * Accumulates into d_layer_input, d_ws, d_bs, d_target_image, d_dists, d_accumulated_color,
 * d_intermediate_outputs (pre-activation adjoints dZ_l); leaves the primal scratch arrays and all
 * int adjoints untouched; d_img_sample_rgba_arr, d_alpha, d_cumprod_alpha, d_weights_samples are
 * left as they were (the reference ends them at zero).  On CUDA failure d_ws is NaN-filled so the
 * host's NaN guard (train_nerf.py:486-489) trips. */
LNB_API void grad_nerf_evaluate_and_march(
    float **layer_input, float **d_layer_input, int layer_input_h, int *d_layer_input_h,
    int layer_input_w, int *d_layer_input_w, float ***ws, float ***d_ws, float **bs, float **d_bs,
    float **target_image, float **d_target_image, int target_image_h, int *d_target_image_h,
    int target_image_w, int *d_target_image_w, int num_weights, int *d_num_weights,
    int **weight_shapes, int **d_weight_shapes, int **bias_shapes, int **d_bias_shapes,
    int **intermediate_output_shapes, int **d_intermediate_output_shapes,
    float ***intermediate_outputs, float ***d_intermediate_outputs, float ***img_sample_rgba_arr,
    float ***d_img_sample_rgba_arr, int num_samples, int *d_num_samples, float **dists,
    float **d_dists, float **alpha, float **d_alpha, float **cumprod_alpha,
    float **d_cumprod_alpha, float **weights_samples, float **d_weights_samples,
    float **accumulated_color, float **d_accumulated_color, float _dreturn);

/* replaces: loma function mlp_fit, scripts/mlp_fit.py:1-147 (call site fit_img.py:515-530).
 * layer_output is unused by the reference and by us. */
LNB_API float mlp_fit(float **layer_input, int layer_input_h, int layer_input_w, float **layer_output,
              float ***ws, float **bs, float **target_image, int target_image_h,
              int target_image_w, int num_weights, int **weight_shapes, int **bias_shapes,
              int **intermediate_output_shapes, float ***intermediate_outputs);

/* replaces: grad_mlp_fit = rev_diff(mlp_fit), scripts/mlp_fit.py:174 (call site
 * fit_img.py:468-498). */
LNB_API void grad_mlp_fit(float **layer_input, float **d_layer_input, int layer_input_h,
                  int *d_layer_input_h, int layer_input_w, int *d_layer_input_w,
                  float **layer_output, float **d_layer_output, float ***ws, float ***d_ws,
                  float **bs, float **d_bs, float **target_image, float **d_target_image,
                  int target_image_h, int *d_target_image_h, int target_image_w,
                  int *d_target_image_w, int num_weights, int *d_num_weights, int **weight_shapes,
                  int **d_weight_shapes, int **bias_shapes, int **d_bias_shapes,
                  int **intermediate_output_shapes, int **d_intermediate_output_shapes,
                  float ***intermediate_outputs, float ***d_intermediate_outputs, float _dreturn);

/* replaces: loma function mult_a_b, scripts/mlp_fit.py:150-172 (required by fit_img.py:360-374).
 * c[i][j] += sum_k a[i][k] * b[k][j]. */
LNB_API void mult_a_b(float **a, int a_h, int a_w, float **b, int b_h, int b_w, float **c);

/* replaces: the forward-mode pair constructor the loma compiler emits into every library (autodiff.py:199-208;
 * `_dfloat make__dfloat(float val, float dval)` in the generated C).  Neither host script calls it; it is exported
 * so that the symbol table of the drop-in matches the generated one. */
typedef struct { float val; float dval; } _dfloat;
LNB_API _dfloat make__dfloat(float val, float dval);

/* ------------------------------------------------------------------------------------------
 * (2) FLAT API
 * ------------------------------------------------------------------------------------------ */

typedef struct lnb_ctx lnb_ctx; /* device, stream, workspace; one per (thread, GPU) */

enum { LNB_OK = 0, LNB_ERR_CUDA = 1, LNB_ERR_ARG = 2, LNB_ERR_UNSUPPORTED = 3 };
enum { LNB_HEAD_NERF = 0,     /* sigmoid on channels 0-2, ReLU on channel 3: nerf.py:147-167 */
       LNB_HEAD_SIGMOID = 1 };/* sigmoid on every channel:                  mlp_fit.py:121-132 */
enum { LNB_RAY_F64 = 0, LNB_RAY_F32 = 1 };
enum { LNB_SEED_VALUE = 0,    /* backward seed _dreturn = args->seed                        */
       LNB_SEED_LOSS = 1 };   /* backward seed = this step's loss (train_nerf.py:477)       */
enum { LNB_PATH_F32 = 0,      /* fp32 CUDA-core kernels, <=1e-5 of the reference: the fused  */
                              /* kernel when the problem fits it, else the layerwise kernels */
       LNB_PATH_TC = 1,       /* tcgen05 tensor-core kernels (bf16 operands, fp32 accumulate);*/
                              /* LNB_ERR_UNSUPPORTED when the problem does not fit them.      */
                              /* Weight gradients are summed in tensor memory in MMA          */
                              /* completion order: their last bits may differ between launches*/
                              /* (colours and losses are bit-stable; LNB_PATH_F32 is fully    */
                              /* reproducible)                                                */
       LNB_PATH_F32_LAYERWISE = 2 }; /* force the layerwise fp32 kernels (all intermediates)  */

/* The MLP: weights in the reference's padded layout (mlp_utils.py:272-313):
 * ws [n_layers][max_in][max_out] with ws[l][k][j] = weight from input k to output j
 * (mlp_utils.py:166-175), bs [n_layers][max_out]; dims[l] = in_l, dims[l+1] = out_l. */
typedef struct {
    int n_layers;
    int dims[LNB_MAX_LAYERS + 1];
    int max_in, max_out;
    int head; /* LNB_HEAD_* */
} lnb_mlp;

/* CAMERA MODE input (nerf only): rays AND sample depths are generated on the device from a pose, so nothing per ray or per
 * sample crosses the bus or sits in HBM -- the device-side form of get_rays (train_nerf.py:23-62) and of the sample
 * construction (train_nerf.py:289-311).  Ray r looks through pixel q = pixels ? pixels[r] : first_pixel + r of a
 * height x width grid, row = q / width, col = q % width, with normalised coordinates i = col / (width-1),
 * j = row / (width-1) (get_rays builds BOTH axes from linspace(0, 1, width), train_nerf.py:37-39);
 *   dir = ((i - cx) / fx, -(j - cy) / fy, -1) @ R^T  (not normalised),  origin = T,  [R | T] = c2w (3 x 4, row-major).
 * Sample depths: stratified == 0: t = linspace(near, far, S) (train_nerf.py:289, endpoint included);
 * stratified == 1: t_s = near + (s + u)(far - near) / S with u = lnb_uniform(seed, q, s) in [0, 1), the counter-based
 * generator documented at lnb_uniform below (SURVEY.md 8d's synthetic sampling; the reference has the jitter commented
 * out).  dists = t[s+1] - t[s], last 1e8.  The exact path forms every value in float64 like numpy does and rounds
 * once to float32; the tensor-core path computes in float32. */
typedef struct {
    double c2w[12];
    double fx, fy, cx, cy;        /* normalized_K[0][0], [1][1], [0][2], [1][2] */
    int width, height;
    long long first_pixel;
    const int *pixels;            /* optional [R] pixel indices (device pointer; host pointer for the *_host calls) */
    double near, far;
    int stratified;
    unsigned long long seed;
} lnb_camera;
/* u = (z >> 40) * 2^-24 with z = mix(seed + 0x9E3779B97F4A7C15 * (q * 4096 + s + 1)), mix = the SplitMix64 finaliser
 * (z ^= z >> 30; z *= 0xBF58476D1CE4E5B9; z ^= z >> 27; z *= 0x94D049BB133111EB; z ^= z >> 31).  Host-callable. */
LNB_API double lnb_uniform(unsigned long long seed, long long pixel, int s);

/* One forward(+backward) problem in the reference's layout.  N = R*S samples.  Pointers are
 * DEVICE pointers for lnb_nerf_step / lnb_fit_step and HOST pointers for the *_host variants.
 * Any output pointer may be NULL (not produced).  Accumulating outputs (+=) are marked. */
typedef struct {
    int R, S;                 /* rays (target rows) and samples per ray; mlp_fit: S = 1        */
    int n_rows;               /* rows of X actually multiplied (layer_input_h); 0 -> R*S        */
    int rows;                 /* rows the reference host declares in intermediate_output_shapes */
                              /* (>= n_rows; bias/activation run over them, nerf.py:95);0->n_rows*/
    int target_w;             /* columns of target (3)                                          */
    const float *X;           /* [n_rows][c_in] pre-encoded features                            */
    const float *ws, *bs;     /* padded weights / biases                                        */
    const float *target;      /* [R][target_w] or NULL (render: no loss)                        */
    const float *dists;       /* [R][S] (nerf only)                                             */
    /* forward outputs */
    float *inter;             /* [n_layers][inter_rows][inter_ld] post-activation layer outputs */
    int inter_rows, inter_ld; /*   (+= onto entry contents when inter_accumulate != 0)          */
    int inter_accumulate;
    float *rgba;              /* [R][S][4]  overwritten                                         */
    float *alpha, *cumprod, *weights; /* [R][S] overwritten                                     */
    float *color;             /* [R][3]  += (nerf.py:284-286) when color_accumulate, else =     */
    int color_accumulate;
    float *loss;              /* [1] overwritten                                                */
    /* backward (performed when want_grad != 0) */
    int want_grad;
    int seed_mode;            /* LNB_SEED_*                                                     */
    float seed;
    float *d_ws, *d_bs;       /* += padded layout                                               */
    float *d_X;               /* += [n_rows][c_in] or NULL                                      */
    float *d_target;          /* += [R][target_w] or NULL                                       */
    float *d_dists;           /* += [R][S] or NULL                                              */
    float *d_color;           /* += [R][3] or NULL  (d_accumulated_color)                       */
    float *d_inter;           /* += [n_layers][inter_rows][inter_ld] dZ_l, or NULL              */
    int path;                 /* LNB_PATH_*                                                     */
    /* RAYS MODE (nerf only; selected when X == NULL): the features are computed on the device  */
    /* from rays and sample depths exactly as the reference host does it: pts = o + d*t          */
    /* (train_nerf.py:289-299), X = positional_encoding_3d(pts, pe_bands) (pos_encoding.py:38-70),*/
    /* dists = [t[s+1]-t[s] ..., 1e8] (train_nerf.py:306-311).  `dists` is then ignored.         */
    const void *rays_o, *rays_d; /* [R][3]                                                      */
    const void *t;            /* [R][S] sample depths along each ray                            */
    int ray_dtype;            /* LNB_RAY_F64 (the reference's get_rays/linspace dtype) or F32   */
    int pe_bands;             /* E; dims[0] must equal 3 + 6E                                   */
    /* CAMERA MODE (selected when X == NULL, rays_o == NULL and cam != NULL): see lnb_camera.      */
    /* `cam` is always a HOST pointer (a few bytes, passed to the kernels by value).               */
    const lnb_camera *cam;
} lnb_step_args;

LNB_API int lnb_abi_version(void);
/* Fills out[0..n) with {sizeof(lnb_mlp), sizeof(lnb_step_args), offsetof(lnb_step_args, X),
 * inter, rgba, loss, want_grad, d_ws, path, rays_o, pe_bands, cam, sizeof(lnb_camera), offsetof(lnb_camera, pixels)};
 * returns how many values exist.  Lets a foreign-
 * language binding verify its struct mirror without a GPU. */
LNB_API int lnb_struct_layout(int *out, int n);
LNB_API int lnb_device_count(void);
/* device < 0: the current CUDA device. */
LNB_API int lnb_create(lnb_ctx **out, int device);
LNB_API void lnb_destroy(lnb_ctx *ctx);
/* stream is a cudaStream_t: NULL = CUDA's legacy default stream, (void*)-1 = the context's own
 * non-blocking stream (the initial setting). */
LNB_API int lnb_set_stream(lnb_ctx *ctx, void *stream);
LNB_API int lnb_synchronize(lnb_ctx *ctx);
LNB_API const char *lnb_last_error(lnb_ctx *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
LNB_API long long lnb_launch_count(lnb_ctx *ctx);
/* Kernel timing for roofline reports: while enabled, every launch of the dominant (fused)
 * kernel is bracketed by CUDA events on the context's stream.  lnb_profile_read synchronises and
 * returns the summed duration, the launch count and the kernel's name. */
LNB_API int lnb_profile(lnb_ctx *ctx, int enable);
LNB_API int lnb_profile_read(lnb_ctx *ctx, double *ms_total, long long *launches, char *name, int name_len);
/* page-locked host memory for the *_host entry points (they copy straight from / to such
 * buffers; pageable buffers bounce through the context's own pinned staging). */
LNB_API void *lnb_host_alloc(size_t bytes);
LNB_API void lnb_host_free(void *p);

/* nerf path: MLP -> compositing -> loss [-> backward].  Device pointers, asynchronous on the
 * context's stream.  Follows scripts/nerf.py:67-304 and its reverse. */
LNB_API int lnb_nerf_step(lnb_ctx *ctx, const lnb_mlp *mlp, const lnb_step_args *args);
/* mlp_fit path: MLP -> loss [-> backward].  scripts/mlp_fit.py:39-147. */
LNB_API int lnb_fit_step(lnb_ctx *ctx, const lnb_mlp *mlp, const lnb_step_args *args);
/* Same, HOST pointers (pinned or pageable): stages H2D, runs, stages D2H, synchronises. */
LNB_API int lnb_nerf_step_host(lnb_ctx *ctx, const lnb_mlp *mlp, const lnb_step_args *args);
LNB_API int lnb_fit_step_host(lnb_ctx *ctx, const lnb_mlp *mlp, const lnb_step_args *args);

/* Positional encoding, pos_encoding.py:4-70: x [n][F] float64 -> out [n][F*(1+2E)] float32,
 * feature = slot*F + coord, slot 0 identity, 2i+1 sin(2^i x), 2i+2 cos(2^i x); float64 range
 * reduction so the result matches the reference's float64-then-cast values. Device pointers. */
LNB_API int lnb_pos_encoding(lnb_ctx *ctx, const double *x, long long n, int F, int E, float *out);
/* Sample generation + encoding, train_nerf.py:289-311 + pos_encoding.py:38-70:
 * pts = o + d*t (float64), X = PE(pts) float32 [R*S][3+6E], dists [R][S] = t[s+1]-t[s], last 1e8.
 * rays_o, rays_d [R][3] float64, t [R][S] float64. Device pointers. */
LNB_API int lnb_sample_encode(lnb_ctx *ctx, const double *rays_o, const double *rays_d, const double *t,
                      int R, int S, int E, float *X, float *dists);

/* The rays and sample depths camera mode generates, written out (any pointer may be NULL): rays_o, rays_d [R][3], t [R][S]
 * float64.  Device pointers (cam itself is a host pointer). */
LNB_API int lnb_camera_rays(lnb_ctx *ctx, const lnb_camera *cam, int R, int S, double *rays_o, double *rays_d, double *t);

/* out[i] = (unsigned char) rint(255 * clamp(rgb[i], 0, 1)), i < n: what the hosts do before writing a PNG / video frame
 * (train_nerf.py:686-700), so that a rendered frame leaves the device as 3 bytes per pixel.  Device pointers. */
LNB_API int lnb_color_to_u8(lnb_ctx *ctx, const float *rgb, long long n, unsigned char *out);

/* c[a_h][b_w] += a[a_h][a_w] * b[a_w][b_w]; device pointers (scripts/mlp_fit.py:150-172). */
LNB_API int lnb_mult_a_b(lnb_ctx *ctx, const float *a, int a_h, int a_w, const float *b, int b_w, float *c);

/* Optimisers the reference hosts apply to the padded arrays, n floats each.
 * lnb_adam_step reproduces AdamOptimizer.update (train_nerf.py:133-161) including its double
 * bias correction; t is the 1-based step count AFTER increment.  lnb_sgd_step: p -= lr * g
 * (fit_img.py:512-513).  Device pointers.  The hyper-parameters are doubles because the
 * reference holds them as Python floats and derives (1-beta), the bias corrections and lr_t in
 * double before they meet the float32 arrays. */
LNB_API int lnb_adam_step(lnb_ctx *ctx, float *param, const float *grad, float *m, float *v, long long n,
                  int t, double lr, double beta1, double beta2, double eps);
/* Same update with the step counter kept on the device (t = *t_dev + 1, then *t_dev += 1), so a
 * whole train step can be captured in a CUDA graph and replayed. */
LNB_API int lnb_adam_step_dev(lnb_ctx *ctx, float *param, const float *grad, float *m, float *v, long long n,
                      int *t_dev, double lr, double beta1, double beta2, double eps);
LNB_API int lnb_sgd_step(lnb_ctx *ctx, float *param, const float *grad, long long n, double lr);

/* ------------------------------------------------------------------------------------------
 * Device-resident training state (the hosts' per-chunk "grad call, then optimiser on the padded
 * arrays" loop, train_nerf.py:395-499 / fit_img.py:468-513, without leaving the GPU).
 * A batch is an lnb_step_args with DEVICE pointers; its ws, bs, d_*, loss and by-product fields are
 * ignored (the trainer supplies them).  lnb_trainer_step = gradient + optimiser update (two kernel
 * launches on the tensor-core path).  For data-parallel training split it:
 * lnb_trainer_grad -> all-reduce lnb_trainer_grad_buffer ([d_ws | d_bs | loss], written, not
 * accumulated) over the ranks -> lnb_trainer_apply.
 * ------------------------------------------------------------------------------------------ */
typedef struct lnb_trainer lnb_trainer;
enum { LNB_OPT_ADAM = 0,  /* AdamOptimizer.update, train_nerf.py:133-161 (double bias correction) */
       LNB_OPT_SGD = 1 }; /* p -= lr * g, fit_img.py:512-513                                      */
LNB_API int lnb_trainer_create(lnb_ctx *ctx, const lnb_mlp *mlp, const float *ws_host, const float *bs_host,
                       int optimizer, double lr, double beta1, double beta2, double eps, lnb_trainer **out);
LNB_API void lnb_trainer_destroy(lnb_trainer *t);
LNB_API int lnb_trainer_step(lnb_trainer *t, const lnb_step_args *batch, int nerf);
/* the same step with the batch in HOST memory (pinned buffers are copied from directly): stages
 * host->device, steps, returns the loss through *loss_out; synchronous */
LNB_API int lnb_trainer_step_host(lnb_trainer *t, const lnb_step_args *batch, int nerf, float *loss_out);
/* PIPELINED host-buffer steps (the hosts' chunk loop, train_nerf.py:275-499, with the copy of batch i+1 under the step
 * of batch i): lnb_trainer_submit_host stages the batch on a copy stream into one of two device staging slots, enqueues the
 * step behind it and returns without waiting for the GPU.  Pinned host buffers are read by the DMA engine directly and must
 * stay unchanged until the submission after next has returned (or until lnb_trainer_wait); pageable buffers are copied to a
 * pinned bounce slot before the call returns.  lnb_trainer_wait blocks until every submitted step has finished and copies
 * the losses of the steps submitted since the previous wait, oldest first, into losses[0 .. *n_out) (at most max_losses of
 * the newest 4096; losses may be NULL).  Errors of a poisoned peer exchange surface at lnb_trainer_wait. */
LNB_API int lnb_trainer_submit_host(lnb_trainer *t, const lnb_step_args *batch, int nerf);
LNB_API int lnb_trainer_wait(lnb_trainer *t, float *losses, int max_losses, int *n_out);
LNB_API int lnb_trainer_grad(lnb_trainer *t, const lnb_step_args *batch, int nerf);
LNB_API int lnb_trainer_apply(lnb_trainer *t);
LNB_API float *lnb_trainer_grad_buffer(lnb_trainer *t, long long *n_floats);
LNB_API float *lnb_trainer_params(lnb_trainer *t, long long *n_w, long long *n_b);
/* Peer-memory gradient all-reduce for the ranks of one box (one process per GPU), fused into the
 * tensor-core step's reduction kernel: every rank exports a 64-byte CUDA IPC handle of its exchange
 * buffer, the host exchanges the handles (any transport), every rank attaches all `world` handles
 * (rank-major, 64 bytes each).  From then on lnb_trainer_step[_host] sums the gradients (and the
 * loss) of all ranks over NVLink inside its second kernel -- no separate collective -- and all ranks
 * must step in lockstep; with LNB_SEED_LOSS the gradient is (sum of losses) x (sum of unit-seed
 * gradients), the reference's gradient of the whole batch.  Steps that cannot exchange in-kernel
 * (fp32 path, wide MLPs, S > 128) return LNB_ERR_UNSUPPORTED instead of updating from local
 * gradients.  A peer that stays silent for LNB_PEER_TIMEOUT_MS (default 5000) poisons that step:
 * gradients, parameters and the returned loss become NaN, lnb_trainer_step_host / lnb_trainer_read
 * return LNB_ERR_CUDA and lnb_trainer_comm_status returns 1 (0 = fine). */
LNB_API int lnb_trainer_comm_export(lnb_trainer *t, void *handle64);
LNB_API int lnb_trainer_comm_attach(lnb_trainer *t, int rank, int world, const void *handles);
LNB_API int lnb_trainer_comm_status(lnb_trainer *t);
/* synchronises; any pointer may be NULL */
LNB_API int lnb_trainer_read(lnb_trainer *t, float *ws_host, float *bs_host, float *loss_host);

/* Process-wide default context used by the compat symbols (created lazily on first call,
 * device from LOMA_NERF_B200_DEVICE or 0). Returns NULL when no CUDA device is usable. */
LNB_API lnb_ctx *lnb_default_ctx(void);

#ifdef __cplusplus
}
#endif
#endif /* LOMA_NERF_B200_H */
