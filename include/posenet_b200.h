/*
 * posenet_b200.h -- C ABI of libposenet_b200.so: the B200 (sm_100a) PoseNet inference hot path
 *
 *     preprocess -> MobileNetV1 backbone + 4 heads -> part candidates -> greedy multi-pose decode
 *
 * The reference (michellelychan/posenet-pytorch) is pure Python and has no FFI / plugin /
 * operator interface; its boundary for this path is the Python API.  Each entry point below
 * therefore names the reference *function* (file:line, relative to the reference checkout)
 * whose arithmetic it replaces; posenet-pytorch_b200/posenet/ binds them with ctypes behind the
 * reference's own names.
 *
 * Conventions
 *   - Plain pointers and sizes only.  Every pointer is a DEVICE pointer unless its name ends
 *     in _host.  The library never allocates device memory: the caller owns all buffers.
 *   - All work is enqueued on `stream` (a cudaStream_t); nothing synchronises.
 *   - Return value: PN_OK (0) or a negative PN_ERR_*; pn_last_error_string() describes the
 *     last failure on the calling thread.
 *   - Activations are NHWC ("pixel-major": [M = n*h*w, C]); `dtype` selects fp32 or bf16
 *     storage.  Head tensors and everything downstream are fp32 NCHW like the reference's.
 */
#ifndef POSENET_B200_H
#define POSENET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PN_ABI_VERSION 5

typedef void *pn_stream_t; /* cudaStream_t */

enum { PN_OK = 0, PN_ERR_ARG = -1, PN_ERR_CUDA = -2, PN_ERR_UNSUPPORTED = -3, PN_ERR_NO_DEVICE = -4 };
enum { PN_F32 = 0, PN_BF16 = 1 };

#define PN_NUM_PARTS 17
#define PN_NUM_EDGES 16
#define PN_HEAD_CHANNELS 115 /* 17 heatmap + 34 offset + 32 fwd + 32 bwd */
#define PN_HEAD_ROWS 128     /* packed head weight rows (115 padded to a UMMA-friendly 128) */

int pn_abi_version(void);
const char *pn_last_error_string(void);
/* PN_OK when the current CUDA device is an sm_100 part the kernels were built for. */
int pn_device_check(void);

/* ---- P1: posenet/utils.py:13-26 (_process_input; cv2.resize INTER_LINEAR + BGR->RGB + x*(2/255)-1)
 * src: uint8 [n, src_h, src_w, 3] BGR HWC.  dst: f32 [n, 3, dst_h, dst_w] RGB NCHW.  Bit-exact with cv2. */
int pn_preprocess_u8(const uint8_t *src, int n, int src_h, int src_w, int dst_h, int dst_w,
                     float *dst, pn_stream_t stream);
/* The resize stage of P1 alone (utils.py:21, cv2.resize INTER_LINEAR on uint8, bit-exact): uint8 [n, src_h, src_w, 3]
 * -> uint8 [n, dst_h, dst_w, 3], still BGR HWC -- the input format of pn_stem_conv_u8 / a uint8 plan, so frames of any
 * size (utils.py:51-55 read_cap at 1280x720) reach the uint8 fast path without a float round trip. */
int pn_resize_u8(const uint8_t *src, int n, int src_h, int src_w, int dst_h, int dst_w, uint8_t *dst,
                 pn_stream_t stream);

/* ---- B2: posenet/models/mobilenet_v1.py:47-54 (InputConv: relu6(conv3x3(x, stride, pad 1) + b), 3 -> cout)
 * x: f32 NCHW [n,3,h,w].  w: f32 [27, cout], row = (ky*3+kx)*3+ci.  y: NHWC [n,ho,wo,cout] of out_dtype. */
int pn_stem_conv(const float *x, const float *w, const float *bias, void *y, int n, int h, int wd,
                 int cout, int stride, int out_dtype, pn_stream_t stream);

/* ---- B2': P1 fused into B2 for already-sized images (identity resize): uint8 BGR HWC [n,h,w,3] in. */
int pn_stem_conv_u8(const uint8_t *img, const float *w, const float *bias, void *y, int n, int h, int wd,
                    int cout, int stride, int out_dtype, pn_stream_t stream);

/* ---- B3: mobilenet_v1.py:60-62,66 (SeperableConv.depthwise + relu6; stride 1|2, dilation 1|2|4,
 * padding ((s-1)+2d)/2 per mobilenet_v1.py:42-44).  x,y: NHWC of dtype.  w: f32 [9, c] tap-major. */
int pn_dwconv3x3(const void *x, const float *w, const float *bias, void *y, int n, int h, int wd, int c,
                 int stride, int dilation, int dtype, pn_stream_t stream);

/* ---- B4: mobilenet_v1.py:63,67 (SeperableConv.pointwise + relu6) == GEMM + bias + clamp(0,6)
 * a: [m,k] dtype.  w: [n,k] dtype (the OIHW 1x1 weight as stored).  bias: f32 [n].  y: [m,n] dtype.
 * bf16 runs on tcgen05 tensor cores (TMA-fed, TMEM accumulators); fp32 is the FFMA parity path. */
int pn_pwconv_gemm(const void *a, const void *w, const float *bias, void *y, int m, int k, int n,
                   int dtype, pn_stream_t stream);

/* ---- B3+B4 fused: mobilenet_v1.py:57-68 (SeperableConv.forward: relu6(pointwise(relu6(depthwise(x))))) as ONE
 * kernel -- the depthwise result is produced straight into the shared-memory A tile of the tcgen05 GEMM and never
 * reaches HBM.  bf16 only.  x: NHWC [n,h,wd,cin]; dw_w f32 [9,cin]; dw_b f32 [cin]; pw_w bf16 [cout,cin];
 * pw_b f32 [cout]; y: NHWC [n,ho,wo,cout].  cin % 8 == 0, cout % 16 == 0, (stride,dilation) in (1,1|2|4), (2,1). */
int pn_sepconv_block(const void *x, const float *dw_w, const float *dw_b, const void *pw_w, const float *pw_b, void *y,
                     int n, int h, int wd, int cin, int cout, int stride, int dilation, pn_stream_t stream);
/* Tile shape / pipeline depths pn_sepconv_block would pick for a block (host arithmetic only; diagnostics). */
int pn_sepconv_describe(int n, int h, int wd, int cin, int cout, int stride, int dilation, char *out_host, int capacity);

/* ---- H1: mobilenet_v1.py:151-154,158-161 (4 head convs + sigmoid on the heatmap) as ONE GEMM.
 * a: [n_img*hw, k] dtype.  w: [PN_HEAD_ROWS, k] dtype, rows heat|offset|fwd|bwd then zero padding.
 * Outputs f32 NCHW: heat [n_img,17,hw], off [n_img,34,hw], fwd/bwd [n_img,32,hw]. */
int pn_heads_gemm(const void *a, const void *w, const float *bias, float *heat, float *off, float *fwd,
                  float *bwd, int n_img, int hw, int k, int dtype, pn_stream_t stream);

/* A strided f32 view [n_img, channels, h, w]; strides in elements (NCHW or channels-last alike). */
typedef struct pn_map {
    const float *ptr;
    int64_t s_img, s_ch, s_y, s_x;
} pn_map;

/* ---- C1: posenet/decode_multi.py:27-34 (build_part_with_score_torch: 3x3 local max, -inf padding,
 * score >= fp32(threshold)).  keys: [n_img, capacity] 64-bit sort keys (high word orders by score
 * descending, low word = flat (part,y,x) index), UNORDERED within an image; counts: int32 [n_img].
 * capacity >= 17*h*w guarantees no candidate is dropped (counts never exceed capacity). */
int pn_candidates(const pn_map *heat, int n_img, int h, int wd, float score_threshold, uint64_t *keys,
                  int capacity, int *counts, pn_stream_t stream);

typedef struct pn_decode_params {
    int output_stride;
    int max_pose_detections;
    double squared_nms_radius; /* nms_radius ** 2, as the reference computes it on the host */
    double min_pose_score;
} pn_decode_params;

/* ---- D1-D4: decode_multi.py:61-148 + decode.py:9-63,131-182 (greedy multi-pose decode, float64).
 * One thread block per image; candidates are consumed in (score desc, flat index asc) order.
 * keys / counts are pn_candidates' outputs: the keys of an image must be pairwise distinct (they are -- the
 * low word is the cell index) and non-zero; the in-kernel radix selection relies on it.
 * Outputs (float64): pose_scores [n_img,P], kp_scores [n_img,P,17], kp_coords [n_img,P,17,2] (y,x),
 * kp_offsets [n_img,P,17,2]; pose_counts int32 [n_img].  The buffers need not be initialised: the rows
 * past pose_counts[i] are zero-padded by the kernel (the reference's np.zeros, decode_multi.py:94-100). */
int pn_decode_greedy(const pn_map *heat, const pn_map *off, const pn_map *fwd, const pn_map *bwd,
                     int n_img, int h, int wd, const uint64_t *keys, int capacity, const int *counts,
                     const pn_decode_params *params, double *pose_scores, double *kp_scores,
                     double *kp_coords, double *kp_offsets, int *pose_counts, pn_stream_t stream);

/* ---- D3: posenet/decode.py:9-63 (traverse_to_targ_keypoint) as a stand-alone call: ONE displacement hop along edge
 * `edge_id` from `source_keypoint_host` (float64 (y, x) image coordinates, HOST pointer) to part `target_keypoint_id`.
 * heat [1,17,h,w], off [1,34,h,w], disp [1,32,h,w] (channels 0..15 dy, 16..31 dx), any strides.
 * out7 (device, float64): score, image_coord y, x, displacement_vector y, x, offset y, x -- the fp32 values widened exactly. */
int pn_traverse_to_targ_keypoint(int edge_id, const double *source_keypoint_host, int target_keypoint_id,
                                 const pn_map *heat, const pn_map *off, const pn_map *disp, int h, int wd,
                                 int output_stride, double *out7, pn_stream_t stream);

/* ---- D2: posenet/decode.py:131-182 (decode_pose) as a stand-alone call for ONE root: backward edges 15..0 through `bwd`, then
 * forward edges 0..15 through `fwd`; a hop needs score[source] > 0.0 and an undecoded (== 0.0) target.
 * root_image_coord_host: float64 (y, x), HOST pointer.  out85 (device, float64): instance_keypoint_scores [17],
 * instance_keypoint_coords [17,2], instance_offsets [17,2] (the root's offset row stays 0, decode.py:150). */
int pn_decode_pose(double root_score, int root_id, const double *root_image_coord_host, const pn_map *heat,
                   const pn_map *off, const pn_map *fwd, const pn_map *bwd, int h, int wd, int output_stride,
                   double *out85, pn_stream_t stream);

/* ---- N4: image_demo.py:50 (`keypoint_coords *= output_scale`, output_scale = (src_h / target_h, src_w / target_w) of
 * utils.py:19) for a batch of pose records ON THE DEVICE: kp_coords float64 [n_img, points_per_img, 2] (y, x), multiplied in
 * place (IEEE float64 multiply, what numpy does).  scales: device float64 [n_img, 2] with one (y, x) factor per image (frames
 * of different sizes), or NULL for the uniform (scale_y, scale_x). */
int pn_scale_keypoint_coords(double *kp_coords, int n_img, int points_per_img, const double *scales, double scale_y,
                             double scale_x, pn_stream_t stream);

/* ---- Whole-network plan: mobilenet_v1.py:130-162 (MobileNetV1.__init__/forward) --------------------
 * The layer table is computed by the caller (posenet/models/mobilenet_v1.py mirrors
 * _to_output_strided_layers, mobilenet_v1.py:8-39) and passed in explicitly. */
typedef struct pn_layer {
    int cin, cout, stride, dilation;
    const float *dw_w, *dw_b; /* [9,cin], [cin]   (NULL for the stem) */
    const void *pw_w;         /* stem: f32 [27,cout]; blocks: [cout,cin] of plan dtype */
    const float *pw_b;        /* [cout] */
} pn_layer;

typedef struct pn_net_desc {
    int dtype;       /* PN_F32 | PN_BF16: activation + pointwise weight storage */
    int n, h, w;     /* batch and input size */
    int input_u8;    /* 0: f32 NCHW input (reference forward()); 1: uint8 BGR HWC of size h x w */
    int num_layers;  /* 14 */
    pn_layer layers[16];
    const void *head_w;    /* [PN_HEAD_ROWS, c_last] plan dtype */
    const float *head_b;   /* [PN_HEAD_ROWS] */
    int flags;             /* PN_PLAN_* */
} pn_net_desc;

/* bf16 plans run every SeperableConv block as one fused kernel (pn_sepconv_block) by default;
 * PN_PLAN_UNFUSED keeps depthwise and pointwise as two kernels (pn_dwconv3x3 + pn_pwconv_gemm). */
#define PN_PLAN_UNFUSED 1

typedef struct pn_plan pn_plan;

/* Bytes of activation arena pn_plan_create needs for `desc` (two ping-pong buffers), and the
 * head map size (out_h, out_w). */
int pn_plan_query(const pn_net_desc *desc, size_t *arena_bytes, int *out_h, int *out_w);
int pn_plan_create(const pn_net_desc *desc, void *arena, size_t arena_bytes, pn_plan **plan);
/* input: f32 NCHW or uint8 HWC as desc->input_u8 says.  Enqueues the kernels of one forward (15 fused, 28 unfused). */
int pn_plan_forward(pn_plan *plan, const void *input, float *heat, float *off, float *fwd, float *bwd,
                    pn_stream_t stream);
/* Same launches as pn_plan_forward with a CUDA event between consecutive kernels; synchronises `stream`
 * and writes the device time of each launch (ms, launch order; see pn_plan_launch_name)
 * into ms_host[0 .. pn_plan_num_launches).  Measurement aid for bench.py's roofline; host pointer. */
int pn_plan_profile(pn_plan *plan, const void *input, float *heat, float *off, float *fwd, float *bwd,
                    float *ms_host, int capacity, pn_stream_t stream);
/* number of kernel launches one pn_plan_forward enqueues */
int pn_plan_num_launches(const pn_plan *plan);
/* name of launch i in launch order: "stem", "dw3" / "pw3" (unfused), "sep3" (fused block 3), "heads" */
const char *pn_plan_launch_name(const pn_plan *plan, int i);
int pn_plan_destroy(pn_plan *plan);

#ifdef __cplusplus
}
#endif
#endif /* POSENET_B200_H */
