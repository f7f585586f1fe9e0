/*
 * posenet_b200_diag.h -- C ABI of libposenet_b200_diag.so: hardware probes used while developing the tensor-pipe
 * depthwise path (csrc/septc.cu).  NOT part of the product library: libposenet_b200.so (include/posenet_b200.h) exports
 * the hot path only; these probes are built into a second library by posenet-pytorch_b200/build.py and are bound by
 * tests/abi.py (`load_diag`) and tools/ only.
 *
 * Timeline traces of the product kernels are compile-time options of the product sources, exported only by a
 * diagnostics build of libposenet_b200.so: -DPN_SEP_TRACE (pn_debug_sep_trace, csrc/sepconv.cu, tools/trace_sep.py) and
 * -DPN_TCS_TRACE (pn_debug_tcs_trace, csrc/septc.cu, tools/trace_septc.py).
 */
#ifndef POSENET_B200_DIAG_H
#define POSENET_B200_DIAG_H

#include "posenet_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

const char *pn_diag_last_error_string(void);

/* Diagnostic for the tensor-pipe depthwise (csrc/diag/dwtc_probe.cu): one 128-position chunk of one 64-channel bf16 image
 * [h, wd, 64] through shifted-descriptor tcgen05 depthwise (block-diagonal tap tiles `diag` [9*16, 64]) and an
 * A-from-TMEM pointwise (`pw_w` [64, 64]); out_dw / out_pw: f32 [128, 64].  Not part of the product path. */
int pn_dwtc_probe(const void *x, int h, int wd, const void *diag, const void *pw_w, const float *dw_bias, float *out_dw,
                  float *out_pw, int wp, int dil, int qoff, int rows_box, int x_org, int y_org, int flags, pn_stream_t stream);
/* pn_debug_umma_cost times 9 * reps tcgen05.mma (M128 x n x K16, bf16) on one SM; layout
 * 0 / 1 / 2 = 128 / 32 / 64-byte swizzle issued by one thread, 3 = 128-byte swizzle issued from warp-uniform code;
 * out_host[0] = cycles until the last issue, out_host[1] = until completion (synchronous, default stream). */
int pn_debug_umma_cost(int n, int layout, int reps, int a_step16, long long *out_host);

#ifdef __cplusplus
}
#endif
#endif /* POSENET_B200_DIAG_H */
