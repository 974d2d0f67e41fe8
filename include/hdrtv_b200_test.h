/* hdrtv_b200_test.h — debug, self-test and micro-probe entry points of the TEST build of the engine
 * (libhdrtv_b200_test.so = the same sources compiled with -DHDRTV_TEST_EXPORTS: the whole product ABI of hdrtv_b200.h plus
 * the symbols below).  The product library libhdrtv_b200.so does not export them.  Used by tests/ and scripts/ only.
 */
#ifndef HDRTV_B200_TEST_H
#define HDRTV_B200_TEST_H
#include "hdrtv_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Parity hook (FP32 contexts): one named conv / linear layer (bias, no activation, the layer's input quantiser) on    */
/* host data: in (Cin,H,W) fp32 -> out (Cout,Ho,Wo) fp32.                                                             */
int hdrtv_debug_layer(hdrtv_t* h, const char* layer, const float* in_host, int cin, int height, int width, int stride,
                      float* out_host);

/* One W8A8 layer through its tcgen05.mma.kind::i8 launch on uint8 codes q [C][H][W] (host): raw S32 accumulators [N][Ho][Wo]  */
/* and the de-quantised output (conv + bias, no activation; [Cout/4][2Ho][2Wo] after PixelShuffle for the up-convs).          */
int hdrtv_debug_conv_i8(hdrtv_t* h, const char* layer, const uint8_t* q_host, int c, int height, int width, int stride,
                        int32_t* acc_host, float* out_host);
int hdrtv_debug_tensor_count(const hdrtv_t* h);
int hdrtv_debug_tensor_info(const hdrtv_t* h, int idx, char* name, int name_cap, int* c, int* height, int* width);
int hdrtv_debug_tensor_read(hdrtv_t* h, int idx, float* dst_host); /* (C,H,W) fp32; synchronises                    */
/* One convolution through both paths on random data: returns max |tcgen05(fp16) - cuda-core(fp32)| in *max_abs.   */
int hdrtv_conv_selftest(hdrtv_t* h, int kind, int cin, int cout, int height, int width, int flags, float* max_abs,
                        float* ref_max);
/* tcgen05 issue-rate probe (design evidence): cycles per M=128 x n x K=16 MMA; layout 0/1 = SWIZZLE_NONE (plane     */
/* pitch / dense), 2 = SWIZZLE_128B; `blocks` concurrent CTAs.                                                      */
int hdrtv_mma_probe(hdrtv_t* h, int n, int layout, int vary, int iters, int blocks, int n_accumulators,
                    float* cycles_per_mma);
/* Micro-probes of the tensor path (csrc/probes.cuh): 0 MMA SS M=128, 1 MMA with A in TMEM, 2 MMA SS M=64,             */
/* 3 tcgen05.ld throughput (nwarps warps x 64 columns), 4 layer-chain round trip (nmma K-steps, `groups` row slots).   */
/* 5 free-running MMA + commit stream.  trace_host (optional, 256 entries): clock64 stamps of probe 4's first 16      */
/* iterations, [iter][group<4][event<4] = issue start, after commit, epilogue woke, epilogue arrived.                  */
int hdrtv_probe(hdrtv_t* h, int kind, int n, int iters, int blocks, int nwarps, int nmma, int groups,
                float* cycles_per_iter, long long* trace_host);
/* Debug timeline of one fused layer-chain launch (after an hdrtv_infer at the current size): clock64 stamps of CTA 0, */
/* [step < 64][row slot < 8][8].                                                                                        */
int hdrtv_chain_trace(hdrtv_t* h, int agcm_plan, int launch_index, long long* trace_host);

#ifdef __cplusplus
}
#endif
#endif /* HDRTV_B200_TEST_H */
