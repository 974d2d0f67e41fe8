/* hdrtv_b200.h — C ABI of the B200-native HDRTVNet++ per-frame SDR->HDR engine.
 *
 * The reference (DanHelmy/hdr-realtime-video-pipeline) has NO native/FFI boundary of its own: its boundary is the
 * Python class HDRTVNetTorch / HDRTVNetTensorRT (src/models/hdrtvnet_torch.py:1513, :8164).  Each entry point below
 * names the reference method whose device work it replaces; the Python mirror of that class
 * (hdr_realtime_video_pipeline_b200/backend.py) binds these symbols with ctypes.
 *
 * Conventions: plain pointers and sizes only; every call returns 0 on success or a negative code, never throws;
 * no hidden synchronisation — work is enqueued on the caller's stream (a cudaStream_t passed as void*);
 * all tensor pointers are caller-owned device (or mapped pinned host) memory; the engine owns only its workspace.
 */
#ifndef HDRTV_B200_H
#define HDRTV_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hdrtv_ctx hdrtv_t;

enum { HDRTV_FP32 = 0, HDRTV_FP16 = 1 };               /* arithmetic/storage type of x, cond, out tensors        */
enum { HDRTV_COND_BICUBIC_AA = 0, HDRTV_COND_ZERO = 1, HDRTV_COND_BILINEAR = 2 }; /* hdrtvnet_torch.py:2265-2294      */
enum { HDRTV_TRANSFER_IDENTITY = 0, HDRTV_TRANSFER_LUT = 1 };

typedef struct {
  int device;    /* CUDA ordinal                                                                                   */
  int precision; /* HDRTV_FP32 (CUDA-core path, <=1e-4 vs reference fp32) | HDRTV_FP16 (tcgen05 path, <=2e-3)      */
} hdrtv_config;

typedef struct {
  const char* name;   /* state-dict key after "module." stripping (hdrtvnet_torch.py:2154-2157), e.g. "LE.HR_conv1.weight" */
  const float* data;  /* host, fp32, contiguous                                                                     */
  int ndim;
  int64_t shape[4];
} hdrtv_tensor_desc;

/* HDRTVNetTorch.__init__ / _load_model (hdrtvnet_torch.py:1532-1673, 2044-2169): create a context on a device,   */
/* then hand it the Ensemble_AGCM_LE state-dict (264 tensors); weights are repacked once into kernel-native layouts. */
int hdrtv_create(const hdrtv_config* cfg, hdrtv_t** out);
void hdrtv_destroy(hdrtv_t* h);
int hdrtv_set_weights(hdrtv_t* h, const hdrtv_tensor_desc* tensors, int n);

/* INT8 layouts (W8A8Conv2d / W8A8Linear, hdrtvnet_torch.py:296-410; loader :1748-1963): the reference's eager INT8     */
/* model is fake-quantisation — int8 weights de-quantised per output channel (pass them to hdrtv_set_weights already  */
/* de-quantised), and a static per-tensor quantiser on every layer INPUT.  This call installs those input quantisers  */
/* (mode 0 none, 1 symmetric [-128,127], 2 asymmetric [0,255] with zero point) by layer name ("LE.HR_conv1", ...).    */
/* Supported on HDRTV_FP32 contexts (CUDA-core path).                                                                  */
int hdrtv_set_act_quant(hdrtv_t* h, const char* const* layers, const float* scales, const float* zeros, const int* modes,
                        int n);

/* HDRTVNetTorch._ensure_buffers (hdrtvnet_torch.py:2198-2233): (re)allocate the per-resolution workspace.          */
/* Called implicitly by preprocess/infer; returns bytes held via hdrtv_workspace_bytes.                              */
int hdrtv_prepare(hdrtv_t* h, int height, int width);
size_t hdrtv_workspace_bytes(const hdrtv_t* h);

/* HDRTVNetTorch.preprocess (hdrtvnet_torch.py:2239-2296): bgr = uint8 HxWx3 BGR on the device; x_out = (1,3,H,W)   */
/* RGB planar, cond_out = (1,3,H/4,W/4), both of the context's precision.                                           */
int hdrtv_preprocess(hdrtv_t* h, const uint8_t* bgr, int height, int width, void* x_out, void* cond_out, int cond_mode,
                     void* stream);

/* HDRTVNetTorch.infer / Ensemble_AGCM_LE.forward (hdrtvnet_torch.py:2302-2346, Ensemble_AGCM_LE_arch.py:889-897;  */
/* TensorRT variant :8992-9106): x, cond as produced by preprocess; out, agcm_out = (1,3,H,W) planar.               */
int hdrtv_infer(hdrtv_t* h, const void* x, const void* cond, int height, int width, void* out, void* agcm_out,
                void* stream);

/* The same inference split for frame pipelining: hdrtv_classify runs the input-only part of the network — staging x */
/* into the tensor-core layout and the AGCM condition classifier + GFM fold (Condition_arch.py:559-569), which depend  */
/* on x / cond alone — so it can run on a side stream while the previous frame's LE network still occupies the GPU;    */
/* hdrtv_infer_ex(skip_classifier = 1) then runs everything else for exactly those x / cond.                           */
/* inputs_consumed_event (cudaEvent_t, optional) is recorded once x, cond and the classifier results have been read. */
int hdrtv_classify(hdrtv_t* h, const void* x, const void* cond, int height, int width, void* stream);
int hdrtv_infer_ex(hdrtv_t* h, const void* x, const void* cond, int height, int width, void* out, void* agcm_out,
                   int skip_classifier, void* inputs_consumed_event, void* stream);
/* Fused front end of the pipelined path: hdrtv_preprocess + hdrtv_classify of one frame in one call (preprocess:2239-2296 */
/* followed by the input-only part of the network).  On the FP16 path the normalise pass also writes the tensor-core      */
/* staging copy of the image, so the frame is read once and no separate staging launch runs.                              */
int hdrtv_preprocess_classify(hdrtv_t* h, const uint8_t* bgr, int height, int width, void* x_out, void* cond_out,
                              int cond_mode, void* stream);

/* _tensor_to_rgb48_bytes (gui_pipeline_worker_feeders.py:193-249): (1,3,H,W) planar of `dtype` -> uint16 HxWx3 RGB */
/* (rgb48le), FP32 clamp*65535+0.5 truncate.  dst may be device memory or a mapped pinned ring slot.                */
/* transfer = HDRTV_TRANSFER_LUT applies a 15361-entry code table (half bit pattern of the clamped value -> code).  */
int hdrtv_pack_rgb48(hdrtv_t* h, const void* src, int dtype, int height, int width, uint16_t* dst, int transfer,
                     void* stream);
int hdrtv_set_transfer_lut(hdrtv_t* h, const uint16_t* lut_host, int n);

/* The whole hot path of one frame in one call (SURVEY §8b `hdrtv_process`): what the playback / export loops do per  */
/* frame with preprocess -> infer -> _tensor_to_rgb48_bytes (gui_pipeline_worker_frame_processing.py:168-331,         */
/* gui_pipeline_worker_feeders.py:193-249, gui_export.py:1034-1104).  bgr = uint8 HxWx3 BGR and rgb48 = uint16 HxWx3  */
/* RGB may each be device memory or (pinned) host memory - detected with cudaPointerGetAttributes; host frames are    */
/* DMA-copied to / from context-owned device buffers.  Three-stage frame pipeline on context-owned streams: copy-in   */
/* (H2D, normalise, condition image, AGCM classifier) | caller's `stream` (AGCM MLP, LE network, pack) | copy-out     */
/* (D2H), so frame k+1's input side and frame k-1's output copy overlap frame k's network.  Nothing synchronises the  */
/* host: `done_event` (cudaEvent_t, optional) is recorded when rgb48 is complete; the bgr frame must stay untouched   */
/* until then.  Results are bit-identical to hdrtv_preprocess + hdrtv_infer + hdrtv_pack_rgb48.                       */
enum { HDRTV_PROCESS_SERIAL = 1,        /* everything on `stream`, no overlap between frames (lowest single-frame latency) */
       HDRTV_PROCESS_INPUT_READY = 2,   /* device frame already complete: the copy-in stream need not wait for `stream`   */
       HDRTV_PROCESS_RESYNC = 4 };      /* other entry points used the context since the last hdrtv_process on `stream`   */
int hdrtv_process(hdrtv_t* h, const uint8_t* bgr, int height, int width, uint16_t* rgb48, int cond_mode, int transfer,
                  int flags, void* done_event, void* stream);
/* Same, plus the frame's output descriptor checksum (ordered output descriptors of the frame-sharded export): */
/* sum_i code[i] * ((i mod 65521) + 1) over the uint16 HxWx3 frame, computed by the pack kernel on the device and copied */
/* to *checksum_out (pinned host or device memory, may be NULL) before done_event is recorded.                            */
int hdrtv_process_ex(hdrtv_t* h, const uint8_t* bgr, int height, int width, uint16_t* rgb48, int cond_mode, int transfer,
                     int flags, uint64_t* checksum_out, void* done_event, void* stream);
int hdrtv_process_flush(hdrtv_t* h, void* stream);          /* `stream` waits for the copy-out stream's pending copies */
const void* hdrtv_process_output(const hdrtv_t* h, int which); /* device (1,3,H,W) out (0) / agcm_out (1) / fp32 HG out (2) of the last frame */

/* _letterbox_bgr (src/gui_scaling.py:228-244), the resize the frame loop applies before preprocess                          */
/* (gui_pipeline_worker_frame_processing.py): aspect-preserving cv2.resize - INTER_AREA when shrinking, INTER_CUBIC when    */
/* enlarging, size = round(src * min(out_w / w, out_h / h)) - centred on a black canvas.  src = uint8 HxWx3 and             */
/* dst = uint8 out_height x out_width x 3, both device memory.  Bytes equal OpenCV's own resize (resize.cpp) operation by  */
/* operation; with the Intel-IPP dispatch of the pip wheel the cubic case differs from cv2 by at most one code.            */
int hdrtv_letterbox_bgr(hdrtv_t* h, const uint8_t* src, int height, int width, uint8_t* dst, int out_height, int out_width,
                        void* stream);

/* HG stage (third HDRTVNet++ stage).  hdrtv_set_hg_weights replaces                                 */
/* model.hg.load_state_dict(hg_state, strict=True) (hdrtvnet_torch.py:2141-2143): the Hallucination_Generator state-dict */
/* (Hallucination_arch.py:53-98, nf = 64; 92 tensors with BatchNorm, or 42 with the FusedBN fold already applied,         */
/* :201-275); eval-mode BatchNorm is folded into the convs here.  n = 0 removes the stage.                                */
/* hdrtv_hg replaces HG_Composite.forward after the base model (HG_Composite_arch.py:86-107) and                          */
/* Hallucination_Generator.forward (Hallucination_arch.py:101-137): highlight mask from base_out, reflect pad to a        */
/* multiple of 32, the 64..512-channel U-Net (K-streamed tcgen05 implicit GEMMs on HDRTV_FP16 contexts), mask blend,      */
/* crop.  base_out = (1,3,H,W) planar of the context's precision (the `out` of hdrtv_infer); out = (1,3,H,W) planar       */
/* FLOAT32 in both precisions, as in the reference (mask.float() * out + img promotes a half model's output).            */
/* Once HG weights are installed, hdrtv_process / hdrtv_process_ex run the stage between the LE network and the pack.     */
/* Outside the mask the stage leaves the frame unchanged (mask * hg + img, mask = 0) and a masked pixel depends on nothing   */
/* further than 186 pixels away: on HDRTV_FP16 contexts the stage-in pass leaves a map of the cells holding masked pixels in */
/* device memory and the U-Net launches compute only the tiles within that reach of them (none: they return at once) - no host */
/* synchronisation, bit-identical output (environment HDRTV_HG_EARLY_OUT=0 forces the dense evaluation).                    */
int hdrtv_set_hg_weights(hdrtv_t* h, const hdrtv_tensor_desc* tensors, int n);
int hdrtv_hg(hdrtv_t* h, const void* base_out, int height, int width, float* out, void* stream);
/* Per-launch device times (ms) of one hdrtv_hg (FP16 contexts), like hdrtv_time_plan; returns the count.                */
int hdrtv_hg_time_plan(hdrtv_t* h, const void* base_out, int height, int width, float* out, float* ms, int cap, char* names,
                       int names_cap, void* stream);

/* HDRTVNetTorch.postprocess (hdrtvnet_torch.py:2352-2368): planar -> uint8 HxWx3 BGR, arithmetic in `dtype`.       */
int hdrtv_pack_bgr24(hdrtv_t* h, const void* src, int dtype, int height, int width, uint8_t* dst, void* stream);

/* Introspection used by tests, smoke() and bench.py.  (Debug / probe entry points live in hdrtv_b200_test.h and are  */
/* exported only by the test build of the library, libhdrtv_b200_test.so.)                                          */
const char* hdrtv_last_error(const hdrtv_t* h);
long hdrtv_launch_count(const hdrtv_t* h);               /* kernels launched by this context so far                 */
/* Per-launch device times (ms) of one FP16 hdrtv_infer, CUDA events between launches; returns the count.          */
int hdrtv_time_plan(hdrtv_t* h, const void* x, const void* cond, int height, int width, void* out, void* agcm_out,
                    float* ms, int cap, char* names, int names_cap, void* stream);
const char* hdrtv_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HDRTV_B200_H */
