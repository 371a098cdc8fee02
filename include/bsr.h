/*
 * bsr.h — C ABI of libbsr.so: B200 (sm_100a) forward pass of the BlindShadowRemoval generator.
 *
 * The reference has no plugin/FFI layer; the seam this library replaces is the Keras model call
 *     gs, rgb, mask22, dif = self.gen(img, uv, reg, chuck=k, training=False)
 * (/root/reference/train_test_GSC.py:422, 807, 871, 901 -> /root/reference/model.py:228-290) and
 *     self.gen(img, uv, reg, frame=F, share=tf.constant(True), chuck=1, training=False)
 * (/root/reference/train_with_TSM.py:433, 676, 718 -> /root/reference/model_with_TSM.py:261-325),
 * plus the element-wise caller glue next to those calls.  Plain pointers and sizes only; all tensors
 * are NHWC float32 at the boundary exactly like the reference's tensors.  Every function returns 0 on
 * success or a negative BSR_E* code; bsr_last_error() gives the message.  A handle is bound to one
 * device and is not thread-safe (one handle per GPU per process, as the reference drives one model
 * from one Python thread).  forward_* calls are asynchronous on the given CUDA stream and perform no
 * allocation: the workspace, the staging of the host / chunk entry points and every tensor map are
 * sized at bsr_create / bsr_load_weights time (the only exception is a handle created with
 * BSR_DEBUG_KEEP=1, whose debug captures allocate).  Consecutive forwards of one handle share that
 * workspace, so the library orders them itself: every forward waits (on the device, no host sync) for
 * the previous forward of the same handle, whatever streams the two were issued on.
 * training=True has no equivalent here (inference only).
 */
#ifndef BSR_H_
#define BSR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bsr_handle bsr_handle;

enum { BSR_VARIANT_GSC = 0, BSR_VARIANT_TSM = 1 };
/* TC16: 16-bit activations/weights, tcgen05 tensor-core convs + attention, fp32 accumulation (the product path).
 *   The storage type is IEEE binary16 (bsr_act_dtype() == "f16"): kind::f16 UMMAs take it at the bfloat16 rate and
 *   its 11-bit significand is what keeps the outputs within north_star's 1e-2 of the fp32 reference; conversions
 *   saturate at +-65504.  BSR_PRECISION_BF16 is the historical name of the same mode.
 * FP32CHECK: fp32 activations/weights on CUDA cores - the check mode north_star asks for (<=1e-4). */
enum { BSR_PRECISION_TC16 = 0, BSR_PRECISION_BF16 = 0, BSR_PRECISION_FP32CHECK = 1 };
enum {
  BSR_OK = 0, BSR_EINVAL = -1, BSR_ECUDA = -2, BSR_ENOMEM = -3, BSR_ESTATE = -4, BSR_EUNSUPPORTED = -5,
  BSR_EDEVICE = -6   /* a kernel's in-kernel watchdog fired (mbarrier wait timed out): outputs are invalid */
};

#define BSR_IMG 256   /* Config.IMG_SIZE, train_test_GSC.py:31 */
#define BSR_FEAT 32   /* bottleneck resolution after three stride-2 convs, model.py:231-234 */

const char* bsr_version(void);
/* "f16" (default build) or "bf16" (-DBSR_ACT_BF16): storage type of the TC16 path. */
const char* bsr_act_dtype(void);
/* Host utility (no device work): the library's own float -> 16-bit storage conversion (round to nearest even,
 * binary16 saturating at +-65504), exposed so the converter's rounding can be tested against NumPy. */
void bsr_convert_h16(const float* in, unsigned short* out, size_t n);
/* Host utility for the checkpoint converter (no device work): CRC-32C (Castagnoli) of `n` bytes, the checksum
 * TensorFlow stores (masked) per leveldb block of `ckpt-N.index` and per tensor in BundleEntryProto.crc32c - what
 * `checkpoint.restore` verifies (train_test_GSC.py:362-365).  crc = 0 starts a new checksum; pass the previous result
 * to continue one. */
unsigned int bsr_crc32c(unsigned int crc, const void* data, size_t n);

/* Replaces `self.gen = Generator()` (train_test_GSC.py:120).  micro_batch = images resident in the
 * workspace at once (forward loops over larger batches); for TSM it must be >= frame. */
int bsr_create(int variant, int precision, int device, int micro_batch, bsr_handle** out);
int bsr_destroy(bsr_handle* h);
const char* bsr_last_error(const bsr_handle* h);   /* h may be NULL: last error of a failed create */

/* Replaces `checkpoint.restore(...).expect_partial()` (train_test_GSC.py:362-365): `blob` is the
 * output of blindshadowremoval_b200.convert (BN-folded canonical fp32 layers); the library packs it
 * into its device layouts (bf16, K-major, channel-padded).  Host pointer. */
int bsr_load_weights(bsr_handle* h, const void* blob, size_t nbytes);

/* Generator.call of model.py:228-290.  Device pointers, NHWC fp32: img[n,256,256,3], uv[n,256,256,3];
 * outputs gs[n,256,256,1], rgb[n,256,256,3], mask22[n,256,256,3], dif[n,256,256,1]; any output may be
 * NULL (every inference caller discards gs and mask22).  `reg`/`chuck` of the reference signature are
 * ignored by model.py and therefore absent. */
int bsr_forward_gsc(bsr_handle* h, const float* img, const float* uv, int n,
                    float* gs, float* rgb, float* mask22, float* dif, void* cuda_stream);

/* Generator.call of model_with_TSM.py:261-325: n = n_chunks*frame images, reg[n,256,256,6]
 * (= [reg_in(3) | reg_out(3)], only channels 0:2 of each half are used, warp.py:139); temporal sharing
 * acts inside each group of `frame` consecutive images; share=0 selects concat([x,x]) (:227-228). */
int bsr_forward_tsm(bsr_handle* h, const float* img, const float* uv, const float* reg,
                    int n_chunks, int frame, int share,
                    float* gs, float* rgb, float* mask22, float* dif, void* cuda_stream);

/* Same as the two calls above but with HOST buffers: copies inputs H2D, runs, copies the non-NULL
 * outputs D2H and synchronises the stream.  This is what a drop-in replacement of the eager Keras
 * call costs end to end (the reference feeds host NumPy batches, dataset.py:296-302). */
int bsr_forward_gsc_host(bsr_handle* h, const float* img, const float* uv, int n,
                         float* gs, float* rgb, float* mask22, float* dif);
int bsr_forward_tsm_host(bsr_handle* h, const float* img, const float* uv, const float* reg,
                         int n_chunks, int frame, int share,
                         float* gs, float* rgb, float* mask22, float* dif);

/* Compact host I/O: the same forward fed with the bytes the dataset really holds
 * instead of the [F,256,256,16] fp32 chunk of dataset.py:296-302.
 *   img_u8 [n,256,256,3] uint8 RGB, expanded on the device as float(u8)/255 (dataset.py:119,159 `/ 255.`);
 *   uv32   [n,32,32,3]  fp32 = tf.image.resize(uv,[32,32]) - the only form model.py:237 consumes;
 *   reg32  [n,32,32,6]  fp32 = tf.image.resize(reg,[32,32]) - the only form warp.py:137 consumes (TSM).
 * Outputs: the four fp32 tensors as above and/or two compact ones, any of them NULL:
 *   rgb_u8  [n,256,256,3] uint8 = rint(clip(con_rgb,0,1)*255), i.e. clip_by_value (train_test_GSC.py:809)
 *           followed by the *255 + cv2.imwrite saturate-cast the loggers apply (utils.py:221,239,180);
 *   dif_f16 [n,256,256,1] IEEE binary16 of `dif` (round to nearest even).
 * 0.21 MB/image H2D instead of 1.57 (3.15 TSM); 0.33 MB/image D2H instead of 1.05 with the compact outputs. */
int bsr_forward_gsc_host_compact(bsr_handle* h, const unsigned char* img_u8, const float* uv32, int n,
                                 float* gs, float* rgb, float* mask22, float* dif,
                                 unsigned char* rgb_u8, unsigned short* dif_f16);
int bsr_forward_tsm_host_compact(bsr_handle* h, const unsigned char* img_u8, const float* uv32, const float* reg32,
                                 int n_chunks, int frame, int share,
                                 float* gs, float* rgb, float* mask22, float* dif,
                                 unsigned char* rgb_u8, unsigned short* dif_f16);

/* Chunk entry = the whole body of the reference's test steps around the generator call: the dataset hands them ONE
 * tensor [n,256,256,C] (dataset.py:296-302) which they split along channels, feed to the generator and post-process:
 *   BSR_CHUNK_GT    C=16  img3|gt3|uv3|reg6|face1         train_test_GSC.py:415-419, 866-870 (UCB / FFHQ with ground truth)
 *   BSR_CHUNK_SFW   C=17  img3|cmap3|mask1|uv3|reg6|face1  train_test_GSC.py:802-806; train_with_TSM.py:671-675 (SFW labels)
 *   BSR_CHUNK_PLAIN C=13  img3|uv3|reg6|face1              train_test_GSC.py:896-900; train_with_TSM.py:713-717
 * then  _, rgb, _, mask_pred = gen(img, uv, reg, ...);  mask_pred *= face;  rgb = clip(rgb, 0, 1)  (:807-809, 871-873,
 * 901-903).  Device pointers; `chunk` is read in place (one de-interleaving pass into a library-owned staging buffer
 * that is allocated on the first chunk call); outputs rgb_clipped[n,256,256,3], mask_pred[n,256,256,1] and, optionally
 * (may be NULL), the raw gs[n,256,256,1] / mask22[n,256,256,3].  frame / share are used by the TSM variant only. */
enum { BSR_CHUNK_GT = 0, BSR_CHUNK_SFW = 1, BSR_CHUNK_PLAIN = 2 };
int bsr_forward_chunk(bsr_handle* h, const float* chunk, int n, int layout, int frame, int share,
                      float* rgb_clipped, float* mask_pred, float* gs, float* mask22, void* cuda_stream);

/* ShareLayer.call of model_with_TSM.py:204-229 on its own (the temporal sharing module; inside bsr_forward_tsm it runs
 * on the 16-bit activations in place): x[n,32,32,C] fp32, reg[n,256,256,6] -> out[n,32,32,2C] fp32 =
 * warp_out(tile(concat(max_f, mean_f)(warp_in(x)))) per group of `frame` consecutive images, or concat([x, x]) when
 * share == 0.  Device pointers, TSM handles only, n <= micro_batch, C <= 291.  x is converted to the handle's
 * activation type first (exact for FP32CHECK handles). */
int bsr_share_layer(bsr_handle* h, const float* x, const float* reg, int n, int C, int frame, int share, float* out,
                    void* cuda_stream);

/* Caller glue, train_test_GSC.py:808-809 / 872-873 / 902-903 (TSM: train_with_TSM.py:677-678):
 * mask_pred = dif*face ; rgb_clipped = clip(rgb, 0, 1).  Device pointers; in-place allowed. */
int bsr_caller_glue(bsr_handle* h, const float* rgb, const float* dif, const float* face, int n,
                    float* rgb_clipped, float* mask_pred, void* cuda_stream);
/* Composite, train_test_GSC.py:711,718: out = clip(pred*m + inp*(1-m), 0, 1) over n_elems floats. */
int bsr_composite(bsr_handle* h, const float* pred, const float* inp, const float* m, size_t n_elems,
                  float* out, void* cuda_stream);

/* Post-processing of FSRNet.test_step (train_test_GSC.py:436-725), the block that follows the generator call in the
 * UCB evaluation `fsr.test` (BASELINE config 2), for n independent samples (frame 0 of n chunks).  Device pointers:
 *   img, gt, rgb [n,256,256,3], dif [n,256,256,1]  input image, ground truth, generator outputs [1] and [3] of frame 0;
 *   sizes [n] int32                                  box[3] - box[1] of face_crop_and_resize (:417), 1..256;
 *   masks [n,7,256,256] uint8 {0,1}                  region masks in the order face+hair, face, mouth, nose, eyebrow,
 *                                                    eye, glasses (:387-393; single channel: the PNGs are grey).
 * Steps: resize everything to [size,size] and zero-pad to 256 (:438-476), mustache / mouth false positives (:479-496),
 * per-pixel threshold with the hair / forehead / mouth-and-below / left-eyebrow rules (:518-570), 4-connected
 * components keeping those >= 0.45 x the largest and < 80 % hair (:590-611, cv2.connectedComponentsWithStats in the
 * reference), nose rule (:650-662), final = clip(pred*m + input*(1-m), 0, 1) (:711, 718), SSIM / PSNR vs gt (:724-725).
 * Outputs: final_out [n,256,256,3]; detected_out [n,256,256] (the {0,1} shadow mask m; may be NULL);
 * metrics [n,2] = (ssim, psnr) (may be NULL).  Everything runs on the device, no host synchronisation; scratch is grown
 * on the first call with a larger n (this entry is not part of the forward path). */
int bsr_postprocess_ucb(bsr_handle* h, int n, const float* img, const float* gt, const float* rgb, const float* dif,
                        const int* sizes, const unsigned char* masks, float* final_out, float* detected_out,
                        float* metrics, void* cuda_stream);

/* Device-side error check.  Kernels never hang on a protocol error: a 2 s watchdog in every mbarrier wait sets a
 * device flag and the kernel drains.  The flag is copied to pinned host memory at the end of every forward; every
 * forward_* entry point first looks at it (no sync) and returns BSR_EDEVICE if an EARLIER forward tripped it.
 * bsr_check() synchronises the handle's last forward and reports the flag of everything issued so far (and clears
 * it): BSR_OK or BSR_EDEVICE.  The *_host entry points do this themselves before returning. */
int bsr_check(bsr_handle* h);

/* Introspection for tests/bench. */
int bsr_launch_count(const bsr_handle* h);          /* kernels launched by the last forward call */
/* Launch-plan counters of the last forward (tests assert that the benchmarked code paths really ran):
 * which = 0: conv launches with resident weights, 1: with per-CTA pinned weights (qkv), 2: staged TMA-store epilogues,
 * 3: fused attention+w launches, 4: micro-batches replayed from a captured CUDA graph, 5: 3x3 convs run on the halo-tile
 * kernel (res conv2). */
int bsr_plan_counter(const bsr_handle* h, int which);
size_t bsr_workspace_bytes(const bsr_handle* h);
/* Copy a named intermediate of the LAST micro-batch to host as fp32 (dense NHWC, logical channels).
 * Names: x1 x2 x3 x_in0 res0..res5 up1 up2 up3 x_in3 clr_up1 clr_up2 clr_up3 bmask dif_small.
 * Returns the element count in *n_elems (call with host_out=NULL to query). */
int bsr_debug_read(bsr_handle* h, const char* name, float* host_out, size_t capacity, size_t* n_elems);
/* Per-layer device time of the last forward when BSR_PROFILE=1 was set at create time: fills up to
 * `capacity` entries; names are static strings. Returns the number of entries. */
int bsr_layer_times(const bsr_handle* h, const char** names, float* ms, int capacity);

#ifdef __cplusplus
}
#endif
#endif  /* BSR_H_ */
