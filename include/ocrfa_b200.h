/* ocrfa_b200 - C ABI of the B200-native OCR-FA recognition hot path.
 *
 * The reference has no FFI layer: its boundary is a set of Python callables. Each entry
 * point below replaces the arithmetic behind one of them; INTEGRATION.md shows the ctypes binding a reference
 * maintainer would add. Paths are relative to /root/reference/.
 *
 *   ocrfa_create / ocrfa_load_weight / ocrfa_finalize_weights
 *        <- get_recognizer(...)                         src/ocrfa/utils.py:665-695   (Model(...), load_state_dict)
 *   ocrfa_preprocess_u8
 *        <- AlignCollate.__call__ + NormalizePAD        src/ocrfa/utils.py:541-550, 575-602
 *   ocrfa_forward
 *        <- Model.forward(inputs, *args)                src/ocrfa/modules/blocks.py:440-447
 *   ocrfa_ctc_greedy
 *        <- recognizer_predict softmax/mask/renorm/max  src/ocrfa/utils.py:628-638
 *           CTCLabelConverter.decode_greedy             src/ocrfa/utils.py:1099-1115
 *           confidence + custom_mean                    src/ocrfa/utils.py:647-660, 446-447
 *   ocrfa_ctc_beam
 *        <- CTCLabelConverter.decode_beamsearch         src/ocrfa/utils.py:1117-1122 (ctcBeamSearch :949-1038)
 *   ocrfa_recognize_host
 *        <- recognizer_predict(...) for one batch       src/ocrfa/utils.py:605-662  (host buffers in, host results out)
 *
 * Conventions: plain pointers and sizes only; `stream` is a cudaStream_t passed as void*; device pointers are
 * prefixed d_, host pointers h_. Every function returns 0 on success or a negative code; the message is
 * available from ocrfa_last_error() (thread-local). Nothing here falls back to the CPU: without a CUDA device
 * the compute entry points fail with OCRFA_ERR_CUDA.
 * Index -> character mapping stays with the caller (converter.character), so Unicode / bidi handling is untouched.
 */
#ifndef OCRFA_B200_H_
#define OCRFA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCRFA_OK 0
#define OCRFA_ERR_ARG (-1)
#define OCRFA_ERR_CUDA (-2)
#define OCRFA_ERR_WEIGHTS (-3)
#define OCRFA_ERR_WORKSPACE (-4)
#define OCRFA_ERR_STATE (-5)

/* arithmetic of the contraction layers */
#define OCRFA_PRECISION_FP32 0 /* CUDA-core fp32: logits within 1e-3 of the reference's fp32 path */
#define OCRFA_PRECISION_BF16 1 /* tcgen05 bf16 inputs, fp32 accumulate: logits within 2e-2 */
#define OCRFA_PRECISION_F16X3 2 /* the host mirror's DEFAULT. tcgen05 on fp16 hi/lo split operands, 3 products per term,
                                   the dominant one summed per K chunk in fp32 registers and compensated for the tensor
                                   core's truncating accumulate: fp32-grade logits (0.5e-6 of their rms from exact
                                   arithmetic; the reference's own fp32 run: 0.6e-6). Results do not depend on how lines
                                   are batched. Experiment switches (read at ocrfa_create / per launch): OCRFA_X3_CHAIN=1
                                   (one 64-wide K chunk per accumulation instead of two), OCRFA_X3_COMP / OCRFA_LSTM_COMP */
#define OCRFA_PRECISION_F16 3 /* tcgen05 fp16 inputs/activations, fp32 accumulate: the bf16 kernels with 3 more mantissa bits
                                (same rate and bytes); activations saturate at +-65504 */

typedef struct ocrfa_handle ocrfa_handle;

typedef struct ocrfa_config {
  int32_t input_channel;  /* 1 (network_params of "generation1", ocr.py:180-185) */
  int32_t output_channel; /* 512 */
  int32_t hidden_size;    /* 512 */
  int32_t num_class;      /* len(converter.character), 185 for the default Farsi model */
  int32_t precision;      /* OCRFA_PRECISION_* */
  int32_t device;         /* CUDA device ordinal */
} ocrfa_config;

int ocrfa_abi_version(void);
const char* ocrfa_last_error(void);

int ocrfa_create(const ocrfa_config* cfg, ocrfa_handle** out);
void ocrfa_destroy(ocrfa_handle* h);

/* CRAFT text detector (src/ocrfa/utils.py:31-98 class OCRFA over modules/blocks.py:27-82 Vgg16BN; called at
 * detectors.py:43-47): a second kind of handle. Load the detector's state dict with ocrfa_load_weight, call
 * ocrfa_finalize_weights, then ocrfa_craft_forward = OCRFA.forward: d_x float32 [N,3,H,W] (mean/variance normalised,
 * H and W multiples of 32 as resize_aspect_ratio produces) -> d_y float32 [N,H/2,W/2,2] (text, link score maps, the
 * reference's y.permute(0,2,3,1)) and, when not NULL, d_feature float32 [N,32,H/2,W/2]. precision: BF16, F16 (single-pass
 * tensor-core modes) or F16X3 (fp32-grade: score maps within 1.4e-6 of their range of the reference's fp32 output, so the
 * thresholded components and hence the boxes are the reference's). */
int ocrfa_craft_create(int32_t precision, int32_t device, ocrfa_handle** out);
size_t ocrfa_craft_workspace_bytes(const ocrfa_handle* h, int32_t N, int32_t H, int32_t W);
int ocrfa_craft_forward(ocrfa_handle* h, const float* d_x, int32_t N, int32_t H, int32_t W, float* d_y, float* d_feature,
                        void* d_ws, size_t ws_bytes, void* stream);

/* Detector preprocessing (detectors.py:33-43): resize_aspect_ratio (src/ocrfa/utils.py:497-524: cv2.resize with
 * INTER_LINEAR to dst_w x dst_h, pasted into a zero canvas of H32 x W32) + normalize_mean_variance (utils.py:479-484:
 * (v - mean*255) / (variance*255) in float32, applied to the canvas INCLUDING its zero padding) + HWC -> CHW, in one
 * kernel: d_img uint8 [src_h][src_w][3] (src_stride bytes per row) -> d_out float32 [3][H32][W32], bit-exact with
 * the reference's x. d_xtab / d_ytab are the per-axis tables of ocrfa_crop_resize_u8 (int32 [dst][2]); mode as in
 * ocrfa_crop_desc. mean255 / var255: 3 host floats each (float32(mean*255), float32(variance*255)). */
int ocrfa_det_preprocess_u8(const uint8_t* d_img, int32_t src_h, int32_t src_w, int32_t src_stride, int32_t dst_h,
                            int32_t dst_w, int32_t H32, int32_t W32, const int32_t* d_xtab, const int32_t* d_ytab,
                            int32_t mode, const float* mean255, const float* var255, float* d_out, void* stream);

/* reformat_input's grey page for a colour ndarray (src/ocrfa/utils.py:1545-1547): cv2.cvtColor(image, COLOR_BGR2GRAY)
 * on uint8 [H][W][3] (img_stride bytes per row) -> uint8 [H][W] (gray_stride bytes per row), bit-exact with OpenCV's
 * 15-bit fixed point (first channel x 3735 + second x 19235 + third x 9798, + 16384, >> 15). */
int ocrfa_bgr2gray_u8(const uint8_t* d_img, int32_t img_stride, int32_t H, int32_t W, uint8_t* d_gray,
                      int32_t gray_stride, void* stream);

/* The pixel work of get_det_boxes_core (src/ocrfa/utils.py:106-153): thresholds, 4-connected labeling with
 * cv2.connectedComponentsWithStats' label numbers and stats, the per-label maximum of the text map, the size (< 10)
 * and score (< text_threshold) filters, niter and the dilation ROI, and the rectangle dilation of each surviving
 * component (minus its link-only pixels) expressed as ROW EXTENTS: for ROI row sy + r the dilated mask's smallest and
 * largest x (min > max = empty row). cv2.minAreaRect depends on a pixel set only through its convex hull, hence only
 * through these extents, so the host finishes each box from (ey - sy) pairs (ocrfa_b200/detection.py).
 * d_text / d_link: float32 maps, element (y, x) at [(y*W + x) * pix_stride] (2 for the CRAFT output's interleaved
 * [H,W,2] layout with d_link = d_text + 1; 1 for planar maps). Thresholds are float32 as cv2.threshold / NumPy compare
 * them. d_labels: int32 [H][W] out (cv2's labels; may be NULL). d_comps: the surviving components in label order.
 * d_extents: int32 [max_rows][2]. d_counts int32 [4]: nLabels (cv2's, background included), components written,
 * extent rows written, overflow flag (1 when max_comps / max_rows were too small: results are then incomplete). */
typedef struct ocrfa_det_comp {
  int32_t label;                    /* k */
  int32_t area;                     /* stats[k, CC_STAT_AREA] */
  int32_t left, top, width, height; /* stats[k, CC_STAT_LEFT / TOP / WIDTH / HEIGHT] */
  int32_t niter;
  int32_t sx, sy, ex, ey;           /* dilation ROI */
  int32_t row_off;                  /* first row of this component in d_extents; it owns ey - sy rows */
  float max_text;                   /* np.max(textmap[labels == k]) */
} ocrfa_det_comp;
size_t ocrfa_det_workspace_bytes(int32_t H, int32_t W, int32_t max_rows);
int ocrfa_det_components(const float* d_text, const float* d_link, int32_t pix_stride, int32_t H, int32_t W,
                         float low_text, float link_threshold, float text_threshold, int32_t* d_labels,
                         ocrfa_det_comp* d_comps, int32_t max_comps, int32_t* d_extents, int32_t max_rows,
                         int32_t* d_counts, void* d_ws, size_t ws_bytes, void* stream);

/* One state-dict entry (float32, host memory, PyTorch layout); an optional "module." prefix is stripped.
 * Integer entries (num_batches_tracked) may be skipped by the caller. */
int ocrfa_load_weight(ocrfa_handle* h, const char* name, const float* h_data, const int64_t* shape, int32_t ndim);
/* Folds BatchNorm into the convolutions, repacks for the kernels and uploads. Call once after all weights. */
int ocrfa_finalize_weights(ocrfa_handle* h, void* stream);

/* T = W/4 + 1 timesteps for an input width W (multiple of 4); 0 on invalid geometry. */
int32_t ocrfa_out_timesteps(int32_t img_h, int32_t img_w);
/* Bytes of caller-provided device scratch ocrfa_forward needs for a batch of B lines of H x W. */
size_t ocrfa_workspace_bytes(const ocrfa_handle* h, int32_t B, int32_t H, int32_t W);

/* uint8 lines already at model height -> float32 [B,1,H,W] in [-1,1]; columns >= widths[b] replicate the last
 * valid column (widths may be NULL: all lines are W wide). in_stride = bytes between rows of d_img. */
int ocrfa_preprocess_u8(const uint8_t* d_img, const int32_t* d_widths, int32_t B, int32_t H, int32_t W,
                        int32_t in_stride, float* d_x, void* stream);

/* AlignCollate for lines that are not yet at model height (src/ocrfa/utils.py:587-598): optional contrast stretch
 * (adjust_contrast_grey, utils.py:450-463, applied when contrast_target > 0 and desc.contrast != 0) followed by
 * Pillow's 8-bit fixed-point BICUBIC resize to img_h rows, bit-exact with PIL. Lines are packed back to back in
 * d_src; the fixed-point coefficient tables are produced on the host (ocrfa_b200/resize.py mirrors Pillow's
 * precompute_coeffs). d_tmp holds the horizontal-pass intermediates, d_lut is [B][256] scratch,
 * d_out is uint8 [B, img_h, img_w] (columns >= out_w zeroed; ocrfa_preprocess_u8 / the stem replicate-pad them). */
typedef struct ocrfa_line_desc {
  int64_t src_off;  /* byte offset of uint8 [src_h][src_w] in d_src */
  int64_t tmp_off;  /* byte offset of the [src_h][out_w] intermediate in d_tmp */
  int32_t src_h, src_w, out_w;
  int32_t kh_off, kh_size, bh_off; /* horizontal: coef int32 [out_w][kh_size], bounds int32 [out_w][2]; kh_size 0 = copy */
  int32_t kv_off, kv_size, bv_off; /* vertical:   coef [img_h][kv_size], bounds [img_h][2];            kv_size 0 = copy */
  int32_t contrast;
} ocrfa_line_desc;
int ocrfa_align_collate_u8(const uint8_t* d_src, const ocrfa_line_desc* d_desc, const int32_t* d_tables, int32_t B,
                           int32_t img_h, int32_t img_w, double contrast_target, uint8_t* d_tmp, uint8_t* d_lut,
                           uint8_t* d_out, void* stream);

/* get_image_list / compute_ratio_and_resize for axis-aligned boxes (src/ocrfa/utils.py:1361-1418): crops n boxes out
 * of one uint8 page image and resizes each with OpenCV's 8-bit INTER_LINEAR arithmetic (what the reference's
 * cv2.resize(..., interpolation=Image.Resampling.LANCZOS) executes: the enum's value is 1), bit-exact with cv2.
 * Sources are boxes of a page image or the warped boxes ocrfa_warp_perspective_u8 produced (each desc carries its own
 * offset and row stride). The per-axis index/weight tables are produced on the host (ocrfa_b200/crops.py). Outputs
 * are packed back to back in d_out: uint8 [dst_h][dst_w] at out_off - the layout ocrfa_align_collate_u8 takes as d_src. */
typedef struct ocrfa_crop_desc {
  int64_t out_off;            /* byte offset of the resized crop in d_out */
  int64_t src_off;            /* byte offset of the box's top-left pixel in d_src */
  int32_t src_stride;         /* bytes between rows of the source */
  int32_t src_w, src_h;       /* box size (already clipped to the source) */
  int32_t dst_w, dst_h;       /* resized size */
  int32_t xtab_off, ytab_off; /* int32 offsets into d_tables: per output column {x0, a0 | a1 << 16}, per row {y0, b0 | b1 << 16} */
  int32_t mode;               /* 0 = bilinear, 1 = copy (equal sizes), 2 = exact 2x shrink (OpenCV switches to INTER_AREA) */
} ocrfa_crop_desc;
int ocrfa_crop_resize_u8(const uint8_t* d_src, const ocrfa_crop_desc* d_desc, const int32_t* d_tables, int32_t n,
                         uint8_t* d_out, void* stream);

/* four_point_transform for rotated boxes (src/ocrfa/utils.py:1218-1231): cv2.warpPerspective(page, M, (dst_w, dst_h))
 * with INTER_LINEAR / BORDER_CONSTANT 0, bit-exact with OpenCV's fixed-point remap. minv is the INVERSE of the
 * reference's getPerspectiveTransform matrix (float64, row-major; ocrfa_b200/crops.py restates OpenCV's LU solve
 * and 3x3 inverse); block_w is the column-block width OpenCV evaluates row starts at. Warped boxes are packed back
 * to back in d_out (uint8 [dst_h][dst_w] at out_off), ready for ocrfa_crop_resize_u8. */
typedef struct ocrfa_warp_desc {
  int64_t out_off;
  int32_t dst_w, dst_h;
  int32_t block_w;
  int32_t reserved;
  double m[9];
} ocrfa_warp_desc;
int ocrfa_warp_perspective_u8(const uint8_t* d_page, int32_t page_h, int32_t page_w, int32_t page_stride,
                              const ocrfa_warp_desc* d_desc, int32_t n, uint8_t* d_out, void* stream);

/* make_rotated_img_list for right angles (src/ocrfa/utils.py:1599-1612): scipy.ndimage.rotate(img, 90*k, reshape=True)
 * of a uint8 image = np.rot90(img, k), k in {1, 2, 3}: d_src uint8 [h][w] contiguous -> d_dst uint8 [w][h] (k odd) or
 * [h][w] (k = 2). */
int ocrfa_rot90_u8(const uint8_t* d_src, int32_t h, int32_t w, int32_t k, uint8_t* d_dst, void* stream);

/* d_x: float32 [B,1,H,W]; d_logits: float32 [B,T,num_class]. */
int ocrfa_forward(ocrfa_handle* h, const float* d_x, int32_t B, int32_t H, int32_t W, float* d_logits, void* d_ws,
                  size_t ws_bytes, void* stream);
/* The two halves of ocrfa_forward, for stage-wise use and parity checks:
 *   features = AdaptiveAvgPool(FeatureExtraction(x))   float32 [B,T,512]   (blocks.py:432-433, 443-446)
 *   logits   = Prediction(SequenceModeling(features))  float32 [B,T,num_class] (blocks.py:434-438, 441-442)
 * (H, W) always name the INPUT line geometry; T = W/4 + 1. */
int ocrfa_forward_features(ocrfa_handle* h, const float* d_x, int32_t B, int32_t H, int32_t W, float* d_features,
                           void* d_ws, size_t ws_bytes, void* stream);
int ocrfa_forward_sequence(ocrfa_handle* h, const float* d_features, int32_t B, int32_t H, int32_t W, float* d_logits,
                           void* d_ws, size_t ws_bytes, void* stream);

/* Same, straight from uint8 lines (normalisation fused into the stem convolution). */
int ocrfa_forward_u8(ocrfa_handle* h, const uint8_t* d_img, const int32_t* d_widths, int32_t B, int32_t H,
                     int32_t W, float* d_logits, void* d_ws, size_t ws_bytes, void* stream);

/* softmax over C, zero the classes with d_ignore_mask[c] != 0, renormalise, argmax, CTC collapse, confidence.
 * Optional outputs may be NULL: d_probs [B,T,C], d_idx [B,T], d_maxp [B,T]. d_collapse_ignore [C] may be NULL
 * (only blank is dropped). d_labels [B,T] int32 (first d_len[b] valid), d_conf [B] float64. */
int ocrfa_ctc_greedy(const float* d_logits, int32_t B, int32_t T, int32_t C, const uint8_t* d_ignore_mask,
                     const uint8_t* d_collapse_ignore, float* d_probs, int32_t* d_idx, float* d_maxp,
                     int32_t* d_labels, int32_t* d_len, double* d_conf, void* stream);

/* CTCLabelConverter.decode_greedy on caller-provided indices (src/ocrfa/utils.py:1099-1115): line b owns
 * d_idx[d_offsets[b] .. d_offsets[b+1]); keeps an index iff it differs from its predecessor in the line and is
 * neither blank (0) nor flagged in d_collapse_ignore [C] (may be NULL). d_labels uses the same offsets. */
int ocrfa_ctc_collapse(const int32_t* d_idx, const int32_t* d_offsets, int32_t B, int32_t C,
                       const uint8_t* d_collapse_ignore, int32_t* d_labels, int32_t* d_len, void* stream);

/* np.argmax(mat, axis=-1) of a [rows, C] float32 matrix (first maximum wins) - the word segmentation input of
 * decode_wordbeamsearch (src/ocrfa/utils.py:1126). */
int ocrfa_argmax_rows(const float* d_mat, int64_t rows, int32_t C, int32_t* d_idx, void* stream);

/* Scratch bytes for ocrfa_ctc_beam. */
size_t ocrfa_beam_workspace_bytes(int32_t B, int32_t T, int32_t C, int32_t beam_width);
/* CTC prefix beam search over probabilities `mat` [B,T,C] (already masked/renormalised), lm = None, empty
 * dictionary. Bit-exact with the reference's float32 arithmetic. beam_width in [1, 64]. */
int ocrfa_ctc_beam(const float* d_probs, int32_t B, int32_t T, int32_t C, int32_t beam_width,
                   const uint8_t* d_collapse_ignore, int32_t* d_labels, int32_t* d_len, void* d_ws, size_t ws_bytes,
                   void* stream);

/* Beam search over ragged row spans of one [rows, C] matrix (the per-word calls of decode_wordbeamsearch,
 * src/ocrfa/utils.py:1138-1147): span i covers rows d_span_row0[i] .. + d_span_len[i] (<= t_max).
 * d_labels is [n_spans, t_max]; scratch = ocrfa_beam_workspace_bytes(n_spans, t_max, C, beam_width). */
int ocrfa_ctc_beam_spans(const float* d_probs, const int32_t* d_span_row0, const int32_t* d_span_len, int32_t n_spans,
                         int32_t t_max, int32_t C, int32_t beam_width, const uint8_t* d_collapse_ignore,
                         int32_t* d_labels, int32_t* d_len, void* d_ws, size_t ws_bytes, void* stream);

/* The ranked candidate list of the final beam state, for the dictionary lookup of word beam search
 * (BeamState.wordsearch, src/ocrfa/utils.py:853-872, reached from ctcBeamSearch(..., dict_list != []), :1036-1037, which
 * decode_wordbeamsearch calls per word, :1138-1147 and :1158-1172): the same search as ocrfa_ctc_beam / _spans
 * (`beam_width` labelings survive every timestep), but after the last timestep ALL labelings of the state are ranked
 * (prTotal descending, creation order on ties = the reference's stable sort) and the first n_cand (the reference uses
 * maxCandidate = 20) are returned, each collapsed like the best labeling is (ignored indices and adjacent duplicates of
 * the raw labeling dropped). d_span_row0 / d_span_len may both be NULL: item i then covers rows [i*t_max, (i+1)*t_max).
 * d_cand_labels [n_items, n_cand, t_max] int32, d_cand_len [n_items, n_cand], d_cand_count [n_items] (how many
 * candidates exist, <= n_cand; an item with no timestep has one empty candidate). beam_width and n_cand in [1, 64].
 * Scratch = ocrfa_beam_candidates_workspace_bytes. The dictionary membership test stays with the caller (strings). */
size_t ocrfa_beam_candidates_workspace_bytes(int32_t n_items, int32_t t_max, int32_t C, int32_t beam_width, int32_t n_cand);
int ocrfa_ctc_beam_candidates(const float* d_probs, const int32_t* d_span_row0, const int32_t* d_span_len, int32_t n_items,
                              int32_t t_max, int32_t C, int32_t beam_width, int32_t n_cand,
                              const uint8_t* d_collapse_ignore, int32_t* d_cand_labels, int32_t* d_cand_len,
                              int32_t* d_cand_count, void* d_ws, size_t ws_bytes, void* stream);

/* Whole recognizer_predict step for one batch with HOST buffers: H2D of the uint8 lines, preprocess, forward,
 * decode (decoder 0 = greedy, 1 = beamsearch), D2H of labels/lengths/confidences, stream synchronised on return.
 * h_ignore_mask [num_class] may be NULL. h_labels [B,T] int32, h_len [B] int32, h_conf [B] float64. */
int ocrfa_recognize_host(ocrfa_handle* h, const uint8_t* h_img, const int32_t* h_widths, int32_t B, int32_t H,
                         int32_t W, const uint8_t* h_ignore_mask, int32_t decoder, int32_t beam_width,
                         int32_t* h_labels, int32_t* h_len, double* h_conf, void* d_ws, size_t ws_bytes,
                         void* stream);
/* The same step on DEVICE buffers, asynchronous on `stream`: d_img uint8 [B,H,W], outputs d_labels [B,T],
 * d_len [B], d_conf [B] stay on the device. Workspace = ocrfa_recognize_workspace_bytes. */
int ocrfa_recognize_device(ocrfa_handle* h, const uint8_t* d_img, const int32_t* d_widths, int32_t B, int32_t H,
                           int32_t W, const uint8_t* d_ignore_mask, int32_t decoder, int32_t beam_width,
                           int32_t* d_labels, int32_t* d_len, double* d_conf, void* d_ws, size_t ws_bytes,
                           void* stream);

/* Extended forms. d_margin / h_margin [B] float32 (may be NULL) receive the DECISION MARGIN of every line's greedy
 * decode (see ocrfa_ctc_margin). flags: OCRFA_FLAG_ASYNC = return without synchronising `stream` - the D2H copies are
 * still in flight; the caller waits on the stream (or an event) before reading the host buffers, which should be pinned,
 * and leaves them and the workspace untouched until then. Two calls on two streams with two workspaces overlap the
 * H2D / D2H copies of one batch with the kernels of the other (what DataLoader(pin_memory=True) + .to(device) is for in
 * the reference, utils.py:728-730, :621). */
#define OCRFA_FLAG_ASYNC 1u
int ocrfa_recognize_host_ex(ocrfa_handle* h, const uint8_t* h_img, const int32_t* h_widths, int32_t B, int32_t H,
                            int32_t W, const uint8_t* h_ignore_mask, int32_t decoder, int32_t beam_width,
                            int32_t* h_labels, int32_t* h_len, double* h_conf, float* h_margin, uint32_t flags,
                            void* d_ws, size_t ws_bytes, void* stream);
int ocrfa_recognize_device_ex(ocrfa_handle* h, const uint8_t* d_img, const int32_t* d_widths, int32_t B, int32_t H,
                              int32_t W, const uint8_t* d_ignore_mask, int32_t decoder, int32_t beam_width,
                              int32_t* d_labels, int32_t* d_len, double* d_conf, float* d_margin, void* d_ws,
                              size_t ws_bytes, void* stream);

/* Decision margin of the greedy decode (recognizer_predict's preds_prob.max(2), src/ocrfa/utils.py:628-638):
 *   d_margin[b] = min over timesteps of (largest - second largest logit among the classes with d_ignore_mask[c] == 0)
 *                 / rms of those logits in the row.
 * softmax and the renormalisation are monotone, so this is the argmax the reference takes on probabilities. A line
 * whose margin exceeds twice the relative arithmetic error of the logits decodes to the same greedy string under any
 * evaluation within that error, the reference's own fp32 run included; smaller margins are near-ties that the
 * reference itself resolves differently when its batch size changes (profiles/r02_ref_self_consistency.json). */
int ocrfa_ctc_margin(const float* d_logits, int32_t B, int32_t T, int32_t C, const uint8_t* d_ignore_mask,
                     float* d_margin, void* stream);

/* Scratch bytes ocrfa_recognize_host needs (forward workspace + staging + decode scratch). */
size_t ocrfa_recognize_workspace_bytes(const ocrfa_handle* h, int32_t B, int32_t H, int32_t W, int32_t beam_width);

/* CUDA-event profile per kernel class (bit c of class_mask selects class c: 0 stem, 1 tcgen05 implicit GEMM,
 * 2 tcgen05 LSTM step, 3 pool/mean/convert, 4 CUDA-core conv, 5 fp32 LSTM cell, 6 greedy decode, 7 beam search).
 * ocrfa_profile_end synchronises the device and returns, per class, summed kernel milliseconds, summed
 * algorithmic FLOPs and launch counts (arrays of 8) since ocrfa_profile_begin. */
int ocrfa_profile_begin(ocrfa_handle* h, uint32_t class_mask);
int ocrfa_profile_end(ocrfa_handle* h, double* ms_by_class, double* flops_by_class, int64_t* launches_by_class);

/* Number of kernels this library has launched on behalf of the handle since creation (bench bookkeeping). */
int64_t ocrfa_launch_count(const ocrfa_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* OCRFA_B200_H_ */
