/*
 * respmon_b200 -- C ABI of the B200-native hot path of kevroy314/respmon.
 *
 * The reference has no FFI: its boundary is the Python call graph of base.py / transforms.py / pyramid.py
 * (SURVEY.md section 8b).  Each entry point below replaces the arithmetic behind one of those calls; the Python
 * drop-in (respmon_b200/monitor.py, respmon_b200/pyramid.py, respmon_b200/transforms.py) binds them with ctypes.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - every function returns int32: RM_OK (0) or a negative rm_status; rm_last_error(h) has the text.
 *     Nothing throws, nothing calls exit().
 *   - data pointers are CALLER-OWNED DEVICE pointers unless a parameter is named host_*.
 *   - every launch is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream).
 *   - a handle is bound to one device and is not thread-safe; distinct handles are independent.
 *   - per-clip failures are data, not errors: see rm_clip_status.
 *   - images are row-major, clips are (T, H, W), batches are (n_clips, T, H, W), all contiguous.
 */
#ifndef RESPMON_B200_H
#define RESPMON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RM_VERSION 100

typedef enum rm_status {
  RM_OK = 0,
  RM_ERR_INVALID = -1,     /* bad argument */
  RM_ERR_CUDA = -2,        /* CUDA runtime error (text in rm_last_error) */
  RM_ERR_UNSUPPORTED = -3, /* shape / parameter outside what the kernels implement */
  RM_ERR_WORKSPACE = -4    /* workspace too small */
} rm_status;

typedef enum rm_dtype { RM_U8 = 0, RM_F32 = 1, RM_F64 = 2,
                        RM_BGR8 = 3 /* 8-bit BGR frames (H, W, 3) as cv2.VideoCapture.read() yields them, base.py:228-230 */ } rm_dtype;

/* per-clip outcome of the batch path; replaces the reference's state flips (base.py:249-253, 451-454, 543-545) */
typedef enum rm_clip_status {
  RM_CLIP_OK = 0,
  RM_CLIP_NO_ROI = 1,      /* locate() returned None (base.py:569-570) */
  RM_CLIP_NO_CORNERS = 2,  /* goodFeaturesToTrack found nothing (base.py:367-368) */
  RM_CLIP_TRACK_LOST = 3,  /* extract_motion returned nan (base.py:373-374, 385-386) */
  RM_CLIP_NO_PEAKS = 4     /* fewer than two accepted peaks: no BPM (base.py:349) */
} rm_clip_status;

/* Hyper-parameters the reference hard-codes (base.py:80-106; locate defaults base.py:548-552). */
typedef struct rm_params {
  int32_t pyramid_levels;      /* 9    base.py:550 */
  int32_t skip_levels_at_top;  /* 4    base.py:550 */
  double freq_min;             /* 0.1  base.py:82 */
  double freq_max;             /* 1.0  base.py:83 */
  double amplification;        /* 500  base.py:549 */
  double temporal_threshold;   /* 0.7  base.py:84 */
  int32_t threshold;           /* 20 = int(round(0.08*255))  base.py:85, 448 */
  int32_t max_corners;         /* 100  base.py:91 */
  double quality_level;        /* 0.3  base.py:92 */
  int32_t min_distance;        /* 7    base.py:93 */
  int32_t block_size;          /* 7    base.py:94 */
  int32_t lk_win;              /* 15   base.py:96 */
  int32_t lk_max_level;        /* 2    base.py:97 */
  int32_t lk_max_iter;         /* 10   base.py:98 */
  double lk_eps;               /* 0.03 base.py:98 */
  double lk_min_eig;           /* 1e-4 (OpenCV default minEigThreshold) */
  double gaussian_cutoff;      /* 10.0 base.py:100 */
  int32_t filter_order;        /* 3    base.py:101 */
  int32_t measure_buffer_len;  /* 128  base.py:88 */
  int32_t measure_init_len;    /* 12   base.py:106 */
  double peak_threshold;       /* 0.3  peakutils.indexes default `thres` (base.py:314) */
} rm_params;

/* 32-byte per-clip result record, the unit of the final all-gather. */
typedef struct rm_result {
  double bpm;        /* freq[-1] (base.py:352) or NaN */
  int32_t x, y, w, h;/* ROI (base.py:456) */
  int32_t status;    /* rm_clip_status */
  int32_t n_peaks;   /* len(peak_indices) of the last window (base.py:345) */
} rm_result;

typedef struct rm_handle rm_handle;

/* ------------------------------------------------------------------ lifetime / host-side helpers (no GPU work) */
int32_t rm_version(void);
int32_t rm_default_params(rm_params* out);
int32_t rm_create(const rm_params* params, int32_t device, rm_handle** out);
int32_t rm_destroy(rm_handle* h);
const char* rm_last_error(rm_handle* h);
/* sizes (w,h) of `n_levels` pyramid levels: the ((n+1)//2) chain of cv2.pyrDown (pyramid.py:13-15). wh_out[2*n_levels]. */
int32_t rm_level_sizes(int32_t W, int32_t H, int32_t n_levels, int32_t* wh_out);
/* bound_low / bound_high of transforms.py:88-90 (argmin over scipy.fftpack.fftfreq). */
int32_t rm_temporal_bounds(int32_t T, double fps, double freq_min, double freq_max, int32_t* lo, int32_t* hi);
/* scipy.signal.butter(order, wn, 'low') coefficients (transforms.py:58-63); b[order+1], a[order+1]. */
int32_t rm_butter_lowpass(int32_t order, double wn, double* host_b, double* host_a);
/* the 256-entry LUT of uint8 -> float -> uint8 (transforms.py:20-29 round trip). */
int32_t rm_lossy_u8_lut(uint8_t* host_lut256);

/* ------------------------------------------------------------------ test data */
typedef struct rm_clip_spec {
  int32_t width, height, n_frames, seed;
  int32_t x0, y0, w0, h0; /* moving patch rectangle */
} rm_clip_spec;
/* Generate n clips on the device; dq8 is (n, T) int32 displacement tables (host-computed, device-resident);
 * specs is a device array; out is (n, T, H, W) uint8.  All clips share W,H,T.  Bit-identical to respmon_b200/synth.py. */
int32_t rm_synth_clips(rm_handle* h, const rm_clip_spec* specs, const int32_t* dq8, int32_t n_clips, uint8_t* out,
                       void* stream);

/* ------------------------------------------------------------------ frame ingest */
/* cv2.cvtColor(frame, COLOR_BGR2GRAY) of next_frame (base.py:230) on interleaved 8-bit BGR pixels. */
int32_t rm_bgr_to_gray(rm_handle* h, const uint8_t* bgr, uint8_t* gray_out, int64_t n_pixels, void* stream);

/* `cropped_image = current_frame[y:y+h, x:x+w]` (base.py:471) for a run of frames of every clip: frames (n_clips,T,H,W) of
 * RM_U8, or (n_clips,T,H,W,3) of RM_BGR8 -- then cv2.cvtColor(BGR2GRAY) of next_frame (base.py:230) is applied to the ROI's
 * pixels only --, roi (n_clips,4) x,y,w,h -> out (n_clips, n_frames, out_h, out_w) uint8, crops top-left aligned.  The
 * measure entry points take `out` as their frames with ROI origin (0,0).  `frames` (here and in rm_crop_frames_ragged) may
 * be PINNED HOST memory (cudaHostAlloc / cudaHostRegister: device-addressable under unified addressing): the crop is then
 * the upload of the measure frames -- only the ROI's rows cross PCIe, as aligned words of one request per row, the ROI
 * never visits the host and no staging copy is made. */
int32_t rm_crop_frames(rm_handle* h, const void* frames, int32_t dtype, int32_t n_clips, int32_t T, int32_t W, int32_t H,
                       const int32_t* roi, int32_t first_frame, int32_t n_frames, uint8_t* out, int32_t out_w,
                       int32_t out_h, void* stream);

/* Ragged batches (BASELINE config 5): one descriptor per clip.  `frame_offset` = byte offset of the clip's first
 * frame from `base` (base may be NULL: then it is the absolute device address), `row_stride` = bytes between rows
 * (W for packed gray frames, 3 W for packed BGR).  Clips of any mix of resolutions go through ONE launch; the descriptor
 * table lives in device memory. */
typedef struct rm_clip_desc {
  int64_t frame_offset;
  int32_t W, H, T, row_stride;
} rm_clip_desc;

/* rm_crop_frames over a ragged batch: out (n_clips, n_frames, out_h, out_w) gets frame[y:y+h, x:x+w] (base.py:471) of
 * frames [first_frame, first_frame+n_frames) of every clip, whatever its resolution; roi (n_clips,4) in each clip's own
 * frame coordinates.  With the crops in one tensor the measure stage (rm_measure_signal, ROI origin 0,0) runs once over
 * all resolution classes. */
int32_t rm_crop_frames_ragged(rm_handle* h, const void* base, int32_t dtype, const rm_clip_desc* descs, int32_t n_clips,
                              const int32_t* roi, int32_t first_frame, int32_t n_frames, uint8_t* out, int32_t out_w,
                              int32_t out_h, void* stream);

/* ------------------------------------------------------------------ single-level ops (API parity: pyramid.py) */
/* uint8_to_float (transforms.py:20-23): u8 -> f64 * (1/255); f32 -> f64 widening. */
int32_t rm_to_f64(rm_handle* h, const void* src, int32_t dtype, double* dst, int64_t n, void* stream);
/* float_to_uint8 (transforms.py:26-29): img * 255 truncated into uint8 (values in [0,1]). */
int32_t rm_f64_to_u8(rm_handle* h, const double* src, uint8_t* dst, int64_t n, void* stream);
/* cv2.pyrDown on float64 images (pyramid.py:14): (n_img, sh, sw) -> (n_img, (sh+1)/2, (sw+1)/2). */
int32_t rm_pyr_down_f64(rm_handle* h, const double* src, double* dst, int64_t n_img, int32_t sw, int32_t sh, void* stream);
/* cv2.pyrUp(src, dstsize=(dw,dh)) on float64 (pyramid.py:25, pyramid.py:55), fused with the caller's add/sub:
 * mode 0: dst = up(src); mode 1: dst = other - up(src) (Laplacian level); mode 2: dst = up(src) + other (collapse). */
int32_t rm_pyr_up_f64(rm_handle* h, const double* src, double* dst, const double* other, int32_t mode, int64_t n_img,
                      int32_t sw, int32_t sh, int32_t dw, int32_t dh, void* stream);

/* ------------------------------------------------------------------ calibrate: fused hot path */
/* Number of doubles per frame in the packed Laplacian record: sum over levels skip..levels-2 of w_l*h_l
 * (1600 at 640x480 with levels=9, skip=4), level `skip` first. */
int32_t rm_lap_record_len(rm_handle* h, int32_t W, int32_t H, int64_t* out);
/* Workspace (bytes) rm_pyramid_build / rm_heatmap need for a batch of n_frames / (n_clips, T).  The heat-map figure
 * depends on the "no_minmax_seed" option (rm_set_option): query it after setting the option. */
int32_t rm_pyramid_workspace_bytes(rm_handle* h, int32_t W, int32_t H, int64_t n_frames, size_t* out);
int32_t rm_heatmap_workspace_bytes(rm_handle* h, int32_t W, int32_t H, int32_t n_clips, int32_t T, size_t* out);

/* create_laplacian_video_pyramid restricted to the levels that are ever read (pyramid.py:31-48 via
 * transforms.py:148,156-170): frames (n_frames,H,W) of dtype -> lap (n_frames, record_len) float64.
 * u8 frames mean gray/255 (transforms.py:20-23). */
int32_t rm_pyramid_build(rm_handle* h, const void* frames, int32_t dtype, int64_t n_frames, int32_t W, int32_t H,
                         double* lap_out, void* workspace, size_t workspace_bytes, void* stream);

/* The same on a window of every clip of a batch, without copying it: frames (n_clips,T,H,W); clip c contributes frames
 * [first_frame, first_frame+n_frames) (the calibration buffer of base.py:429-434) -> lap (n_clips, n_frames, record_len).
 * Workspace as for n_clips*n_frames frames. */
int32_t rm_pyramid_build_clips(rm_handle* h, const void* frames, int32_t dtype, int32_t n_clips, int32_t T,
                               int32_t first_frame, int32_t n_frames, int32_t W, int32_t H, double* lap_out,
                               void* workspace, size_t workspace_bytes, void* stream);

/* temporal_bandpass_filter_fft (transforms.py:82-102) on every column of (n_clips, T, record_len), in place allowed. */
int32_t rm_temporal_bandpass(rm_handle* h, const double* lap, double* bp_out, int32_t n_clips, int32_t T,
                             int64_t record_len, double fps, void* stream);

/* collapse (pyramid.py:51-69) + global min/max clip (transforms.py:184-192) + time average, normalise, truncate
 * (base.py:562-564): bp (n_clips,T,record_len) -> heat (n_clips,H,W) uint8.
 * minmax_out (n_clips,4) = raw min, raw max, avg min, avg max (nullable). */
int32_t rm_heatmap(rm_handle* h, const double* bp, int32_t n_clips, int32_t T, int32_t W, int32_t H, uint8_t* heat_out,
                   double* minmax_out, void* workspace, size_t workspace_bytes, void* stream);

/* Tail of locate() (base.py:566-575): cv2.threshold(heat, threshold, 255, THRESH_BINARY), cv2.findContours(RETR_EXTERNAL),
 * max by cv2.contourArea (ties: the contour cv2 lists first), cv2.boundingRect.  heat (n_clips,H,W) uint8 ->
 * roi_out (n_clips,4) int32 x,y,w,h; status_out (n_clips; nullable) = RM_CLIP_OK or RM_CLIP_NO_ROI (locate() -> None,
 * base.py:569-570; the box is then 0,0,0,0). */
int32_t rm_roi_workspace_bytes(rm_handle* h, int32_t W, int32_t H, int32_t n_clips, size_t* out);
int32_t rm_roi_select(rm_handle* h, const uint8_t* heat, int32_t n_clips, int32_t W, int32_t H, int32_t threshold,
                      int32_t* roi_out, int32_t* status_out, void* workspace, size_t workspace_bytes, void* stream);

/* Stand-alone tail of eulerian_magnification_bandpass on a materialised volume (transforms.py:184-192) plus the time
 * average of base.py:562: raw (T,hw) -> clipped_out (T,hw; nullable), avg_out (hw; nullable; mean of the clipped volume,
 * or of raw when clipped_out is null), minmax_out (2; nullable).  workspace: 64 bytes. */
int32_t rm_volume_clip_mean(rm_handle* h, const double* raw, double* clipped_out, double* avg_out, double* minmax_out,
                            int32_t T, int64_t hw, double threshold, void* workspace, void* stream);

/* ------------------------------------------------------------------ measure */
/* Workspace for rm_measure_flow when every ROI is at most max_roi_w x max_roi_h. */
int32_t rm_measure_workspace_bytes(rm_handle* h, int32_t max_roi_w, int32_t max_roi_h, int32_t n_clips, int32_t n_frames,
                                   size_t* out);
/* extract_motion 'flow' over a whole clip (base.py:360-407): for each clip, frames [first_frame, first_frame+n_frames)
 * of (n_clips,T,H,W) uint8 cropped to roi (n_clips,4): LUT crop (transforms.py:26-29), Shi-Tomasi corners on the first
 * frame, pyramidal LK frame to frame with lost points dropped, mean displacement, rolling 2-D PCA.
 * Outputs: data_out (n_clips,n_frames) f64 (the `data` deque, base.py:478), motion_out (n_clips,n_frames,2) f32
 * (row 0 unused; `motion_data`, base.py:389), npts_out (n_clips) corners found, status_io (n_clips; in: RM_CLIP_OK for
 * clips to process, anything else skips the clip; out: NO_CORNERS / TRACK_LOST where they happen). */
int32_t rm_measure_flow(rm_handle* h, const uint8_t* frames, int32_t n_clips, int32_t T, int32_t W, int32_t H,
                        const int32_t* roi, int32_t max_roi_w, int32_t max_roi_h, int32_t first_frame, int32_t n_frames,
                        double* data_out, float* motion_out, int32_t* npts_out, int32_t* status_io, void* workspace,
                        size_t workspace_bytes, void* stream);
/* Diagnostic variant: additionally writes the tracked points after every frame, pts_out (n_clips,n_frames,128,2) f32,
 * NaN padded (the reference's `motion_key_points`, base.py:382). */
int32_t rm_measure_flow_debug(rm_handle* h, const uint8_t* frames, int32_t n_clips, int32_t T, int32_t W, int32_t H,
                              const int32_t* roi, int32_t max_roi_w, int32_t max_roi_h, int32_t first_frame,
                              int32_t n_frames, double* data_out, float* motion_out, int32_t* npts_out,
                              int32_t* status_io, float* pts_out, void* workspace, size_t workspace_bytes, void* stream);
/* extract_motion 'average' (base.py:355-358): mean of the float crop. */
int32_t rm_measure_average(rm_handle* h, const uint8_t* frames, int32_t n_clips, int32_t T, int32_t W, int32_t H,
                           const int32_t* roi, int32_t first_frame, int32_t n_frames, double* data_out, void* stream);
/* measure() for every frame of every clip (base.py:340-352, 312-338): rolling window of the last measure_buffer_len
 * samples of data (n_clips,n_frames): Butterworth filtfilt, peak picking, Gaussian-fit gate, BPM.
 * bpm_out (n_clips,n_frames) f64 (NaN where the reference appends nothing), filtered_out (n_clips, measure_buffer_len)
 * and peaks_out (n_clips, measure_buffer_len) int32 (-1 terminated) for the LAST frame's window, npeaks_out (n_clips). */
int32_t rm_signal_bpm(rm_handle* h, const double* data, int32_t n_clips, int32_t n_frames, double fps, double* bpm_out,
                      double* filtered_out, int32_t* peaks_out, int32_t* npeaks_out, const int32_t* status, void* stream);
/* rm_measure_flow followed by rm_signal_bpm as one pipeline (the 'measure' branch of run(), base.py:464-495, over whole
 * clips): identical results, but the tracker walks the frames in chunks (option "measure_chunks", default 4) on `stream`
 * while the signal stage of the finished chunks runs underneath on a stream owned by the handle; `stream` is joined
 * before the call returns control of it.  Arguments as in the two separate calls; workspace as rm_measure_flow. */
int32_t rm_measure_signal(rm_handle* h, const uint8_t* frames, int32_t n_clips, int32_t T, int32_t W, int32_t H,
                          const int32_t* roi, int32_t max_roi_w, int32_t max_roi_h, int32_t first_frame, int32_t n_frames,
                          double fps, double* data_out, float* motion_out, int32_t* npts_out, int32_t* status_io,
                          double* bpm_out, double* filtered_out, int32_t* peaks_out, int32_t* npeaks_out, void* workspace,
                          size_t workspace_bytes, void* stream);
/* Live streams (the reference's actual use, base.py:464-495 frame by frame): a cohort of n_clips cameras whose measure
 * states started on the same frame, fed a few frames at a time.  `frames` is a ring of ring_len ROI crops per camera
 * (n_clips, ring_len, H, W), the crop of absolute measure frame f in slot f % ring_len (rm_crop_to_ring); each call tracks
 * the new frames [f_begin, f_end) from the tracker state the handle carries over from the previous call (f_begin = 0:
 * corners are detected on frame 0), writes their motion / data / BPM at their absolute positions in the (n_clips, cap)
 * arrays and leaves filtered / peaks / npeaks of the window ending at f_end - 1.  Identical to rm_measure_signal on the
 * whole clip for any block sizes.  One handle per cohort; f_end - f_begin < ring_len; workspace as
 * rm_measure_workspace_bytes(max_roi_w, max_roi_h, n_clips, ring_len). */
int32_t rm_measure_signal_stream(rm_handle* h, const uint8_t* frames, int32_t n_clips, int32_t ring_len, int32_t W,
                                 int32_t H, const int32_t* roi, int32_t max_roi_w, int32_t max_roi_h, int32_t cap,
                                 int32_t f_begin, int32_t f_end, double fps, double* data_out, float* motion_out,
                                 int32_t* npts_out, int32_t* status_io, double* bpm_out, double* filtered_out,
                                 int32_t* peaks_out, int32_t* npeaks_out, void* workspace, size_t workspace_bytes,
                                 void* stream);
/* The same for motion_extraction_method = 'average' (base.py:355-358): the mean of the float crop of the new frames
 * [f_begin, f_end), read from the crop ring (roi = (0, 0, w, h) per camera inside its ring slot), is written to data_out
 * at the frames' absolute positions and measure() (base.py:340-352) runs on the windows ending at them.  No state is
 * carried between calls. */
int32_t rm_measure_average_stream(rm_handle* h, const uint8_t* ring, int32_t n_clips, int32_t ring_len, int32_t W, int32_t H,
                                  const int32_t* roi, int32_t cap, int32_t f_begin, int32_t f_end, double fps,
                                  double* data_out, int32_t* status_io, double* bpm_out, double* filtered_out,
                                  int32_t* peaks_out, int32_t* npeaks_out, void* stream);
/* crop = frame[y:y+h, x:x+w] (base.py:471) of k new frames per camera into the crop ring: frames (n_clips, k, H, W), roi
 * in frame coordinates, ring (n_clips, ring_len, ring_h, ring_w); block frame j goes to slot (f_first + j) % ring_len. */
int32_t rm_crop_to_ring(rm_handle* h, const uint8_t* frames, int32_t n_clips, int32_t k, int32_t W, int32_t H,
                        const int32_t* roi, uint8_t* ring, int32_t ring_len, int32_t ring_w, int32_t ring_h,
                        int32_t f_first, void* stream);
/* assemble the 32-byte records: last finite BPM per clip, ROI, status, n_peaks. */
int32_t rm_pack_results(rm_handle* h, const double* bpm, const int32_t* roi, const int32_t* status, const int32_t* npeaks,
                        int32_t n_clips, int32_t n_frames, rm_result* out, void* stream);

/* ------------------------------------------------------------------ bookkeeping */
/* Number of kernel launches this handle has issued since creation (bench.py reports it as gpu_launches). */
int64_t rm_launch_count(rm_handle* h);

/* the same from the (n_clips, cap) histories of a live cohort of which the first n_valid frames have been measured. */
int32_t rm_pack_results_stream(rm_handle* h, const double* bpm, const int32_t* roi, const int32_t* status,
                               const int32_t* npeaks, int32_t n_clips, int32_t cap, int32_t n_valid, rm_result* out,
                               void* stream);

/* Switches (results never depend on them beyond float64 rounding; they select between equivalent code paths or schedules).
 *   "measure_chunks" (1..16)     frame chunks of rm_measure_signal (default 4).
 *   "temporal_sparse" (0/1)      rm_temporal_bandpass evaluates only the bins the mask keeps (2 K T multiply-adds per column
 *                                instead of two FFTs; same sums as the any-T kernel) where that is cheaper (default 1).
 *   "pyramid_mode" (0/1)         1 (default): uint8 frames whose rows are 16-byte multiples take the one-pass fused pyramid
 *                                kernel (TMA rows, levels 0..4 as integers, the rest in shared memory); 0: always level 3
 *                                through HBM + the tail kernel.  Bit-identical records.
 *   "pyramid_cfg" (0..3)         fused kernel: (ring stages, warps per CTA) = by frame width (0, default) / (4, 18) / (2, 18) /
 *                                (2, 24).
 *   "pyramid_g4" (0..2)          fused kernel: the level-4 image of a frame lives in shared memory (1), in the frame's
 *                                record, where its Laplacian replaces it (2), or wherever more frame slots fit an SM (0,
 *                                default: the record where shared memory would hold a single frame slot, 1080p).
 *   "force_global_lk" (0/1)      track from global memory even when the ROI fits shared memory (the path used for ROIs
 *                                too large to stage).
 *   "force_generic_front" (0/1)  uint8 frames take the float pyramid front kernel instead of the integer one.
 *   "no_minmax_seed" (0/1)       rm_heatmap evaluates every tile-frame (no pruning) and materialises level 2; changes
 *                                what rm_heatmap_workspace_bytes returns, so set it before sizing the workspace. */
int32_t rm_set_option(rm_handle* h, const char* host_name, int64_t value);

/* Per-kernel device timing, the kernel-granular analogue of the reference's tools.Benchmarker (tools.py:60-82): while
 * enabled every launch is bracketed by CUDA events on its own stream.  rm_profile_collect waits for them, folds them into
 * a per-kernel table and returns its size; rm_profile_entry reads row i (name is a static string). */
int32_t rm_profile_enable(rm_handle* h, int32_t on);
int32_t rm_profile_reset(rm_handle* h);
int32_t rm_profile_collect(rm_handle* h);
/* timeline form, valid before rm_profile_collect: i < 0 returns the number of recorded launches; otherwise start / end of
 * launch i in ms relative to the start of launch 0. */
int32_t rm_profile_slot(rm_handle* h, int32_t i, const char** host_name, double* host_start_ms, double* host_end_ms);
int32_t rm_profile_entry(rm_handle* h, int32_t i, const char** host_name, double* host_total_ms, int64_t* host_launches);

#ifdef __cplusplus
}
#endif
#endif /* RESPMON_B200_H */
