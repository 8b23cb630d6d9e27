/*
 * dib.h -- C ABI of libdib.so, the B200 (sm_100a) motion-blur synthesis library.
 *
 * The reference (mohammed-amr/detectInBlur) has no FFI layer: its hot path is Python calling torch
 * (models/blur_functions.py, motion_blur/generate_PSF.py, models/net_transforms.py).  These entry points are what a
 * binding for that path binds instead; each one cites the reference code it replaces.  The thin Python
 * layer in detectinblur_b200/ loads this library with ctypes and keeps the reference's call signatures.
 *
 * Conventions
 *   - every function returns DIB_OK (0) or a negative dib_status; dib_last_error() gives the message of the
 *     last failure on the calling thread;
 *   - all device buffers are caller-owned (torch allocations passed as raw pointers); the library never
 *     allocates, frees or keeps device memory, and holds no mutable global state, so calls on different
 *     streams / devices may run concurrently;
 *   - every launcher takes the cudaStream_t to launch on (as void*), and works on the caller's current device;
 *   - "host" pointers are plain host memory read before the call returns.
 */
#ifndef DIB_H_
#define DIB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIB_ABI_VERSION 1

#if defined(__GNUC__)
#define DIB_API __attribute__((visibility("default")))
#else
#define DIB_API
#endif

typedef enum dib_status {
    DIB_OK = 0,
    DIB_ERR_INVALID = -1,     /* bad argument (message says which) */
    DIB_ERR_CUDA = -2,        /* a CUDA runtime call or launch failed */
    DIB_ERR_UNSUPPORTED = -3, /* valid request this build cannot serve (e.g. reflect pad on a 64-px side) */
    DIB_ERR_CAPACITY = -4     /* caller-provided buffer too small */
} dib_status;

typedef enum dib_dtype { DIB_F32 = 0, DIB_F16 = 1, DIB_F64 = 2 } dib_dtype;

/* Boundary semantics of manual_blur (models/blur_functions.py:17-69). */
typedef enum dib_pad_mode {
    DIB_PAD_REFLECT128 = 0,   /* :43-59  pad (63,64,63,64) reflect,  centre 63 (every side > 64)   */
    DIB_PAD_ZERO128 = 1,      /* :55-56  pad (63,64,63,64) zeros,    centre 63 (a side < 64)       */
    DIB_PAD_REPLICATE256 = 2  /* :17-33  pad (127,128,127,128) replicate, centre 127 (PSF side > 129) */
} dib_pad_mode;

/* One nonzero PSF tap, in the reference's accumulation order (row-major nonzero, blur_functions.py:63). */
typedef struct dib_tap {
    int16_t y;   /* PSF row    (0 .. side-1) */
    int16_t x;   /* PSF column (0 .. side-1) */
    float w;     /* normalised weight (a half PSF's weight widened exactly to fp32) */
} dib_tap;

/* Per-PSF summary produced by dib_compact_taps. */
typedef struct dib_psf_meta {
    int32_t count;            /* nonzero taps of the normalised PSF (may exceed max_taps: then truncated) */
    int16_t ymin, ymax;       /* tap extents in PSF coordinates; utils.py:372-380 (expand_targets) */
    int16_t xmin, xmax;
    float sum;                /* psf.sum() in the PSF dtype, widened (blur_functions.py:98) */
    int32_t support;          /* cells with psf > 0: the support used by the PCA, transforms.py:366 */
    int32_t prog_chunks;      /* chunks of the tiled-kernel program (0: none was built) */
    int32_t prog_steps;       /* total weight vectors (window steps) of that program */
    int32_t flags;            /* DIB_META_* */
    int32_t prog_segs;        /* segments (column-group runs) of that program */
    int16_t prog_group_w;     /* PSF columns per group of that program (2 or 4; every step executes all of them) */
    int16_t prog_shear;       /* columns per row the program's groups move: the tiled kernel's tiles are sheared alike */
    double sy, sx;            /* sums over the support of y, x         (transforms.py:367-371) */
    double syy, sxx, sxy;     /* sums of y*y, x*x, y*x over the support (transforms.py:373-376) */
} dib_psf_meta;

#define DIB_META_TRUNCATED 1      /* more taps than max_taps */
#define DIB_META_NO_PROGRAM 2     /* tiled program not built (capacity / extents): use the generic kernel */

/*
 * Tap set: the device-side product of dib_compact_taps, consumed by dib_blur_batch.  Layout of the
 * caller-owned buffer (all sections 256-byte aligned, sizes from dib_tapset_layout):
 *   meta   : dib_psf_meta[n]
 *   taps   : dib_tap[n][max_taps]
 *   prog   : uint8[n][prog_bytes]     program of the tiled kernel (opaque)
 *   sched  : 256 bytes                work-distribution words of the tiled kernel (zeroed by dib_compact_taps, left
 *                                     zero by every dib_blur_batch; one blur call per tap set may be in flight at a time)
 */
typedef struct dib_tapset_layout {
    size_t meta_offset, taps_offset, prog_offset, prog_bytes_per_psf, sched_offset, total_bytes;
} dib_tapset_layout;

/* One image of a blur batch (a CHW plane stack anywhere in device memory). */
typedef struct dib_image {
    const void* src;          /* C x H x W input, element pitches below */
    void* dst;                /* C x H x W output (must not alias src) */
    const void* noise;        /* optional pre-drawn N(0,1) tensor, dst layout/dtype (blur_functions.py:74) or NULL */
    int32_t C, H, W;
    int32_t psf_index;        /* which PSF of the tap set; < 0: copy src to dst through the epilogue only */
    int64_t src_row_pitch, src_chan_pitch;   /* in elements */
    int64_t dst_row_pitch, dst_chan_pitch;   /* in elements */
    int32_t pad_mode;         /* dib_pad_mode */
    int32_t epilogue;         /* DIB_EPI_* flags */
    float noise_sd;           /* sqrt(noise_var) (blur_functions.py:73-74) */
    float gamma;              /* exponent for DIB_EPI_GAMMA */
    float mean[4];            /* per-channel mean for DIB_EPI_NORMALIZE (net_transforms.py:135-139) */
    float std[4];             /* per-channel std  (true division, as the reference) */
} dib_image;

#define DIB_EPI_NOISE 1       /* out += noise * noise_sd     (blur_functions.py:74)                     */
#define DIB_EPI_CLAMP 2       /* out = clamp(out, 0, 1)       (blur_functions.py:74; only with noise there) */
#define DIB_EPI_GAMMA 4       /* out = pow(out, gamma)        (transforms.py:182 gammaFunc; unused upstream) */
#define DIB_EPI_NORMALIZE 8   /* out = (out - mean[c]) / std[c]  (net_transforms.py:135-139)             */
#define DIB_EPI_PHILOX 16     /* with DIB_EPI_NOISE and noise == NULL: draw N(0,1) in-kernel (Philox4x32-10) */

#define DIB_MAX_BATCH 32      /* images per dib_blur_batch call (descriptors travel as kernel parameters) */

/* Kernel selection for dib_blur_batch. */
#define DIB_ALGO_AUTO 0       /* tiled kernel where eligible, generic kernel for the rest */
#define DIB_ALGO_GENERIC 1    /* exact-order kernel: per tap one rounded multiply and one rounded add, bit-identical
                                 to the reference loop for fp32 and fp16 */
#define DIB_ALGO_TILED 2      /* fail instead of falling back */
/* Flags OR-ed into `algo`.  DIB_ALGO_OVERLAP: the caller asserts that this batch does not depend on the work launched just
 * before it on the stream (typically the previous, independent batch), so the tiled kernel may start on SMs that kernel has
 * already vacated (programmatic dependent launch: its tail overlaps this launch's ramp-up) instead of waiting for the whole
 * grid.  Launches that may overlap and share a tap set must use different scheduler slots: DIB_ALGO_SLOT(k), k = a launch
 * counter modulo 4.  Without the flag a launch is ordered after everything before it on the stream, as usual. */
#define DIB_ALGO_OVERLAP 0x100
/* DIB_ALGO_DEVICE_PLAN: plan the launch on the device.  No host copy of the per-PSF summaries is needed (meta_host may be
 * NULL), so nothing has to be read back between dib_compact_taps and dib_blur_batch and the whole chain rasterise -> compact
 * -> blur can be enqueued without a host synchronisation (and captured in a CUDA graph): both tiled kernels and the exact-
 * order kernel are launched and each takes, from the summaries in the tap set, the images that are its own (half images:
 * the masked kernel for small PSFs, the exact-order kernel -- the reference's half loop -- for the others).  max_taps must
 * hold every PSF of the tap set (a truncated PSF falls to the exact-order kernel with its first
 * max_taps taps -- size the list for the worst case, 4096 covers every 128 x 128 motion PSF). */
#define DIB_ALGO_DEVICE_PLAN 0x200
#define DIB_ALGO_SLOT(k) (((k) & 3) << 12)

DIB_API int dib_abi_version(void);
DIB_API const char* dib_last_error(void);
/* Number of SMs / compute capability of the current device (major*10+minor); for launch planning and logs. */
DIB_API int dib_device_info(int* sm_count, int* cc);

/*
 * Tap compaction.  Replaces `psf_GPU/psf_GPU.sum()` (blur_functions.py:98) and `psf_GPU.nonzero()` with the
 * two per-tap device->host index reads of the loop (blur_functions.py:63-67), for a whole batch in one launch.
 *   psfs        n dense side x side PSFs, `psf_stride` elements apart, dtype DIB_F32 or DIB_F16
 *   normalize   flags: DIB_COMPACT_NORMALIZE (1) divides by the PSF's sum in the PSF dtype (blur_image_list; without it the PSF
 *               is taken as already normalised, manual_blur); DIB_COMPACT_DENSE_ONLY (2) gives every PSF the dense sheared
 *               program of the TMA-staged tiled kernel, also the small ones that would otherwise get the masked kernel's
 *   tapset      caller-owned device buffer of dib_tapset_layout(...).total_bytes
 * Taps come out in row-major order; with normalize = 0 the weights are the PSF's own values, with normalize = 1 the sum is
 * accumulated in fp64 and rounded once to the PSF dtype -- torch's result for every PSF on the fp16 grid that sums to at most
 * 1 (all stored / generated PSFs); a general fp32 PSF can differ from torch's reduction tree by one ulp of the sum (the
 * wrapper's exact mode normalises such PSFs with torch and passes normalize = 0).  meta carries counts, extents, PCA moments.
 */
#define DIB_COMPACT_NORMALIZE 1
#define DIB_COMPACT_DENSE_ONLY 2
#define DIB_COMPACT_MASKED_ONLY 4   /* measurements: the masked kernel's program for every PSF it can hold (several chunks) */
DIB_API int dib_tapset_layout_for(int n_psfs, int max_taps, dib_tapset_layout* out);
DIB_API int dib_compact_taps(const void* psfs, int psf_dtype, int n_psfs, int side, int64_t psf_stride, int normalize,
                     void* tapset, int max_taps, void* stream);

/*
 * Batched sparse-PSF blur.  Replaces manual_blur's pad / per-tap roll-multiply-add loop / crop
 * (blur_functions.py:17-69), its noise+clamp epilogue (:72-74), blur_image_list's Python loop (:92-100) and,
 * with DIB_EPI_NORMALIZE, GeneralizedRCNNTransform.normalize (net_transforms.py:135-139).
 *   images      host array of n_images (<= DIB_MAX_BATCH) descriptors
 *   tapset      device buffer filled by dib_compact_taps;  meta_host: host copy of its meta section
 *               (n_psfs entries) used for launch planning, or NULL to force the generic kernel
 *   io_dtype    DIB_F32 or DIB_F16 (src, dst and noise share it)
 *   philox_seed / philox_offset   counter-based RNG stream for DIB_EPI_PHILOX
 *   launches    if not NULL, receives the number of kernels launched
 */
DIB_API int dib_blur_batch(const dib_image* images, int n_images, void* tapset, int n_psfs, int max_taps,
                   const dib_psf_meta* meta_host, int io_dtype, int algo, uint64_t philox_seed,
                   uint64_t philox_offset, int* launches, void* stream);

/*
 * PSF rasterisation.  Replaces PSF.fit (motion_blur/generate_PSF.py:31-77), PSF.centerPSF (:106-123), the
 * [64:192] crop (transforms.py:334-335) and the float16 cast of the stored bank (dataset_utils/generate_PSFs.py:60).
 *   traj        n x iters complex128 samples (re, im interleaved doubles), canvas-centred as Trajectory.x
 *   fractions   n exposure fractions (double)
 *   canvas      256 in the reference;  center: apply centerPSF;  out_side: canvas or 128 (central crop)
 *   out         n x out_side x out_side in out_dtype (DIB_F64 / DIB_F32 / DIB_F16), bit-identical to the
 *               reference's fp64 raster (then rounded once)
 *   offsets     optional n x 2 int32 (offsetX, offsetY) of centerPSF, or NULL
 *   scratch     n * canvas * canvas doubles of caller-owned device scratch
 */
DIB_API int dib_rasterize_psf(const double* traj, const double* fractions, int n, int iters, int canvas, int center,
                      int out_side, void* out, int out_dtype, int32_t* offsets, double* scratch, void* stream);

/*
 * Camera-shake trajectories on the device: the random walk of motion_blur/generate_trajectory.py:38-98 (Trajectory.fit)
 * with a counter-based generator instead of numpy's sequential global stream -- statistically, not bit-wise, equal to
 * the reference (it removes ~18 ms of host Python per PSF from the on-the-fly path).  Trajectory k of the call is a pure
 * function of (seed, index_k), index_k = indices[k] when `indices` (n device uint64) is given, else first_index + k.
 *   expl        n per-trajectory `expl` parameters (0.005 / 0.001 / 0.00005 on the blur path)
 *   out         n x iters complex128 samples (re = column, im = row), start point at canvas / 2, the layout
 *               dib_rasterize_psf consumes
 *   big_count   optional n int32: number of impulsive shakes (Trajectory.big_expl_count)
 */
DIB_API int dib_generate_trajectories(uint64_t seed, uint64_t first_index, const uint64_t* indices, int n, int iters,
                              double max_len, double canvas, const double* expl, double* out, int32_t* big_count,
                              void* stream);

/*
 * One image of a fused normalize + bilinear resize + zero-padded batch pass (dib_resize_batch).
 */
typedef struct dib_resize_image {
    const void* src;          /* C x in_h x in_w source (the blurred image), rows contiguous */
    void* dst;                /* this image's C planes inside the padded batch tensor */
    int32_t C, in_h, in_w;
    int32_t out_h, out_w;     /* resized extent, floor(in * scale_factor) as torch computes it */
    int32_t pad_h, pad_w;     /* extent of the destination plane to fill: zeros outside out_h x out_w */
    int32_t normalize;        /* != 0: (x - mean[c]) / std[c] */
    int64_t src_row_pitch, src_chan_pitch, dst_row_pitch, dst_chan_pitch;    /* in elements */
    float mean[4], std[4];
} dib_resize_image;

/*
 * Normalize, resize (bilinear, align_corners=False, recomputed scale) and batch up to DIB_MAX_BATCH images in one
 * pass.  Replaces GeneralizedRCNNTransform.forward's per-image normalize (models/net_transforms.py:135-139), resize
 * (:36-48, :151-175) and batch_images (:218-249).  *launches (optional) receives the number of kernels launched.
 */
DIB_API int dib_resize_batch(const dib_resize_image* images, int n_images, int io_dtype, int* launches, void* stream);

/*
 * Device-side reader of the packed sparse PSF bank: expands stored taps into the dense PSFs the reference's reader
 * yields.  Replaces, for a whole batch, transforms.py:301-309 (np.load of a 131 KB float16[256,256] file per image, then
 * the [64:192, 64:192] crop) followed by engine.py:84 (torch.HalfTensor(blur_dict["psf"]).to(device), one dense
 * upload per image): the caller uploads only the taps.
 *   taps        device array of packed taps: y | x << 8 | (fp16 bits) << 16, coordinates on the bank's 256 x 256 canvas,
 *               row-major nonzero order per PSF
 *   offsets     device array of n + 1 tap offsets (PSF k owns taps[offsets[k] .. offsets[k+1]))
 *   crop_lo     first canvas row / column of the crop (64 for the reference reader, 0 for the whole canvas)
 *   out         n x out_side x out_side dense PSFs of out_dtype (zero filled here; fp16 values are copied bit for bit)
 */
DIB_API int dib_unpack_psfs(const uint32_t* taps, const int64_t* offsets, int n, int crop_lo, int out_side, void* out,
                    int out_dtype, void* stream);

/*
 * Order-independent 64-bit checksum of a buffer's raw element bits (for cross-shard verification; the
 * multi-GPU path all-gathers one value per rank, mirroring utils.all_gather, utils.py:536-576).
 * `out` is one uint64 in device memory; accumulate != 0 adds to its current value.
 */
DIB_API int dib_checksum(const void* data, int dtype, int64_t n_elements, uint64_t* out, int accumulate, void* stream);

/*
 * uint8 <-> float image planes.  dib_u8_to_float replaces torchvision's to_tensor scaling (byte / 255, the input of
 * engine.py:80's upload) on the device, so that a caller can upload bytes instead of floats; dib_float_to_u8 is the uint8
 * result of the --cpu_blur path (motion_blur/blur_image.py:147: 255 * x, clipped, truncated).  `rows` counts image rows over
 * all channels (C * H); the float side may be pitched (`*_row_pitch` in elements), the uint8 side is dense.
 */
DIB_API int dib_u8_to_float(const uint8_t* src, void* dst, int dst_dtype, int64_t rows, int W, int64_t dst_row_pitch, void* stream);
DIB_API int dib_float_to_u8(const void* src, int src_dtype, uint8_t* dst, int64_t rows, int W, int64_t src_row_pitch, void* stream);

/*
 * FP32 FMA-pipe probe used by bench.py for the compute roofline: launches `iters` dependent-chain-free FFMA
 * batches on every SM and writes the FMA count to *fma_count; the caller times it with CUDA events.
 */
DIB_API int dib_fp32_probe(int iters, float* sink, uint64_t* fma_count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIB_H_ */
