// dib_blur_batch: validation, kernel selection and launch planning for the batched blur (include/dib.h).
//
// Stands where the reference has blur_image_list's Python loop (models/blur_functions.py:92-100) calling
// manual_blur (:11-89) once per image: one call here blurs the whole batch with at most two launches (tiled
// kernel for eligible images, exact-order kernel for the rest).
#include "dib_common.cuh"

namespace dib {
int launch_generic(const dib_image* images, int n_images, const dib_tap* taps, const dib_psf_meta* meta, int max_taps,
                   int io_dtype, uint32_t skip_mask, uint32_t planned_mask, uint64_t seed, uint64_t offset, cudaStream_t st);
// dense sheared program (blur_tiled.cu): large PSFs, float32 images
int launch_tiled(const dib_image* images, const int* order, int n_sel, const dib_psf_meta* meta_host, const uint8_t* prog,
                 SchedWords* sched, uint64_t seed, uint64_t offset, int io_dtype, bool overlap_prev, const dib_psf_meta* meta_dev,
                 cudaStream_t st);
namespace mk {
// masked program (blur_masked.cu): PSFs whose support fits one chunk, float32 and float16 images
int launch_tiled(const dib_image* images, const int* order, int n_sel, const dib_psf_meta* meta_host, const uint8_t* prog,
                 SchedWords* sched, uint64_t seed, uint64_t offset, int io_dtype, bool overlap_prev, const dib_psf_meta* meta_dev,
                 cudaStream_t st);
}  // namespace mk

enum { kNotTiled = 0, kMasked = 1, kDense = 2 };

enum { kPlanned = 3 };      // device-planned: a tiled kernel takes the image if its PSF has a program (decided on the device)

// which tiled kernel, if any, takes this image
static int tiled_kind(const dib_image& im, const dib_psf_meta* meta_host, int io_dtype, bool device_plan) {
    if (device_plan) {
        // the image-side conditions only; which kernel takes the image is decided on the device from its PSF's summary
        if (im.psf_index < 0) return kNotTiled;
        if ((im.pad_mode != DIB_PAD_REFLECT128 && im.pad_mode != DIB_PAD_ZERO128) || im.H <= 64 || im.W <= 64) return kNotTiled;
        if (io_dtype == DIB_F16) {
            // half images: the masked kernel if the PSF is small, else the exact-order kernel (the dense kernel is float32 only)
            if ((reinterpret_cast<uintptr_t>(im.dst) & 15u) || (im.dst_row_pitch & 7) || (im.dst_chan_pitch & 7)) return kNotTiled;
            if ((reinterpret_cast<uintptr_t>(im.src) & 1u) || (im.epilogue & ~DIB_EPI_NORMALIZE)) return kNotTiled;
            return kPlanned;
        }
        if ((reinterpret_cast<uintptr_t>(im.src) | reinterpret_cast<uintptr_t>(im.dst)) & 3u) return kNotTiled;
        return kPlanned;
    }
    if (meta_host == nullptr || im.psf_index < 0) return kNotTiled;
    // reflect-101 or zero padding about centre 63; tiny images (the reference's own zero-padding case) stay on the generic kernel
    if ((im.pad_mode != DIB_PAD_REFLECT128 && im.pad_mode != DIB_PAD_ZERO128) || im.H <= 64 || im.W <= 64) return kNotTiled;
    const dib_psf_meta& m = meta_host[im.psf_index];
    if (m.count <= 0 || m.prog_chunks <= 0 || (m.flags & (DIB_META_NO_PROGRAM | DIB_META_TRUNCATED))) return kNotTiled;
    const int kind = m.prog_group_w == 0 ? kMasked : kDense;
    if (io_dtype == DIB_F16) {
        // half I/O (masked kernel only): fp32 accumulation, rounded to half once (the reference's half loop rounds after every
        // tap -- callers that need its bits pass DIB_ALGO_GENERIC).  Needs 16-byte-aligned destination rows and at most the
        // normalize epilogue.  Half images with a large PSF go through the wrapper's casts around the float32 path.
        if (kind != kMasked) return kNotTiled;
        if ((reinterpret_cast<uintptr_t>(im.dst) & 15u) || (im.dst_row_pitch & 7) || (im.dst_chan_pitch & 7)) return kNotTiled;
        if (reinterpret_cast<uintptr_t>(im.src) & 1u) return kNotTiled;
        if (im.epilogue & ~DIB_EPI_NORMALIZE) return kNotTiled;
    } else if ((reinterpret_cast<uintptr_t>(im.src) | reinterpret_cast<uintptr_t>(im.dst)) & 3u) {
        return kNotTiled;
    }
    return kind;
}
}  // namespace dib

extern "C" int dib_blur_batch(const dib_image* images, int n_images, void* tapset, int n_psfs, int max_taps,
                              const dib_psf_meta* meta_host, int io_dtype, int algo, uint64_t philox_seed,
                              uint64_t philox_offset, int* launches, void* stream) {
    using namespace dib;
    if (launches) *launches = 0;
    DIB_CHECK_ARG(images != nullptr, "dib_blur_batch: images is NULL");
    DIB_CHECK_ARG(n_images >= 0 && n_images <= DIB_MAX_BATCH, "dib_blur_batch: n_images %d outside [0, %d]", n_images, DIB_MAX_BATCH);
    DIB_CHECK_ARG(io_dtype == DIB_F32 || io_dtype == DIB_F16, "dib_blur_batch: io_dtype must be DIB_F32 or DIB_F16");
    const bool overlap_prev = (algo & DIB_ALGO_OVERLAP) != 0;
    const int sched_slot = (algo >> 12) & (kSchedSlots - 1);
    const bool device_plan = (algo & DIB_ALGO_DEVICE_PLAN) != 0;
    DIB_CHECK_ARG((algo & ~(0xff | DIB_ALGO_OVERLAP | DIB_ALGO_DEVICE_PLAN | DIB_ALGO_SLOT(3))) == 0, "dib_blur_batch: unknown algo flags 0x%x", algo);
    algo &= 0xff;
    DIB_CHECK_ARG(algo == DIB_ALGO_AUTO || algo == DIB_ALGO_GENERIC || algo == DIB_ALGO_TILED, "dib_blur_batch: unknown algo %d", algo);
    if (n_images == 0) return DIB_OK;
    bool any_psf = false;
    for (int k = 0; k < n_images; ++k) {
        const dib_image& im = images[k];
        DIB_CHECK_ARG(im.src != nullptr && im.dst != nullptr, "dib_blur_batch: image %d has a NULL buffer", k);
        DIB_CHECK_ARG(im.src != im.dst, "dib_blur_batch: image %d: dst must not alias src", k);
        DIB_CHECK_ARG(im.C >= 1 && im.H >= 1 && im.W >= 1, "dib_blur_batch: image %d has an empty shape %dx%dx%d", k, im.C, im.H, im.W);
        DIB_CHECK_ARG(im.src_row_pitch >= im.W && im.dst_row_pitch >= im.W, "dib_blur_batch: image %d: row pitch smaller than W", k);
        DIB_CHECK_ARG(im.psf_index < n_psfs, "dib_blur_batch: image %d: psf_index %d >= n_psfs %d", k, im.psf_index, n_psfs);
        DIB_CHECK_ARG(im.pad_mode >= DIB_PAD_REFLECT128 && im.pad_mode <= DIB_PAD_REPLICATE256, "dib_blur_batch: image %d: bad pad_mode", k);
        if (im.psf_index >= 0 && im.pad_mode == DIB_PAD_REFLECT128 && (im.H <= 64 || im.W <= 64)) {
            // the reference raises here: torch reflect padding needs pad (64) < dim  (blur_functions.py:57-59)
            set_error("dib_blur_batch: image %d (%dx%d): reflect padding of 64 needs every side > 64", k, im.H, im.W);
            return DIB_ERR_UNSUPPORTED;
        }
        if ((im.epilogue & DIB_EPI_NORMALIZE) && im.C > 4) {
            set_error("dib_blur_batch: image %d: fused normalize supports at most 4 channels", k);
            return DIB_ERR_UNSUPPORTED;
        }
        any_psf |= im.psf_index >= 0;
    }
    DIB_CHECK_ARG(!any_psf || (tapset != nullptr && n_psfs > 0 && max_taps > 0), "dib_blur_batch: tap set missing");
    const dib_tapset_layout L = tapset_layout(n_psfs > 0 ? n_psfs : 1, max_taps > 0 ? max_taps : 1);
    uint8_t* base = static_cast<uint8_t*>(tapset);
    const dib_psf_meta* meta_dev = reinterpret_cast<const dib_psf_meta*>(base + L.meta_offset);
    const dib_tap* taps = reinterpret_cast<const dib_tap*>(base + L.taps_offset);
    const uint8_t* prog = base + L.prog_offset;
    SchedWords* sched = reinterpret_cast<SchedWords*>(base + L.sched_offset) + sched_slot;
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    // split the batch: a tiled kernel where eligible (heaviest PSFs first), exact-order kernel for the rest
    int order[2][DIB_MAX_BATCH];
    int n_kind[2] = {0, 0};
    uint32_t tiled_mask = 0;
    uint32_t planned_mask = 0;
    if (algo != DIB_ALGO_GENERIC) {
        for (int k = 0; k < n_images; ++k) {
            const int kind = tiled_kind(images[k], meta_host, io_dtype, device_plan);
            if (kind == kPlanned) {             // listed for both tiled kernels; each takes its own on the device
                order[0][n_kind[0]++] = k;
                if (io_dtype == DIB_F32) order[1][n_kind[1]++] = k;
                planned_mask |= 1u << k;
            } else if (kind != kNotTiled) {
                order[kind - 1][n_kind[kind - 1]++] = k;
                tiled_mask |= 1u << k;
            } else if (algo == DIB_ALGO_TILED) {
                set_error("dib_blur_batch: image %d is not eligible for a tiled kernel (reflect or zero mode, sides > 64, PSF with a "
                          "tiled program and host meta; half images: a PSF within 18 x 20 cells, 16-byte-aligned destination rows, "
                          "normalize epilogue at most)", k);
                return DIB_ERR_UNSUPPORTED;
            }
        }
        // insertion sort by tap count, descending (stable): the dynamic scheduler hands out long tiles first
        for (int q = 0; q < 2 && !device_plan; ++q) {
            for (int a = 1; a < n_kind[q]; ++a) {
                const int v = order[q][a];
                const int cv = meta_host[images[v].psf_index].count;
                int b = a - 1;
                while (b >= 0 && meta_host[images[order[q][b]].psf_index].count < cv) {
                    order[q][b + 1] = order[q][b];
                    --b;
                }
                order[q][b + 1] = v;
            }
        }
    }
    const dib_psf_meta* plan_meta = device_plan ? meta_dev : nullptr;
    const int n_sel = device_plan ? 0 : n_kind[0] + n_kind[1];
    int nl = 0;
    if (n_kind[0] > 0) {
        const int rc = mk::launch_tiled(images, order[0], n_kind[0], meta_host, prog, sched, philox_seed, philox_offset, io_dtype, overlap_prev,
                                        plan_meta, st);
        if (rc != DIB_OK) return rc;
        ++nl;
    }
    if (n_kind[1] > 0) {
        // its own scheduler words: the two tiled launches of one call may be co-resident -- and are meant to be: they work on
        // different images of the batch, so the second one never waits for the first (it starts on the SMs that one vacates)
        const int rc = launch_tiled(images, order[1], n_kind[1], meta_host, prog, sched + kSchedSlots, philox_seed, philox_offset, io_dtype,
                                    overlap_prev || n_kind[0] > 0, plan_meta, st);
        if (rc != DIB_OK) return rc;
        ++nl;
    }
    if (n_sel < n_images) {
        const int rc = launch_generic(images, n_images, taps, meta_dev, max_taps, io_dtype, tiled_mask, planned_mask, philox_seed, philox_offset, st);
        if (rc != DIB_OK) return rc;
        ++nl;
    }
    if (launches) *launches = nl;
    return DIB_OK;
}
