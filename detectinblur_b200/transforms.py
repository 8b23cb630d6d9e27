"""Mirror of the blur part of the reference's ``transforms.py``: ``BlurImage`` (:186-463) and
``add_jpeg_artifact_to_image`` (:467-493).

``BlurImage`` keeps the reference's constructor, ``__call__(image, target=None, blur_dict={})`` contract, blur_dict keys
and -- importantly -- its consumption of python's ``random`` stream (selection draws, including the ones the stored-PSF
branch makes and discards, transforms.py:252-298) and of numpy's global stream (two trajectories per PSF,
transforms.py:316-317), so a seeded data pipeline picks the same blur for the same sample.

What changed underneath:
  * on-the-fly PSFs are rasterised by the CUDA kernel (psf_ops.rasterize_psfs) instead of the ~45 ms Python splat loop.
    DataLoader workers must not touch CUDA, so there the transform only draws the trajectory and DEFERS rasterisation
    (``psf_backend="defer"``; blur_dict carries "trajectory"/"fraction" and psf=None); the main process completes the
    batch in one launch with ``complete_blur_dicts`` / ``upload_psfs`` (called by blur_image_list automatically);
  * ``blur_image_in_transform=True`` (``--cpu_blur``) keeps the Fourier path's semantics (edge padding, per-image
    min-max contrast stretch, uint8 truncation; motion_blur/blur_image.py) but evaluates them as a tap sum on the CUDA
    kernels (detectinblur_b200.motion_blur.BlurImageHandler); it needs a CUDA-capable process and raises otherwise.
    The FFT implementation only survives as the timed CPU baseline (oracle/).
"""
import math
import random

import numpy as np
import torch

from . import psf_bank
from . import psf_ops
from .motion_blur.generate_trajectory import Trajectory

PARAMS = [0.005, 0.001, 0.00005]                 # transforms.py:248
FRACTIONS = [1 / 18, 1 / 10, 1 / 5, 1 / 2, 1]    # transforms.py:249
LEHE_WEIGHTS = [0.0625, 0.0625, 0.0625, 0.375, 0.375]


def sigmoid(x):
    return 1 / (1 + math.exp(-x))


def psf_principal_components(psf):
    """transforms.py:366-385 on a host PSF: (theta_rad, scale_factor_lambda1, scale_factor_lambda2)."""
    ys, xs = np.nonzero(psf > 0)
    yp, xp = ys - ys.mean(), xs - xs.mean()
    cov = (yp * xp).mean()
    var_x, var_y = (xp * xp).mean(), (yp * yp).mean()
    root = math.sqrt(math.pow((var_x - var_y) / 2, 2) + math.pow(cov, 2))
    lambda1 = (var_x + var_y) / 2 + root
    lambda2 = (var_x + var_y) / 2 - root
    s1 = 1 - (sigmoid(math.sqrt(lambda1) / 10) - 0.5) * 0.6
    s2 = 1 - (sigmoid(math.sqrt(lambda2) / 10) - 0.5) * 0.6
    return -math.atan2(lambda1 - var_x, -cov), s1, s2


def _in_worker():
    info = torch.utils.data.get_worker_info()
    return info is not None


class BlurImage(object):
    def __init__(self, prob=0.5, blur_type=None, blur_exposure=None, use_stored_psfs=False, stored_psf_directory=None,
                 blur_image_in_transform=True, dont_center_psf=False, low_exposure=False, high_exposure=False,
                 dilate_psf=False, LEHE_blur_seg=False, psf_backend="auto", device=None, trajectory_backend="host",
                 trajectory_seed=0):
        self.prob = prob
        self.blur_type = blur_type
        self.blur_exposure = blur_exposure
        self.use_stored_psf = use_stored_psfs
        self.stored_psf_directory = stored_psf_directory
        self.blur_image_in_transform = blur_image_in_transform
        self.dont_center_psf = dont_center_psf
        self.LEHE_blur_seg = LEHE_blur_seg
        self.low_exposure = low_exposure
        self.high_exposure = high_exposure
        self.dilate_psf = dilate_psf
        if psf_backend not in ("auto", "cuda", "defer"):
            raise ValueError("psf_backend must be 'auto', 'cuda' or 'defer'")
        self.psf_backend = psf_backend
        self.device = device
        # "host": the reference's walk on numpy's global stream (bit-identical PSFs for a seeded pipeline).
        # "cuda": opt-in -- the walk is drawn on the GPU from a counter-based generator (statistically equal, ~18 ms of
        # host Python per PSF gone); the sample's key comes from python's `random`, numpy's stream is not consumed.
        if trajectory_backend not in ("host", "cuda"):
            raise ValueError("trajectory_backend must be 'host' or 'cuda'")
        self.trajectory_backend = trajectory_backend
        self.trajectory_seed = trajectory_seed
        if self.blur_image_in_transform:
            print("Blurring internally in transform on the GPU (detectinblur_b200; the reference's CPU Fourier path is replaced).")
        else:
            print("Not blurring internally on CPU.")
        self.count = 0

    # -- selection draws (python `random`), in the reference's order --------------------------------------------
    def _draw_fraction_index(self, stored):
        if self.high_exposure:
            return random.choice([3, 4]) if stored else random.choice(range(len(FRACTIONS[3:]))) + 3
        if self.low_exposure:
            return random.choice([0, 1, 2]) if stored else random.choice(range(len(FRACTIONS[:3])))
        if self.LEHE_blur_seg:
            pop = [0, 1, 2, 3, 4] if stored else range(len(FRACTIONS))
            return random.choices(pop, weights=LEHE_WEIGHTS)[0]
        return random.choice([0, 1, 2, 3, 4]) if stored else random.choice(range(len(FRACTIONS)))

    def _backend(self):
        if self.psf_backend != "auto":
            return self.psf_backend
        return "defer" if (_in_worker() or not torch.cuda.is_available()) else "cuda"

    def __call__(self, image, target=None, blur_dict={}):
        if "preBlurred" in blur_dict and blur_dict["preBlurred"]:            # transforms.py:225-235
            blur_dict.update(blurring=False, psf=[0], inverseWarp=None, theta_rad=0, scale_factor_lambda1=1,
                             scale_factor_lambda2=1, param_index=None, fraction_index=None)
            return image, target, blur_dict

        threshold = 1 - 0.0625 if self.LEHE_blur_seg else self.prob          # :238-241
        if not (random.random() < threshold):                                # :244
            blur_dict.update(blurring=False, psf=[0], theta_rad=0, scale_factor_lambda1=1, scale_factor_lambda2=1,
                             param_index=None, fraction_index=None)          # :455-461
            return image, target, blur_dict

        # first round of draws (:252-273); the stored branch below redraws and discards these
        fraction_index = None
        if self.blur_exposure is not None:
            fraction = self.blur_exposure
        else:
            fraction_index = self._draw_fraction_index(stored=False)
            fraction = FRACTIONS[fraction_index]
        param_index = None
        if self.blur_type is not None:
            param = self.blur_type
        else:
            param_index = random.choice(range(len(PARAMS)))
            param = PARAMS[param_index]

        deferred = None
        if self.use_stored_psf:                                              # :276-309
            param_index = self.blur_type if self.blur_type is not None else random.choice([1, 2, 3])
            fraction_index = self.blur_exposure if self.blur_exposure is not None else self._draw_fraction_index(stored=True)
            psf_index = random.randint(0, 12000 - 1)
            # the reference's file read + 128 crop (:301-309), served from a packed bank when one is present
            psf = psf_bank.load_stored_psf(self.stored_psf_directory, param_index, fraction_index, psf_index)
            blur_dict["stored_psf_source"] = (self.stored_psf_directory, param_index, fraction_index, psf_index)
        else:                                                                # :311-335
            center = not self.dont_center_psf
            side = 128 if center else 256
            if self.trajectory_backend == "cuda":
                key = random.getrandbits(62)
                if self._backend() == "cuda":
                    dev = self.device if self.device is not None else torch.device("cuda")
                    traj = psf_ops.generate_trajectories(1, param, self.trajectory_seed, dev, indices=[key])
                    psf = psf_ops.rasterize_psfs(traj, [fraction], dev, canvas=256, center=center, out_side=side,
                                                 dtype=torch.float64)[0].cpu().numpy()
                else:
                    psf = None
                    deferred = {"trajectory": None, "trajectory_key": (self.trajectory_seed, key), "expl": param,
                                "fraction": fraction, "center": center, "side": side}
            else:
                trajectory = Trajectory(canvas=256, max_len=96, expl=param).fit().fit()   # two walks drawn, the second kept
                if self._backend() == "cuda":
                    dev = self.device if self.device is not None else torch.device("cuda")
                    psf = psf_ops.rasterize_psfs(trajectory.x[None], [fraction], dev, canvas=256, center=center, out_side=side,
                                                 dtype=torch.float64)[0].cpu().numpy()
                else:
                    psf = None
                    deferred = {"trajectory": trajectory.x, "fraction": fraction, "center": center, "side": side}

        if self.dilate_psf:                                                  # :338-342 (host scipy, as the reference)
            if psf is None:
                raise RuntimeError("dilate_psf needs the dense PSF: use psf_backend='cuda' (not available in DataLoader workers)")
            import scipy.ndimage
            sigma = np.random.uniform(low=0, high=3)
            psf = scipy.ndimage.gaussian_filter(psf, sigma)
            psf = psf / psf.max()

        output_image = image
        if self.blur_image_in_transform:                                     # :344-359, the Fourier path's result on the GPU
            if psf is None:
                raise RuntimeError("blur_image_in_transform=True needs a CUDA-capable process (psf_backend='cuda'); "
                                   "DataLoader workers cannot blur -- use --gpu_blur semantics (blur_image_in_transform=False)")
            output_image = blur_pil_image(image, psf, self.device)
            self.pilImageResult = output_image

        blur_dict["blurring"] = True
        blur_dict["psf"] = psf
        if psf is not None:
            theta, s1, s2 = psf_principal_components(psf)                    # :366-385
            blur_dict["theta_rad"], blur_dict["scale_factor_lambda1"], blur_dict["scale_factor_lambda2"] = theta, s1, s2
        else:
            blur_dict["deferred_psf"] = deferred
            blur_dict["theta_rad"] = blur_dict["scale_factor_lambda1"] = blur_dict["scale_factor_lambda2"] = None
        self.count += 1

        if self.blur_type is not None:                                       # :418-428
            blur_dict["param_index"] = np.argmin(np.abs(np.asarray(PARAMS) - self.blur_type))
            if self.use_stored_psf:
                blur_dict["param_index"] = blur_dict["param_index"] - 1
        else:
            blur_dict["param_index"] = param_index
            if self.use_stored_psf:
                blur_dict["param_index"] = blur_dict["param_index"] - 1
        if self.blur_exposure is not None:                                   # :436-446
            blur_dict["fraction_index"] = np.argmin(np.abs(np.asarray(FRACTIONS) - self.blur_exposure))
            if self.blur_exposure < 1 / 90:
                blur_dict["fraction_index"] = -1
        else:
            blur_dict["fraction_index"] = fraction_index
        return output_image, target, blur_dict


def complete_blur_dicts(blur_dicts, device, dtype=torch.float16):
    """Finish, in the main process, the blur_dicts whose PSF a DataLoader worker deferred: one batched rasterisation
    launch for all of them, then the dense PSF (numpy, as the reference's blur_dict["psf"]) and the PCA scalars
    (transforms.py:366-385) are filled in place.  Returns the list of PSF tensors on ``device`` (None where not blurring),
    i.e. what engine.py:84 builds with ``torch.HalfTensor(blur_dict["psf"]).to(device)`` per image."""
    device = torch.device(device)
    todo = [k for k, bd in enumerate(blur_dicts) if bd.get("blurring") and bd.get("psf") is None and bd.get("deferred_psf")]
    psfs = [None] * len(blur_dicts)
    by_shape = {}
    for k in todo:
        d = blur_dicts[k]["deferred_psf"]
        by_shape.setdefault((d["center"], d["side"], d["trajectory"] is None, d.get("trajectory_key", (0, 0))[0]), []).append(k)
    for (center, side, on_device, seed), members in by_shape.items():
        frac = [blur_dicts[k]["deferred_psf"]["fraction"] for k in members]
        if on_device:      # trajectory_backend="cuda": the walks are drawn here, one launch for the group
            traj = psf_ops.generate_trajectories(len(members), [blur_dicts[k]["deferred_psf"]["expl"] for k in members], seed, device,
                                                 indices=[blur_dicts[k]["deferred_psf"]["trajectory_key"][1] for k in members])
        else:
            traj = np.stack([blur_dicts[k]["deferred_psf"]["trajectory"] for k in members])
        dense = psf_ops.rasterize_psfs(traj, frac, device, canvas=256, center=center, out_side=side, dtype=torch.float64)
        host = dense.cpu().numpy()
        for j, k in enumerate(members):
            bd = blur_dicts[k]
            bd["psf"] = host[j]
            bd["theta_rad"], bd["scale_factor_lambda1"], bd["scale_factor_lambda2"] = psf_principal_components(host[j])
            del bd["deferred_psf"]
            psfs[k] = dense[j].to(dtype)
    # stored PSFs that a packed bank holds go up as taps only: one pinned copy for the batch, expanded on the device
    by_dir = {}
    for k, bd in enumerate(blur_dicts):
        src = bd.get("stored_psf_source")
        if psfs[k] is None and bd.get("blurring") and src is not None and np.asarray(bd["psf"]).shape == (128, 128):
            bank = psf_bank.bank_for(src[0])
            if bank.words(*src[1:]) is not None:
                by_dir.setdefault(src[0], []).append(k)
    for directory, members in by_dir.items():
        dense = psf_bank.bank_for(directory).upload([blur_dicts[k]["stored_psf_source"][1:] for k in members], device, dtype=dtype)
        for j, k in enumerate(members):
            psfs[k] = dense[j]
    for k, bd in enumerate(blur_dicts):
        if psfs[k] is None and bd.get("blurring"):
            psfs[k] = torch.as_tensor(np.asarray(bd["psf"]), dtype=dtype, device=device)
        elif psfs[k] is None:
            psfs[k] = torch.as_tensor(np.asarray(bd.get("psf", [0]), dtype=np.float32), dtype=dtype, device=device)
    return psfs


upload_psfs = complete_blur_dicts


def blur_pil_image(image, psf, device=None):
    """The ``--cpu_blur`` branch (transforms.py:349-359): PIL image in, blurred uint8 PIL image out, through the
    BlurImageHandler mirror -- same borders, contrast stretch and truncation as the reference's Fourier path."""
    from .motion_blur.blur_image import BlurImageHandler
    handler = BlurImageHandler(image_path=None, PSFs=[np.asarray(psf).astype(np.float32)], pillowImage=image, device=device)
    if not handler.blur_image():
        print("Error in blurring.")
    return handler.pilImageResult


def add_jpeg_artifact_to_image(image_GPU, jpeg_compressor, quality):
    """transforms.py:467-493: pad to a multiple of 16 (reflect), run the caller's DiffJPEG module, crop back.
    The compressor itself (models/jpeg/*) is outside the blur hot path and is used as passed in."""
    image_GPU = image_GPU.unsqueeze(0)
    w0, h0 = image_GPU.shape[3], image_GPU.shape[2]
    wp, hp = 16 - w0 % 16, 16 - h0 % 16
    left, right = math.floor(wp / 2), math.ceil(wp / 2)
    top, bottom = math.floor(hp / 2), math.ceil(hp / 2)
    padded = torch.nn.functional.pad(image_GPU, (left, right, top, bottom), mode='reflect')
    ph, pw = padded.shape[2], padded.shape[3]
    jpeg_compressor.setQuality(quality)
    jpeg_compressor.setRes(ph, pw)
    comp = jpeg_compressor(padded.float())
    image_GPU = comp[:, :, top:ph - bottom, left:pw - right].cpu()
    return image_GPU.half().detach().squeeze()
