import os, random, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import detectinblur_b200.blur_functions as bf
import detectinblur_b200.psf_ops as ops
from detectinblur_b200 import _lib
from detectinblur_b200.motion_blur import Trajectory
np.random.seed(1); random.seed(1)
dev = torch.device("cuda")
PARAMS = [0.005, 0.001, 0.00005]; EXPOSURES = [1 / 25, 1 / 10, 1 / 5, 1 / 2, 1]
trajs, fracs = [], []
for p in PARAMS:
    for e in EXPOSURES:
        for _ in range(2):
            trajs.append(Trajectory(canvas=256, max_len=96, expl=p).fit().fit().x); fracs.append(e)
pool = ops.rasterize_psfs(np.stack(trajs), np.array(fracs), dev, dtype=torch.float16).float()
torch.manual_seed(0)
for shape in [(2, 70, 449), (3, 70, 449), (2, 70, 448), (2, 70, 460), (2, 200, 449)]:
    for pi in (24, 4, 14):
        for dil in (False, True):
            x = torch.rand(shape, device=dev).half()
            psf = pool[pi:pi + 1]
            if dil:
                psf = torch.nn.functional.conv2d(psf[:, None], torch.ones((1, 1, 5, 5), device=dev) / 25, padding=2)[:, 0]
            psf = psf.half().contiguous()
            ref = bf.blur_batch([x.float()], ops.compact_taps(psf.float(), normalize=True), [0], exact=True)[0]
            for planned in (False, True):
                ts = ops.compact_taps(psf, normalize=True, **({"sync": False, "max_taps": 4096} if planned else {}))
                l0 = bf.launch_count()
                got = bf.blur_batch([x], ts, [0])[0]
                err = float((got.float() - ref).abs().max())
                m = ops.compact_taps(psf, normalize=True).meta[0]
                flag = "" if err < 2e-3 else "   <<<<<<"
                print(shape, "psf", pi, "dilated" if dil else "plain", "planned" if planned else "host", "launches", bf.launch_count() - l0,
                      "taps", m.count, "chunks", m.prog_chunks, "G", m.prog_group_w, "err %.3g" % err, flag)
