"""Which tiled kernel should take which PSF?  Times one blur of 8 x 3x800x1333 fp32 per (param, exposure) cell with the
masked kernel's program (DIB_MASKED_ONLY=1), the dense kernel's (DIB_DENSE_ONLY=1) and the default routing."""
import os, sys, json, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

def child():
    import numpy as np, random, torch
    import detectinblur_b200.blur_functions as bf
    import detectinblur_b200.psf_ops as ops
    from detectinblur_b200.motion_blur.generate_trajectory import Trajectory
    PARAMS = [0.005, 0.001, 0.00005]
    EXPOSURES = [1 / 25, 1 / 10, 1 / 5, 1 / 2, 1]
    np.random.seed(7); random.seed(7)
    dev = torch.device("cuda")
    B = 8
    imgs = torch.rand((B, 3, 800, 1333), generator=torch.Generator().manual_seed(1)).to(dev)
    outs = torch.zeros((B, 3, 800, 1336), device=dev)
    res = {}
    for p in PARAMS:
        for e in EXPOSURES:
            traj = np.stack([Trajectory(canvas=256, max_len=96, expl=p).fit().fit().x for _ in range(B)])
            psfs = ops.rasterize_psfs(traj, [e] * B, dev, dtype=torch.float16).float()
            ts = ops.compact_taps(psfs, normalize=True)
            plan = bf.prepare_blur([imgs[i] for i in range(B)], ts, list(range(B)), outs=[outs[i, :, :, :1333] for i in range(B)])
            for _ in range(3):
                plan.run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                plan.run()
            e1.record(); torch.cuda.synchronize()
            kinds = sorted({m.prog_group_w for m in ts.meta})
            res["P%gE%g" % (p, e)] = [round(e0.elapsed_time(e1) / 10 * 1e3, 1), int(np.mean(ts.counts)), kinds, int(np.mean([m.prog_chunks for m in ts.meta]))]
    print(json.dumps(res))

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        out = {}
        for name, env in (("masked", {"DIB_MASKED_ONLY": "1"}), ("dense", {"DIB_DENSE_ONLY": "1"}), ("default", {})):
            r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **env), capture_output=True, text=True)
            out[name] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else r.stderr[-500:]
        cells = list(out["default"].keys()) if isinstance(out["default"], dict) else []
        for c in cells:
            print(c, "taps", out["default"][c][1], "| masked us", out["masked"][c][0], "chunks", out["masked"][c][3], "| dense us", out["dense"][c][0],
                  "chunks", out["dense"][c][3], "| default", out["default"][c][0], out["default"][c][2])
