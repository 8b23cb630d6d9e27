"""Kernel-development harness: times the tiled blur kernel alone on BASELINE configs 2 and 3 (device-resident inputs,
rotating batches) and prints the tap programs' shape.  DIB_LIB_PATH selects an alternative build of libdib.so.

    python tools/exp/kbench.py [cfg2|cfg3|cfg2b32 ...] [--steps N]
"""
import os
import sys
import json

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import detectinblur_b200.blur_functions as bf
import detectinblur_b200.psf_ops as ops


def run(name, steps):
    batch = None
    wl = name
    if "b" in name[3:]:
        wl, batch = name.split("b")[0], int(name.split("b")[1])
    spec = bench.workload_spec(wl, batch)
    B = spec["batch"]
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device="cpu").manual_seed(1337)
    n_rot = 3
    batches = [torch.rand((B, 3, 800, 1333), generator=gen).to(dev) for _ in range(n_rot)]
    outs = [torch.zeros((B, 3, 800, 1336), device=dev) for _ in range(n_rot)]
    traj, fr = bench.make_trajectories(spec, seed=0)
    psfs = ops.rasterize_psfs(traj, fr, dev, dtype=torch.float16).float()
    ts = ops.compact_taps(psfs, normalize=True)
    prog = [(m.count, m.prog_group_w, m.prog_shear, m.prog_steps, m.prog_segs, m.prog_chunks) for m in ts.meta]
    slots = sum(p[1] * p[3] for p in prog)
    plans = [bf.prepare_blur([batches[r][i] for i in range(B)], ts, list(range(B)), outs=[outs[r][i, :, :, :1333] for i in range(B)])
             for r in range(n_rot)]
    res = {}
    for overlap in (False, True):
        for k in range(5):
            plans[k % n_rot].run(overlap=overlap)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(steps):
            plans[k % n_rot].run(overlap=overlap)
        e1.record()
        torch.cuda.synchronize()
        res["overlap" if overlap else "ordered"] = 1000.0 * e0.elapsed_time(e1) / steps
    taps = sum(p[0] for p in prog)
    fma_floor_us = 2.0 * taps * 3 * 800 * 1333 / 71.3e12 * 1e6
    hbm_floor_us = B * 2 * 3 * 800 * 1333 * 4 / 6543.1e9 * 1e6
    floor = max(fma_floor_us, hbm_floor_us)
    # check one output against the exact kernel on a crop
    chk = bf.blur_batch([batches[0][0][:, :200, :300].contiguous()], ts, [0], exact=True)[0]
    got = bf.blur_batch([batches[0][0][:, :200, :300].contiguous()], ts, [0])[0]
    err = (chk - got).abs().max().item()
    print(json.dumps({"workload": name, "lib": os.environ.get("DIB_LIB_PATH", "default"), "us_ordered": round(res["ordered"], 2),
                      "us_overlap": round(res["overlap"], 2), "floor_us": round(floor, 2), "frac_ordered": round(floor / res["ordered"], 3),
                      "frac_overlap": round(floor / res["overlap"], 3), "fill_eff": round(taps / max(slots, 1), 3), "err": err,
                      "prog(count,G,k,steps,segs,chunks)": prog}))


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["cfg2", "cfg3"]
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 200
    for n in names:
        run(n, steps if n.startswith("cfg2") else max(20, steps // 5))
