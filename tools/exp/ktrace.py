"""Timeline of CTA 0 of the tiled kernel (needs the DIB_TRACE build: make VARIANT=_trace DEFS=-DDIB_TRACE).
    DIB_LIB_PATH=detectinblur_b200/libdib_trace.so python tools/exp/ktrace.py cfg3 > gpurun_out/trace.json"""
import os, sys, json, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch, bench
import detectinblur_b200.blur_functions as bf
import detectinblur_b200.psf_ops as ops
from detectinblur_b200 import _lib

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
spec = bench.workload_spec(name, None)
B = spec["batch"]
dev = torch.device("cuda", 0)
gen = torch.Generator(device="cpu").manual_seed(1337)
dt = torch.float16 if spec.get("dtype") == "f16" or name.endswith("h") else torch.float32
imgs = torch.rand((B, 3, 800, 1333), generator=gen).to(dev).to(dt)
outs = torch.zeros((B, 3, 800, 1336), device=dev, dtype=dt)
traj, fr = bench.make_trajectories(spec, seed=0)
psfs = ops.rasterize_psfs(traj, fr, dev, dtype=torch.float16).float()
ts = ops.compact_taps(psfs, normalize=True)
plan = bf.prepare_blur([imgs[i] for i in range(B)], ts, list(range(B)), outs=[outs[i, :, :, :1333] for i in range(B)])
lib = _lib.lib
fn = lib.dib_debug_trace_masked if ts.meta[0].prog_group_w == 0 else lib.dib_debug_trace
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
fn.restype = ctypes.c_int
buf = (ctypes.c_uint64 * 65536)()
for _ in range(3):
    plan.run()
torch.cuda.synchronize()
fn(buf, 65536)          # reset
plan.run()
torch.cuda.synchronize()
n = fn(buf, 65536)
ev = np.frombuffer(buf, dtype=np.uint64)[:n]
rec = [(int(e >> 24), int((e >> 6) & 0x3f), int(e & 0x3f), int((e >> 12) & 0xfff)) for e in ev]
rec.sort()
t0 = rec[0][0]
print(json.dumps({"workload": name, "n": n, "events": [(t - t0, w, e, a) for t, w, e, a in rec]}))
