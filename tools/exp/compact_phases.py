"""Phase timestamps of dib_compact_taps, block 0 (needs the -DDIB_COMPACT_TIMING build):
    make -C detectinblur_b200/csrc VARIANT=_ct DEFS=-DDIB_COMPACT_TIMING
    DIB_LIB_PATH=detectinblur_b200/libdib_ct.so python tools/exp/compact_phases.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, bench
import detectinblur_b200.psf_ops as ops
from detectinblur_b200 import _lib
dev = torch.device("cuda")
names = ["stage load", "sum", "pass 1 (count)", "pass 2 (write taps)", "block reductions", "masked build", "dense build", "meta"]
fn = _lib.lib.dib_debug_compact_times
fn.argtypes = [ctypes.c_void_p]
for name in ("cfg2", "cfg3"):
    spec = bench.workload_spec(name, None)
    traj, fr = bench.make_trajectories(spec, seed=0)
    psfs = ops.rasterize_psfs(traj, fr, dev, dtype=torch.float16).float()
    for first in (0, 1):
        p = psfs[first:first + 1].contiguous()
        for _ in range(3):
            ops.compact_taps(p, normalize=True, sync=False)
        buf = (ctypes.c_uint64 * 16)()
        fn(buf)
        t = list(buf)[:9]
        d = list(buf)[9:15]
        if d[0] > t[6]:
            print("   dense build: init %d, ranking passes %d, cost + choice %d, occupancy %d, band walks %d, offsets %d, records + weights %d" % (
                d[0] - t[6], d[1] - d[0], d[2] - d[1], d[3] - d[2], d[4] - d[3], d[5] - d[4], t[7] - d[5]))
        print(name, "psf", first, "total cycles", t[8] - t[0], {names[k]: t[k + 1] - t[k] for k in range(8)})
