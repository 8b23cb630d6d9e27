"""A small end-to-end case for compute-sanitizer runs (not collected by pytest):

    compute-sanitizer --tool memcheck python tests/sanitize_case.py

Lives under tests/ because it draws its PSFs with the oracle (test infrastructure)."""
import torch, numpy as np, sys
sys.path.insert(0, ".")
import detectinblur_b200.blur_functions as bf, detectinblur_b200.psf_ops as ops
from oracle import psf_oracle as po
np.random.seed(3)
psfs=[]
for frac,expl in ((1/10,0.005),(1,0.00005)):
    p16,_=po.stored_psf(expl,frac,np.random); psfs.append(po.crop128(p16).astype(np.float32))
ts=ops.compact_taps(torch.from_numpy(np.stack(psfs)).cuda(), normalize=True)
imgs=[torch.rand(3,150,500).cuda(), torch.rand(2,100,230).cuda()]
out=bf.blur_batch(imgs, ts, [0,1])
ex=bf.blur_batch(imgs, ts, [0,1], exact=True)
torch.cuda.synchronize()
print("max err", max((a-b).abs().max().item() for a,b in zip(out,ex)), ts.counts)
x=ops.rasterize_psfs(np.stack([po.trajectory(256,2000,96,0.005,np.random)]), [0.2], "cuda")
print(x.sum().item())
