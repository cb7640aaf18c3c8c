import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dclip_b200 import losses as L
import workloads as synthetic
dev = torch.device("cuda:0")
for B, d in ((1024, 768), (4096, 256), (640, 512), (4096, 768)):
    x = [t.to(torch.bfloat16).to(dev) for t in synthetic.loss_inputs(B, d, seed=7)]
    ref = None
    bad = [0, 0, 0]
    for it in range(25):
        out = L.distill_loss_fwd_bwd(*x, path="tensor")
        cur = [o.clone() for o in out]
        if ref is None:
            ref = cur
        else:
            for k in range(3):
                if not torch.equal(ref[k], cur[k]):
                    bad[k] += 1
                    if bad[k] == 1:
                        diff = (ref[k].float() - cur[k].float()).abs()
                        nz = diff.nonzero()
                        print(f"  B={B} d={d} out{k} it={it}: {nz.shape[0]} elements differ, max {float(diff.max()):.3e}; first at {nz[0].tolist()}, rows {sorted(set(nz[:, 0].tolist()))[:8] if nz.dim() == 2 else ''}")
    print(f"B={B} d={d} fused_tail={os.environ.get('DCLIP_GMAT_FUSED_TAIL', '1')}: mismatching runs (losses, d_img, d_txt) = {bad}", flush=True)
