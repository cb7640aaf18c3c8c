"""Times the Flickr-shaped Recall@K evaluation (BASELINE configs[1]: 1k images x 5k captions, d = 512) on both rank paths."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from dclip_b200 import retrieval as R
import workloads as synthetic
dev = torch.device("cuda:0")
cap, img, image_ids, cap_ids = synthetic.retrieval_inputs(1000, 5, 512, 7.0, seed=44)
cap_d, img_d = torch.from_numpy(cap).to(dev), torch.from_numpy(img).to(dev)
for path in ("fp32", "tensor"):
    for _ in range(3):
        m = R.retrieval_metrics_from_embeddings(cap_d, img_d, image_ids, cap_ids, path=path)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        m = R.retrieval_metrics_from_embeddings(cap_d, img_d, image_ids, cap_ids, path=path)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        R.retrieval_metrics_from_embeddings(cap_d, img_d, image_ids, cap_ids, path=path)
        torch.cuda.synchronize()
    rows = sorted(((e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0), key=lambda r: -r[2])
    print(f"# path={path}: {dt * 1e3:.3f} ms per eval (wall, incl. the one device->host read); t2i R@1 {m['t2i']['R@1']:.4f}; device time {sum(r[2] for r in rows):.1f} us in {sum(r[1] for r in rows)} launches")
    for k, c, t in rows[:12]:
        print(f"   {t:8.1f} us  x{c}  {k[:100]}")
