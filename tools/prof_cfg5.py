"""BASELINE configs[4] on one GPU (10 M x 768 fp16 gallery, 100 k queries, top-10): phase timing with CUDA events."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda:0")
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
r = bench.sharded_gallery_bench(dev, 0, 1, ng, nq)
print({k: r[k] for k in ("queries_per_s", "ms", "tflops_aggregate", "frac_of_bf16_sustained_peak", "planted_R@1")})
