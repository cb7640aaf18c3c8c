#!/usr/bin/env python
"""bench.py — DCLIP hot-path benchmark on B200 (contract in the task brief).

Headline workload (BASELINE.json configs[3], "ViT-L/14 distillation contrastive+KL loss, global
batch 4096, d=768"): one step = distill-loss forward+backward (symmetric InfoNCE + cosine feature
matching) over the global batch; metric = pairs/s.  N > 1 shards the global
batch over the ranks (strong scaling: the global batch is fixed at 4096) with the operand
all-gather + lse all-gather choreography of dclip_b200.losses.sharded_distill_loss_fwd_bwd.

The second BASELINE metric (retrieval queries/s, top-10) is reported in the same JSON line under
"retrieval" (Flickr-shaped eval, configs[1], and an exact top-10 search over a gallery slice).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

B_GLOBAL = 4096
DIM = 768
TEMPERATURE = 0.05
METRIC = "distill-loss fwd+bwd pairs/sec"
UNIT = "pairs/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--path", default=os.environ.get("DCLIP_BENCH_PATH", "auto"), choices=["auto", "fp32", "tensor"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32", "fp16"])
    ap.add_argument("--batch", type=int, default=B_GLOBAL)
    ap.add_argument("--dim", type=int, default=DIM)
    ap.add_argument("--no-retrieval", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel/collective from Python instead of one CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gallery", type=int, default=10_000_000, help="total gallery rows of the sharded top-10 workload (configs[4])")
    ap.add_argument("--queries", type=int, default=100_000)
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index=0, period=0.002):      # the timed loop lasts only tens of ms: sample every 2 ms
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        # median of the upper half = clock under load (idle samples between steps pull the plain median down)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_reference_step(s_img, s_txt, t_img, t_txt):
    from oracle import reference_math as om
    return om.distill_total_with_grads(s_img, s_txt, t_img, t_txt, TEMPERATURE)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the reference is
    pure Python/PyTorch and cannot travel to the GPU box; the port is pinned to it by tests/golden)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B, d = args.batch, args.dim
    s_img, s_txt, t_img, t_txt = synthetic.loss_inputs(B, d, seed=42)
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_reference_step(s_img, s_txt, t_img, t_txt)
    steps = max(1, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(s_img, s_txt, t_img, t_txt)
    dt = (time.perf_counter() - t0) / steps
    val = B / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"distill-loss fwd+bwd, global batch {B}, d={d}, tau={TEMPERATURE} (BASELINE configs[3])",
                   "global_batch": B, "dim": d},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps} full steps of the B={B} d={d} workload, torch fp32 on {cores} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from dclip_b200 import _lib
    from dclip_b200 import losses as L
    from oracle import synthetic          # synthetic input recipe only (not the checker, not timed)

    lib = _lib.load_library()
    tdtype = {"bf16": torch.bfloat16, "fp32": torch.float32, "fp16": torch.float16}[args.dtype]
    B, d = args.batch, args.dim
    assert B % world == 0
    m = B // world
    path = args.path
    if path == "auto":
        path = "tensor" if (lib.dclip_device_supports_tensor_path(local_rank) and tdtype != torch.float32
                            and os.environ.get("DCLIP_DISABLE_TENSOR", "0") != "1"
                            and lib.dclip_tensor_path_ready()) else "fp32"

    host = [t.to(tdtype)[rank * m:(rank + 1) * m].contiguous().pin_memory() for t in synthetic.loss_inputs(B, d, seed=42)]
    resident = [t.to(dev) for t in host]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2
    launches_per_step = lib.dclip_distill_loss_launch_count(m, B, d, _lib.path_code(path), 1 if world > 1 else 0)

    def step_fn(s_img, s_txt, t_img, t_txt):
        if world == 1:
            return L.distill_loss_fwd_bwd(s_img, s_txt, t_img, t_txt, temperature=TEMPERATURE, path=path)
        return L.sharded_distill_loss_fwd_bwd(s_img, s_txt, t_img, t_txt, temperature=TEMPERATURE, path=path)

    graphed = None
    if not args.no_graph:
        # the public API's CUDA-graph wrapper: one graph launch per step (kernels + NCCL collectives)
        graphed = L.GraphedStep(step_fn, resident)
        step_inputs = graphed.static_in           # the resident inputs ARE the graph's static inputs
    else:
        step_inputs = resident

    def step_resident():
        if graphed is not None:
            graphed.graph.replay()
            return graphed.static_out
        return step_fn(*step_inputs)

    host_step = L.HostDistillStep(B, d, tdtype, dev, temperature=TEMPERATURE, path=path) if world == 1 else None

    def step_e2e():
        if host_step is not None:
            # public host-buffer API: pinned host -> device copies (students first), the contrastive part overlapped
            # with the teachers' copy, device -> host read of the four losses; all inside the timed region
            return host_step(*host)[0]
        for dst, src in zip(step_inputs, host):
            dst.copy_(src, non_blocking=True)       # pinned host -> device, inside the timed region
        losses = step_resident()[0]
        return losses.cpu()          # device->host read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N > 1: after the host barrier the ranks still enqueue their step tens of microseconds apart (host jitter),
    # and max-over-ranks timing of an isolated 0.15 ms step would mostly measure that.  A device-side gate (one
    # peer-memory signal + wait, dclip_p2p_push) right before the start event releases all GPUs together.
    gate = None
    if world > 1:
        try:
            from dclip_b200.peer import PeerBuffer
            gate = PeerBuffer(256, dev)
        except RuntimeError:
            gate = None

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        total_ms = 0.0
        for _ in range(steps):
            flush.zero_()                                    # L2 flush between timed iterations (untimed)
            barrier()
            if gate is not None:
                gate.push([], -1, 3, wait_after=True)        # device-side start gate (untimed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)     # max over ranks, per step
            total_ms += float(t)
        barrier()
        return total_ms / steps

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(step_resident, args.steps, max(3, args.warmup))
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(step_e2e, max(5, args.steps // 2), 3)

    # dominant kernel (backward grad contraction) timed alone through the C-ABI building block
    roof = None
    if rank == 0:
        roof = kernel_roofline(L, _lib, resident, B, m, d, path, rank, load_peaks())

    extra = {}
    if not args.no_retrieval:
        big = sharded_gallery_bench(dev, rank, world, args.gallery, args.queries)     # every rank takes part
        if rank == 0:
            extra["retrieval"] = {"sharded_top10": big}
            if world == 1:
                extra["retrieval"].update(retrieval_bench(dev))
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(B, d)

    if rank == 0:
        line = {
            "metric": METRIC, "value": B / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": {"bf16": "bf16", "fp32": "f32", "fp16": "f16"}[args.dtype] +
            ("/f16-operands,f32-accumulate" if path == "tensor" else "/f32-math"),
            "data": "synthetic",
            "config": {"workload": f"distill-loss fwd+bwd, global batch {B}, d={d}, tau={TEMPERATURE} (BASELINE configs[3])",
                       "global_batch": B, "dim": d, "rows_per_gpu": m, "path": path, "l2_flush": "256 MiB write between timed steps",
                       "launch": "python" if args.no_graph else "cuda-graph (dclip_b200.losses.GraphedStep)",
                       "parallelism": f"dp{world} (batch-sharded, operand all-gather over NVLink peer memory)" if world > 1 else "single GPU",
                       "start": "host barrier + device-side peer-memory gate before the start event" if gate is not None else "host barrier"},
            "clocks": clocks,
            "e2e": {"value": B / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host) * world,
                    "d2h_bytes_per_step": 16 * world},
            "gpu_launches": int(launches_per_step) * args.steps,
            "roofline": roof, "cpu_baseline": cpu,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        # tearing NCCL down while captured graphs still reference its streams can block: drop the graph,
        # drain the device, leave through a barrier and exit without the (optional) communicator destroy.
        graphed = None
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


# DRAM traffic of the dominant kernel per launch, from the committed ncu --set full capture of this very
# launch shape (profiles/r1_ncu_full_loss_kernels.txt: dram__bytes_read.sum + dram__bytes_write.sum); the
# operands (2 x 6.3 MB fp16) are read once, everything else lives in L2.
NCU_DRAM_BYTES_PER_LAUNCH = {"tensor": 12.69e6}


def kernel_roofline(L, _lib, resident, B, m, d, path, rank, peaks):
    """Times the dominant kernel alone — the InfoNCE backward contraction, both directions in one launch
    exactly as the step launches it — with CUDA events on the launching stream, and relates its ALGORITHMIC
    FLOPs (dU = G V and dV = G^T U: 2 x 2*B*B*d; the score-tile recompute is not counted) to the measured
    bf16 peak."""
    if m != B:   # the roofline entry is taken at N = 1
        return None
    from dclip_b200._lib import check, ptr, stream_ptr, workspace
    lib = _lib.load_library()
    ops = L.CudaLossOps(path)
    dev = resident[0].device
    u_all = torch.empty((B, d), dtype=ops.op_torch_dtype, device=dev)
    v_all = torch.empty((B, d), dtype=ops.op_torch_dtype, device=dev)
    ops.prep(resident[0], None, u_all)
    ops.prep(resident[1], None, v_all)
    lse_u, _ = ops.rowlse(u_all, v_all, TEMPERATURE, 0, True)
    lse_v, _ = ops.rowlse(v_all, u_all, TEMPERATURE, 0, False)
    du = torch.empty((B, d), dtype=torch.float32, device=dev)
    dv = torch.empty((B, d), dtype=torch.float32, device=dev)
    pc = _lib.path_code(path)
    ws = workspace(lib.dclip_infonce_grad_pair_workspace_bytes(B, d, pc), dev)

    def launch():
        check(lib.dclip_infonce_grad_pair(ptr(u_all), ptr(v_all), B, d, ops.op_dtype, TEMPERATURE, ptr(lse_u), ptr(lse_v),
                                          1.0 / (2 * B), ptr(du), ptr(dv), ptr(ws), ws.numel(), pc, stream_ptr(dev)),
              "infonce_grad_pair")

    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        launch()
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / n
    flops = 4.0 * B * B * d
    achieved = flops / sec / 1e12
    peak = peaks["bf16_tflops"]
    return {"bound": "tensor", "kernel": "k_infonce_bwd_pair (CTA pairs, cta_group::2; both directions in one launch: dU = G V, dV = G^T U; score tiles recomputed)",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_source": peaks["source"] + " bf16 burst (kernel timed alone, back to back, L2-warm)",
            "algorithmic_flops_per_launch": flops, "executed_flops_per_launch": 2.0 * flops,
            "tensor_cores": path == "tensor", "avg_launch_us": sec * 1e6, "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(path)}


def sharded_gallery_bench(dev, rank, world, n_gallery, n_query):
    """BASELINE configs[4]: cosine top-10 over an N x 768 fp16 gallery sharded over the ranks (rank r holds
    rows [r N/P, (r+1) N/P)), every rank scores all queries against its shard on the tcgen05 path, the
    per-GPU exact top-10 lists are all-gathered and merged by (-score, +index).  Strong scaling."""
    import math

    import torch.distributed as dist

    from dclip_b200 import retrieval as R
    d = 768
    per = n_gallery // world
    lo = rank * per
    hi = n_gallery if rank == world - 1 else lo + per
    g = torch.empty((hi - lo, d), dtype=torch.float16, device=dev)
    gen = torch.Generator(device=dev).manual_seed(42 + rank)
    for a0 in range(0, hi - lo, 1 << 20):          # generated on the device, shard by shard
        b0 = min(hi - lo, a0 + (1 << 20))
        x = torch.randn(b0 - a0, d, device=dev, generator=gen)
        g[a0:b0] = (x / x.norm(dim=1, keepdim=True)).half()
    # queries = perturbed gallery rows of rank 0's shard (known nearest neighbour: built-in R@1 check)
    qgen = torch.Generator(device=dev).manual_seed(7)
    planted = torch.randperm(min(per, 1 << 20), device=dev, generator=qgen)[:n_query]
    if rank == 0:
        q = g[planted].float() + 0.1 / math.sqrt(d) * torch.randn(n_query, d, device=dev, generator=qgen)
        q = (q / q.norm(dim=1, keepdim=True)).half()
    else:
        q = torch.empty((n_query, d), dtype=torch.float16, device=dev)
    if world > 1:
        dist.broadcast(q, src=0)
    gallery = R.TensorGallery(g)

    def search():
        s, i = R.topk_ip(q, g, 10, g_index_offset=lo, path="tensor", gallery=gallery)
        if world > 1:
            all_s = torch.empty((world * n_query, 10), dtype=s.dtype, device=dev)
            all_i = torch.empty((world * n_query, 10), dtype=i.dtype, device=dev)
            dist.all_gather_into_tensor(all_s, s)
            dist.all_gather_into_tensor(all_i, i)
            s, i = R.topk_merge(all_s.view(world, n_query, 10), all_i.view(world, n_query, 10))
        return s, i

    search()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s, i = search()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t) * 1e-3
    r1 = float((i[:, 0] == planted).float().mean())
    del g, gallery
    torch.cuda.empty_cache()
    peaks = load_peaks()
    tflops = 2.0 * n_query * n_gallery * d / sec / 1e12
    return {"metric": "retrieval queries/sec top-10", "queries_per_s": n_query / sec, "ms": sec * 1e3, "planted_R@1": r1,
            "tflops_aggregate": tflops, "frac_of_bf16_sustained_peak": tflops / (world * peaks["bf16_tflops_sustained"]),
            "workload": f"exact cosine top-10, {n_gallery} x {d} fp16 gallery sharded over {world} GPU(s), {n_query} queries "
                        "(BASELINE configs[4]); tcgen05 candidates + exact fp32 re-scoring, bit-exact contract"}


def retrieval_bench(dev):
    """Second BASELINE metric: retrieval queries/s.  (a) Flickr-shaped Recall@K eval, 1k images x 5k
    captions, d=512 (configs[1]; 6000 ranked queries per eval).  (b) exact top-10 over a 1M x 768
    fp16 gallery slice, 1024 queries (configs[4] scaled to one GPU's slice of the bench budget)."""
    import numpy as np

    from dclip_b200 import retrieval as R
    from oracle import synthetic
    out = {}
    cap, img, image_ids, cap_ids = synthetic.retrieval_inputs(1000, 5, 512, 1.0, seed=42)
    cap_d, img_d = torch.from_numpy(cap).to(dev), torch.from_numpy(img).to(dev)
    for _ in range(2):
        R.retrieval_metrics_from_embeddings(cap_d, img_d, image_ids, cap_ids)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        m = R.retrieval_metrics_from_embeddings(cap_d, img_d, image_ids, cap_ids)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    out["flickr_eval"] = {"queries_per_s": 6000 / dt, "ms_per_eval": dt * 1e3, "t2i_R@1": m["t2i"]["R@1"],
                          "i2t_R@1": m["i2t"]["R@1"], "workload": "1k images x 5k captions, d=512, R@1/5/10 both directions"}
    ng, nq, dd = 1_000_000, 4096, 768
    g = torch.randn(ng, dd, device=dev, generator=torch.Generator(device=dev).manual_seed(42))
    g = (g / g.norm(dim=1, keepdim=True)).half()
    idx = torch.randperm(ng, device=dev)[:nq]
    q = g[idx].float() + 0.1 / np.sqrt(dd) * torch.randn(nq, dd, device=dev)
    q = (q / q.norm(dim=1, keepdim=True)).half()
    gallery = R.TensorGallery(g)
    for name, path in (("top10", "auto"), ("top10_fp32_path", "fp32")):
        s, i = R.topk_ip(q, g, 10, path=path, gallery=gallery)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s, i = R.topk_ip(q, g, 10, path=path, gallery=gallery)
        e1.record()
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) * 1e-3
        out[name] = {"queries_per_s": nq / sec, "ms": sec * 1e3, "planted_R@1": float((i[:, 0] == idx).float().mean()),
                     "tflops": 2.0 * nq * ng * dd / sec / 1e12, "path": path,
                     "workload": f"exact IP top-10, {ng} x {dd} fp16 gallery, {nq} queries, 1 GPU"}
    return out


def cpu_baseline(B, d):
    from oracle import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    s_img, s_txt, t_img, t_txt = synthetic.loss_inputs(B, d, seed=42)
    cpu_reference_step(s_img, s_txt, t_img, t_txt)
    n = 0
    t0 = time.perf_counter()
    while True:
        cpu_reference_step(s_img, s_txt, t_img, t_txt)
        n += 1
        el = time.perf_counter() - t0
        if el > 12.0 or n >= 40:
            break
    return {"value": B * n / el, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} full steps (B={B}, d={d}) of the oracle port (torch fp32 CPU, {cores} threads), {el:.1f} s"}


if __name__ == "__main__":
    main()
