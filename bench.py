#!/usr/bin/env python
"""bench.py — DCLIP hot-path benchmark on B200 (contract in the task brief).

Headline workload (BASELINE.json configs[3], "ViT-L/14 distillation contrastive+KL loss, global batch 4096,
d=768"): one step = distill-loss forward+backward (symmetric InfoNCE + cosine feature matching, SURVEY.md §8a
a9-a11) over the global batch; metric = pairs/s.  N > 1 shards the global batch over the ranks (strong scaling:
the global batch stays 4096) with the peer-memory operand / statistics exchange of
dclip_b200.losses.sharded_distill_loss_fwd_bwd; every rank re-computes the objective on ONE GPU and the line
carries `parity_check` (sharded == single-GPU losses and gradients).

Extra keys of the same JSON line (they explain the headline, the driver reads the top-level fields):
  roofline / cpu_baseline      dominant kernel vs the measured bf16 peak; the reference's own code on the host cores
  retrieval.sharded_top10      BASELINE configs[4]: 10 M x 768 fp16 gallery sharded over the ranks, 100 k queries,
                               + 256 sampled queries re-run on the exact FP32 path (bit-identical or the run fails)
  retrieval.flickr_eval ...    configs[1] (N = 1) with the reference's calculate_retrieval_metrics timed beside it
  teacher_step.cfg0 / .cfg2    BASELINE configs[0] / [2]: meta-teacher cross-attention + aggregation + InfoNCE fwd+bwd
                               (one CUDA graph), us/step, roofline fraction, reference module on the host cores
  weak_scaling                 N > 1: 4096 rows PER GPU (the shape on which the NVLink exchange can reach 85 %)

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
    ap.add_argument("--no-teacher", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel/collective from Python instead of one CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gallery", type=int, default=10_000_000, help="total gallery rows of the sharded top-10 workload (configs[4])")
    ap.add_argument("--queries", type=int, default=100_000)
    return ap.parse_args()


def shared_config(B, d, world):
    """The `config` object: identical in the GPU arm and in the reference arm of the same invocation."""
    return {"workload": f"distill-loss fwd+bwd, global batch {B}, d={d}, tau={TEMPERATURE} (BASELINE configs[3])",
            "global_batch": B, "dim": d, "temperature": TEMPERATURE, "rows_per_gpu": B // max(world, 1),
            "l2_flush": "between timed steps: 256 MiB write, then a 256 MiB read of another buffer, so L2 is cold AND clean (no dirty "
                        "flush lines to write back inside the timed step) - GPU arm; none needed on the host arm"}


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
        # the steps are short bursts between untimed L2 flushes and barriers: the clock UNDER LOAD is the median of
        # the upper half of the samples; the plain median is reported beside it
        upper = s[len(s) // 2:]
        return {"sm_mhz": upper[len(upper) // 2] if upper else None, "sm_mhz_median_all": s[len(s) // 2] if s else None,
                "sm_mhz_min": s[0] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------------------------
# reference arm (host cores): the reference's own code where it is staged (oracle/_ref), else the oracle port
# ------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import workloads as synthetic
    from oracle import reference_arm
    arm = reference_arm.load()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B, d = args.batch, args.dim
    inputs = synthetic.loss_inputs(B, d, seed=42)
    for _ in range(max(1, min(args.warmup, 2))):
        arm.loss_step(*inputs, TEMPERATURE)
    steps = max(1, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(steps):
        arm.loss_step(*inputs, TEMPERATURE)
    dt = (time.perf_counter() - t0) / steps
    val = B / dt
    what = ("the reference's cosine_distillation_loss x2 + compute_contrastive_loss + backward (training/CLIP_image_distillation.py:532-576, "
            "staged unmodified in oracle/_ref)") if arm.kind == "reference" else "the oracle port (oracle/reference_math.py; no staged reference found)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": shared_config(B, d, args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": arm.kind,
                         "sample": f"{steps} full steps of the B={B} d={d} workload: {what}; torch fp32 on {cores} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    # stdout carries exactly ONE JSON line: libraries that print to file descriptor 1 (NCCL's version banner) go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from dclip_b200 import _lib
    from dclip_b200 import losses as L
    import workloads as synthetic          # seeded input recipes (neutral module: the GPU arm never imports oracle/)

    lib = _lib.load_library()
    tdtype = {"bf16": torch.bfloat16, "fp32": torch.float32, "fp16": torch.float16}[args.dtype]
    B, d = args.batch, args.dim
    assert B % world == 0
    m = B // world
    full_cpu = [t.to(tdtype) for t in synthetic.loss_inputs(B, d, seed=42)]
    path = args.path
    if path == "auto":       # the library's own rule (dclip_resolve_loss_path): tcgen05 path for 16-bit inputs
        probe = torch.empty((1, d), dtype=tdtype, device=dev)
        path = "tensor" if _lib.resolve_loss_path("auto", probe, TEMPERATURE, global_batch=B) == _lib.PATH_TENSOR else "fp32"

    host = [t[rank * m:(rank + 1) * m].contiguous().pin_memory() for t in full_cpu]
    resident = [t.to(dev) for t in host]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2
    flush_rd = torch.zeros(32 * 1024 * 1024, dtype=torch.int64, device=dev)   # 256 MiB, only ever read
    def launches_per_step_now():
        # N > 1: 2 = the peer exchange runs inside the kernels (a peer buffer was mapped, 16-bit inputs), 1 = separate push kernels / NCCL
        sharded = 0
        if world > 1:
            in_kernel = (any(b is not None for b, _ in L._PEER_BUFFERS.values()) and args.dtype != "fp32"
                         and os.environ.get("DCLIP_SHARDED_PUSH", "0") != "1")
            sharded = 2 if in_kernel else 1
        return lib.dclip_distill_loss_launch_count(m, B, d, _lib.path_code(path), sharded)

    def step_fn(s_img, s_txt, t_img, t_txt):
        if world == 1:
            return L.distill_loss_fwd_bwd(s_img, s_txt, t_img, t_txt, temperature=TEMPERATURE, path=path)
        return L.sharded_distill_loss_fwd_bwd(s_img, s_txt, t_img, t_txt, temperature=TEMPERATURE, path=path)

    # ---- parity of the sharded step against the single-GPU step (N > 1): same seed -> the N = 1 values are constants ----
    parity = None
    if world > 1:
        full_dev = [t.to(dev) for t in full_cpu]
        ref_losses, ref_di, ref_dt = L.distill_loss_fwd_bwd(*full_dev, temperature=TEMPERATURE, path=path)
        got_losses, got_di, got_dt = step_fn(*resident)
        torch.cuda.synchronize()
        sl = slice(rank * m, (rank + 1) * m)
        loss_err = float(((got_losses - ref_losses).abs() / ref_losses.abs().clamp_min(1e-6)).max())
        g_err = max(float((got_di.float() - ref_di[sl].float()).norm() / ref_di[sl].float().norm()),
                    float((got_dt.float() - ref_dt[sl].float()).norm() / ref_dt[sl].float().norm()))
        ok = loss_err < 1e-5 and g_err < (2e-3 if path == "tensor" else 2e-5)
        flag = torch.tensor([1 if ok else 0, ], device=dev)
        errs = torch.tensor([loss_err, g_err], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        parity = {"status": "ok" if int(flag) == 1 else "FAILED", "max_rel_loss_diff_vs_1gpu": float(errs[0]),
                  "max_rel_grad_diff_vs_1gpu": float(errs[1]), "losses_sharded": [float(x) for x in got_losses.cpu()],
                  "losses_1gpu": [float(x) for x in ref_losses.cpu()]}
        del full_dev, ref_di, ref_dt

    graphed = None
    if not args.no_graph:
        # the public API's CUDA-graph wrapper: one graph launch per step (kernels + peer-memory exchanges)
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
            # with the teachers' copies, the four losses delivered into pinned host memory (returned as a CPU tensor);
            # all inside the timed region
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
            flush.zero_()                                    # L2 flush between timed iterations (untimed) ...
            flush_rd.sum()                                   # ... then fill L2 with clean lines of a read-only buffer
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
        if roof is not None:      # the whole step against the same peak: forward 2*B*B*d + backward 4*B*B*d, nothing else counted
            step_tf = 6.0 * B * B * d / (ms * 1e-3) / 1e12
            roof["step"] = {"algorithmic_flops": 6.0 * B * B * d, "tflops": step_tf, "frac": step_tf / roof["peak"], "us": ms * 1e3}

    extra = {}
    if world > 1:
        weak = weak_scaling_bench(L, dev, rank, world, d, tdtype, path, timed, args)
        if rank == 0:
            extra["weak_scaling"] = weak
    if not args.no_retrieval:
        big = sharded_gallery_bench(dev, rank, world, args.gallery, args.queries)     # every rank takes part
        if rank == 0:
            extra["retrieval"] = {"sharded_top10": big}
            if world == 1:
                extra["retrieval"].update(retrieval_bench(dev, not args.no_cpu_baseline))
    if rank == 0 and world == 1 and not args.no_teacher:
        extra["teacher_step"] = teacher_step_bench(dev, not args.no_cpu_baseline)
    if rank == 0 and world == 1 and args.dtype == "bf16" and not args.no_teacher:
        extra["fp32_inputs"] = fp32_input_bench(L, dev, B, d)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(B, d)

    failed = (parity is not None and parity["status"] != "ok") or (
        "retrieval" in extra and extra["retrieval"]["sharded_top10"].get("fp32_recheck", {}).get("status") == "FAILED")
    if rank == 0:
        line = {
            "metric": METRIC, "value": B / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": {"bf16": "bf16", "fp32": "f32", "fp16": "f16"}[args.dtype] +
            ("/f16-operands,f32-accumulate" if path == "tensor" else "/f32-math"),
            "data": "synthetic",
            "config": shared_config(B, d, world),
            "details": {"path": path, "launch": "python" if args.no_graph else "cuda-graph (dclip_b200.losses.GraphedStep)",
                        "parallelism": f"dp{world} (batch-sharded, operand exchange over NVLink peer memory)" if world > 1 else "single GPU",
                        "start": "host barrier + device-side peer-memory gate before the start event" if gate is not None else "host barrier"},
            "clocks": clocks,
            "e2e": {"value": B / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host) * world,
                    # N = 1: the four losses + a sequence word, stored by the last block of the step's last kernel into the pinned
                    # host buffer the call waits on (HostDistillStep); N > 1: a 16-byte D2H copy of the losses
                    "d2h_bytes_per_step": 20 if world == 1 else 16 * world},
            "gpu_launches": int(launches_per_step_now()) * args.steps,
            "roofline": roof, "cpu_baseline": cpu,
        }
        if parity is not None:
            line["parity_check"] = parity["status"]
            line["parity"] = parity
        line.update(extra)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        # tearing NCCL down while captured graphs still reference its streams can block: drop the graph,
        # drain the device, leave through a barrier and exit without the (optional) communicator destroy.
        graphed = None
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(1 if failed else 0)
    if failed:
        raise SystemExit(1)


# DRAM traffic of the dominant kernel per launch comes from the committed ncu --set full capture of this very launch
# shape (dram__bytes_read.sum + dram__bytes_write.sum); bench.py cannot read DRAM counters itself.
NCU_TRAFFIC = {"tensor": {"bytes": 89.75e6, "source": "profiles/r2_ncu_full_loss_kernels.txt (ncu --set full of k_gemm_pair, B=4096 d=768; caches flushed before "
                                                       "the kernel, so G 33.5 MB + operands 12.6 MB + raw student / teacher rows 25.2 MB are all DRAM reads "
                                                       "there: algorithmic bytes 71.3 MB read + 12.6 MB written; inside the step G and the operands are L2 hits)"}}


def kernel_roofline(L, _lib, resident, B, m, d, path, rank, peaks):
    """Times the dominant kernel alone - the CTA-pair GEMM that forms BOTH gradients from the materialised coefficient
    matrix (dU = G V, dV = G^T U) and finishes them in its epilogue (normalise backward + cosine-term gradient), launched
    through the C-ABI building block exactly as the step launches it - with CUDA events on the launching stream, and
    relates its ALGORITHMIC FLOPs (2 x 2*B*B*d; nothing is recomputed in this kernel) to the measured bf16 peak.  The
    coefficient kernel (score tiles recomputed once, 2*B*B*d) that feeds it is timed beside it."""
    if m != B or path != "tensor":   # the roofline entry is taken at N = 1 on the tensor path
        return None
    import ctypes

    from dclip_b200._lib import check, dtype_code, ptr, stream_ptr
    lib = _lib.load_library()
    ops = L.CudaLossOps(path)
    s_img, s_txt, t_img, t_txt = resident
    dev = s_img.device
    u_all = torch.empty((B, d), dtype=ops.op_torch_dtype, device=dev)
    v_all = torch.empty((B, d), dtype=ops.op_torch_dtype, device=dev)
    inv_u, cos_i = ops.prep(s_img, t_img, u_all)
    inv_v, cos_t = ops.prep(s_txt, t_txt, v_all)
    tinv_i = (1.0 / t_img.float().norm(dim=1).clamp_min(1e-12)).contiguous()
    tinv_t = (1.0 / t_txt.float().norm(dim=1).clamp_min(1e-12)).contiguous()
    lse_u, _ = ops.rowlse(u_all, v_all, TEMPERATURE, 0, True)
    lse_v, _ = ops.rowlse(v_all, u_all, TEMPERATURE, 0, False)
    coeffs = torch.empty(lib.dclip_infonce_coeffs_bytes(B), dtype=torch.uint8, device=dev)
    dots = torch.empty(lib.dclip_infonce_coeffs_dot_bytes(B) // 4, dtype=torch.float32, device=dev)
    d_img, d_txt = torch.empty_like(s_img), torch.empty_like(s_txt)

    class Tail(ctypes.Structure):
        _fields_ = [(n, ctypes.c_void_p) for n in ("s_img", "s_txt", "t_img", "t_txt", "inv_u", "inv_v", "cos_img", "cos_txt", "tinv_img", "tinv_txt")] + \
                   [("w_img_over_B", ctypes.c_float), ("w_txt_over_B", ctypes.c_float), ("grad_scale", ctypes.c_float), ("dtype", ctypes.c_int),
                    ("d_s_img", ctypes.c_void_p), ("d_s_txt", ctypes.c_void_p)]
    tail = Tail(ptr(s_img), ptr(s_txt), ptr(t_img), ptr(t_txt), ptr(inv_u), ptr(inv_v), ptr(cos_i), ptr(cos_t), ptr(tinv_i), ptr(tinv_t),
                1.0 / B, 1.0 / B, 1.0, dtype_code(s_img), ptr(d_img), ptr(d_txt))

    def stage1():
        check(lib.dclip_infonce_coeffs(ptr(u_all), ptr(v_all), B, d, TEMPERATURE, ptr(lse_u), ptr(lse_v), ptr(coeffs), ptr(dots), stream_ptr(dev)),
              "infonce_coeffs")

    def stage2():
        check(lib.dclip_infonce_grad_from_coeffs(ptr(coeffs), ptr(dots), ptr(u_all), ptr(v_all), B, d, TEMPERATURE, 1.0 / (2 * B), None, None,
                                                 ctypes.addressof(tail), stream_ptr(dev)), "infonce_grad_from_coeffs")

    def timed_alone(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / n

    stage1()
    sec1 = timed_alone(stage1)
    sec = timed_alone(stage2)
    flops = 4.0 * B * B * d
    achieved = flops / sec / 1e12
    peak = peaks["bf16_tflops"]
    tr = NCU_TRAFFIC.get(path)
    return {"bound": "tensor", "kernel": "k_gemm_pair (CTA pairs, cta_group::2, 256 x 384 tiles; both gradient GEMMs dU = G V, dV = G^T U over the "
                                         "materialised fp16 coefficient matrix in one launch, gradient tail fused into the epilogue)",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_source": peaks["source"] + " bf16 burst (kernel timed alone, back to back, L2-warm)",
            "algorithmic_flops_per_launch": flops, "executed_flops_per_launch": flops,
            "tensor_cores": True, "avg_launch_us": sec * 1e6,
            "traffic": tr["bytes"] if tr else None, "traffic_source": tr["source"] if tr else None,
            "coefficient_kernel": {"kernel": "k_stream_tc<3> (score tiles recomputed once -> fp16 G + radial sums)", "avg_launch_us": sec1 * 1e6,
                                   "flops_per_launch": 2.0 * B * B * d, "tflops": 2.0 * B * B * d / sec1 / 1e12, "frac": 2.0 * B * B * d / sec1 / 1e12 / peak},
            "backward": {"what": "coefficient kernel + gradient GEMMs (the whole backward contraction, algorithmic 4*B*B*d)",
                         "us": (sec + sec1) * 1e6, "frac": flops / (sec + sec1) / 1e12 / peak}}


def weak_scaling_bench(L, dev, rank, world, d, tdtype, path, timed, args):
    """N > 1 only: 4096 rows PER GPU (global batch 4096 N), the same sharded step.  Per-rank tile work is what one GPU
    does at N = 1 times N in the column direction; reported so that the exchange is judged on a shape where the
    compute can hide it.  Rank 0 returns the dict."""
    import workloads as synthetic
    m = 4096
    Bw = m * world
    inputs = [t.to(tdtype)[rank * m:(rank + 1) * m].contiguous().to(dev) for t in synthetic.loss_inputs(Bw, d, seed=43)]

    def fn():
        return L.sharded_distill_loss_fwd_bwd(*inputs, temperature=TEMPERATURE, path=path)

    graphed = None if args.no_graph else L.GraphedStep(lambda *a: L.sharded_distill_loss_fwd_bwd(*a, temperature=TEMPERATURE, path=path), inputs)
    ms = timed((lambda: graphed.graph.replay()) if graphed is not None else fn, max(5, args.steps // 2), 3)
    return {"rows_per_gpu": m, "global_batch": Bw, "ms_per_step": ms, "pairs_per_s": Bw / (ms * 1e-3), "scaling": "weak",
            "note": "ideal time grows with N (B^2 tiles / N per GPU = 4096 x 4096 N): compare pairs/s per GPU x B/4096"}


def sharded_gallery_bench(dev, rank, world, n_gallery, n_query):
    """BASELINE configs[4]: cosine top-10 over an N x 768 fp16 gallery sharded over the ranks (rank r holds
    rows [r N/P, (r+1) N/P)), every rank scores all queries against its shard on the tcgen05 path, the
    per-GPU exact top-10 lists are all-gathered and merged by (-score, +index).  Strong scaling.  256 sampled
    queries are re-run on the exact FP32 path over the same shards: (score, index) must be bit-identical."""
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
        dist.broadcast(planted, src=0)
    gallery = R.TensorGallery(g)

    def search(queries, path):
        s, i = R.topk_ip(queries, g, 10, g_index_offset=lo, path=path, gallery=gallery if path == "tensor" else None)
        nq = queries.shape[0]
        if world > 1:
            all_s = torch.empty((world * nq, 10), dtype=s.dtype, device=dev)
            all_i = torch.empty((world * nq, 10), dtype=i.dtype, device=dev)
            dist.all_gather_into_tensor(all_s, s)
            dist.all_gather_into_tensor(all_i, i)
            s, i = R.topk_merge(all_s.view(world, nq, 10), all_i.view(world, nq, 10))
        return s, i

    search(q, "tensor")
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s, i = search(q, "tensor")
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t) * 1e-3
    r1 = float((i[:, 0] == planted).float().mean())
    # exactness at scale: a fixed sample of the queries again on the FP32 CUDA-core path (every score the contract's fmaf chain)
    nchk = min(256, n_query)
    sel = torch.arange(0, n_query, max(1, n_query // nchk), device=dev)[:nchk]
    s32, i32 = search(q[sel].contiguous(), "fp32")
    same = bool(torch.equal(i32, i[sel]) and torch.equal(s32, s[sel]))
    recheck = {"status": "ok" if same else "FAILED", "queries": int(sel.numel()),
               "what": "tcgen05 path vs exact FP32 path over the same sharded gallery: top-10 scores and indices bit-identical"}
    del g, gallery
    torch.cuda.empty_cache()
    peaks = load_peaks()
    tflops = 2.0 * n_query * n_gallery * d / sec / 1e12
    per_gpu = tflops / world
    return {"metric": "retrieval queries/sec top-10", "queries_per_s": n_query / sec, "ms": sec * 1e3, "planted_R@1": r1,
            "fp32_recheck": recheck,
            "tflops_aggregate": tflops, "frac_of_bf16_sustained_peak": tflops / (world * peaks["bf16_tflops_sustained"]),
            "roofline": {"bound": "tensor", "kernel": "k_stream_tc<2,*> (tcgen05 score tiles + fused top-32 candidate lists), per GPU",
                         "achieved": per_gpu, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": per_gpu / peaks["bf16_tflops_sustained"],
                         "peak_source": peaks["source"] + " bf16 sustained (a 0.1-1 s kernel under the power cap)",
                         "algorithmic_flops_per_launch": 2.0 * n_query * n_gallery * d / world,
                         "traffic": None, "traffic_source": "profiles/ (ncu capture of one shard; the gallery is read once per wave of CTAs)"},
            "workload": f"exact cosine top-10, {n_gallery} x {d} fp16 gallery sharded over {world} GPU(s), {n_query} queries "
                        "(BASELINE configs[4]); tcgen05 candidates + exact fp32 re-scoring, bit-exact contract"}


def retrieval_bench(dev, with_cpu):
    """Second BASELINE metric: retrieval queries/s.  (a) Flickr-shaped Recall@K eval, 1k images x 5k
    captions, d=512 (configs[1]; 6000 ranked queries per eval), the reference's own evaluator timed beside it.
    (b) exact top-10 over a 1M x 768 fp16 gallery slice, 4096 queries, with the host stand-in for faiss."""
    import numpy as np

    from dclip_b200 import retrieval as R
    import workloads as synthetic
    out = {}
    cap, img, image_ids, cap_ids = synthetic.retrieval_inputs(1000, 5, 512, 7.0, seed=44)      # the tests/golden/retrieval_cfg2 recipe
    cap_d, img_d = torch.from_numpy(cap).to(dev), torch.from_numpy(img).to(dev)
    res = {}
    for rank_path in ("fp32", "tensor"):
        for _ in range(2):
            R.retrieval_metrics_from_embeddings(cap_d, img_d, image_ids, cap_ids, path=rank_path)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            m = R.retrieval_metrics_from_embeddings(cap_d, img_d, image_ids, cap_ids, path=rank_path)
        torch.cuda.synchronize()
        res[rank_path] = ((time.perf_counter() - t0) / reps, m)
    (dt32, m32), (dt, m) = res["fp32"], res["tensor"]
    out["flickr_eval"] = {"queries_per_s": 6000 / dt, "ms_per_eval": dt * 1e3, "ms_per_eval_fp32_path": dt32 * 1e3, "t2i_R@1": m["t2i"]["R@1"],
                          "i2t_R@1": m["i2t"]["R@1"], "metrics_identical_on_both_paths": m == m32,
                          "path": "tensor (tcgen05 score tiles, exact re-scoring of the band around each target score; ranks bit-exact)",
                          "workload": "1k images x 5k captions, d=512, R@1/5/10 both directions (BASELINE configs[1])"}
    if with_cpu:
        from oracle import reference_arm
        arm = reference_arm.load()
        cores = os.cpu_count() or 1
        arm.retrieval_eval(cap, img, image_ids, cap_ids)
        t0 = time.perf_counter()
        mc = arm.retrieval_eval(cap, img, image_ids, cap_ids)
        ct = time.perf_counter() - t0
        same = all(mc[dn][k] == m[dn][k] for dn in ("t2i", "i2t") for k in ("R@1", "R@5", "R@10"))
        out["flickr_eval"]["cpu_baseline"] = {"value": 6000 / ct, "unit": "ranked queries/s", "cores": cores, "kind": arm.kind, "ms_per_eval": ct * 1e3,
                                              "sample": "one full eval: normalise + chunked similarity + calculate_retrieval_metrics (eval_scripts/flickr30k_eval.py:16-88, :243-266)",
                                              "recall_at_k_identical_to_gpu": same}
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
    if with_cpu:
        # faiss is absent and unpinned: the stand-in BASELINE.md names - Q @ G^T + top-k on the host cores over
        # the same 1 M-row slice, 512 queries (a few seconds), extrapolated linearly to the 10 M-row gallery
        from oracle import reference_arm
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        gq = q[:512].float().cpu().numpy()
        gg = g.float().cpu().numpy()
        reference_arm.topk_stand_in(gq[:32], gg, 10)
        t0 = time.perf_counter()
        reference_arm.topk_stand_in(gq, gg, 10)
        ct = time.perf_counter() - t0
        out["top10"]["cpu_baseline"] = {"value": 512 / ct, "unit": "queries/s at 1 M rows", "cores": cores, "kind": "port",
                                        "extrapolated_to_10M_rows_queries_per_s": 512 / ct / 10.0,
                                        "sample": f"512 queries x 1 M x 768 fp32, torch matmul + topk on {cores} threads, {ct:.1f} s (stand-in for faiss IndexFlatIP.search; extrapolated)"}
    return out


def fp32_input_bench(L, dev, B, d):
    """The same objective for fp32 inputs (the reference's own dtype): `path="tensor"` rounds the L2-normalised rows to fp16
    operands once (fp32 accumulate, fp32 gradients: the 1e-3 tier), `path="fp32"` is the exact CUDA-core tier (1e-5) that
    `path="auto"` selects for fp32 inputs.  Device-timed, eager launches, L2 flushed between steps."""
    import workloads as synthetic
    inp = [t.to(dev) for t in synthetic.loss_inputs(B, d, seed=42)]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = {"workload": f"distill-loss fwd+bwd, B={B}, d={d}, fp32 inputs and gradients"}
    ref = None
    for path, reps in (("tensor", 10), ("fp32", 4)):
        for _ in range(2):
            res = L.distill_loss_fwd_bwd(*inp, temperature=TEMPERATURE, path=path)
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = L.distill_loss_fwd_bwd(*inp, temperature=TEMPERATURE, path=path)
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        out[f"ms_per_step_{path}_path"] = tot / reps
        if path == "tensor":
            ref = res
        else:
            out["tensor_vs_fp32_path_rel_grad_diff"] = float((ref[1] - res[1]).norm() / res[1].norm())
            out["tensor_vs_fp32_path_rel_loss_diff"] = float(((ref[0] - res[0]).abs() / res[0].abs().clamp_min(1e-6)).max())
    return out


def teacher_flops(B, T, R, d):
    """Algorithmic FLOPs of the meta-teacher step as the reference uses it (weight gradients only):
    forward 8 d^2 N + 8 B T R d, backward 10 d^2 N + 16 B T R d, N = B (T + R)."""
    n = B * (T + R)
    return 18.0 * d * d * n + 24.0 * B * T * R * d


def teacher_step_bench(dev, with_cpu):
    """BASELINE configs[0] (B=32, R=8, T=77, d=512, fp32) and configs[2] (B=32, R=36, T=77, d=768, bf16): cross-attention ->
    aggregation x2 -> combine -> InfoNCE -> backward to the 12 parameter tensors (train_contrastive_teacher.py:347-361),
    replayed as ONE CUDA graph; the second figure adds the fused Adam step (:362)."""
    from dclip_b200 import losses as L
    from dclip_b200.optim import FusedAdam
    from dclip_b200.patch_text_aggregation import GraphedTeacherStep, PatchTextAggregation
    import workloads as synthetic
    peaks = load_peaks()
    out = {}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    flush_rd = torch.zeros(32 * 1024 * 1024, dtype=torch.int64, device=dev)
    for name, (B, T, R, d, dtype) in (("cfg0", (32, 77, 8, 512, torch.float32)), ("cfg2", (32, 77, 36, 768, torch.bfloat16))):
        params = synthetic.perturbed_cross_attention_params(d, 8, 42)
        text, image = synthetic.xattn_inputs(B, T, R, d, 42, ragged=True)
        txt = torch.randn(B, d, generator=torch.Generator().manual_seed(43))

        def loss_fn(mod, t, i, x):
            return L.compute_contrastive_loss(mod.global_embedding(t, i), x, temperature=TEMPERATURE)

        res = {}
        for with_opt in (False, True):
            mod = PatchTextAggregation(embed_dim=d, num_heads=8)
            mod.cross_modal_attention.load_state_dict(params)
            mod = mod.to(dev).to(dtype)
            inputs = [text.to(dev).to(dtype), image.to(dev).to(dtype), txt.to(dev).to(dtype)]
            opt = FusedAdam(mod.parameters(), lr=1e-5) if with_opt else None
            gs = GraphedTeacherStep(mod, loss_fn, inputs, optimizer=opt)
            inputs = gs.static_in          # inputs resident in the graph's own buffers (where a training loop's H2D copies land)
            for _ in range(3):
                gs(*inputs)
            torch.cuda.synchronize()
            reps, tot = 20, 0.0
            for _ in range(reps):
                flush.zero_()
                flush_rd.sum()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                loss = gs(*inputs)
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            res["us_per_step_with_adam" if with_opt else "us_per_step"] = tot / reps * 1e3
            if not with_opt:
                res["loss"] = float(loss)
        us = res["us_per_step"]
        fl = teacher_flops(B, T, R, d)
        res.update({"pairs_per_s": B / (us * 1e-6), "algorithmic_gflop": fl / 1e9,
                    "roofline": {"bound": "tensor" if dtype != torch.float32 else "tensor (fp32 tier: GEMMs on tcgen05 through bf16 x 3 operand splitting = 6x the MMA work; attention cores on CUDA cores)",
                                 "achieved": fl / (us * 1e-6) / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                                 "frac": fl / (us * 1e-6) / 1e12 / peaks["bf16_tflops"], "peak_source": peaks["source"] + " bf16 burst",
                                 "traffic": None},
                    "workload": f"meta-teacher step fwd+bwd, B={B} T={T} R={R} d={d} {str(dtype).replace('torch.', '')} "
                                f"(BASELINE configs[{0 if name == 'cfg0' else 2}]), one CUDA graph, inputs resident in the graph's input buffers, "
                                f"L2 flushed between steps"})
        if with_cpu:
            from oracle import reference_arm
            arm = reference_arm.load()
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            arm.teacher_step(params, text, image, txt)
            n, t0 = 0, time.perf_counter()
            while n < 10 and time.perf_counter() - t0 < 6.0:
                arm.teacher_step(params, text, image, txt)
                n += 1
            ct = (time.perf_counter() - t0) / n
            res["cpu_baseline"] = {"value": B / ct, "unit": "pairs/s", "ms_per_step": ct * 1e3, "cores": cores, "kind": arm.kind,
                                   "sample": f"{n} steps of the reference CrossModalAttention + aggregation + contrastive loss + backward, torch fp32 on {cores} threads"}
        out[name] = res
    return out


def cpu_baseline(B, d):
    import workloads as synthetic
    from oracle import reference_arm
    arm = reference_arm.load()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    inputs = synthetic.loss_inputs(B, d, seed=42)
    arm.loss_step(*inputs, TEMPERATURE)
    n = 0
    t0 = time.perf_counter()
    while True:
        arm.loss_step(*inputs, TEMPERATURE)
        n += 1
        el = time.perf_counter() - t0
        if el > 12.0 or n >= 40:
            break
    what = "the reference's own loss functions + backward (oracle/_ref, unmodified)" if arm.kind == "reference" else "the oracle port"
    return {"value": B * n / el, "unit": UNIT, "cores": cores, "kind": arm.kind,
            "sample": f"{n} full steps (B={B}, d={d}) of {what} (torch fp32 CPU, {cores} threads), {el:.1f} s"}


if __name__ == "__main__":
    main()
