"""Times the meta-teacher step (cross-attention + aggregation + InfoNCE, fwd+bwd) at the BASELINE
cfg1 / cfg3 shapes on the GPU and, optionally, the oracle port on the host cores."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dclip_b200 import losses as L
from dclip_b200.patch_text_aggregation import PatchTextAggregation
from oracle import reference_math as om
import workloads as synthetic

def run(B, T, R, d, dtype, cpu=False):
    dev = torch.device("cuda:0")
    params = om.perturbed_cross_attention_params(d, 8, 3)
    m = PatchTextAggregation(embed_dim=d, num_heads=8)
    m.cross_modal_attention.load_state_dict(params)
    m = m.to(dev).to(dtype)
    text, image = synthetic.xattn_inputs(B, T, R, d, 3, ragged=True)
    txt = torch.randn(B, d, generator=torch.Generator().manual_seed(1))
    tg, ig, xg = text.to(dev).to(dtype), image.to(dev).to(dtype), txt.to(dev).to(dtype)
    def step():
        for p in m.parameters(): p.grad = None
        g = m.global_embedding(tg, ig)
        loss = L.compute_contrastive_loss(g, xg, path="fp32")
        loss.backward()
        return loss
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    line = f"teacher step B={B} T={T} R={R} d={d} {dtype}: {ms:.3f} ms/step = {B/ms*1e3:.0f} pairs/s"
    # the same step replayed as one CUDA graph
    from dclip_b200.patch_text_aggregation import GraphedTeacherStep
    ref_grads = [p.grad.clone() for p in m.parameters()]
    ref_loss = float(step())
    gs = GraphedTeacherStep(m, lambda mod, t, i, x: L.compute_contrastive_loss(mod.global_embedding(t, i), x, path="fp32"), [tg, ig, xg])
    for _ in range(3): gs(tg, ig, xg)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n): gl = gs(tg, ig, xg)
    e1.record(); torch.cuda.synchronize()
    gms = e0.elapsed_time(e1) / n
    same = all(torch.equal(a, p.grad) for a, p in zip(ref_grads, m.parameters())) and float(gl) == ref_loss
    line += f" | CUDA graph: {gms:.3f} ms/step (bit-identical to eager: {same})"
    if cpu:
        torch.set_num_threads(os.cpu_count())
        om.teacher_step(params, text, image, txt)
        t0 = time.perf_counter()
        for _ in range(3): om.teacher_step(params, text, image, txt)
        c = (time.perf_counter() - t0) / 3 * 1e3
        line += f" | CPU oracle port ({os.cpu_count()} threads): {c:.1f} ms/step"
    print(line, flush=True)

if len(sys.argv) > 1 and sys.argv[1] == "bf16":
    run(32, 77, 36, 768, torch.bfloat16)
else:
    run(32, 77, 8, 512, torch.float32, cpu=True)
    run(32, 77, 36, 768, torch.bfloat16, cpu=True)
    run(32, 77, 36, 768, torch.float32)
