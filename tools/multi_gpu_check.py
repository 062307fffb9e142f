"""Determinism check of SURVEY §8e on real GPUs (run under torchrun with 2 ranks): the CFG-sharded loop (one CFG half per
GPU, one NCCL pair exchange per step) and the prompt-sharded loop (one prompt per GPU, no per-step collective) must return
latents BIT-IDENTICAL to the single-GPU loop of the same build.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import s2v_b200
from s2v_b200 import parallel

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
bf16 = torch.bfloat16
torch.manual_seed(0)
model = s2v_b200.CogVideoXTransformer3DModel(num_attention_heads=4, num_layers=3, time_embed_dim=64, text_embed_dim=64,
                                             use_rotary_positional_embeddings=True).to(bf16)
s2v_b200.inject_lora(model, 8, 4.0)
g = torch.Generator().manual_seed(1)
with torch.no_grad():
    for n, p in model.named_parameters():
        p.copy_((0.05 * torch.randn(p.shape, generator=g)).to(p.dtype))
model = model.to(dev)
sched = s2v_b200.CogVideoXDDIMScheduler.for_cogvideox(1.0)
sched.set_timesteps(4)
steps = list(sched._timesteps_host)
pipe = s2v_b200.CustomCogVideoXPipeline(None, None, model, None, sched)
h, w, Fr = 16, 24, 3
results = {}
for P in (1, 2):
    lat = torch.randn(P, Fr, 16, h, w, generator=g).to(bf16).to(dev)
    pe = (0.2 * torch.randn(2 * P, 226, 64, generator=g)).to(bf16).to(dev)
    ref = (0.7 * torch.randn(P, 1, 16, h, w, generator=g)).to(bf16).to(dev)
    cos, sin = pipe.rotary_tables(h * 8, w * 8, Fr, dev)
    n = cos.shape[0] // (Fr + 1)

    def model_fn(x, e, r, t):
        B = x.shape[0]
        rr = r if 2 * r.shape[0] == B else r            # eval=True doubles ref inside the model when B == 2 * Br
        ev = (2 * r.shape[0] == B)
        return model(hidden_states=x.contiguous(), encoder_hidden_states=e.contiguous(), ref_img_states=rr.contiguous(),
                     timestep=torch.full((B,), float(t), device=dev), image_rotary_emb=(cos[n:], sin[n:]),
                     ref_image_rotary_emb=(cos[:n], sin[:n]), return_dict=False, eval=ev)[0]

    def step_fn(noise2, t, l, gs):
        return sched.step_cfg(noise2.contiguous(), t, l.contiguous(), gs)

    single = parallel.sharded_denoise(parallel.plan(P, 1, 0), P, lat, pe, ref, steps, model_fn, step_fn, lambda i: 6.0)
    sharded = parallel.sharded_denoise(parallel.plan(P, world, rank), P, lat, pe, ref, steps, model_fn, step_fn, lambda i: 6.0)
    torch.cuda.synchronize()
    results[f"P{P}_{parallel.plan(P, world, rank).mode}"] = bool(torch.equal(single, sharded))
# ---- the public call in CFG-parallel mode (pipe.enable_cfg_parallel): every rank pair owns world/2 ... prompts, one CFG half per GPU
if world % 2 == 0:
    pairs = world // 2
    spc = parallel.plan(pairs, world, rank, mode="cfg")
    gp = torch.Generator().manual_seed(50 + rank // 2)                   # both ranks of a pair draw the same prompt
    lat = torch.randn(1, Fr, 16, h, w, generator=gp).to(bf16).to(dev)
    pos = (0.2 * torch.randn(1, 226, 64, generator=gp)).to(bf16).to(dev)
    neg = (0.2 * torch.randn(1, 226, 64, generator=gp)).to(bf16).to(dev)
    ref = (0.7 * torch.randn(1, 1, 16, h, w, generator=gp)).to(bf16).to(dev)
    kw = dict(prompt_embeds=pos, negative_prompt_embeds=neg, ref_img_states=ref, latents=lat, height=h * 8, width=w * 8, num_frames=(Fr - 1) * 4 + 1,
              num_inference_steps=4, guidance_scale=6.0, use_dynamic_cfg=True, output_type="latent", return_dict=False)
    single = pipe(**kw)[0].clone()
    pipe.enable_cfg_parallel(spc, record_times=True)
    sharded = pipe(**kw)[0].clone()
    xt = pipe._cfg_xchg.times_ms()
    pipe.disable_cfg_parallel()
    results["pipe_call_cfg_parallel"] = bool(torch.equal(single, sharded))
    gathered = parallel.gather_latents(sharded, pairs, spc)
    results["gather_in_prompt_order"] = bool(gathered.shape[0] == pairs and torch.equal(gathered[rank // 2], sharded[0]))
    results["pair_exchange_ms_median"] = sorted(xt)[len(xt) // 2]
ok = torch.tensor([int(all(bool(v) for v in results.values()))], device=dev)
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"world": world, "bit_identical_to_single_gpu": results, "all_ranks_ok": bool(ok.item())}))
dist.destroy_process_group()
sys.exit(0 if ok.item() else 1)
