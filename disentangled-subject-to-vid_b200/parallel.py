"""Multi-GPU partitioning of the denoising loop on one NVLink/NVSwitch box: one process per GPU (torchrun),
`torch.distributed` for the plumbing.

The reference is single-GPU (no torch.distributed anywhere); this is new design.  The 2P sequences of a
step (P prompts x {uncond, cond}) never interact inside the transformer, so the path shards with NO per-step
collective when every rank owns whole prompts:

  prompt sharding   rank r owns prompts  r, r+W, r+2W, ...  with both CFG halves local  -> zero traffic during the
                    loop, ONE all-gather of the final latents (2.25 MB per 49x480x720 sample).
  CFG sharding      ranks (2j, 2j+1) form a pair owning prompts j, j + W/2, ...; each rank runs ONE CFG half of them, the pair
                    exchanges the bf16 model output once per step (one all_gather_into_tensor inside the pair group, 2.25 MB
                    per sample at 480x720, 6 MB at 720x1280, on a side stream) and both ranks apply the identical bit-exact
                    CFG+DDIM kernel redundantly, so latents stay replicated without a broadcast.  Halves the per-step latency
                    of a video; the default when there are fewer prompts than ranks (BASELINE configs[3]: 2 prompts on 4 GPUs).

Weights are replicated (5B bf16 = 11 GB of 180 GB).  Results are bit-identical to the 1-GPU run of the same build
because per-sample reduction orders do not depend on the batch composition.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


@dataclass
class ShardPlan:
    mode: str                 # "prompt" | "cfg"
    world: int
    rank: int
    prompts: List[int]        # global prompt indices this rank works on
    cfg_half: Optional[int]   # None (both halves local) or 0 = uncond / 1 = cond
    pair_ranks: Optional[List[int]]  # ranks sharing my prompt (cfg mode)


def plan(num_prompts: int, world: int, rank: int, mode: Optional[str] = None) -> ShardPlan:
    """Partition `num_prompts` guided samples over `world` ranks.

    mode None picks prompt sharding when there is at least one prompt per rank, else CFG sharding; "cfg" forces CFG sharding
    (rank pair (2j, 2j+1) owns prompts j, j + W/2, j + W, ...: half the per-step latency per video for one small exchange).
    One prompt on more than two ranks would need head-parallel attention inside a sequence — not part of this path."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    if mode not in (None, "prompt", "cfg"):
        raise ValueError("mode must be None, 'prompt' or 'cfg'")
    if mode is None:
        mode = "prompt" if num_prompts >= world else "cfg"
    if mode == "prompt":
        if num_prompts % world:
            raise ValueError(f"{num_prompts} prompts do not divide over {world} ranks")
        return ShardPlan("prompt", world, rank, list(range(rank, num_prompts, world)), None, None)
    pairs = world // 2
    if world % 2 or pairs == 0 or num_prompts % pairs:
        raise ValueError(f"cannot place {num_prompts} prompts on {world} ranks with CFG sharding: need an even number of ranks and "
                         f"prompts % (ranks / 2) == 0 (the two CFG halves of a prompt are the only independent units inside it)")
    j = rank // 2
    return ShardPlan("cfg", world, rank, list(range(j, num_prompts, pairs)), rank % 2, [2 * j, 2 * j + 1])


def select_prompt_embeds(prompt_embeds_2p: torch.Tensor, num_prompts: int, sp: ShardPlan) -> torch.Tensor:
    """prompt_embeds_2p is ordered [negative x P, positive x P] (S/custom_cogvideox_pipe.py:196); returns this rank's rows in
    the same [neg..., pos...] order (or the single half in cfg mode)."""
    neg, pos = prompt_embeds_2p[:num_prompts], prompt_embeds_2p[num_prompts:]
    idx = torch.tensor(sp.prompts, device=prompt_embeds_2p.device)
    if sp.mode == "prompt":
        return torch.cat([neg.index_select(0, idx), pos.index_select(0, idx)], dim=0)
    return (neg if sp.cfg_half == 0 else pos).index_select(0, idx)


class PairExchange:
    """Per-step exchange of the two CFG halves inside a rank pair (cfg mode): ONE `all_gather_into_tensor` of the model output
    ([P_local, ...] per rank -> [2 * P_local, ...] = (uncond, cond), 2.25 MB bf16 per 49 x 480 x 720 sample) into a buffer that is
    allocated once, issued on a side stream so that the compute stream only waits on an event (whatever the caller enqueues
    between `start()` and `finish()` overlaps the transfer).  `times_ms()` returns the device time of every exchange."""

    def __init__(self, sp: ShardPlan, record_times: bool = False):
        self.sp = sp
        self.group = None
        self.buf = None
        self.side = None
        self.record = record_times
        self._ev = []
        self._pending = None
        if sp.mode == "cfg":
            # every rank must create every pair group, in the same order
            for j in range(sp.world // 2):
                g = dist.new_group(ranks=[2 * j, 2 * j + 1])
                if j == sp.rank // 2:
                    self.group = g

    def start(self, my_half: torch.Tensor):
        my_half = my_half.contiguous()
        shape = (2 * my_half.shape[0],) + tuple(my_half.shape[1:])     # rank order inside the pair = (uncond, cond) along dim 0
        if self.buf is None or tuple(self.buf.shape) != shape or self.buf.dtype != my_half.dtype or self.buf.device != my_half.device:
            self.buf = torch.empty(shape, dtype=my_half.dtype, device=my_half.device)
        if my_half.is_cuda:
            if self.side is None:
                self.side = torch.cuda.Stream(device=my_half.device)
            cur = torch.cuda.current_stream()
            self.side.wait_stream(cur)
            with torch.cuda.stream(self.side):
                e0 = torch.cuda.Event(enable_timing=self.record)
                e1 = torch.cuda.Event(enable_timing=self.record)
                e0.record()
                dist.all_gather_into_tensor(self.buf, my_half, group=self.group)
                e1.record()
            my_half.record_stream(self.side)
            self._pending = e1
            if self.record:
                self._ev.append((e0, e1))
        else:
            dist.all_gather_into_tensor(self.buf, my_half, group=self.group)
            self._pending = None

    def finish(self) -> torch.Tensor:
        """-> [2 * P_local, ...] ordered (uncond x P_local, cond x P_local), like the reference's CFG batch."""
        if self._pending is not None:
            torch.cuda.current_stream().wait_event(self._pending)
        return self.buf

    def both_halves(self, my_half: torch.Tensor) -> torch.Tensor:
        self.start(my_half)
        return self.finish()

    def times_ms(self):
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in self._ev]


def _owner_index(num_prompts: int, sp: ShardPlan) -> List[int]:
    """For every global prompt p, the flat position of its latents in the all-gathered [world, P_local, ...] buffer."""
    per = len(sp.prompts)
    pos = [0] * num_prompts
    for r in range(sp.world):
        o = plan(num_prompts, sp.world, r, sp.mode)
        if o.mode == "cfg" and o.cfg_half == 1:
            continue            # both ranks of a pair hold identical latents; the even rank's copy is kept
        for j, p_ in enumerate(o.prompts):
            pos[p_] = r * per + j
    return pos


def gather_latents(local: torch.Tensor, num_prompts: int, sp: ShardPlan) -> torch.Tensor:
    """The single collective of the path: ONE all-gather of the final latents (or decoded frames) [P_local, ...] into global
    prompt order on every rank."""
    if sp.world == 1:
        return local
    local = local.contiguous()
    buf = torch.empty((sp.world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, local)
    if not hasattr(sp, "_owner") or sp._owner[0] != num_prompts:
        sp._owner = (num_prompts, torch.tensor(_owner_index(num_prompts, sp), device=local.device))
    return buf.index_select(0, sp._owner[1].to(local.device))


def sharded_denoise(sp: ShardPlan, num_prompts: int, latents: torch.Tensor, prompt_embeds_2p: torch.Tensor,
                    ref_img_states: torch.Tensor, timesteps: Sequence[int], model_fn: Callable, step_fn: Callable,
                    guidance_fn: Callable[[int], float]) -> torch.Tensor:
    """Run the loop on this rank's shard and return the gathered final latents.

    model_fn(latent_in [B,...], prompt_embeds [B,...], ref [Br,...], t) -> model output [B,...]
    step_fn(noise_pred_2 [2P_local,...] (uncond first), t, latents [P_local,...], guidance) -> next latents
    Device / backend agnostic (the gloo CPU tests drive it with stand-in callables)."""
    idx = torch.tensor(sp.prompts, device=latents.device)
    lat = latents.index_select(0, idx).contiguous()
    ref = ref_img_states.index_select(0, idx).contiguous() if ref_img_states.shape[0] == num_prompts else ref_img_states
    pe = select_prompt_embeds(prompt_embeds_2p, num_prompts, sp)
    xchg = PairExchange(sp) if sp.mode == "cfg" else None
    for i, t in enumerate(timesteps):
        if sp.mode == "prompt":
            noise = model_fn(torch.cat([lat, lat], dim=0), pe, ref, t)
        else:
            noise = xchg.both_halves(model_fn(lat, pe, ref, t))
        lat = step_fn(noise, t, lat, guidance_fn(i))
    return gather_latents(lat, num_prompts, sp)
