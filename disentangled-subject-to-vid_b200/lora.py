"""LoRA adapters in the PEFT module layout.

The reference injects `peft.LoraConfig(r=128, lora_alpha=64, target_modules=[...])` into the transformer
(S/inference.py:218-225 -> D/loaders/peft.py:112-148).  peft is an un-vendored third-party dependency; its published
layer contract is what the B200 engine consumes: a wrapped layer exposes `base_layer`, `lora_A[adapter].weight`,
`lora_B[adapter].weight`, `scaling[adapter]` and computes `base(x) + lora_B(lora_A(x)) * scaling` (dropout 0, no bias).

`inject_lora` builds that layout without peft (for standalone use and tests); `read_linear` reads it from EITHER this
module's wrappers or real peft layers attached to stock diffusers modules, so the engine works unchanged behind
S/inference.py.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterable, Optional

import torch
import torch.nn as nn

# S/inference.py:222
REFERENCE_TARGETS = ["to_k", "to_q", "to_v", "to_out.0", "proj", "text_proj", "norm1.linear", "norm2.linear", "ff.net.2"]


class LoraWrapped(nn.Module):
    """PEFT-layout LoRA wrapper around nn.Linear / nn.Conv2d.  Holds parameters only: the arithmetic is executed by the
    fused sm_100a GEMM (LoRA-B rides along as extra K columns), never by this module."""

    def __init__(self, base: nn.Module, r: int, alpha: float, adapter: str = "default"):
        super().__init__()
        self.base_layer = base
        dt, dev = base.weight.dtype, base.weight.device
        if isinstance(base, nn.Linear):
            a, b = nn.Linear(base.in_features, r, bias=False), nn.Linear(r, base.out_features, bias=False)
        elif isinstance(base, nn.Conv2d):
            a = nn.Conv2d(base.in_channels, r, base.kernel_size, base.stride, base.padding, bias=False)
            b = nn.Conv2d(r, base.out_channels, (1, 1), (1, 1), bias=False)
        else:
            raise TypeError(f"cannot wrap {type(base).__name__} with LoRA")
        nn.init.zeros_(b.weight)  # PEFT init_lora_weights=True: B = 0
        self.lora_A = nn.ModuleDict({adapter: a.to(device=dev, dtype=dt)})
        self.lora_B = nn.ModuleDict({adapter: b.to(device=dev, dtype=dt)})
        self.scaling = {adapter: alpha / r}
        self.r = {adapter: r}
        self.active_adapter = adapter

    @property
    def weight(self):
        return self.base_layer.weight

    @property
    def bias(self):
        return self.base_layer.bias

    def forward(self, *args, **kwargs):
        raise RuntimeError("LoraWrapped holds parameters for the fused B200 engine; it has no eager forward")


def _matches(name: str, targets: Iterable[str]) -> bool:
    return any(name == t or name.endswith("." + t) for t in targets)


def invalidate_packed(model: nn.Module):
    """Drop every packed-weight snapshot that hangs off `model` (the engine concatenates q|k|v and the LoRA factors, so they are
    copies): this package's `_engine` / `_pb` / `_ws` caches and, on a stock model bound with `attach()`, the bound engine is
    re-packed.  Called by everything that changes weights or adapters."""
    for m in model.modules():
        if getattr(m, "_engine", None) is not None:
            m._engine = None
        if getattr(m, "_pb", None) is not None:
            m._pb = None
        if getattr(m, "_ws", None) is not None and not isinstance(m._ws, dict):
            m._ws = None
    eng = getattr(model, "_s2v_engine", None)
    if eng is not None:
        eng.repack()


def inject_lora(model: nn.Module, r: int, alpha: float, targets: Iterable[str] = REFERENCE_TARGETS, adapter: str = "default"):
    """Wrap every nn.Linear / nn.Conv2d whose dotted name matches a target suffix (PEFT's matching rule)."""
    targets = list(targets)
    names = [n for n, m in model.named_modules() if _matches(n, targets) and isinstance(m, (nn.Linear, nn.Conv2d))]
    for n in names:
        parent_name, _, child = n.rpartition(".")
        parent = model.get_submodule(parent_name) if parent_name else model
        wrapped = LoraWrapped(parent[int(child)] if child.isdigit() else getattr(parent, child), r, alpha, adapter)
        if child.isdigit():
            parent[int(child)] = wrapped
        else:
            setattr(parent, child, wrapped)
    invalidate_packed(model)
    return names


def load_lora_state_dict(model: nn.Module, state: Dict[str, torch.Tensor], adapter: str = "default", prefix: str = "transformer.",
                         strict: bool = True):
    """Load a `pytorch_lora_weights_transformer.safetensors`-style dict.  Accepted key forms (after stripping `prefix`):
    `<module>.lora_{A,B}.weight` (the file format, README.md:70-75) and `<module>.lora_{A,B}.<adapter>.weight` (the PEFT form that
    `convert_unet_state_dict_to_peft` + `set_peft_model_state_dict` produce, S/inference.py:94-100).  Like the reference loader,
    unexpected keys are reported: with `strict` (default) any key that is not a LoRA factor of an adapted module, or an adapted
    module left without both factors, raises — a checkpoint that silently loads nothing would run with B = 0 (no subject)."""
    own = dict(model.named_modules())
    loaded, unexpected = {}, []
    for k, v in state.items():
        k = k[len(prefix):] if k.startswith(prefix) else k
        parts = k.split(".")
        if len(parts) >= 3 and parts[-1] == "weight" and parts[-2] in ("lora_A", "lora_B"):
            base, ab, ad = ".".join(parts[:-2]), parts[-2], adapter
        elif len(parts) >= 4 and parts[-1] == "weight" and parts[-3] in ("lora_A", "lora_B"):
            base, ab, ad = ".".join(parts[:-3]), parts[-3], parts[-2]
        else:
            unexpected.append(k)
            continue
        layer = own.get(base)
        if layer is None or not hasattr(layer, ab) or ad not in getattr(layer, ab):
            raise KeyError(f"LoRA key {k!r} does not match an adapted module (adapter {ad!r})")
        getattr(layer, ab)[ad].weight.data.copy_(v)
        loaded.setdefault(base, set()).add(ab)
    adapted = [n for n, m in own.items() if hasattr(m, "lora_A") and hasattr(m, "base_layer")]
    incomplete = [n for n in adapted if loaded.get(n, set()) != {"lora_A", "lora_B"}]
    if strict and (unexpected or incomplete or not loaded):
        raise KeyError(f"load_lora_state_dict: {len(loaded)} of {len(adapted)} adapted modules received both factors; "
                       f"unexpected keys: {unexpected[:5]}{'...' if len(unexpected) > 5 else ''}; "
                       f"modules without both factors: {incomplete[:5]}{'...' if len(incomplete) > 5 else ''}")
    invalidate_packed(model)
    return sum(len(v) for v in loaded.values())


@dataclass
class LinearParams:
    weight: torch.Tensor            # [out, in] (Conv2d weights are flattened to [out, in*kh*kw])
    bias: Optional[torch.Tensor]
    lora_a: Optional[torch.Tensor]  # [r, in]
    lora_b: Optional[torch.Tensor]  # [out, r]
    scale: float


def read_linear(layer: nn.Module, adapter: Optional[str] = None) -> LinearParams:
    """Parameters of a plain or LoRA-wrapped (this module's or peft's) Linear / Conv2d, as 2-D views (no copies)."""
    base = getattr(layer, "base_layer", layer)
    w = base.weight
    w2 = w.reshape(w.shape[0], -1)
    la = lb = None
    scale = 0.0
    if hasattr(layer, "lora_A") and len(layer.lora_A) > 0 and not getattr(layer, "disable_adapters", False):
        if adapter is None:
            act = getattr(layer, "active_adapter", None)
            act = act[0] if isinstance(act, (list, tuple)) and act else act
            adapter = act if isinstance(act, str) and act in layer.lora_A else next(iter(layer.lora_A.keys()))
        a_w = layer.lora_A[adapter].weight
        b_w = layer.lora_B[adapter].weight
        la, lb = a_w.reshape(a_w.shape[0], -1), b_w.reshape(b_w.shape[0], -1)
        scale = float(layer.scaling[adapter])
    return LinearParams(w2, base.bias, la, lb, scale)
