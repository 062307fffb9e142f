"""Minimal stand-in for the PEFT LoRA layer layout (peft is NOT installed here and is un-vendored in the reference).

Reproduces what `transformer.add_adapter(LoraConfig(r, lora_alpha, target_modules=[...]))` leaves behind
(S/inference.py:218-225 -> D/loaders/peft.py:112-148 -> peft.inject_adapter_in_model): the target nn.Linear /
nn.Conv2d is replaced by a wrapper exposing `base_layer`, `lora_A["default"]`, `lora_B["default"]`,
`scaling["default"]`, whose forward is PEFT's published `base(x) + B(A(x)) * scaling` (dropout 0, no LoRA bias).
Target matching is PEFT's suffix rule: a module matches if its dotted name equals a target or ends with "."+target.
"""
import torch
import torch.nn as nn

TARGETS = ["to_k", "to_q", "to_v", "to_out.0", "proj", "text_proj", "norm1.linear", "norm2.linear", "ff.net.2"]


class LoraLayer(nn.Module):
    def __init__(self, base: nn.Module, r: int, alpha: float):
        super().__init__()
        self.base_layer = base
        if isinstance(base, nn.Linear):
            a = nn.Linear(base.in_features, r, bias=False)
            b = nn.Linear(r, base.out_features, bias=False)
        elif isinstance(base, nn.Conv2d):
            a = nn.Conv2d(base.in_channels, r, base.kernel_size, base.stride, base.padding, bias=False)
            b = nn.Conv2d(r, base.out_channels, (1, 1), (1, 1), bias=False)
        else:
            raise TypeError(type(base))
        dt = base.weight.dtype
        self.lora_A = nn.ModuleDict({"default": a.to(dt)})
        self.lora_B = nn.ModuleDict({"default": b.to(dt)})
        self.scaling = {"default": alpha / r}
        self.r = {"default": r}

    @property
    def weight(self):
        return self.base_layer.weight

    @property
    def bias(self):
        return self.base_layer.bias

    def forward(self, x, *args, **kwargs):
        y = self.base_layer(x)
        return y + self.lora_B["default"](self.lora_A["default"](x)) * self.scaling["default"]


def matches(name: str, targets=TARGETS) -> bool:
    return any(name == t or name.endswith("." + t) for t in targets)


def inject(model: nn.Module, r: int, alpha: float, targets=TARGETS):
    names = [n for n, m in model.named_modules() if matches(n, targets) and isinstance(m, (nn.Linear, nn.Conv2d))]
    for n in names:
        parent_name, _, child = n.rpartition(".")
        parent = model.get_submodule(parent_name) if parent_name else model
        base = getattr(parent, child) if not child.isdigit() else parent[int(child)]
        wrapped = LoraLayer(base, r, alpha)
        if child.isdigit():
            parent[int(child)] = wrapped
        else:
            setattr(parent, child, wrapped)
    return names


def load_flat_params(model: nn.Module, params: dict):
    """Load a flat oracle-style dict ("<module>.weight", "<module>.lora_A.weight") into a (possibly wrapped) model."""
    sd = {}
    wrapped = {n for n, m in model.named_modules() if isinstance(m, LoraLayer)}
    for k, v in params.items():
        mod, _, leaf = k.rpartition(".")
        if mod.endswith(".lora_A") or mod.endswith(".lora_B"):
            base, _, ab = mod.rpartition(".")
            sd[f"{base}.{ab}.default.{leaf}"] = v
        elif mod in wrapped:
            sd[f"{mod}.base_layer.{leaf}"] = v
        else:
            sd[k] = v
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [m for m in missing if "pos_embedding" not in m]
    assert not missing and not unexpected, (missing, unexpected)
