"""disentangled-subject-to-vid_b200 — B200-native (sm_100a) implementation of the CogVideoX subject-to-video
denoising hot path of carpedkm/disentangled-subject-to-vid, behind the reference's own pipeline / module surface.

The directory name is not a Python identifier; import it as `s2v_b200` (the alias module at the repo root) or with
importlib.import_module("disentangled-subject-to-vid_b200").

    csrc/ + include/s2v_b200.h   hand-written sm_100a kernels (tcgen05 / TMEM / TMA) behind a C ABI
    ops                          ctypes wrappers: torch tensors -> raw pointers
    engine                       weight packing, workspaces, the fused block / model forward
    modules, pipeline, scheduler the reference's operator surface (same names, arguments, errors)
    lora                         PEFT-layout LoRA adapters, read in place
    vae                          3D causal VAE decoder (implicit-GEMM convs) behind AutoencoderKLCogVideoX.decode
    parallel                     prompt / CFG sharding over the GPUs of one node (torch.distributed)
    t5                           T5 v1.1 prompt encoder behind an attached transformers.T5EncoderModel
"""
from . import _lib, engine, lora, modules, ops, parallel, pipeline, scheduler, t5, tables, vae  # noqa: F401
from .lora import inject_lora, load_lora_state_dict  # noqa: F401
from .modules import (  # noqa: F401
    Attention,
    CogVideoXAttnProcessor2_0,
    CogVideoXBlock,
    CogVideoXTransformer3DModel,
    attach,
)
from .pipeline import CogVideoXPipelineOutput, CustomCogVideoXPipeline, export_to_video, postprocess_video  # noqa: F401
from .scheduler import CogVideoXDDIMScheduler, CogVideoXDPMScheduler  # noqa: F401
from .t5 import attach_t5  # noqa: F401
from .vae import AutoencoderKLCogVideoX, attach_vae, encode_reference_image  # noqa: F401

__version__ = "0.1.0"
