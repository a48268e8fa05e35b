"""Types that are part of the reference's surface: `LlavaConfig(PhiConfig)` with
`model_type = "llava_phi"` (language_model/llava_phi.py:34-35, registered :2001-2002) and `CausalOutputWithMask`
(:37-49).  With transformers installed they subclass the real Hugging Face classes (so `AutoConfig` / checkpoints of
the reference resolve); without it they are plain stand-ins with the same fields - the hot path needs neither."""
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

try:   # pragma: no cover - depends on the environment
    from transformers import AutoConfig, PhiConfig
    from transformers.modeling_outputs import CausalLMOutputWithPast
    _HAVE_HF = True
except Exception:   # transformers absent or broken: stand-ins
    _HAVE_HF = False

    class PhiConfig:   # the defaults of transformers' PhiConfig == Phi-1.5
        model_type = "phi"

        def __init__(self, **kw):
            self.hidden_size, self.num_hidden_layers, self.num_attention_heads = 2048, 24, 32
            self.intermediate_size, self.vocab_size, self.layer_norm_eps = 8192, 51200, 1e-5
            self.rope_theta, self.partial_rotary_factor = 10000.0, 0.5
            for k, v in kw.items():
                setattr(self, k, v)

    @dataclass
    class CausalLMOutputWithPast:
        loss: Optional[torch.FloatTensor] = None
        logits: torch.FloatTensor = None
        past_key_values: Optional[Tuple] = None
        hidden_states: Optional[Tuple] = None
        attentions: Optional[Tuple] = None


class LlavaConfig(PhiConfig):
    model_type = "llava_phi"


@dataclass
class CausalOutputWithMask(CausalLMOutputWithPast):
    loss: Optional[torch.FloatTensor] = None
    logits: torch.FloatTensor = None
    past_key_values: Optional[Tuple[Tuple[torch.FloatTensor]]] = None
    hidden_states: Optional[Tuple[torch.FloatTensor]] = None
    attentions: Optional[Tuple[torch.FloatTensor]] = None
    loss_mask: Optional[torch.FloatTensor] = None
    loss_dice: Optional[torch.FloatTensor] = None
    loss_SEG_class: Optional[torch.FloatTensor] = None
    loss_class_name_class: Optional[torch.FloatTensor] = None
    loss_region_class: Optional[torch.FloatTensor] = None
    loss_llm: Optional[torch.FloatTensor] = None


def register():
    """AutoConfig.register("llava_phi", LlavaConfig) like llava_phi.py:2001 (idempotent; no-op without transformers)."""
    if not _HAVE_HF:
        return False
    try:
        AutoConfig.register("llava_phi", LlavaConfig)
    except ValueError:
        pass   # already registered (e.g. by the reference package in the same process)
    return True
