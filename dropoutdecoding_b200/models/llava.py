"""LLaVA-1.5 wrapper with the reference's API (reference: models/llava.py), running the
Dropout-Decoding hot path on the B200 engine.

Kept from the reference: class name, `from_pretrained` / `(config)` construction, `generate(...)`
and `forward(...)` signatures, the `settings` dict, the introspection attributes
(models/llava.py:55-72) and the public helper methods.  Replaced: everything inside -- the
per-token ensemble runs as two batched passes over one shared KV cache through libdropdec.so
(see engine.py) instead of K+1 sequential HF forwards on deep-copied caches.

Differences a user can observe, all deliberate:
  * batches of B > 1 prompts of equal length are decoded together (the reference is B = 1 only);
  * the mask draws come from an explicit Philox stream (seed 24 = models/llava.py:16) instead of
    the process-global torch generator; `uniforms=` reproduces the reference draws exactly;
  * prompts may carry one `<image>` id (transformers 4.44 processors) or pre-expanded
    placeholders (transformers >= 4.45).
"""
from __future__ import annotations

from collections import Counter
from typing import Dict, List, Optional, Tuple

import torch
from transformers import LlavaForConditionalGeneration
from transformers.models.llava.modeling_llava import LlavaCausalLMOutputWithPast

from .. import ops
from ._mixin import DropoutDecodeCache, DropoutDecodingMixin, decoder_weights_from_hf  # noqa: F401
from .config import settings  # noqa: F401

seed = 24  # models/llava.py:16


def select_by_vote(outputs_all):
    """Reference signature (models/llava.py:22-36): outputs_all is a list of objects whose [0] is the
    logits [1, T, V]; returns (winning output, its index).  Runs dd_aggregate on the device."""
    last = torch.stack([o[0][-1].reshape(-1) for o in outputs_all])[None].float().contiguous()
    _, winner, _, _ = ops.aggregate(last, "vote")
    idx = int(winner[0])
    return outputs_all[idx], idx


def select_by_average(outputs_all):
    """Reference signature (models/llava.py:37-52): member 0's output with its last-position logits
    replaced by the fp32 mean over members."""
    last = torch.stack([o[0][-1].reshape(-1) for o in outputs_all])[None].float().contiguous()
    _, _, _, mean = ops.aggregate(last, "mean")
    out = outputs_all[0]
    out[0][-1] = mean[0].reshape(out[0][-1].shape).to(out[0].dtype)
    return out


class CustomLlavaForConditionalGeneration(DropoutDecodingMixin, LlavaForConditionalGeneration):
    """LLaVA-1.5 flavour: top-5 ids (models/llava.py:408), cumulative masks (:342-346), vote on logits,
    mask seed 24 (:16)."""
    topk = 5
    mask_mode = "cumulative"
    vote_on = "logits"
    default_seed = seed
    output_cls = LlavaCausalLMOutputWithPast

    def __init__(self, config):
        super().__init__(config)
        self._dd_init_state()
