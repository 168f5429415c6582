"""LLaVA-NeXT wrapper with the reference's API (reference: models/llavanext.py).

Differences from the LLaVA-1.5 flavour, all taken from the reference:
  * top-10 ids per visual token (models/llavanext.py:652);
  * the attention mask is RESTORED between members (:546) -> independent masks;
  * settings['use_random'][0] selects "epis_no_overlap" (no relevant-token restore, :547-550);
  * module seed 506 (:18-21); select_by_vote returns the output only (:26-39);
  * anyres: the number of visual tokens depends on the image size (up to ~2.9k).
"""
from __future__ import annotations

import torch
from transformers import LlavaNextForConditionalGeneration
from transformers.models.llava_next.modeling_llava_next import LlavaNextCausalLMOutputWithPast

from .. import ops
from ._mixin import DropoutDecodeCache, DropoutDecodingMixin  # noqa: F401
from .config import settings

seed = 506  # models/llavanext.py:18


def argmax(output):                                           # models/llavanext.py:23-24
    return int(torch.argmax(output).item())


def select_by_vote(outputs_all):
    """Reference signature (models/llavanext.py:26-39): returns the winning output only."""
    last = torch.stack([o[0][-1].reshape(-1) for o in outputs_all])[None].float().contiguous()
    _, winner, _, _ = ops.aggregate(last, "vote")
    return outputs_all[int(winner[0])]


class CustomLlavaNextForConditionalGeneration(DropoutDecodingMixin, LlavaNextForConditionalGeneration):
    topk = 10
    mask_mode = "independent"
    vote_on = "logits"
    default_seed = seed
    output_cls = LlavaNextCausalLMOutputWithPast

    def __init__(self, config):
        super().__init__(config)
        self._dd_init_state()
        self.do_step = True                                   # models/llavanext.py:492 (case-study flag)

    def _keep_relevant(self) -> bool:
        return not bool(settings['use_random'][0])            # models/llavanext.py:547-550

    def _visual_features(self, pixel_values: torch.Tensor, image_sizes=None, **kw) -> torch.Tensor:
        if image_sizes is None:
            raise ValueError("LLaVA-NeXT needs image_sizes (anyres patch packing, models/llavanext.py:389-427)")
        feats = self.model.get_image_features(pixel_values=pixel_values, image_sizes=image_sizes, return_dict=True,
                                              **kw).pooler_output
        feats = list(feats) if not torch.is_tensor(feats) else [feats]
        n = {f.shape[0] for f in feats}
        if len(n) != 1:
            raise ValueError("batched LLaVA-NeXT prompts must yield the same number of visual tokens "
                             f"(got {sorted(n)}); decode such images one at a time")
        return torch.stack(feats)
