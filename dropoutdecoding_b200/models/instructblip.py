"""InstructBLIP wrapper with the reference's API (reference: models/instructblip.py).

Flavour (all from the reference):
  * 32 Q-Former tokens projected to the LM are the "visual tokens"; they sit first in the
    sequence: start_img_pos = 0, end_img_pos = 31 (models/instructblip.py:599-602, 648-651);
  * top-10 ids (:187); DETERMINISTIC masks: drop the tokens whose epistemic uncertainty is at or above
    the (1 - prob) quantile, then restore the relevant ones (:447-460); restored between members (:121);
  * members run the decoder without the head, the vote is on argmax of the final HIDDEN state and
    lm_head is applied to the winner only (:125-140);
  * generate() prepends BOS to the returned ids (:680-697);
  * module seed 5217 (:17) -- unused by the deterministic masks.
The reference hands the image span to its patched LM through module globals (:23-27); here it is
ordinary state on the wrapper.  The HF-4.44 aliasing quirk (the in-place edited attention mask
leaking into the next step's preliminary forward, SURVEY 8a note) is NOT reproduced: every step
starts from an all-ones mask, as the LLaVA wrappers do.
"""
from __future__ import annotations

from typing import Optional

import torch
from transformers import InstructBlipForConditionalGeneration, LlamaForCausalLM

from .. import ops
from ._mixin import DropoutDecodeCache, DropoutDecodingMixin  # noqa: F401
from .config import settings  # noqa: F401
from .llava import select_by_average  # noqa: F401   (models/instructblip.py:15)
from .llavanext import select_by_vote  # noqa: F401  (models/instructblip.py:14)

seed = 5217  # models/instructblip.py:17


class CustomLlamaForCausalLM(LlamaForCausalLM):
    """The reference patches the ensemble into this inner LM (models/instructblip.py:28-165).  Here
    the ensemble lives in the engine driven by the outer wrapper; the class is kept so that
    `model.language_model` has the reference's type and state attributes."""

    def __init__(self, config):
        super().__init__(config)
        self.start_image_pos, self.end_image_pos = [], []
        self.start_generation_pos = -1
        self.token_entropies, self.token_ventropies, self.logits_mask_prob = [], [], []
        self.image_features = None
        self.first_generation = True
        self.vision_uncert_dict = None
        self.image_logits = None


class CustomInstructBlipForConditionalGeneration(DropoutDecodingMixin, InstructBlipForConditionalGeneration):
    topk = 10
    mask_mode = "quantile"
    vote_on = "hidden"
    default_seed = seed

    def __init__(self, config):
        super().__init__(config)
        self.language_model = CustomLlamaForCausalLM._from_config(config.text_config)   # :583-585
        self._dd_init_state()

    def _lm(self):
        return self.language_model.model

    def _head(self):
        return self.language_model.lm_head

    @torch.no_grad()
    def generate(self, pixel_values: torch.FloatTensor, qformer_input_ids: Optional[torch.LongTensor] = None,
                 qformer_attention_mask: Optional[torch.LongTensor] = None,
                 input_ids: Optional[torch.LongTensor] = None, attention_mask: Optional[torch.LongTensor] = None,
                 interpolate_pos_encoding: bool = False, max_new_tokens: Optional[int] = None,
                 eos_token_id: Optional[int] = None, pad_token_id: Optional[int] = None,
                 use_cuda_graph: bool = True, **generate_kwargs) -> torch.LongTensor:
        """Reference: models/instructblip.py:588-697.  Returns [B, 1 + T]: BOS followed by the new tokens
        (the LM is driven with inputs_embeds, so the prompt is not echoed)."""
        dev = self._head().weight.device
        dt = self._head().weight.dtype
        pixel_values = pixel_values.to(dev, dt)
        B = pixel_values.shape[0]
        feats = self.get_image_features(pixel_values, qformer_input_ids=qformer_input_ids.to(dev),
                                        qformer_attention_mask=None if qformer_attention_mask is None
                                        else qformer_attention_mask.to(dev),
                                        interpolate_pos_encoding=interpolate_pos_encoding,
                                        return_dict=True).pooler_output               # [B, 32, d]  :608-633
        n_vis = feats.shape[1]
        tc = self.config.text_config
        if input_ids is None:
            input_ids = torch.full((B, 1), tc.bos_token_id, dtype=torch.long)
        input_ids = input_ids.to(dev)
        if attention_mask is not None and not bool(attention_mask.all()):
            raise ValueError("padded prompts are not supported: pass prompts of equal length")
        img_tok = getattr(self.config, "image_token_id", None) or getattr(self.config, "image_token_index", None)
        if img_tok is not None and bool((input_ids == img_tok).any()):
            text_ids = input_ids[:, (input_ids[0] != img_tok)]      # transformers >= 4.46 processors pre-insert 32 placeholders
        else:
            text_ids = input_ids
        emb = self.get_input_embeddings()(text_ids).to(feats.dtype)
        merged = torch.cat([feats, emb], dim=1)                     # :661-664 visual tokens first
        S0 = merged.shape[1]
        self.start_image_pos, self.end_image_pos = [0], [n_vis - 1]  # :599-602, 648-649
        self.start_generation_pos = S0                               # :650-651
        if max_new_tokens is None:
            max_new_tokens = int(generate_kwargs.get("max_length", 20))   # :668-672 (max_length counts new tokens here)
        if eos_token_id is None:
            eos_token_id = getattr(self.language_model.generation_config, "eos_token_id", None) or tc.eos_token_id
        if isinstance(eos_token_id, (list, tuple)):
            eos_token_id = eos_token_id[0]
        if pad_token_id is None:
            pad_token_id = eos_token_id if eos_token_id is not None else 0

        eng = self._engine(B, n_vis, S0 + max_new_tokens, est_seq=S0 + max_new_tokens // 2)
        eng.set_uniforms(None)
        hidden = eng.prefill(merged.to(torch.bfloat16), [0] * B)
        eng.uncertainty_stage(hidden[:, :n_vis])                    # :83-96
        eng.first_token(hidden[:, -1])
        self._publish_uncertainty(eng)
        lm = self.language_model
        lm.vision_uncert_dict, lm.image_features = self.vision_uncert_dict, self.image_features
        lm.start_image_pos, lm.end_image_pos, lm.start_generation_pos = [0], [n_vis - 1], S0

        steps_left = max_new_tokens - 1
        chunk = 16 if eos_token_id is not None else max(steps_left, 1)
        while steps_left > 0:
            n = min(chunk, steps_left)
            eng.decode(n, use_graph=use_cuda_graph)
            steps_left -= n
            if eos_token_id is not None and bool((eng.generated() == eos_token_id).any(-1).all()):
                break
        new = eng.generated().long()
        self.masked_numbers = lm.masked_numbers = eng.n_masked[0].tolist()   # last step's members
        self.logits_mask_prob, self.token_entropies, self.token_ventropies = eng.token_bookkeeping(0)   # :141-157
        lm.logits_mask_prob, lm.token_entropies, lm.token_ventropies = (self.logits_mask_prob, self.token_entropies,
                                                                        self.token_ventropies)
        if eos_token_id is not None:
            done = (new == eos_token_id).cumsum(-1) > 0
            after = torch.cat([torch.zeros_like(done[:, :1]), done[:, :-1]], dim=-1)
            new = new.masked_fill(after, pad_token_id)
            new = new[:, :int((~after).any(0).sum())]
        # BOS first, "to be consistent with other generation models" (:680-695)
        archs = getattr(tc, "architectures", None) or [""]
        bos = 2 if archs[0] == "LLaMAForCausalLM" else tc.bos_token_id
        return torch.cat([torch.full((B, 1), bos, dtype=torch.long, device=dev), new], dim=-1)

    def forward(self, *args, **kwargs):
        # the reference does not touch the outer forward (training path); keep the stock one
        return InstructBlipForConditionalGeneration.forward(self, *args, **kwargs)
