"""Shared host logic of the three Dropout-Decoding wrappers (reference: models/llava.py,
models/llavanext.py, models/instructblip.py): prompt/visual merge, generate(), the reference-style
forward(), the introspection state and the public helper methods.  The per-model classes only say
where the decoder / vision features live and which flavour of the algorithm the reference uses
(top-k size, mask mode, vote target)."""
from __future__ import annotations

import os
import warnings
from typing import Dict, List, Optional, Tuple

import torch
from transformers.modeling_outputs import CausalLMOutputWithPast

from .. import ops
from ..engine import DecoderConfig, DecoderWeights, DropoutDecodeEngine
from .config import settings

def resolve_device_map(device_map) -> Optional[torch.device]:
    """The reference drivers load with `device_map="auto"` (accelerate's layer placement over all visible GPUs,
    chair_test/chair_test.py:189-213).  This engine keeps one whole model per GPU (one process per GPU), so every form
    that names ONE device is honoured, "auto" / "balanced" / "sequential" / "cuda" mean this process's GPU
    (LOCAL_RANK, default 0), and a map that spreads modules over several devices is refused."""
    if device_map is None:
        return None
    if isinstance(device_map, torch.device):
        return device_map
    if isinstance(device_map, int):
        return torch.device(f"cuda:{device_map}")
    if isinstance(device_map, str):
        if device_map in ("auto", "balanced", "balanced_low_0", "sequential", "cuda"):
            return torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}")
        return torch.device(device_map)
    if isinstance(device_map, dict):
        devs = {torch.device(f"cuda:{v}") if isinstance(v, int) else torch.device(v) for v in device_map.values()}
        if len(devs) == 1:
            return devs.pop()
        raise ValueError(f"device_map places the model on {len(devs)} devices {sorted(map(str, devs))}: Dropout Decoding "
                         f"keeps the whole model on one GPU (shard sequences or ensemble members over GPUs with "
                         f"dropoutdecoding_b200.distributed instead)")
    raise ValueError(f"unsupported device_map {device_map!r}")


def decoder_weights_from_hf(lm, lm_head, device) -> DecoderWeights:
    """Fuse q|k|v and gate|up of an HF Llama/Mistral decoder into single weight streams.  The HF
    parameters are re-pointed at views of the fused tensors, so no weight exists twice.  (The caller has made
    the whole model bf16 first -- DropoutDecodingMixin.decoder_weights -- so the aliased parameters keep the
    dtype of their neighbours.)"""
    c = lm.config
    D = getattr(c, "head_dim", None) or c.hidden_size // c.num_attention_heads
    rope_theta = getattr(c, "rope_theta", None)
    if rope_theta is None:
        rp = getattr(c, "rope_parameters", None) or {}
        rope_theta = rp.get("rope_theta", 10000.0)
    cfg = DecoderConfig(c.num_hidden_layers, c.hidden_size, c.num_attention_heads, c.num_key_value_heads, D,
                        c.intermediate_size, lm_head.weight.shape[0], c.rms_norm_eps, float(rope_theta))
    cv = lambda t: t.detach().to(device=device, dtype=torch.bfloat16).contiguous()
    layers = []
    for layer in lm.layers:
        a, m = layer.self_attn, layer.mlp
        wqkv = cv(torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0))
        wgu = cv(torch.cat([m.gate_proj.weight, m.up_proj.weight], 0))
        nq, nk = a.q_proj.weight.shape[0], a.k_proj.weight.shape[0]
        a.q_proj.weight.data, a.k_proj.weight.data, a.v_proj.weight.data = wqkv[:nq], wqkv[nq:nq + nk], wqkv[nq + nk:]
        nf = m.gate_proj.weight.shape[0]
        m.gate_proj.weight.data, m.up_proj.weight.data = wgu[:nf], wgu[nf:]
        layers.append(dict(ln1=cv(layer.input_layernorm.weight), ln2=cv(layer.post_attention_layernorm.weight),
                           wqkv=wqkv, wo=cv(a.o_proj.weight), wgu=wgu, wd=cv(m.down_proj.weight)))
    return DecoderWeights(cfg, cv(lm.embed_tokens.weight), layers, cv(lm.norm.weight), cv(lm_head.weight))


class DropoutDecodeCache:
    """What forward() returns as `past_key_values`: a handle on the engine that owns the (single,
    shared) KV cache.  The reference hands back the winner's deep-copied DynamicCache
    (models/llava.py:373); here the winner's rows were committed into the shared cache instead."""

    def __init__(self, engine: DropoutDecodeEngine):
        self.engine = engine

    def get_seq_length(self, layer_idx: int = 0) -> int:
        return int(self.engine.seq_lens.max())


class DropoutDecodingMixin:
    """Mix in BEFORE the transformers class: `class X(DropoutDecodingMixin, HFClass)`."""
    topk = 5                    # models/llava.py:408
    mask_mode = "cumulative"    # models/llava.py:342-346 (mask not restored between members)
    vote_on = "logits"          # InstructBLIP: "hidden" (models/instructblip.py:125-140)
    default_seed = 24           # models/llava.py:16
    output_cls = CausalLMOutputWithPast

    def _dd_init_state(self):
        # introspection state, same names as the reference (models/llava.py:57-72)
        self.start_generation_pos = 0
        self.start_image_pos: List[int] = []
        self.end_image_pos: List[int] = []
        self.is_first_generation = False
        self.image_features = None          # (values, ids) of the per-visual-token top-k
        self.logits_mask_prob: List[float] = []
        self.all_outputs: list = []
        self.int_array: list = []
        self.token_entropies: List[float] = []
        self.token_ventropies: List[float] = []
        self.vision_uncert_dict: Optional[Dict[str, torch.Tensor]] = None
        self.ag_mask_ids = None
        self.image_logits = None
        self.masked_numbers: List[int] = []
        self.mask_seed = self.default_seed
        self._weights: Optional[DecoderWeights] = None
        self._engines: Dict[Tuple, DropoutDecodeEngine] = {}

    # ------------------------------------------------------------------------------------------
    # ---- hooks the per-model classes override --------------------------------------------------
    def _lm(self):
        """The decoder stack (embed_tokens, layers, norm)."""
        return self.model.language_model

    def _head(self):
        return self.lm_head

    def _image_token_id(self):
        return self.config.image_token_id if hasattr(self.config, "image_token_id") else self.config.image_token_index

    def _visual_features(self, pixel_values: torch.Tensor, **kw) -> torch.Tensor:
        """[B, n_vis, d] projected visual tokens.  The vision tower + projector are library modules and
        launch-bound when run eagerly (~6 ms for one 336 px image on B200 against < 1 ms of math), so their forward is
        captured once per input shape and replayed from a CUDA graph (DD_VISION_GRAPH=0: eager)."""
        def eager(px):
            feats = self.model.get_image_features(pixel_values=px, return_dict=True, **kw).pooler_output
            return torch.stack(list(feats)) if not torch.is_tensor(feats) else feats

        if not pixel_values.is_cuda or os.environ.get("DD_VISION_GRAPH", "1") != "1" or torch.is_grad_enabled():
            return eager(pixel_values)
        graphs = self.__dict__.setdefault("_vision_graphs", {})
        key = (tuple(pixel_values.shape), pixel_values.dtype, pixel_values.device.index,
               tuple(sorted((k, str(v)) for k, v in kw.items())))
        entry = graphs.get(key)
        if entry is None:
            try:
                static_in = pixel_values.clone()
                side = torch.cuda.Stream(device=pixel_values.device)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):                      # warm-up off the capture (lazy inits, autotuning)
                    for _ in range(2):
                        eager(static_in)
                torch.cuda.current_stream().wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    static_out = eager(static_in)
                entry = (static_in, graph, static_out)
            except Exception as e:                                  # a tower that cannot be captured stays eager
                warnings.warn(f"vision tower not captured in a CUDA graph ({type(e).__name__}: {e}); running it eagerly")
                torch.cuda.synchronize()
                entry = False
            graphs[key] = entry
        if entry is False:
            return eager(pixel_values)
        static_in, graph, static_out = entry
        static_in.copy_(pixel_values)
        graph.replay()
        return static_out

    def _keep_relevant(self) -> bool:
        return True

    # -------------------------------------------------------------------------------------------
    def decoder_weights(self) -> DecoderWeights:
        if self._weights is None:
            dev = self._head().weight.device
            if dev.type != "cuda":
                raise ops.DropDecError("Dropout Decoding runs on a CUDA device only (no CPU path): move the "
                                       "model to cuda first")
            if self._head().weight.dtype != torch.bfloat16:
                # chair_test / pope_test load with torch_dtype=torch.float16 (chair_test/chair_test.py:189-213).  The
                # engine computes in bf16 (fp32 accumulation); the WHOLE model is converted once, so that the HF
                # modules the wrapper still calls (vision tower, projector, text-only fallbacks) and the fused decoder
                # weights agree on one dtype.  fp16 -> bf16 drops 3 mantissa bits of every weight.
                warnings.warn(f"Dropout Decoding runs in bfloat16: converting the model from "
                              f"{self._head().weight.dtype} to torch.bfloat16 (weights are rounded once)", stacklevel=3)
                self.to(torch.bfloat16)
            self._weights = decoder_weights_from_hf(self._lm(), self._head(), dev)
        return self._weights

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, *model_args, device_map=None, **kwargs):
        """`Custom*.from_pretrained(path, torch_dtype=torch.float16, device_map="auto")` of the reference drivers
        (chair_test/chair_test.py:185-214, pope_test/pope_test.py:180-197): torch_dtype is passed through (fp16 / fp32
        checkpoints are converted to bf16 at the first Dropout-Decoding call, with a warning), device_map is resolved
        to ONE device (resolve_device_map) and the model is moved there."""
        dev = resolve_device_map(device_map)
        model = super().from_pretrained(pretrained_model_name_or_path, *model_args, **kwargs)
        if dev is not None:
            model = model.to(dev)
        return model

    def _engine(self, batch: int, n_vis: int, max_seq: int, est_seq: int) -> DropoutDecodeEngine:
        probs = tuple(float(p) for p in settings['voting_numbers'])
        agg = "mean" if settings.get('use_avg') else "vote"
        key = (batch, n_vis, max_seq, probs, agg, self.mask_mode, self._keep_relevant())
        eng = self._engines.get(key)
        if eng is None:
            self._engines.clear()            # one live KV cache at a time
            eng = DropoutDecodeEngine(self.decoder_weights(), batch, n_vis, max_seq, probs=probs,
                                      mode=self.mask_mode, keep_relevant=self._keep_relevant(), topk=self.topk,
                                      aggregate=agg, seed=self.mask_seed, max_steps=max_seq, est_seq=est_seq,
                                      vote_on=self.vote_on)
            self._engines[key] = eng
        return eng

    # ------------------------------------------------------------------------------------------
    def _merge(self, input_ids: torch.Tensor, pixel_values: torch.Tensor, **vision_kwargs):
        """[text | visual | text] embeddings; records start/end_image_pos like
        _merge_input_ids_with_image_features (models/llava.py:74-78)."""
        img_tok = self._image_token_id()
        feats = self._visual_features(pixel_values, **vision_kwargs)                    # [B, n_vis, d]
        B, n_vis, d = feats.shape
        is_img = input_ids == img_tok
        counts = is_img.sum(-1)
        if not bool((counts == counts[0]).all()) or int(counts[0]) not in (1, n_vis):
            raise ValueError(
                f"The input provided to the model are wrong. The number of image tokens is {int(is_img.sum())} while"
                f" the number of image given to the model is {B}. This prevents correct indexing and breaks batch "
                f"generation.")                                  # models/llava.py:134-138
        start = is_img.float().argmax(-1)
        if not bool((start == start[0]).all()):
            raise ValueError("batched prompts must place the image at the same position")
        start = int(start[0])
        emb = self.get_input_embeddings()(input_ids.masked_fill(is_img, 0)).to(feats.dtype)
        tail = start + int(counts[0])
        merged = torch.cat([emb[:, :start], feats, emb[:, tail:]], dim=1)
        self.start_image_pos = [start]
        self.end_image_pos = [start + n_vis - 1]
        return merged, start, n_vis

    @torch.no_grad()
    def generate(self, input_ids: torch.LongTensor = None, pixel_values: torch.FloatTensor = None,
                 attention_mask: Optional[torch.Tensor] = None, max_new_tokens: int = 20, num_beams: int = 1,
                 pad_token_id: Optional[int] = None, eos_token_id: Optional[int] = None,
                 uniforms: Optional[torch.Tensor] = None, use_cuda_graph: bool = True, **kwargs) -> torch.LongTensor:
        """Greedy Dropout Decoding (reference call: chair_test/chair_test.py:341-346).  Returns
        [B, S_in + T] token ids, prompt included.  Without pixel_values this is the parent's generate."""
        if pixel_values is None:
            return super().generate(input_ids=input_ids, attention_mask=attention_mask, max_new_tokens=max_new_tokens,
                                    num_beams=num_beams, pad_token_id=pad_token_id, eos_token_id=eos_token_id, **kwargs)
        if num_beams != 1:
            raise ValueError("Dropout Decoding is greedy: num_beams must be 1")
        if attention_mask is not None and not bool(attention_mask.all()):
            raise ValueError("padded prompts are not supported: pass prompts of equal length")
        dev = self._head().weight.device
        input_ids = input_ids.to(dev, non_blocking=True)
        pixel_values = pixel_values.to(dev, self._head().weight.dtype, non_blocking=True)
        if eos_token_id is None:
            eos_token_id = getattr(self.generation_config, "eos_token_id", None)
        if isinstance(eos_token_id, (list, tuple)):
            eos_token_id = eos_token_id[0] if eos_token_id else None
        if pad_token_id is None:
            pad_token_id = eos_token_id if eos_token_id is not None else 0
        vis_kw = {k: kwargs[k] for k in ("vision_feature_layer", "vision_feature_select_strategy", "image_sizes")
                  if kwargs.get(k) is not None}

        merged, start, n_vis = self._merge(input_ids, pixel_values, **vis_kw)
        B, S0, _ = merged.shape
        self.is_first_generation = True
        self.start_generation_pos = S0                                   # models/llava.py:285-286
        eng = self._engine(B, n_vis, S0 + max_new_tokens, est_seq=S0 + max_new_tokens // 2)
        eng.set_uniforms(uniforms)
        hidden = eng.prefill(merged.to(torch.bfloat16), [start] * B)
        eng.uncertainty_stage(hidden[:, start:start + n_vis])             # models/llava.py:306-314
        eng.first_token(hidden[:, -1])
        self._publish_uncertainty(eng)

        steps_left = max_new_tokens - 1
        # EOS is checked between decode calls; the asynchronous batch schedule re-synchronises the sequences at the
        # end of every call, so it gets longer calls
        # (one sequence on the speculative path: the host sees every pass's input token anyway -- no chunking)
        host_stop = eng.spec_active and eos_token_id is not None
        chunk = (64 if eng.spec_batch else 16) if (eos_token_id is not None and not host_stop) else steps_left
        while steps_left > 0:
            n = min(chunk, steps_left)
            done = eng.decode(n, use_graph=use_cuda_graph, stop_token=eos_token_id if host_stop else None)
            steps_left -= n
            if done < n or (eos_token_id is not None and bool((eng.generated() == eos_token_id).any(-1).all())):
                break
        self.is_first_generation = False
        new = eng.generated().long()
        self.masked_numbers = eng.n_masked[0].tolist()
        self.logits_mask_prob, self.token_entropies, self.token_ventropies = eng.token_bookkeeping(0)
        if eos_token_id is not None:                                    # finished rows emit pad (HF greedy)
            done = (new == eos_token_id).cumsum(-1) > 0
            after = torch.cat([torch.zeros_like(done[:, :1]), done[:, :-1]], dim=-1)
            new = new.masked_fill(after, pad_token_id)
            keep = int((~after).any(0).sum())
            new = new[:, :keep]
        return torch.cat([input_ids, new], dim=-1)

    def _publish_uncertainty(self, eng: DropoutDecodeEngine) -> None:
        """Expose the stage-1 results under the reference's names (models/llava.py:747-756, 408-410)."""
        self.vision_uncert_dict = {
            "variance_per_token": eng.var.clone(), "epis_uncert_per_token": eng.epi.clone(),
            "alea_uncert_per_token": eng.alea.clone(), "variance": eng.var.mean(-1),
            "epis_uncert": eng.epi.mean(-1), "alea_uncert": eng.alea.mean(-1)}
        self.image_features = (None, eng.topk_ids.long())

    # ------------------------------------------------------------------------------------------
    # forward(): the reference's per-token entry point (models/llava.py:155-388), for callers that
    # drive the greedy loop themselves.  Prefill iff input_ids.shape[-1] != 1 (:218-226).
    # ------------------------------------------------------------------------------------------
    max_cache_tokens = 1024     # decode budget of a cache created by forward() (generate sizes its own)

    @torch.no_grad()
    def forward(self, input_ids: torch.LongTensor = None, pixel_values: torch.FloatTensor = None,
                attention_mask: Optional[torch.Tensor] = None, position_ids: Optional[torch.LongTensor] = None,
                past_key_values=None, inputs_embeds: Optional[torch.FloatTensor] = None,
                vision_feature_layer=None, vision_feature_select_strategy: Optional[str] = None,
                labels: Optional[torch.LongTensor] = None, use_cache: Optional[bool] = None,
                output_attentions: Optional[bool] = None, output_hidden_states: Optional[bool] = None,
                return_dict: Optional[bool] = None, cache_position: Optional[torch.LongTensor] = None,
                logits_to_keep=0, **kwargs):
        is_dd_cache = isinstance(past_key_values, DropoutDecodeCache)
        if (pixel_values is None and not is_dd_cache) or labels is not None or inputs_embeds is not None:
            # not a Dropout-Decoding call (text-only / training): the stock forward, with everything transformers'
            # generate() hands over (it picks cache_position / logits_to_keep by looking at THIS signature)
            if cache_position is not None:
                kwargs["cache_position"] = cache_position
            return super().forward(input_ids=input_ids, pixel_values=pixel_values, attention_mask=attention_mask,
                                   position_ids=position_ids, past_key_values=past_key_values,
                                   inputs_embeds=inputs_embeds, labels=labels, use_cache=use_cache,
                                   logits_to_keep=logits_to_keep, **kwargs)
        dev = self._head().weight.device
        input_ids = input_ids.to(dev)
        if input_ids.shape[-1] != 1:                                       # ---- prefill branch
            self.is_first_generation = True                                # :218-224
            self.logits_mask_prob, self.token_entropies, self.token_ventropies = [], [], []
            vis_kw = {k: v for k, v in (("vision_feature_layer", vision_feature_layer),
                                        ("vision_feature_select_strategy", vision_feature_select_strategy),
                                        ("image_sizes", kwargs.get("image_sizes"))) if v is not None}
            merged, start, n_vis = self._merge(input_ids, pixel_values.to(dev, self._head().weight.dtype), **vis_kw)
            B, S0, _ = merged.shape
            self.start_generation_pos = S0                                 # :285-286
            eng = self._engine(B, n_vis, S0 + self.max_cache_tokens, est_seq=S0 + 128)
            eng.set_uniforms(None)
            hidden = eng.prefill(merged.to(torch.bfloat16), [start] * B)
            eng.uncertainty_stage(hidden[:, start:start + n_vis])          # :306-314
            self._publish_uncertainty(eng)
            # full-sequence logits as the reference returns them (:382-388); library GEMM, prefill side
            logits = torch.nn.functional.linear(hidden, eng.w.lm_head).float()
            self.image_logits = logits[:, start:start + n_vis]
            return self.output_cls(loss=None, logits=logits, past_key_values=DropoutDecodeCache(eng))
        # ---- decode branch: one ensemble step for the token the caller picked (:336-376)
        if not is_dd_cache:
            raise ValueError("decode step needs the past_key_values returned by the prefill forward")
        self.is_first_generation = False
        eng = past_key_values.engine
        eng.cur_token.copy_(input_ids[:, 0].to(torch.int32))
        eng.decode(1, use_graph=True)
        self.masked_numbers = eng.n_masked[0].tolist()                     # :338, 661-662
        self.logits_mask_prob, self.token_entropies, self.token_ventropies = eng.token_bookkeeping(0)
        return self.output_cls(loss=None, logits=eng.step_logits()[:, None, :],
                                           past_key_values=past_key_values)

    # ---- public helpers kept from the reference -----------------------------------------------
    def calculate_vision_uncertainty(self, logits: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Reference signature (models/llava.py:710-756) takes materialised logits [B, L, V]; the
        hot path never builds them (see project_vision_uncertainty).  This helper exists for API
        compatibility and evaluates the same formulas with torch ops on the logits' device."""
        probs = torch.softmax(logits.float(), dim=-1)
        p_avg = probs.mean(dim=1, keepdim=True)
        epi = (probs * (torch.log(probs + 1e-10) - torch.log(p_avg + 1e-10))).sum(-1)
        alea = -(probs * torch.log(probs + 1e-10)).sum(-1)
        var = probs.var(dim=-1)
        return {"variance_per_token": var, "epis_uncert_per_token": epi, "alea_uncert_per_token": alea,
                "variance": var.mean(-1), "epis_uncert": epi.mean(-1), "alea_uncert": alea.mean(-1)}

    def project_vision_uncertainty(self, vis_hidden: torch.Tensor, round_logits: bool = True):
        """Fused lm_head projection + uncertainty of final-norm visual hidden states [B, n_vis, d]
        (dd_uncertainty_project): the logits are never materialised."""
        w = self.decoder_weights()
        epi, alea, var, ids = ops.uncertainty_project(vis_hidden.to(torch.bfloat16).contiguous(), w.lm_head,
                                                      topk=self.topk, round_logits=round_logits)
        return {"variance_per_token": var, "epis_uncert_per_token": epi, "alea_uncert_per_token": alea,
                "variance": var.mean(-1), "epis_uncert": epi.mean(-1), "alea_uncert": alea.mean(-1)}, ids

    def get_image_logits(self, logits, start_image_pos, end_image_pos):      # models/llava.py:412-426
        return logits[:, start_image_pos:end_image_pos + 1]

    def get_topk_token_id(self, image_logits, topk=5):                      # models/llava.py:428-441
        return torch.topk(image_logits, topk, dim=-1)

    def get_overlap_image_tokens(self, logits, id=-1):                       # models/llava.py:443-482
        max_id = torch.argmax(logits, dim=-1)[0] if id == -1 else torch.tensor([id], device=logits.device)
        ids = self.image_features[1][0]
        hit = (ids == max_id.reshape(-1)[-1]).any(dim=1)
        return torch.nonzero(hit).squeeze() + self.start_image_pos[0]

    def lowest_percent_kl_indices(self, image_logits, logits, percent=0.1):
        """models/llava.py:758-776 / models/instructblip.py:559-578: indices (inside the visual span) of the
        int(percent * N) visual tokens whose projected distribution is closest in KL to the current one.  Needs the
        materialised image logits [1, N, V] (`self.image_logits`, which only the reference-style forward() prefill
        builds -- the generate() hot path never does); an ablation helper, evaluated with torch ops."""
        il = image_logits.squeeze(0).float()
        cur = logits.reshape(1, -1).float()
        kl = torch.nn.functional.kl_div(torch.log_softmax(il, dim=-1), torch.softmax(cur, dim=-1).expand_as(il),
                                        reduction="none").sum(dim=-1)
        return torch.topk(kl, int(percent * kl.numel()), largest=False).indices

    def get_image_attention_mask(self, logits, attention_mask, method="epis", prob=0.5, ids=[]):
        """'epis' / 'epis_no_overlap' / 'epis_kl' of the reference (models/llava.py:589-683, instructblip.py:464-485)
        for ONE member, on the device sampler: zeros `attention_mask` in place at the drawn visual positions (cumulative
        with whatever is already zero) and restores the relevant tokens ('epis': those whose top-k ids contain the
        preliminary argmax; 'epis_kl': the lowest-KL tenth, from the materialised `self.image_logits`)."""
        if method not in ("epis", "epis_no_overlap", "epis_kl"):
            raise NotImplementedError(f"mask method {method!r} is dead code in the reference and not provided")
        if method == "epis_kl" and self.image_logits is None:
            raise ValueError("'epis_kl' needs the materialised image logits: run the reference-style forward() prefill "
                             "(generate() never builds them)")
        epi = self.vision_uncert_dict["epis_uncert_per_token"][:1].float().contiguous()
        n_vis, start = epi.shape[1], self.start_image_pos[0]
        y = torch.argmax(logits.reshape(-1, logits.shape[-1])[-1:], dim=-1).int()
        keep = method == "epis"
        self._mask_calls = getattr(self, "_mask_calls", 0)
        bm, cnt = ops.sample_masks(epi, self.image_features[1][:1].int().contiguous(), y, [prob], mode="independent",
                                   keep_relevant=False, seed=self.mask_seed, offset=4 * self._mask_calls)
        self._mask_calls += 1
        words = bm[0, 0].cpu().numpy().astype("uint32")
        import numpy as np
        bits = ((words[:, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool).reshape(-1)[:n_vis]
        drop = torch.from_numpy(bits).to(attention_mask.device)
        attention_mask[:, start:start + n_vis][:, drop] = 0
        if keep:
            rel = self.get_overlap_image_tokens(logits).reshape(-1)
            attention_mask[:, rel] = 1
            self.masked_numbers.append(int((attention_mask[0] == 0).sum()))
        elif method == "epis_kl":
            low = self.lowest_percent_kl_indices(self.image_logits[:1], logits.reshape(1, 1, -1)[:, -1:])
            attention_mask[:, (low + start).to(attention_mask.device)] = 1
        return attention_mask
